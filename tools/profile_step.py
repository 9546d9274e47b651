"""Diagnostic: torch.profiler table of one update (which ATen ops still run next to the rorl kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as Bn
from rorl_b200.algorithm.sac_full_length_rnn_redq_sep_optim import SACFullLengthRNNREDQ_SEP_OPTIM
torch.manual_seed(0); np.random.seed(0)
dev = torch.device("cuda:0")
alg = SACFullLengthRNNREDQ_SEP_OPTIM(dict(Bn.HP), Bn.model_kwargs(Bn.ENCODER, False), Bn.model_kwargs(Bn.ENCODER, True), Bn.T_LEN, device=dev)
alg.replay_buffer._init_memory_buffer(Bn.template_transition())
rng = np.random.RandomState(1000)
for _ in range(Bn.N_TRAJ):
    alg.replay_buffer.push_trajectory_array(Bn.synth_trajectory(rng))
for _ in range(3):
    alg.train_one_batch(sync=False)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    alg.train_one_batch(sync=False)
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=110, max_name_column_width=45, max_shapes_column_width=70))
