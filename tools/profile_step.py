"""Diagnostic: kernel-level table of ONE eager update (torch.profiler device events, aggregated by kernel name), and
the ATen ops behind the non-rorl kernels.  RORL_BENCH_ENCODER / RORL_BENCH_ALGO select the configuration."""
import collections
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench as Bn
from rorl_b200.utility.alg_init import alg_class

torch.manual_seed(0)
np.random.seed(0)
dev = torch.device("cuda:0")
cls = alg_class("sac_rnn_full_horizon_redQ_sep_optim" if Bn.ALGO == "sac" else "td3_rnn_full_horizon_redQ_sep_optim")
alg = cls(dict(Bn.HP, use_cuda_graph=False), Bn.model_kwargs(Bn.ENCODER, False), Bn.model_kwargs(Bn.ENCODER, True), Bn.T_LEN, device=dev)
alg.replay_buffer._init_memory_buffer(Bn.template_transition())
rng = np.random.RandomState(1000)
for _ in range(Bn.N_TRAJ):
    alg.replay_buffer.push_trajectory_array(Bn.synth_trajectory(rng))
for _ in range(3):
    alg.train_one_batch(sync=False)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    alg.train_one_batch(sync=False)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if str(e.device_type).endswith("CUDA"):
        name = re.sub(r"<.*", "", e.name.replace("void ", "")).strip()[:64]
        agg[name][0] += 1
        agg[name][1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v[1] for v in agg.values())
print(f"one eager update: {sum(v[0] for v in agg.values())} device launches, {tot / 1e3:.2f} ms of device time")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{100 * t / tot:6.2f}% {t:9.1f} us {n:4d}  {k}")
print()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=70, max_name_column_width=40, max_shapes_column_width=60))

# who launches the non-rorl kernels: innermost CPU op -> chain of its parents (up to 4), with input shapes
chains = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    ks = [k for k in getattr(e, "kernels", []) if not k.name.startswith("void rorl::") and not k.name.startswith("rorl::")]
    if not ks or str(e.device_type).endswith("CUDA"):
        continue
    if any(getattr(c, "kernels", []) for c in (e.cpu_children or [])):
        continue                                   # not the innermost launcher
    chain, p = [e.name], e.cpu_parent
    while p is not None and len(chain) < 4:
        chain.append(p.name)
        p = p.cpu_parent
    key = " <- ".join(chain) + "  " + str(e.input_shapes)[:70]
    chains[key][0] += len(ks)
    chains[key][1] += sum(k.duration for k in ks)
print("\nnon-rorl launches by launching op:")
for k, (n, t) in sorted(chains.items(), key=lambda kv: -kv[1][1])[:70]:
    print(f"{t:8.1f} us {n:4d}  {k}")

# individual launches in time order with their durations (spot kernels that are slow for their shape)
if os.environ.get("RORL_PROFILE_LAUNCHES"):
    evs = [e for e in prof.events() if str(e.device_type).endswith("CUDA")]
    evs.sort(key=lambda e: e.time_range.start)
    print("\nlaunch sequence (us):")
    for e in evs:
        t = e.device_time if hasattr(e, "device_time") else e.cuda_time
        name = re.sub(r"\(.*", "", e.name.replace("void ", "")).strip()[:70]
        print(f"{t:8.1f}  {name}")
