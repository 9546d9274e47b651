"""Data-parallel equivalence ON GPUs: the same batch updated by one process and by N trajectory-
sharded NCCL ranks must leave the same gradients (<= 1e-5 relative: summation order only), logged scalars, guard state
and updated parameters.  Needs >= 2 GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_dp_gpu.py -m gpu`); skipped on a
single-GPU box."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
S, A, H = 5, 3, 64
LENS = [40] * 8


def _kw(enc, value):
    return dict(state_dim=S, action_dim=A, embedding_size=32, embedding_hidden=[H, H], embedding_activations=['elu', 'elu', 'linear'],
                embedding_layer_type=['fc', enc, 'fc'], uni_model_hidden=[H, H], uni_model_activations=['elu', 'elu', 'linear'],
                uni_model_layer_type=(['efc-8'] * 3 if value else ['fc'] * 3), fix_rnn_length=0, uni_model_input_mapping_dim=32,
                reward_input=False, last_action_input=True, last_state_input=True, separate_encoder=True)


HP = dict(gamma=0.99, sac_tau=0.995, policy_update_per=1, redq_m=2, policy_lr=3e-4, value_lr=1e-3, rnn_policy_lr=1e-5, rnn_value_lr=1e-5,
          alpha_lr=1e-4, target_entropy_ratio=1.0, sac_batch_size=sum(LENS) - 1, max_buffer_transition_num=1000, use_cuda_graph=False)


def _fill(buf, Transition, rng):
    for Tn in LENS:
        last_s, last_a, last_r = np.zeros((1, S)), np.zeros((1, A)), np.zeros((1, 1))
        s = rng.standard_normal((1, S))
        for t in range(Tn):
            a = np.tanh(rng.standard_normal((1, A)))
            ns = rng.standard_normal((1, S))
            r = float(rng.standard_normal())
            done = t == Tn - 1
            buf.mem_push(Transition(state=s, last_state=last_s, last_action=last_a, action=a, next_state=ns, reward=r, logp=None, mask=1,
                                    done=done, timeout=done, start=(t == 0), reward_input=last_r))
            last_s, last_a, last_r, s = s, a, np.array([[r]]), ns


def _run(rank, world, port, enc, algo, out, graph):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from rorl_b200.algorithm.data_parallel import shard_rows
    from rorl_b200.buffers.transition_buffer.replay_memory import Transition
    from rorl_b200.utility.alg_init import alg_class
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    group = None
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        group = dist.group.WORLD
    torch.manual_seed(3)
    np.random.seed(50)
    cls = alg_class("sac_rnn_full_horizon_redQ_sep_optim" if algo == "sac" else "td3_rnn_full_horizon_redQ_sep_optim")
    alg = cls(dict(HP, use_cuda_graph=graph), _kw(enc, False), _kw(enc, True), max(LENS), device=dev, dist_group=group)
    _fill(alg.replay_buffer, Transition, np.random.RandomState(4))
    calls = [0]

    def noise(like):                  # the draw a single process would make for the FULL batch, sliced to this rank's rows
        calls[0] += 1
        g = torch.Generator(device="cpu").manual_seed(1000 + calls[0])
        full = torch.randn((len(LENS),) + tuple(like.shape[1:]), generator=g)
        return full[lo:hi].to(like.device)
    noise.graph_safe = False
    alg.policy.noise_fn = alg.target_policy.noise_fn = noise
    logs = []
    for step in range(3 if graph else 2):
        plan = alg.replay_buffer.plan_trajs(HP["sac_batch_size"], None, nest_stack_trajs=False)   # same draw on every rank
        b_dev, v_dev = alg.replay_buffer.gather_device(plan)
        lo, hi = shard_rows(b_dev.shape[0], rank, world)
        batch = alg.replay_buffer.array_to_transition(b_dev[lo:hi].contiguous())
        logs.append(alg.update_on_batch(batch, plan.total_size, v_dev[lo:hi].contiguous(), plan.lens[lo:hi]))
    torch.cuda.synchronize()
    res = {"logs": logs, "vgrad": alg.value_arena.grad.cpu(), "pgrad": alg.policy_arena.grad_full.cpu(), "v": alg.value_arena.flat.cpu(),
           "p": alg.policy_arena.flat.cpu(), "t": alg.target_arena.flat.cpu(), "guard": alg.Q_guard.state.cpu(), "alpha": alg.log_sac_alpha.detach().cpu()}
    torch.save(res, f"{out}.{world}.{rank}")
    if world > 1:
        dist.destroy_process_group()


@pytest.mark.parametrize("enc,algo", [("smamba_s16_c4_b1", "sac"), ("gilr", "td3")])
def test_dp_matches_single_process(enc, algo, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world = min(4, torch.cuda.device_count())
    world = 4 if world >= 4 else 2
    out = str(tmp_path / "dp")
    port = 29600 + os.getpid() % 1000
    mp.start_processes(_run, args=(1, port, enc, algo, out, False), nprocs=1, join=True, start_method="spawn")
    mp.start_processes(_run, args=(world, port + 1, enc, algo, out, False), nprocs=world, join=True, start_method="spawn")
    ref = torch.load(f"{out}.1.0")
    worst = 0.0
    for r in range(world):
        got = torch.load(f"{out}.{world}.{r}")
        for k in ("vgrad", "pgrad", "v", "p", "t", "alpha"):
            e = float((got[k] - ref[k]).abs().max() / (ref[k].abs().max() + 1e-30))
            tol = 1e-5 if k.endswith("grad") else 2e-5
            assert e <= tol, (k, r, e)
            worst = max(worst, e)
        assert float((got["guard"][:2] - ref["guard"][:2]).abs().max()) <= 1e-6 * float(ref["guard"][:2].abs().max())
        for a, b in zip(got["logs"], ref["logs"]):
            for k in ("critic_loss", "actor_loss", "alpha_loss", "target_q_max", "log_prob", "clip_min", "clip_max"):
                if k in b:
                    assert abs(a[k] - b[k]) <= 1e-5 * max(1.0, abs(b[k])), (k, a[k], b[k])
    print(f"{enc} {algo}: world {world} vs 1 -- worst relative difference {worst:.2e}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "dp_equivalence.log"), "a") as f:
        f.write(f"{enc} {algo}: world {world} vs 1, eager, 2 updates: worst relative difference over gradients / parameters {worst:.3e}\n")


def test_dp_graph_replay_matches_eager(tmp_path):
    """Graph-replayed data-parallel updates (NCCL between replayed segments) leave what the eager schedule leaves."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    out_e, out_g = str(tmp_path / "e"), str(tmp_path / "g")
    port = 29700 + os.getpid() % 1000

    def graph_safe_run(rank, world, port, out, graph):
        pass
    mp.start_processes(_run_graphsafe, args=(2, port, out_e, False), nprocs=2, join=True, start_method="spawn")
    mp.start_processes(_run_graphsafe, args=(2, port + 1, out_g, True), nprocs=2, join=True, start_method="spawn")
    for r in range(2):
        a, b = torch.load(f"{out_e}.{r}"), torch.load(f"{out_g}.{r}")
        for k in ("v", "p", "t"):
            e = float((a[k] - b[k]).abs().max() / (b[k].abs().max() + 1e-30))
            assert e <= 2e-4, (k, r, e)


def _run_graphsafe(rank, world, port, out, graph):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from rorl_b200.buffers.transition_buffer.replay_memory import Transition
    from rorl_b200.utility.alg_init import alg_class
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.manual_seed(3)
    np.random.seed(50)
    alg = alg_class("sac_rnn_full_horizon_redQ_sep_optim")(dict(HP, use_cuda_graph=graph), _kw("smamba_s16_c4_b1", False), _kw("smamba_s16_c4_b1", True),
                                                           max(LENS), device=dev, dist_group=dist.group.WORLD)
    _fill(alg.replay_buffer, Transition, np.random.RandomState(4 + rank))

    def noise(like):
        return torch.sin(torch.arange(like.numel(), device=like.device, dtype=torch.float32) * 12.9898).reshape(like.shape)
    noise.graph_safe = True
    alg.policy.noise_fn = alg.target_policy.noise_fn = noise
    for _ in range(5):
        alg.train_one_batch()
    if graph:
        assert any(isinstance(v, dict) for v in alg._graphs.values()), "no graph was captured"
    torch.cuda.synchronize()
    torch.save({"v": alg.value_arena.flat.cpu(), "p": alg.policy_arena.flat.cpu(), "t": alg.target_arena.flat.cpu()}, f"{out}.{rank}")
    dist.destroy_process_group()
