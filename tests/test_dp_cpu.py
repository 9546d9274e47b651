"""Host-side data-parallel logic on CPU with gloo, world size 2: bucket planning tiles the arena,
the bucketed + overlapped gradient all-reduce equals the single-process gradient of the same globally normalised
masked loss, rows shard contiguously, and the count / guard collectives make every rank agree."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_rows_and_buckets():
    from rorl_b200.algorithm.data_parallel import plan_buckets, shard_rows
    for n, w in ((256, 8), (33, 4), (5, 8), (32, 1)):
        blocks = [shard_rows(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        assert max(h - l for l, h in blocks) - min(h - l for l, h in blocks) <= 1
    sizes = [10, 3, 128, 7, 64, 1, 300]
    offs, n = [], 0
    for s in sizes:
        offs.append(n)
        n += (s + 3) // 4 * 4
    for be in (1, 16, 200, 10 ** 6):
        bk = plan_buckets(offs, sizes, n, be)
        assert bk[0][0] == 0 and bk[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(bk, bk[1:]))
        assert sorted(i for _, _, m in bk for i in m) == list(range(len(sizes)))


def _make_model(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ELU(), torch.nn.Linear(16, 16), torch.nn.ELU(), torch.nn.Linear(16, 1))


def _flatten(model):
    """A miniature FlatArena: parameters and gradients as views of flat buffers."""
    params = list(model.parameters())
    offs, n = [], 0
    for p in params:
        offs.append(n)
        n += (p.numel() + 3) // 4 * 4
    flat, grad = torch.zeros(n), torch.zeros(n)
    with torch.no_grad():
        for p, o in zip(params, offs):
            flat[o:o + p.numel()].copy_(p.reshape(-1))
            p.data = flat[o:o + p.numel()].view(p.shape)
            p.grad = grad[o:o + p.numel()].view(p.shape)
    return params, offs, flat, grad


def _batch():
    g = torch.Generator().manual_seed(5)
    x = torch.randn(8, 11, 6, generator=g)
    y = torch.randn(8, 11, 1, generator=g)
    mask = (torch.rand(8, 11, 1, generator=g) < 0.7).float()
    return x, y, mask


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rorl_b200.algorithm.data_parallel import BucketedGradSync, shard_rows, sync_count_and_guard
    model = _make_model(0)
    params, offs, flat, grad = _flatten(model)
    x, y, mask = _batch()
    lo, hi = shard_rows(x.shape[0], rank, world)
    xs, ys, ms = x[lo:hi], y[lo:hi], mask[lo:hi]
    n_valid = ms.sum().reshape(1).clone()
    gmin = torch.tensor([float(ys.min())], dtype=torch.float64)
    gmax = torch.tensor([float(ys.max())], dtype=torch.float64)
    sync_count_and_guard(n_valid, gmin, gmax)
    sync = BucketedGradSync(params, offs, grad, bucket_bytes=256)       # several buckets on this tiny model
    assert len(sync.buckets) > 2
    for _ in range(2):                                                     # twice: the hooks re-arm correctly
        grad.zero_()
        sync.begin()
        loss = (ms * (model(xs) - ys) ** 2).sum() / n_valid               # normalised by the GLOBAL count
        loss.backward()
        sync.finish()
    torch.save({"grad": grad.clone(), "n_valid": n_valid, "gmin": gmin, "gmax": gmax}, f"{out}.{rank}")
    dist.destroy_process_group()


def test_bucketed_allreduce_matches_single_process(tmp_path):
    world = 2
    port = 29500 + os.getpid() % 2000
    out = str(tmp_path / "dp")
    mp.start_processes(_worker, args=(world, port, out), nprocs=world, join=True, start_method="spawn")
    res = [torch.load(f"{out}.{r}") for r in range(world)]
    model = _make_model(0)
    params, offs, flat, grad = _flatten(model)
    x, y, mask = _batch()
    ((mask * (model(x) - y) ** 2).sum() / mask.sum()).backward()
    for r in res:
        assert torch.equal(r["grad"], res[0]["grad"])                     # replicas agree bit for bit
        assert float(r["n_valid"]) == float(mask.sum())
        assert float(r["gmin"]) == float(y.min()) and float(r["gmax"]) == float(y.max())
        err = float((r["grad"] - grad).abs().max() / grad.abs().max())
        assert err < 1e-5, err                                            # summation order only


def test_update_stage_schedule_places_the_exchanges():
    """The update is four device segments; which data-parallel exchange follows which segment depends on the mode:
    single process -> none; eager + data parallel -> only the count / guard sync (gradient buckets are reduced by the
    hooks inside the backward); CUDA-graph replay + data parallel -> one all-reduce per gradient arena between the
    replayed segments (NCCL stays outside the captured graphs)."""
    from types import SimpleNamespace
    from rorl_b200.algorithm.full_length_update import FullLengthRNNUpdate
    calls = []

    class Stub:
        parameter = SimpleNamespace(no_alpha_auto_tune=False)
        value_arena = SimpleNamespace(grad="value_grad")
        policy_arena = SimpleNamespace(grad="policy_grad", grad_full="policy_grad_full")
        alpha_arena = SimpleNamespace(grad="alpha_grad")

        def __init__(self, group):
            self.dist_group = group

        def _sync_guard_and_count(self):
            calls.append("count")

        def _allreduce(self, t):
            calls.append(t)

    def exchanges(group, graph_mode):
        calls.clear()
        out = []
        for _, comm in FullLengthRNNUpdate._stages(Stub(group), graph_mode):
            before = len(calls)
            if comm is not None:
                comm()
            out.append(tuple(calls[before:]))
        return out

    assert exchanges(None, False) == [(), (), (), ()]
    assert exchanges(None, True) == [(), (), (), ()]
    assert exchanges("g", False) == [("count",), (), (), ()]
    assert exchanges("g", True) == [("count",), ("value_grad",), ("policy_grad_full",), ()]
