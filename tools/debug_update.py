"""Stage-by-stage comparison of the CUDA update path against the oracle on a golden update case (GPU box)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import rel_err  # noqa: E402
from test_update_gpu import build  # noqa: E402
from test_oracle_golden import run_oracle_update  # noqa: E402
from oracle import model as OM  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "sac_smamba"
g, cfg, alg = build(tag)
_, _, ora = run_oracle_update(tag)
hp = cfg["hp"]
td3 = cfg["case"]["algo"] == "td3"

np.random.seed(cfg["np_seed_run"])
tr, total, valid, lens = alg.replay_buffer.sample_trajs_device(hp["sac_batch_size"], None, nest_stack_trajs=alg.allow_nest_stack)
np.random.seed(cfg["np_seed_run"])
otr, ototal, ovalid, olens = ora.replay.sample_trajs(hp["sac_batch_size"], nest_stack_trajs=ora.allow_nest_stack)
for n in tr._fields:
    a, b = getattr(tr, n), getattr(otr, n)
    if a is not None:
        print(f"sampler {n:13s} equal={np.array_equal(a.cpu().numpy(), b.astype(np.float32))} shape={tuple(a.shape)}")
print("sampler valid equal", np.array_equal(valid.cpu().numpy(), ovalid.astype(np.float32)), "lens", np.array_equal(lens, olens), total, ototal)

f32 = lambda a: torch.from_numpy(np.array(a)).float()
ob = {n: f32(getattr(otr, n)) for n in otr._fields if getattr(otr, n) is not None}
ovalid_t = f32(ovalid)
side = OM.Side(ob["start"], ovalid_t)
noise0 = torch.from_numpy(g["noise/0"])
dev = alg.device
B = tr.state.shape[0]
# ---- policy forward on (state, last_state, last_action)
hid = alg.policy.make_init_state(B, dev)
hid.set_rnn_start(tr.start), hid.set_mask(valid)
alg.policy.noise_fn = lambda like: noise0.to(dev)
with torch.no_grad():
    am, emb, asamp, logp, _, _ = alg.policy.forward(tr.state, tr.last_state, tr.last_action, hid, tr.reward_input)
    om, oemb, osamp, ologp = OM.policy_forward(ora.policy, ora.pspec, ob["state"], ob["last_state"], ob["last_action"], side,
                                               ob["reward_input"], noise0, td3)
print("policy emb", rel_err(emb, oemb), "mean", rel_err(am, om), "sample", rel_err(asamp, osamp), "logp", rel_err(logp, ologp))
# encoder pieces
with torch.no_grad():
    ein = alg.policy.get_embedding_input(tr.state, tr.last_state, tr.last_action, tr.reward_input)
    oein = OM.embedding_input(ora.policy, ora.pspec, ob["state"], ob["last_state"], ob["last_action"], ob["reward_input"])
    print("policy emb input", rel_err(ein, oein))
    net = alg.policy.embedding_network
    x = net.layer_list[0](ein)
    x = net.activation_list[0](x)
    p = ora.policy["embedding_model"]
    ox = torch.nn.functional.elu(torch.nn.functional.linear(oein, p["layer_list.0.weight"], p["layer_list.0.bias"]))
    print("pre-fc", rel_err(x, ox))
    lid = net.layer_type[1]
    if lid.startswith("smamba"):
        y, _ = net.layer_list[1](x, None, tr.start, valid)
        oy = OM.smamba_layer(p, "layer_list.1.", ox, side, lid)
    elif lid == "gilr":
        y, _ = net.layer_list[1](x, None, tr.start)
        oy = OM.gilr_layer(p, "layer_list.1.", ox, side)
    elif lid == "lru":
        y, _ = net.layer_list[1](x, None, tr.start, None)
        oy = OM.lru_layer(p, "layer_list.1.", ox, side)
    else:
        y, _ = net.layer_list[1](x, None)
        oy = OM.gru_layer(p, "layer_list.1.", ox, side)
    print("encoder layer", lid, rel_err(y, oy))
# ---- value forward
hv = alg.values[0].make_init_state(B, dev)
hv.set_rnn_start(tr.start), hv.set_mask(valid)
with torch.no_grad():
    q, qemb, _, _ = alg.values[0].forward(tr.state, tr.last_state, tr.last_action, tr.action, hv, tr.reward_input)
    oq, oqemb = OM.value_forward(ora.value, ora.vspec, ob["state"], ob["last_state"], ob["last_action"], ob["action"], side, ob["reward_input"])
print("value emb", rel_err(qemb, oqemb), "q", rel_err(q, oq), tuple(q.shape), tuple(oq.shape))
# ---- fused reductions vs torch
with torch.no_grad():
    E = q.shape[0]
    sel = np.array([3, 5])
    alg.Q_guard.reset()
    y = alg._target_Q(q, sel, logp, tr.reward, tr.done, tr.timeout, tr.mask)
    alpha = alg.log_sac_alpha.exp()
    m = q[sel].min(dim=0).values - alpha * logp
    done = tr.done.clone(); done[tr.timeout > 0] = 0
    yref = tr.reward + (1 - done) * hp["gamma"] * m
    print("target_Q kernel vs torch", rel_err(y, yref), "stats", alg._stats[:3].tolist(), float(tr.mask.sum()), float(yref.abs().max()))
    print("guard", alg.Q_guard.state.tolist(), float((yref * tr.mask).min()), float((yref * tr.mask).max()), float(m.min()), float(m.max()))
    M = q[0].numel()
    dq = torch.empty((E, M), device=dev)
    import rorl_b200._native as N
    N.call("rorl_q_loss_fwd_bwd", N.ptr(q.contiguous()), N.ptr(y.contiguous()), N.ptr(tr.mask.contiguous()), N.ptr(alg._stats[1:2]),
           N.ptr(alg._stats[2:3]), N.ptr(dq), N.ptr(alg._work), E, M, N.stream())
    nv = tr.mask.sum()
    lref = (((q - y.unsqueeze(0)) ** 2).sum(0) * tr.mask).sum() / nv
    dref = 2 * (q - y.unsqueeze(0)) * tr.mask / nv
    print("q_loss kernel vs torch", float(alg._stats[2]), float(lref), "dq", rel_err(dq.view_as(q), dref))
