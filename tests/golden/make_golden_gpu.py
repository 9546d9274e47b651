"""Generate the fixtures that need a GPU by running the UNMODIFIED reference ON THE GPU BOX.

    /usr/local/graft/bin/gpurun -- 'python tests/golden/make_golden_gpu.py gpurun_out/golden_gpu'
    cp gpurun_out/golden_gpu/*.npz tests/golden/            (back in the build container; then commit)

The reference's cgpt encoder only runs on CUDA (flash-attn 2.x: `flash_attn.modules.mha.MHA` with ALiBi under bf16
autocast, ref: offpolicy_rnn/models/flash_attention/TransformerFlashAttention.py:64-121).  /root/reference does not
exist on the GPU box; the byte-identical staged copy `oracle/_ref/` (oracle/make_ref.py) is imported instead.

Fixtures:
  layer_cgpt_{ln,rms}.npz   RNNBase(['fc', cgpt_h2_l2_p0.0[_rms], 'fc']) forward + all gradients on row-packed sequences
  step_cgpt.npz             the same stack decoded one token at a time against flash-attn's kv-cache (rollout path)
  update_sac_cgpt.npz       one full train_one_batch() of SACFullLengthRNNREDQ_SEP_OPTIM with a cgpt encoder (GradScaler on)
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.refload import REF_HP, build_algorithm, install_algo_stubs, load_reference  # noqa: E402

OUT = sys.argv[1] if len(sys.argv) > 1 else HERE
os.makedirs(OUT, exist_ok=True)
load_reference(gpu_semantics=True)
dev = torch.device("cuda:0")


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().float().cpu().numpy()
        out[k] = v
    np.savez_compressed(os.path.join(OUT, name), **out)
    print("wrote", name, sum(np.asarray(v).nbytes for v in out.values()) // 1024, "KiB")


def flat_sd(sd, prefix=""):
    return {f"{prefix}{k}/{n}": t.detach().float().cpu().clone().numpy() for k, m in sd.items() for n, t in m.items()}


def gen_layers():
    from offpolicy_rnn.models.rnn_base import RNNBase
    for tag, lid in (("cgpt_ln", "cgpt_h2_l2_p0.0_ml512"), ("cgpt_rms", "cgpt_h2_l2_p0.0_ml512_rms")):
        torch.manual_seed(7)
        net = RNNBase(12, 8, [128, 128], ['elu', 'elu', 'linear'], ['fc', lid, 'fc'])
        with torch.no_grad():
            for p in net.parameters():
                if p.dim() == 1 or p.abs().max() == 0:
                    p.add_(0.1 * torch.randn_like(p))
        net.to(dev)
        net.train()
        B, L = 3, 300
        seq = np.zeros((B, L), dtype=np.int64)
        seq[0, :2] = (200, 100)
        seq[1, :3] = (1, 150, 99)          # 50 padding tokens at the end of the row
        seq[2, :1] = (212,)
        x = torch.randn(B, L, 12).to(dev).requires_grad_()
        hid = net.make_init_state(B, dev)
        hid.set_attention_concat_mask(torch.from_numpy(seq).to(torch.int).to(dev))
        y, _, _ = net.meta_forward(x, hid)
        dy = torch.randn(B, L, 8).to(dev)
        params = dict(net.named_parameters())
        grads = torch.autograd.grad(y, [x] + list(params.values()), dy, allow_unused=True)
        arrs = {"x": x, "seqlens": seq, "y": y, "dy": dy, "dx": grads[0]}
        for (n, p), gr in zip(params.items(), grads[1:]):
            arrs["p/" + n] = p
            if gr is not None:
                arrs["g/" + n] = gr
        save(f"layer_{tag}.npz", layer_id=np.array(lid), **arrs)


def gen_egilr():
    """`egilr-3` on the reference's GPU path (Triton scan: width % 256 == 0; its CPU path cannot take the zero hidden)."""
    from offpolicy_rnn.models.rnn_base import RNNBase
    lid = "egilr-3"
    torch.manual_seed(17)
    net = RNNBase(12, 2, [256, 256], ['elu', 'elu', 'linear'], ['fc', lid, 'efc-3'])
    for l in net.layer_list:
        if hasattr(l, 'desire_ndim'):
            l.desire_ndim = 4
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() == 1 or p.abs().max() == 0:
                p.add_(0.1 * torch.randn_like(p))
    net.to(dev)
    for l in net.layer_list:
        if hasattr(l, 'to') and hasattr(l, 'device'):
            l.to(dev)
    B, L = 3, 23
    x = torch.randn(B, L, 12).to(dev).requires_grad_()
    start = torch.zeros(B, L, 1)
    start[:, 0] = 1
    start[1, 9] = 1
    start = start.to(dev)
    hid = net.make_init_state(B, dev)
    hid.set_rnn_start(start)
    y, h_out, _ = net.meta_forward(x, hid)
    dy = torch.randn_like(y)
    params = dict(net.named_parameters())
    grads = torch.autograd.grad(y, [x] + list(params.values()), dy, allow_unused=True)
    arrs = {"x": x, "start": start, "mask": torch.ones(B, L, 1), "y": y, "dy": dy, "dx": grads[0], "h_out": h_out[0]}
    for (n, p), gr in zip(params.items(), grads[1:]):
        arrs["p/" + n] = p
        if gr is not None:
            arrs["g/" + n] = gr
    save("layer_egilr.npz", layer_id=np.array(lid), **arrs)


def gen_step():
    """Rollout: one token per call against the kv-cache (ref: rnn_base.py:437-452; flash-attn MHA inference path)."""
    from offpolicy_rnn.models.rnn_base import RNNBase
    lid = "cgpt_h2_l2_p0.0_ml64_rms"
    torch.manual_seed(11)
    net = RNNBase(10, 6, [128, 128], ['elu', 'elu', 'linear'], ['fc', lid, 'fc'])
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() == 1 or p.abs().max() == 0:
                p.add_(0.1 * torch.randn_like(p))
    net.to(dev)
    net.eval()
    B, T = 3, 37
    x = torch.randn(B, T, 10).to(dev)
    with torch.no_grad():
        y_full, _, _ = net.meta_forward(x, net.make_init_state(B, dev))
        hid = net.make_init_state(B, dev)
        ys = []
        for t in range(T):
            y, hid, _ = net.meta_forward(x[:, t:t + 1], hid)
            ys.append(y)
    arrs = {"x": x, "y_steps": torch.cat(ys, dim=1), "y_full": y_full}
    for n, p in net.named_parameters():
        arrs["p/" + n] = p
    save("step_cgpt.npz", layer_id=np.array(lid), **arrs)


def fill_buffer(buf, Transition, rng, lens, S, A):
    for T in lens:
        last_s, last_a, last_r = np.zeros((1, S)), np.zeros((1, A)), np.zeros((1, 1))
        s = rng.standard_normal((1, S))
        for t in range(T):
            a = np.tanh(rng.standard_normal((1, A)))
            ns = rng.standard_normal((1, S))
            r = float(rng.standard_normal())
            done = t == T - 1
            buf.mem_push(Transition(state=s, last_state=last_s, last_action=last_a, action=a, next_state=ns, reward=r,
                                    logp=None, mask=1, done=done, timeout=done, start=(t == 0), reward_input=last_r))
            last_s, last_a, last_r, s = s, a, np.array([[r]]), ns


def gen_update():
    install_algo_stubs()
    from offpolicy_rnn.buffers.transition_buffer.replay_memory import Transition
    c = dict(algo="sac", enc="cgpt_h2_l1_p0.0_ml128_rms", hidden=128, emb=32, lens=[40, 33, 25, 37, 12], S=5, A=3, calls=1)
    H, E = c["hidden"], c["emb"]
    kw = lambda value: dict(state_dim=c["S"], action_dim=c["A"], embedding_size=E, embedding_hidden=[H, H],
                            embedding_activations=['elu', 'elu', 'linear'], embedding_layer_type=['fc', c["enc"], 'fc'],
                            uni_model_hidden=[64, 64], uni_model_activations=['elu', 'elu', 'linear'],
                            uni_model_layer_type=(['efc-8'] * 3 if value else ['fc'] * 3), fix_rnn_length=0,
                            uni_model_input_mapping_dim=E, reward_input=False, last_action_input=True, last_state_input=True,
                            separate_encoder=True)
    torch.manual_seed(5)
    np.random.seed(5)
    hp = dict(REF_HP, sac_batch_size=sum(c["lens"]) - 1, max_buffer_transition_num=1000)
    A = build_algorithm("SACFullLengthRNNREDQ_SEP_OPTIM", hp, kw(False), kw(True), max(c["lens"]), c["A"], device="cuda:0", perturb=0.05)
    hp = vars(A.parameter)
    skip = type(A)._get_skip_len(A)
    fill_buffer(A.replay_buffer, Transition, np.random.RandomState(9), c["lens"], c["S"], c["A"])
    arrs = {}
    arrs.update(flat_sd(A.policy.state_dict(), "init/policy/"))
    arrs.update(flat_sd(A.values[0].state_dict(), "init/value/"))
    noises = []
    orig = torch.randn_like

    def rec(t, *a, **k):
        out = orig(t, *a, **k)
        noises.append(out.detach().clone())
        return out

    np.random.seed(21)
    torch.manual_seed(22)
    torch.randn_like = rec
    try:
        for call in range(c["calls"]):
            snap = {}
            opt = A.optimizer_value
            orig_step = opt.step

            def step_and_snap(*a, _o=orig_step, **k):
                inv = 1.0 / (A.amp_scalar_critic.get_scale() if A.amp_scalar_critic is not None else 1.0)
                # GradScaler.step() has already un-scaled .grad when it calls optimizer.step()
                snap.update({f"{k2}/{n}": (p.grad.detach().float().cpu().clone().numpy() if p.grad is not None else None)
                             for k2, m in A.values[0].contextual_modules.items() for n, p in m.named_parameters()})
                return _o(*a, **k)

            opt.step = step_and_snap
            log = A.train_one_batch()
            opt.step = orig_step
            A.grad_num += 1
            for k, v in log.items():
                if isinstance(v, tuple):
                    v = v[0]
                arrs[f"c{call}/log/{k}"] = np.array(float(v))
            for k, v in snap.items():
                if v is not None:
                    arrs[f"c{call}/vgrad/{k}"] = v
            for k2, m in A.policy.contextual_modules.items():
                for n, p in m.named_parameters():
                    if p.grad is not None:
                        arrs[f"c{call}/pgrad/{k2}/{n}"] = p.grad.detach().float().cpu().clone().numpy()
            arrs.update(flat_sd(A.policy.state_dict(), f"c{call}/policy/"))
            arrs.update(flat_sd(A.values[0].state_dict(), f"c{call}/value/"))
            arrs.update(flat_sd(A.target_values[0].state_dict(), f"c{call}/target/"))
            arrs[f"c{call}/log_alpha"] = A.log_sac_alpha.detach().cpu().clone().numpy()
    finally:
        torch.randn_like = orig
    for i, nz in enumerate(noises):
        arrs[f"noise/{i}"] = nz.cpu().numpy()
    cfg = dict(case=c, hp=hp, cls="SACFullLengthRNNREDQ_SEP_OPTIM", policy_kwargs=A.policy_args, value_kwargs=A.value_args, skip=skip,
               allow_nest_stack=bool(A.allow_nest_stack), n_noise=len(noises), n_hdraw=0, np_seed_fill=9, np_seed_run=21,
               flash_attn=__import__("flash_attn").__version__, gpu=torch.cuda.get_device_name(0))
    save("update_sac_cgpt.npz", cfg=np.array(json.dumps(cfg)), **arrs)


if __name__ == "__main__":
    for fn in (gen_layers, gen_egilr, gen_step, gen_update):
        try:
            fn()
        except Exception as e:      # keep going: each fixture is independent
            import traceback
            traceback.print_exc()
            print("FAILED", fn.__name__, e)
