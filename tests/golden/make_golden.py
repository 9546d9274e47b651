"""Generate the committed golden fixtures by running the UNMODIFIED reference (/root/reference) on CPU.

Run here (the container that has /root/reference):   python tests/golden/make_golden.py
Nothing on the GPU box runs this; tests read the .npz files it writes next to itself.

Fixtures:
  ops_*.npz      op-level in/out/grad of the reference's own restatements (scan_cpu, complex_scan_cpu,
                 selective_scan_ref, layernorm_cpu)
  layer_*.npz    RNNBase(['fc', <ID>, 'fc']) forward + input/parameter gradients for each encoder ID
  step_*.npz     smamba rollout path: L single steps through Mamba.step() with a carried hidden (reference CPU path)
  sampler_*.npz  NestedMemoryArray.sample_trajs outputs for seeded buffers
  update_*.npz   one or two full train_one_batch() calls of the reference algorithm classes (App. D harness)
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.refload import REF_HP, build_algorithm, install_algo_stubs, load_reference  # noqa: E402

load_reference(gpu_semantics=True)
torch.set_num_threads(4)


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = v
    np.savez_compressed(os.path.join(HERE, name), **out)
    print("wrote", name, sum(np.asarray(v).nbytes for v in out.values()) // 1024, "KiB")


def flat_sd(sd, prefix=""):
    return {f"{prefix}{k}/{n}": t.detach().clone().numpy() for k, m in sd.items() for n, t in m.items()}


# ------------------------------------------------------------------------------------------------
# op-level
# ------------------------------------------------------------------------------------------------
def gen_ops():
    from offpolicy_rnn.models.gilr.scan_triton.real_rnn_tie_input_gate_cpu import scan_cpu
    from offpolicy_rnn.models.lru.scan_triton.complex_rnn_cpu import complex_scan_cpu
    from offpolicy_rnn.models.smamba.mamba_ssm.ops.selective_scan_interface_new import selective_scan_ref
    from offpolicy_rnn.models.smamba.mamba_ssm.ops.triton.layernorm_cpu import layer_norm_fn, rms_norm_fn

    g = torch.Generator().manual_seed(1)
    rn = lambda *s: torch.randn(*s, generator=g)
    # gilr
    B, L, C = 3, 37, 8
    v = rn(B, L, C).requires_grad_()
    f = torch.sigmoid(rn(B, L, C))
    f[:, [0, 5, 6, 20]] = 0.0
    f.requires_grad_()
    h, _ = scan_cpu(v, f, torch.zeros(B, 1, C))
    dh = rn(B, L, C)
    dv, df = torch.autograd.grad(h, (v, f), dh)
    save("ops_gilr.npz", v=v, f=f, h=h, dh=dh, dv=dv, df=df)
    # lru
    B, L, C = 2, 29, 8
    vr, vi = rn(B, L, C).requires_grad_(), rn(B, L, C).requires_grad_()
    fr = (0.9 * torch.cos(rn(B, L, C))).requires_grad_()
    fi = (0.9 * torch.sin(rn(B, L, C))).requires_grad_()
    h0r, h0i = rn(B, 1, C), rn(B, 1, C)
    hr, hi, _, _ = complex_scan_cpu(vr, vi, fr, fi, h0r, h0i)
    gr, gi = rn(B, L, C), rn(B, L, C)
    grads = torch.autograd.grad((hr, hi), (vr, vi, fr, fi), (gr, gi))
    save("ops_lru.npz", vr=vr, vi=vi, fr=fr, fi=fi, h0r=h0r, h0i=h0i, hr=hr, hi=hi, gr=gr, gi=gi,
         dvr=grads[0], dvi=grads[1], dfr=grads[2], dfi=grads[3])
    # selective scan (reference layout)
    for tag, (B, D, L, N) in {"a": (2, 8, 45, 16), "b": (1, 4, 70, 32)}.items():
        u, delta, z = rn(B, D, L).requires_grad_(), (0.5 * rn(B, D, L)).requires_grad_(), rn(B, D, L).requires_grad_()
        A = (-torch.exp(0.3 * rn(D, N))).requires_grad_()
        Bm, Cm = rn(B, N, L).requires_grad_(), rn(B, N, L).requires_grad_()
        Dk, bias = rn(D).requires_grad_(), rn(D).requires_grad_()
        st = torch.zeros(B, 1, L)
        st[:, :, 0] = 1
        st[0, :, 17] = 1
        st[0, :, 18] = 1
        st[-1, :, 33] = 1
        start = st.expand(B, D, L).contiguous()
        out, last = selective_scan_ref(u, delta, A, Bm, Cm, start, Dk, z=z, delta_bias=bias, delta_softplus=True,
                                       return_last_state=True)
        dout = rn(B, D, L)
        gs = torch.autograd.grad(out, (u, delta, A, Bm, Cm, Dk, z, bias), dout)
        save(f"ops_selscan_{tag}.npz", u=u, delta=delta, A=A, B=Bm, C=Cm, D=Dk, z=z, bias=bias, start=start, out=out,
             last=last, dout=dout, du=gs[0], ddelta=gs[1], dA=gs[2], dB=gs[3], dC=gs[4], dD=gs[5], dz=gs[6], dbias=gs[7])
    # add + norm
    x, r = rn(10, 24).requires_grad_(), rn(10, 24).requires_grad_()
    w, b = rn(24).requires_grad_(), rn(24).requires_grad_()
    y, res = layer_norm_fn(x, w, b, residual=r, eps=1e-8, prenorm=True, residual_in_fp32=True)
    dy, dres = rn(10, 24), rn(10, 24)
    gs = torch.autograd.grad((y, res), (x, r, w, b), (dy, dres))
    y2 = rms_norm_fn(x, w, None, residual=r, eps=1e-8, prenorm=False, residual_in_fp32=True)
    gs2 = torch.autograd.grad(y2, (x, r, w), dy)
    save("ops_addnorm.npz", x=x, r=r, w=w, b=b, y=y, res=res, dy=dy, dres=dres, dx=gs[0], dr=gs[1], dw=gs[2], db=gs[3],
         y_rms=y2, dx_rms=gs2[0], dr_rms=gs2[1], dw_rms=gs2[2])


# ------------------------------------------------------------------------------------------------
# layer-level
# ------------------------------------------------------------------------------------------------
LAYER_IDS = {"gilr": "gilr", "lru": "lru", "gru": "gru", "smamba_rms": "smamba_s16_c4_b2",
             "smamba_ln": "smamba_s32_c8_b1_nln", "smamba_ff": "smamba_s16_c8_b1_ff",
             # s6 layer: reference CPU path = selective_scan_cpu, explicit conv left state
             "mamba_ff": "mamba_s16_c4", "mamba_noff": "mamba_s32_c8_noff", "mamba_h0": "mamba_s16_c4",
             # remaining encoder IDs; *_h0 = a carried non-zero state entering the call
             "gilr_lstm": "gilr_lstm", "gilr_lstm_h0": "gilr_lstm", "conv1d": "conv1d_8", "conv1d_h0": "conv1d"}


def gen_layers(only=None):
    from offpolicy_rnn.models.rnn_base import RNNBase
    for tag, lid in LAYER_IDS.items():
        if only and tag not in only:
            continue
        torch.manual_seed(7)
        net = RNNBase(12, 8, [16, 16], ['elu', 'elu', 'linear'], ['fc', lid, 'fc'])
        with torch.no_grad():   # make every parameter non-trivial (zero biases hide bugs)
            for p in net.parameters():
                if p.dim() == 1 or p.abs().max() == 0:
                    p.add_(0.1 * torch.randn_like(p))
        B, L = 3, 41
        x = torch.randn(B, L, 12, requires_grad=True)
        start = torch.zeros(B, L, 1)
        start[:, 0] = 1
        start[1, 15] = 1
        start[2, 30:33] = 1
        mask = torch.ones(B, L, 1)
        mask[2, 28:33] = 0
        hid = net.make_init_state(B)
        if 'gru' not in lid:
            hid.set_rnn_start(start)
            hid.set_mask(mask)
        if tag.endswith("_h0"):     # carried (non-zero) state entering the call
            hid[0] = 0.3 * torch.randn_like(hid[0])
        h_in = [h.clone() for h in hid] if (lid.startswith('mamba') or lid.startswith('conv1d') or lid == 'gilr_lstm') else None
        y, h_out, _ = net.meta_forward(x, hid)
        dy = torch.randn_like(y)
        params = dict(net.named_parameters())
        grads = torch.autograd.grad(y, [x] + list(params.values()), dy, allow_unused=True)
        arrs = {"x": x, "start": start, "mask": mask, "y": y, "dy": dy, "dx": grads[0]}
        if h_in is not None:
            arrs["h_in"], arrs["h_out"] = h_in[0], h_out[0]
        for (n, p), gr in zip(params.items(), grads[1:]):
            arrs["p/" + n] = p
            if gr is not None:   # e.g. GILRLayer.layer_norm is constructed but never used
                arrs["g/" + n] = gr
        save(f"layer_{tag}.npz", layer_id=np.array(lid), **arrs)


ENSEMBLE_LAYER_IDS = {"elru": "elru-3", "econv1d": "econv1d_4-3", "egilr_lstm": "egilr_lstm-3", "elru_h0": "elru-3"}


def gen_ensemble_layers():
    """Ensemble encoder IDs (one independent encoder per ensemble member, output [E, B, L, C]):
    RNNBase(['fc', <ID>, 'efc-3']) forward + all gradients on the reference's CPU path.  (`egilr-E` cannot run on the
    reference's CPU path -- its zero hidden has the wrong shape for scan_cpu -- and is generated on the GPU box by
    make_golden_gpu.py.)"""
    from offpolicy_rnn.models.rnn_base import RNNBase
    for tag, lid in ENSEMBLE_LAYER_IDS.items():
        torch.manual_seed(17)
        net = RNNBase(12, 2, [16, 16], ['elu', 'elu', 'linear'], ['fc', lid, 'efc-3'])
        for l in net.layer_list:
            if hasattr(l, 'desire_ndim'):
                l.desire_ndim = 4
        with torch.no_grad():
            for p in net.parameters():
                if p.dim() == 1 or p.abs().max() == 0:
                    p.add_(0.1 * torch.randn_like(p))
        B, L = 3, 23
        x = torch.randn(B, L, 12, requires_grad=True)
        start = torch.zeros(B, L, 1)
        start[:, 0] = 1
        start[1, 9] = 1
        mask = torch.ones(B, L, 1)
        mask[2, 15:18] = 0
        hid = net.make_init_state(B)
        hid.set_rnn_start(start)
        hid.set_mask(mask)
        arrs = {}
        if tag.endswith("_h0"):
            hid[0] = 0.3 * torch.randn_like(hid[0])
            arrs["h_in"] = hid[0].clone()
        y, h_out, _ = net.meta_forward(x, hid)
        dy = torch.randn_like(y)
        params = dict(net.named_parameters())
        grads = torch.autograd.grad(y, [x] + list(params.values()), dy, allow_unused=True)
        arrs.update({"x": x, "start": start, "mask": mask, "y": y, "dy": dy, "dx": grads[0], "h_out": h_out[0]})
        for (n, p), gr in zip(params.items(), grads[1:]):
            arrs["p/" + n] = p
            if gr is not None:
                arrs["g/" + n] = gr
        save(f"layer_{tag}.npz", layer_id=np.array(lid), **arrs)


def gen_steps():
    """Rollout step path of the smamba layer: the UNMODIFIED reference on CPU walks Mamba.step() once per time step
    (ref: smamba/mamba.py:133-159,257-305) -- its natural CPU path, no patching (a fresh import would be needed to
    undo _refload's forward_sequential routing, so the original forward is recovered from the class dict)."""
    import importlib
    import offpolicy_rnn.models.smamba.mamba as smamba
    src = importlib.util.find_spec("offpolicy_rnn.models.smamba.mamba").origin
    ns = {}
    orig_forward = None
    # re-exec the module source in a scratch namespace to get the original Mamba.forward function object
    spec = importlib.util.spec_from_file_location("_smamba_orig", src, submodule_search_locations=None)
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = "offpolicy_rnn.models.smamba"
    spec.loader.exec_module(mod)
    orig_forward = mod.Mamba.forward
    patched = smamba.Mamba.forward
    smamba.Mamba.forward = orig_forward
    try:
        from offpolicy_rnn.models.rnn_base import RNNBase
        for tag, lid in (("smamba_rms", "smamba_s16_c4_b2"), ("smamba_ln16", "smamba_s32_c16_b1_nln")):
            torch.manual_seed(11)
            net = RNNBase(12, 8, [16, 16], ['elu', 'elu', 'linear'], ['fc', lid, 'fc'])
            with torch.no_grad():
                for p in net.parameters():
                    if p.dim() == 1 or p.abs().max() == 0:
                        p.add_(0.1 * torch.randn_like(p))
            B, L = 3, 7
            x = torch.randn(B, L, 12)
            hid = net.make_init_state(B)
            hid[0] = 0.3 * torch.randn_like(hid[0])
            h_in = hid[0].clone()
            ys, h = [], hid
            with torch.no_grad():
                for t in range(L):
                    y, h, _ = net.meta_forward(x[:, t:t + 1], h)
                    ys.append(y)
            arrs = {"x": x, "h_in": h_in, "y": torch.cat(ys, dim=1), "h_out": h[0]}
            for n, p in net.named_parameters():
                arrs["p/" + n] = p
            save(f"step_{tag}.npz", layer_id=np.array(lid), **arrs)
    finally:
        smamba.Mamba.forward = patched
    gen_steps_linear()


def gen_steps_linear():
    """Rollout step path (L == 1 calls with the hidden carried between them) of the gilr / lru / gru / gilr_lstm / conv1d
    encoders: the UNMODIFIED reference on CPU (scan_cpu / complex_scan_cpu / torch.nn.GRU), incl. a mid-sequence reset
    (ref: gilr/gilr.py:44-67, lru/lru.py:70-174, rnn_base.py:424-454).  SURVEY.md 8 f2."""
    from offpolicy_rnn.models.rnn_base import RNNBase
    for lid in ("gilr", "lru", "gru", "gilr_lstm", "conv1d_4"):
        torch.manual_seed(13)
        net = RNNBase(12, 8, [16, 16], ['elu', 'elu', 'linear'], ['fc', lid, 'fc'])
        with torch.no_grad():
            for p in net.parameters():
                if p.dim() == 1 or p.abs().max() == 0:
                    p.add_(0.1 * torch.randn_like(p))
        B, L = 3, 9
        x = torch.randn(B, L, 12)
        start = torch.zeros(B, L, 1)
        start[1, 4] = 1                       # an episode boundary in the middle of the rollout (row 1)
        hid = net.make_init_state(B)
        hid[0] = 0.3 * torch.randn_like(hid[0])
        h_in = hid[0].clone()
        ys, h = [], hid
        with torch.no_grad():
            for t in range(L):
                if lid != "gru":
                    h.set_rnn_start(start[:, t:t + 1])
                    h.set_mask(torch.ones(B, 1, 1))
                y, h, _ = net.meta_forward(x[:, t:t + 1], h)
                ys.append(y)
        arrs = {"x": x, "start": start, "h_in": h_in, "y": torch.cat(ys, dim=1), "h_out": h[0]}
        for n, p in net.named_parameters():
            arrs["p/" + n] = p
        save(f"step_{lid.split('_')[0] if lid.startswith('conv1d') else lid}.npz", layer_id=np.array(lid), **arrs)


# ------------------------------------------------------------------------------------------------
# sampler
# ------------------------------------------------------------------------------------------------
def fill_buffer_discrete(buf, Transition, rng, lens, S, A):
    """Discrete action space: `action` is the index [1, 1], `last_action` its one-hot [1, A] (what the last-action encoder
    of the discrete models takes, ref: contextual_sac_discrete_policy.py:35-36)."""
    for T in lens:
        last_s, last_a, last_r = np.zeros((1, S)), np.zeros((1, A)), np.zeros((1, 1))
        s = rng.standard_normal((1, S))
        for t in range(T):
            ai = int(rng.randint(A))
            ns = rng.standard_normal((1, S))
            r = float(rng.standard_normal())
            done = t == T - 1
            buf.mem_push(Transition(state=s, last_state=last_s, last_action=last_a, action=np.array([[float(ai)]]), next_state=ns, reward=r,
                                    logp=None, mask=1, done=done, timeout=done, start=(t == 0), reward_input=last_r))
            last_a = np.zeros((1, A))
            last_a[0, ai] = 1.0
            last_s, last_r, s = s, np.array([[r]]), ns


def fill_buffer(buf, Transition, rng, lens, S, A):
    """Synthetic trajectories in the shape SAC.train() pushes them (ref: algorithm/sac.py:337-351)."""
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


def gen_sampler(only=None):
    from offpolicy_rnn.buffers.transition_buffer.nested_replay_memory import NestedMemoryArray
    from offpolicy_rnn.buffers.transition_buffer.replay_memory import Transition
    cases = {
        "a": dict(skip_extra=0, max_step=20, lens=[5, 3, 6, 4, 20, 9, 1, 12], batch=30, nest=True, S=3, A=2),
        "b": dict(skip_extra=4, max_step=20, lens=[5, 3, 6, 4, 20, 9, 2, 12, 7], batch=40, nest=True, S=2, A=1),
        "c": dict(skip_extra=0, max_step=12, lens=[12, 12, 5, 12, 7], batch=35, nest=False, S=2, A=2),
        "d": dict(skip_extra=16, max_step=50, lens=[50] * 6, batch=6 * 50 - 1, nest=True, S=3, A=2),
        # sampler options: randomised masks (equalised per trajectory / global) and random truncation
        "e": dict(skip_extra=0, max_step=20, lens=[5, 3, 6, 4, 20, 9, 1, 12], batch=30, nest=True, S=3, A=2,
                  randomize_mask=True, equalize=True, valid_num=14),
        "f": dict(skip_extra=4, max_step=20, lens=[5, 3, 6, 4, 20, 9, 2, 12, 7], batch=20, nest=True, S=2, A=1,
                  random_trunc=True),
        "g": dict(skip_extra=0, max_step=12, lens=[12, 12, 5, 12, 7], batch=35, nest=False, S=2, A=2,
                  randomize_mask=True, equalize=False, valid_num=9),
        "h": dict(skip_extra=1, max_step=16, lens=[7, 3, 6, 4, 9, 1, 12], batch=25, nest=True, S=2, A=2,
                  randomize_mask=True, equalize=False, valid_num=6),
        # a row filled to the full row length: the one case in which the reference's global mask randomisation actually
        # reaches the batch (its `mask.reshape((-1,))` is a view only then)
        "i": dict(skip_extra=0, max_step=14, lens=[14, 6, 7], batch=26, nest=True, S=2, A=2,
                  randomize_mask=True, equalize=False, valid_num=10),
    }
    for tag, c in cases.items():
        if only and tag not in only:
            continue
        buf = NestedMemoryArray(500, c["max_step"], additional_history_len=c["skip_extra"])
        fill_buffer(buf, Transition, np.random.RandomState(3), c["lens"], c["S"], c["A"])
        np.random.seed(11)
        arrs = {}
        for call in range(2):   # second call exercises the cached-array reuse path
            tr, total, valid, lens = buf.sample_trajs(c["batch"], None, randomize_mask=c.get("randomize_mask", False),
                                                      valid_number_post_randomized=c.get("valid_num", 0),
                                                      equalize_data_of_each_traj=c.get("equalize", True),
                                                      random_trunc_traj=c.get("random_trunc", False),
                                                      nest_stack_trajs=c["nest"])
            for n in tr._fields:
                v = getattr(tr, n)
                if v is not None:
                    arrs[f"c{call}/{n}"] = np.array(v, copy=True)
            arrs[f"c{call}/valid"] = np.array(valid, copy=True)
            arrs[f"c{call}/lens"] = lens
            arrs[f"c{call}/total"] = np.array(total)
            arrs[f"c{call}/rng_next"] = np.array(np.random.get_state()[1][:4], dtype=np.int64)
            arrs[f"c{call}/rng_pos"] = np.array(np.random.get_state()[2])
        save(f"sampler_{tag}.npz", cfg=np.array(json.dumps(c)), **arrs)


# ------------------------------------------------------------------------------------------------
# full update (App. D harness)
# ------------------------------------------------------------------------------------------------
UPDATE_CASES = {
    "sac_smamba": dict(algo="sac", enc="smamba_s16_c4_b2_nln", hidden=16, lens=[14, 9, 20, 6, 11], S=3, A=2, calls=2),
    "sac_gru": dict(algo="sac", enc="gru", hidden=16, lens=[14, 9, 20, 6, 11], S=3, A=2, calls=2),
    "td3_gilr": dict(algo="td3", enc="gilr", hidden=16, lens=[14, 9, 20, 6, 11], S=3, A=2, calls=1),
    "td3_lru": dict(algo="td3", enc="lru", hidden=16, lens=[14, 9, 20, 6, 11], S=3, A=2, calls=1),
    # the plain (non-REDQ, non-SEP_OPTIM) classes: min over all 8 members for target and actor, ONE AdamW group per
    # model (ref: sac.py:81-90); the TD3 one takes the target action from the frozen target policy.  The TD3 case also
    # carries weight decay: gilr's unused layer_norm keeps .grad None, so torch.optim.AdamW must not decay it.
    "sac_ensembleq": dict(algo="sac", cls="SACFullLengthRNNEnsembleQ", enc="lru", hidden=16, lens=[14, 9, 20, 6, 11], S=3, A=2, calls=2),
    "td3_ensembleq": dict(algo="td3", cls="TD3FullLengthRNNEnsembleQ", enc="gilr", hidden=16, lens=[14, 9, 20, 6, 11], S=3, A=2, calls=2,
                          hp=dict(policy_l2_norm=1e-2, value_l2_norm=1e-2)),
    "sac_ensembleq_sep": dict(algo="sac", cls="SACFullLengthRNNENSEMBLEQ_SEP_OPTIM", enc="gilr", hidden=16, lens=[14, 9, 20, 6, 11], S=3, A=2, calls=1),
    # mid-size: every projection is wide enough (K >= 32) for the tensor-core GEMM path
    "sac_smamba_mid": dict(algo="sac", enc="smamba_s16_c4_b1_nln", hidden=64, emb=32, lens=[40, 33, 25, 37], S=5, A=3, calls=1),
    "td3_gilr_mid": dict(algo="td3", enc="gilr", hidden=64, emb=32, lens=[40, 33, 25, 37], S=5, A=3, calls=1),
    # conv1d encoder with nest-stacked trajectories: isolation between neighbours rests on skip_len = d_conv + 1 blanks
    "sac_conv1d": dict(algo="sac", enc="conv1d_4", hidden=16, lens=[5, 3, 6, 4, 9, 7], S=3, A=2, calls=1, max_len=20),
    # gradient clipping (ref: sac_full_length_rnn_ensembleQ.py:239-250,274-287): global norm only / norm + value (+ A_log 1e-3)
    "sac_gru_clipnorm": dict(algo="sac", enc="gru", hidden=16, lens=[14, 9, 20, 6, 11], S=3, A=2, calls=1,
                             hp=dict(policy_max_gradnorm=0.05, value_max_gradnorm=0.5)),
    "sac_smamba_clipval": dict(algo="sac", enc="smamba_s16_c4_b2_nln", hidden=16, lens=[14, 9, 20, 6, 11], S=3, A=2, calls=1,
                               hp=dict(policy_max_gradnorm=0.05, value_max_gradnorm=0.5, policy_embedding_max_gradnorm=2e-3,
                                       value_embedding_max_gradnorm=2e-2)),
    # utd = 2 with one policy update per call (ref :311,405)
    "sac_gru_utd2": dict(algo="sac", enc="gru", hidden=16, lens=[14, 9, 20, 6, 11], S=3, A=2, calls=2, hp=dict(utd=2, policy_utd=1)),
    # random carried state shared by the target-policy and actor passes (ref :345-351); the draws are recorded (gru: the one
    # encoder without reset flags, so the carried state actually reaches the outputs)
    # discrete action space: categorical policy, one Q per action, expectation over actions in target
    # and actor, fixed alpha (ref: sac_full_length_rnn_ensembleQ.py:134-185, sac_full_length_rnn_redq.py:52-88, sac.py:72-74)
    "sac_discrete": dict(algo="sac", enc="gilr", hidden=16, lens=[14, 9, 20, 6, 11], S=3, A=4, calls=2, discrete=True, hp=dict(sac_alpha=0.2)),
    "sac_gru_rndhidden": dict(algo="sac", enc="gru", hidden=16, lens=[14, 9, 20, 6, 11], S=3, A=2, calls=1, hp=dict(randomize_first_hidden=True)),
}


def model_kwargs(c, value):
    H = c["hidden"]
    E = c.get("emb", 8)
    return dict(state_dim=c["S"], action_dim=c["A"], embedding_size=E, embedding_hidden=[H, H],
                embedding_activations=['elu', 'elu', 'linear'], embedding_layer_type=['fc', c["enc"], 'fc'],
                uni_model_hidden=[H, H], uni_model_activations=['elu', 'elu', 'linear'],
                uni_model_layer_type=(['efc-8'] * 3 if value else ['fc'] * 3), fix_rnn_length=0,
                uni_model_input_mapping_dim=E, reward_input=False, last_action_input=True, last_state_input=True,
                separate_encoder=True)


HP = REF_HP


def gen_updates(only=None):
    install_algo_stubs()
    from offpolicy_rnn.buffers.transition_buffer.replay_memory import Transition

    for tag, c in UPDATE_CASES.items():
        if only and tag not in only:
            continue
        torch.manual_seed(5)
        np.random.seed(5)
        cls_name = c.get("cls") or ("SACFullLengthRNNREDQ_SEP_OPTIM" if c["algo"] == "sac" else "TD3FullLengthRNNREDQ_SEP_OPTIM")
        hp = dict(HP, sac_batch_size=sum(c["lens"]) - 1, max_buffer_transition_num=1000)
        hp.update(c.get("hp", {}))
        pk, vk = model_kwargs(c, False), model_kwargs(c, True)
        A = build_algorithm(cls_name, hp, pk, vk, c.get("max_len", max(c["lens"])), c["A"], perturb=0.05, discrete=c.get("discrete", False))
        hp, pk = vars(A.parameter), A.policy_args
        skip = type(A)._get_skip_len(A)
        (fill_buffer_discrete if c.get("discrete") else fill_buffer)(A.replay_buffer, Transition, np.random.RandomState(9), c["lens"], c["S"], c["A"])

        arrs = {}
        arrs.update(flat_sd(A.policy.state_dict(), "init/policy/"))
        arrs.update(flat_sd(A.values[0].state_dict(), "init/value/"))
        noises = []
        orig_randn_like = torch.randn_like

        def rec_randn_like(t, *a, **k):
            out = orig_randn_like(t, *a, **k)
            noises.append(out.detach().clone())
            return out

        hdraws = []
        orig_randn = torch.rand

        def rec_randn(*a, **k):
            out = orig_randn(*a, **k)
            hdraws.append(out.detach().clone())
            return out

        np.random.seed(21)
        torch.manual_seed(22)
        torch.randn_like = rec_randn_like
        if hp["randomize_first_hidden"]:
            torch.rand = rec_randn
        try:
            for call in range(c["calls"]):
                snap = {}
                orig_step = A.optimizer_value.step

                def step_and_snap(*a, _orig=orig_step, _snap=snap, **k):
                    _snap.update({f"{k2}/{n}": (p.grad.detach().clone().numpy() if p.grad is not None else None)
                                  for k2, m in A.values[0].contextual_modules.items() for n, p in m.named_parameters()})
                    return _orig(*a, **k)

                A.optimizer_value.step = step_and_snap
                log = A.train_one_batch()
                A.optimizer_value.step = orig_step
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
                            arrs[f"c{call}/pgrad/{k2}/{n}"] = p.grad.detach().clone().numpy()
                arrs.update(flat_sd(A.policy.state_dict(), f"c{call}/policy/"))
                arrs.update(flat_sd(A.values[0].state_dict(), f"c{call}/value/"))
                arrs.update(flat_sd(A.target_values[0].state_dict(), f"c{call}/target/"))
                arrs[f"c{call}/log_alpha"] = A.log_sac_alpha.detach().clone().numpy()
        finally:
            torch.randn_like = orig_randn_like
            torch.rand = orig_randn
        for i, nz in enumerate(noises):
            arrs[f"noise/{i}"] = nz.numpy()
        for i, nz in enumerate(hdraws):
            arrs[f"hdraw/{i}"] = nz.numpy()
        cfg = dict(case=c, hp=hp, cls=cls_name, policy_kwargs=pk, value_kwargs=vk, skip=skip, allow_nest_stack=bool(A.allow_nest_stack),
                   n_noise=len(noises), n_hdraw=len(hdraws), np_seed_fill=9, np_seed_run=21)
        save(f"update_{tag}.npz", cfg=np.array(json.dumps(cfg)), **arrs)


# ------------------------------------------------------------------------------------------------
# full size (BASELINE.json config 2: 32 trajectories x 1000 steps, width 256) -- scalars only
# ------------------------------------------------------------------------------------------------
FULLSIZE_BIG = 16384          # parameters with at least this many elements are regenerated from their name (helpers.det_uniform)


def gen_full_size(tags=None):
    """The reference's own train_one_batch at the benchmark shape.  Inputs are reproducible without being stored: the
    replay rows come from bench.synth_trajectory(RandomState(1000)), every large weight matrix is helpers.det_uniform(name)
    scaled to the reference initialiser's range, the Gaussian draws are helpers.det_normal(i).  The fixture keeps the
    small parameters, the per-parameter ranges, the returned log and one gradient norm per parameter."""
    sys.path.insert(0, os.path.dirname(HERE))
    from helpers import det_normal, det_uniform
    import bench
    install_algo_stubs()
    from offpolicy_rnn.buffers.transition_buffer.replay_memory import Transition
    cases = {"sac_smamba": dict(algo="sac", enc="smamba_s32_c16_b2_nln", n_traj=32, t_len=1000),
             "td3_gilr": dict(algo="td3", enc="gilr", n_traj=32, t_len=1000)}
    for tag, c in cases.items():
        if tags and tag not in tags:
            continue
        torch.manual_seed(5)
        np.random.seed(5)
        S, Ad, n_traj, t_len = bench.S_DIM, bench.A_DIM, c["n_traj"], c["t_len"]
        cls_name = "SACFullLengthRNNREDQ_SEP_OPTIM" if c["algo"] == "sac" else "TD3FullLengthRNNREDQ_SEP_OPTIM"
        hp = dict(REF_HP, sac_batch_size=n_traj * t_len - 1, max_buffer_transition_num=n_traj * t_len + 8)
        pk, vk = bench.model_kwargs(c["enc"], False), bench.model_kwargs(c["enc"], True)
        A = build_algorithm(cls_name, hp, pk, vk, t_len, Ad, perturb=0.05)
        hp, pk = vars(A.parameter), A.policy_args
        arrs, bounds = {}, {}
        with torch.no_grad():
            for side, model in (("policy", A.policy), ("value", A.values[0])):
                for mod, m in model.contextual_modules.items():
                    for n, p in m.named_parameters():
                        key = f"{side}/{mod}/{n}"
                        if p.numel() >= FULLSIZE_BIG:
                            bounds[key] = float(p.abs().max())
                            p.copy_(torch.from_numpy(det_uniform(key, p.shape, bounds[key])))
                        else:
                            arrs["init/" + key] = p.detach().clone().numpy()
        A._value_update(tau=0.0)
        A.target_policy.copy_weight_from(A.policy, tau=0.0)
        rng = np.random.RandomState(1000)
        for _ in range(n_traj):
            rows = bench.synth_trajectory(rng, t_len)
            for t in range(t_len):
                r = rows[t]
                last = t == t_len - 1
                A.replay_buffer.mem_push(Transition(
                    state=r[None, 0:S], last_state=r[None, S:2 * S], last_action=r[None, 2 * S:2 * S + Ad],
                    action=r[None, 2 * S + Ad:2 * S + 2 * Ad], next_state=r[None, 2 * S + 2 * Ad:3 * S + 2 * Ad],
                    reward=float(r[3 * S + 2 * Ad]), logp=None, mask=1, done=last, timeout=last, start=(t == 0),
                    reward_input=r[None, 3 * S + 2 * Ad + 4:3 * S + 2 * Ad + 5]))
        draws = [0]
        orig_randn_like = torch.randn_like

        def det_randn_like(t, *a, **k):
            out = torch.from_numpy(det_normal(draws[0], t.shape)).to(t.dtype)
            draws[0] += 1
            return out

        np.random.seed(21)
        torch.randn_like = det_randn_like
        vnorm = {}
        orig_step = A.optimizer_value.step

        def step_and_snap(*a, **k):
            for mod, m in A.values[0].contextual_modules.items():
                for n, p in m.named_parameters():
                    if p.grad is not None:
                        vnorm[f"{mod}/{n}"] = float(p.grad.double().norm())
            return orig_step(*a, **k)

        A.optimizer_value.step = step_and_snap
        try:
            import time
            t0 = time.time()
            log = A.train_one_batch()
            print(tag, "reference update took", round(time.time() - t0, 1), "s")
        finally:
            torch.randn_like = orig_randn_like
            A.optimizer_value.step = orig_step
        for k, v in log.items():
            arrs[f"log/{k}"] = np.array(float(v[0] if isinstance(v, tuple) else v))
        for k, v in vnorm.items():
            arrs[f"vgnorm/{k}"] = np.array(v)
        for mod, m in A.policy.contextual_modules.items():
            for n, p in m.named_parameters():
                if p.grad is not None:
                    arrs[f"pgnorm/{mod}/{n}"] = np.array(float(p.grad.double().norm()))
        arrs["log_alpha"] = A.log_sac_alpha.detach().clone().numpy()
        cfg = dict(case=c, hp=hp, cls=cls_name, policy_kwargs=pk, value_kwargs=vk, bounds=bounds, n_draws=draws[0],
                   np_seed_run=21, traj_seed=1000, S=S, A=Ad)
        save(f"fullsize_{tag}.npz", cfg=np.array(json.dumps(cfg)), **arrs)


if __name__ == "__main__":
    which = sys.argv[1:] or ["ops", "layers", "steps", "sampler", "updates"]
    if "fullsize" in which or any(w.startswith("fullsize_") for w in which):
        gen_full_size([w[len("fullsize_"):] for w in which if w.startswith("fullsize_")] or None)
    if "ops" in which:
        gen_ops()
    if "layers" in which:
        gen_layers([w[len("layer_"):] for w in which if w.startswith("layer_")] or None)
    if "layers" in which or "ensemble" in which:
        gen_ensemble_layers()
    if "steps" in which:
        gen_steps()
    if "sampler" in which:
        gen_sampler([w[len("sampler_"):] for w in which if w.startswith("sampler_")] or None)
    if "updates" in which or any(w.startswith("update_") for w in which):
        gen_updates([w[len("update_"):] for w in which if w.startswith("update_")] or None)
