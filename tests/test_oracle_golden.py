"""Pin the oracle (oracle/*.py) against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only.  Tolerances: sampler bit-exact; fp32 arithmetic 1e-5
(oracle and reference run the same PyTorch CPU ops in slightly different association orders)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import T, assert_close, cfg_of, load_npz, nested_sd
from oracle import model as OM
from oracle import ops as O
from oracle import sampler as OS
from oracle import update as OU

TOL = 2e-5


def test_gilr_scan_op():
    g = load_npz("ops_gilr.npz")
    v, f = T(g["v"], grad=True), T(g["f"], grad=True)
    h, _ = O.gilr_scan(v, f)
    assert_close(h, g["h"], TOL, "h")
    dv, df = torch.autograd.grad(h, (v, f), T(g["dh"]))
    assert_close(dv, g["dv"], TOL, "dv")
    assert_close(df, g["df"], TOL, "df")


def test_lru_scan_op():
    g = load_npz("ops_lru.npz")
    vr, vi, fr, fi = (T(g[k], grad=True) for k in ("vr", "vi", "fr", "fi"))
    hr, hi = O.lru_scan(vr, vi, fr, fi, T(g["h0r"]), T(g["h0i"]))
    assert_close(hr, g["hr"], TOL, "hr")
    assert_close(hi, g["hi"], TOL, "hi")
    gs = torch.autograd.grad((hr, hi), (vr, vi, fr, fi), (T(g["gr"]), T(g["gi"])))
    for got, k in zip(gs, ("dvr", "dvi", "dfr", "dfi")):
        assert_close(got, g[k], TOL, k)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_selective_scan_op(tag):
    g = load_npz(f"ops_selscan_{tag}.npz")
    names = ("u", "delta", "A", "B", "C", "D", "z", "bias")
    u, delta, A, Bm, Cm, Dk, z, bias = (T(g[k], grad=True) for k in names)
    out, last = O.selective_scan(u, delta, A, Bm, Cm, T(g["start"]), Dk, z=z, delta_bias=bias, delta_softplus=True,
                                 return_last_state=True)
    assert_close(out, g["out"], TOL, "out")
    assert_close(last, g["last"], TOL, "last")
    gs = torch.autograd.grad(out, (u, delta, A, Bm, Cm, Dk, z, bias), T(g["dout"]))
    for got, k in zip(gs, ("du", "ddelta", "dA", "dB", "dC", "dD", "dz", "dbias")):
        assert_close(got, g[k], 1e-4, k)


def test_addnorm_op():
    g = load_npz("ops_addnorm.npz")
    x, r, w, b = (T(g[k], grad=True) for k in ("x", "r", "w", "b"))
    y, res = O.add_norm(x, w, b, r, eps=1e-8, prenorm=True)
    assert_close(y, g["y"], TOL, "y")
    assert_close(res, g["res"], TOL, "res")
    gs = torch.autograd.grad((y, res), (x, r, w, b), (T(g["dy"]), T(g["dres"])))
    for got, k in zip(gs, ("dx", "dr", "dw", "db")):
        assert_close(got, g[k], TOL, k)
    y2 = O.add_norm(x, w, None, r, eps=1e-8, prenorm=False, is_rms=True)
    assert_close(y2, g["y_rms"], TOL, "y_rms")
    gs = torch.autograd.grad(y2, (x, r, w), T(g["dy"]))
    for got, k in zip(gs, ("dx_rms", "dr_rms", "dw_rms")):
        assert_close(got, g[k], TOL, k)


@pytest.mark.parametrize("tag", ["gilr", "lru", "gru", "smamba_rms", "smamba_ln", "smamba_ff", "mamba_ff", "mamba_noff", "mamba_h0",
                                 "gilr_lstm", "gilr_lstm_h0", "conv1d", "conv1d_h0"])
def test_encoder_layer(tag):
    g = load_npz(f"layer_{tag}.npz")
    lid = str(g["layer_id"])
    p = {k[2:]: T(v, grad=True) for k, v in g.items() if k.startswith("p/")}
    x = T(g["x"], grad=True)
    side = OM.Side() if lid == "gru" else OM.Side(T(g["start"]), T(g["mask"]))
    if "h_in" in g:
        side.h0 = {1: T(g["h_in"])}
    y = OM.rnn_base(p, ['fc', lid, 'fc'], ['elu', 'elu', 'linear'], x, side)
    assert_close(y, g["y"], TOL, "y")
    if "h_out" in g:        # s6 layer: [ssm state | conv window], batch-first (ref: s6/mamba.py:187-190)
        assert side.h_out.shape == g["h_out"].shape
        assert_close(side.h_out, g["h_out"], TOL, "h_out")
    names = [k[2:] for k in g if k.startswith("g/")]
    gs = torch.autograd.grad(y, [x] + [p[n] for n in names], T(g["dy"]))
    assert_close(gs[0], g["dx"], 1e-4, "dx")
    for n, got in zip(names, gs[1:]):
        assert_close(got, g["g/" + n], 2e-4, n)


@pytest.mark.parametrize("tag", ["smamba_rms", "smamba_ln16"])
def test_smamba_rollout_step(tag):
    """Oracle restatement of the L == 1 rollout path against the reference's own step loop (carried hidden)."""
    g = load_npz(f"step_{tag}.npz")
    lid = str(g["layer_id"])
    p = {k[2:]: T(v) for k, v in g.items() if k.startswith("p/")}
    x, h = T(g["x"]), T(g["h_in"])
    ys = []
    for t in range(x.shape[1]):
        e = F.elu(F.linear(x[:, t:t + 1], p['layer_list.0.weight'], p['layer_list.0.bias']))
        e, h = OM.smamba_step(p, 'layer_list.1.', e, h, lid)
        ys.append(F.linear(F.elu(e), p['layer_list.2.weight'], p['layer_list.2.bias']))
    assert_close(torch.cat(ys, dim=1), g["y"], TOL, "y")
    assert_close(h, g["h_out"], TOL, "h_out")


@pytest.mark.parametrize("tag", ["cgpt_ln", "cgpt_rms"])
def test_cgpt_layer_oracle_vs_reference_on_flash_attn(tag):
    """Pins oracle/attention.py (fp32 restatement of the cgpt decoder) against the UNMODIFIED reference run with
    flash-attn 2.8.3 in bf16 autocast on a B200 (tests/golden/make_golden_gpu.py).  The reference's attention region
    computes in bf16, so the two agree to bf16 rounding: 1e-2 (BASELINE.json's tolerance for this path)."""
    g = load_npz(f"layer_{tag}.npz")
    lid = str(g["layer_id"])
    p = {k[2:]: T(v, grad=True) for k, v in g.items() if k.startswith("p/")}
    x = T(g["x"], grad=True)
    side = OM.Side(attention_concat_mask=torch.from_numpy(g["seqlens"]).to(torch.int))
    y = OM.rnn_base(p, ['fc', lid, 'fc'], ['elu', 'elu', 'linear'], x, side)
    assert_close(y, g["y"], 1e-2, "y")
    names = [k[2:] for k in g if k.startswith("g/")]
    gs = torch.autograd.grad(y, [x] + [p[n] for n in names], T(g["dy"]))
    assert_close(gs[0], g["dx"], 1e-2, "dx")
    for n, got in zip(names, gs[1:]):
        assert_close(got, g["g/" + n], 2e-2, n)


def test_cgpt_update_oracle_vs_reference_on_flash_attn():
    """The oracle's full SAC update with a cgpt encoder against the reference's train_one_batch on a B200 with flash-attn
    and GradScaler (update_sac_cgpt.npz): logged scalars and gradients at the bf16 tolerance."""
    g, cfg, upd = run_oracle_update("sac_cgpt")
    log = upd.train_one_batch()
    for k in ("critic_loss", "actor_loss", "log_prob", "target_q_max", "q1_l2_norm_square"):
        ref = float(g[f"c0/log/{k}"])
        assert abs(log[k] - ref) <= 1e-2 * max(1.0, abs(ref)), (k, log[k], ref)
    for k, v in g.items():
        if k.startswith("c0/vgrad/"):
            mod, name = k[len("c0/vgrad/"):].split("/", 1)
            assert_close(upd.value_grads[mod][name], v, 2e-2, k)
        if k.startswith("c0/pgrad/"):
            mod, name = k[len("c0/pgrad/"):].split("/", 1)
            assert_close(upd.policy_grads[mod][name], v, 2e-2, k)


def _fill_discrete(buf, rng, lens, S, A):
    for Tn in lens:
        last_s, last_a, last_r = np.zeros((1, S)), np.zeros((1, A)), np.zeros((1, 1))
        s = rng.standard_normal((1, S))
        for t in range(Tn):
            ai = int(rng.randint(A))
            ns = rng.standard_normal((1, S))
            r = float(rng.standard_normal())
            done = t == Tn - 1
            buf.mem_push(OS.Transition(state=s, last_state=last_s, last_action=last_a, action=np.array([[float(ai)]]), next_state=ns, reward=r,
                                       logp=None, mask=1, done=done, timeout=done, start=(t == 0), reward_input=last_r))
            last_a = np.zeros((1, A))
            last_a[0, ai] = 1.0
            last_s, last_r, s = s, np.array([[r]]), ns


def _fill(buf, rng, lens, S, A):
    for Tn in lens:
        last_s, last_a, last_r = np.zeros((1, S)), np.zeros((1, A)), np.zeros((1, 1))
        s = rng.standard_normal((1, S))
        for t in range(Tn):
            a = np.tanh(rng.standard_normal((1, A)))
            ns = rng.standard_normal((1, S))
            r = float(rng.standard_normal())
            done = t == Tn - 1
            buf.mem_push(OS.Transition(state=s, last_state=last_s, last_action=last_a, action=a, next_state=ns, reward=r,
                                       logp=None, mask=1, done=done, timeout=done, start=(t == 0), reward_input=last_r))
            last_s, last_a, last_r, s = s, a, np.array([[r]]), ns


@pytest.mark.parametrize("tag", ["a", "b", "c", "d", "e", "f", "g", "h", "i"])
def test_sampler_bit_exact(tag):
    g = load_npz(f"sampler_{tag}.npz")
    c = cfg_of(g)
    buf = OS.RefNestedReplay(500, c["max_step"], additional_history_len=c["skip_extra"])
    _fill(buf, np.random.RandomState(3), c["lens"], c["S"], c["A"])
    np.random.seed(11)
    for call in range(2):
        tr, total, valid, lens = buf.sample_trajs(c["batch"], nest_stack_trajs=c["nest"], randomize_mask=c.get("randomize_mask", False),
                                                  valid_number_post_randomized=c.get("valid_num", 0),
                                                  equalize_data_of_each_traj=c.get("equalize", True),
                                                  random_trunc_traj=c.get("random_trunc", False))
        for n in OS.FIELDS:
            v = getattr(tr, n)
            key = f"c{call}/{n}"
            if v is None:
                assert key not in g
                continue
            assert np.array_equal(np.asarray(v), g[key]), n
        assert np.array_equal(valid, g[f"c{call}/valid"])
        assert np.array_equal(lens, g[f"c{call}/lens"])
        assert total == int(g[f"c{call}/total"])
        st = np.random.get_state()
        assert np.array_equal(np.array(st[1][:4], dtype=np.int64), g[f"c{call}/rng_next"])
        assert st[2] == int(g[f"c{call}/rng_pos"])


def run_oracle_update(tag):
    g = load_npz(f"update_{tag}.npz")
    cfg = cfg_of(g)
    c, hp = cfg["case"], cfg["hp"]
    pol = nested_sd(g, "init/policy/")
    val = nested_sd(g, "init/value/")
    buf = OS.RefNestedReplay(1000, c.get("max_len", max(c["lens"])), additional_history_len=cfg["skip"])   # ref: sac_full_length_rnn_ensembleQ.py:41
    (_fill_discrete if c.get("discrete") else _fill)(buf, np.random.RandomState(cfg["np_seed_fill"]), c["lens"], c["S"], c["A"])
    noises = [T(g[f"noise/{i}"]) for i in range(cfg["n_noise"])]
    it = iter(noises)
    hit = iter([T(g[f"hdraw/{i}"]) for i in range(cfg.get("n_hdraw", 0))])

    def hidden_fn(spec, batch):          # the reference's torch.rand draws, one per recurrent layer of the embedding network
        return {i: next(hit) * 2 - 1 for i, t in enumerate(spec.emb_types) if OM.is_rnn(t)}

    def noise_fn(shape):
        n = next(it)
        assert tuple(n.shape) == tuple(shape), (n.shape, shape)
        return n

    pk = dict(cfg["policy_kwargs"])
    vk = dict(cfg["value_kwargs"])
    cls = cfg.get("cls", "REDQ_SEP_OPTIM")
    upd = OU.RefUpdate(pol, val, OM.ModelSpec(**pk), OM.ModelSpec(**vk), hp, buf, noise_fn, algo=c["algo"], redq="REDQ" in cls,
                       allow_nest_stack=cfg["allow_nest_stack"], sep_optim=cls.endswith("SEP_OPTIM"), hidden_fn=hidden_fn,
                       discrete=bool(c.get("discrete", False)))
    np.random.seed(cfg["np_seed_run"])
    return g, cfg, upd


UPDATE_TAGS = ["sac_smamba", "sac_gru", "td3_gilr", "td3_lru", "sac_ensembleq", "td3_ensembleq", "sac_ensembleq_sep", "sac_smamba_mid",
               "td3_gilr_mid", "sac_conv1d", "sac_gru_clipnorm", "sac_smamba_clipval", "sac_gru_utd2", "sac_gru_rndhidden", "sac_discrete"]


@pytest.mark.parametrize("tag", UPDATE_TAGS)
def test_full_update(tag):
    g, cfg, upd = run_oracle_update(tag)
    for call in range(cfg["case"]["calls"]):
        log = upd.train_one_batch()
        for k in ("critic_loss", "actor_loss", "alpha_loss", "log_prob", "log_alpha", "target_q_max", "clip_min", "clip_max",
                  "q1_l2_norm_square", "policy_l2_norm_square", "value_grad_norm", "policy_grad_norm"):
            key = f"c{call}/log/{k}"
            if key in g and k in log:
                ref = float(g[key])
                assert abs(log[k] - ref) <= 1e-4 * max(1.0, abs(ref)), (k, log[k], ref)
        assert log["real_batch_size"] == int(g[f"c{call}/log/real_batch_size"])
        for k, v in g.items():
            if k.startswith(f"c{call}/vgrad/"):
                mod, name = k[len(f"c{call}/vgrad/"):].split("/", 1)
                assert_close(upd.value_grads[mod][name], v, 5e-4, k)
            if k.startswith(f"c{call}/pgrad/"):
                mod, name = k[len(f"c{call}/pgrad/"):].split("/", 1)
                assert_close(upd.policy_grads[mod][name], v, 5e-4, k)
        for which, sd in (("policy", upd.policy), ("value", upd.value), ("target", upd.target)):
            for k, v in g.items():
                pre = f"c{call}/{which}/"
                if k.startswith(pre):
                    mod, name = k[len(pre):].split("/", 1)
                    assert_close(sd[mod][name], v, 1e-4, k)
        assert abs(upd.log_alpha.item() - float(g[f"c{call}/log_alpha"][0])) < 1e-6
