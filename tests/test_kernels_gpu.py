"""GPU parity of every C-ABI kernel (called through rorl_b200.kernels -> ctypes -> librorl_b200.so)
against (1) the committed golden fixtures produced by the unmodified reference and (2) the pinned
oracle on the same seeded inputs.  Tolerance: 1e-3 relative fp32 (BASELINE.json), in practice ~1e-5.
"""
import numpy as np
import pytest
import torch

from helpers import T, assert_close, load_npz

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module")
def K():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import rorl_b200.kernels as K
    return K


def G(a, grad=False):
    return T(a, "cuda", grad)


# ------------------------------------------------------------------------------------------------ gilr
def test_gilr_golden(K):
    g = load_npz("ops_gilr.npz")
    v, f = G(g["v"], True), G(g["f"], True)
    h = K.real_scan_tie_input_gate(v, f)
    assert_close(h, g["h"], TOL, "h")
    dv, df = torch.autograd.grad(h, (v, f), G(g["dh"]))
    assert_close(dv, g["dv"], TOL, "dv")
    assert_close(df, g["df"], TOL, "df")


@pytest.mark.parametrize("shape", [(1, 1, 4), (2, 5, 36), (4, 257, 256), (3, 130, 100), (32, 1002, 256)])
@pytest.mark.parametrize("fused", [False, True])
def test_gilr_vs_oracle(K, shape, fused):
    from oracle import ops as O
    B, L, C = shape
    gen = torch.Generator().manual_seed(B * 1000 + L)
    uv, uf = torch.randn(B, L, C, generator=gen), torch.randn(B, L, C, generator=gen) + 1.0
    start = (torch.rand(B, L, 1, generator=gen) < 0.05).float()
    start[:, 0] = 1
    if L > 3:
        start[0, 2:4] = 1      # consecutive resets
    dh = torch.randn(B, L, C, generator=gen)
    # oracle
    a, b = uv.clone().requires_grad_(), uf.clone().requires_grad_()
    if fused:
        v, f = torch.tanh(a), torch.sigmoid(b) * (1 - start)
    else:
        v, f = a, torch.sigmoid(b).detach().requires_grad_()
        b = f
    h_ref, _ = O.gilr_scan(v, f)
    ga_ref, gb_ref = torch.autograd.grad(h_ref, (a, b), dh)
    # kernel
    if fused:
        x, y = uv.cuda().requires_grad_(), uf.cuda().requires_grad_()
        h = K.gilr_fused_scan(x, y, start.cuda())
    else:
        x, y = uv.cuda().requires_grad_(), f.detach().cuda().requires_grad_()
        h = K.real_scan_tie_input_gate(x, y)
    assert_close(h, h_ref, TOL, "h")
    gx, gy = torch.autograd.grad(h, (x, y), dh.cuda())
    assert_close(gx, ga_ref, TOL, "d_v")
    assert_close(gy, gb_ref, TOL, "d_f")


# ------------------------------------------------------------------------------------------------ lru
def test_lru_golden(K):
    g = load_npz("ops_lru.npz")
    vr, vi, fr, fi = (G(g[k], True) for k in ("vr", "vi", "fr", "fi"))
    hr, hi = K.complex_scan(vr, vi, fr, fi, G(g["h0r"]), G(g["h0i"]), None)
    assert_close(hr, g["hr"], TOL, "hr")
    assert_close(hi, g["hi"], TOL, "hi")
    gs = torch.autograd.grad((hr, hi), (vr, vi, fr, fi), (G(g["gr"]), G(g["gi"])))
    for got, k in zip(gs, ("dvr", "dvi", "dfr", "dfi")):
        assert_close(got, g[k], TOL, k)


@pytest.mark.parametrize("shape", [(1, 1, 4), (2, 67, 40), (4, 257, 256), (8, 1002, 256)])
@pytest.mark.parametrize("detach", [False, True])
def test_lru_vs_oracle(K, shape, detach):
    from oracle import ops as O
    B, L, C = shape
    gen = torch.Generator().manual_seed(L)
    mag = 0.9 + 0.099 * torch.rand(C, generator=gen)
    ph = 6.28 * torch.rand(C, generator=gen)
    start = (torch.rand(B, L, 1, generator=gen) < 0.03).float()
    fr = (mag * torch.cos(ph)).expand(B, L, C) * (1 - start)
    fi = (mag * torch.sin(ph)).expand(B, L, C) * (1 - start)
    vr, vi = torch.randn(B, L, C, generator=gen), torch.randn(B, L, C, generator=gen)
    h0r, h0i = torch.randn(B, 1, C, generator=gen), torch.randn(B, 1, C, generator=gen)
    gd = (torch.rand(B, L, 1, generator=gen) < 0.1).float() if detach else None
    gr, gi = torch.randn(B, L, C, generator=gen), torch.randn(B, L, C, generator=gen)
    cpu_in = [t.clone().contiguous().requires_grad_() for t in (vr, vi, fr, fi)]
    hr_ref, hi_ref = O.lru_scan(*cpu_in, h0r, h0i, gd)
    ref = torch.autograd.grad((hr_ref, hi_ref), cpu_in, (gr, gi))
    gpu_in = [t.clone().contiguous().cuda().requires_grad_() for t in (vr, vi, fr, fi)]
    hr, hi = K.complex_scan(*gpu_in, h0r.cuda(), h0i.cuda(), None if gd is None else gd.cuda())
    assert_close(hr, hr_ref, TOL, "hr")
    assert_close(hi, hi_ref, TOL, "hi")
    got = torch.autograd.grad((hr, hi), gpu_in, (gr.cuda(), gi.cuda()))
    for a, b, n in zip(got, ref, ("dvr", "dvi", "dfr", "dfi")):
        assert_close(a, b, TOL, n)


# ------------------------------------------------------------------------------------------------ selective scan
@pytest.mark.parametrize("tag", ["a", "b"])
def test_selscan_golden(K, tag):
    g = load_npz(f"ops_selscan_{tag}.npz")
    names = ("u", "delta", "A", "B", "C", "D", "z", "bias")
    u, delta, A, Bm, Cm, Dk, z, bias = (G(g[k], True) for k in names)
    out, last = K.selective_scan_fn(u, delta, A, Bm, Cm, G(g["start"]), Dk, z=z, delta_bias=bias, delta_softplus=True,
                                    return_last_state=True)
    assert_close(out, g["out"], TOL, "out")
    assert_close(last, g["last"], TOL, "last")
    gs = torch.autograd.grad(out, (u, delta, A, Bm, Cm, Dk, z, bias), G(g["dout"]))
    for got, k in zip(gs, ("du", "ddelta", "dA", "dB", "dC", "dD", "dz", "dbias")):
        assert_close(got, g[k], TOL, k)


@pytest.mark.parametrize("cfg", [
    dict(B=2, D=64, L=130, N=16, z=True, sp=True),
    dict(B=1, D=36, L=33, N=32, z=False, sp=False),
    dict(B=2, D=96, L=70, N=64, z=True, sp=True),
    dict(B=3, D=128, L=16, N=32, z=True, sp=True),
    dict(B=2, D=512, L=1018, N=32, z=True, sp=True),
])
def test_selscan_vs_oracle(K, cfg):
    from oracle import ops as O
    B, D, L, N = cfg["B"], cfg["D"], cfg["L"], cfg["N"]
    gen = torch.Generator().manual_seed(L + N)
    rn = lambda *s: torch.randn(*s, generator=gen)
    u, delta = rn(B, D, L), 0.5 * rn(B, D, L) - 1.0
    z = rn(B, D, L) if cfg["z"] else None
    A = -torch.exp(0.5 * rn(D, N))
    Bm, Cm = rn(B, N, L), rn(B, N, L)
    Dk, bias = rn(D), 0.3 * rn(D)
    if not cfg["sp"]:
        delta = delta.abs() * 0.2
    st = (torch.rand(B, 1, L, generator=gen) < 0.02).float()
    st[:, :, 0] = 1
    if L > 40:
        st[0, :, 31:33] = 1       # reset across the 32-step tile / 16-step chunk boundaries
        st[0, :, 15] = 1
        st[0, :, 16] = 1
    start = st.expand(B, D, L).contiguous()
    dout = rn(B, D, L)
    ins = [u, delta, A, Bm, Cm, Dk, bias] + ([z] if z is not None else [])
    cpu = [t.clone().requires_grad_() for t in ins]
    out_ref, last_ref = O.selective_scan(cpu[0], cpu[1], cpu[2], cpu[3], cpu[4], start, cpu[5],
                                         z=cpu[7] if z is not None else None, delta_bias=cpu[6],
                                         delta_softplus=cfg["sp"], return_last_state=True)
    ref = torch.autograd.grad(out_ref, cpu, dout)
    gpu = [t.clone().cuda().requires_grad_() for t in ins]
    out, last = K.selective_scan_fn(gpu[0], gpu[1], gpu[2], gpu[3], gpu[4], start.cuda(), gpu[5],
                                    z=gpu[7] if z is not None else None, delta_bias=gpu[6],
                                    delta_softplus=cfg["sp"], return_last_state=True)
    assert_close(out, out_ref, TOL, "out")
    assert_close(last, last_ref, TOL, "last_state")
    got = torch.autograd.grad(out, gpu, dout.cuda())
    for a, b, n in zip(got, ref, ["du", "ddelta", "dA", "dB", "dC", "dD", "dbias", "dz"]):
        assert_close(a, b, TOL, n)


def test_selscan_strided_inputs(K):
    """u / z as column slices of one wider projection output, B / C as slices of x_dbl (the layout the
    smamba mixer feeds the kernel)."""
    from oracle import ops as O
    B, L, D, N, R = 2, 50, 64, 32, 16
    gen = torch.Generator().manual_seed(3)
    xz = torch.randn(B, L, 2 * D, generator=gen)
    xdbl = torch.randn(B, L, R + 2 * N, generator=gen)
    delta = torch.randn(B, L, D, generator=gen) * 0.3
    A = -torch.exp(0.3 * torch.randn(D, N, generator=gen))
    start = torch.zeros(B, L, 1)
    start[:, 0] = 1
    start[1, 20] = 1
    ref = O.selective_scan(xz[..., :D].transpose(1, 2), delta.transpose(1, 2), A, xdbl[..., R:R + N].transpose(1, 2),
                           xdbl[..., R + N:].transpose(1, 2), start.transpose(1, 2).expand(B, D, L), None,
                           z=xz[..., D:].transpose(1, 2), delta_softplus=True).transpose(1, 2)
    xzg, xdg = xz.cuda(), xdbl.cuda()
    y = K.selective_scan_tm(xzg[..., :D], delta.cuda(), A.cuda(), xdg[..., R:R + N], xdg[..., R + N:], None,
                            xzg[..., D:], None, start.cuda(), True)
    assert_close(y, ref, TOL, "y")


# ------------------------------------------------------------------------------------------------ conv
@pytest.mark.parametrize("D,L,masked", [(96, 300, True), (97, 300, True), (96, 131, False), (130, 7, True), (512, 1019, True)])
@pytest.mark.parametrize("K_", [2, 4, 8, 16])
def test_conv1d_silu_vs_oracle(K, K_, D, L, masked):
    """D even: the channel-pair (f32x2) kernels; D odd: the one-channel kernels; L below / across / many segments."""
    from oracle import ops as O
    B = 3 if D < 512 else 2
    gen = torch.Generator().manual_seed(K_)
    x = torch.randn(B, L, D, generator=gen)
    w, b = 0.3 * torch.randn(D, 1, K_, generator=gen), 0.1 * torch.randn(D, generator=gen)
    mask = (torch.rand(B, L, 1, generator=gen) > 0.1).float() if masked else torch.ones(B, L, 1)
    dy = torch.randn(B, L, D, generator=gen)
    cx, cw, cb = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    ref = O.causal_conv1d_silu(cx.transpose(1, 2), cw, cb, mask.transpose(1, 2)).transpose(1, 2)
    rg = torch.autograd.grad(ref, (cx, cw, cb), dy)
    gx, gw, gb = x.cuda().requires_grad_(), w.cuda().requires_grad_(), b.cuda().requires_grad_()
    y = K.causal_conv1d_silu(gx, gw, gb, mask.cuda() if masked else None)
    assert_close(y, ref, TOL, "y")
    gg = torch.autograd.grad(y, (gx, gw, gb), dy.cuda())
    for a, r, n in zip(gg, rg, ("dx", "dw", "db")):
        assert_close(a, r, TOL, n)


# ------------------------------------------------------------------------------------------------ add + norm
def test_addnorm_golden(K):
    g = load_npz("ops_addnorm.npz")
    x, r, w, b = (G(g[k], True) for k in ("x", "r", "w", "b"))
    y, res = K.layer_norm_fn(x, w, b, residual=r, eps=1e-8, prenorm=True, residual_in_fp32=True)
    assert_close(y, g["y"], TOL, "y")
    assert torch.equal(res.cpu(), torch.from_numpy(g["res"])), "residual must be the exact fp32 sum"
    gs = torch.autograd.grad((y, res), (x, r, w, b), (G(g["dy"]), G(g["dres"])))
    for got, k in zip(gs, ("dx", "dr", "dw", "db")):
        assert_close(got, g[k], TOL, k)
    y2 = K.rms_norm_fn(x, w, None, residual=r, eps=1e-8, prenorm=False, residual_in_fp32=True)
    assert_close(y2, g["y_rms"], TOL, "y_rms")
    gs = torch.autograd.grad(y2, (x, r, w), G(g["dy"]))
    for got, k in zip(gs, ("dx_rms", "dr_rms", "dw_rms")):
        assert_close(got, g[k], TOL, k)


@pytest.mark.parametrize("C", [16, 256, 512])
@pytest.mark.parametrize("rms", [False, True])
def test_addnorm_vs_oracle(K, C, rms):
    from oracle import ops as O
    rows = 5000
    gen = torch.Generator().manual_seed(C)
    x, r = torch.randn(rows, C, generator=gen), torch.randn(rows, C, generator=gen)
    w, b = torch.randn(C, generator=gen), torch.randn(C, generator=gen)
    dy, dres = torch.randn(rows, C, generator=gen), torch.randn(rows, C, generator=gen)
    cin = [t.clone().requires_grad_() for t in (x, r, w, b)]
    y_ref, res_ref = O.add_norm(cin[0], cin[2], None if rms else cin[3], cin[1], 1e-8, True, rms)
    ref = torch.autograd.grad((y_ref, res_ref), cin[:3] + ([] if rms else [cin[3]]), (dy, dres))
    gin = [t.clone().cuda().requires_grad_() for t in (x, r, w, b)]
    y, res = K.layer_norm_fn(gin[0], gin[2], None if rms else gin[3], residual=gin[1], eps=1e-8, prenorm=True,
                             is_rms_norm=rms)
    assert_close(y, y_ref, TOL, "y")
    got = torch.autograd.grad((y, res), gin[:3] + ([] if rms else [gin[3]]), (dy.cuda(), dres.cuda()))
    for a, c, n in zip(got, ref, ("dx", "dr", "dw", "db")):
        assert_close(a, c, TOL, n)


# ------------------------------------------------------------------------------------------------ gru
@pytest.mark.parametrize("shape", [(3, 41, 16, 16), (5, 17, 12, 32), (2, 9, 64, 64), (6, 50, 128, 128), (4, 33, 256, 256),
                                   (32, 1002, 256, 256), (150, 7, 64, 256)])
@pytest.mark.parametrize("with_h0", [False, True])
def test_gru_vs_torch_cpu(K, shape, with_h0):
    """Persistent cluster GRU (csrc/gru.cu + tensor-core input GEMM) vs torch.nn.GRU on CPU, the reference's own
    `gru` layer (ref: offpolicy_rnn/models/rnn_base.py:59,245-247,454): outputs, final state, input / state /
    parameter gradients."""
    from rorl_b200.models.gru.gru import GRULayer
    B, L, I, H = shape
    if B * L > 20000 and with_h0:
        pytest.skip("one large case is enough")
    torch.manual_seed(B * 100 + L)
    ref = torch.nn.GRU(I, H, batch_first=True)
    mine = GRULayer(I, H, batch_first=True)
    mine.load_state_dict(ref.state_dict())
    mine.cuda()
    x = torch.randn(B, L, I)
    h0 = 0.5 * torch.randn(1, B, H) if with_h0 else None
    dy, dhl = torch.randn(B, L, H), torch.randn(1, B, H)
    xr = x.clone().requires_grad_()
    h0r = None if h0 is None else h0.clone().requires_grad_()
    yr, hr = ref(xr, h0r)
    (yr * dy).sum().backward(retain_graph=True) if False else ((yr * dy).sum() + (hr * dhl).sum()).backward()
    xg = x.cuda().requires_grad_()
    h0g = None if h0 is None else h0.cuda().requires_grad_()
    yg, hg = mine(xg, h0g)
    ((yg * dy.cuda()).sum() + (hg * dhl.cuda()).sum()).backward()
    assert_close(yg, yr, TOL, "out")
    assert_close(hg, hr, TOL, "h_n")
    assert_close(xg.grad, xr.grad, TOL, "dx")
    if with_h0:
        assert_close(h0g.grad, h0r.grad, TOL, "dh0")
    for (n, p), (_, q) in zip(mine.named_parameters(), ref.named_parameters()):
        assert_close(p.grad, q.grad, TOL, n)


def test_selscan_full_size_chunk_composition(K):
    """BASELINE.json's full shape (32 x 1018 x 512, N = 32), checked through a size-independent property: the scan of
    the whole sequence equals the scan of its two halves with the state handed over (h0 in, last_state out), and a
    reset makes everything after it independent of what came before."""
    B, L, D, N = 32, 1018, 512, 32
    g = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
    u, delta, z = rn(B, L, D), 0.5 * rn(B, L, D) - 1, rn(B, L, D)
    Bm, Cm = rn(B, L, N), rn(B, L, N)
    A = -torch.exp(0.3 * rn(D, N))
    Dk, bias = rn(D), 0.3 * rn(D)
    start = torch.zeros(B, L, device="cuda")
    start[:, 0] = 1
    start[3, 500] = 1
    y, last = K.selective_scan_tm(u, delta, A, Bm, Cm, Dk, z, bias, start, True, True)
    k = 509
    sl = lambda t, a, b: t[:, a:b].contiguous()
    y1, l1 = K.selective_scan_tm(sl(u, 0, k), sl(delta, 0, k), A, sl(Bm, 0, k), sl(Cm, 0, k), Dk, sl(z, 0, k), bias, sl(start, 0, k), True, True)
    y2, l2 = K.selective_scan_tm(sl(u, k, L), sl(delta, k, L), A, sl(Bm, k, L), sl(Cm, k, L), Dk, sl(z, k, L), bias, sl(start, k, L), True, True, l1)
    assert_close(torch.cat((y1, y2), dim=1), y, 1e-6, "y")
    assert_close(l2, last, 1e-6, "last_state")
    # reset at t = 500 in row 3: the tail of that row does not depend on the inputs before it
    u2 = u.clone()
    u2[3, :500] = rn(500, D)
    y_alt = K.selective_scan_tm(u2, delta, A, Bm, Cm, Dk, z, bias, start, True)
    assert torch.equal(y_alt[3, 500:], y[3, 500:])
    assert torch.equal(y_alt[:3], y[:3]) and torch.equal(y_alt[4:], y[4:])


@pytest.mark.parametrize("shape", [(3, 41, 16), (2, 130, 64), (32, 1002, 256), (1, 1, 8)])
@pytest.mark.parametrize("with_h0", [False, True])
def test_lru_fused_matches_materialised_path(K, shape, with_h0):
    """The fused LRU scan (lambda, gamma as [C] vectors, reset flags [B, L]) against the same recurrence fed with the
    materialised gamma * u and lambda * (1 - start) tensors through the oracle's CPU scan, incl. d lambda, d gamma."""
    from oracle import ops as O
    B, L, C = shape
    if B * L > 20000 and with_h0:
        pytest.skip("one large case is enough")
    gen = torch.Generator().manual_seed(B * 7 + L)
    rn = lambda *s: torch.randn(*s, generator=gen)
    u_re, u_im = rn(B, L, C), rn(B, L, C)
    mag, theta = 0.9 + 0.099 * torch.rand(C, generator=gen), 6.28 * torch.rand(C, generator=gen)
    lam_re, lam_im, gamma = mag * torch.cos(theta), mag * torch.sin(theta), 0.1 + torch.rand(C, generator=gen)
    start = (torch.rand(B, L, 1, generator=gen) < 0.03).float()
    start[:, 0] = 1
    h0r, h0i = (rn(B, 1, C), rn(B, 1, C)) if with_h0 else (None, None)
    if with_h0:
        start[:, 0] = 0                      # otherwise the carried state is wiped at once
    gr, gi = rn(B, L, C), rn(B, L, C)
    cin = [t.clone().requires_grad_() for t in (u_re, u_im, lam_re, lam_im, gamma)]
    keep = 1 - start
    hr_ref, hi_ref = O.lru_scan(cin[4] * cin[0], cin[4] * cin[1], (cin[2] * keep).expand(B, L, C), (cin[3] * keep).expand(B, L, C), h0r, h0i, None)
    ref = torch.autograd.grad((hr_ref, hi_ref), cin, (gr, gi))
    gin = [t.clone().cuda().requires_grad_() for t in (u_re, u_im, lam_re, lam_im, gamma)]
    hr, hi = K.lru_fused_scan(*gin, start.cuda(), None if h0r is None else h0r.cuda(), None if h0i is None else h0i.cuda())
    assert_close(hr, hr_ref, TOL, "h_re")
    assert_close(hi, hi_ref, TOL, "h_im")
    got = torch.autograd.grad((hr, hi), gin, (gr.cuda(), gi.cuda()))
    for a, b, n in zip(got, ref, ("du_re", "du_im", "dlam_re", "dlam_im", "dgamma")):
        assert_close(a, b, TOL, n)


# ------------------------------------------------------------------------------------------------ tanh-Gaussian head
@pytest.mark.parametrize("M,A", [(1000, 6), (32 * 1019, 6), (77, 1), (300, 17)])
def test_tanh_gaussian_head(K, M, A):
    """One-kernel policy head (csrc/head.cu) against the elementwise torch graph of the reference's process_model_out
    (ref: contextual_sac_policy_single_head.py:105-123) in float64: outputs and the gradient w.r.t. the head output,
    with logstd values outside the clamp range (zero gradient there) and large |sample| (softplus tails)."""
    import numpy as np
    gen = torch.Generator(device="cuda").manual_seed(A)
    out = torch.randn(M, 2 * A, device="cuda", generator=gen) * 3.0
    out[::7, 0] = 5.0                      # above MAX_LOG_STD
    out[1::7, 0] = -25.0                   # below MIN_LOG_STD
    out.requires_grad_()
    noise = torch.randn(M, A, device="cuda", generator=gen)
    am, asamp, lp = K.tanh_gaussian_head(out, noise, -20.0, 2.0)
    d_am, d_as, d_lp = (torch.randn_like(t) for t in (am, asamp, lp))
    got = torch.autograd.grad((am, asamp, lp), out, (d_am, d_as, d_lp))[0]
    o64 = out.detach().double().requires_grad_()
    logstd, logit = o64.chunk(2, dim=-1)
    logstd = torch.clamp(logstd, -20.0, 2.0)
    sample = logit + noise.double() * logstd.exp()
    ref_lp = (-0.5 * noise.double().pow(2) - (logstd + 0.5 * np.log(2 * np.pi))).sum(-1, keepdim=True)
    ref_lp = ref_lp - (2 * (-sample - torch.nn.functional.softplus(-2 * sample) + np.log(2))).sum(-1, keepdim=True)
    ref = torch.autograd.grad((torch.tanh(logit), torch.tanh(sample), ref_lp), o64, (d_am.double(), d_as.double(), d_lp.double()))[0]
    assert_close(am, torch.tanh(logit), 1e-6, "action_mean")
    assert_close(asamp, torch.tanh(sample), 1e-6, "action_sample")
    assert_close(lp, ref_lp, 1e-6, "log_prob")
    assert_close(got, ref, 1e-5, "d_out")
    # only the log-prob gradient (the alpha / entropy path)
    got2 = torch.autograd.grad(K.tanh_gaussian_head(out, noise, -20.0, 2.0)[2], out, d_lp)[0]
    ref2 = torch.autograd.grad(ref_lp_of(o64, noise), o64, d_lp.double())[0]
    assert_close(got2, ref2, 1e-5, "d_out (log_prob only)")


def ref_lp_of(o64, noise):
    import numpy as np
    logstd, logit = o64.chunk(2, dim=-1)
    logstd = torch.clamp(logstd, -20.0, 2.0)
    sample = logit + noise.double() * logstd.exp()
    lp = (-0.5 * noise.double().pow(2) - (logstd + 0.5 * np.log(2 * np.pi))).sum(-1, keepdim=True)
    return lp - (2 * (-sample - torch.nn.functional.softplus(-2 * sample) + np.log(2))).sum(-1, keepdim=True)


# ------------------------------------------------------------------------------------------------ fused SSM core
@pytest.mark.parametrize("B,L,Dn,R,Ns", [(3, 203, 128, 4, 16), (4, 1019, 512, 16, 32), (2, 77, 64, 8, 64)])
def test_ssm_core_matches_unfused_graph(K, B, L, Dn, R, Ns):
    """x_proj + dt_proj + A = -exp(A_log) + scan as one autograd node against the same computation composed of the
    separately tested pieces (tensor-core linear, narrow linear, selective_scan_tm on column slices, torch exp): output
    and every gradient (xs, x_proj / dt_proj weights, A_log, D, z, dt bias)."""
    gen = torch.Generator(device="cuda").manual_seed(B * L)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=gen)
    xs, z = rn(B, L, Dn), rn(B, L, Dn)
    Wx, Wdt = 0.1 * rn(R + 2 * Ns, Dn), 0.3 * rn(Dn, R)
    A_log = torch.log(torch.arange(1, Ns + 1, device="cuda", dtype=torch.float32).repeat(Dn, 1)) + 0.05 * rn(Dn, Ns)
    Dk, bias = rn(Dn), 0.5 * rn(Dn)
    start = torch.zeros(B, L, 1, device="cuda")
    start[:, 0] = 1
    start[1, L // 2] = 1
    dy = rn(B, L, Dn)
    leaves = [t.clone().requires_grad_() for t in (xs, Wx, Wdt, A_log, Dk, z, bias)]
    assert K.ssm_core_ok(leaves[0], leaves[5], leaves[1], leaves[2], leaves[3])
    y = K.ssm_core(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], leaves[5], leaves[6], start)
    got = torch.autograd.grad(y, leaves, dy)
    ref_leaves = [t.clone().requires_grad_() for t in (xs, Wx, Wdt, A_log, Dk, z, bias)]
    x_, Wx_, Wdt_, A_, D_, z_, b_ = ref_leaves
    x_dbl = K.linear(x_, Wx_)
    delta = K.linear(x_dbl[..., :R], Wdt_)
    y_ref = K.selective_scan_tm(x_, delta, -torch.exp(A_), x_dbl[..., R:R + Ns], x_dbl[..., R + Ns:], D_, z_, b_, start, True)
    ref = torch.autograd.grad(y_ref, ref_leaves, dy)
    assert_close(y, y_ref, 1e-5, "y")
    for a, r, n in zip(got, ref, ("dxs", "dWx", "dWdt", "dA_log", "dD", "dz", "dbias")):
        assert_close(a, r, 2e-5, n)
