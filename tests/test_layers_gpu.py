"""Encoder layers on the CUDA path against the layer-level golden fixtures recorded from the UNMODIFIED reference
(tests/golden/layer_*.npz: RNNBase(['fc', <ID>, 'fc']) forward, input and parameter gradients, and for the s6
`mamba_*` layer the returned hidden state incl. a carried non-zero state).  fp32 tolerance 1e-3 (BASELINE.json)."""
import pytest
import torch

from helpers import T, assert_close, load_npz

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.mark.parametrize("tag", ["gilr", "lru", "gru", "smamba_rms", "smamba_ln", "smamba_ff", "mamba_ff", "mamba_noff", "mamba_h0",
                                 "gilr_lstm", "gilr_lstm_h0", "conv1d", "conv1d_h0"])
def test_layer_golden(tag):
    from rorl_b200.models.rnn_base import RNNBase
    g = load_npz(f"layer_{tag}.npz")
    lid = str(g["layer_id"])
    net = RNNBase(12, 8, [16, 16], ['elu', 'elu', 'linear'], ['fc', lid, 'fc'])
    sd = {k[2:]: T(v) for k, v in g.items() if k.startswith("p/")}
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    net.cuda()
    x = T(g["x"], "cuda", grad=True)
    hid = net.make_init_state(x.shape[0], x.device)
    if lid != "gru":
        hid.set_rnn_start(T(g["start"], "cuda"))
        hid.set_mask(T(g["mask"], "cuda"))
    if tag.endswith("_h0"):
        hid[0] = T(g["h_in"], "cuda")
    y, h_out, _ = net.meta_forward(x, hid)
    assert_close(y, g["y"], TOL, "y")
    if "h_out" in g:
        assert tuple(h_out[0].shape) == g["h_out"].shape
        assert_close(h_out[0], g["h_out"], TOL, "h_out")
    params = dict(net.named_parameters())
    names = [k[2:] for k in g if k.startswith("g/")]
    gs = torch.autograd.grad(y, [x] + [params[n] for n in names], T(g["dy"], "cuda"))
    assert_close(gs[0], g["dx"], TOL, "dx")
    for n, got in zip(names, gs[1:]):
        assert_close(got, g["g/" + n], TOL, n)


@pytest.mark.parametrize("tag", ["elru", "econv1d", "egilr_lstm", "elru_h0", "egilr"])
def test_ensemble_layer_golden(tag):
    """Ensemble encoder IDs (`elru-E`, `egilr-E`, `egilr_lstm-E`, `econv1d_K-E`: one independent encoder per member,
    output [E, B, L, C]) against the UNMODIFIED reference: RNNBase(['fc', ID, 'efc-3']) forward, returned hidden, input
    and parameter gradients (tests/golden/layer_e*.npz; `egilr` comes from the reference's Triton path on the GPU box)."""
    import os
    from helpers import GOLDEN
    from rorl_b200.models.rnn_base import RNNBase
    if not os.path.exists(os.path.join(GOLDEN, f"layer_{tag}.npz")):
        pytest.skip(f"layer_{tag}.npz is generated on the GPU box (tests/golden/make_golden_gpu.py)")
    g = load_npz(f"layer_{tag}.npz")
    lid = str(g["layer_id"])
    width = g["p/layer_list.0.weight"].shape[0]
    net = RNNBase(12, 2, [width, width], ['elu', 'elu', 'linear'], ['fc', lid, 'efc-3'])
    for l in net.layer_list:
        if hasattr(l, 'desire_ndim'):
            l.desire_ndim = 4
    missing, unexpected = net.load_state_dict({k[2:]: T(v) for k, v in g.items() if k.startswith("p/")}, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    net.cuda()
    x = T(g["x"], "cuda", grad=True)
    hid = net.make_init_state(x.shape[0], x.device)
    hid.set_rnn_start(T(g["start"], "cuda"))
    hid.set_mask(T(g["mask"], "cuda"))
    if "h_in" in g:
        hid[0] = T(g["h_in"], "cuda")
    y, h_out, _ = net.meta_forward(x, hid)
    assert tuple(y.shape) == g["y"].shape
    assert_close(y, g["y"], TOL, "y")
    assert tuple(h_out[0].shape) == g["h_out"].shape
    assert_close(h_out[0], g["h_out"], TOL, "h_out")
    params = dict(net.named_parameters())
    names = [k[2:] for k in g if k.startswith("g/")]
    gs = torch.autograd.grad(y, [x] + [params[n] for n in names], T(g["dy"], "cuda"))
    assert_close(gs[0], g["dx"], TOL, "dx")
    for n, got in zip(names, gs[1:]):
        assert_close(got, g["g/" + n], TOL, n)


@pytest.mark.parametrize("tag", ["smamba_rms", "smamba_ln16"])
def test_smamba_rollout_step_golden(tag):
    """The L == 1 rollout path (conv window roll + one step of the scan kernel with the carried SSM state) against the
    reference's own Mamba.step loop on CPU (tests/golden/step_*.npz)."""
    from rorl_b200.models.rnn_base import RNNBase
    g = load_npz(f"step_{tag}.npz")
    lid = str(g["layer_id"])
    net = RNNBase(12, 8, [16, 16], ['elu', 'elu', 'linear'], ['fc', lid, 'fc'])
    net.load_state_dict({k[2:]: T(v) for k, v in g.items() if k.startswith("p/")})
    net.cuda()
    x = T(g["x"], "cuda")
    h = net.make_init_state(x.shape[0], x.device)
    h[0] = T(g["h_in"], "cuda")
    ys = []
    with torch.no_grad():
        for t in range(x.shape[1]):
            y, h, _ = net.meta_forward(x[:, t:t + 1], h)
            ys.append(y)
    assert_close(torch.cat(ys, dim=1), g["y"], TOL, "y")
    assert tuple(h[0].shape) == g["h_out"].shape
    assert_close(h[0], g["h_out"], TOL, "h_out")


@pytest.mark.parametrize("tag", ["gilr", "lru", "gru", "gilr_lstm", "conv1d"])
def test_linear_rollout_step_golden(tag):
    """L == 1 rollout calls with the hidden carried between them (and a mid-rollout reset) against the UNMODIFIED
    reference's CPU path for the same loop (tests/golden/step_{gilr,lru,gru,gilr_lstm,conv1d}.npz).  SURVEY.md 8 f2."""
    from rorl_b200.models.rnn_base import RNNBase
    g = load_npz(f"step_{tag}.npz")
    lid = str(g["layer_id"])
    net = RNNBase(12, 8, [16, 16], ['elu', 'elu', 'linear'], ['fc', lid, 'fc'])
    net.load_state_dict({k[2:]: T(v) for k, v in g.items() if k.startswith("p/")})
    net.cuda()
    x, start = T(g["x"], "cuda"), T(g["start"], "cuda")
    h = net.make_init_state(x.shape[0], x.device)
    h[0] = T(g["h_in"], "cuda")
    ys = []
    with torch.no_grad():
        for t in range(x.shape[1]):
            if lid != "gru":
                h.set_rnn_start(start[:, t:t + 1])
                h.set_mask(torch.ones(x.shape[0], 1, 1, device="cuda"))
            y, h, _ = net.meta_forward(x[:, t:t + 1], h)
            ys.append(y)
    assert_close(torch.cat(ys, dim=1), g["y"], TOL, "y")
    assert tuple(h[0].shape) == g["h_out"].shape
    assert_close(h[0], g["h_out"], TOL, "h_out")


@pytest.mark.parametrize("lid,width", [("gilr", 64), ("lru", 64), ("gru", 64), ("mamba_s16_c4", 64), ("mamba_s32_c16_noff", 32),
                                       ("gilr_lstm", 64), ("conv1d_8", 64)])
def test_carried_state_composition(lid, width):
    """Size-independent property of every layer that carries its state: running a sequence in two pieces, handing the
    hidden state over, gives what the single pass gives (outputs and final state).  Exercises the carried-state
    paths the update itself never takes (h0 of the scan kernels, the s6 conv window, the gilr correction term)."""
    from rorl_b200.models.rnn_base import RNNBase
    torch.manual_seed(5)
    net = RNNBase(24, 16, [width, width], ['elu', 'elu', 'linear'], ['fc', lid, 'fc']).cuda()
    B, L, k = 6, 203, 77
    x = torch.randn(B, L, 24, device="cuda")
    start = torch.zeros(B, L, 1, device="cuda")
    start[:, 0] = 1
    start[1, 120] = 1
    start[2, 40] = 1

    def run(xs, st, hid):
        if lid != "gru":
            hid.set_rnn_start(st)
        with torch.no_grad():
            y, h, _ = net.meta_forward(xs, hid)
        return y, h

    y_full, h_full = run(x, start, net.make_init_state(B, x.device))
    y1, h1 = run(x[:, :k], start[:, :k], net.make_init_state(B, x.device))
    carry = net.make_init_state(B, x.device)
    hv = h1[0]
    carry[0] = hv if hv.shape[0] == 1 else hv.transpose(0, 1)          # the s6 layer returns its hidden batch-first
    y2, h2 = run(x[:, k:], start[:, k:], carry)
    assert_close(torch.cat((y1, y2), dim=1), y_full, 1e-4, "y")
    assert_close(h2[0].reshape(-1), h_full[0].reshape(-1), 1e-4, "final hidden")
