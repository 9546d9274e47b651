"""Deterministic reduction kernels (csrc/reduce.cu) against float64 sums: bias-gradient column sums, the fused
ELU-backward + column sum, partial-stack sums and the narrow-input weight gradient."""
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    import rorl_b200.kernels as K
    return K


@pytest.mark.parametrize("M,N", [(1, 4), (63, 12), (1000, 80), (32576, 256), (32576, 128), (5000, 1028), (70000, 2052)])
def test_colsum(K, M, N):
    g = torch.Generator(device="cuda").manual_seed(M + N)
    x = torch.randn(M, N, device="cuda", generator=g)
    assert rel_err(K.colsum(x), x.double().sum(0)) < 2e-6
    wide = torch.randn(M, 2 * N, device="cuda", generator=g)
    sl = wide[:, N:]                                   # row-strided view
    assert rel_err(K.colsum(sl), sl.double().sum(0)) < 2e-6
    a, b = K.colsum(x), K.colsum(x)
    assert torch.equal(a, b), "must be deterministic"


def test_colsum_grouped_and_leading(K):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(8, 3001, 256, device="cuda", generator=g)
    assert rel_err(K.colsum(x), x.double().sum(1)) < 2e-6
    parts = torch.randn(8, 32, 1018, 64, device="cuda", generator=g)
    assert rel_err(K.sum_leading(parts), parts.double().sum(0)) < 1e-6
    assert K.sum_leading(parts[:1]).data_ptr() == parts.data_ptr()
    odd = torch.randn(5, 7, 3, device="cuda", generator=g)          # not a multiple of 4: ATen path
    assert rel_err(K.sum_leading(odd), odd.double().sum(0)) < 1e-6


@pytest.mark.parametrize("shape", [(777, 64), (8, 2000, 256), (32576, 256)])
def test_elu_bwd_colsum(K, shape):
    g = torch.Generator(device="cuda").manual_seed(shape[-1])
    pre = torch.randn(*shape, device="cuda", generator=g)
    y = torch.nn.functional.elu(pre)
    dy = torch.randn(*shape, device="cuda", generator=g)
    gout, db = K.elu_bwd_colsum(dy, y)
    ref = dy.double() * torch.where(pre > 0, torch.ones_like(pre), pre.exp()).double()
    assert rel_err(gout, ref) < 1e-6
    assert rel_err(db, ref.sum(-2)) < 2e-6


@pytest.mark.parametrize("M,N,Kk", [(32576, 128, 9), (32576, 128, 6), (32576, 512, 16), (5000, 300, 1), (100, 8, 3), (2049, 1030, 13)])
def test_skinny_wgrad(K, M, N, Kk):
    gen = torch.Generator(device="cuda").manual_seed(M + Kk)
    g = torch.randn(M, N, device="cuda", generator=gen)
    big = torch.randn(M, 45, device="cuda", generator=gen)
    x = big[:, 7:7 + Kk]                               # column slice of a wider batch row, like the replay fields
    dW = K.skinny_wgrad(g, x)
    assert rel_err(dW, g.double().t() @ x.double()) < 2e-6


def test_linear_skinny_autograd(K):
    gen = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(32, 1018, 9, device="cuda", generator=gen, requires_grad=True)
    W = torch.randn(128, 9, device="cuda", generator=gen, requires_grad=True)
    b = torch.randn(128, device="cuda", generator=gen, requires_grad=True)
    dy = torch.randn(32, 1018, 128, device="cuda", generator=gen)
    y = K.linear(x, W, b)
    assert rel_err(y, torch.nn.functional.linear(x.detach().double(), W.detach().double(), b.detach().double())) < 1e-6
    sl = torch.randn(32, 1018, 80, device="cuda", generator=gen)[..., :16]          # dt_proj: column slice, no bias
    W2 = torch.randn(512, 16, device="cuda", generator=gen)
    assert rel_err(K.linear(sl, W2), sl.double() @ W2.double().t()) < 1e-6
    got = torch.autograd.grad(y, (x, W, b), dy)
    xd, Wd, bd = (t.detach().double().requires_grad_() for t in (x, W, b))
    ref = torch.autograd.grad(torch.nn.functional.linear(xd, Wd, bd), (xd, Wd, bd), dy.double())
    for a, r, n in zip(got, ref, "x W b".split()):
        assert rel_err(a, r) < 1e-5, n


@pytest.mark.parametrize("elu,M", [(False, 32 * 1019), (True, 32 * 1019), (True, 77)])
def test_skinny_encoders_write_one_buffer(elu, M):
    """cat_i(x_i W_i^T + b_i) [+ ELU] with every projection writing its column block of one buffer, against torch."""
    import rorl_b200.kernels as K
    gen = torch.Generator(device="cuda").manual_seed(3)
    dims = (9, 9, 6)
    xs = [torch.randn(M, k, device="cuda", generator=gen, requires_grad=(i == 2)) for i, k in enumerate(dims)]
    Ws = [torch.randn(128, k, device="cuda", generator=gen, requires_grad=True) for k in dims]
    bs = [torch.randn(128, device="cuda", generator=gen, requires_grad=True) for _ in dims]
    assert K.skinny_encoders_ok(xs, Ws)
    y = K.skinny_encoders(xs, Ws, bs, elu=elu)
    dy = torch.randn(M, 384, device="cuda", generator=gen)
    got = torch.autograd.grad(y, [xs[2]] + Ws + bs, dy)
    xd = [x.detach().double().requires_grad_(i == 2) for i, x in enumerate(xs)]
    Wd = [w.detach().double().requires_grad_() for w in Ws]
    bd = [b.detach().double().requires_grad_() for b in bs]
    yr = torch.cat([torch.nn.functional.linear(a, w, b) for a, w, b in zip(xd, Wd, bd)], dim=-1)
    if elu:
        yr = torch.nn.functional.elu(yr)
    ref = torch.autograd.grad(yr, [xd[2]] + Wd + bd, dy.double())
    assert rel_err(y, yr) < 1e-5
    for a, r in zip(got, ref):
        assert a.shape == r.shape
        assert rel_err(a, r) < 1e-4


@pytest.mark.parametrize("elu", [False, True])
def test_skinny_encoders_read_batch_slices_in_place(elu):
    """The projections' inputs are column AND time slices of one [B, L + 1, C] batch tensor (rows not mergeable into one
    stride): read where they lie (seg_rows / seg_stride addressing), forward, data gradient, weight and bias gradients."""
    import rorl_b200.kernels as K
    gen = torch.Generator(device="cuda").manual_seed(4)
    B, L = 5, 203
    batch = torch.randn(B, L + 1, 30, device="cuda", generator=gen)
    act = torch.randn(B, L, 6, device="cuda", generator=gen, requires_grad=True)
    xs = [batch[:, 1:, 0:9], batch[:, :-1, 9:18], act]
    assert not xs[0].is_contiguous() and xs[0].reshape(-1, 9).data_ptr() != xs[0].data_ptr()      # reshape would copy
    Ws = [torch.randn(128, x.shape[-1], device="cuda", generator=gen, requires_grad=True) for x in xs]
    bs = [torch.randn(128, device="cuda", generator=gen, requires_grad=True) for _ in xs]
    y = K.skinny_encoders(xs, Ws, bs, elu=elu)
    assert y.shape == (B, L, 384)
    dy = torch.randn(B, L, 384, device="cuda", generator=gen)
    got = torch.autograd.grad(y, [act] + Ws + bs, dy)
    xd = [x.detach().double() for x in xs]
    xd[2].requires_grad_()
    Wd = [w.detach().double().requires_grad_() for w in Ws]
    bd = [b.detach().double().requires_grad_() for b in bs]
    yr = torch.cat([torch.nn.functional.linear(a, w, b) for a, w, b in zip(xd, Wd, bd)], dim=-1)
    if elu:
        yr = torch.nn.functional.elu(yr)
    ref = torch.autograd.grad(yr, [xd[2]] + Wd + bd, dy.double())
    assert rel_err(y, yr) < 1e-5
    for a, r in zip(got, ref):
        assert a.shape == r.shape
        assert rel_err(a, r) < 1e-4


@pytest.mark.parametrize("M,N,Kk", [(1000, 128, 6), (32608, 128, 9), (77, 256, 16), (3, 128, 1)])
def test_skinny_dgrad(M, N, Kk):
    import rorl_b200.kernels as K
    gen = torch.Generator(device="cuda").manual_seed(5)
    g = torch.randn(M, N + 8, device="cuda", generator=gen)[:, :N]         # row-strided
    W = torch.randn(N, Kk, device="cuda", generator=gen)
    got = K.skinny_dgrad(g, W)
    assert rel_err(got, g.double() @ W.double()) < 1e-5
