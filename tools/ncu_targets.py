"""Launch the update's dominant kernels at the benchmark shapes a few times each, for `ncu --set full` captures
(tools/gpu_ncu.sh with NCU_K='gemm_bf16x3|selscan|conv1d_silu|colsum|skinny_linear').  The LAST launch of every kernel is
the one to read (earlier ones are warm-ups)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import rorl_b200.kernels as K

dev = torch.device("cuda:0")
torch.manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev)
B, L, D, Ns, R = 32, 1019, 512, 32, 16
M = B * L
reps = 2
# tensor-core GEMMs: in_proj half (TN), efc-8 L1 (grouped TN), weight gradient (NT)
a, w, b = rn(M, 256), rn(512, 256), rn(512)
a8, w8, b8 = rn(M, 384), rn(8, 256, 384), rn(8, 256)
g, x = rn(M, 512), rn(M, 256)
for _ in range(reps):
    K.gemm_tn(a, w, b)
for _ in range(reps):
    K.gemm_tn(a8, w8, b8, 1)
for _ in range(reps):
    K.gemm_nt(g, x)
# fused SSM core forward + backward (selscan fwd with checkpoints, selscan bwd, and the GEMMs around them)
xs, z = rn(B, L, D).requires_grad_(), rn(B, L, D).requires_grad_()
Wx, Wdt = (0.05 * rn(R + 2 * Ns, D)).requires_grad_(), (0.2 * rn(D, R)).requires_grad_()
A_log = torch.log(torch.arange(1, Ns + 1, device=dev, dtype=torch.float32).repeat(D, 1)).requires_grad_()
Dk, bias = rn(D).requires_grad_(), (0.5 * rn(D) - 3).requires_grad_()
start = torch.zeros(B, L, 1, device=dev)
start[:, :19] = 1
dy = rn(B, L, D)
for _ in range(reps):
    y = K.ssm_core(xs, Wx, Wdt, A_log, Dk, z, bias, start)
    torch.autograd.grad(y, (xs, Wx, Wdt, A_log, Dk, z, bias), dy)
# conv, reductions
cw, cb = rn(D, 1, 16).requires_grad_(), rn(D).requires_grad_()
mask = torch.ones(B, L, device=dev)
for _ in range(reps):
    yc = K.causal_conv1d_silu(xs, cw, cb, mask)
    torch.autograd.grad(yc, (xs, cw, cb), dy)
for _ in range(reps):
    K.colsum(a)
torch.cuda.synchronize()
print("done")
