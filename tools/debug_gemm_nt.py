import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rorl_b200.kernels as K
torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)
dev = "cuda"
R, M, N = 128, 128, 128
A = torch.ones(R, M, device=dev); B = torch.ones(R, N, device=dev)
D = K.gemm_nt(A, B); torch.cuda.synchronize()
print("ones:", D[:2, :6].tolist(), "expect", R, "nonzero frac", float((D != 0).float().mean()))
A = torch.arange(M, device=dev, dtype=torch.float32).repeat(R, 1)      # A[r, m] = m
B = torch.zeros(R, N, device=dev); B[0] = 1.0                           # only row 0
D = K.gemm_nt(A, B); torch.cuda.synchronize()
print("D[m,n]=m? col0:", D[:10, 0].tolist(), D[120:, 5].tolist())
B = torch.zeros(R, N, device=dev); B[0] = torch.arange(N, device=dev, dtype=torch.float32)
A = torch.zeros(R, M, device=dev); A[0] = 1.0
D = K.gemm_nt(A, B); torch.cuda.synchronize()
print("D[m,n]=n? row0:", D[0, :10].tolist(), D[3, 120:].tolist())
A = torch.randn(R, M, device=dev); B = torch.randn(R, N, device=dev)
D = K.gemm_nt(A, B); ref = A.double().t() @ B.double()
print("rand err", float((D - ref).abs().max() / ref.abs().max()))
for r_only in (0, 1, 7, 8, 31, 32, 100):
    A2 = torch.zeros(R, M, device=dev); B2 = torch.zeros(R, N, device=dev)
    A2[r_only] = A[r_only]; B2[r_only] = B[r_only]
    D = K.gemm_nt(A2, B2); ref = A2.double().t() @ B2.double()
    print("single reduction row", r_only, "err", float((D - ref).abs().max() / ref.abs().max()))
R = 1000
A = torch.randn(R, 80, device=dev); B = torch.randn(R, 512, device=dev)
D = K.gemm_nt(A, B); ref = A.double().t() @ B.double()
print("R=1000 80x512 err", float((D - ref).abs().max() / ref.abs().max()), "splits", K.N.lib().rorl_gemm_nt_splits(80, 512, R, 1))
