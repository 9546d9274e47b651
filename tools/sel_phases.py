"""Diagnostic: time the selective-scan forward with phases disabled (rorl_selscan_debug bit mask)."""
import os, sys, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rorl_b200.kernels as K
import rorl_b200._native as N
dev = torch.device("cuda:0")
B, L, D, Ns = 32, 1018, 512, 32
rn = lambda *s: torch.randn(*s, device=dev)
u, delta, z = rn(B, L, D), 0.5 * rn(B, L, D) - 1, rn(B, L, D)
Bm, Cm = rn(B, L, Ns), rn(B, L, Ns)
A = -torch.exp(0.3 * rn(D, Ns)); Dk, bias = rn(D), rn(D)
start = torch.zeros(B, L, device=dev); start[:, :18] = 1
lib = N.lib()
lib.rorl_selscan_debug.argtypes = [ctypes.c_int]; lib.rorl_selscan_debug.restype = None
def t(n=20):
    f = lambda: K.selective_scan_tm(u, delta, A, Bm, Cm, Dk, z, bias, start, True)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
with torch.no_grad():
    for mask, name in [(0, "full"), (1, "no softplus"), (16, "no start ldg"), (2, "no silu(z)"), (8, "no y store"), (4, "no scan loop"),
                       (1 | 2 | 16, "scan + plain copies"), (1 | 2 | 8 | 16, "scan only, no store"), (4 | 1 | 2 | 16, "copies only")]:
        lib.rorl_selscan_debug(mask)
        print(f"{name:28s} {t():8.1f} us")
lib.rorl_selscan_debug(0)
