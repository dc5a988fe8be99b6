"""Per-k-block SM-clock trace of CTA 0 of the tcgen05 GEMM (bench-only, gemm_debug bit 4 = 16).
Rows: 0 producer (stage free, about to issue), 1 MMA thread: stage landed, 2 MMA thread: A slot filled,
3 MMA thread: MMAs + commits issued, 4 dequant warp: stage landed, 5 dequant warp: A slot free, 6 dequant warp: A slot handed over."""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200q_pkg import pkg
lib = pkg._lib.load()
lib.b200q_debug_gemm_trace.argtypes = [ctypes.c_void_p]
dev = torch.device("cuda", 0)
K, N, M = 4096, 11008, int(os.environ.get("M", "4096"))
p = torch.randint(0, 256, (N, K // 2), device=dev, dtype=torch.uint8)
s = torch.rand(N, device=dev) * 0.01 + 0.001
z = torch.randint(0, 16, (N,), device=dev).float()
x = torch.randn(M, K, device=dev).to(torch.bfloat16)
for _ in range(2): pkg._lib.linear_fwd(x, p, s, z)
pkg._lib.tune("gemm_debug", 16 | int(os.environ.get("DBG", "0")))
pkg._lib.linear_fwd(x, p, s, z)
torch.cuda.synchronize()
buf = np.zeros(8 * 96, dtype=np.int64)
pkg._lib.check(lib.b200q_debug_gemm_trace(buf.ctypes.data), "trace")
t = buf.reshape(8, 96)
t0 = t[0, 0]
names = ["prod:free", "mma:landed", "mma:Afull", "mma:issued", "dq:landed", "dq:Afree", "dq:handed"]
print("k-block " + " ".join(f"{n:>11s}" for n in names))
for kb in list(range(0, 12)) + list(range(40, 52)):
    row = []
    for r in range(7):
        v = t[r, kb]
        row.append(f"{(v - t0) if v else -1:11d}")
    print(f"{kb:7d} " + " ".join(row))
d = np.diff(t[3, 16:80].astype(np.int64))
print("steady state: clk per k-block (MMA issue to MMA issue), median", int(np.median(d)), "mean", int(d.mean()))
