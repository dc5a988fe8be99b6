import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import int4_oracle as oracle
from b200q_pkg import pkg
def cuda(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
for (M,N,K) in [(5,3000,6144),(1,11008,4096),(8,2371,2048),(2,200,1024)]:
    rng = np.random.default_rng(1000 * M + N + K)
    packed = rng.integers(0, 256, size=(N, K // 2), dtype=np.uint8)
    scales = (rng.random(N, dtype=np.float32) * 0.01 + 0.001).astype(np.float32)
    zps = rng.integers(0, 16, size=N).astype(np.float32)
    x = (rng.standard_normal((M, K)) * rng.choice([1e-3, 1.0, 300.0], size=(M, 1))).astype(np.float32)
    ref = oracle.reference_quantized_linear(x, packed, scales, zps, acc=np.float64)
    for path in (5, 2, 1):
        pkg._lib.tune("force_path", path)
        y = pkg._lib.linear_fwd(cuda(x), cuda(packed), cuda(scales), cuda(zps)).cpu().numpy()
        pkg._lib.tune("force_path", -1)
        rel = [float(np.abs(ref[m]-y[m]).max()/np.abs(ref[m]).max()) for m in range(M)]
        print(M,N,K,"path",path,"rel err per row", ["%.1e"%r for r in rel])
