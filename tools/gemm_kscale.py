"""tcgen05 GEMM at M = 16 (32-token tiles): time vs K separates the per-k-block cost from the fixed cost."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200q_pkg import pkg
dev = torch.device("cuda", 0)
N = 11008
for M in (16, 64):
    for K in (512, 1024, 2048, 4096, 8192):
        p = torch.randint(0, 256, (N, K // 2), device=dev, dtype=torch.uint8)
        s = torch.rand(N, device=dev) * 0.01 + 0.001
        z = torch.randint(0, 16, (N,), device=dev).float()
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        pkg._lib.tune("force_path", 3)
        for _ in range(3): pkg._lib.linear_fwd(x, p, s, z)
        torch.cuda.synchronize()
        y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        side = torch.cuda.Stream(dev); side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            pkg._lib.linear_fwd(x, p, s, z, out=y)
        torch.cuda.current_stream(dev).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):                  # the host cost of a call (3 tensor-map encodes + 2 launches) stays out
            for _ in range(20): pkg._lib.linear_fwd(x, p, s, z, out=y)
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): g.replay()
        e1.record(); torch.cuda.synchronize()
        print(json.dumps({"M": M, "K": K, "kblocks": K // 64, "us": round(e0.elapsed_time(e1) * 10, 2)}), flush=True)
