"""tcgen05 GEMM ablations (bench-only tuning key gemm_debug: 1 no weight loads, 2 no activation loads, 4 no MMAs,
8 no tcgen05.st), 4096 -> 11008 bf16, calls captured in a CUDA graph so that the host cost of a call stays out."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200q_pkg import pkg
dev = torch.device("cuda", 0)
K, N = 4096, 11008
p = torch.randint(0, 256, (N, K // 2), device=dev, dtype=torch.uint8)
s = torch.rand(N, device=dev) * 0.01 + 0.001
z = torch.randint(0, 16, (N,), device=dev).float()
for M in (16, 4096):
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for dbg in (0, 1, 2, 3, 4, 8, 12, 15):
        pkg._lib.tune("gemm_debug", dbg)
        for _ in range(2): pkg._lib.linear_fwd(x, p, s, z, out=y)
        torch.cuda.synchronize()
        side = torch.cuda.Stream(dev); side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            pkg._lib.linear_fwd(x, p, s, z, out=y)
        torch.cuda.current_stream(dev).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10): pkg._lib.linear_fwd(x, p, s, z, out=y)
        for _ in range(2): g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): g.replay()
        e1.record(); torch.cuda.synchronize()
        print(json.dumps({"M": M, "gemm_debug": dbg, "us": round(e0.elapsed_time(e1) * 20, 2)}), flush=True)
pkg._lib.tune("gemm_debug", -1)
