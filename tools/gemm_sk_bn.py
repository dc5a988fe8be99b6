"""Stream-K x token-tile height on the prefill shapes that under-fill their waves (CUDA-graph timing)."""
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
for M in (256, 512, 1024):
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for sk, bn in ((0, -1), (1, 256), (1, 192), (1, 128), (0, 192), (0, 128)):
        pkg._lib.tune("gemm_sk", sk); pkg._lib.tune("gemm_bn", bn)
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
        us = e0.elapsed_time(e1) * 20
        print(json.dumps({"M": M, "sk": sk, "bn": bn, "us": round(us, 2), "TFLOPs": round(2.0 * M * N * K / us / 1e6, 1)}), flush=True)
pkg._lib.tune("gemm_sk", -1); pkg._lib.tune("gemm_bn", -1)
