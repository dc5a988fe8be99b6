"""Activation preparation of the tcgen05 GEMM in isolation (as far as the C ABI allows): a linear with ONE weight-row tile
(N = 128), so the preparation pass over x [M, K] dominates the call.  B200Q_LIB=<other build> for an A/B on one box."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200q_pkg import pkg
from bench_moe import time_steps
_lib = pkg._lib
dev = torch.device("cuda", 0)
for (M, K, dt) in [(32768, 14336, torch.bfloat16), (4096, 11008, torch.bfloat16), (32768, 4096, torch.bfloat16), (4096, 11008, torch.float32)]:
    N = 128
    g = torch.Generator(device=dev); g.manual_seed(1)
    p = torch.randint(0, 256, (N, K // 2), generator=g, device=dev, dtype=torch.uint8)
    s = torch.rand(N, device=dev) * 0.01 + 0.001
    z = torch.randint(0, 16, (N,), device=dev).float()
    x = torch.randn(M, K, device=dev).to(dt)
    ms = time_steps(torch, None, dev, lambda: _lib.linear_fwd(x, p, s, z), 20, 5)
    print(json.dumps({"lib": os.environ.get("B200Q_LIB", "default"), "M": M, "K": K, "x": str(dt).split(".")[1], "ms": round(ms, 4)}), flush=True)
