"""Token-tile height sweep of the tcgen05 GEMM (tuning key gemm_bn)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200q_pkg import pkg
_lib = pkg._lib
dev = torch.device("cuda", 0)
K, N = 4096, 11008
p = torch.randint(0, 256, (N, K // 2), device=dev, dtype=torch.uint8)
s = torch.rand(N, device=dev) * 0.01 + 0.001
z = torch.randint(0, 16, (N,), device=dev).float()
for M in (256, 384, 512, 768, 1024, 1536, 2048, 4096):
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    row = {"M": M}
    for bn in (128, 192, 256, -1):
        _lib.tune("gemm_bn", bn)
        for _ in range(3):
            y = _lib.linear_fwd(x, p, s, z)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            y = _lib.linear_fwd(x, p, s, z)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        row[f"bn{bn}"] = round(2.0 * M * N * K / ms / 1e9, 0)
    print(json.dumps(row), flush=True)
_lib.tune("gemm_bn", -1)
