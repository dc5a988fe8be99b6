"""Prefill GEMM 4096->11008 (bf16) at M = 256 / 512 / 1024 with stream-K off and on: the command whose kernel
times (ncu --metrics gpu__time_duration.sum) are quoted in profiles/."""
import os, sys
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
    for sk in (0, 1):
        pkg._lib.tune("gemm_sk", sk)
        for _ in range(3):
            y = pkg._lib.linear_fwd(x, p, s, z)
        torch.cuda.synchronize()
print("ok")
