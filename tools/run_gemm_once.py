"""One prefill GEMM (4096->11008, M=2048, bf16) a few times: the command profiled by ncu for profiles/."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200q_pkg import pkg
dev = torch.device("cuda", 0)
K, N, M = 4096, 11008, int(os.environ.get("M", "2048"))
p = torch.randint(0, 256, (N, K // 2), device=dev, dtype=torch.uint8)
s = torch.rand(N, device=dev) * 0.01 + 0.001
z = torch.randint(0, 16, (N,), device=dev).float()
x = torch.randn(M, K, device=dev).to(torch.bfloat16)
if os.environ.get("SK"): pkg._lib.tune("gemm_sk", int(os.environ["SK"]))
for _ in range(4):
    y = pkg._lib.linear_fwd(x, p, s, z)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
