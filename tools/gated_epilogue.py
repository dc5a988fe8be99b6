"""Is the gated (SiLU-gate) epilogue of the grouped tcgen05 GEMM what makes the first GEMM of the Mixtral layer slower per
FLOP than the second?  Same weights [E, 2F, K/2], same rows: b200q_moe_grouped_gated_fwd (h [R, F]) against
b200q_moe_grouped_fwd (y [R, 2F], plain epilogue), ms per call."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200q_pkg import pkg
from bench_moe import time_steps
_lib = pkg._lib
dev = torch.device("cuda", 0)
E, F, K = 8, 14336, 4096
g = torch.Generator(device=dev); g.manual_seed(1)
p = torch.randint(0, 256, (E, 2 * F, K // 2), generator=g, device=dev, dtype=torch.uint8)
s = torch.rand(E, 2 * F, device=dev) * 0.01 + 0.001
z = torch.randint(0, 16, (E, 2 * F), device=dev).float()
for R in (4096, 32768):
    x = (torch.randn(R, K, device=dev) * 0.1).to(torch.bfloat16)
    offs = torch.arange(0, E + 1, device=dev, dtype=torch.int32) * (R // E)
    a = time_steps(torch, None, dev, lambda: _lib.moe_grouped_gated_fwd(x, p, s, z, offs), 5, 3)
    b = time_steps(torch, None, dev, lambda: _lib.moe_grouped_fwd(x, p, s, z, offs), 5, 3)
    fl = 2.0 * R * 2 * F * K
    print(json.dumps({"R": R, "gated_ms": round(a, 4), "plain_ms": round(b, 4), "gated_TFLOPs": round(fl / a / 1e9, 1), "plain_TFLOPs": round(fl / b / 1e9, 1)}), flush=True)
