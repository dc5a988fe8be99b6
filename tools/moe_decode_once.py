"""A few MoE decode calls (T = 1, 4) for an ncu launch list: which kernels one QuantizedMoE.forward_routed call launches
and how long each runs (serialised, cold caches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200q_pkg import pkg
from bench_moe import build_local_moe, make_inputs, E, TOPK
dev = torch.device("cuda", 0)
layer = build_local_moe(torch, pkg, list(range(E)), dev)
for T in (1, 4):
    x, logits = make_inputs(torch, 0, T, "random", dev)
    for _ in range(3):
        y = layer.forward_routed(x, logits, top_k=TOPK)
    torch.cuda.synchronize()
    print(T, float(y.float().abs().sum()))
