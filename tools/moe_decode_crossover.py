"""Mixtral layer, T = 1..16 tokens: the decode call (b200q_moe_decode_fwd: grouped GEMVs) against the prefill-shaped path
(route -> gather -> grouped tcgen05 GEMMs -> combine), ms per layer call (CUDA events, eager)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200q_pkg import pkg
from bench_moe import build_local_moe, make_inputs, time_steps, E, TOPK
dev = torch.device("cuda", 0)
layer = build_local_moe(torch, pkg, list(range(E)), dev)
w13, w2 = layer.stacked_weights()
for T in (1, 2, 4, 6, 8, 10, 12, 16):
    x, logits = make_inputs(torch, 0, T, "random", dev)
    lg = logits.float().contiguous()
    pkg._lib.tune("moe_dec_hm", 0)
    a = time_steps(torch, None, dev, lambda: pkg._lib.moe_decode_fwd(x, lg, TOPK, w13, w2), 30, 5)
    pkg._lib.tune("moe_dec_hm", 1)
    c = time_steps(torch, None, dev, lambda: pkg._lib.moe_decode_fwd(x, lg, TOPK, w13, w2), 30, 5)
    pkg._lib.tune("moe_dec_hm", -1)
    b = time_steps(torch, None, dev, lambda: layer.forward_dispatched(x, pkg.route(logits, TOPK)), 30, 5)
    print(json.dumps({"T": T, "decode_call_ms": round(a, 4), "decode_call_mid_batch_kernel_ms": round(c, 4), "grouped_gemm_ms": round(b, 4)}))
