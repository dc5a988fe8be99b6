"""Mixtral layer decode call (b200q_moe_decode_fwd) at T = 1..3: grid rows = min(E, T k) with device-side expert ranks
(tuning key moe_dec_compact = 1, default) against one grid row per expert (0); ms per call, CUDA graph, bf16."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200q_pkg import pkg
from bench_moe import build_local_moe, E, TOPK, D
dev = torch.device("cuda", 0)
layer = build_local_moe(torch, pkg, list(range(E)), dev)
w13, w2 = layer.stacked_weights()
for T in (1, 2, 3, 4):
    logits = pkg.make_logits(T, E, "random", "cpu", 42).to(dev).float().contiguous()
    x = torch.randn(T, D, device=dev).to(torch.bfloat16)
    out = {"T": T}
    ys = []
    for mode in (0, 1):
        pkg._lib.tune("moe_dec_compact", mode)
        call = lambda: pkg._lib.moe_decode_fwd(x, logits, TOPK, w13, w2)
        for _ in range(3): y = call()
        torch.cuda.synchronize()
        side = torch.cuda.Stream(dev); side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side): call()
        torch.cuda.current_stream(dev).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g): y2 = call()
        for _ in range(5): g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): g.replay()
        e1.record(); torch.cuda.synchronize()
        out["compact_ms" if mode else "per_expert_ms"] = round(e0.elapsed_time(e1) / 50, 4)
        ys.append(y2.clone())
    out["bit_identical"] = bool(torch.equal(ys[0], ys[1]))
    pkg._lib.tune("moe_dec_compact", 1)
    print(json.dumps(out))
