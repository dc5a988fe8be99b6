import sys, os, json
sys.path.insert(0, "/root/repo")
import torch
from b200q_pkg import pkg
from bench_moe import build_local_moe, E, TOPK, D, FFN
dev = torch.device("cuda", 0)
layer = build_local_moe(torch, pkg, list(range(E)), dev)
for T in (1, 4, 16):
    logits = pkg.make_logits(T, E, "random", "cpu", 42).to(dev)
    x = torch.randn(T, D, device=dev).to(torch.bfloat16)
    for _ in range(3): y = layer.forward_routed(x, logits, top_k=TOPK)
    torch.cuda.synchronize()
    side = torch.cuda.Stream(dev); side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        layer.forward_routed(x, logits, top_k=TOPK)
    torch.cuda.current_stream(dev).wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        y2 = layer.forward_routed(x, logits, top_k=TOPK)
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(T, "graph ms", e0.elapsed_time(e1)/50, "equal", bool(torch.equal(y, y2)))
