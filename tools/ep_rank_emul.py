"""One expert-parallel rank's expert phase at N = 8, emulated on one GPU: 4111 received rows, one local Mixtral expert,
64 ranges of which one is non-empty (what b200q_ep_plan produces after merging).  Prints ms per call; run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel list."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200q_pkg import pkg
from bench_moe import build_local_moe, D
dev = torch.device("cuda", 0)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 4111
moe = build_local_moe(torch, pkg, [0], dev)
rows = torch.randn(R, D, device=dev).to(torch.bfloat16)
starts = torch.zeros(64, dtype=torch.int32, device=dev); ends = torch.zeros(64, dtype=torch.int32, device=dev)
ends[0] = R
rexp = (torch.arange(64, device=dev) // 8).to(torch.int32)
offsets = torch.tensor([0, R], dtype=torch.int32, device=dev)
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
n = 3 if os.environ.get("NCU") else 20
a = timed(lambda: moe.forward_ranges(rows, starts, ends, rexp, all_rows_covered=True), n)
b = timed(lambda: moe.forward_grouped(rows, offsets), n)
fl = R * 3 * 2 * 4096 * 14336
print(json.dumps({"rows": R, "mapped_ms": round(a, 4), "mapped_TFLOPs": round(fl / a / 1e9, 1), "offsets_ms": round(b, 4), "offsets_TFLOPs": round(fl / b / 1e9, 1)}))
