"""Prefill GEMM throughput (tcgen05 path): TFLOP/s = 2*M*N*K / time, CUDA events, weights + activations resident."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200q_pkg import pkg
_lib = pkg._lib
dev = torch.device("cuda", 0)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0}
for (K, N) in [(4096, 11008), (11008, 4096)]:
    g = torch.Generator(device=dev); g.manual_seed(1)
    p = torch.randint(0, 256, (N, K // 2), generator=g, device=dev, dtype=torch.uint8)
    s = torch.rand(N, device=dev) * 0.01 + 0.001
    z = torch.randint(0, 16, (N,), device=dev).float()
    for dtype in (torch.bfloat16, torch.float32):
        for M in (64, 256, 512, 1024, 2048, 4096):
            x = torch.randn(M, K, device=dev).to(dtype)
            for _ in range(3):
                y = _lib.linear_fwd(x, p, s, z)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps):
                y = _lib.linear_fwd(x, p, s, z)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            tf = 2.0 * M * N * K / ms / 1e9
            print(json.dumps({"K": K, "N": N, "M": M, "dtype": str(dtype), "ms": round(ms, 4), "TFLOPs": round(tf, 1),
                              "frac_of_measured_bf16_peak": round(tf / peaks["bf16_tflops"], 3)}), flush=True)
