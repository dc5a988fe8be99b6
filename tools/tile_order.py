"""tcgen05 GEMM tile order (tuning key gemm_mt_major): 0 = all token tiles of a weight tile first (every weight tile
streams ALL activations again), 1 = all weight tiles of a token tile first (the CTAs of a wave share few activation tiles
and stream the weights).  Mixtral layer (grouped GEMMs) and the dense prefill GEMM, ms per call, CUDA events."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200q_pkg import pkg
from bench_moe import build_local_moe, make_inputs, time_steps, E, TOPK
_lib = pkg._lib
dev = torch.device("cuda", 0)
for (K, N) in [(11008, 4096), (4096, 11008)]:
    g = torch.Generator(device=dev); g.manual_seed(1)
    p = torch.randint(0, 256, (N, K // 2), generator=g, device=dev, dtype=torch.uint8)
    s = torch.rand(N, device=dev) * 0.01 + 0.001
    z = torch.randint(0, 16, (N,), device=dev).float()
    for dtype in (torch.bfloat16, torch.float32):
        for M in (1024, 2048, 4096, 8192):
            x = torch.randn(M, K, device=dev).to(dtype)
            out = {"K": K, "N": N, "M": M, "x": str(dtype).split(".")[1]}
            ys = []
            for mode in (0, 1):
                _lib.tune("gemm_mt_major", mode)
                ms = time_steps(torch, None, dev, lambda: _lib.linear_fwd(x, p, s, z), 10, 3)
                out[f"order{mode}_ms"] = round(ms, 4)
                ys.append(_lib.linear_fwd(x, p, s, z))
            out["bit_identical"] = bool(torch.equal(ys[0], ys[1]))
            out["TFLOPs_best"] = round(2.0 * M * N * K / min(out["order0_ms"], out["order1_ms"]) / 1e9, 1)
            print(json.dumps(out), flush=True)
    del p, x
_lib.tune("gemm_mt_major", -1)
layer = build_local_moe(torch, pkg, list(range(E)), dev)
for routing in ("random", "skewed"):
    for T in (2048, 8192, 16384):
        x, logits = make_inputs(torch, 0, T, routing, dev)
        out = {"moe_T": T, "routing": routing}
        ys = []
        for mode in (0, 1):
            _lib.tune("gemm_mt_major", mode)
            ms = time_steps(torch, None, dev, lambda: layer.forward_routed(x, logits, top_k=TOPK), 5, 3)
            out[f"order{mode}_ms"] = round(ms, 4)
            ys.append(layer.forward_routed(x, logits, top_k=TOPK))
        out["bit_identical"] = bool(torch.equal(ys[0], ys[1]))
        out["tokens_per_s_best"] = round(T / min(out["order0_ms"], out["order1_ms"]) * 1e3)
        print(json.dumps(out), flush=True)
_lib.tune("gemm_mt_major", -1)
