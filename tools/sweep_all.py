"""SURVEY section 8(d) configurations in one run (1 GPU): decode GEMV M in {1,2,4,8,16} on both Llama shapes,
prefill GEMM M in {512..4096} (bf16 and fp32 activations), Mixtral MoE layer T in {1,4,16,512,2048,8192,16384} with
random and skewed routing.  One JSON line per configuration (device-resident timing, CUDA events, CUDA graph for the
decode shapes); the table in profiles/ is generated from this output."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200q_pkg import pkg
from bench import measured_peaks
_lib = pkg._lib
lib = _lib.load()
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
peaks, _ = measured_peaks()
HBM, TF = float(peaks["hbm_gbs"]), float(peaks["bf16_tflops"])


def timed(fn, steps, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def pool(N, K, n):
    out = []
    for i in range(n):
        g = torch.Generator(device=dev); g.manual_seed(i)
        out.append((torch.randint(0, 256, (N, K // 2), generator=g, device=dev, dtype=torch.uint8),
                    torch.rand(N, generator=g, device=dev) * 0.01 + 0.001,
                    torch.randint(0, 16, (N,), generator=g, device=dev).float()))
    return out


def decode(K, N, batch=(1, 2, 4, 8, 16)):
    layers = pool(N, K, 24)
    for M in batch:
        x = torch.randn(M, K, device=dev)
        y = torch.empty(M, N, device=dev)
        ws = torch.zeros(max(lib.b200q_linear_ws_bytes(M, N, K), 16), dtype=torch.uint8, device=dev)

        def launch_all(sp):
            for i, (p, s, z) in enumerate(layers):
                nxt = layers[(i + 1) % len(layers)][0]
                _lib.check(lib.b200q_linear_fwd_next(x.data_ptr(), 0, p.data_ptr(), s.data_ptr(), z.data_ptr(), y.data_ptr(), 0,
                                                     M, N, K, ws.data_ptr(), ws.numel(), 1, sp, nxt.data_ptr(), nxt.numel()), "fwd")
        side = torch.cuda.Stream(dev); side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            launch_all(side.cuda_stream)
        torch.cuda.current_stream(dev).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            launch_all(torch.cuda.current_stream(dev).cuda_stream)
        us = timed(g.replay, 50) * 1e3 / len(layers)
        nbytes = N * K // 2 + 8 * N + 4 * M * K + 4 * M * N
        print(json.dumps({"config": "decode", "K": K, "N": N, "M": M, "us_per_launch": round(us, 3), "GBps": round(nbytes / us / 1e3, 1),
                          "frac_hbm_peak": round(nbytes / us / 1e3 / HBM, 3)}), flush=True)
    del layers


def prefill(K, N):
    p, s, z = pool(N, K, 1)[0]
    for dt in (torch.bfloat16, torch.float32):
        for M in (256, 512, 1024, 2048, 4096):
            x = torch.randn(M, K, device=dev).to(dt)
            ms = timed(lambda: _lib.linear_fwd(x, p, s, z), 20)
            tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
            print(json.dumps({"config": "prefill", "K": K, "N": N, "M": M, "x": str(dt).replace("torch.", ""), "ms": round(ms, 4),
                              "TFLOPs": round(tf, 1), "frac_bf16_peak": round(tf / TF, 3)}), flush=True)


def moe():
    from bench_moe import build_local_moe, E, TOPK, D, FFN, FLOPS_PER_TOKEN
    layer = build_local_moe(torch, pkg, list(range(E)), dev)
    for dist_name in ("random", "skewed"):
        for T in (1, 4, 16, 512, 2048, 8192, 16384):
            logits = pkg.make_logits(T, E, dist_name, "cpu", 42).to(dev)
            x = torch.randn(T, D, device=dev).to(torch.bfloat16)
            ms = timed(lambda: layer.forward_routed(x, logits, top_k=TOPK), 20 if T >= 2048 else 50)
            rec = {"config": "moe", "routing": dist_name, "T": T, "ms": round(ms, 4), "tokens_per_s": round(T / (ms * 1e-3), 1)}
            if T >= 512:
                tf = FLOPS_PER_TOKEN * T / (ms * 1e-3) / 1e12
                rec.update({"TFLOPs": round(tf, 1), "frac_bf16_peak": round(tf / TF, 3)})
            else:
                hit = int(torch.unique(pkg.route(logits, TOPK).expert_indices).numel())
                nbytes = hit * 3 * FFN * D // 2
                rec.update({"experts_hit": hit, "GBps": round(nbytes / ms / 1e6, 1), "frac_hbm_peak": round(nbytes / ms / 1e6 / HBM, 3)})
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["decode", "prefill", "moe"]
    if "prefill" in what:       # first: the multi-millisecond MoE / fp32 blocks leave the part power-capped for a while
        prefill(4096, 11008); prefill(11008, 4096)
    if "decode" in what:
        decode(4096, 11008); decode(11008, 4096)
        decode(4096, 14336, (1,)); decode(14336, 4096, (1,))        # one Mixtral expert's projections
    if "moe" in what:
        moe()
