"""Mixtral-8x7B INT4 MoE layer benchmark (BASELINE.json configs[3] at 1 GPU, configs[4] expert-parallel
at 2/4/8 GPUs).  Used by bench.py (`--workload moe`, the default for --gpus > 1).

Layer: E = 8 experts, top-2, d = 4096, ffn = 14336, gated MLP  sum_k p_k * w2_e(silu(w1_e x) * (w3_e x)),
weights randn * 0.02 (fp16) quantised per row to INT4 on the GPU; T = 16384 tokens per step in total
(T / N per rank), "random" router logits (routing.py:68; the reference's default "skewed" recipe sends
47 % of the assignments to expert 0 and caps any expert-parallel speed-up at ~2x);
activations bf16, fp32 combine.  tokens/s = T / (max over ranks of the device time per step).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

E, TOPK, D, FFN = 8, 2, 4096, 14336
T_GLOBAL = 16384
FLOPS_PER_TOKEN = TOPK * 3 * 2 * D * FFN          # 704,643,072


def build_local_moe(torch, pkg, experts, dev):
    w1, w3, w2 = [], [], []
    for e in experts:
        g = torch.Generator(device=dev)
        g.manual_seed(1000 + e)
        w1.append((torch.randn(FFN, D, generator=g, device=dev) * 0.02).half())
        w3.append((torch.randn(FFN, D, generator=g, device=dev) * 0.02).half())
        w2.append((torch.randn(D, FFN, generator=g, device=dev) * 0.02).half())
    moe = pkg.QuantizedMoE.from_gated_fp16_weights(w1, w3, w2)
    del w1, w3, w2
    moe.stacked_weights()
    return moe


def time_steps(torch, dist, dev, fn, steps, warmup):
    for _ in range(max(warmup, 3)):
        fn()
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def cpu_moe_baseline(tokens=8):
    """The reference-composed CPU layer (routing.py + dequantize_weights + matmul) on a few
    tokens of a Mixtral-size layer would need 8 x 3 x 58.7 M-weight dequantisations per call (minutes); the
    bounded sample keeps the Mixtral d / ffn but only the 2 experts that `tokens` = 8 tokens of one routing
    draw hit most, i.e. it times 2 experts x 3 projections and scales to the per-token cost."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import int4_oracle as oracle
    try:
        import c_oracle
    except Exception:
        c_oracle = None
    rng = np.random.default_rng(0)
    x = rng.standard_normal((tokens, D), dtype=np.float32)

    def proj(n, k):
        packed = rng.integers(0, 256, size=(n, k // 2), dtype=np.uint8)
        s = (rng.random(n, dtype=np.float32) * 0.004 + 0.001).astype(np.float32)
        z = rng.integers(0, 16, size=n).astype(np.float32)
        return c_oracle.Linear(packed, s, z) if c_oracle else (lambda a: oracle.reference_quantized_linear(a, packed, s, z))

    w1, w3, w2 = proj(FFN, D), proj(FFN, D), proj(D, FFN)
    t0 = time.perf_counter()
    reps = 2
    for _ in range(reps):
        g, u = w1(x), w3(x)
        h = (oracle.silu(g) * u).astype(np.float32)
        w2(h)
    dt = (time.perf_counter() - t0) / reps          # one expert, `tokens` rows
    # a token visits TOPK experts; with `tokens` rows per expert call: tokens / (TOPK * dt) tokens/s
    threads = c_oracle.max_threads() if c_oracle else 1
    return tokens / (TOPK * dt), threads, f"{reps} x one Mixtral expert (w1, w3, silu-gate, w2) on {tokens} rows, C restatement of dequantize_weights + matmul"


def run_moe(args):
    import numpy as np
    import torch
    from b200q_pkg import pkg
    from bench import ClockSampler, measured_peaks

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    assert E % world == 0 and T_GLOBAL % world == 0

    experts = pkg.shard_experts(E, rank, world)
    moe = build_local_moe(torch, pkg, experts, dev)
    layer = pkg.ExpertParallelMoE(moe, E, TOPK)
    t_loc = T_GLOBAL // world
    g = torch.Generator(device=dev)
    g.manual_seed(42 + rank)
    x = torch.randn(t_loc, D, generator=g, device=dev).to(torch.bfloat16)
    logits = torch.randn(t_loc, E, generator=g, device=dev)

    out = layer(x, logits)
    assert out.shape == (t_loc, D) and bool(torch.isfinite(out).all())

    with ClockSampler(local_rank) as clocks:
        ms = time_steps(torch, dist, dev, lambda: layer(x, logits), args.steps, args.warmup)
    tokens_per_s = T_GLOBAL / (ms * 1e-3)

    # end to end: activations and router logits start in pinned host memory, the result returns to the host
    xh = x.cpu().pin_memory()
    lh = logits.cpu().pin_memory()
    oh = torch.empty(t_loc, D).pin_memory()

    def e2e_step():
        xd = xh.to(dev, non_blocking=True)
        ld = lh.to(dev, non_blocking=True)
        oh.copy_(layer(xd, ld), non_blocking=True)

    e2e_ms = time_steps(torch, dist, dev, e2e_step, max(3, min(args.steps, 20)), 3)
    stats = dict(layer.last_stats)
    if dist is not None:
        dist.barrier()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peaks, peak_src = measured_peaks()
    peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    tflops = FLOPS_PER_TOKEN * T_GLOBAL / (ms * 1e-3) / 1e12
    launches_per_step = 12 + (2 if world > 1 else 0)
    line = {
        "metric": "mixtral_moe_int4_layer_tokens_per_s", "value": tokens_per_s, "unit": "tokens/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "Mixtral-8x7B INT4 MoE layer (8 experts, top-2, d=4096, ffn=14336), "
                               f"{T_GLOBAL} tokens/step, random routing" + (", expert-parallel" if world > 1 else ""),
                   "tokens_per_step": T_GLOBAL, "experts_per_rank": E // world,
                   "l2": "inputs larger than L2: 705 MB of INT4 expert weights + 134 MB of activations per step",
                   "parallelism": f"ep{world}" if world > 1 else "single GPU"},
        "roofline": {"bound": "tensor", "achieved": tflops / world, "peak": peak, "unit": "TFLOP/s",
                     "frac": tflops / world / peak, "traffic": None, "peak_source": peak_src + " bf16 sustained",
                     "algorithmic_flops_per_step": FLOPS_PER_TOKEN * T_GLOBAL},
        "e2e": {"value": T_GLOBAL / (e2e_ms * 1e-3), "unit": "tokens/s",
                "h2d_bytes_per_step": t_loc * D * 2 + t_loc * E * 4, "d2h_bytes_per_step": t_loc * D * 4,
                "api": "ExpertParallelMoE.forward on pinned host activations"},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks.summary(),
        "ep": stats,
    }
    # the same sub-object the N = 1 (GEMV-headline) line carries: `moe.value` is one consistent series over N = 1, 2, 4, 8
    line["moe"] = {k: line[k] for k in ("metric", "value", "unit", "n_gpus", "ms_per_step", "dtype", "roofline")}
    line["moe"]["workload"] = line["config"]["workload"]
    if not args.no_cpu:
        v, threads, sample = cpu_moe_baseline()
        line["cpu_baseline"] = {"value": v, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
