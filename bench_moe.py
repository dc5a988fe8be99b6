"""Mixtral-8x7B INT4 MoE layer benchmark (BASELINE.json configs[3] at 1 GPU, configs[4] expert-parallel
at 2/4/8 GPUs).  Used by bench.py (`--workload moe`, the default for --gpus > 1 and under torch.distributed.run).

Layer: E = 8 experts, top-2, d = 4096, ffn = 14336, gated MLP  sum_k p_k * w2_e(silu(w1_e x) * (w3_e x)),
weights randn * 0.02 (fp16) quantised per row to INT4 on the GPU; T = 16384 tokens per step in total
(T / N per rank); activations bf16, fp32 combine.  tokens/s = T / (max over ranks of the device time per step).
Headline routing: "random" logits (routing.py:68).  The reference's default "skewed" recipe (routing.py:54-66: 47 % of
the assignments go to expert 0, which caps plain expert parallelism at ~2x) is reported alongside, with
the two hottest experts replicated on every rank.

Before any timing the expert-parallel output is checked against the unsharded layer on the same tokens (every rank,
all its tokens) and against the float64 oracle on sampled tokens (rank 0)."""
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
EXPERT_BYTES = 3 * FFN * D // 2                   # packed INT4 bytes of one expert (w1, w3, w2)


def expert_weights(torch, e, dev):
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + e)
    return ((torch.randn(FFN, D, generator=g, device=dev) * 0.02).half(),
            (torch.randn(FFN, D, generator=g, device=dev) * 0.02).half(),
            (torch.randn(D, FFN, generator=g, device=dev) * 0.02).half())


def build_local_moe(torch, pkg, experts, dev):
    w1, w3, w2 = [], [], []
    for e in experts:
        a, b, c = expert_weights(torch, e, dev)
        w1.append(a); w3.append(b); w2.append(c)
    moe = pkg.QuantizedMoE.from_gated_fp16_weights(w1, w3, w2)
    del w1, w3, w2
    moe.stacked_weights()
    return moe


def make_inputs(torch, rank, t_loc, routing, dev):
    """Rank `rank`'s tokens and router logits (any rank can regenerate any other rank's)."""
    g = torch.Generator(device=dev)
    g.manual_seed(42 + rank)
    x = torch.randn(t_loc, D, generator=g, device=dev).to(torch.bfloat16)
    logits = torch.randn(t_loc, E, generator=g, device=dev)
    if routing == "skewed":                        # routing.py:54-66: log of 1/(e+1) probabilities + N(0, 0.5) noise
        probs = 1.0 / (torch.arange(E, device=dev, dtype=torch.float32) + 1)
        logits = torch.log(probs / probs.sum() + 1e-10).unsqueeze(0) + logits * 0.5
    return x, logits


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


def oracle_check(torch, full, x, logits, n_sample=64):
    """float64 oracle (composed from the reference's primitives) on sampled tokens of one rank."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import int4_oracle as oracle
    try:
        import c_oracle
    except Exception:
        c_oracle = None
    rng = np.random.default_rng(0)
    sel = np.sort(rng.choice(x.shape[0], size=min(n_sample, x.shape[0]), replace=False))
    xs = x[sel].float().cpu().numpy()
    lg = logits[sel].float().cpu().numpy()
    idx, w = oracle.softmax_topk(lg, TOPK)
    ref = np.zeros((len(sel), D), dtype=np.float64)

    def lin(mod, a):
        p, s, z = mod.packed_weights.contiguous().cpu().numpy(), mod.scales.contiguous().cpu().numpy(), mod.zero_points.contiguous().cpu().numpy()
        if c_oracle is not None:
            return c_oracle.Linear(p, s, z)(a.astype(np.float32)).astype(np.float64)
        return oracle.reference_quantized_linear(a, p, s, z, acc=np.float64)

    for e in range(E):
        rows, slots = np.nonzero(idx == e)
        if len(rows) == 0:
            continue
        a = xs[rows]
        h = (oracle.silu(lin(full.experts[e], a)) * lin(full.experts_up[e], a))
        # the product keeps h in bf16 between the two GEMMs (activations are bf16): same rounding here
        h = torch.from_numpy(h.astype(np.float32)).to(torch.bfloat16).float().numpy()
        y = lin(full.experts_down[e], h)
        y = torch.from_numpy(y.astype(np.float32)).to(torch.bfloat16).float().numpy().astype(np.float64)
        np.add.at(ref, rows, y * w[rows, slots][:, None])
    return sel, ref


def cpu_moe_reference(steps=2, tokens=512):
    """The reference's CPU path for the stated layer, composed from ITS OWN primitives (oracle/_ref/reference:
    routing.py simulate_routing-style top-k, create_expert_inputs, dequantize_weights + matmul per projection,
    combine_expert_outputs) on `tokens` tokens of the Mixtral-size layer: every call dequantises all 24 projections
    (as QuantizedMoE.forward does, moe_int4_module.py:71-72).  Returns (tokens/s, cores, kind, sample)."""
    import numpy as np
    import torch
    torch.set_num_threads(os.cpu_count() or 1)       # (torch.distributed.run exports OMP_NUM_THREADS=1)
    ref_root = os.path.join(ROOT, "oracle", "_ref", "reference")
    kind = "reference"
    if os.path.isdir(os.path.join(ref_root, "python")):
        sys.path.insert(0, ref_root)
        from python.quantize import dequantize_weights
        from benchmark.moe_grouped_gemm.routing import RoutingResult, create_expert_inputs, combine_expert_outputs
    else:
        kind = "port"
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import int4_oracle as o
        dequantize_weights = lambda p, s, z: torch.from_numpy(o.dequantize_weights(p.numpy(), s.numpy(), z.numpy()))
        RoutingResult = create_expert_inputs = combine_expert_outputs = None
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(0)
    mk = lambda n, k: (torch.randint(0, 256, (n, k // 2), dtype=torch.uint8, generator=g), torch.rand(n, generator=g) * 0.004 + 0.001,
                       torch.randint(0, 16, (n,), generator=g).float())
    w1 = [mk(FFN, D) for _ in range(E)]
    w3 = [mk(FFN, D) for _ in range(E)]
    w2 = [mk(D, FFN) for _ in range(E)]
    x = torch.randn(tokens, D, generator=g)
    logits = torch.randn(tokens, E, generator=g)

    def step():
        probs = torch.softmax(logits, dim=-1)                              # routing.py:72-76
        wts, idx = torch.topk(probs, TOPK, dim=-1)
        wts = wts / wts.sum(dim=-1, keepdim=True)
        if RoutingResult is not None:
            counts = [int((idx == e).sum()) for e in range(E)]
            offs = [0]
            for c in counts[:-1]:
                offs.append(offs[-1] + c)
            rr = RoutingResult(idx, wts, counts, offs)
            xs, perm = create_expert_inputs(x, rr, E, TOPK)
        else:
            flat = idx.reshape(-1)
            order = torch.argsort(flat, stable=True)
            xs = [x[(order[flat[order] == e]) // TOPK] for e in range(E)]
            perm = torch.argsort(order)
        ys = []
        for e in range(E):
            a = xs[e]
            gte = a @ dequantize_weights(*w1[e]).T
            up = a @ dequantize_weights(*w3[e]).T
            h = torch.nn.functional.silu(gte) * up
            ys.append(h @ dequantize_weights(*w2[e]).T)
        if RoutingResult is not None:
            return combine_expert_outputs(ys, rr, perm, TOPK)
        y = torch.cat(ys)[perm].reshape(tokens, TOPK, -1)
        return (y * wts[:, :, None]).sum(1)

    step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return tokens / dt, torch.get_num_threads(), kind, (
        f"{steps} x the Mixtral-size layer (8 experts x 3 projections dequantised per call + matmul) on {tokens} tokens, "
        f"{'the reference itself (oracle/_ref/reference)' if kind == 'reference' else 'numpy port of the reference'}; "
        f"{dt:.2f} s per call")


def moe_decode(torch, pkg, dev, hbm_peak):
    """BASELINE.json configs[3], decode: T = 1, 4, 8, 16 tokens through QuantizedMoE.forward_routed (random routing): the
    decode call (route + grouped gate/up GEMV + grouped down GEMV + combine; exact-integer kernel below 1.5 rows per
    expert on average, mid-batch kernel above).  HBM-bound: algorithmic bytes = (#distinct experts hit) x 88 MB of packed
    INT4 weights."""
    layer = build_local_moe(torch, pkg, list(range(E)), dev)
    out = []
    for T in (1, 4, 8, 16):
        x, logits = make_inputs(torch, 0, T, "random", dev)
        ms = time_steps(torch, None, dev, lambda: layer.forward_routed(x, logits, top_k=TOPK), 50, 5)
        hit = int(torch.unique(pkg.route(logits, TOPK).expert_indices).numel())
        nb = hit * EXPERT_BYTES
        out.append({"T": T, "experts_hit": hit, "ms": round(ms, 4), "tokens_per_s": round(T / (ms * 1e-3), 1),
                    "GBps": round(nb / ms / 1e6, 1), "frac_hbm_peak": round(nb / ms / 1e6 / hbm_peak, 4)})
    return out


def moe_prefill_sweep(torch, pkg, dev, tf_peak):
    """BASELINE.json configs[3], prefill: T = 512, 2048, 8192 tokens through QuantizedMoE.forward_routed on one GPU
    (random routing; T = 16384 is the `moe` headline).  Tensor-bound: FLOPs = T x 704,643,072."""
    layer = build_local_moe(torch, pkg, list(range(E)), dev)
    out = []
    for T in (512, 2048, 8192):
        x, logits = make_inputs(torch, 0, T, "random", dev)
        ms = time_steps(torch, None, dev, lambda: layer.forward_routed(x, logits, top_k=TOPK), 10, 3)
        tf = T * FLOPS_PER_TOKEN / (ms * 1e-3) / 1e12
        out.append({"T": T, "ms": round(ms, 4), "tokens_per_s": round(T / (ms * 1e-3), 1), "TFLOPs": round(tf, 1),
                    "frac_bf16_sustained_peak": round(tf / tf_peak, 4)})
    return out


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
    t_loc = T_GLOBAL // world
    full = build_local_moe(torch, pkg, list(range(E)), dev) if world > 1 else None     # the unsharded layer: parity check only

    def make_layer(replicated):
        mine = pkg.local_expert_list(E, rank, world, replicated)
        moe = full if (full is not None and len(mine) == E) else build_local_moe(torch, pkg, mine, dev)
        return pkg.ExpertParallelMoE(moe, E, TOPK, replicated=replicated)

    results, parity = {}, {}
    skew_rep = (0, 1)                                # the two hottest experts of the skewed recipe
    plans = [("random", ()), ("skewed", ())] + ([("skewed", skew_rep)] if world > 1 else [])
    clocks_summary = None
    layer_cache = {}
    for routing, replicated in plans:
        if replicated not in layer_cache:
            layer_cache[replicated] = make_layer(replicated)
        layer = layer_cache[replicated]
        x, logits = make_inputs(torch, rank, t_loc, routing, dev)
        out = layer(x, logits)
        assert out.shape == (t_loc, D) and bool(torch.isfinite(out).all())
        key = routing + ("+replicated" if replicated else "")
        if world > 1:
            # ---- parity of the CUDA + NCCL path: every rank, all its tokens, against the unsharded layer
            xs_all = [make_inputs(torch, r, t_loc, routing, dev) for r in range(world)]
            ref = full.forward_routed(torch.cat([a for a, _ in xs_all]), torch.cat([b for _, b in xs_all]), top_k=TOPK)
            diff = float((out - ref[rank * t_loc:(rank + 1) * t_loc]).abs().max())
            scale = float(ref.abs().max())
            t = torch.tensor([diff], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            parity[key] = {"max_abs_diff_vs_unsharded": float(t.item()), "ref_abs_max": scale}
            assert float(t.item()) <= 2e-3 * scale, f"EP output differs from the unsharded layer: {float(t.item())} (|y|max {scale})"
            del xs_all, ref
        if rank == 0 and not args.no_cpu and routing == "random":
            src = full if world > 1 else layer.local_moe
            sel, oref = oracle_check(torch, src, x, logits)
            err = float(np.abs(out[sel].double().cpu().numpy() - oref).max())
            parity.setdefault(key, {})["max_abs_err_vs_f64_oracle_64_tokens"] = err
            parity[key]["oracle_abs_max"] = float(np.abs(oref).max())
            assert err <= 2e-2 * float(np.abs(oref).max()), f"MoE layer differs from the oracle: {err}"
        with ClockSampler(local_rank) as clocks:
            ms = time_steps(torch, dist, dev, lambda: layer(x, logits), args.steps if routing == "random" else max(3, args.steps // 2), args.warmup)
        results[key] = {"ms_per_step": ms, "tokens_per_s": T_GLOBAL / (ms * 1e-3), "ep": dict(layer.last_stats)}
        if world > 1:                              # one more call with CUDA events around the phases (not part of the timing)
            layer.profile = True
            layer(x, logits)
            layer.profile = False
            results[key]["ep"] = dict(layer.last_stats)
        if routing == "random":
            clocks_summary = clocks.summary()
            # end to end: activations and router logits start in pinned host memory, the result returns to the host
            xh, lh = x.cpu().pin_memory(), logits.cpu().pin_memory()
            oh = torch.empty(t_loc, D).pin_memory()

            def e2e_step():
                oh.copy_(layer(xh.to(dev, non_blocking=True), lh.to(dev, non_blocking=True)), non_blocking=True)

            e2e_ms = time_steps(torch, dist, dev, e2e_step, max(3, min(args.steps, 20)), 3)
    if dist is not None:
        dist.barrier()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peaks, peak_src = measured_peaks()
    peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    ms = results["random"]["ms_per_step"]
    tflops = FLOPS_PER_TOKEN * T_GLOBAL / (ms * 1e-3) / 1e12
    # kernels per layer call: top-k, 3 permutation kernels, gather, 2 x (activation prep, grouped tcgen05 GEMM, fix-up), combine
    # (+ the plan kernel; NCCL's own kernels are not counted)
    launches_per_step = 12 + (1 if world > 1 else 0)
    line = {
        "metric": "mixtral_moe_int4_layer_tokens_per_s", "value": T_GLOBAL / (ms * 1e-3), "unit": "tokens/s", "n_gpus": world,
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
        "clocks": clocks_summary,
        "ep": results["random"]["ep"],
        "parity": parity,
        "skewed": {k: {"tokens_per_s": v["tokens_per_s"], "ms_per_step": v["ms_per_step"], "ep": v["ep"]} for k, v in results.items() if k != "random"},
    }
    if world > 1:
        line["skewed"]["note"] = ("the reference's default routing recipe (47 % of the assignments to expert 0); "
                                  "'+replicated': experts 0 and 1 held by every rank (hot-expert replication)")
    # the same sub-object the N = 1 (GEMV-headline) line carries: `moe.value` is one consistent series over N = 1, 2, 4, 8
    line["moe"] = {k: line[k] for k in ("metric", "value", "unit", "n_gpus", "ms_per_step", "dtype", "roofline")}
    line["moe"]["workload"] = line["config"]["workload"]
    if not args.no_cpu:
        v, threads, kind, sample = cpu_moe_reference()
        line["cpu_baseline"] = {"value": v, "unit": "tokens/s", "cores": threads, "kind": kind, "sample": sample}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
