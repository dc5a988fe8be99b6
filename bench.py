#!/usr/bin/env python
"""bench.py -- the round-end measurement contract.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload gemv|moe]

N = 1 (default workload "gemv"): BASELINE.json configs[1], the Llama-7B up-projection decode GEMV
  (4096 -> 11008, INT4, M = 1, fp32 activations).  One STEP = one fused dequantize-linear launch on
  each layer of a pool of POOL independently-seeded layers (POOL x 22.5 MB >> the 126 MB L2, so
  every launch streams its weights from HBM), replayed as one CUDA graph.
  value   = algorithmic GB/s (B(M) = N*K/2 + 8N + 4MK + 4MN per launch) with all
            inputs resident in HBM, CUDA events on the launching stream;
  e2e     = the same metric through the public API (QuantizedLinear.forward) with the activations in
            pinned HOST memory: H2D copy of x and D2H copy of y inside the timed region, every launch;
  roofline / cpu_baseline: see DESIGN.md "Measurement".
N > 1 (default workload "moe"): BASELINE.json configs[4], the Mixtral-8x7B INT4 MoE layer,
  expert-parallel over N GPUs (one process per GPU, NCCL all-to-all dispatch / combine).

--impl reference times the CPU restatement of the reference path (oracle/int4_oracle.py: dequantize
to fp32 + matmul, python/quantize.py:127-202) on the host cores, one layer per step.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K_IN, N_OUT = 4096, 11008        # Llama-7B MLP up / gate projection
POOL = 24                        # layers in the rotating weight pool (24 x 22.5 MB = 541 MB)


def gemv_bytes(M, N, K):
    return N * K // 2 + 8 * N + 4 * M * K + 4 * M * N


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake_slowdown": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def oracle_module():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import int4_oracle
    return int4_oracle


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(n) if n else 1
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_gemv(steps, warmup, M=1):
    """The reference CPU path (dequantize_weights to a full fp32 matrix, then F.linear:
    python/quantize.py:127-202) timed on the host cores, one 4096->11008 layer per step.  Uses the C
    restatement (oracle/oracle.c, pthreads over all online cores) when it is built, else the numpy one.
    Returns (GB/s, seconds per layer, threads, description)."""
    import numpy as np
    oracle = oracle_module()
    rng = np.random.default_rng(42)
    packed = rng.integers(0, 256, size=(N_OUT, K_IN // 2), dtype=np.uint8)
    scales = (rng.random(N_OUT, dtype=np.float32) * 0.004 + 0.002).astype(np.float32)
    zps = rng.integers(0, 16, size=N_OUT).astype(np.float32)
    x = rng.standard_normal((M, K_IN), dtype=np.float32)
    try:
        import c_oracle
        lin = c_oracle.Linear(packed, scales, zps)
        fn, threads, what = (lambda: lin(x)), c_oracle.max_threads(), "C restatement (oracle/oracle.c, pthreads)"
    except Exception:
        fn = lambda: oracle.reference_quantized_linear(x, packed, scales, zps)
        threads, what = blas_threads(), "numpy restatement (oracle/int4_oracle.py)"
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = (time.perf_counter() - t0) / steps
    return gemv_bytes(M, N_OUT, K_IN) / dt / 1e9, dt, threads, what


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 200))
    warmup = max(1, min(args.warmup, 5))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    if (args.workload or ("gemv" if world == 1 else "moe")) == "moe":
        # same metric / config as the product arm at this N: the Mixtral layer, composed from the reference's
        # primitives on the host cores (bounded sample: one expert's three projections on 8 rows, scaled per token)
        from bench_moe import cpu_moe_baseline, T_GLOBAL, E as MOE_E
        best, cores, sample = 0.0, 1, ""
        t0 = time.perf_counter()
        for _ in range(min(steps, 5)):
            v, cores, sample = cpu_moe_baseline()
            best = max(best, v)
        dt = (time.perf_counter() - t0) / min(steps, 5)
        line = {
            "impl": "reference", "metric": "mixtral_moe_int4_layer_tokens_per_s", "value": best, "unit": "tokens/s",
            "n_gpus": args.gpus, "steps": min(steps, 5), "warmup": 0, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Mixtral-8x7B INT4 MoE layer (8 experts, top-2, d=4096, ffn=14336), "
                                   f"{T_GLOBAL} tokens/step, random routing" + (", expert-parallel" if world > 1 else ""),
                       "arm": "the reference's CPU path composed per expert: dequantize_weights + matmul (C restatement, all host cores)",
                       "tokens_per_step": T_GLOBAL, "experts_per_rank": MOE_E // max(world, 1)},
            "cpu_baseline": {"value": best, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": best, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return
    gbs, dt, cores, what = cpu_reference_gemv(steps, warmup)
    line = {
        "impl": "reference", "metric": "int4_gemv_hbm_gbps", "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"Llama-7B MLP INT4 decode GEMV M=1 ({K_IN}->{N_OUT}), configs[1]",
                   "arm": "the reference's CPU path: dequantize_weights + F.linear (C restatement, all host cores)",
                   "M": 1, "K": K_IN, "N": N_OUT},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} x one {K_IN}->{N_OUT} layer forward, M=1: {what} of "
                                   "python/quantize.py dequantize_weights + F.linear"},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def make_pool(torch, pkg, n_layers, device):
    """POOL independently-seeded Llama up-projection layers, quantised on the GPU (bit-exact with
    the reference's quantize_weights)."""
    layers = []
    for i in range(n_layers):
        g = torch.Generator(device=device)
        g.manual_seed(42 + i)
        # nn.Linear's default init range for in_features=4096: U(-1/64, 1/64)
        w = (torch.rand(N_OUT, K_IN, generator=g, device=device) * 2 - 1) / 64.0
        layers.append(pkg.quantize_weights(w))
        del w
    return layers


def run_gemv(args):
    import numpy as np
    import torch
    from b200q_pkg import pkg
    _lib = pkg._lib
    lib = _lib.load()

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
    M = args.m
    layers = make_pool(torch, pkg, POOL, dev)
    x = torch.randn(M, K_IN, device=dev)
    y = torch.empty(M, N_OUT, device=dev)
    ws_bytes = lib.b200q_linear_ws_bytes(M, N_OUT, K_IN)
    ws = torch.zeros(max(ws_bytes, 16), dtype=torch.uint8, device=dev)

    # a decode loop knows which fused linear follows: each call names the next layer's packed weights so the
    # kernel can pull them into L2 behind its own weight stream (b200q_linear_fwd_next; --no-hint turns it off)
    def launch_all(stream_ptr):
        for i, (p, s, z) in enumerate(layers):
            nxt = None if args.no_hint else layers[(i + 1) % len(layers)][0]
            _lib.check(lib.b200q_linear_fwd_next(x.data_ptr(), _lib.F32, p.data_ptr(), s.data_ptr(), z.data_ptr(),
                                                 y.data_ptr(), _lib.F32, M, N_OUT, K_IN, ws.data_ptr(), ws.numel(),
                                                 _lib.FLAG_STATIC_WEIGHTS, stream_ptr,
                                                 nxt.data_ptr() if nxt is not None else None,
                                                 nxt.numel() if nxt is not None else 0), "b200q_linear_fwd_next")

    # correctness spot check before timing (oracle used as the checker only)
    oracle = oracle_module()
    p0, s0, z0 = (t.cpu().numpy() for t in layers[0])
    stream = torch.cuda.current_stream(dev)
    _lib.check(lib.b200q_linear_fwd(x.data_ptr(), _lib.F32, layers[0][0].data_ptr(), layers[0][1].data_ptr(),
                                    layers[0][2].data_ptr(), y.data_ptr(), _lib.F32, M, N_OUT, K_IN, ws.data_ptr(),
                                    ws.numel(), 0, stream.cuda_stream), "b200q_linear_fwd")
    rows = np.arange(0, N_OUT, 43)
    ref = oracle.reference_quantized_linear(x.cpu().numpy(), p0[rows], s0[rows], z0[rows], acc=np.float64)
    err = float(np.abs(y.cpu().numpy()[:, rows] - ref).max())
    assert err < 1e-3, f"GEMV parity check failed: max abs err {err}"

    graph = None
    if not args.no_graph:
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            launch_all(side.cuda_stream)       # warm-up on the capture stream (sets func attributes)
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            launch_all(torch.cuda.current_stream(dev).cuda_stream)

    def step():
        if graph is not None:
            graph.replay()
        else:
            launch_all(torch.cuda.current_stream(dev).cuda_stream)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = args.steps * POOL            # fused dequantize-linear calls; each is ONE kernel (gemv_res_kernel)
    per_launch_s = ms * 1e-3 / launches
    bytes_per_launch = gemv_bytes(M, N_OUT, K_IN)
    gbs_per_gpu = bytes_per_launch / per_launch_s / 1e9
    value = gbs_per_gpu * world

    # ---- end to end through the public API with HOST buffers: QuantizedLinear.forward_host (one C-ABI call,
    # b200q_linear_fwd_host, enqueues the H2D copy of x, the kernels and the D2H copy of y).  The 24 calls of a
    # step are captured in a CUDA graph, as a decode loop would; every replay copies x from and y to pinned memory.
    mods = []
    for (p, s, z) in layers:
        m = pkg.QuantizedLinear(K_IN, N_OUT)
        m.packed_weights, m.scales, m.zero_points = p, s, z
        m._weights_settled = True
        mods.append(m)
    xh = torch.randn(M, K_IN).pin_memory()
    yhs = [torch.empty(M, N_OUT).pin_memory() for _ in mods]

    def e2e_all():
        for m, yh in zip(mods, yhs):
            m.forward_host(xh, out=yh)

    e2e_graph = None
    if not args.no_graph:
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            e2e_all()
        torch.cuda.current_stream(dev).wait_stream(side)
        e2e_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(e2e_graph):
            e2e_all()

    def e2e_step():
        if e2e_graph is not None:
            e2e_graph.replay()
        else:
            e2e_all()

    e2e_steps = max(3, min(args.steps, 100))
    for _ in range(3):
        e2e_step()
    torch.cuda.synchronize(dev)
    ref0 = oracle.reference_quantized_linear(xh.numpy(), p0[rows], s0[rows], z0[rows], acc=np.float64)
    e2e_err = float(np.abs(yhs[0].numpy()[:, rows] - ref0).max())
    assert e2e_err < 1e-3, f"e2e parity check failed: max abs err {e2e_err}"
    if dist is not None:
        dist.barrier()
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms2 = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms2], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms2 = float(t.item())
    e2e_gbs = bytes_per_launch / (ms2 * 1e-3 / (e2e_steps * POOL)) / 1e9 * world

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peaks, peak_src = measured_peaks()
    peak = float(peaks["hbm_gbs"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "gemv_traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    cpu_gbs, cpu_dt, cpu_threads, cpu_what = cpu_reference_gemv(steps=20, warmup=2, M=M) if not args.no_cpu else (None, None, None, None)
    line = {
        "metric": "int4_gemv_hbm_gbps", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": f"Llama-7B MLP INT4 decode GEMV M={M} ({K_IN}->{N_OUT}), configs[1]", "M": M, "K": K_IN,
            "N": N_OUT, "launches_per_step": POOL, "next_layer_l2_hint": not args.no_hint,
            "l2": f"inputs larger than L2: weight pool of {POOL} layers = {POOL * N_OUT * K_IN // 2 / 1e6:.0f} MB "
                  "cycled round-robin (L2 is 126 MB)",
            "cuda_graph": graph is not None, "pdl": True,
            "parallelism": "replicas only" if world > 1 else "single GPU",
        },
        "us_per_launch": per_launch_s * 1e6,
        "roofline": {"bound": "hbm", "achieved": gbs_per_gpu, "peak": peak, "unit": "GB/s",
                     "frac": gbs_per_gpu / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": bytes_per_launch},
        "e2e": {"value": e2e_gbs, "unit": "GB/s", "h2d_bytes_per_step": POOL * M * K_IN * 4,
                "d2h_bytes_per_step": POOL * M * N_OUT * 4, "steps": e2e_steps,
                "api": "QuantizedLinear.forward_host -> b200q_linear_fwd_host (pinned host x and y, weights resident; x is pulled "
                       "over PCIe by a staging kernel and the GEMV epilogue stores y straight into the pinned buffer), "
                       "24 calls per step replayed as a CUDA graph", "us_per_call": ms2 * 1e3 / (e2e_steps * POOL)},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "parity": {"max_abs_err_vs_f64_oracle": err},
    }
    if cpu_gbs is not None:
        line["cpu_baseline"] = {"value": cpu_gbs, "unit": "GB/s", "cores": cpu_threads, "kind": "port",
                                "sample": f"20 x one {K_IN}->{N_OUT} layer forward, M={M}: {cpu_what} of "
                                          "dequantize_weights + F.linear", "ms_per_layer": cpu_dt * 1e3}
    # the second clause of BASELINE.json's metric (MoE layer tokens/s) at this GPU count, as a sub-object: the
    # 1-GPU point of the expert-parallel scaling curve that `--gpus 2/4/8` (default workload "moe") continues
    if world == 1 and not args.no_moe:
        try:
            del layers, mods, graph, e2e_graph
            torch.cuda.empty_cache()
            import io
            import contextlib
            from bench_moe import run_moe
            buf = io.StringIO()
            margs = argparse.Namespace(steps=20, warmup=3, no_cpu=True)
            with contextlib.redirect_stdout(buf):
                run_moe(margs)
            ml = json.loads(buf.getvalue().strip().splitlines()[-1])
            line["moe"] = {k: ml[k] for k in ("metric", "value", "unit", "n_gpus", "ms_per_step", "dtype", "roofline")}
            line["moe"]["workload"] = ml["config"]["workload"]
            line["moe"]["note"] = ("1-GPU base of the expert-parallel series: `bench.py --gpus 2|4|8` report this metric "
                                   "(same layer, same 16384 tokens per step, strong scaling) as their headline value")
        except Exception as e:      # the headline line must not depend on the extra measurement
            line["moe"] = {"error": repr(e)[:200]}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=["gemv", "moe"])
    ap.add_argument("--m", type=int, default=1, help="batch rows for the gemv workload (1..16)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-hint", action="store_true", help="gemv workload: do not name the next layer's weights (no L2 prefetch)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-moe", action="store_true", help="skip the MoE-layer sub-measurement of the N=1 gemv line")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    workload = args.workload or ("gemv" if args.gpus <= 1 and int(os.environ.get("WORLD_SIZE", "1")) <= 1 else "moe")
    if workload == "gemv":
        return run_gemv(args)
    from bench_moe import run_moe       # expert-parallel Mixtral layer
    return run_moe(args)


if __name__ == "__main__":
    main()
