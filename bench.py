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

--impl reference times the reference's own CPU path (the real reference staged under oracle/_ref by
oracle/make_ref.py: python/module.py forward = quantize.py dequantize_weights + F.linear; the oracle port only
when that staging is absent) on the host cores, one layer per step; for the MoE workload a bounded sample.
The N = 1 line also carries decode_sweep, groupwise_decode, prefill, ref_gpu_kernel, moe, moe_decode, moe_prefill.
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


REF_ROOT = os.path.join(ROOT, "oracle", "_ref", "reference")      # staged copy of the reference (oracle/make_ref.py)


def cpu_reference_gemv(steps, warmup, M=1):
    """The reference's CPU path timed on the host cores, one 4096->11008 layer per step: the REAL reference
    (oracle/_ref/reference: python/module.py QuantizedLinear.forward on a CPU tensor -> quantize.py dequantize_weights +
    F.linear, :127-202) when it is staged, else the C restatement (oracle/oracle.c).
    Returns (GB/s, seconds per layer, threads, kind, description)."""
    import numpy as np
    if os.path.isdir(os.path.join(REF_ROOT, "python")):
        import torch
        torch.set_num_threads(os.cpu_count() or 1)   # (torch.distributed.run exports OMP_NUM_THREADS=1)
        sys.path.insert(0, REF_ROOT)
        from python.module import QuantizedLinear as RefQuantizedLinear
        torch.manual_seed(42)
        ql = RefQuantizedLinear.from_linear(torch.nn.Linear(K_IN, N_OUT, bias=False))       # BASELINE.json configs[0]
        x = torch.randn(M, K_IN) if M > 1 else torch.randn(K_IN)
        fn = lambda: ql(x)
        threads, kind = torch.get_num_threads(), "reference"
        what = "the reference itself (python/module.py forward on CPU = quantize.py dequantize_weights + F.linear, torch CPU threads)"
    else:
        oracle = oracle_module()
        rng = np.random.default_rng(42)
        packed = rng.integers(0, 256, size=(N_OUT, K_IN // 2), dtype=np.uint8)
        scales = (rng.random(N_OUT, dtype=np.float32) * 0.004 + 0.002).astype(np.float32)
        zps = rng.integers(0, 16, size=N_OUT).astype(np.float32)
        x = rng.standard_normal((M, K_IN), dtype=np.float32)
        kind = "port"
        try:
            import c_oracle
            lin = c_oracle.Linear(packed, scales, zps)
            fn, threads, what = (lambda: lin(x)), c_oracle.max_threads(), "C restatement (oracle/oracle.c, pthreads) of dequantize_weights + F.linear"
        except Exception:
            fn = lambda: oracle.reference_quantized_linear(x, packed, scales, zps)
            threads, what = os.cpu_count() or 1, "numpy restatement (oracle/int4_oracle.py) of dequantize_weights + F.linear"
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = (time.perf_counter() - t0) / steps
    return gemv_bytes(M, N_OUT, K_IN) / dt / 1e9, dt, threads, kind, what


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 50))
    warmup = max(1, min(args.warmup, 3))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    if (args.workload or default_workload(args)) == "moe":
        # same metric / config as the product arm at this N: the Mixtral layer composed from the reference's own
        # primitives on the host cores; bounded sample = 512 tokens of the same layer (every call dequantises all 24
        # projections, as the reference's QuantizedMoE.forward does)
        from bench_moe import cpu_moe_reference, T_GLOBAL, E as MOE_E
        nsteps = max(1, min(steps, 3))
        t0 = time.perf_counter()
        v, cores, kind, sample = cpu_moe_reference(steps=nsteps, tokens=512)
        dt = (time.perf_counter() - t0) / (nsteps + 1)
        line = {
            "impl": "reference", "metric": "mixtral_moe_int4_layer_tokens_per_s", "value": v, "unit": "tokens/s",
            "n_gpus": args.gpus, "steps": nsteps, "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Mixtral-8x7B INT4 MoE layer (8 experts, top-2, d=4096, ffn=14336), "
                                   f"{T_GLOBAL} tokens/step, random routing" + (", expert-parallel" if world > 1 else ""),
                       "arm": "the reference's CPU path composed per expert (dequantize_weights + matmul, routing.py dispatch / combine); "
                              "each step is a bounded sample of the workload: 512 of its tokens",
                       "tokens_per_step": T_GLOBAL, "sample_tokens_per_step": 512, "experts_per_rank": MOE_E // max(world, 1)},
            "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return
    gbs, dt, cores, kind, what = cpu_reference_gemv(steps, warmup)
    line = {
        "impl": "reference", "metric": "int4_gemv_hbm_gbps", "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"Llama-7B MLP INT4 decode GEMV M=1 ({K_IN}->{N_OUT}), configs[1]",
                   "arm": what, "M": 1, "K": K_IN, "N": N_OUT},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": kind,
                         "sample": f"{steps} x one {K_IN}->{N_OUT} layer forward, M=1: {what}"},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def make_pool(torch, pkg, n_layers, device, K=K_IN, N=N_OUT):
    """POOL independently-seeded layers, quantised on the GPU (bit-exact with the reference's quantize_weights)."""
    layers = []
    for i in range(n_layers):
        g = torch.Generator(device=device)
        g.manual_seed(42 + i)
        # nn.Linear's default init range: U(-1/sqrt(K), 1/sqrt(K))
        w = (torch.rand(N, K, generator=g, device=device) * 2 - 1) / (K ** 0.5)
        layers.append(pkg.quantize_weights(w))
        del w
    return layers


def time_decode(torch, _lib, lib, layers, M, K, N, dev, hint=True, graph=True, reps=4, iters=20):
    """us per fused dequantize-linear launch over the layer pool (every launch streams its weights from HBM)."""
    x = torch.randn(M, K, device=dev)
    y = torch.empty(M, N, device=dev)
    ws = torch.zeros(max(lib.b200q_linear_ws_bytes(M, N, K), 16), dtype=torch.uint8, device=dev)

    def launch_all(sp):
        for _ in range(reps):
            for i, (p, s, z) in enumerate(layers):
                nxt = layers[(i + 1) % len(layers)][0] if hint else None
                _lib.check(lib.b200q_linear_fwd_next(x.data_ptr(), _lib.F32, p.data_ptr(), s.data_ptr(), z.data_ptr(), y.data_ptr(),
                                                     _lib.F32, M, N, K, ws.data_ptr(), ws.numel(), _lib.FLAG_STATIC_WEIGHTS, sp,
                                                     nxt.data_ptr() if nxt is not None else None,
                                                     nxt.numel() if nxt is not None else 0), "b200q_linear_fwd_next")

    g = None
    if graph:
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            launch_all(side.cuda_stream)
        torch.cuda.current_stream(dev).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            launch_all(torch.cuda.current_stream(dev).cuda_stream)
    run = g.replay if g is not None else (lambda: launch_all(torch.cuda.current_stream(dev).cuda_stream))
    for _ in range(3):
        run()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) * 1e3 / (iters * reps * len(layers))


def decode_sweep(torch, pkg, dev, peak, first_pool=None):
    """BASELINE.json configs[1]: M = 1 .. 16 on both Llama-7B MLP shapes, fp32 activations, next-layer hint on."""
    _lib = pkg._lib
    lib = _lib.load()
    out = []
    for (K, N) in ((4096, 11008), (11008, 4096)):
        layers = first_pool if (first_pool is not None and (K, N) == (K_IN, N_OUT)) else make_pool(torch, pkg, POOL, dev, K, N)
        for M in (1, 2, 3, 4, 8, 12, 16):
            us = time_decode(torch, _lib, lib, layers, M, K, N, dev)
            nb = gemv_bytes(M, N, K)
            out.append({"K": K, "N": N, "M": M, "us_per_launch": round(us, 3), "GBps": round(nb / us / 1e3, 1),
                        "frac_hbm_peak": round(nb / us / 1e3 / peak, 4),
                        "kernel": "gemv_dec (exact-integer IMMA)" if M <= 2 else "gemv_hm (three-digit IMMA form, eight tokens per instruction)"})
        if layers is not first_pool:
            del layers
            torch.cuda.empty_cache()
    return out


def groupwise_decode(torch, pkg, dev, peak):
    """group-wise scales (one scale / zero point per 128 columns) on the decode path, 4096 -> 11008, fp32
    activations: us per b200q_linear_groupwise_bias_fwd call over the 24-layer pool (CUDA graph, static-weights flag and
    next-layer hint as in the per-row sweep)."""
    _lib = pkg._lib
    lib = _lib.load()
    K, N, G = K_IN, N_OUT, 128
    layers = []
    for i in range(POOL):
        g = torch.Generator(device=dev); g.manual_seed(100 + i)
        layers.append((torch.randint(0, 256, (N, K // 2), generator=g, device=dev, dtype=torch.uint8),
                       torch.rand(N, K // G, generator=g, device=dev) * 0.01 + 0.001,
                       torch.randint(0, 16, (N, K // G), generator=g, device=dev).float()))
    out = []
    for M in (1, 4, 8, 16):
        x = torch.randn(M, K, device=dev); y = torch.empty(M, N, device=dev)
        def launch_all(sp):
            for i, (p, s, z) in enumerate(layers):
                nxt = layers[(i + 1) % len(layers)][0]
                _lib.check(lib.b200q_linear_groupwise_bias_fwd(x.data_ptr(), 0, p.data_ptr(), s.data_ptr(), z.data_ptr(), None, G, y.data_ptr(), 0,
                                                               M, N, K, 1, sp, nxt.data_ptr(), nxt.numel()), "groupwise")
        side = torch.cuda.Stream(dev); side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            launch_all(side.cuda_stream)
        torch.cuda.current_stream(dev).wait_stream(side)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            launch_all(torch.cuda.current_stream(dev).cuda_stream)
        for _ in range(3): gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): gr.replay()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (10 * POOL)
        nb = N * K // 2 + 8 * N * (K // G) + 4 * M * K + 4 * M * N
        out.append({"K": K, "N": N, "group_size": G, "M": M, "us_per_launch": round(us, 3), "GBps": round(nb / us / 1e3, 1),
                    "frac_hbm_peak": round(nb / us / 1e3 / peak, 4)})
    return out


def prefill_sweep(torch, pkg, dev, peak_tf):
    """BASELINE.json configs[2]: M = 512 .. 4096 through the tcgen05 GEMM, bf16 and fp32 activations (the reference API's
    dtype; hi + lo split = twice the MMAs).  FLOPs = 2 M N K (dequantisation not counted); includes the activation
    preparation pass.  One layer, repeated calls: the 22.5 MB of weights stay in L2, the tensor pipe is the bound."""
    _lib = pkg._lib
    out = []
    for (K, N) in ((4096, 11008), (11008, 4096)):
        p, s, z = make_pool(torch, pkg, 1, dev, K, N)[0]
        for dt in (torch.bfloat16, torch.float32):
            for M in (256, 512, 1024, 2048, 4096):
                x = torch.randn(M, K, device=dev).to(dt)
                for _ in range(3):
                    _lib.linear_fwd(x, p, s, z)
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    _lib.linear_fwd(x, p, s, z)
                e1.record()
                torch.cuda.synchronize(dev)
                ms = e0.elapsed_time(e1) / 10
                tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
                out.append({"K": K, "N": N, "M": M, "x": str(dt).replace("torch.", ""), "ms": round(ms, 4), "TFLOPs": round(tf, 1),
                            "frac_bf16_peak": round(tf / peak_tf, 4)})
        del p, s, z
    return out


def ref_gpu_kernel(torch, dev, layers):
    """The reference's own CUDA kernel (csrc/quantized_linear_kernel.cu:90-279) built for sm_100 from its sources by
    oracle/make_ref.py (baseline/_ref/), timed on the same layer pool: us per launch at M = 1 and 16."""
    import glob
    import importlib.util
    so = glob.glob(os.path.join(ROOT, "baseline", "_ref", "fused_quant_linear_cuda*.so"))
    if not so:
        return {"unavailable": "baseline/_ref/fused_quant_linear_cuda*.so not built (python oracle/make_ref.py --cuda)"}
    try:
        spec = importlib.util.spec_from_file_location("fused_quant_linear_cuda", so[0])
        ext = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ext)
        out = {"what": "reference kernel rebuilt for sm_100 (TORCH_CUDA_ARCH_LIST=10.0), its pybind forward(), stream 0, 24-layer pool"}
        for M in (1, 16):
            x = torch.randn(M, K_IN, device=dev)
            for (p, s, z) in layers[:3]:
                ext.forward(x, p, s, z)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for (p, s, z) in layers:
                ext.forward(x, p, s, z)
            e1.record()
            torch.cuda.synchronize(dev)
            us = e0.elapsed_time(e1) * 1e3 / len(layers)
            out[f"M{M}_us_per_launch"] = round(us, 2)
            out[f"M{M}_GBps"] = round(gemv_bytes(M, N_OUT, K_IN) / us / 1e3, 1)
        return out
    except Exception as e:
        return {"unavailable": repr(e)[:200]}


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
    # kernel can pull them into L2 behind its own work (b200q_linear_fwd_next; --no-hint turns it off)
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
    launches = args.steps * POOL            # fused dequantize-linear calls; each is ONE kernel (gemv_dec_kernel)
    per_launch_s = ms * 1e-3 / launches
    bytes_per_launch = gemv_bytes(M, N_OUT, K_IN)
    gbs_per_gpu = bytes_per_launch / per_launch_s / 1e9
    value = gbs_per_gpu * world

    # ---- end to end through the public API with HOST buffers: QuantizedLinear.forward_host (one C-ABI call,
    # b200q_linear_fwd_host, enqueues the H2D copy of x, the kernels and the D2H copy of y).  Measured twice: the 24
    # calls of a step captured in a CUDA graph, as a decode loop would, and EAGER -- every call pays the Python /
    # ctypes cost of the API it names.
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

    def timed(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        t_ms = e0.elapsed_time(e1)
        if dist is not None:
            tt = torch.tensor([t_ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_ms = float(tt.item())
        return t_ms

    e2e_steps = max(3, min(args.steps, 100))
    ms2 = timed(e2e_graph.replay if e2e_graph is not None else e2e_all, e2e_steps)
    ref0 = oracle.reference_quantized_linear(xh.numpy(), p0[rows], s0[rows], z0[rows], acc=np.float64)
    e2e_err = float(np.abs(yhs[0].numpy()[:, rows] - ref0).max())
    assert e2e_err < 1e-3, f"e2e parity check failed: max abs err {e2e_err}"
    e2e_gbs = bytes_per_launch / (ms2 * 1e-3 / (e2e_steps * POOL)) / 1e9 * world
    eager_steps = max(3, min(args.steps, 20))
    ms3 = timed(e2e_all, eager_steps)                                       # eager, host buffers
    xd = torch.randn(M, K_IN, device=dev)
    ms4 = timed(lambda: [m(xd) for m in mods], eager_steps)               # eager, device buffers: QuantizedLinear.forward

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peaks, peak_src = measured_peaks()
    peak = float(peaks["hbm_gbs"])
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "gemv_traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                tj = json.load(f)
            traffic = tj.get("dram_bytes_per_launch")
            traffic_note = {k: tj[k] for k in tj if k != "dram_bytes_per_launch"}
        except Exception:
            traffic = None
    if not args.no_cpu:
        cpu_gbs, cpu_dt, cpu_threads, cpu_kind, cpu_what = cpu_reference_gemv(steps=20, warmup=2, M=M)
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
                     "frac": gbs_per_gpu / peak, "traffic": traffic, "traffic_capture": traffic_note, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": bytes_per_launch},
        "e2e": {"value": e2e_gbs, "unit": "GB/s", "h2d_bytes_per_step": POOL * M * K_IN * 4,
                "d2h_bytes_per_step": POOL * M * N_OUT * 4, "steps": e2e_steps,
                "api": "QuantizedLinear.forward_host -> b200q_linear_fwd_host (pinned host x and y, weights resident; x is pulled "
                       "over PCIe by a staging kernel and the GEMV epilogue stores y straight into the pinned buffer), "
                       "24 calls per step replayed as a CUDA graph", "us_per_call": ms2 * 1e3 / (e2e_steps * POOL),
                "eager_us_per_call": ms3 * 1e3 / (eager_steps * POOL),
                "eager_value": bytes_per_launch / (ms3 * 1e-3 / (eager_steps * POOL)) / 1e9 * world,
                "eager_note": "the same 24 forward_host calls per step issued from Python without a graph: includes the per-call "
                              "Python / ctypes cost of the API",
                "eager_device_us_per_call": ms4 * 1e3 / (eager_steps * POOL),
                "eager_device_note": "QuantizedLinear.forward on a device tensor, no graph (host overhead per call when it exceeds the kernel)"},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "parity": {"max_abs_err_vs_f64_oracle": err},
    }
    if not args.no_cpu:
        line["cpu_baseline"] = {"value": cpu_gbs, "unit": "GB/s", "cores": cpu_threads, "kind": cpu_kind,
                                "sample": f"20 x one {K_IN}->{N_OUT} layer forward, M={M}: {cpu_what}", "ms_per_layer": cpu_dt * 1e3}
    if world == 1 and not args.no_sweeps:
        # the other single-GPU configurations of BASELINE.json, driver-visible: configs[1] (all M, both shapes),
        # configs[2] (prefill), the same-box reference GPU kernel
        try:
            line["ref_gpu_kernel"] = ref_gpu_kernel(torch, dev, layers)
            line["decode_sweep"] = decode_sweep(torch, pkg, dev, peak, first_pool=layers)
            del mods, graph, e2e_graph
            line["groupwise_decode"] = groupwise_decode(torch, pkg, dev, peak)
            line["prefill"] = prefill_sweep(torch, pkg, dev, float(peaks["bf16_tflops"]))
        except Exception as e:
            line["sweep_error"] = repr(e)[:300]
    # the second clause of BASELINE.json's metric (MoE layer tokens/s) at this GPU count, as a sub-object: the
    # 1-GPU point of the expert-parallel scaling curve that `--gpus 2/4/8` (default workload "moe") continues
    if world == 1 and not args.no_moe:
        try:
            del layers
            torch.cuda.empty_cache()
            import io
            import contextlib
            from bench_moe import run_moe, moe_decode
            buf = io.StringIO()
            margs = argparse.Namespace(steps=20, warmup=3, no_cpu=args.no_cpu)
            with contextlib.redirect_stdout(buf):
                run_moe(margs)
            ml = json.loads(buf.getvalue().strip().splitlines()[-1])
            line["moe"] = {k: ml[k] for k in ("metric", "value", "unit", "n_gpus", "ms_per_step", "dtype", "roofline", "parity", "skewed") if k in ml}
            line["moe"]["workload"] = ml["config"]["workload"]
            line["moe"]["note"] = ("1-GPU base of the expert-parallel series: `bench.py --gpus 2|4|8` report this metric "
                                   "(same layer, same 16384 tokens per step, strong scaling) as their headline value")
            if "cpu_baseline" in ml:
                line["moe"]["cpu_baseline"] = ml["cpu_baseline"]
            line["moe_decode"] = moe_decode(torch, pkg, dev, peak)
            if not args.no_sweeps:
                from bench_moe import moe_prefill_sweep
                line["moe_prefill"] = moe_prefill_sweep(torch, pkg, dev, float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])))
        except Exception as e:      # the headline line must not depend on the extra measurement
            line["moe"] = {"error": repr(e)[:200]}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def default_workload(args):
    """N = 1 run directly: the GEMV line (with the MoE layer as a sub-object).  Under torch.distributed.run (any N,
    also 1) or with --gpus > 1: the MoE layer, so that the points of a scaling series share one metric."""
    under_launcher = "WORLD_SIZE" in os.environ or "TORCHELASTIC_RUN_ID" in os.environ
    return "moe" if (args.gpus > 1 or under_launcher) else "gemv"


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
    ap.add_argument("--no-sweeps", action="store_true", help="skip decode_sweep / prefill / ref_gpu_kernel of the N=1 gemv line")
    args = ap.parse_args()
    # The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner when the
    # library's own communicator comes up): send file descriptor 1 to stderr for the duration of the run and give
    # sys.stdout the real one back -- the only writer left on it is the final print of the line.
    sys.stdout.flush()
    real_out = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_out, "w", buffering=1)
    if args.impl == "reference":
        return run_reference(args)
    workload = args.workload or default_workload(args)
    if workload == "gemv":
        return run_gemv(args)
    from bench_moe import run_moe       # expert-parallel Mixtral layer
    return run_moe(args)


if __name__ == "__main__":
    main()
