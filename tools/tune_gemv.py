"""Sweep the decode-GEMV launch heuristics / ablations on a B200 and print us per launch.

    python tools/tune_gemv.py [--quick]

Not part of the product: a measurement aid whose output is summarised under profiles/."""
import argparse
import itertools
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200q_pkg import pkg  # noqa: E402

_lib = pkg._lib
lib = _lib.load()
KEYS = ["gemv_warps", "gemv_slabs", "gemv_stages", "gemv_pdl", "gemv_ctas", "gemv_debug", "force_path", "gemv_occ2", "gemv_pf", "gemv_xprep", "gemv_early", "gemv_res"]


def make_pool(N, K, n_layers, dev):
    layers = []
    for i in range(n_layers):
        g = torch.Generator(device=dev)
        g.manual_seed(i)
        p = torch.randint(0, 256, (N, K // 2), generator=g, device=dev, dtype=torch.uint8)
        s = torch.rand(N, generator=g, device=dev) * 0.01 + 0.001
        z = torch.randint(0, 16, (N,), generator=g, device=dev).float()
        layers.append((p, s, z))
    return layers


def measure(layers, M, N, K, dev, steps=100, graph=True, flags=_lib.FLAG_STATIC_WEIGHTS):
    x = torch.randn(M, K, device=dev)
    y = torch.empty(M, N, device=dev)
    ws = torch.zeros(max(lib.b200q_linear_ws_bytes(M, N, K), 16), dtype=torch.uint8, device=dev)

    use_next = os.environ.get("NEXT", "1") == "1"

    def launch_all(sp):
        for i in range(max(int(os.environ.get('GRAPH_LEN', '24')), len(layers))):               # always 24+ launches per graph replay
            p, s, z = layers[i % len(layers)]
            nxt = layers[(i + (0 if os.environ.get("NEXT_SELF") else 1)) % len(layers)][0]
            _lib.check(lib.b200q_linear_fwd_next(x.data_ptr(), 0, p.data_ptr(), s.data_ptr(), z.data_ptr(), y.data_ptr(), 0,
                                                 M, N, K, ws.data_ptr(), ws.numel(), flags, sp,
                                                 nxt.data_ptr() if use_next else None, nxt.numel() if use_next else 0), "linear_fwd")

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
    run = (lambda: g.replay()) if g is not None else (lambda: launch_all(torch.cuda.current_stream(dev).cuda_stream))
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (steps * max(int(os.environ.get('GRAPH_LEN', '24')), len(layers)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "tune_gemv.jsonl"))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    shapes = [(4096, 11008)] if args.quick else [(4096, 11008), (11008, 4096)]
    out = open(args.out, "w")
    for (K, N) in shapes:
        layers = make_pool(N, K, int(os.environ.get("POOL", "24")), dev)
        nbytes = lambda M: N * K // 2 + 8 * N + 4 * M * K + 4 * M * N
        configs = [({}, 1, True), ({}, 2, True), ({}, 4, True), ({}, 8, True),
                   ({"gemv_pf": 0}, 1, True), ({"gemv_early": 31}, 1, True), ({"gemv_early": 9}, 1, True),
                   ({"force_path": 2}, 1, True), ({"force_path": 2, "gemv_xprep": 1}, 1, True)]
        for tune, M, graph in configs:
            for k in KEYS:
                _lib.tune(k, -1)
            for k, v in tune.items():
                _lib.tune(k, v)
            try:
                us = measure(layers, M, N, K, dev, graph=graph)
                rec = {"K": K, "N": N, "M": M, "tune": tune, "graph": graph, "us": round(us, 3),
                       "GBps": round(nbytes(M) / us / 1e3, 1)}
            except RuntimeError as e:
                rec = {"K": K, "N": N, "M": M, "tune": tune, "graph": graph, "error": str(e)[:120]}
            print(json.dumps(rec), flush=True)
            out.write(json.dumps(rec) + "\n")
        del layers
    for k in KEYS:
        _lib.tune(k, -1)


if __name__ == "__main__":
    main()
