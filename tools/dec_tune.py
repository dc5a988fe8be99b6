"""Decode-kernel tuning sweep: us per launch (24-layer pool, CUDA graph, next-layer hint) for a list of
(M, K, N, {tuning key: value}) configurations.   python tools/dec_tune.py "<json list>"   (or the built-in list)"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200q_pkg import pkg
_lib = pkg._lib
lib = _lib.load()
dev = torch.device("cuda", 0)
KEYS = ["hm_min_m", "hm_max_m", "hm_i3", "hm_waves", "gemm_sk", "gemm_bn", "gemv_early", "gemv_pf", "gemv_slots", "gemv_bufs", "gemv_pdl", "gemv_ctas", "force_path"]
pools = {}


POOL_LAYERS = int(os.environ.get("POOL_LAYERS", "24"))


def pool(N, K):
    if (N, K) not in pools:
        pools.clear()
        out = []
        for i in range(POOL_LAYERS):
            g = torch.Generator(device=dev); g.manual_seed(i)
            out.append((torch.randint(0, 256, (N, K // 2), generator=g, device=dev, dtype=torch.uint8),
                        torch.rand(N, generator=g, device=dev) * 0.01 + 0.001, torch.randint(0, 16, (N,), generator=g, device=dev).float()))
        pools[(N, K)] = out
    return pools[(N, K)]


def measure(M, K, N, tune, hint=True, reps=8):
    layers = pool(N, K)
    for k in KEYS:
        try:
            _lib.tune(k, -1)
        except RuntimeError:
            pass                                            # an older build (B200Q_LIB) without the key
    for k, v in tune.items(): _lib.tune(k, v)
    XD = int(os.environ.get("XD", "0")); TD = [torch.float32, torch.float16, torch.bfloat16][XD]
    x = torch.randn(M, K, device=dev).to(TD); y = torch.empty(M, N, device=dev, dtype=TD)
    ws = torch.zeros(max(lib.b200q_linear_ws_bytes(M, N, K), 16), dtype=torch.uint8, device=dev)
    reps = reps * 24 // len(layers)
    def launch_all(sp):
        for r in range(reps):
            for i, (p, s, z) in enumerate(layers):
                nxt = layers[(i + 1) % len(layers)][0]
                _lib.check(lib.b200q_linear_fwd_next(x.data_ptr(), XD, p.data_ptr(), s.data_ptr(), z.data_ptr(), y.data_ptr(), XD, M, N, K,
                                                     ws.data_ptr(), ws.numel(), 1, sp, nxt.data_ptr() if hint else None, nxt.numel() if hint else 0), "fwd")
    side = torch.cuda.Stream(dev); side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        launch_all(side.cuda_stream)
    torch.cuda.current_stream(dev).wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        launch_all(torch.cuda.current_stream(dev).cuda_stream)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(10): g.replay()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / (10 * len(layers) * reps))
    for k in KEYS:
        try:
            _lib.tune(k, -1)
        except RuntimeError:
            pass                                            # an older build (B200Q_LIB) without the key
    return best


if __name__ == "__main__":
    cfgs = json.loads(sys.argv[1]) if len(sys.argv) > 1 else [
        [1, 4096, 11008, {}], [1, 4096, 11008, {"gemv_pf": 2}], [1, 4096, 11008, {"gemv_pf": 3}], [1, 4096, 11008, {"gemv_pf": 0}],
        [1, 4096, 11008, {"gemv_early": 5}], [1, 4096, 11008, {"gemv_early": 5, "gemv_pf": 2}], [1, 4096, 11008, {"gemv_early": 2}],
        [1, 4096, 11008, {"gemv_early": 2, "gemv_pf": 2}], [1, 4096, 11008, {"gemv_early": 1, "gemv_pf": 2}], [1, 4096, 11008, {"gemv_early": 91, "gemv_pf": 2}],
        [1, 4096, 11008, {"gemv_early": 0, "gemv_pf": 2}], [1, 4096, 11008, {"gemv_early": 90, "gemv_pf": 2}], [1, 4096, 11008, {"gemv_early": 93, "gemv_pf": 2}],
        [1, 4096, 11008, {"gemv_slots": 0}], [1, 4096, 11008, {"gemv_pdl": 0}],
        [2, 4096, 11008, {}], [2, 4096, 11008, {"gemv_pf": 2}], [4, 4096, 11008, {}], [4, 4096, 11008, {"gemv_pf": 2}],
        [1, 11008, 4096, {}], [1, 11008, 4096, {"gemv_pf": 2}], [1, 11008, 4096, {"gemv_early": 2, "gemv_pf": 2}], [1, 11008, 4096, {"gemv_early": 1, "gemv_pf": 2}],
        [1, 11008, 4096, {"gemv_early": 91, "gemv_pf": 2}]]
    for M, K, N, tune in cfgs:
        us = measure(M, K, N, tune)
        nb = N * K // 2 + 8 * N + 4 * M * K + 4 * M * N
        print(json.dumps({"M": M, "K": K, "N": N, "tune": tune, "us": round(us, 3), "GBps": round(nb / us / 1e3, 1)}), flush=True)
