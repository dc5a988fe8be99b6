"""Group-wise scales (G columns per scale / zero point) on the decode path: us per launch of b200q_linear_groupwise_fwd over
a 24-layer pool (CUDA graph), mid-batch decode kernel (default) against the reference-speed SIMT kernel (force_path 1)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200q_pkg import pkg
_lib = pkg._lib
lib = _lib.load()
dev = torch.device("cuda", 0)


def measure(M, K, N, G, force):
    layers = []
    for i in range(24):
        g = torch.Generator(device=dev); g.manual_seed(i)
        layers.append((torch.randint(0, 256, (N, K // 2), generator=g, device=dev, dtype=torch.uint8),
                       torch.rand(N, K // G, generator=g, device=dev) * 0.01 + 0.001, torch.randint(0, 16, (N, K // G), generator=g, device=dev).float()))
    x = torch.randn(M, K, device=dev); y = torch.empty(M, N, device=dev)
    _lib.tune("force_path", force)
    def launch_all(sp):
        for i, (p, s, z) in enumerate(layers):
            nxt = layers[(i + 1) % len(layers)][0]
            _lib.check(lib.b200q_linear_groupwise_bias_fwd(x.data_ptr(), 0, p.data_ptr(), s.data_ptr(), z.data_ptr(), None, G, y.data_ptr(), 0, M, N, K,
                                                           1, sp, nxt.data_ptr(), nxt.numel()), "fwd")
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
    _lib.tune("force_path", -1)
    return e0.elapsed_time(e1) * 1e3 / (10 * 24)


if __name__ == "__main__":
    for (K, N) in ((4096, 11008), (11008, 4096)):
        for G in (128, 256):
            for M in (1, 4, 8, 16):
                a, b = measure(M, K, N, G, -1), measure(M, K, N, G, 1)
                nb = N * K // 2 + 8 * N * (K // G) + 4 * M * K + 4 * M * N
                print(json.dumps({"K": K, "N": N, "G": G, "M": M, "decode_kernel_us": round(a, 2), "simt_kernel_us": round(b, 2),
                                  "GBps": round(nb / a / 1e3, 1)}), flush=True)
