"""Per-phase SM-clock timestamps of the decode GEMV (bench-only instrumentation, gemv_debug bit 3)."""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200q_pkg import pkg
_lib = pkg._lib; lib = _lib.load()
lib.b200q_debug_read_prof.argtypes = [ctypes.c_void_p]
dev = torch.device("cuda", 0)
names = ["start", "bar-init", "issued", "pdl_wait", "amax", "-", "bf-loaded", "main-done", "red-bar", "epilogue"]
for (K, N, M) in [(4096, 11008, 1), (4096, 11008, 4), (11008, 4096, 1)]:
    layers = []
    for i in range(12):
        g = torch.Generator(device=dev); g.manual_seed(i)
        layers.append((torch.randint(0, 256, (N, K // 2), generator=g, device=dev, dtype=torch.uint8),
                       torch.rand(N, device=dev) * 0.01, torch.randint(0, 16, (N,), device=dev).float()))
    x = torch.randn(M, K, device=dev); y = torch.empty(M, N, device=dev)
    ws = torch.zeros(max(lib.b200q_linear_ws_bytes(M, N, K), 16), dtype=torch.uint8, device=dev)
    _lib.tune("force_path", 2); _lib.tune("gemv_debug", 8); _lib.tune("gemv_pf", 0)
    _lib.tune("gemv_xprep", int(os.environ.get("XPREP", "0")))
    sp = torch.cuda.current_stream().cuda_stream
    for rep in range(3):
        for (p, s, z) in layers:
            _lib.check(lib.b200q_linear_fwd(x.data_ptr(), 0, p.data_ptr(), s.data_ptr(), z.data_ptr(), y.data_ptr(), 0,
                                            M, N, K, ws.data_ptr(), ws.numel(), 1, sp), "fwd")
    torch.cuda.synchronize()
    buf = np.zeros(256 * 16, dtype=np.int64)
    _lib.check(lib.b200q_debug_read_prof(buf.ctypes.data), "read")
    t = buf.reshape(256, 16)[:148, :10]
    rel = t - t[:, :1]
    print(f"K={K} N={N} M={M}: cycles since CTA start (median / max over 148 CTAs)")
    for i, n in enumerate(names):
        print(f"   {n:10s} {int(np.median(rel[:, i])):7d} {int(rel[:, i].max()):7d}")
    _lib.tune("gemv_debug", -1)
