"""Wall-clock (globaltimer) gaps between consecutive decode-GEMV launches inside a CUDA graph."""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200q_pkg import pkg
_lib = pkg._lib; lib = _lib.load()
RES = os.environ.get("RES", "1") == "1"          # resident-slab kernel (gemv_res.cu) or the ring kernel (gemv.cu)
dbg_wall = lib.b200q_debug_wall_res if RES else lib.b200q_debug_wall
dbg_prof = lib.b200q_debug_read_prof_res if RES else lib.b200q_debug_read_prof
dbg_wall.argtypes = [ctypes.c_void_p, ctypes.c_int]
dbg_prof.argtypes = [ctypes.c_void_p]
dev = torch.device("cuda", 0)
K, N, M = int(os.environ.get("K", "4096")), int(os.environ.get("N", "11008")), int(os.environ.get("M", "1"))
layers = []
for i in range(24):
    g = torch.Generator(device=dev); g.manual_seed(i)
    layers.append((torch.randint(0, 256, (N, K // 2), generator=g, device=dev, dtype=torch.uint8),
                   torch.rand(N, device=dev) * 0.01, torch.randint(0, 16, (N,), device=dev).float()))
x = torch.randn(M, K, device=dev); y = torch.empty(M, N, device=dev)
ws = torch.zeros(max(lib.b200q_linear_ws_bytes(M, N, K), 16), dtype=torch.uint8, device=dev)
_lib.tune("force_path", 5 if RES else 2)
_lib.tune("gemv_debug", int(os.environ.get("DBG", "8")))
if os.environ.get("PDL"): _lib.tune("gemv_pdl", int(os.environ["PDL"]))
if os.environ.get("OCC2"): _lib.tune("gemv_occ2", 1)
_lib.tune("gemv_pf", int(os.environ.get("PF", "0")))
if os.environ.get("EARLY"): _lib.tune("gemv_early", int(os.environ["EARLY"]))
USE_NEXT = os.environ.get("PF", "0") != "0"
if os.environ.get("XPREP"): _lib.tune("gemv_xprep", int(os.environ["XPREP"]))
def launch_all(sp):
    for i, (p, s, z) in enumerate(layers):
        nxt = layers[(i + 1) % len(layers)][0]
        _lib.check(lib.b200q_linear_fwd_next(x.data_ptr(), 0, p.data_ptr(), s.data_ptr(), z.data_ptr(), y.data_ptr(), 0,
                                             M, N, K, ws.data_ptr(), ws.numel(), 1, sp,
                                             nxt.data_ptr() if USE_NEXT else None, nxt.numel() if USE_NEXT else 0), "fwd")
side = torch.cuda.Stream(dev); side.wait_stream(torch.cuda.current_stream(dev))
with torch.cuda.stream(side):
    launch_all(side.cuda_stream)
torch.cuda.current_stream(dev).wait_stream(side)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    launch_all(torch.cuda.current_stream(dev).cuda_stream)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
_lib.check(dbg_wall(None, 1), "reset")
g.replay(); torch.cuda.synchronize()
buf = np.zeros(64 * 4, dtype=np.uint64)
_lib.check(dbg_wall(buf.ctypes.data, 0), "read")
w = buf.reshape(64, 4).astype(np.int64)
w = w[w[:, 3] > 0]
w = w[np.argsort(w[:, 0])]
t0 = w[0, 0]
print("launch: first CTA start, last CTA start, first CTA end, last CTA end (ns from the first launch); gap to next first start")
for i in range(len(w)):
    gap = (w[i + 1, 0] - w[i, 3]) if i + 1 < len(w) else 0
    print(f"{i:2d}: {w[i,0]-t0:7d} {w[i,1]-t0:7d} {w[i,2]-t0:7d} {w[i,3]-t0:7d}   period {(w[i+1,0]-w[i,0]) if i+1<len(w) else 0:6d}  gap {gap:6d}")

if int(os.environ.get("DBG", "8")) & 16:
    pb = np.zeros(256 * 16, dtype=np.int64)
    _lib.check(dbg_prof(pb.ctypes.data), "read prof")
    t = pb.reshape(256, 16)[:148, :15]
    ref = w[-2, 3] if len(w) > 1 else t[:, 0].min()          # last CTA end of the previous launch
    names = ["start", "bar-init", "issued", "pdl_wait", "amax", "-", "bf-loaded", "main-done", "red-bar", "epilogue", "tile0", "tile1", "tile2", "tile3", "tile4"]
    print("last launch, ns relative to the LAST CTA end of the previous launch (min / median / max over CTAs)")
    for i, n in enumerate(names):
        if n == "-": continue
        c = t[:, i] - ref
        print(f"   {n:10s} {int(c.min()):7d} {int(np.median(c)):7d} {int(c.max()):7d}")
    life = t[:, 9] - t[:, 0]
    print(f"   lifetime   {int(life.min()):7d} {int(np.median(life)):7d} {int(life.max()):7d}")
