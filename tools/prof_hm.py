"""Phase stamps of the fp16 HMMA decode kernel (gemv_hm.cu), -DB200Q_PROF build (make -C csrc prof).

    B200Q_LIB=<pkg>/libb200q_prof.so python tools/prof_hm.py [M] [K] [N]

Prints, for the LAST launch of a 24-launch graph replay: median over CTAs of every phase stamp relative to the CTA's
own start (ns, %globaltimer), the spread of CTA starts / ends, and the CUDA-event period per launch."""
import ctypes, json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200q_pkg import pkg
_lib = pkg._lib
lib = _lib.load()
lib.b200q_debug_read_prof_hm.argtypes = [ctypes.c_void_p]
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
N = int(sys.argv[3]) if len(sys.argv) > 3 else 11008
dev = torch.device("cuda", 0)
NAMES = ["start", "setup", "pdl wait", "x loaded", "B frags", "tile0", "tile1", "tile2", "tile3", "tile4", "loop end",
         "barrier", "fold pass0", "fold: sums", "end"]
layers = []
for i in range(24):
    g = torch.Generator(device=dev); g.manual_seed(i)
    layers.append((torch.randint(0, 256, (N, K // 2), generator=g, device=dev, dtype=torch.uint8),
                   torch.rand(N, generator=g, device=dev) * 0.01 + 0.001, torch.randint(0, 16, (N,), generator=g, device=dev).float()))
XD = int(os.environ.get("XD", "0")); TD = [torch.float32, torch.float16, torch.bfloat16][XD]
x = torch.randn(M, K, device=dev).to(TD); y = torch.empty(M, N, device=dev, dtype=TD)
ws = torch.zeros(1 << 20, dtype=torch.uint8, device=dev)
DBG = int(os.environ.get("DBG", "1"))
for dbg in (0, DBG):
    _lib.tune("gemv_debug", dbg); _lib.tune("force_path", 7)
    def launch_all(sp):
        for i, (p, s, z) in enumerate(layers):
            nxt = layers[(i + 1) % 24][0]
            _lib.check(lib.b200q_linear_fwd_next(x.data_ptr(), XD, p.data_ptr(), s.data_ptr(), z.data_ptr(), y.data_ptr(), XD, M, N, K,
                                                 ws.data_ptr(), ws.numel(), 1, sp, nxt.data_ptr(), nxt.numel()), "fwd")
    side = torch.cuda.Stream(dev); side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        launch_all(side.cuda_stream)
    torch.cuda.current_stream(dev).wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        launch_all(torch.cuda.current_stream(dev).cuda_stream)
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"M": M, "K": K, "N": N, "stamps": bool(dbg), "us_per_launch": round(e0.elapsed_time(e1) * 1e3 / (50 * 24), 3)}))
buf = np.zeros(256 * 16, dtype=np.int64)
_lib.check(lib.b200q_debug_read_prof_hm(buf.ctypes.data), "read")
grid = min(148, (N + 15) // 16)
st = buf.reshape(256, 16)[:grid]
rel = st - st[:, :1]
t0 = st[:, 0].min()
print(f"CTA starts: spread {int(st[:, 0].max() - t0)} ns; ends: first {int(st[:, 14].min() - t0)} last {int(st[:, 14].max() - t0)} ns after the first start")
for i, n in enumerate(NAMES):
    col = rel[:, i]
    if (st[:, i] > 0).all():
        print(f"  {i:2d} {n:14s} median {int(np.median(col)):6d}  min {int(col.min()):6d}  max {int(col.max()):6d}")
_lib.tune("gemv_debug", -1); _lib.tune("force_path", -1)
