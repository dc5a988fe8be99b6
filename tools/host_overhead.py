"""Host cost per eager call of the decode path (M = 1, 4096 -> 11008): where the microseconds between two launches go.
Loops of 3000 calls without a synchronisation inside (the host is the bottleneck: a launch takes 5.6 us on the GPU)."""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200q_pkg import pkg
_lib = pkg._lib
lib = _lib.load()
dev = torch.device("cuda", 0)
K, N, M = 4096, 11008, 1
lin = torch.nn.Linear(K, N, bias=False).cuda()
ql = pkg.QuantizedLinear.from_linear(lin)
x = torch.randn(M, K, device=dev)
ext = _lib.torch_ext()
n = 3000


def loop(fn, name):
    for _ in range(200): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name:60s} host {1e6 * (t1 - t0) / n:6.2f} us/call   (+ drain {1e6 * (t2 - t1) / n:5.2f})")


loop(lambda: ql(x), "QuantizedLinear.__call__")
loop(lambda: ql._forward_cuda(x), "QuantizedLinear._forward_cuda")
if ext is not None:
    p, s, z = ql.packed_weights, ql.scales, ql.zero_points
    loop(lambda: ext.linear_forward(x, p, s, z, None, None, 1, None), "b200q_torch.linear_forward (compiled binding)")
    loop(lambda: ext.forward(x, p, s, z), "b200q_torch.forward (reference signature)")
y = torch.empty(M, N, device=dev)
ws = torch.zeros(1 << 20, dtype=torch.uint8, device=dev)
sp = torch.cuda.current_stream(dev).cuda_stream
args = (x.data_ptr(), 0, ql.packed_weights.data_ptr(), ql.scales.data_ptr(), ql.zero_points.data_ptr(), y.data_ptr(), 0, M, N, K,
        ws.data_ptr(), ws.numel(), 1, sp, None, 0)
f = lib.b200q_linear_fwd_next
loop(lambda: f(*args), "b200q_linear_fwd_next via ctypes, preallocated output")
loop(lambda: torch.empty(M, N, device=dev), "torch.empty alone")
loop(lambda: None, "empty lambda")
