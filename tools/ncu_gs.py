"""ncu target: launches of the mid-batch decode kernel with group-wise scales (GS instances of gemv_hm.cu) over a
24-layer pool (no graph), the bench's configuration: static-weights flag and next-layer hint.  M K N G from argv."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200q_pkg import pkg
_lib = pkg._lib
lib = _lib.load()
M, K, N, G = [int(a) for a in (sys.argv[1:5] + ["1", "4096", "11008", "128"][len(sys.argv) - 1:])]
dev = torch.device("cuda", 0)
layers = []
for i in range(24):
    g = torch.Generator(device=dev); g.manual_seed(i)
    layers.append((torch.randint(0, 256, (N, K // 2), generator=g, device=dev, dtype=torch.uint8),
                   torch.rand(N, K // G, generator=g, device=dev) * 0.01 + 0.001,
                   torch.randint(0, 16, (N, K // G), generator=g, device=dev).float()))
x = torch.randn(M, K, device=dev); y = torch.empty(M, N, device=dev)
sp = torch.cuda.current_stream(dev).cuda_stream
for r in range(2):
    for i, (p, s, z) in enumerate(layers):
        nxt = layers[(i + 1) % 24][0]
        _lib.check(lib.b200q_linear_groupwise_bias_fwd(x.data_ptr(), 0, p.data_ptr(), s.data_ptr(), z.data_ptr(), None, G, y.data_ptr(), 0,
                                                       M, N, K, 1, sp, nxt.data_ptr(), nxt.numel()), "fwd")
torch.cuda.synchronize()
print("ok")
