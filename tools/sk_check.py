"""Stream-K correctness sweep of the tcgen05 GEMM: shapes with few, long tiles (a tile cut between many CTAs), both
activation dtypes, default dispatch and forced token-tile heights; max error relative to the largest output vs float64."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200q_pkg import pkg
_lib = pkg._lib
dev = torch.device("cuda", 0)
shapes = [(512, 11008, 4096), (4096, 8192, 128), (2048, 11008, 256), (4096, 11008, 128), (4096, 14336, 128), (1024, 14336, 384), (300, 11008, 1000)]
for dt in (torch.bfloat16, torch.float32):
    for (M, K, N) in shapes:
        g = torch.Generator(device=dev); g.manual_seed(M + K + N)
        p = torch.randint(0, 256, (N, K // 2), generator=g, device=dev, dtype=torch.uint8)
        s = torch.rand(N, device=dev) * 0.01 + 0.001
        z = torch.randint(0, 16, (N,), device=dev).float()
        x = torch.randn(M, K, device=dev).to(dt)
        ref = x.double() @ pkg.dequantize_weights(p, s, z).double().T
        out = {"x": str(dt).split(".")[1], "M": M, "K": K, "N": N}
        for name, sk, bn in (("default", -1, -1), ("sk_bn256", 1, 256), ("sk_bn192", 1, 192), ("no_sk", 0, -1)):
            _lib.tune("gemm_sk", sk); _lib.tune("gemm_bn", bn); _lib.tune("force_path", 3)
            errs = []
            for rep in range(3):
                y = _lib.linear_fwd(x, p, s, z, out_dtype=torch.float32)
                errs.append(float((y.double() - ref).abs().max() / ref.abs().max()))
            out[name] = float("%.2e" % max(errs))
        _lib.tune("gemm_sk", -1); _lib.tune("gemm_bn", -1); _lib.tune("force_path", -1)
        print(json.dumps(out), flush=True)
