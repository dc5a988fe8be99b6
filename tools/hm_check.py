"""fp16 HMMA decode kernel (force_path=7) against a float64 dequantize + matmul on the GPU, plus timing against the
default dispatch.   python tools/hm_check.py"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from b200q_pkg import pkg
_lib = pkg._lib
lib = _lib.load()
dev = torch.device("cuda", 0)
DT = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def ref64(x, p, s, z):
    lo = (p & 15).double(); hi = (p >> 4).double()
    q = torch.stack([lo, hi], dim=-1).reshape(p.shape[0], -1)
    w = (q - z.double()[:, None]) * s.double()[:, None]
    return x.double() @ w.T


def run(M, K, N, dtype, force, bias=None):
    g = torch.Generator(device=dev); g.manual_seed(M * 7 + K + N)
    p = torch.randint(0, 256, (N, K // 2), generator=g, device=dev, dtype=torch.uint8)
    s = torch.rand(N, generator=g, device=dev) * 0.01 + 0.001
    z = torch.randint(0, 16, (N,), generator=g, device=dev).float()
    x = torch.randn(M, K, generator=g, device=dev).to(dtype)
    y = torch.empty(M, N, device=dev, dtype=dtype)
    ws = torch.zeros(max(lib.b200q_linear_ws_bytes(M, N, K), 16), dtype=torch.uint8, device=dev)
    _lib.tune("force_path", force)
    rc = lib.b200q_linear_bias_fwd(x.data_ptr(), DT[dtype], p.data_ptr(), s.data_ptr(), z.data_ptr(), bias.data_ptr() if bias is not None else None,
                                   y.data_ptr(), DT[dtype], M, N, K, ws.data_ptr(), ws.numel(), 1, torch.cuda.current_stream().cuda_stream, None, 0)
    _lib.tune("force_path", -1)
    _lib.check(rc, "fwd")
    torch.cuda.synchronize()
    r = ref64(x, p, s, z)
    if bias is not None: r = r + bias.double()
    err = (y.double() - r).abs().max().item(); sc = r.abs().max().item()
    return err, sc


if __name__ == "__main__":
    worst = {}
    for (K, N) in [(4096, 11008), (11008, 4096), (4096, 4096), (1024, 1000), (256, 100), (2048, 50)]:
        for M in [1, 2, 3, 4, 7, 8, 9, 12, 16]:
            for dtype in [torch.float32, torch.bfloat16, torch.float16]:
                try:
                    err, sc = run(M, K, N, dtype, 7)
                    worst[str(dtype)] = max(worst.get(str(dtype), 0.0), err / sc)
                except RuntimeError as e:
                    print(json.dumps({"K": K, "N": N, "M": M, "dtype": str(dtype), "error": str(e)[:100]}), flush=True)
    print(json.dumps({"worst_rel_err_vs_f64": worst}), flush=True)
    import dec_tune
    for (K, N) in [(4096, 11008), (11008, 4096)]:
        for M in [int(v) for v in os.environ.get("HM_MS", "3,4,8,16").split(",")]:
            a = dec_tune.measure(M, K, N, {"force_path": 7})
            b = dec_tune.measure(M, K, N, {"hm_min_m": 99})
            nb = N * K // 2 + 8 * N + 4 * M * K + 4 * M * N
            print(json.dumps({"K": K, "N": N, "M": M, "hm_us": round(a, 3), "old_us": round(b, 3), "hm_frac": round(nb / a / 1e3 / 6532.5, 3)}), flush=True)
