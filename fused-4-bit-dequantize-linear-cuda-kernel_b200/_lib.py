"""ctypes binding of libb200q.so (C ABI in include/b200q.h).

This is the whole "torch extension": torch supplies device memory (``data_ptr``) and the current
CUDA stream; every computation happens inside the shared library.  There is no CPU path -- if the
library is missing, or a tensor is not on an sm_100 GPU, the call raises.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

F32, F16, BF16 = 0, 1, 2
FLAG_NONE, FLAG_STATIC_WEIGHTS = 0, 1

_DTYPE_CODE = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}

_here = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200Q_LIB") or os.path.join(_here, "libb200q.so")   # B200Q_LIB: tools/ pick the -DB200Q_PROF build

_c = ctypes
_vp, _i64, _i32, _sz, _u32 = _c.c_void_p, _c.c_int64, _c.c_int, _c.c_size_t, _c.c_uint

# name -> (restype, argtypes); must list every symbol include/b200q.h declares
SIGNATURES = {
    "b200q_version": (_i32, []),
    "b200q_last_error_string": (_c.c_char_p, []),
    "b200q_device_check": (_i32, [_i32]),
    "b200q_sm_count": (_i32, []),
    "b200q_quantize_rows": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    "b200q_quantize_rows_given": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    "b200q_minmax_ws_bytes": (_sz, []),
    "b200q_minmax": (_i32, [_vp, _i64, _vp, _vp, _sz, _vp]),
    "b200q_dequantize_rows": (_i32, [_vp, _vp, _vp, _i64, _i64, _vp, _vp]),
    "b200q_linear_ws_bytes": (_sz, [_i64, _i64, _i64]),
    "b200q_linear_fwd": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _i32, _i64, _i64, _i64, _vp, _sz, _u32, _vp]),
    "b200q_linear_fwd_next": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _i32, _i64, _i64, _i64, _vp, _sz, _u32, _vp, _vp, _sz]),
    "b200q_linear_bias_fwd": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i64, _i64, _vp, _sz, _u32, _vp, _vp, _sz]),
    "b200q_linear_groupwise_fwd": (_i32, [_vp, _i32, _vp, _vp, _vp, _i64, _vp, _i32, _i64, _i64, _i64, _vp]),
    "b200q_linear_groupwise_bias_fwd": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _i32, _i64, _i64, _i64, _u32, _vp, _vp, _sz]),
    "b200q_linear_gated_fwd": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _i32, _i64, _i64, _i64, _vp, _sz, _u32, _vp, _vp, _sz]),
    "b200q_linear_fwd_host": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i64, _i64, _vp, _sz, _u32, _vp]),
    "b200q_tune_set": (_i32, [_c.c_char_p, _i32]),
    "b200q_moe_topk": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    "b200q_moe_permute_ws_bytes": (_sz, [_i64, _i32, _i32]),
    "b200q_moe_permute": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200q_moe_permute_mapped": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200q_moe_route": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200q_moe_gather_rows": (_i32, [_vp, _i32, _vp, _i64, _i32, _i64, _vp, _vp]),
    "b200q_moe_grouped_ws_bytes": (_sz, [_i64, _i32, _i64, _i64]),
    "b200q_moe_grouped_fwd": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _i64, _i64, _i64, _vp, _sz, _vp]),
    "b200q_moe_grouped_gated_fwd": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _i64, _i64, _i64, _vp, _sz, _vp]),
    "b200q_moe_grouped_fwd_ranges": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _i64, _i64, _i64, _vp, _sz, _vp]),
    "b200q_moe_grouped_fwd_mapped": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _i32, _i64, _i64, _i64, _vp, _sz, _vp]),
    "b200q_ep_plan": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200q_ep_unique_id": (_i32, [_vp]),
    "b200q_ep_comm_create": (_i32, [_vp, _i32, _i32, _c.POINTER(_vp)]),
    "b200q_ep_comm_destroy": (_i32, [_vp]),
    "b200q_ep_allgather_i32": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "b200q_ep_exchange": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _i64, _vp]),
    "b200q_moe_decode_ws_bytes": (_sz, [_i64, _i32, _i32, _i64, _i64]),
    "b200q_moe_decode_fwd": (_i32, [_vp, _i32, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _sz, _vp]),
    "b200q_moe_silu_mul": (_i32, [_vp, _i32, _i64, _i64, _vp, _vp]),
    "b200q_moe_combine": (_i32, [_vp, _i32, _vp, _vp, _i64, _i32, _i64, _vp, _i32, _vp]),
}

_lib = None
_lock = threading.Lock()


def load() -> ctypes.CDLL:
    """Load libb200q.so (built by ``__graft_entry__.build()`` / ``make -C csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(there is no CPU or PyTorch fallback for the INT4 kernels)")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                try:
                    fn = getattr(lib, name)
                except AttributeError:
                    if "B200Q_LIB" in os.environ:       # A/B runs against an older build (tools/): bind what it has
                        continue
                    raise
                fn.restype = res
                fn.argtypes = args
            _lib = lib
            # bench hook: B200Q_TUNE="gemm_bn=256,gemv_early=91" applies b200q_tune_set keys at load time
            for item in filter(None, os.environ.get("B200Q_TUNE", "").split(",")):
                key, _, val = item.partition("=")
                lib.b200q_tune_set(key.strip().encode(), int(val))
    return _lib


_torch_ext = None


def torch_ext():
    """The compiled torch binding (csrc/torch_binding.cpp -> b200q_torch.so: checks, allocation, stream lookup and the C-ABI
    call in one C++ function), or None when it has not been built -- the ctypes wrappers below then do the same work."""
    global _torch_ext
    if _torch_ext is None:
        path = os.path.join(_here, "b200q_torch.so")
        if os.environ.get("B200Q_NO_TORCH_EXT") or os.environ.get("B200Q_LIB") or not os.path.exists(path):
            _torch_ext = False
        else:
            try:
                load()                                   # libb200q.so first (B200Q_TUNE is applied there)
                import importlib.util
                spec = importlib.util.spec_from_file_location("b200q_torch", path)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                _torch_ext = mod
            except Exception:
                _torch_ext = False
    return _torch_ext or None


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().b200q_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPE_CODE[t.dtype]
    except KeyError:
        raise RuntimeError(f"unsupported dtype {t.dtype}; expected float32, float16 or bfloat16") from None


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (libb200q has no CPU path)")


_ws_cache = {}


def workspace(device: torch.device, nbytes: int, tag: str) -> torch.Tensor:
    """Zero-initialised scratch, private to (device, stream, entry point): every entry point of
    b200q.h that takes a workspace expects to find it as it left it (zeroed tickets / counters)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream, tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def tune(key: str, value: int) -> None:
    check(load().b200q_tune_set(key.encode(), int(value)), "b200q_tune_set")


# ------------------------------------------------------------------------------------ wrappers
def linear_fwd(x: torch.Tensor, packed: torch.Tensor, scales: torch.Tensor, zps: torch.Tensor,
               out_dtype=None, flags: int = FLAG_NONE, out: torch.Tensor | None = None,
               next_packed: torch.Tensor | None = None, bias: torch.Tensor | None = None) -> torch.Tensor:
    """y[M,N] = x[M,K] @ dequant(packed, scales, zps)^T (+ bias [N] f32) on the current stream of x's device.
    next_packed: packed weights of the fused linear that follows on this stream (L2 prefetch hint)."""
    ext = torch_ext()
    if ext is not None and out is None:
        return ext.linear_forward(x, packed, scales, zps, bias, out_dtype, flags, next_packed)
    lib = load()
    M, K = x.shape
    N = packed.shape[0]
    out_dtype = out_dtype or x.dtype
    with torch.cuda.device(x.device):
        y = out if out is not None else torch.empty((M, N), dtype=out_dtype, device=x.device)
        ws_bytes = lib.b200q_linear_ws_bytes(M, N, K)
        ws = workspace(x.device, ws_bytes, "linear") if ws_bytes else None
        check(lib.b200q_linear_bias_fwd(x.data_ptr(), dtype_code(x), packed.data_ptr(), scales.data_ptr(),
                                        zps.data_ptr(), bias.data_ptr() if bias is not None else None,
                                        y.data_ptr(), dtype_code(y), M, N, K,
                                        ws.data_ptr() if ws is not None else None, ws.numel() if ws is not None else 0,
                                        flags, stream_ptr(x.device),
                                        next_packed.data_ptr() if next_packed is not None else None,
                                        next_packed.numel() if next_packed is not None else 0), "b200q_linear_fwd")
    return y


def linear_groupwise_fwd(x: torch.Tensor, packed: torch.Tensor, scales: torch.Tensor, zps: torch.Tensor, group_size: int,
                         out_dtype=None, flags: int = FLAG_NONE, bias: torch.Tensor | None = None,
                         next_packed: torch.Tensor | None = None) -> torch.Tensor:
    """y = x @ dequant(packed, scales [N, K/G], zps [N, K/G])^T (+ bias [N] f32) with one scale / zero point per G columns.
    next_packed: packed weights of the fused linear that follows on this stream (L2 prefetch hint)."""
    lib = load()
    M, K = x.shape
    N = packed.shape[0]
    out_dtype = out_dtype or x.dtype
    with torch.cuda.device(x.device):
        y = torch.empty((M, N), dtype=out_dtype, device=x.device)
        check(lib.b200q_linear_groupwise_bias_fwd(x.data_ptr(), dtype_code(x), packed.data_ptr(), scales.data_ptr(), zps.data_ptr(),
                                                  bias.data_ptr() if bias is not None else None, group_size, y.data_ptr(), dtype_code(y),
                                                  M, N, K, flags, stream_ptr(x.device),
                                                  next_packed.data_ptr() if next_packed is not None else None,
                                                  next_packed.numel() if next_packed is not None else 0),
              "b200q_linear_groupwise_bias_fwd")
    return y


def linear_gated_fwd(x: torch.Tensor, packed13: torch.Tensor, scales13: torch.Tensor, zps13: torch.Tensor,
                     out_dtype=None, flags: int = FLAG_NONE, next_packed: torch.Tensor | None = None) -> torch.Tensor:
    """h[M,F] = silu(x w_gate^T) * (x w_up^T); packed13 [2F,K/2] with gate / up rows interleaved (2f: gate, 2f+1: up)."""
    lib = load()
    M, K = x.shape
    F = packed13.shape[0] // 2
    out_dtype = out_dtype or x.dtype
    with torch.cuda.device(x.device):
        h = torch.empty((M, F), dtype=out_dtype, device=x.device)
        ws_bytes = lib.b200q_linear_ws_bytes(M, 2 * F, K)
        ws = workspace(x.device, ws_bytes, "linear") if ws_bytes else None
        check(lib.b200q_linear_gated_fwd(x.data_ptr(), dtype_code(x), packed13.data_ptr(), scales13.data_ptr(), zps13.data_ptr(),
                                         h.data_ptr(), dtype_code(h), M, F, K, ws.data_ptr() if ws is not None else None,
                                         ws.numel() if ws is not None else 0, flags, stream_ptr(x.device),
                                         next_packed.data_ptr() if next_packed is not None else None,
                                         next_packed.numel() if next_packed is not None else 0), "b200q_linear_gated_fwd")
    return h


def linear_fwd_host(x_host: torch.Tensor, x_dev: torch.Tensor, packed, scales, zps, y_dev: torch.Tensor,
                    y_host: torch.Tensor, flags: int = FLAG_NONE) -> torch.Tensor:
    """Host (pinned) activations in, host result out: H2D copy, fused kernel, D2H copy, all enqueued by one C
    call on the current stream of the weights' device (x_dev / y_dev are the caller's staging buffers)."""
    lib = load()
    M, K = x_host.shape
    N = packed.shape[0]
    dev = packed.device
    with torch.cuda.device(dev):
        ws_bytes = lib.b200q_linear_ws_bytes(M, N, K)
        ws = workspace(dev, ws_bytes, "linear") if ws_bytes else None
        check(lib.b200q_linear_fwd_host(x_host.data_ptr(), dtype_code(x_host), x_dev.data_ptr(), packed.data_ptr(),
                                        scales.data_ptr(), zps.data_ptr(), y_dev.data_ptr(), y_host.data_ptr(),
                                        dtype_code(y_host), M, N, K, ws.data_ptr() if ws is not None else None,
                                        ws.numel() if ws is not None else 0, flags, stream_ptr(dev)),
              "b200q_linear_fwd_host")
    return y_host


def quantize_rows(w: torch.Tensor):
    lib = load()
    N, K = w.shape
    with torch.cuda.device(w.device):
        packed = torch.empty((N, K // 2), dtype=torch.uint8, device=w.device)
        scales = torch.empty((N,), dtype=torch.float32, device=w.device)
        zps = torch.empty((N,), dtype=torch.float32, device=w.device)
        check(lib.b200q_quantize_rows(w.data_ptr(), N, K, packed.data_ptr(), scales.data_ptr(),
                                      zps.data_ptr(), stream_ptr(w.device)), "b200q_quantize_rows")
    return packed, scales, zps


def quantize_rows_given(w: torch.Tensor, scales: torch.Tensor, zps: torch.Tensor) -> torch.Tensor:
    lib = load()
    N, K = w.shape
    with torch.cuda.device(w.device):
        packed = torch.empty((N, K // 2), dtype=torch.uint8, device=w.device)
        check(lib.b200q_quantize_rows_given(w.data_ptr(), N, K, scales.data_ptr(), zps.data_ptr(),
                                            packed.data_ptr(), stream_ptr(w.device)), "b200q_quantize_rows_given")
    return packed


def minmax(v: torch.Tensor) -> torch.Tensor:
    lib = load()
    with torch.cuda.device(v.device):
        out = torch.empty((2,), dtype=torch.float32, device=v.device)
        nb = lib.b200q_minmax_ws_bytes()
        ws = workspace(v.device, nb, "minmax")
        check(lib.b200q_minmax(v.data_ptr(), v.numel(), out.data_ptr(), ws.data_ptr(), ws.numel(),
                               stream_ptr(v.device)), "b200q_minmax")
    return out


def dequantize_rows(packed: torch.Tensor, scales: torch.Tensor, zps: torch.Tensor) -> torch.Tensor:
    lib = load()
    N, Kh = packed.shape
    with torch.cuda.device(packed.device):
        out = torch.empty((N, Kh * 2), dtype=torch.float32, device=packed.device)
        check(lib.b200q_dequantize_rows(packed.data_ptr(), scales.data_ptr(), zps.data_ptr(), N, Kh * 2,
                                        out.data_ptr(), stream_ptr(packed.device)), "b200q_dequantize_rows")
    return out


def moe_topk(logits: torch.Tensor, k: int):
    lib = load()
    T, E = logits.shape
    with torch.cuda.device(logits.device):
        idx = torch.empty((T, k), dtype=torch.int32, device=logits.device)
        w = torch.empty((T, k), dtype=torch.float32, device=logits.device)
        check(lib.b200q_moe_topk(logits.data_ptr(), T, E, k, idx.data_ptr(), w.data_ptr(),
                                 stream_ptr(logits.device)), "b200q_moe_topk")
    return idx, w


def moe_permute(idx: torch.Tensor, E: int):
    """idx [T,k] int32 -> counts [E], offsets [E+1], sorted_slot [T*k], inv_perm [T*k] (all int32)."""
    lib = load()
    T, k = idx.shape
    dev = idx.device
    with torch.cuda.device(dev):
        counts = torch.empty((E,), dtype=torch.int32, device=dev)
        offsets = torch.empty((E + 1,), dtype=torch.int32, device=dev)
        sorted_slot = torch.empty((T * k,), dtype=torch.int32, device=dev)
        inv_perm = torch.empty((T * k,), dtype=torch.int32, device=dev)
        nb = lib.b200q_moe_permute_ws_bytes(T, E, k)
        ws = workspace(dev, nb, "permute")
        check(lib.b200q_moe_permute(idx.data_ptr(), T, E, k, counts.data_ptr(), offsets.data_ptr(),
                                    sorted_slot.data_ptr(), inv_perm.data_ptr(), ws.data_ptr(), ws.numel(),
                                    stream_ptr(dev)), "b200q_moe_permute")
    return counts, offsets, sorted_slot, inv_perm


def moe_route(logits: torch.Tensor, k: int, remap: torch.Tensor | None = None):
    """top-k + stable permutation in one C call: logits [T,E] f32 -> (idx [T,k] i32, weights [T,k] f32, counts [E],
    offsets [E+1], sorted_slot [T*k], inv_perm [T*k]); remap [E] i32: sort by remap[expert] (counts stay per expert)."""
    lib = load()
    T, E = logits.shape
    dev = logits.device
    with torch.cuda.device(dev):
        idx = torch.empty((T, k), dtype=torch.int32, device=dev)
        w = torch.empty((T, k), dtype=torch.float32, device=dev)
        counts = torch.empty((E,), dtype=torch.int32, device=dev)
        offsets = torch.empty((E + 1,), dtype=torch.int32, device=dev)
        sorted_slot = torch.empty((T * k,), dtype=torch.int32, device=dev)
        inv_perm = torch.empty((T * k,), dtype=torch.int32, device=dev)
        nb = lib.b200q_moe_permute_ws_bytes(T, E, k)
        ws = workspace(dev, nb, "permute")
        check(lib.b200q_moe_route(logits.data_ptr(), remap.data_ptr() if remap is not None else None, T, E, k, idx.data_ptr(),
                                  w.data_ptr(), counts.data_ptr(), offsets.data_ptr(), sorted_slot.data_ptr(), inv_perm.data_ptr(),
                                  ws.data_ptr(), ws.numel(), stream_ptr(dev)), "b200q_moe_route")
    return idx, w, counts, offsets, sorted_slot, inv_perm


def moe_gather_rows(x: torch.Tensor, sorted_slot: torch.Tensor, k: int) -> torch.Tensor:
    lib = load()
    rows = sorted_slot.numel()
    d = x.shape[1]
    with torch.cuda.device(x.device):
        xs = torch.empty((rows, d), dtype=x.dtype, device=x.device)
        check(lib.b200q_moe_gather_rows(x.data_ptr(), dtype_code(x), sorted_slot.data_ptr(), rows, k, d,
                                        xs.data_ptr(), stream_ptr(x.device)), "b200q_moe_gather_rows")
    return xs


def moe_grouped_fwd(xs: torch.Tensor, packed: torch.Tensor, scales: torch.Tensor, zps: torch.Tensor,
                    offsets: torch.Tensor, out_dtype=None) -> torch.Tensor:
    """xs [R,K] grouped by expert, packed [E,N,K/2], scales/zps [E,N], offsets [E+1] int32 (device)."""
    lib = load()
    R, K = xs.shape
    E, N = packed.shape[0], packed.shape[1]
    out_dtype = out_dtype or xs.dtype
    with torch.cuda.device(xs.device):
        y = torch.empty((R, N), dtype=out_dtype, device=xs.device)
        nb = lib.b200q_moe_grouped_ws_bytes(R, E, N, K)
        ws = workspace(xs.device, nb, "grouped") if nb else None
        check(lib.b200q_moe_grouped_fwd(xs.data_ptr(), dtype_code(xs), packed.data_ptr(), scales.data_ptr(),
                                        zps.data_ptr(), offsets.data_ptr(), E, y.data_ptr(), dtype_code(y),
                                        R, N, K, ws.data_ptr() if ws is not None else None,
                                        ws.numel() if ws is not None else 0, stream_ptr(xs.device)),
              "b200q_moe_grouped_fwd")
    return y


def moe_grouped_gated_fwd(xs: torch.Tensor, packed13: torch.Tensor, scales13: torch.Tensor, zps13: torch.Tensor,
                          offsets: torch.Tensor, out_dtype=None) -> torch.Tensor:
    """h [R,F] = silu(xs w1^T) * (xs w3^T) per expert group in one grouped GEMM; packed13 [E,2F,K/2] with the rows
    of w1 and w3 interleaved (2f: w1[f], 2f+1: w3[f]).  Needs K % 128 == 0."""
    lib = load()
    R, K = xs.shape
    E, N2 = packed13.shape[0], packed13.shape[1]
    out_dtype = out_dtype or xs.dtype
    with torch.cuda.device(xs.device):
        h = torch.empty((R, N2 // 2), dtype=out_dtype, device=xs.device)
        nb = lib.b200q_moe_grouped_ws_bytes(R, E, N2, K)
        ws = workspace(xs.device, nb, "grouped") if nb else None
        check(lib.b200q_moe_grouped_gated_fwd(xs.data_ptr(), dtype_code(xs), packed13.data_ptr(), scales13.data_ptr(),
                                              zps13.data_ptr(), offsets.data_ptr(), E, h.data_ptr(), dtype_code(h),
                                              R, N2 // 2, K, ws.data_ptr() if ws is not None else None,
                                              ws.numel() if ws is not None else 0, stream_ptr(xs.device)),
              "b200q_moe_grouped_gated_fwd")
    return h


def moe_grouped_fwd_ranges(xs: torch.Tensor, packed: torch.Tensor, scales: torch.Tensor, zps: torch.Tensor,
                           starts: torch.Tensor, ends: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """Like moe_grouped_fwd for explicit [starts[e], ends[e]) row ranges; writes only covered rows of `out`."""
    lib = load()
    R, K = xs.shape
    E, N = packed.shape[0], packed.shape[1]
    with torch.cuda.device(xs.device):
        nb = lib.b200q_moe_grouped_ws_bytes(R, E, N, K)
        ws = workspace(xs.device, nb, "grouped") if nb else None
        check(lib.b200q_moe_grouped_fwd_ranges(xs.data_ptr(), dtype_code(xs), packed.data_ptr(), scales.data_ptr(),
                                               zps.data_ptr(), starts.data_ptr(), ends.data_ptr(), E, out.data_ptr(),
                                               dtype_code(out), R, N, K, ws.data_ptr() if ws is not None else None,
                                               ws.numel() if ws is not None else 0, stream_ptr(xs.device)),
              "b200q_moe_grouped_fwd_ranges")
    return out


def moe_grouped_fwd_mapped(xs: torch.Tensor, packed: torch.Tensor, scales: torch.Tensor, zps: torch.Tensor,
                           starts: torch.Tensor, ends: torch.Tensor, range_expert: torch.Tensor, gated: bool,
                           out_dtype=None, zero_fill: bool = True) -> torch.Tensor:
    """Rows [starts[v], ends[v]) of xs through expert range_expert[v] of packed [E,N,K/2]; gated: h [R, N/2].
    zero_fill=False: rows covered by no range are left uninitialised (callers whose ranges cover every row)."""
    lib = load()
    R, K = xs.shape
    E, N = packed.shape[0], packed.shape[1]
    out_dtype = out_dtype or xs.dtype
    with torch.cuda.device(xs.device):
        y = (torch.zeros if zero_fill else torch.empty)((R, N // 2 if gated else N), dtype=out_dtype, device=xs.device)
        nb = lib.b200q_moe_grouped_ws_bytes(R, E, N, K)
        ws = workspace(xs.device, nb, "grouped") if nb else None
        check(lib.b200q_moe_grouped_fwd_mapped(xs.data_ptr(), dtype_code(xs), packed.data_ptr(), scales.data_ptr(),
                                               zps.data_ptr(), starts.data_ptr(), ends.data_ptr(), range_expert.data_ptr(),
                                               starts.numel(), E, 1 if gated else 0, y.data_ptr(), dtype_code(y), R, N, K,
                                               ws.data_ptr() if ws is not None else None,
                                               ws.numel() if ws is not None else 0, stream_ptr(xs.device)),
              "b200q_moe_grouped_fwd_mapped")
    return y


def moe_decode_fwd(x: torch.Tensor, logits: torch.Tensor, k: int, w13, w2) -> torch.Tensor:
    """Whole routed gated layer for T <= 16 tokens in one C call: w13 = (packed [E,2F,d/2] interleaved, scales, zps),
    w2 = (packed [E,d,F/2], scales, zps) -> out [T,d] f32."""
    lib = load()
    T, d = x.shape
    E, F = w13[0].shape[0], w13[0].shape[1] // 2
    dev = x.device
    with torch.cuda.device(dev):
        out = torch.empty((T, d), dtype=torch.float32, device=dev)
        nb = lib.b200q_moe_decode_ws_bytes(T, E, k, d, F)
        ws = workspace(dev, nb, "moe_decode")
        check(lib.b200q_moe_decode_fwd(x.data_ptr(), dtype_code(x), logits.data_ptr(), T, E, k, w13[0].data_ptr(), w13[1].data_ptr(),
                                       w13[2].data_ptr(), w2[0].data_ptr(), w2[1].data_ptr(), w2[2].data_ptr(), d, F,
                                       out.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr(dev)), "b200q_moe_decode_fwd")
    return out


def moe_silu_mul(gu: torch.Tensor) -> torch.Tensor:
    lib = load()
    R, F2 = gu.shape
    F = F2 // 2
    with torch.cuda.device(gu.device):
        h = torch.empty((R, F), dtype=gu.dtype, device=gu.device)
        check(lib.b200q_moe_silu_mul(gu.data_ptr(), dtype_code(gu), R, F, h.data_ptr(), stream_ptr(gu.device)),
              "b200q_moe_silu_mul")
    return h


def moe_combine(y: torch.Tensor, inv_perm: torch.Tensor, weights: torch.Tensor, k: int,
                out_dtype=torch.float32) -> torch.Tensor:
    lib = load()
    F = y.shape[1]
    T = inv_perm.numel() // k
    with torch.cuda.device(y.device):
        out = torch.empty((T, F), dtype=out_dtype, device=y.device)
        check(lib.b200q_moe_combine(y.data_ptr(), dtype_code(y), inv_perm.data_ptr(), weights.data_ptr(), T, k, F,
                                    out.data_ptr(), dtype_code(out), stream_ptr(y.device)), "b200q_moe_combine")
    return out
