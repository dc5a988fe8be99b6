"""Packed-INT4 quantisation format: drop-in for the reference's python/quantize.py.

Same names, arguments and results as the reference functions (python/quantize.py:38-124,
127-173, 176-202), computed by libb200q's sm_100a kernels.  Packing, scales and zero points are
bit-exact with the reference; there is no CPU implementation here -- tensors that live on the host
are moved to the current CUDA device, processed there, and the results moved back.
"""
from __future__ import annotations

import torch

from . import _lib


def _to_device(t: torch.Tensor):
    if t.is_cuda:
        return t, None
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: the b200 INT4 path has no CPU implementation")
    return t.cuda(), t.device


def quantize_weights(weight_fp32: torch.Tensor, num_bits: int = 4):
    """[N,K] float32 -> (packed uint8 [N,K/2], scales f32 [N], zero_points f32 [N]).

    Asymmetric per-row INT4, low nibble = even column (python/quantize.py:38-124).
    """
    assert weight_fp32.ndim == 2, "Weight must be 2D [output_dim, input_dim]"
    assert weight_fp32.shape[1] % 2 == 0, "input_dim must be even for packing"
    if num_bits != 4:
        raise ValueError("only num_bits=4 is implemented (the reference's packing is 4-bit only)")
    w, home = _to_device(weight_fp32)
    w = w.to(torch.float32).contiguous()
    packed, scales, zps = _lib.quantize_rows(w)
    if home is not None:
        return packed.to(home), scales.to(home), zps.to(home)
    return packed, scales, zps


def dequantize_weights(packed_uint8: torch.Tensor, scales: torch.Tensor, zero_points: torch.Tensor):
    """(packed, scales, zero_points) -> [N,K] float32, (q - zp) * scale (python/quantize.py:127-173)."""
    p, home = _to_device(packed_uint8)
    s = scales.to(p.device, torch.float32).contiguous()
    z = zero_points.to(p.device, torch.float32).contiguous()
    out = _lib.dequantize_rows(p.contiguous(), s, z)
    return out.to(home) if home is not None else out


def reference_quantized_linear(input: torch.Tensor, packed_weights: torch.Tensor,
                               scales: torch.Tensor, zero_points: torch.Tensor):
    """Unfused form: materialise the fp32 weights, then a plain matmul (python/quantize.py:176-202).

    Kept for API compatibility; the product path is ``QuantizedLinear`` / ``fused_quant_linear_cuda``.
    """
    x, home = _to_device(input)
    w = dequantize_weights(packed_weights.to(x.device), scales.to(x.device), zero_points.to(x.device))
    y = torch.nn.functional.linear(x, w)
    return y.to(home) if home is not None else y


def quantize_weights_grouped(weight_fp32: torch.Tensor, group_size: int = 128):
    """Group-wise variant: the reference's per-row formulas (python/quantize.py:73-109) applied to every
    `group_size` consecutive columns.  [N,K] float32 -> (packed uint8 [N,K/2], scales f32 [N,K/G], zero_points f32 [N,K/G]).
    The nibble order inside a row is unchanged, so `packed` differs from the per-row result only through the codes."""
    assert weight_fp32.ndim == 2 and group_size % 8 == 0 and weight_fp32.shape[1] % group_size == 0
    N, K = weight_fp32.shape
    packed, scales, zps = quantize_weights(weight_fp32.reshape(N * (K // group_size), group_size))
    return packed.reshape(N, K // 2), scales.reshape(N, K // group_size), zps.reshape(N, K // group_size)


def dequantize_weights_grouped(packed_uint8: torch.Tensor, scales: torch.Tensor, zero_points: torch.Tensor):
    """Inverse of quantize_weights_grouped: (q - zp[n, k // G]) * scale[n, k // G] -> [N,K] float32."""
    N, Kh = packed_uint8.shape
    ngrp = scales.shape[1]
    out = dequantize_weights(packed_uint8.reshape(N * ngrp, Kh // ngrp), scales.reshape(-1), zero_points.reshape(-1))
    return out.reshape(N, 2 * Kh)
