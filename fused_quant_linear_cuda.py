"""``fused_quant_linear_cuda`` -- same module name and ``forward`` signature as the reference's
pybind11 extension (csrc/quantized_linear.cpp:22-28), backed by libb200q.so (sm_100a).

Contract mirrored from csrc/quantized_linear_kernel.cu:293-378: CUDA tensors, contiguous input and
packed weights, dtypes f32 / u8 / f32 / f32, packed.size(1) == K/2, 1-D input -> 1-D output, a fresh
output tensor, ``RuntimeError`` on violation.  Differences (all stricter or safer): scales and
zero points must be contiguous, the launch is on torch's current stream under a device guard, and
launch errors are reported.
"""
import torch

from b200q_pkg import pkg as _pkg

_lib = _pkg._lib


def _check(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def forward(input, packed_weights, scales, zero_points):
    """Fused INT4 dequantize + linear forward: input [K] or [M,K] f32 -> [N] or [M,N] f32."""
    ext = _lib.torch_ext()
    if ext is not None:          # compiled binding (csrc/torch_binding.cpp): the same checks and messages, in C++
        return ext.forward(input, packed_weights, scales, zero_points)
    squeeze = False
    if input.dim() == 1:
        input = input.unsqueeze(0)
        squeeze = True
    _check(input.is_cuda, "input must be a CUDA tensor")
    _check(packed_weights.is_cuda, "packed_weights must be a CUDA tensor")
    _check(scales.is_cuda, "scales must be a CUDA tensor")
    _check(zero_points.is_cuda, "zero_points must be a CUDA tensor")
    _check(input.is_contiguous(), "input must be contiguous")
    _check(packed_weights.is_contiguous(), "packed_weights must be contiguous")
    _check(scales.is_contiguous(), "scales must be contiguous")
    _check(zero_points.is_contiguous(), "zero_points must be contiguous")
    _check(input.dtype == torch.float32, "input must be float32")
    _check(packed_weights.dtype == torch.uint8, "packed_weights must be uint8")
    _check(scales.dtype == torch.float32, "scales must be float32")
    _check(zero_points.dtype == torch.float32, "zero_points must be float32")
    _check(input.dim() == 2 and packed_weights.dim() == 2, "input must be 1-D or 2-D, packed_weights 2-D")
    input_dim = input.size(1)
    _check(packed_weights.size(1) == input_dim // 2 and input_dim % 2 == 0,
           "packed_weights dim 1 must be input_dim / 2")
    _check(scales.numel() == packed_weights.size(0) and zero_points.numel() == packed_weights.size(0),
           "scales / zero_points must have one entry per output row")
    out = _lib.linear_fwd(input, packed_weights, scales, zero_points, out_dtype=torch.float32)
    return out.squeeze(0) if squeeze else out
