"""``moe_int4_cuda`` -- same module name and 7-argument ``forward`` as the reference extension
(csrc/moe_int4_kernel.cu:93-141), with the semantics that kernel intends:

    out[t, :] = inputs[t, :] @ dequant(W[e])^T   for input_offsets[e] <= t < input_offsets[e] + tokens_per_expert[e]

(the reference kernel itself only fills columns 0..255 from each expert's first token -- see
SURVEY.md section 2.2 -- so parity is defined against python/quantize.py's dequantize + matmul).
Rows covered by no expert stay zero (moe_int4_kernel.cu:109).  No ``.item()`` calls, no device
synchronisation: offsets are read on the device.
"""
import torch

from b200q_pkg import pkg as _pkg

_lib = _pkg._lib


def forward(packed_weights, scales, zero_points, inputs, expert_ids, tokens_per_expert, input_offsets):
    for name, t in (("packed_weights", packed_weights), ("scales", scales), ("zero_points", zero_points),
                    ("inputs", inputs), ("tokens_per_expert", tokens_per_expert), ("input_offsets", input_offsets)):
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")
    if packed_weights.dim() != 3 or inputs.dim() != 2:
        raise RuntimeError("packed_weights must be [E, F, K/2] and inputs [T, K]")
    E, F, Kh = packed_weights.shape
    if inputs.size(1) != 2 * Kh:
        raise RuntimeError("packed_weights dim 2 must be hidden_dim / 2")
    if tokens_per_expert.numel() != E or input_offsets.numel() != E:
        raise RuntimeError("tokens_per_expert and input_offsets must have one entry per expert")
    x = inputs.contiguous()
    starts = input_offsets.to(torch.int32).contiguous()
    ends = (starts + tokens_per_expert.to(torch.int32)).contiguous()
    out = torch.zeros((x.size(0), F), dtype=x.dtype, device=x.device)
    if x.size(0) == 0:
        return out
    # expert_ids is accepted and ignored, exactly as in the reference (it is never read there)
    return _lib.moe_grouped_fwd_ranges(x, packed_weights.contiguous(), scales.contiguous().float(),
                                       zero_points.contiguous().float(), starts, ends, out)
