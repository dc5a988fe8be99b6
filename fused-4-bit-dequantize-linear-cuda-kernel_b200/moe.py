"""INT4 MoE modules: drop-ins for the reference's two MoE surfaces.

* ``QuantizedMoEExpert`` / ``QuantizedMoE`` -- benchmark/moe_grouped_gemm/moe_int4_module.py:21-130
  (per-row INT4 experts, ``forward(List[Tensor]) -> List[Tensor]``), plus the full routed layer the
  north star asks for: ``forward_routed(x, router_logits)`` = router top-k -> permutation ->
  grouped GEMM (w1||w3, SiLU-gate) -> grouped GEMM (w2) -> weighted combine, with no host sync.
* ``MoEINT4`` / ``quantize_weights_moe`` -- python/moe_int4_module.py:19-146 (per-expert scalar
  scale, 7-argument ``moe_int4_cuda.forward``).

Every forward runs libb200q's sm_100a kernels; there is no PyTorch or CPU fallback.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib
from .quantize import quantize_weights
from .routing import DeviceRouting, route


class QuantizedMoEExpert(nn.Module):
    """One expert projection with packed INT4 weights (moe_int4_module.py:21-81)."""

    _gen = 0    # bumped whenever a weight buffer is replaced or moved: a QuantizedMoE re-stacks lazily when it changes

    def __setattr__(self, name, value):
        if name in ("packed_weights", "scales", "zero_points"):
            object.__setattr__(self, "_gen", self._gen + 1)
        super().__setattr__(name, value)

    def _apply(self, fn, *args, **kwargs):
        object.__setattr__(self, "_gen", self._gen + 1)
        return super()._apply(fn, *args, **kwargs)

    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.register_buffer("packed_weights",
                             torch.zeros(out_features, in_features // 2, dtype=torch.uint8))
        self.register_buffer("scales", torch.zeros(out_features, dtype=torch.float32))
        self.register_buffer("zero_points", torch.zeros(out_features, dtype=torch.float32))

    @classmethod
    def from_fp16(cls, weight: torch.Tensor) -> "QuantizedMoEExpert":
        assert weight.shape[1] % 2 == 0, "in_features must be even for INT4 packing"
        expert = cls(weight.shape[1], weight.shape[0])
        packed, scales, zp = quantize_weights(weight.float())      # moe_int4_module.py:55
        expert.packed_weights = packed
        expert.scales = scales
        expert.zero_points = zp
        return expert

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.shape[0] == 0:                                        # moe_int4_module.py:65-68
            return torch.empty(0, self.out_features, device=x.device, dtype=torch.float16)
        _lib.require_cuda(x, "x")
        # reference: x @ dequant(W).T.to(x.dtype) -> result in x.dtype (moe_int4_module.py:71-72)
        # (buffers of a gated QuantizedMoE are strided views of the interleaved w1 / w3 stack: compact them for this call)
        return _lib.linear_fwd(x.contiguous(), self.packed_weights.contiguous(), self.scales.contiguous(),
                               self.zero_points.contiguous(), out_dtype=x.dtype)

    @property
    def weight_memory_bytes(self) -> int:
        return self.packed_weights.numel() + self.scales.numel() * 4 + self.zero_points.numel() * 4


_BUFS = ("packed_weights", "scales", "zero_points")


def _aliases(t: torch.Tensor, view: torch.Tensor) -> bool:
    return (t.device == view.device and t.data_ptr() == view.data_ptr() and t.shape == view.shape
            and t.stride() == view.stride() and t.dtype == view.dtype)


class QuantizedMoE(nn.Module):
    """MoE layer with INT4 experts (moe_int4_module.py:84-130).

    Reference mode: ``experts`` holds one ``hidden -> ffn`` projection per expert and
    ``forward(list) -> list`` applies them.  Gated mode (``from_gated_fp16_weights``): ``experts``
    holds w1 (gate), ``experts_up`` w3 and ``experts_down`` w2, and ``forward_routed`` runs the
    whole layer ``sum_k p_k * w2_e( silu(w1_e x) * (w3_e x) )``.

    Weights exist ONCE on the device: the grouped kernels want all experts in one ``[E, N, K/2]`` tensor (w1 and w3
    interleaved row by row for the fused SiLU-gate), so that tensor is the storage and the per-expert buffers
    (``experts.{i}.packed_weights`` ... -- the reference's state_dict keys, moe_int4_module.py:38-45) are views into
    it.  ``load_state_dict`` copies through the views; replacing an expert or moving the module re-stacks lazily.
    """

    def __init__(self, num_experts: int, hidden_dim: int, ffn_dim: int, gated: bool = False):
        super().__init__()
        self.num_experts = num_experts
        self.hidden_dim = hidden_dim
        self.ffn_dim = ffn_dim
        self.gated = gated
        self.experts = nn.ModuleList(
            [QuantizedMoEExpert(hidden_dim, ffn_dim) for _ in range(num_experts)])
        if gated:
            self.experts_up = nn.ModuleList(
                [QuantizedMoEExpert(hidden_dim, ffn_dim) for _ in range(num_experts)])
            self.experts_down = nn.ModuleList(
                [QuantizedMoEExpert(ffn_dim, hidden_dim) for _ in range(num_experts)])
        self._stacked = None   # (w13 triple, w2 triple or None): the storage the expert buffers are views of
        self._stamp = None

    @classmethod
    def from_fp16_weights(cls, weights: List[torch.Tensor]) -> "QuantizedMoE":
        assert len(weights) > 0
        moe = cls(len(weights), weights[0].shape[1], weights[0].shape[0])
        for i, w in enumerate(weights):
            moe.experts[i] = QuantizedMoEExpert.from_fp16(w)
        return moe

    @classmethod
    def from_gated_fp16_weights(cls, w1: List[torch.Tensor], w3: List[torch.Tensor],
                                w2: List[torch.Tensor]) -> "QuantizedMoE":
        """w1[e], w3[e]: [ffn, hidden]; w2[e]: [hidden, ffn] (Mixtral naming)."""
        assert len(w1) == len(w3) == len(w2) > 0
        moe = cls(len(w1), w1[0].shape[1], w1[0].shape[0], gated=True)
        for i in range(len(w1)):
            moe.experts[i] = QuantizedMoEExpert.from_fp16(w1[i])
            moe.experts_up[i] = QuantizedMoEExpert.from_fp16(w3[i])
            moe.experts_down[i] = QuantizedMoEExpert.from_fp16(w2[i])
        return moe

    def forward(self, expert_inputs: List[torch.Tensor]) -> List[torch.Tensor]:
        return [expert(x) for expert, x in zip(self.experts, expert_inputs)]   # moe_int4_module.py:123-125

    # ------------------------------------------------------------------ fused routed layer
    @property
    def fused_gate(self) -> bool:
        """The SiLU-gate runs in the epilogue of the first grouped GEMM (rows of w1 and w3 interleaved) whenever the
        tcgen05 path takes the shape; otherwise w1||w3 are concatenated and b200q_moe_silu_mul follows."""
        return self.gated and self.experts[0].in_features % 128 == 0

    def _views(self, w13, w2):
        """(module, buffer name, view into the stacked storage) for every per-expert buffer."""
        out = []
        F = self.ffn_dim
        for e in range(self.num_experts):
            for j, name in enumerate(_BUFS):
                if not self.gated:
                    out.append((self.experts[e], name, w13[j][e]))
                elif self.fused_gate:
                    out.append((self.experts[e], name, w13[j][e, 0::2]))
                    out.append((self.experts_up[e], name, w13[j][e, 1::2]))
                else:
                    out.append((self.experts[e], name, w13[j][e, :F]))
                    out.append((self.experts_up[e], name, w13[j][e, F:]))
                if self.gated:
                    out.append((self.experts_down[e], name, w2[j][e]))
        return out

    def _restack(self):
        """Move every expert's buffers into the stacked tensors and turn them into views of those."""
        first = self.experts[0]
        dev = first.packed_weights.device
        E, F, d = self.num_experts, self.ffn_dim, self.hidden_dim
        n13 = 2 * F if self.gated else F
        w13 = (torch.empty((E, n13, d // 2), dtype=torch.uint8, device=dev),
               torch.empty((E, n13), dtype=torch.float32, device=dev),
               torch.empty((E, n13), dtype=torch.float32, device=dev))
        w2 = None
        if self.gated:
            w2 = (torch.empty((E, d, F // 2), dtype=torch.uint8, device=dev),
                  torch.empty((E, d), dtype=torch.float32, device=dev),
                  torch.empty((E, d), dtype=torch.float32, device=dev))
        for mod, name, view in self._views(w13, w2):
            view.copy_(mod._buffers[name].to(dev))
            mod._buffers[name] = view                       # the old tensor is released: one copy on the device
        self._stacked = (w13, w2)
        self._stamp = self._expert_stamp()

    def _expert_stamp(self):
        mods = list(self.experts) + (list(self.experts_up) + list(self.experts_down) if self.gated else [])
        return tuple((id(m), m._gen) for m in mods)

    def stacked_weights(self):
        """[E,N,K/2] / [E,N] tensors for the grouped kernels: w1 and w3 stacked along N (interleaved row by row when
        `fused_gate`, else concatenated), and w2.  These ARE the weights; the per-expert buffers alias them."""
        # cheap check per call (identity and generation counter of every expert module); the full aliasing check only
        # when that changed
        if self._stacked is None or self._stamp != self._expert_stamp():
            st = self._stacked
            if st is None or not all(_aliases(mod._buffers[name], view) for mod, name, view in self._views(*st)):
                self._restack()
            else:
                self._stamp = self._expert_stamp()
        return self._stacked

    def forward_routed(self, x: torch.Tensor, router_logits: torch.Tensor, top_k: int = 2,
                       routing: Optional[DeviceRouting] = None) -> torch.Tensor:
        """x [T,hidden], router_logits [T,E] -> [T, hidden] (gated) or [T, ffn] (single projection).

        Everything stays on the current stream; no host synchronisation.  Output is fp32, like the
        reference's combine (routing.py:186-187 multiplies by fp32 routing weights).
        """
        _lib.require_cuda(x, "x")
        # (decode call vs grouped tcgen05 GEMMs, Mixtral layer: 0.069 / 0.22 ms at T = 1, 0.26 / 0.47 at T = 8, 0.30 / 0.49 at
        # T = 16 -- groups of three and more rows run on the mid-batch kernel: tools/moe_decode_crossover.py)
        if (routing is None and self.fused_gate and x.shape[0] <= 16
                and self.hidden_dim % 256 == 0
                and self.ffn_dim % 256 == 0 and max(self.hidden_dim, self.ffn_dim) <= 16384
                and self.num_experts <= 256 and top_k <= 8):
            # decode-sized call: the whole layer behind one C call, both expert GEMVs grouped on the resident decode kernel
            w13, w2 = self.stacked_weights()
            return _lib.moe_decode_fwd(x.contiguous(), router_logits.to(torch.float32).contiguous(), top_k, w13, w2)
        dr = routing if routing is not None else route(router_logits, top_k)
        return self.forward_dispatched(x, dr)

    def forward_dispatched(self, x: torch.Tensor, dr: DeviceRouting) -> torch.Tensor:
        xs = _lib.moe_gather_rows(x.contiguous(), dr.sorted_slot, dr.top_k)
        y = self.forward_grouped(xs, dr.offsets)
        return _lib.moe_combine(y, dr.inv_perm, dr.expert_weights, dr.top_k, out_dtype=torch.float32)

    def forward_grouped(self, xs: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
        """Rows of xs grouped by expert (offsets [E+1] int32 on the device) -> expert outputs."""
        w13, w2 = self.stacked_weights()
        if self.fused_gate:
            h = _lib.moe_grouped_gated_fwd(xs, w13[0], w13[1], w13[2], offsets)
        else:
            gu = _lib.moe_grouped_fwd(xs, w13[0], w13[1], w13[2], offsets)
            if not self.gated:
                return gu
            h = _lib.moe_silu_mul(gu)
        return _lib.moe_grouped_fwd(h, w2[0], w2[1], w2[2], offsets)

    def forward_ranges(self, rows: torch.Tensor, starts: torch.Tensor, ends: torch.Tensor,
                       range_expert: torch.Tensor, all_rows_covered: bool = False) -> torch.Tensor:
        """Rows [starts[v], ends[v]) of `rows` go through expert range_expert[v] (int32 device tensors; ranges may be
        empty and need not be sorted) -- what an expert-parallel rank runs on the rows it received, which arrive ordered
        (source rank, expert).  Rows covered by no range come back as zeros (unless the caller vouches that there are
        none: all_rows_covered skips the memset)."""
        w13, w2 = self.stacked_weights()
        z = not all_rows_covered
        if self.fused_gate:
            h = _lib.moe_grouped_fwd_mapped(rows, w13[0], w13[1], w13[2], starts, ends, range_expert, gated=True, zero_fill=False)
        else:
            gu = _lib.moe_grouped_fwd_mapped(rows, w13[0], w13[1], w13[2], starts, ends, range_expert, gated=False, zero_fill=z)
            if not self.gated:
                return gu
            h = _lib.moe_silu_mul(gu)
        return _lib.moe_grouped_fwd_mapped(h, w2[0], w2[1], w2[2], starts, ends, range_expert, gated=False, zero_fill=z)

    @property
    def total_memory_bytes(self) -> int:
        total = sum(e.weight_memory_bytes for e in self.experts)
        if self.gated:
            total += sum(e.weight_memory_bytes for e in self.experts_up)
            total += sum(e.weight_memory_bytes for e in self.experts_down)
        return total


# ---------------------------------------------------------------------------- MoEINT4 surface
def quantize_weights_moe(weights_list):
    """python/moe_int4_module.py:19-80: ONE scale / zero point per expert (from the expert's global
    min / max), stored broadcast as [E, F] so the per-row kernels apply unchanged.

    Arithmetic follows the reference on the CPU: scale = (max - min) / 15 in fp32, zp =
    clamp(round(-min / scale), 0, 15), q = clamp(round(w / scale + zp), 0, 15) with a true fp32
    division.
    """
    num_experts = len(weights_list)
    ffn_dim, hidden_dim = weights_list[0].shape
    device = weights_list[0].device
    if not weights_list[0].is_cuda:
        raise RuntimeError("quantize_weights_moe expects CUDA tensors (as the reference does)")
    packed = torch.empty(num_experts, ffn_dim, hidden_dim // 2, dtype=torch.uint8, device=device)
    scales = torch.empty(num_experts, ffn_dim, dtype=torch.float32, device=device)
    zps = torch.empty(num_experts, ffn_dim, dtype=torch.float32, device=device)
    # a 0-dim CUDA divisor keeps torch on the IEEE division kernel (a Python scalar divisor would be
    # turned into a multiplication by the reciprocal on CUDA, which is not what the CPU reference does)
    fifteen = torch.full((), 15.0, dtype=torch.float32, device=device)
    for e, w in enumerate(weights_list):
        w32 = w.float().contiguous()
        mm = _lib.minmax(w32)                                   # [min, max] on the device
        scale = (mm[1] - mm[0]) / fifteen                       # python/moe_int4_module.py:49
        zp = torch.clamp(torch.round(-mm[0] / scale), 0, 15)    # :50-51
        scales[e] = scale
        zps[e] = zp
        packed[e] = _lib.quantize_rows_given(w32, scales[e], zps[e])
    return packed, scales, zps


class MoEINT4(nn.Module):
    """python/moe_int4_module.py:83-146."""

    def __init__(self, num_experts, hidden_dim, ffn_dim):
        super().__init__()
        self.num_experts = num_experts
        self.hidden_dim = hidden_dim
        self.ffn_dim = ffn_dim
        self.packed_dim = hidden_dim // 2
        self.register_buffer("packed_weights",
                             torch.zeros(num_experts, ffn_dim, self.packed_dim, dtype=torch.uint8))
        self.register_buffer("scales", torch.zeros(num_experts, ffn_dim, dtype=torch.float32))
        self.register_buffer("zero_points", torch.zeros(num_experts, ffn_dim, dtype=torch.float32))

    @classmethod
    def from_weights(cls, weights_list):
        module = cls(len(weights_list), weights_list[0].shape[1], weights_list[0].shape[0])
        packed, scales, zp = quantize_weights_moe(weights_list)
        module.packed_weights = packed
        module.scales = scales
        module.zero_points = zp
        return module

    def forward(self, inputs, expert_ids, tokens_per_expert, input_offsets):
        import moe_int4_cuda   # repo-root shim with the reference's extension name
        return moe_int4_cuda.forward(self.packed_weights, self.scales, self.zero_points, inputs,
                                     expert_ids, tokens_per_expert, input_offsets)
