"""MoE routing / dispatch / combine: drop-in for benchmark/moe_grouped_gemm/routing.py.

``simulate_routing``, ``create_expert_inputs`` and ``combine_expert_outputs`` keep the reference's
signatures and return types (routing.py:26-93, 96-149, 152-189); the work is done by libb200q's
router / permutation / gather / combine kernels.  ``route`` is the sync-free device-side form the
fused layer uses (no Python loop over T*k items, no ``.cpu()``).

Within-expert row order: the reference sorts with ``torch.argsort`` (unstable, routing.py:128), so
only the per-expert *set* of rows is defined there; here the order is the stable one (ascending
flat assignment index t*k+s).  The combined output does not depend on it.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch

from . import _lib


@dataclass
class DeviceRouting:
    """Everything the expert GEMMs need, resident on the GPU (no host copies)."""
    expert_indices: torch.Tensor   # [T,k] int32
    expert_weights: torch.Tensor   # [T,k] f32, renormalised over the k winners
    counts: torch.Tensor           # [E] int32
    offsets: torch.Tensor          # [E+1] int32, offsets[E] = T*k
    sorted_slot: torch.Tensor      # [T*k] int32: flat assignment (t*k+s) at each sorted position
    inv_perm: torch.Tensor         # [T*k] int32: sorted position of flat assignment t*k+s
    num_experts: int
    top_k: int


@dataclass
class RoutingResult:
    """Same four fields as the reference (routing.py:17-23) plus the device-side routing."""
    expert_indices: torch.Tensor       # [T,k] int64
    expert_weights: torch.Tensor       # [T,k] f32
    tokens_per_expert: List[int]
    expert_token_offsets: List[int]
    device_routing: Optional[DeviceRouting] = field(default=None, repr=False)


def make_logits(num_tokens: int, num_experts: int, distribution: str = "skewed",
                device: str = "cuda", seed: int = 42) -> torch.Tensor:
    """The reference's synthetic router logits (routing.py:51-69), same RNG calls on `device`."""
    torch.manual_seed(seed)
    if distribution == "uniform":
        logits = torch.zeros(num_tokens, num_experts, device=device)
    elif distribution == "skewed":
        expert_probs = 1.0 / (torch.arange(num_experts, device=device, dtype=torch.float32) + 1)
        expert_probs = expert_probs / expert_probs.sum()
        logits = torch.log(expert_probs + 1e-10).unsqueeze(0).expand(num_tokens, -1)
        logits = logits + torch.randn_like(logits) * 0.5
    elif distribution == "random":
        logits = torch.randn(num_tokens, num_experts, device=device)
    else:
        raise ValueError(f"Unknown distribution: {distribution}")
    return logits


def route(logits: torch.Tensor, top_k: int) -> DeviceRouting:
    """softmax -> top-k -> renormalise -> histogram / offsets / stable permutation, all on the GPU."""
    _lib.require_cuda(logits, "logits")
    logits = logits.to(torch.float32).contiguous()
    E = logits.shape[1]
    idx, w, counts, offsets, sorted_slot, inv_perm = _lib.moe_route(logits, top_k)      # one C call: top-k + permutation
    return DeviceRouting(idx, w, counts, offsets, sorted_slot, inv_perm, E, top_k)


def routing_result(dr: DeviceRouting) -> RoutingResult:
    counts = dr.counts.tolist()                     # one D2H copy (the reference loops over T*k items)
    offs = [0]
    for c in counts[:-1]:
        offs.append(offs[-1] + c)
    return RoutingResult(dr.expert_indices.to(torch.int64), dr.expert_weights, counts, offs, dr)


def simulate_routing(num_tokens: int, num_experts: int, top_k: int, distribution: str = "skewed",
                     device: str = "cuda", seed: int = 42) -> RoutingResult:
    """routing.py:26-93.  Logits may be generated on the CPU (device="cpu") for RNG parity with a
    CPU run of the reference; routing itself always runs on the current CUDA device."""
    logits = make_logits(num_tokens, num_experts, distribution, device, seed)
    return routing_result(route(logits.cuda() if not logits.is_cuda else logits, top_k))


def _device_routing(routing: RoutingResult, num_experts: int, device) -> DeviceRouting:
    if routing.device_routing is not None:
        return routing.device_routing
    # a RoutingResult built elsewhere (e.g. by the reference): rebuild the permutation on the GPU
    idx = routing.expert_indices.to(device=device, dtype=torch.int32).contiguous()
    counts, offsets, sorted_slot, inv_perm = _lib.moe_permute(idx, num_experts)
    dr = DeviceRouting(idx, routing.expert_weights.to(device=device, dtype=torch.float32).contiguous(),
                       counts, offsets, sorted_slot, inv_perm, num_experts, idx.shape[1])
    routing.device_routing = dr
    return dr


def create_expert_inputs(x: torch.Tensor, routing: RoutingResult, num_experts: int,
                         top_k: int) -> Tuple[List[torch.Tensor], torch.Tensor]:
    """routing.py:96-149: per-expert row tensors (views of one gathered buffer) + inverse permutation."""
    _lib.require_cuda(x, "x")
    dr = _device_routing(routing, num_experts, x.device)
    xs = _lib.moe_gather_rows(x.contiguous(), dr.sorted_slot, top_k)
    expert_inputs, off = [], 0
    for e in range(num_experts):
        c = routing.tokens_per_expert[e]
        expert_inputs.append(xs[off:off + c])
        off += c
    return expert_inputs, dr.inv_perm.to(torch.int64)


def combine_expert_outputs(expert_outputs: List[torch.Tensor], routing: RoutingResult,
                           permutation: torch.Tensor, top_k: int) -> torch.Tensor:
    """routing.py:152-189: un-permute and sum the k expert outputs of every token, weighted."""
    concat = torch.cat(expert_outputs, dim=0).contiguous()
    _lib.require_cuda(concat, "expert_outputs")
    w = routing.expert_weights.to(device=concat.device, dtype=torch.float32).contiguous()
    inv = permutation.to(device=concat.device, dtype=torch.int32).contiguous()
    # the reference multiplies by fp32 weights, so the result is fp32 whatever the expert dtype is
    return _lib.moe_combine(concat, inv, w, top_k, out_dtype=torch.float32)


def get_expert_sizes_for_benchmark(num_tokens: int, num_experts: int, hidden_dim: int, ffn_dim: int,
                                   distribution: str = "skewed", device: str = "cuda"):
    """routing.py:192-222."""
    routing = simulate_routing(num_tokens, num_experts, 2, distribution, device)
    return routing.tokens_per_expert, hidden_dim, ffn_dim
