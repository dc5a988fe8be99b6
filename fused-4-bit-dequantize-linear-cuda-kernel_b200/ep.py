"""Expert-parallel INT4 MoE layer over the GPUs of one NVSwitch box (one process per GPU).

Not in the reference (it is single-GPU); required by the north star.  Rank r owns
experts [r*E/N, (r+1)*E/N); tokens are data-parallel (T/N per rank); the router is replicated.

    route (top-k, histogram, stable permutation)            -- libb200q kernels, no host sync
    all_gather of the [E] histogram                         -- every rank derives all split sizes
    DISPATCH  all_to_all_single of the expert-sorted rows   -- NCCL over NVLink 5 / NVSwitch
    regroup (source-major -> expert-major), grouped INT4 GEMMs (w1||w3, SiLU-gate, w2), un-regroup
    COMBINE   all_to_all_single back, weighted sum of each token's k rows

The only host synchronisation per layer is the copy of the N x E histogram (needed for the NCCL
split sizes).  All arithmetic runs in the `ops` object: `CudaOps` (libb200q, the product) by default;
the CPU tests inject an oracle-backed stand-in to exercise this host logic under gloo.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .routing import DeviceRouting, route


def dispatch_plan(counts_all: np.ndarray, rank: int, world: int):
    """counts_all [N, E]: tokens-per-expert histogram of every rank.  Returns
    (send_splits, recv_splits, regroup, inverse, local_offsets):
      send_splits[r']  rows this rank sends to rank r' (its expert-sorted rows are already grouped by
                       destination because experts are owned in contiguous blocks);
      recv_splits[s]   rows received from rank s, ordered (source s, expert e, sender's stable order);
      regroup          gather index: rows_by_expert = received[regroup] is ordered (expert, source);
      inverse          gather index back: received_order = rows_by_expert[inverse];
      local_offsets    [E_loc + 1] exclusive offsets of the local experts in rows_by_expert."""
    n, E = counts_all.shape
    assert n == world and E % world == 0
    e_loc = E // world
    mine = counts_all[:, rank * e_loc:(rank + 1) * e_loc].astype(np.int64)          # [N, E_loc]
    send_splits = counts_all[rank].reshape(world, e_loc).sum(axis=1).astype(np.int64)
    recv_splits = mine.sum(axis=1)
    src_start = np.concatenate([[0], np.cumsum(mine.reshape(-1))[:-1]]).reshape(world, e_loc)
    per_expert = mine.sum(axis=0)
    local_offsets = np.concatenate([[0], np.cumsum(per_expert)]).astype(np.int32)
    regroup = np.empty(int(recv_splits.sum()), dtype=np.int32)
    pos = 0
    for e in range(e_loc):
        for s in range(world):
            c = int(mine[s, e])
            regroup[pos:pos + c] = np.arange(src_start[s, e], src_start[s, e] + c, dtype=np.int32)
            pos += c
    inverse = np.empty_like(regroup)
    inverse[regroup] = np.arange(regroup.size, dtype=np.int32)
    return send_splits.tolist(), recv_splits.tolist(), regroup, inverse, local_offsets


class CudaOps:
    """libb200q kernels (the product path)."""

    def route(self, logits, top_k) -> DeviceRouting:
        return route(logits, top_k)

    def gather(self, x, index, k):
        return _lib.moe_gather_rows(x, index, k)

    def experts(self, moe, xs, offsets):
        return moe.forward_grouped(xs, offsets)

    def combine(self, y, inv_perm, weights, k):
        return _lib.moe_combine(y, inv_perm, weights, k, out_dtype=torch.float32)


class ExpertParallelMoE(torch.nn.Module):
    """`local_moe` holds this rank's E/N experts (a QuantizedMoE, gated or single-projection)."""

    def __init__(self, local_moe, num_experts: int, top_k: int = 2, group=None, ops=None):
        super().__init__()
        self.local_moe = local_moe
        self.num_experts = num_experts
        self.top_k = top_k
        self.group = group
        self.ops = ops if ops is not None else CudaOps()
        self.last_stats = {}

    def forward(self, x: torch.Tensor, router_logits: torch.Tensor) -> torch.Tensor:
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        k, E, ops = self.top_k, self.num_experts, self.ops
        dr = ops.route(router_logits, k)
        xs = ops.gather(x, dr.sorted_slot, k)                     # rows sorted by global expert id
        if world == 1:
            y = ops.experts(self.local_moe, xs, dr.offsets)
            return ops.combine(y, dr.inv_perm, dr.expert_weights, k)
        counts_all = torch.empty((world * E,), dtype=torch.int32, device=dr.counts.device)
        dist.all_gather_into_tensor(counts_all, dr.counts.contiguous(), group=self.group)
        send_splits, recv_splits, regroup, inverse, local_offsets = dispatch_plan(
            counts_all.cpu().numpy().reshape(world, E), rank, world)
        dev = x.device
        recv = torch.empty((sum(recv_splits), x.shape[1]), dtype=x.dtype, device=dev)
        dist.all_to_all_single(recv, xs, recv_splits, send_splits, group=self.group)          # DISPATCH
        regroup_t = torch.from_numpy(regroup).to(dev, non_blocking=True)
        offsets_t = torch.from_numpy(local_offsets).to(dev, non_blocking=True)
        xg = ops.gather(recv, regroup_t, 1)                       # expert-major rows of the local experts
        yg = ops.experts(self.local_moe, xg, offsets_t)
        inverse_t = torch.from_numpy(inverse).to(dev, non_blocking=True)
        yr = ops.gather(yg, inverse_t, 1)                         # back to the order the rows arrived in
        back = torch.empty((xs.shape[0], yg.shape[1]), dtype=yg.dtype, device=dev)
        dist.all_to_all_single(back, yr, send_splits, recv_splits, group=self.group)          # COMBINE
        self.last_stats = {"sent_rows": int(sum(send_splits)), "recv_rows": int(sum(recv_splits)),
                           "bytes_out": int(sum(send_splits) - send_splits[rank]) * x.shape[1] * x.element_size()}
        return ops.combine(back, dr.inv_perm, dr.expert_weights, k)


def shard_experts(num_experts: int, rank: int, world: int) -> List[int]:
    assert num_experts % world == 0, "the number of experts must divide evenly over the ranks"
    per = num_experts // world
    return list(range(rank * per, (rank + 1) * per))
