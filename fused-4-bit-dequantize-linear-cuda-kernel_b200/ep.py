"""Expert-parallel INT4 MoE layer over the GPUs of one NVSwitch box (one process per GPU).

Not in the reference (it is single-GPU); required by the north star.  Rank r owns
experts [r*E/N, (r+1)*E/N); tokens are data-parallel (T/N per rank); the router is replicated.  Experts named in
`replicated` are held by EVERY rank and serve their tokens where those live (hot-expert replication: the
reference's default "skewed" routing recipe sends 47 % of the assignments to expert 0, routing.py:54-66).

    route: top-k, then the stable permutation by (destination rank, expert)      -- libb200q kernels, no host sync
    all-gather of the [E] histogram                                              -- NCCL, 32 bytes per rank
    b200q_ep_plan: split sizes + one [start, end) range per (local expert, source rank) of the rows to receive
    DISPATCH  variable-split all-to-all of the sorted rows                       -- NCCL over NVLink 5 / NVSwitch
    grouped INT4 GEMMs over those ranges (w1||w3 with the SiLU-gate, w2)         -- rows are never re-sorted
    COMBINE   all-to-all back, weighted sum of each token's k rows

The only host synchronisation per layer is the copy of the 2 N split sizes (NCCL's send / recv counts are host
arguments).  All arithmetic and communication runs in the `ops` object: `CudaOps` (libb200q + its NCCL entry points,
the product) by default; the CPU tests inject an oracle-backed stand-in to exercise this host logic under gloo.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .routing import DeviceRouting


def shard_experts(num_experts: int, rank: int, world: int) -> List[int]:
    assert num_experts % world == 0, "the number of experts must divide evenly over the ranks"
    per = num_experts // world
    return list(range(rank * per, (rank + 1) * per))


def local_expert_list(num_experts: int, rank: int, world: int, replicated: Sequence[int] = ()) -> List[int]:
    """Global ids of the experts a rank holds, in the order of its weight tensor: the owned block, then the
    replicated experts it does not own."""
    own = shard_experts(num_experts, rank, world)
    return own + [e for e in sorted(set(replicated)) if e not in own]


def virtual_ids(num_experts: int, rank: int, world: int, replicated: Sequence[int] = ()) -> np.ndarray:
    """vid[e] = position of expert e when this rank's assignments are sorted by (destination rank, expert);
    a replicated expert's destination is the rank itself."""
    per = num_experts // world
    rep = set(replicated)
    order = sorted(range(num_experts), key=lambda e: (rank if e in rep else e // per, e))
    vid = np.empty(num_experts, dtype=np.int64)
    vid[order] = np.arange(num_experts)
    return vid


def ep_plan_host(counts_all: np.ndarray, rank: int, world: int, replicated: Sequence[int], local_index: np.ndarray):
    """numpy statement of b200q_ep_plan (csrc/ep.cu): counts_all [world, E] in original expert ids ->
    (send_rows [world], recv_rows [world], range_starts, range_ends, range_expert [E * world])."""
    n, E = counts_all.shape
    assert n == world and E % world == 0
    per = E // world
    rep = np.zeros(E, dtype=bool)
    rep[list(replicated)] = True
    owner = np.arange(E) // per
    dest_me = np.where(rep, rank, owner)
    send = np.array([counts_all[rank][dest_me == r].sum() for r in range(world)], dtype=np.int64)
    starts = np.zeros(E * world, dtype=np.int32)
    ends = np.zeros(E * world, dtype=np.int32)
    recv = np.zeros(world, dtype=np.int64)
    off = 0
    for s in range(world):
        mine = np.where(rep, s, owner) == rank
        for e in np.nonzero(mine)[0]:
            c = int(counts_all[s, e])
            li = int(local_index[e])
            if li >= 0:
                starts[li * world + s] = off
                ends[li * world + s] = off + c
            off += c
            recv[s] += c
    for li in range(E):                      # adjacent ranges of one expert are merged (fewer ragged last tiles)
        cur = -1
        for s in range(world):
            v = li * world + s
            if ends[v] <= starts[v]:
                continue
            if cur >= 0 and ends[cur] == starts[v]:
                ends[cur] = ends[v]
                starts[v] = ends[v] = 0
            else:
                cur = v
    return send, recv, starts, ends, (np.arange(E * world) // world).astype(np.int32)


def dispatch_plan(counts_all: np.ndarray, rank: int, world: int):
    """Round-1 form (kept for callers of the old name): (send_splits, recv_splits, range_starts, range_ends,
    range_expert) without replication, local experts = the owned block."""
    E = counts_all.shape[1]
    lidx = np.full(E, -1, dtype=np.int64)
    lidx[shard_experts(E, rank, world)] = np.arange(E // world)
    send, recv, st, en, rx = ep_plan_host(np.asarray(counts_all), rank, world, (), lidx)
    return send.tolist(), recv.tolist(), st, en, rx


class TorchCommOps:
    """Communication through torch.distributed (gloo on the CPU tests, NCCL if asked for)."""

    group = None

    def allgather_counts(self, counts: torch.Tensor, world: int) -> torch.Tensor:
        out = torch.empty((world * counts.numel(),), dtype=torch.int32, device=counts.device)
        dist.all_gather_into_tensor(out, counts.contiguous(), group=self.group)
        return out

    def exchange(self, rows: torch.Tensor, send_rows, recv_rows) -> torch.Tensor:
        out = torch.empty((int(sum(recv_rows)), rows.shape[1]), dtype=rows.dtype, device=rows.device)
        dist.all_to_all_single(out, rows.contiguous(), [int(v) for v in recv_rows], [int(v) for v in send_rows], group=self.group)
        return out


class CudaOps(TorchCommOps):
    """libb200q kernels and NCCL entry points (the product path)."""

    def __init__(self, group=None, c_comm: bool = True):
        self.group = group
        self.c_comm = c_comm
        self._comm = None
        self._pinned = None

    # ---- arithmetic
    def route(self, logits, top_k, vid: Optional[torch.Tensor]) -> DeviceRouting:
        _lib.require_cuda(logits, "logits")
        logits = logits.to(torch.float32).contiguous()
        E = logits.shape[1]
        idx, w, counts, offsets, sorted_slot, inv_perm = _lib.moe_route(logits, top_k, vid)     # one C call
        return DeviceRouting(idx, w, counts, offsets, sorted_slot, inv_perm, E, top_k)

    def gather(self, x, index, k):
        return _lib.moe_gather_rows(x, index, k)

    def experts(self, moe, xs, offsets):
        return moe.forward_grouped(xs, offsets)

    def experts_ranges(self, moe, rows, starts, ends, range_expert):
        return moe.forward_ranges(rows, starts, ends, range_expert, all_rows_covered=True)   # every received row has an expert here

    def combine(self, y, inv_perm, weights, k):
        return _lib.moe_combine(y, inv_perm, weights, k, out_dtype=torch.float32)

    def plan(self, counts_all, rank, world, E, replicated_t, local_index_t):
        lib = _lib.load()
        dev = counts_all.device
        with torch.cuda.device(dev):
            splits = torch.empty((2 * world,), dtype=torch.int32, device=dev)
            starts = torch.empty((E * world,), dtype=torch.int32, device=dev)
            ends = torch.empty_like(starts)
            rexp = torch.empty_like(starts)
            _lib.check(lib.b200q_ep_plan(counts_all.data_ptr(), world, rank, E,
                                         replicated_t.data_ptr() if replicated_t is not None else None,
                                         local_index_t.data_ptr(), splits.data_ptr(), starts.data_ptr(), ends.data_ptr(),
                                         rexp.data_ptr(), _lib.stream_ptr(dev)), "b200q_ep_plan")
        return splits, starts, ends, rexp

    def host_splits(self, splits: torch.Tensor, world: int):
        """The one host synchronisation of the layer: 2 N int32 through pinned memory."""
        if self._pinned is None or self._pinned.numel() != splits.numel():
            self._pinned = torch.empty(splits.numel(), dtype=torch.int32).pin_memory()
        self._pinned.copy_(splits, non_blocking=True)
        torch.cuda.current_stream(splits.device).synchronize()
        h = self._pinned.tolist()
        return h[:world], h[world:]

    # ---- communication: libb200q's own NCCL communicator (b200q_ep_*), torch.distributed only hands the id around
    def _ensure_comm(self, dev):
        if self._comm is None:
            lib = _lib.load()
            world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
            uid = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                _lib.check(lib.b200q_ep_unique_id(uid.data_ptr()), "b200q_ep_unique_id")
            uid_d = uid.to(dev)
            dist.broadcast(uid_d, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
            uid = uid_d.cpu()
            comm = ctypes.c_void_p()
            with torch.cuda.device(dev):
                _lib.check(lib.b200q_ep_comm_create(uid.data_ptr(), rank, world, ctypes.byref(comm)), "b200q_ep_comm_create")
            self._comm = comm
        return self._comm

    def allgather_counts(self, counts, world):
        if not self.c_comm:
            return super().allgather_counts(counts, world)
        lib = _lib.load()
        dev = counts.device
        comm = self._ensure_comm(dev)
        counts = counts.contiguous()
        out = torch.empty((world * counts.numel(),), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.b200q_ep_allgather_i32(comm, counts.data_ptr(), out.data_ptr(), counts.numel(), _lib.stream_ptr(dev)),
                       "b200q_ep_allgather_i32")
        return out

    def exchange(self, rows, send_rows, recv_rows):
        if not self.c_comm:
            return super().exchange(rows, send_rows, recv_rows)
        lib = _lib.load()
        dev = rows.device
        comm = self._ensure_comm(dev)
        rows = rows.contiguous()
        world = len(send_rows)
        out = torch.empty((int(sum(recv_rows)), rows.shape[1]), dtype=rows.dtype, device=dev)
        hs = (ctypes.c_int64 * world)(*[int(v) for v in send_rows])
        hr = (ctypes.c_int64 * world)(*[int(v) for v in recv_rows])
        with torch.cuda.device(dev):
            _lib.check(lib.b200q_ep_exchange(comm, world, rows.data_ptr(), hs, out.data_ptr(), hr,
                                             rows.shape[1] * rows.element_size(), _lib.stream_ptr(dev)), "b200q_ep_exchange")
        return out

    def close(self):
        if self._comm is not None:
            _lib.load().b200q_ep_comm_destroy(self._comm)
            self._comm = None


class ExpertParallelMoE(torch.nn.Module):
    """`local_moe` holds this rank's experts (a QuantizedMoE, gated or single-projection) in the order of
    `local_expert_list(num_experts, rank, world, replicated)`."""

    def __init__(self, local_moe, num_experts: int, top_k: int = 2, group=None, ops=None, replicated: Sequence[int] = ()):
        super().__init__()
        self.local_moe = local_moe
        self.num_experts = num_experts
        self.top_k = top_k
        self.group = group
        self.replicated = tuple(sorted(set(int(e) for e in replicated)))
        self.ops = ops if ops is not None else CudaOps(group)
        self.last_stats = {}
        self.profile = False         # True: CUDA events around the phases of the next forward -> last_stats["phases_ms"]
        self._static = None          # (device, vid, replicated mask, local index) as device tensors

    def _tables(self, dev, rank, world):
        if self._static is None or self._static[0] != dev:
            E = self.num_experts
            vid = torch.from_numpy(virtual_ids(E, rank, world, self.replicated)).to(device=dev, dtype=torch.int32)
            rep = torch.zeros(E, dtype=torch.int32)
            rep[list(self.replicated)] = 1
            lidx = torch.full((E,), -1, dtype=torch.int32)
            for i, e in enumerate(local_expert_list(E, rank, world, self.replicated)):
                lidx[e] = i
            self._static = (dev, vid, rep.to(dev) if self.replicated else None, lidx.to(dev), rep.numpy(), lidx.numpy())
        return self._static

    def forward(self, x: torch.Tensor, router_logits: torch.Tensor) -> torch.Tensor:
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        k, E, ops = self.top_k, self.num_experts, self.ops
        if world == 1:
            dr = ops.route(router_logits, k, None)
            xs = ops.gather(x, dr.sorted_slot, k)
            y = ops.experts(self.local_moe, xs, dr.offsets)
            return ops.combine(y, dr.inv_perm, dr.expert_weights, k)
        _, vid, rep_t, lidx_t, rep_h, lidx_h = self._tables(x.device, rank, world)
        marks = []

        def mark(name):
            if self.profile and x.is_cuda:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((name, ev))

        mark("start")
        dr = ops.route(router_logits, k, vid)                     # rows sorted by (destination rank, expert)
        xs = ops.gather(x, dr.sorted_slot, k)
        mark("route+gather")
        counts_all = ops.allgather_counts(dr.counts, world)       # [world * E], original expert ids
        if hasattr(ops, "plan"):
            splits, starts, ends, rexp = ops.plan(counts_all, rank, world, E, rep_t, lidx_t)
            send_rows, recv_rows = ops.host_splits(splits, world)
        else:                                                     # CPU stand-in: the numpy statement of the same plan
            send_rows, recv_rows, starts, ends, rexp = ep_plan_host(counts_all.cpu().numpy().reshape(world, E), rank, world,
                                                                    self.replicated, lidx_h)
        mark("histogram all-gather + plan + host sync")
        recv = ops.exchange(xs, send_rows, recv_rows)             # DISPATCH
        mark("dispatch all-to-all")
        yg = ops.experts_ranges(self.local_moe, recv, starts, ends, rexp)
        mark("expert GEMMs")
        back = ops.exchange(yg, recv_rows, send_rows)             # COMBINE
        mark("combine all-to-all")
        out = ops.combine(back, dr.inv_perm, dr.expert_weights, k)
        mark("weighted combine")
        self.last_stats = {"sent_rows": int(sum(send_rows)), "recv_rows": int(sum(recv_rows)),
                           "bytes_out": int(sum(send_rows) - send_rows[rank]) * x.shape[1] * x.element_size()}
        if marks:
            torch.cuda.synchronize(x.device)
            self.last_stats["phases_ms"] = {marks[i][0]: round(marks[i - 1][1].elapsed_time(marks[i][1]), 4) for i in range(1, len(marks))}
        return out
