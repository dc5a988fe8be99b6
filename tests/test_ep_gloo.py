"""Expert-parallel host logic on CPU: world_size 2, gloo.  The arithmetic is injected (an oracle-backed
stand-in for the CUDA kernels); what is tested is the plan / split-size / all-to-all plumbing of
ep.ExpertParallelMoE -- with and without replicated (hot) experts -- against the unsharded oracle layer."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_ep_plan_host_covers_every_received_row_once(pkg):
    """The numpy statement of b200q_ep_plan: split sizes add up, the ranges tile the received rows exactly, in
    (source, expert) order, with and without replicated experts; sender and receiver agree."""
    rng = np.random.default_rng(0)
    world, E = 4, 8
    for replicated in [(), (0,), (1, 6), tuple(range(8))]:
        counts = rng.integers(0, 7, size=(world, E))
        counts[2, 3] = 0
        plans = []
        for rank in range(world):
            local = pkg.local_expert_list(E, rank, world, replicated)
            lidx = np.full(E, -1)
            lidx[local] = np.arange(len(local))
            plans.append(pkg.ep_plan_host(counts, rank, world, replicated, lidx))
        for rank in range(world):
            send, recv, st, en, rx = plans[rank]
            assert send.sum() == counts[rank].sum()
            for peer in range(world):
                assert send[peer] == plans[peer][1][rank]          # what rank sends to peer is what peer expects from rank
            spans = sorted((int(a), int(b)) for a, b in zip(st, en) if b > a)
            assert sum(b - a for a, b in spans) == recv.sum()
            assert all(spans[i][1] <= spans[i + 1][0] for i in range(len(spans) - 1))
            assert rx.tolist() == [v // world for v in range(E * world)]
            # virtual ids: a permutation that groups experts by destination
            vid = pkg.virtual_ids(E, rank, world, replicated)
            assert sorted(vid.tolist()) == list(range(E))
            dest = [rank if e in replicated else e // (E // world) for e in range(E)]
            order = np.argsort(vid)
            assert all(dest[order[i]] <= dest[order[i + 1]] for i in range(E - 1))


class OracleOps:
    """CPU stand-in for the libb200q kernels, built on oracle/int4_oracle.py (test infrastructure); communication is
    torch.distributed over gloo (ep.TorchCommOps)."""

    def __init__(self, oracle, pkg):
        from b200q_pkg import pkg as p
        self.o, self.pkg = oracle, pkg
        self._comm = p.ep.TorchCommOps()

    def route(self, logits, top_k, vid):
        o = self.o
        idx, w = o.softmax_topk(logits.numpy(), top_k)
        E = logits.shape[1]
        v = np.arange(E) if vid is None else vid.numpy()
        idx_v = v[idx]
        counts_v, _ = o.histogram_offsets(idx_v, E)
        offsets = np.concatenate([[0], np.cumsum(counts_v)])
        sorted_slot, inv = o.permutation(idx_v)
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt)
        return self.pkg.DeviceRouting(t(idx, torch.int32), t(w, torch.float32), t(np.asarray(counts_v)[v], torch.int32),
                                      t(offsets, torch.int32), t(sorted_slot, torch.int32), t(inv, torch.int32), E, top_k)

    def gather(self, x, index, k):
        return x[(index.long() // k)]

    def _expert(self, moe, e, xe):
        o = self.o
        g = o.expert_forward(xe, *moe["w1"][e]); u = o.expert_forward(xe, *moe["w3"][e])
        return o.expert_forward((o.silu(g) * u).astype(np.float32), *moe["w2"][e])

    def experts(self, moe, xs, offsets):
        offs = offsets.numpy()
        out = np.zeros((xs.shape[0], moe["w2"][0][0].shape[0]), dtype=np.float32)
        for e in range(len(moe["w1"])):
            if offs[e + 1] > offs[e]:
                out[offs[e]:offs[e + 1]] = self._expert(moe, e, xs[offs[e]:offs[e + 1]].numpy())
        return torch.from_numpy(out)

    def experts_ranges(self, moe, rows, starts, ends, range_expert):
        out = np.zeros((rows.shape[0], moe["w2"][0][0].shape[0]), dtype=np.float32)
        for a, b, e in zip(np.asarray(starts), np.asarray(ends), np.asarray(range_expert)):
            if b > a:
                out[a:b] = self._expert(moe, int(e), rows[a:b].numpy())
        return torch.from_numpy(out)

    def combine(self, y, inv_perm, weights, k):
        T = inv_perm.numel() // k
        un = y[inv_perm.long()].reshape(T, k, -1)
        return (un * weights[:, :, None]).sum(dim=1)

    def allgather_counts(self, counts, world):
        return self._comm.allgather_counts(counts, world)

    def exchange(self, rows, send_rows, recv_rows):
        return self._comm.exchange(rows, send_rows, recv_rows)


def _worker(rank, world, port, tmp, replicated):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import int4_oracle as oracle
    from b200q_pkg import pkg
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    E, d, F, T, k = 4, 32, 64, 24, 2
    rng = np.random.default_rng(5)
    mk = lambda n, kk: [oracle.quantize_weights((rng.standard_normal((n, kk)) * 0.05).astype(np.float32)) for _ in range(E)]
    w1, w3, w2 = mk(F, d), mk(F, d), mk(d, F)
    x = rng.standard_normal((T, d)).astype(np.float32)
    logits = rng.standard_normal((T, E)).astype(np.float32)
    logits[:, 3] -= 4.0                       # an almost empty expert: ragged, possibly zero-sized blocks
    ref = oracle.moe_gated(x, logits, w1, w3, w2, k)
    mine = pkg.local_expert_list(E, rank, world, replicated)
    local = {"w1": [w1[e] for e in mine], "w3": [w3[e] for e in mine], "w2": [w2[e] for e in mine]}
    layer = pkg.ExpertParallelMoE(local, E, k, ops=OracleOps(oracle, pkg), replicated=replicated)
    lo, hi = rank * T // world, (rank + 1) * T // world
    out = layer(torch.from_numpy(x[lo:hi]), torch.from_numpy(logits[lo:hi]))
    err = float(np.abs(out.numpy() - ref[lo:hi]).max())
    with open(os.path.join(tmp, f"rank{rank}.txt"), "w") as f:
        f.write(f"{err}\n{layer.last_stats['sent_rows']}\n{layer.last_stats['bytes_out']}\n")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("replicated", [(), (0,), (1, 2)])
def test_expert_parallel_two_ranks_gloo(tmp_path, replicated):
    world = 2
    port = 29500 + (os.getpid() % 400) + 7 * len(replicated)
    mp.spawn(_worker, args=(world, port, str(tmp_path), replicated), nprocs=world, join=True)
    moved = 0
    for r in range(world):
        err, sent, bytes_out = open(tmp_path / f"rank{r}.txt").read().split()
        assert float(err) < 1e-5, f"rank {r}: max abs err {err}"
        assert int(sent) == 24          # T/world * k rows leave every rank's router
        moved += int(bytes_out)
    if len(replicated) == 0:
        assert moved > 0
