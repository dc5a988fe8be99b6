"""Expert-parallel host logic on CPU: world_size 2, gloo.  The arithmetic is injected (an oracle-backed
stand-in for the CUDA kernels); what is tested is the split-size / regroup / all-to-all plumbing of
ep.ExpertParallelMoE against the unsharded oracle layer."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_dispatch_plan_is_a_permutation(pkg):
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 7, size=(4, 8))
    counts[2, 3] = 0
    for rank in range(4):
        send, recv, regroup, inverse, offs = pkg.dispatch_plan(counts, rank, 4)
        assert send == counts[rank].reshape(4, 2).sum(axis=1).tolist()
        assert recv == counts[:, 2 * rank:2 * rank + 2].sum(axis=1).tolist()
        assert sorted(regroup.tolist()) == list(range(sum(recv)))
        assert np.array_equal(regroup[inverse], np.arange(sum(recv)))
        assert offs.tolist() == [0, counts[:, 2 * rank].sum(), counts[:, 2 * rank:2 * rank + 2].sum()]
        # rows_by_expert = received[regroup] is expert-major: the owning expert of every row is sorted
        owner = np.concatenate([np.repeat(np.arange(2), counts[s, 2 * rank:2 * rank + 2]) for s in range(4)])
        assert np.all(np.diff(owner[regroup]) >= 0)


class OracleOps:
    """CPU stand-in for the libb200q kernels, built on oracle/int4_oracle.py (test infrastructure)."""

    def __init__(self, oracle, pkg):
        self.o, self.pkg = oracle, pkg

    def route(self, logits, top_k):
        o = self.o
        idx, w = o.softmax_topk(logits.numpy(), top_k)
        E = logits.shape[1]
        counts, _ = o.histogram_offsets(idx, E)
        offsets = np.concatenate([[0], np.cumsum(counts)])
        sorted_slot, inv = o.permutation(idx)
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt)
        return self.pkg.DeviceRouting(t(idx, torch.int32), t(w, torch.float32), t(counts, torch.int32),
                                      t(offsets, torch.int32), t(sorted_slot, torch.int32), t(inv, torch.int32), E, top_k)

    def gather(self, x, index, k):
        return x[(index.long() // k)]

    def experts(self, moe, xs, offsets):
        o = self.o
        offs = offsets.numpy()
        out = np.zeros((xs.shape[0], moe["w2"][0][0].shape[0]), dtype=np.float32)
        for e in range(len(moe["w1"])):
            xe = xs[offs[e]:offs[e + 1]].numpy()
            if len(xe) == 0:
                continue
            g = o.expert_forward(xe, *moe["w1"][e]); u = o.expert_forward(xe, *moe["w3"][e])
            out[offs[e]:offs[e + 1]] = o.expert_forward((o.silu(g) * u).astype(np.float32), *moe["w2"][e])
        return torch.from_numpy(out)

    def combine(self, y, inv_perm, weights, k):
        T = inv_perm.numel() // k
        un = y[inv_perm.long()].reshape(T, k, -1)
        return (un * weights[:, :, None]).sum(dim=1)


def _worker(rank, world, port, tmp):
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
    mine = pkg.shard_experts(E, rank, world)
    local = {"w1": [w1[e] for e in mine], "w3": [w3[e] for e in mine], "w2": [w2[e] for e in mine]}
    layer = pkg.ExpertParallelMoE(local, E, k, ops=OracleOps(oracle, pkg))
    lo, hi = rank * T // world, (rank + 1) * T // world
    out = layer(torch.from_numpy(x[lo:hi]), torch.from_numpy(logits[lo:hi]))
    err = float(np.abs(out.numpy() - ref[lo:hi]).max())
    with open(os.path.join(tmp, f"rank{rank}.txt"), "w") as f:
        f.write(f"{err}\n{layer.last_stats['sent_rows']}\n")
    dist.barrier()
    dist.destroy_process_group()


def test_expert_parallel_two_ranks_gloo(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        err, sent = open(tmp_path / f"rank{r}.txt").read().split()
        assert float(err) < 1e-5, f"rank {r}: max abs err {err}"
        assert int(sent) == 24          # T/world * k rows leave every rank's router
