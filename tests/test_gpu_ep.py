"""Expert parallelism on the GPU: the device-side plan (b200q_ep_plan) against its numpy statement, the mapped
grouped GEMM against the oracle, and -- on a box with >= 2 GPUs -- the whole CUDA + NCCL layer
(b200q_ep_* communicator, dispatch, grouped GEMMs over the received ranges, combine) against the unsharded layer and
the float64 oracle.  Single-GPU boxes run the first two and skip the third."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("world,E,replicated", [(2, 8, ()), (4, 8, (0,)), (8, 8, (0, 1)), (8, 8, ()), (4, 16, (3, 9, 15)), (8, 64, ())])
def test_ep_plan_kernel_matches_numpy_statement(pkg, world, E, replicated):
    rng = np.random.default_rng(world * 100 + E)
    counts = rng.integers(0, 900, size=(world, E)).astype(np.int32)
    counts[0, 1] = 0
    lib = pkg._lib.load()
    for rank in range(world):
        local = pkg.local_expert_list(E, rank, world, replicated)
        lidx = np.full(E, -1, dtype=np.int32)
        lidx[local] = np.arange(len(local))
        rep = np.zeros(E, dtype=np.int32)
        rep[list(replicated)] = 1
        send, recv, st, en, rx = pkg.ep_plan_host(counts, rank, world, replicated, lidx)
        C, R, L = cuda(counts.reshape(-1)), cuda(rep), cuda(lidx)
        splits = torch.full((2 * world,), -7, dtype=torch.int32, device="cuda")
        starts = torch.full((E * world,), -7, dtype=torch.int32, device="cuda")
        ends, rexp = starts.clone(), starts.clone()
        pkg._lib.check(lib.b200q_ep_plan(C.data_ptr(), world, rank, E, R.data_ptr() if replicated else None, L.data_ptr(),
                                         splits.data_ptr(), starts.data_ptr(), ends.data_ptr(), rexp.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream), "plan")
        s = splits.cpu().numpy()
        assert np.array_equal(s[:world], send) and np.array_equal(s[world:], recv)
        assert np.array_equal(starts.cpu().numpy(), st) and np.array_equal(ends.cpu().numpy(), en)
        assert np.array_equal(rexp.cpu().numpy(), rx)


@pytest.mark.parametrize("K,N,gated", [(256, 128, False), (256, 256, True), (90, 64, False), (1024, 512, True)])
def test_mapped_grouped_gemm_matches_oracle(oracle, pkg, K, N, gated):
    """Ranges with an explicit weight expert, unsorted, with gaps and empty ranges: every covered row against the
    float64 oracle, uncovered rows untouched (the Python wrapper hands in zeros)."""
    rng = np.random.default_rng(K + N)
    E, R = 3, 300
    packed = rng.integers(0, 256, size=(E, N, K // 2), dtype=np.uint8)
    scales = (rng.random((E, N), dtype=np.float32) * 0.01 + 0.001).astype(np.float32)
    zps = rng.integers(0, 16, size=(E, N)).astype(np.float32)
    xs = rng.standard_normal((R, K), dtype=np.float32)
    starts = np.array([200, 0, 40, 40, 130, 290], dtype=np.int32)
    ends = np.array([280, 37, 40, 120, 131, 300], dtype=np.int32)
    rexp = np.array([2, 0, 1, 1, 0, 2], dtype=np.int32)
    y = pkg._lib.moe_grouped_fwd_mapped(cuda(xs), cuda(packed), cuda(scales), cuda(zps), cuda(starts), cuda(ends), cuda(rexp),
                                        gated=gated, out_dtype=torch.float32).cpu().numpy()
    covered = np.zeros(R, dtype=bool)
    for a, b, e in zip(starts, ends, rexp):
        if b <= a:
            continue
        covered[a:b] = True
        full = oracle.reference_quantized_linear(xs[a:b], packed[e], scales[e], zps[e], acc=np.float64)
        ref = oracle.silu(full[:, 0::2]) * full[:, 1::2] if gated else full
        assert np.abs(y[a:b] - ref).max() <= 2e-4 * np.abs(ref).max() + 1e-6
    assert not y[~covered].any()


def _ep_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import int4_oracle as oracle
    import torch.distributed as dist
    from b200q_pkg import pkg
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    E, d, F, T, k = 8, 256, 512, 96 * world, 2
    rng = np.random.default_rng(11)
    mk = lambda n, kk: [(rng.standard_normal((n, kk)) * 0.05).astype(np.float16) for _ in range(E)]
    w1, w3, w2 = mk(F, d), mk(F, d), mk(d, F)
    x = rng.standard_normal((T, d)).astype(np.float32)
    results = {}
    for routing in ("random", "skewed"):
        logits = rng.standard_normal((T, E)).astype(np.float32)
        if routing == "skewed":
            logits += np.log(1.0 / (np.arange(E) + 1.0)).astype(np.float32) * 2.0
        q = lambda ws: [oracle.quantize_weights(a.astype(np.float32)) for a in ws]
        ref = oracle.moe_gated(x, logits, q(w1), q(w3), q(w2), k, acc=np.float64)
        t = lambda ws, ids: [torch.from_numpy(ws[e]).to(dev) for e in ids]
        full = pkg.QuantizedMoE.from_gated_fp16_weights(t(w1, range(E)), t(w3, range(E)), t(w2, range(E)))
        lo, hi = rank * T // world, (rank + 1) * T // world
        xd, ld = torch.from_numpy(x[lo:hi]).to(dev), torch.from_numpy(logits[lo:hi]).to(dev)
        unsharded = full.forward_routed(torch.from_numpy(x).to(dev), torch.from_numpy(logits).to(dev), top_k=k)[lo:hi]
        for replicated in ((), (0, 1)):
            mine = pkg.local_expert_list(E, rank, world, replicated)
            local = pkg.QuantizedMoE.from_gated_fp16_weights(t(w1, mine), t(w3, mine), t(w2, mine))
            for c_comm in (True, False):
                layer = pkg.ExpertParallelMoE(local, E, k, replicated=replicated, ops=pkg.ep.CudaOps(None, c_comm=c_comm))
                out = layer(xd, ld)
                out2 = layer(xd, ld)
                key = f"{routing}/rep{len(replicated)}/{'b200q' if c_comm else 'torch'}"
                results[key] = (float((out - unsharded).abs().max()), float(np.abs(out.cpu().numpy() - ref[lo:hi]).max()),
                                bool(torch.equal(out, out2)), layer.last_stats["bytes_out"])
                layer.ops.close()
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        for kx, v in results.items():
            f.write(f"{kx} {v[0]} {v[1]} {int(v[2])} {v[3]}\n")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (NCCL expert parallelism)")
def test_expert_parallel_cuda_nccl_matches_unsharded_and_oracle(tmp_path):
    """VERDICT r1 'parity hole 1': the path that produces the scaling numbers -- CudaOps + NCCL dispatch / combine --
    against the unsharded forward_routed on the same tokens and against the float64 oracle, random and skewed routing,
    with and without replicated hot experts, through libb200q's own communicator and through torch.distributed."""
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    port = 29700 + (os.getpid() % 200)
    mp.spawn(_ep_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        lines = open(tmp_path / f"rank{r}.txt").read().strip().splitlines()
        assert len(lines) == 8
        for line in lines:
            key, d_unsharded, d_oracle, same, bytes_out = line.split()
            assert float(d_unsharded) <= 2e-5, f"rank {r} {key}: differs from the unsharded layer by {d_unsharded}"
            assert float(d_oracle) <= 2e-4, f"rank {r} {key}: differs from the float64 oracle by {d_oracle}"
            assert same == "1", f"rank {r} {key}: not deterministic"
