#!/bin/bash
# round-2 first GPU pass: new decode kernel sanity, decode tests, sweep in both TMA forms
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python - > gpurun_out/r2_sanity.log 2>&1 <<'PY'
import numpy as np, torch, sys, os
sys.path.insert(0, "oracle")
import int4_oracle as oracle
from b200q_pkg import pkg
rng = np.random.default_rng(0)
for (M, N, K, t3) in [(1, 64, 256, 0), (1, 11008, 4096, 0), (1, 11008, 4096, 1), (2, 4096, 11008, 0), (4, 11008, 4096, 0), (16, 11008, 4096, 0), (16, 4096, 11008, 1)]:
    packed = rng.integers(0, 256, size=(N, K // 2), dtype=np.uint8)
    s = (rng.random(N, dtype=np.float32) * 0.01 + 0.001).astype(np.float32)
    z = rng.integers(0, 16, size=N).astype(np.float32)
    x = rng.standard_normal((M, K), dtype=np.float32)
    pkg._lib.tune("gemv_tma3d", t3)
    try:
        y = pkg._lib.linear_fwd(torch.from_numpy(x).cuda(), torch.from_numpy(packed).cuda(), torch.from_numpy(s).cuda(), torch.from_numpy(z).cuda())
        torch.cuda.synchronize()
        y = y.cpu().numpy()
        rows = np.arange(0, N, max(1, N // 200))
        ref = oracle.reference_quantized_linear(x, packed[rows], s[rows], z[rows], acc=np.float64)
        print(M, N, K, "tma3d", t3, "rel err", float(np.abs(y[:, rows] - ref).max() / np.abs(ref).max()), flush=True)
    except Exception as e:
        print(M, N, K, "tma3d", t3, "ERROR", repr(e)[:300], flush=True)
PY
echo "sanity rc=$?" >> gpurun_out/r2_sanity.log
timeout 1200 python -m pytest tests/test_gpu_linear.py tests/test_gpu_parity_r2.py -x -q -m gpu > gpurun_out/r2_pytest1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
timeout 300 python tools/sweep_all.py decode > gpurun_out/r2_sweep_2d.jsonl 2> gpurun_out/r2_sweep_2d.err
B200Q_TUNE=gemv_tma3d=1 timeout 300 python tools/sweep_all.py decode > gpurun_out/r2_sweep_3d.jsonl 2> gpurun_out/r2_sweep_3d.err
tail -3 gpurun_out/r2_sanity.log gpurun_out/r2_pytest1.log
cat gpurun_out/r2_sweep_2d.jsonl gpurun_out/r2_sweep_3d.jsonl
