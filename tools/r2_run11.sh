#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest11.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest11.log
timeout 600 python bench.py --no-moe --no-sweeps --no-cpu > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err; echo "bench rc=$?" >> gpurun_out/r2_bench11.err
B200Q_NO_TORCH_EXT=1 timeout 600 python bench.py --no-moe --no-sweeps --no-cpu > gpurun_out/r2_bench11_ctypes.json 2>> gpurun_out/r2_bench11.err
timeout 300 python - > gpurun_out/r2_moedec11.log 2>&1 <<'PY'
import torch, json, sys
sys.path.insert(0, '.')
from b200q_pkg import pkg
from bench_moe import moe_decode
from bench import measured_peaks
print(json.dumps(moe_decode(torch, pkg, torch.device('cuda', 0), float(measured_peaks()[0]['hbm_gbs']))))
PY
