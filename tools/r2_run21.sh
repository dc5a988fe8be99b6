#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest21.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest21.log
CFG='[[1,4096,11008,{}],[2,4096,11008,{}],[4,4096,11008,{}],[8,4096,11008,{}],[16,4096,11008,{}],[1,11008,4096,{}],[2,11008,4096,{}],[4,11008,4096,{}],[8,11008,4096,{}],[1,4096,14336,{}],[2,4096,14336,{}],[4,4096,14336,{}],[8,4096,14336,{}],[1,14336,4096,{}],[2,14336,4096,{}],[4,14336,4096,{}],[8,14336,4096,{}]]'
timeout 300 python tools/dec_tune.py "$CFG" 2>&1 | cut -c1-150 > gpurun_out/r2_tune21.jsonl
timeout 300 python tools/moe_decode_graph.py > gpurun_out/r2_moegraph21.log 2>&1
timeout 600 python - > gpurun_out/r2_moedec21.log 2>&1 <<'PY'
import sys, json
sys.path.insert(0, '.')
import torch
from b200q_pkg import pkg
import bench_moe
from bench import measured_peaks
peaks, _ = measured_peaks()
dev = torch.device('cuda', 0)
print(json.dumps(bench_moe.moe_decode(torch, pkg, dev, float(peaks['hbm_gbs']))))
PY
