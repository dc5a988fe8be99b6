#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CFG='[[1,4096,11008,{}],[2,4096,11008,{}],[4,4096,11008,{}],[1,11008,4096,{}],[1,4096,14336,{}],[4,4096,14336,{}],[1,14336,4096,{}],[4,14336,4096,{}],[1,4096,28672,{}]]'
for v in ring prechunk; do
  if [ $v = ring ]; then unset B200Q_LIB; else export B200Q_LIB=$PWD/tools/bin/libb200q_$v.so; fi
  echo "== $v" >> gpurun_out/r2_tune20.jsonl
  timeout 300 python tools/dec_tune.py "$CFG" 2>&1 | cut -c1-150 >> gpurun_out/r2_tune20.jsonl
done
unset B200Q_LIB
timeout 300 python tools/moe_decode_graph.py > gpurun_out/r2_moegraph20.log 2>&1
timeout 600 python - > gpurun_out/r2_moedec20.log 2>&1 <<'PY'
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
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_moedec_launches.csv python tools/moe_decode_once.py > gpurun_out/r2_ncu20.log 2>&1
