#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CFG='[[1,4096,11008,{}],[2,4096,11008,{}],[4,4096,11008,{}],[16,4096,11008,{}],[1,11008,4096,{}],[4,11008,4096,{}],[1,4096,14336,{}],[1,14336,4096,{}]]'
for v in ring chunks prechunk ring; do
  if [ $v = ring ]; then unset B200Q_LIB; else export B200Q_LIB=$PWD/tools/bin/libb200q_$v.so; fi
  echo "== $v" >> gpurun_out/r2_tune19.jsonl
  timeout 300 python tools/dec_tune.py "$CFG" 2>&1 | cut -c1-150 >> gpurun_out/r2_tune19.jsonl
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_throttle_reasons.active,power.draw --format=csv >> gpurun_out/r2_tune19.jsonl
