#!/bin/bash
# steady-state DRAM traffic of the decode kernel in the timed configuration: application replay, caches left alone
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
HINT=$1 timeout 900 ncu --replay-mode application --cache-control none --clock-control none \
  --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum \
  -k regex:gemv_dec -s 24 -c 24 --csv --log-file gpurun_out/r02_gemv_dec_steady_$2.csv python tools/ncu_dec.py 1 4096 11008 > gpurun_out/r02_ncu_steady_$2.log 2>&1
echo "rc=$?" >> gpurun_out/r02_ncu_steady_$2.log
