#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest8.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest8.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke8.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke8.log
timeout 900 python bench.py > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err; echo "bench rc=$?" >> gpurun_out/r2_bench8.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench8_ref.json 2>> gpurun_out/r2_bench8.err
