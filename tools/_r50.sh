cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest50.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest50.log
timeout 900 python bench.py > gpurun_out/r2_bench50.json 2> gpurun_out/r2_bench50.err
echo "bench rc=$?" >> gpurun_out/r2_bench50.err
