"""bench_gemm.py (4096 -> 11008, bf16) with the token-tile height forced: BN=<n> python tools/bench_gemm_bn.py"""
import os, sys, runpy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200q_pkg import pkg
if os.environ.get("BN"): pkg._lib.tune("gemm_bn", int(os.environ["BN"]))
sys.argv = ["bench_gemm.py"]
runpy.run_path(os.path.join(ROOT, "tools", "bench_gemm.py"), run_name="__main__")
