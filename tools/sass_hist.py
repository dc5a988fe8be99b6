"""SASS opcode histogram of every kernel in the shipped library (cuobjdump -sass): the evidence that the hot kernels use
the instructions DESIGN.md says they do (IMMA / LDSM / UTMALDG in the decode kernel, UTCHMMA / STTM / LDTM / UTMALDG in the
tcgen05 GEMM).   python tools/sass_hist.py [lib.so] > profiles/r02_sass_histogram.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "fused-4-bit-dequantize-linear-cuda-kernel_b200", "libb200q.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kernels, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[0-9T]\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
    if m and cur is not None:
        op = m.group(1)
        base = op.split(".")[0]
        # keep the qualifiers that identify the data path
        key = op if base in ("IMMA", "HMMA", "UTCHMMA", "UTCIMMA", "UTCQMMA", "LDSM", "UTMALDG", "UTMAPF", "UBLKCP", "UBLKPF", "STTM", "LDTM", "SYNCS",
                             "UTCBAR", "REDUX", "CREDUX", "ATOMS", "RED", "ACQBULK") else base
        cur[key] += 1
print(f"# {os.path.relpath(lib, ROOT)}: SASS opcode histogram per kernel (static instruction counts)")
KEY = ("IMMA", "HMMA", "UTC", "LDSM", "UTMA", "UBLK", "STTM", "LDTM", "SYNCS", "REDUX", "CREDUX", "ATOMS", "RED", "ACQBULK")
for name, c in kernels.items():
    dn = demangle(name)
    dn = re.sub(r"\(anonymous namespace\)::", "", dn)
    dn = re.sub(r"\(.*", "", dn)
    total = sum(c.values())
    print(f"\n## {dn}   ({total} instructions)")
    marked = {k: v for k, v in c.items() if k.startswith(KEY)}
    if marked:
        print("   data path: " + ", ".join(f"{k} x{v}" for k, v in sorted(marked.items())))
    print("   top: " + ", ".join(f"{k} {v}" for k, v in c.most_common(12)))
