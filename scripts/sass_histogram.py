"""SASS opcode histogram per kernel of lib/libspmvb.so (cuobjdump -sass): the instructions that prove the data path -
UBLKCP (cp.async.bulk = TMA 1-D), SYNCS (mbarrier arrive / try_wait), REDG / RED (fused accum_results), LDS / LDG, the
fp64 / fp32 multiply and add, SHFL (segmented reduction keyed on the end-of-row bit).  Usage:
    python scripts/sass_histogram.py [lib] > profiles/r2/sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "spmv-fpga_b200", "lib", "libspmvb.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
kern, hist = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ["UBLKCP", "SYNCS", "RED", "ATOM", "LDS", "STS", "LDG", "STG", "DMUL", "DADD", "DFMA", "FMUL", "FADD", "FFMA", "SHFL", "ACQBULK",
       "UTMA", "BAR"]
for k in sorted(hist):
    if not any(t in k for t in ("spmv_", "zero_rows", "scale_", "sumsq", "cg_", "dot_kernel")):
        continue
    h = hist[k]
    total = sum(h.values())
    groups = collections.OrderedDict()
    for key in KEY:
        n = sum(v for op, v in h.items() if op.split(".")[0].startswith(key))
        if n:
            groups[key] = n
    print("%s: %d instructions" % (k, total))
    print("    " + "  ".join("%s %d" % kv for kv in groups.items()))
    detail = [(op, v) for op, v in h.items() if any(op.startswith(t) for t in ("UBLKCP", "SYNCS", "RED", "ATOMG", "LDG", "STG", "LDS"))]
    print("    " + "  ".join("%s x%d" % kv for kv in sorted(detail)))
