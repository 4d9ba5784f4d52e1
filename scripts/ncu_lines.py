"""Per CUDA source line: dynamic warp-instruction count and stall samples of the profiled kernel."""
import csv, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
h = rows[hi]
ie, isamp = h.index("Instructions Executed"), h.index("# Samples")
seen = set()
inst = defaultdict(int); samp = defaultdict(int); text = {}
for r in rows[hi + 1:]:
    if len(r) <= ie or r[0] == "Line No" or not r[ie].isdigit():
        continue
    key = (r[2])
    if key in seen:
        continue
    seen.add(key)
    ln = r[0]
    inst[ln] += int(r[ie]); samp[ln] += int(r[isamp]) if r[isamp].isdigit() else 0
    text[ln] = r[1]
tot = sum(inst.values())
print("total warp instrs", tot)
for ln, n in sorted(inst.items(), key=lambda kv: -kv[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("%6s %9d %5.1f%% samp %5d | %s" % (ln, n, 100.0 * n / tot, samp[ln], text[ln].strip()[:100]))
