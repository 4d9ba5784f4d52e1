"""Print the hottest SASS instructions (by stall samples) of a kernel from an .ncu-rep (run where ncu is installed)."""
import csv, subprocess, sys
rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
ie, isamp, src = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
st = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
seen, d2 = set(), []
for r in rows[hi + 1:]:
    if len(r) <= max(ie, isamp, src) or r[0] in seen or r[0] == "Address":
        continue
    seen.add(r[0]); d2.append(r)
tot = sum(int(r[isamp]) for r in d2 if r[isamp].isdigit())
inst = sum(int(r[ie]) for r in d2 if r[ie].isdigit())
print("total samples", tot, "static instrs", len(d2), "warp instrs executed", inst)
agg = {}
for r in d2:
    for i in st:
        if r[i].isdigit():
            agg[h[i]] = agg.get(h[i], 0) + int(r[i])
print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for r in sorted([r for r in d2 if r[isamp].isdigit()], key=lambda r: -int(r[isamp]))[:n]:
    stalls = {h[i]: int(r[i]) for i in st if r[i].isdigit() and int(r[i]) > 0}
    best = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
    print(r[isamp], r[ie], r[src][:72], best)
