"""Prints the most-sampled SASS instructions (with stall reasons) of one kernel in an ncu report.
usage: python tools/ncu_hot.py report.ncu-rep launch_index [top_n]"""
import csv
import subprocess
import sys

rep, idx = sys.argv[1], int(sys.argv[2])
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:150])
hdr = rows[1]
data = []
for d in rows[2:]:
    if len(d) != len(hdr) or not d[hdr.index("# Samples")].isdigit():
        if data:
            break  # a second table (another view) follows
        continue
    data.append(d)
isrc, isamp = hdr.index("Source"), hdr.index("# Samples")
stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(d[isamp]) for d in data)
print("total samples", tot)
agg = {}
for d in data:
    for i, h in stalls:
        agg[h] = agg.get(h, 0) + int(d[i] or 0)
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for d in sorted(data, key=lambda d: -int(d[isamp]))[:topn]:
    why = {h[6:]: int(d[i]) for i, h in stalls if int(d[i] or 0)}
    print(f"{int(d[isamp]):6d} {100*int(d[isamp])/max(tot,1):5.1f}%  {d[isrc].strip()[:70]:70s} {why}")
