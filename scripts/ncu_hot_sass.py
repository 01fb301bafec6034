"""Top stall sites of a kernel from an .ncu-rep (source page, SASS view): python scripts/ncu_hot_sass.py rep [kernel-index] [n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
# split per kernel
blocks = out.split('"Kernel Name",')
want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
blk = blocks[1 + want]
lines = blk.splitlines()
print("kernel:", lines[0][:100])
rows = list(csv.reader(lines[1:]))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]] or 0) for r in rows[1:] if len(r) == len(hdr))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: 0 for h in stall_cols}
data = []
for k, r in enumerate(rows[1:]):
    if len(r) != len(hdr):
        continue
    s = int(r[ix["# Samples"]] or 0)
    for h in stall_cols:
        agg[h] += int(r[ix[h]] or 0)
    data.append((s, k, r))
print("total samples", tot, {h.replace("stall_", ""): round(100 * v / max(tot, 1), 1) for h, v in sorted(agg.items(), key=lambda x: -x[1])[:9]})
for s, k, r in sorted(data, reverse=True)[:n]:
    top = sorted(((int(r[ix[h]] or 0), h.replace("stall_", "")) for h in stall_cols), reverse=True)[:2]
    print(f"{100*s/tot:5.1f}%  #{k:5d}  {r[ix['Source']].strip()[:90]:90s} {top}")
