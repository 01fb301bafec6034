"""Executed instructions and stall samples per CUDA source line of one kernel launch in an .ncu-rep
(source page, needs -lineinfo):   python scripts/ncu_by_line.py rep [n_lines] [launch-index]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# the export is a sequence of blocks: ["File Path", path], ["Function Name", f], header row, then per source line a row with
# a line number followed by its SASS rows (empty line number)
cur_file = None
hdr = None
agg = collections.defaultdict(lambda: [0, 0, ""])      # (file, line) -> [inst, samples, text]
fp64 = collections.Counter()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Kernel Name" or r[0] == "ID":
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        # two columns are called "Source": first = cuda source text, second = SASS
        idx_inst = r.index("Instructions Executed")
        idx_samp = r.index("# Samples")
        continue
    if hdr is None or len(r) <= idx_inst:
        continue
    if r[0] != "":
        key = (cur_file, int(r[0]))
        agg[key][2] = r[1].strip()
        cur_key = key
    else:
        try:
            agg[cur_key][0] += int(r[idx_inst] or 0)
            agg[cur_key][1] += int(r[idx_samp] or 0)
            op = r[3].split()[0] if r[3].split() else ""
            if op.startswith("@"):
                op = r[3].split()[1]
            fp64[op.split(".")[0]] += int(r[idx_inst] or 0)
        except (ValueError, IndexError):
            pass
tot_i = sum(v[0] for v in agg.values()) or 1
tot_s = sum(v[1] for v in agg.values()) or 1
print(f"total warp-instructions {tot_i}  samples {tot_s}")
print("opcode mix (executed):", {k: round(100 * v / tot_i, 1) for k, v in fp64.most_common(16)})
for (f, ln), (i, s, txt) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    print(f"{100*i/tot_i:5.2f}% inst {100*s/tot_s:5.2f}% stall  {f}:{ln:<5d} {txt[:110]}")
