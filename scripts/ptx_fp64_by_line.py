"""Static fp64 operation count per source line of one kernel, from the -lineinfo PTX (developer tool; no GPU needed).

    python scripts/ptx_fp64_by_line.py [mangled-kernel-name-fragment] [file.cu]
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "thetis_b200", "csrc")
name = sys.argv[1] if len(sys.argv) > 1 else "swe_stage_kernelILb1ELi3E"
cu = sys.argv[2] if len(sys.argv) > 2 else "tb_kernels.cu"
with tempfile.TemporaryDirectory() as tmp:
    ptx = os.path.join(tmp, "k.ptx")
    subprocess.run(["nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
                    "-I", os.path.join(ROOT, "include"), "-ptx", "-o", ptx, os.path.join(CSRC, cu)], check=True)
    s = open(ptx).read()
files = dict(re.findall(r'\.file\s+(\d+)\s+"([^"]+)"', s))
i = re.search(r"\.entry\s+\S*" + re.escape(name), s).start()
j = s.find(".entry", i + 10)
body = s[i:j if j > 0 else len(s)]
cur, cnt = None, collections.defaultdict(collections.Counter)
for line in body.splitlines():
    m = re.match(r"\s*\.loc\s+(\d+)\s+(\d+)", line)
    if m:
        cur = (os.path.basename(files[m.group(1)]), int(m.group(2)))
        continue
    m = re.match(r"\s+(?:@!?%p\d+\s+)?((?:mul|add|sub|fma\.rn)\.f64|rsqrt\.approx\.ftz\.f64|rcp\.approx\.ftz\.f64)", line)
    if m:
        cnt[cur][m.group(1).split(".")[0]] += 1
src = {f: open(os.path.join(CSRC, f)).read().splitlines() for f in ("tb_kernels.cu", "tb_tracer.cu", "tb_device.cuh")}
tot = 0
for k in sorted(cnt):
    n = sum(v for o, v in cnt[k].items() if o in ("mul", "add", "sub", "fma"))
    tot += n
    text = src[k[0]][k[1] - 1].strip()[:88] if k[0] in src else ""
    print(f"{k[0]}:{k[1]:4d} {n:4d} {dict(cnt[k])}  | {text}")
print("fp64 arithmetic (static):", tot)
