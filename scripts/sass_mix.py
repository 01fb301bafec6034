"""Static SASS instruction mix per kernel of libthetis_b200.so (developer tool)."""
import re, subprocess, sys, collections
lib = sys.argv[1] if len(sys.argv) > 1 else "thetis_b200/libthetis_b200.so"
pat = sys.argv[2] if len(sys.argv) > 2 else "swe_stage"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur = None
mix = collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        mix[cur][m.group(2)] += 1
for f, c in mix.items():
    if pat not in f:
        continue
    tot = sum(c.values())
    dp = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    print(f"{f}: total {tot}  fp64 {dp}  MUFU {c['MUFU']}  CALL {c['CALL']}  MOV {c['MOV']}  IMAD {c['IMAD']}  LDS {c['LDS']}")
    print("   ", dict(c.most_common(14)))
