"""Static SASS instruction mix per kernel of libthetis_b200.so (developer tool)."""
import re, subprocess, sys, collections
lib = sys.argv[1] if len(sys.argv) > 1 else "thetis_b200/libthetis_b200.so"
pat = sys.argv[2] if len(sys.argv) > 2 else "swe_stage"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur = None
mix = collections.defaultdict(collections.Counter)
special = collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)(\S*)", line)
    if m and cur:
        mix[cur][m.group(2)] += 1
        if ".SYS" in m.group(3):
            special[cur]["sys-scope ld/st/fence (cross-GPU flags)"] += 1
for f, c in mix.items():
    if pat not in f:
        continue
    tot = sum(c.values())
    dp = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    print(f"{f}: total {tot}  fp64 {dp}  MUFU {c['MUFU']}  CALL {c['CALL']}  MOV {c['MOV']}  IMAD {c['IMAD']}  LDS {c['LDS']}")
    print("   ", dict(c.most_common(14)))

# TMA / async-copy / mbarrier / system-scope instructions of every kernel of the library
NAMES = {"LDGSTS": "LDGSTS (cp.async)", "SYNCS": "SYNCS (mbarrier)", "UBLKCP": "UBLKCP (TMA bulk copy)",
         "UBLKPF": "UBLKPF (bulk L2 prefetch)"}
print(f"\n# TMA / async-copy / mbarrier / system-scope instructions per kernel (cuobjdump -sass {lib})")
for f in sorted(mix):
    c = mix[f]
    parts = [f"DFMA={c['DFMA']}"] if c["DFMA"] else []
    parts += [f"{NAMES[k]}={c[k]}" for k in sorted(NAMES) if c[k]]
    parts += [f"{k}={v}" for k, v in special[f].items()]
    if len(parts) > (1 if c["DFMA"] else 0) or c["DFMA"] > 50:
        print(f"{f}: " + ", ".join(parts))
