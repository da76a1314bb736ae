"""SASS opcode census of libualm.so per kernel: the mnemonics that prove what the hardware path is (B200_PROFILING.md: TMA = UTMALDG / UBLKCP,
cp.async = LDGSTS, mbarrier = SYNCS, fp64 = DFMA / DADD / DMUL, local memory = LDL / STL).  python tools/sass_census.py > profiles/sass_census_r02.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "uneven_planner_b200", "libualm.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WATCH = ["UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "LDG", "STG", "LDS", "STS", "ATOMS", "RED", "ATOMG", "SHFL", "BAR", "DFMA", "DADD", "DMUL", "FFMA", "FADD", "FMUL", "MUFU", "LDL", "STL",
         "HMMA", "UTCHMMA"]
cur, agg = None, collections.OrderedDict()
for l in out.split("\n"):
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        agg[cur] = collections.Counter()
        continue
    m = re.match(r"^\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
    if m and cur:
        agg[cur][m.group(2).split(".")[0]] += 1
        agg[cur]["_n"] += 1
print("# SASS census of %s (sm_100a), instructions per kernel; watched opcodes only when present" % os.path.relpath(so, ROOT))
for k, a in agg.items():
    print("%-70s %7d  " % (k[:70], a["_n"]) + " ".join("%s=%d" % (o, a[o]) for o in WATCH if a[o]))
