"""Attribute ncu pc-sampling stall samples of one kernel to source lines (needs -lineinfo).

usage: ncu_by_line.py <cubin> <kernel-substring> <ncu --page source --csv --print-source sass dump> [kernel instance] [top]
The cubin comes from `cuobjdump -xelf all libualm.so` (or the .o); instruction -> line from the `//## File ..., line N` marks of
`nvdisasm -g`, instruction order = row order of the ncu SASS page.
"""
import collections, csv, re, subprocess, sys

cubin, kname, dump = sys.argv[1:4]
inst = int(sys.argv[4]) if len(sys.argv) > 4 else 0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
lines = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l][0]
where, cur = [], ("?", 0)
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.lstrip().startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
    elif re.match(r"^\s+/\*[0-9a-f]{4,6}\*/", l):
        where.append(cur)
rows = list(csv.reader(open(dump)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
blk = rows[starts[inst]:starts[inst + 1]]
hdr, data = blk[1], [r for r in blk[2:] if r]
col = {h: i for i, h in enumerate(hdr)}
assert len(data) == len(where), (len(data), len(where))
stalls = [h for h in hdr if h.startswith("stall_")]
agg = collections.defaultdict(collections.Counter)
for k, r in enumerate(data):
    a = agg[where[k]]
    a["samples"] += int(r[col["# Samples"]] or 0)
    a["inst"] += int(r[col["Instructions Executed"]] or 0)
    for s in stalls:
        a[s] += int(r[col[s]] or 0)
tot = sum(a["samples"] for a in agg.values()); toti = sum(a["inst"] for a in agg.values())
print(f"samples {tot}; warp instructions {toti}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    n = max(a["samples"], 1)
    best = sorted(stalls, key=lambda s: -a[s])[:3]
    print(f"{f}:{ln:<5d} samples {100*a['samples']/tot:5.1f} %  inst {100*a['inst']/max(toti,1):5.1f} %   " + "  ".join(f"{s[6:]} {100*a[s]/n:.0f}%" for s in best))
