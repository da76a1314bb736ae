"""Attribute ncu pc-sampling stall samples of one kernel to its (noinline) device functions.

usage: ncu_by_function.py <cubin> <kernel-substring> <ncu --page source --csv --print-source sass dump> [kernel instance]
The cubin comes from `cuobjdump -xelf all libualm.so`; function boundaries from the `$kernel$function:` labels of nvdisasm.
"""
import collections, csv, re, subprocess, sys

cubin, kname, dump = sys.argv[1:4]
inst = int(sys.argv[4]) if len(sys.argv) > 4 else 0
lines = subprocess.run(["nvdisasm", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l][0]
funcs, idx = [(0, kname + " (kernel body)")], 0
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.lstrip().startswith(".section"):
        break
    if l.startswith("$") and l.endswith(":"):
        funcs.append((idx, l[:-1].split("$")[-1]))
    elif re.match(r"^\s+/\*[0-9a-f]{4,6}\*/", l):
        idx += 1

def dem(n):
    out = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    return re.sub(r"\(.*", "", out).replace("ualm::", "")

rows = list(csv.reader(open(dump)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
blk = rows[starts[inst]:starts[inst + 1]]
hdr, data = blk[1], [r for r in blk[2:] if r]
col = {h: i for i, h in enumerate(hdr)}
assert len(data) == idx, (len(data), idx)
stalls = ["stall_barrier", "stall_long_sb", "stall_wait", "stall_short_sb", "stall_branch_resolving", "stall_no_inst", "stall_selected"]
agg = collections.defaultdict(collections.Counter)
bounds = [f[0] for f in funcs] + [10 ** 9]
fi = 0
for k, r in enumerate(data):
    while k >= bounds[fi + 1]:
        fi += 1
    a = agg[funcs[fi][1]]
    a["samples"] += int(r[col["# Samples"]] or 0)
    a["inst"] += int(r[col["Instructions Executed"]] or 0)
    for s in stalls:
        a[s] += int(r[col[s]] or 0)
tot = sum(a["samples"] for a in agg.values()); totb = sum(a["stall_barrier"] for a in agg.values()); toti = sum(a["inst"] for a in agg.values())
print(f"samples {tot}, of which barrier (idle helper warps) {totb}; instructions {toti}")
print("| device function | samples excl. barrier | instructions | long_sb | wait | short_sb | branch | no_inst | selected |")
print("|---|---|---|---|---|---|---|---|---|")
for nm, a in sorted(agg.items(), key=lambda kv: -(kv[1]["samples"] - kv[1]["stall_barrier"]))[:28]:
    nb = max(a["samples"] - a["stall_barrier"], 1)
    f = lambda s: "%.0f %%" % (100 * a[s] / nb)
    print(f"| `{dem(nm)[:60]}` | {100*nb/(tot-totb):.1f} % | {100*a['inst']/toti:.1f} % | {f('stall_long_sb')} | {f('stall_wait')} | {f('stall_short_sb')} | "
          f"{f('stall_branch_resolving')} | {f('stall_no_inst')} | {f('stall_selected')} |")
