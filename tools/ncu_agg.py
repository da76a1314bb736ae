"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name: count, total, mean, and mean by position."""
import csv, sys, collections, re
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ik]); v = float(r[iv].replace(",", "")); u = r[iu]
    v_us = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    agg.setdefault(name, []).append(v_us)
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    n = len(v); q = max(1, n // 4)
    print("%-60s n=%5d total %9.1f us (%4.1f%%) mean %7.1f  first-quarter mean %7.1f  last-quarter mean %7.1f" % (k[:60], n, sum(v), 100 * sum(v) / tot, sum(v) / n, sum(v[:q]) / q, sum(v[-q:]) / q))
