"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time and share per kernel."""
import csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = {}
for r in rows[hi + 1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", ""))
    us = v / 1e3 if r[iu] in ("ns", "nsecond") else v * 1e3 if r[iu] in ("ms", "msecond") else v
    name = re.sub(r"\(.*", "", r[ik]).strip()
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values()); n = sum(a[0] for a in agg.values())
print("# launches %d, total %.1f us" % (n, tot))
print("# kernel | launches | total us | share")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-90s %5d %10.1f %6.1f%%" % (k[:90], a[0], a[1], 100 * a[1] / tot))
