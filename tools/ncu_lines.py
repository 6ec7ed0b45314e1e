"""Rank the CUDA-C lines of an `ncu --page source --csv --print-source cuda,sass` export by executed instructions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
hdr = rows[hi[0]]
iI = hdr.index('Instructions Executed'); iT = hdr.index('Thread Instructions Executed'); iN = hdr.index('# Samples')
agg = []
for r in rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))]:
    if r and r[0].isdigit():
        try: agg.append((int(r[iI]), int(r[iT]), int(r[iN]), int(r[0]), r[1].strip()[:100]))
        except Exception: pass
tot = sum(a[0] for a in agg); tt = sum(a[1] for a in agg); ts = sum(a[2] for a in agg)
print("warp instructions %d, thread instructions %d, avg active threads %.1f" % (tot, tt, tt / tot))
for a in sorted(agg, key=lambda x: -x[0])[:top]:
    print("%5.1f%%i %5.1f%%s thr %4.1f L%-4d %s" % (100 * a[0] / tot, 100 * a[2] / ts, a[1] / max(a[0], 1), a[3], a[4]))
