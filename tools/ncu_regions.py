"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export of pileup_tile_kernel by code region of summary.cu:
warp instructions, threads per instruction (divergence) and stall samples per region, then the hottest lines."""
import csv, sys, os
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
hdr = rows[hi[0]]
iI = hdr.index('Instructions Executed'); iT = hdr.index('Thread Instructions Executed'); iN = hdr.index('# Samples')
sc = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'pepper-thesis_b200', 'csrc', 'summary.cu')).read().split('\n')
def find(s):
    return next((i + 1 for i, l in enumerate(src) if s in l), 10 ** 9)
marks = sorted([(1, 'helpers'), (find('// K0: per-tile work lists'), 'K0'), (find('struct TileCtx'), 'K1 ctx/record_event'),
         (find('__device__ __forceinline__ bool insert_quality_pass'), 'insert_quality/masks'),
         (find('__device__ __forceinline__ void count_mismatch'), 'count_mismatch'),
         (find('struct OpTable'), 'fetch/stage_round'), (find('__device__ __forceinline__ bool next_piece'), 'next_piece'),
         (find('__device__ __forceinline__ void push_exception'), 'push_exception'),
         (find('__device__ void scan_slices'), 'scan_slices'), (find('__device__ void accumulate_entry'), 'accumulate_entry'),
         (find('__device__ void record_entry'), 'record_entry'), (find('__device__ void record_other_snp'), 'record_other'),
         (find('__device__ void for_each_entry'), 'for_each_entry'), (find('pileup_tile_kernel(const SumParams p)'), 'kernel setup'),
         (find('// ---- phase B'), 'phase B'), (find('// ---- phase C'), 'phase C'), (find('// K2'), 'K2')])
agg, lines = {}, []
end = hi[1] - 1 if len(hi) > 1 else len(rows)
for r in rows[hi[0] + 1:end]:
    if r and r[0].isdigit():
        try: ln = int(r[0]); I = int(r[iI]); T = int(r[iT]); N = int(r[iN])
        except Exception: continue
        name = [m[1] for m in marks if m[0] <= ln][-1]
        a = agg.setdefault(name, [0, 0, 0]); a[0] += I; a[1] += T; a[2] += N
        st = {hdr[c]: int(r[c]) for c in sc if r[c].isdigit() and int(r[c]) > 0}
        lines.append((N, I, T, ln, r[1].strip()[:95], st))
tot = sum(a[0] for a in agg.values()); ts = sum(a[2] for a in agg.values())
print("total warp instructions %.1f M, samples %d" % (tot / 1e6, ts))
for k, a in agg.items():
    print("%-24s inst %7.1f M (%4.1f%%)  thr/inst %4.1f  samples %4.1f%%" % (k, a[0] / 1e6, 100 * a[0] / tot, a[1] / max(a[0], 1), 100 * a[2] / max(ts, 1)))
for N, I, T, ln, s, st in sorted(lines, key=lambda x: -x[1])[:top]:
    t3 = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print("%5.1f%%i %5.1f%%s thr %4.1f L%-4d %-95s %s" % (100 * I / tot, 100 * N / max(ts, 1), T / max(I, 1), ln, s, [(k[6:], v) for k, v in t3]))
