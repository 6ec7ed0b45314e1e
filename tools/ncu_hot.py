"""Summarise an `ncu --page raw` + `--page source` CSV pair: key metrics and the hottest source lines."""
import csv, sys
raw, src = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 24
rows = list(csv.reader(open(raw))); hdr = rows[0]; vals = rows[2]
keys = ('gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed_op_shared_atom.sum', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed')
for h, v in zip(hdr, vals):
    if h in keys or ('issue_stalled' in h and 'per_issue_active' in h and float(v or 0) > 0.3):
        print(h, '=', v)
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
hdr = rows[hi[0]]; iN = hdr.index('# Samples'); iI = hdr.index('Instructions Executed')
sc = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = toti = 0; agg = []
for r in rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))]:
    if r and r[0].isdigit():
        try: smp = int(r[iN]); ins = int(r[iI])
        except Exception: continue
        st = {hdr[c]: int(r[c]) for c in sc if r[c].isdigit() and int(r[c]) > 0}
        agg.append((smp, ins, int(r[0]), r[1].strip()[:100], st)); tot += smp; toti += ins
print("total samples", tot, "inst", toti)
for smp, ins, ln, s, st in sorted(agg, key=lambda x: -x[0])[:top]:
    t3 = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print("%5.1f%%s %5.1f%%i L%-4d %-100s %s" % (100 * smp / tot, 100 * ins / toti, ln, s, [(k[6:], v) for k, v in t3]))
