"""Throughput of the widened rows: ingest (host C++), candidate filter and polisher summary
(device). One JSON line per row; the CPU reference beside each where one can run here."""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np, torch
import bamio
from pepper_thesis_b200 import capi, synth, ingest, polish, candidate_filter as CF, device as dev
from pepper_thesis_b200.pipeline import Predictions

out = []
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm = float(peaks.get("hbm_gbs", 6650.0))

# ---- row 1: BAM + FASTA -> packed batch ------------------------------------------------------------------------------
L = 4_000_000
b = synth.generate("ont_r9", L, 30.0, seed=5, region_size=L, margin=0)
recs = []
for i in range(b.n_reads):
    bo, n = int(b.read_base_off[i]), int(b.read_len[i]); co, k = int(b.read_cigar_off[i]), int(b.read_n_ops[i])
    recs.append(dict(tid=0, pos=int(b.read_pos[i]), mapq=60, flag=0x10 if b.read_flags[i] & 1 else 0, name="r%d" % i,
                     cigar=[(int(c) & 15, int(c) >> 4) for c in b.cigar[co:co + k]], seq=bytes(b.bases[bo:bo + n]).decode(),
                     qual=bytes(b.quals[bo:bo + n]), tags=b""))
recs.sort(key=lambda r: r["pos"])
d = tempfile.mkdtemp()
bam, fa = os.path.join(d, "t.bam"), os.path.join(d, "t.fa")
bamio.write_bam(bam, [("chrS", L)], recs)
bamio.write_fasta(fa, [("chrS", bytes(b.ref[:L]).decode())])
bh, fh = ingest.BAMHandler(bam), ingest.FASTAHandler(fa)
starts = list(range(0, L, 100000)); ends = [min(L - 1, s + 100000) for s in starts]
for threads in (1, os.cpu_count() or 1):
    t0 = time.perf_counter(); got = ingest.ingest_regions(bh, fh, "chrS", starts, ends, min_mapq=1, threads=threads); dt = time.perf_counter() - t0
    out.append(dict(row="ingest", metric="Mbp/s of 30x ONT BAM -> packed batch (get_reads clipping + reference fetch)", value=round(L / dt / 1e6, 2),
                    threads=threads, bam_MB=round(os.path.getsize(bam) / 1e6, 1), bam_MBps=round(os.path.getsize(bam) / dt / 1e6, 1),
                    reads=int(got.batch.n_reads), bases_per_s=round(float(got.batch.read_len.sum()) / dt / 1e6, 1), unit="Mbp/s"))

# ---- row 4: polisher summary -----------------------------------------------------------------------------------------
pb = synth.generate("ont_r9", 8_000_000, 50.0, seed=6)
db = dev.DeviceBatch(pb); torch.cuda.synchronize()
for _ in range(2): s = polish.PolishSummary(db)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); s = polish.PolishSummary(db); im, pos, ids, regs = s.chunks(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
alg = 1 * int(pb.read_len.astype(np.int64).sum()) + 4 * pb.n_ops + 32 * pb.n_reads + s.n_rows * (10 + 16) + im.numel()
out.append(dict(row="polisher_summary", metric="Mbp/s polisher pileup summary + 1000/50 chunking (50x ONT, device-resident reads)",
                value=round(pb.candidate_bp / ms / 1e3, 1), unit="Mbp/s", ms=round(ms, 3), rows=int(s.n_rows), chunks=int(im.shape[0]),
                roofline=dict(bound="hbm", achieved=round(alg / ms / 1e6, 1), peak=hbm, unit="GB/s", frac=round(alg / ms / 1e6 / hbm, 4),
                              algorithmic_bytes="1 B/base + 4 B/op + 32 B/read + 26 B/row + chunk images")))
# CPU reference beside it (unmodified summary_generator.cpp, one region, one core)
try:
    import importlib.util
    rd = os.path.join(ROOT, "oracle", "_ref"); f = [x for x in os.listdir(rd) if x.startswith("pv_ref_polisher")][0]
    spec = importlib.util.spec_from_file_location("pv_ref_polisher", os.path.join(rd, f)); m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    r = 3; ro, rl = int(pb.region_ref_off[r]), int(pb.region_ref_len[r])
    t0 = time.perf_counter()
    m.polisher_summary(pb.read_pos, pb.read_base_off, pb.read_len, pb.read_cigar_off, pb.read_n_ops, pb.read_flags, pb.read_mapq, pb.bases, pb.cigar,
                       int(pb.region_read_begin[r]), int(pb.region_read_begin[r + 1]), bytes(pb.ref[ro:ro + rl]).decode(),
                       int(pb.region_ref_start[r]), int(pb.region_ref_end[r]))
    dt = time.perf_counter() - t0
    out[-1]["cpu_baseline"] = dict(value=round(100000 / dt / 1e6, 3), unit="Mbp/s", cores=1, kind="reference", sample="1 region (100 kbp, 50x), unmodified summary_generator.cpp")
except Exception as ex:
    out[-1]["cpu_baseline"] = dict(unavailable=str(ex))

# ---- row 3: candidate filter -----------------------------------------------------------------------------------------
K = 2_000_000
rng = np.random.RandomState(1)
region = np.sort(rng.randint(0, pb.n_regions, K)).astype(np.int32)
position = (pb.region_cand_start[region] + rng.randint(0, 100000, K)).astype(np.int64)
allele = np.zeros((K, 64), np.uint8); allele[:, 0] = ord("1"); allele[:, 1] = np.frombuffer(b"ACGT", np.uint8)[rng.randint(0, 4, K)]
probs = rng.rand(K, 3).astype(np.float32); probs /= probs.sum(1, keepdims=True)
t = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (position, region, rng.randint(3, 126, K).astype(np.int32), rng.randint(1, 60, K).astype(np.int32),
                                                                allele, np.full(K, 2, np.uint8), probs, pb.region_ref_start, pb.region_ref_off, pb.region_ref_len)]
ref_d = db.t["ref"]; flags = torch.zeros(K, dtype=torch.uint8, device="cuda")
import ctypes as C
o = CF.FilterOptions().as_struct(); lib = capi.load()
def run():
    capi.check(lib.pv_candidate_filter(K, *[x.data_ptr() for x in t], None, ref_d.data_ptr(), C.byref(o), flags.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
run(); torch.cuda.synchronize(); e0.record(); run(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
byts = K * (8 + 4 + 4 + 4 + 64 + 1 + 12 + 1 + 20)
out.append(dict(row="candidate_filter", metric="M candidates/s stage-3 decision (device-resident predictions)", value=round(K / ms / 1e3, 1), unit="Mcand/s",
                ms=round(ms, 3), roofline=dict(bound="hbm", achieved=round(byts / ms / 1e6, 1), peak=hbm, unit="GB/s", frac=round(byts / ms / 1e6 / hbm, 4),
                                               algorithmic_bytes="118 B/candidate (record + allele + probs + 20 reference bytes + flag)")))
for r in out:
    print(json.dumps(r))
