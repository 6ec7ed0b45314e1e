"""BAM -> packed batch on the device against the host ingest, and BAM -> candidates (decode + summary kernels), on a synthetic
ONT BAM written here. One JSON line. Usage: python tools/bench_ingest_gpu.py [Mbp] [coverage]"""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import numpy as np, torch
import bamio, fast_bam
from pepper_thesis_b200 import capi, synth, ingest, ingest_gpu, device as dev

def run(mbp=8.0, cov=30.0, group_mbp=None, crc=True, host_arm=True):
    """Dict of the row's measurements (see the module docstring)."""
    L = int(mbp * 1e6)
    t0 = time.perf_counter()
    b = synth.generate("ont_r9", L, cov, seed=5, region_size=L, margin=0)
    d = tempfile.mkdtemp()
    bam, fa = os.path.join(d, "t.bam"), os.path.join(d, "t.fa")
    info = fast_bam.write_bam_from_batch(bam, b, "chrS", L)
    bamio.write_fasta(fa, [("chrS", bytes(b.ref[:L]).decode())])
    t_make = time.perf_counter() - t0
    bh, fh = ingest.BAMHandler(bam), ingest.FASTAHandler(fa)
    starts = list(range(0, L, 100000)); ends = [min(L - 1, s + 100000) for s in starts]
    thr = synth.PROFILES["ont_r9"].thresholds
    out = dict(row="ingest_gpu", mbp=mbp, coverage=cov, bam_MB=round(info["file_bytes"] / 1e6, 1), inflated_MB=round(info["inflated_bytes"] / 1e6, 1),
               records=info["records"], make_seconds=round(t_make, 1), verify_crc=crc)
    # host ingest, all threads
    t0 = time.perf_counter(); host = ingest.ingest_regions(bh, fh, "chrS", starts, ends, min_mapq=1, threads=os.cpu_count() or 1); dt = time.perf_counter() - t0
    out["host_ingest"] = dict(value=round(L / dt / 1e6, 2), unit="Mbp/s", threads=os.cpu_count(), seconds=round(dt, 3))
    # device ingest
    torch.cuda.synchronize()
    best = None
    for rep in range(4):
        capi.load().pv_profile_reset(); capi.load().pv_profile_enable(1)
        t0 = time.perf_counter(); got = ingest_gpu.ingest_regions_gpu(bh, fh, "chrS", starts, ends, min_mapq=1, verify_crc=crc); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        prof = capi.profile_collect(); capi.load().pv_profile_enable(0)
        if rep and (best is None or dt < best[0]):
            best = (dt, dict(got.stats), prof["bam_decode"])
    dt, stats, (kms, kl) = best
    out["device_ingest"] = dict(value=round(L / dt / 1e6, 2), unit="Mbp/s", seconds=round(dt, 4), host_plan_s=round(stats["host_plan_s"], 4),
                                device_s=round(stats["device_s"], 4), kernels_ms=round(kms, 3), kernel_launches=kl,
                                compressed_GBps=round(stats["compressed_bytes"] / dt / 1e9, 2), inflated_GBps=round(stats["inflated_bytes"] / dt / 1e9, 2),
                                trace_ms=stats.get("trace_ms"), blocks=stats["bgzf_blocks"], segments=stats["chain_segments"], reads=stats["reads"])
    # equality with the host ingest (every large array)
    gb = got.batch.to_host()
    same = all(np.array_equal(getattr(gb, n), getattr(host.batch, n)) for n in ("read_pos", "read_len", "read_base_off", "read_cigar_off", "bases", "quals", "cigar", "region_read_begin", "ref"))
    out["identical_to_host_ingest"] = bool(same)
    # BAM -> candidates: decode + summary chain
    ws = dev.SummaryWorkspace.for_batch(got.batch, max(8192, int(L / 1000 * 8)))
    def once():
        g = ingest_gpu.ingest_regions_gpu(bh, fh, "chrS", starts, ends, min_mapq=1, verify_crc=crc)
        dev.summary_regions(g.batch, thr, ws)
        return int(ws.count.item())
    once(); torch.cuda.synchronize()
    t0 = time.perf_counter(); k = once(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out["bam_to_candidates"] = dict(value=round(L / dt / 1e6, 2), unit="Mbp/s", seconds=round(dt, 4), candidates=k, status=ws.status())
    # streamed: the contig in groups of regions, host share of group i+1 under the device work of group i (decode + summary kernels)
    if group_mbp is None:
        group_mbp = float(os.environ.get("PV_INGEST_GROUP_MBP", "32"))
    per = max(1, int(group_mbp * 1e6 / 100000))
    groups = [(starts[i:i + per], ends[i:i + per]) for i in range(0, len(starts), per)]
    from pepper_thesis_b200 import models, pipeline
    hot = pipeline.HotPath(None, thr, "cuda", group_regions=per)  # its summary half only: workspaces pooled across groups
    trace = []
    def streamed():
        total = 0
        del trace[:]
        t_last = time.perf_counter()
        for g in ingest_gpu.stream_regions_gpu(bh, fh, "chrS", groups, min_mapq=1, verify_crc=crc):
            t_in = time.perf_counter()
            w, kk = hot.summarize(g.batch)
            total += kk
            t_sum = time.perf_counter()
            trace.append((round((t_in - t_last) * 1e3, 1), round((t_sum - t_in) * 1e3, 1), round(g.stats["host_plan_s"] * 1e3, 1), round(g.stats["device_s"] * 1e3, 1)))
            t_last = t_sum
        return total
    streamed(); streamed(); torch.cuda.synchronize()                # twice: the pool of page-locked buffers fills on the way
    t0 = time.perf_counter(); k2 = streamed(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out["bam_to_candidates_streamed"] = dict(value=round(L / dt / 1e6, 2), unit="Mbp/s", seconds=round(dt, 4), groups=len(groups), group_mbp=group_mbp, candidates=k2,
                                             same_candidate_count=bool(k2 == k), per_group_ms_ingest_summary_hoststage_devicestage=list(trace))
    import shutil
    shutil.rmtree(d, ignore_errors=True)
    return out


if __name__ == "__main__":
    print(json.dumps(run(float(sys.argv[1]) if len(sys.argv) > 1 else 8.0, float(sys.argv[2]) if len(sys.argv) > 2 else 30.0,
                         crc=os.environ.get("PV_INGEST_CRC", "1") == "1")))
