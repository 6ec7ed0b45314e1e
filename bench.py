#!/usr/bin/env python
"""bench.py -- Mbp/s of pileup-summary + TransducerGRU inference (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one process per GPU under torchrun)
    python bench.py --impl reference [--gpus N --steps K --warmup W]   the reference's CPU path on the host cores

A step = one pass of the hot path (summary kernels -> int16 windows -> LSTM model) over one batch of synthetic
regions of the chr20-scale workload (BASELINE.json configs[1]: 64 Mbp, 50x ONT R9 Guppy5 SUP preset). Under torchrun
every rank owns its own 64 Mbp block of regions (weak scaling, no data-path collective).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Mbp/s pileup-summary + GRU inference"
# BASELINE.json configs[1] by default; PV_BENCH_PRESET / PV_BENCH_COVERAGE select the other presets' shapes for extra
# measurements (configs[2]: hifi at 35x, configs[3]: ont_r10 at 40x) -- the driver's bench line is always the default.
PRESET = os.environ.get("PV_BENCH_PRESET", "ont_r9")
COVERAGE = float(os.environ.get("PV_BENCH_COVERAGE", "50"))
PRESET_NAME = {"ont_r9": "ONT R9 Guppy5 SUP", "ont_r10": "ONT R10 Q20", "hifi": "HiFi"}[PRESET]
REGION_BP = 100000
LSTM_FLOP_PER_WINDOW = 2 * 80664064
LSTM_DEC_STEP_FLOP_PER_WINDOW = 2 * 2 * 1024 * 768   # one decoder step launch: 2 directions x [1024 x (256+512)] MACs


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mbp", type=float, default=float(os.environ.get("PV_BENCH_MBP", "64")),
                    help="Mbp of contig per GPU per step (default: the chr20-scale 64)")
    ap.add_argument("--cpu-regions", type=int, default=0, help="regions in the CPU-baseline sample (0 = 2 per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ---- clocks ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- CPU reference arm ----------------------------------------------------------------------------------------------
def _cpu_region_worker(args):
    """One region through the UNMODIFIED reference C++ (oracle/_ref) or, if it is not built, the C port."""
    seed, region, use_ref = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as O
    from pepper_thesis_b200 import synth
    b = synth.generate(PRESET, 10 ** 9, COVERAGE, seed=seed, first_region=region, num_regions=1, threads=1)
    thr = synth.PROFILES[PRESET].thresholds
    if use_ref:
        rs = O.ref_build_reads(b, 0)                  # type_read construction is not timed (BAM decode excluded on both sides)
        t0 = time.perf_counter()
        out = O.ref_run(b, 0, thr, rs)
        dt = time.perf_counter() - t0
    else:
        t0 = time.perf_counter()
        out = O.port_summary(b, 0, thr)
        dt = time.perf_counter() - t0
    return dt, len(out["position"]), b.candidate_bp


def cpu_reference_step(n_regions, cores, first_region=0, pool=None):
    """Summary on `n_regions` regions with one worker per core (ImageGenerationUI.py:326-328) + the LSTM model in eager
    fp32 PyTorch with all threads, batch 512 (predict_distributed_cpu.py:102-147). Returns a dict."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as O
    import torch
    import model_port as MP
    use_ref = O.have_ref()
    jobs = [(1, first_region + r, use_ref) for r in range(n_regions)]
    t0 = time.perf_counter()
    res = list(pool.map(_cpu_region_worker, jobs))
    wall_sum = time.perf_counter() - t0                 # includes input synthesis in the workers ...
    busy = sum(r[0] for r in res)                       # ... so the summary time is taken from the workers' own timers
    k = sum(r[1] for r in res)
    bp = sum(r[2] for r in res)
    t_summary = busy / min(cores, n_regions)            # perfect packing of the measured per-region times on the cores
    torch.set_num_threads(cores)
    model = MP.TorchVariantModule(MP.variant_state_dict(0)).eval()
    n_win = min(max(k, 512), 1024)
    x = -torch.randint(0, 50, (n_win, 33, 26)).float()
    with torch.no_grad():
        model(x[:512])
        t1 = time.perf_counter()
        for i in range(0, n_win, 512):
            model(x[i:i + 512])
        t_inf_sample = time.perf_counter() - t1
    t_infer = t_inf_sample * k / n_win
    return dict(bp=bp, candidates=k, t_summary=t_summary, t_infer=t_infer, wall=wall_sum, use_ref=use_ref,
                mbps=bp / (t_summary + t_infer) / 1e6, summary_mbps=bp / t_summary / 1e6,
                infer_wps=n_win / t_inf_sample, n_regions=n_regions)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ProcessPoolExecutor
    cores = os.cpu_count() or 1
    n_regions = args.cpu_regions or min(2 * cores, 64)
    with ProcessPoolExecutor(max_workers=cores) as pool:
        for _ in range(max(0, min(args.warmup, 1))):
            cpu_reference_step(min(n_regions, cores), cores, pool=pool)
        t0 = time.perf_counter()
        steps = [cpu_reference_step(n_regions, cores, first_region=s * n_regions, pool=pool) for s in range(max(1, args.steps))]
        wall = time.perf_counter() - t0
    mbps = float(np.mean([s["mbps"] for s in steps]))
    kind = "reference" if steps[0]["use_ref"] else "port"
    sample = ("%d regions x 100 kbp at %gx per step: summary = %s, one worker per core; model = eager fp32 PyTorch "
              "re-declaration of the reference TransducerGRU (nn.LSTM/nn.Linear), %d threads, batch 512, extrapolated "
              "from a <=1024-window sample to the %d candidates found" % (
                  n_regions, COVERAGE, "unmodified reference region_summary.cpp (oracle/_ref)" if kind == "reference"
                  else "C port (oracle/region_summary_port.c)", cores, steps[0]["candidates"]))
    line = {"impl": "reference", "metric": METRIC, "value": round(mbps, 4), "unit": "Mbp/s", "n_gpus": args.gpus,
            "steps": len(steps), "warmup": min(args.warmup, 1), "ms_per_step": round(1e3 * wall / len(steps), 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": "chr20-scale synthetic 64 Mbp, %gx %s preset (bounded sample of it)" % (COVERAGE, PRESET_NAME),
                       "regions_per_step": n_regions},
            "cpu_baseline": {"value": round(mbps, 4), "unit": "Mbp/s", "cores": cores, "kind": kind, "sample": sample,
                             "summary_mbps": round(float(np.mean([s["summary_mbps"] for s in steps])), 3),
                             "infer_windows_per_s": round(float(np.mean([s["infer_wps"] for s in steps])), 1)},
            "e2e": {"value": round(mbps, 4), "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---- our arm --------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from pepper_thesis_b200 import capi, device as dev, models, pipeline, synth
    from pepper_thesis_b200 import nativebuild as build

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")      # keeps NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if rank == 0:
        build.build_all()
    if world > 1:
        dist.barrier()
    lib = capi.load()                                   # raises when the CUDA library is missing: no fallback

    # ---- workload: this rank's block of regions of the synthetic contig ---------------------------------------------
    n_regions = max(1, int(round(args.mbp * 1e6 / REGION_BP)))
    contig_len = n_regions * world * REGION_BP + 1000
    t0 = time.time()
    gen_threads = max(1, min(64, (os.cpu_count() or 1) // max(1, world)))        # ranks share the host cores
    batch = synth.generate(PRESET, contig_len, COVERAGE, seed=1, first_region=rank * n_regions, num_regions=n_regions,
                           threads=gen_threads)
    # host buffers in the compact wire forms: 2-bit bases (+ exceptions), bit-packed qualities, 16-bit CIGAR; only what
    # is uploaded lives in page-locked memory (3.2 GB per rank instead of 10 GB)
    # the batch's smallest base quality travels as metadata (PvReadBatch.min_qual, like qual_bits): when it clears both
    # quality thresholds the tile kernel never loads a quality (PV_NO_ALLQ=1 switches that off)
    batch.scan_min_qual(gen_threads)
    batch.pack_wire(threads=gen_threads, pinned=True).pin_uploaded()
    gen_s = time.time() - t0
    thr = synth.PROFILES[PRESET].thresholds
    bp = batch.candidate_bp

    model = models.TransducerGRU(26, 1, 256, 28, 3, True)
    model.load_state_dict(models.random_variant_state_dict(0))
    hp = pipeline.HotPath(model, thr, device, group_regions=int(os.environ.get("PV_BENCH_HOST_GROUP", "128")))

    # resident copy for the kernel-only number (inputs in HBM before the timed region starts). Device-resident groups
    # are larger than the host-path groups: there is no upload to overlap, and K0/K2/sort/K3 are launch-latency bound.
    res_group = int(os.environ.get("PV_BENCH_RESIDENT_GROUP", "160"))
    groups = [(r0, min(n_regions, r0 + res_group)) for r0 in range(0, n_regions, res_group)]
    resident = [dev.DeviceBatch(batch.region_range_view(*g), device, non_blocking=False) for g in groups]
    torch.cuda.synchronize()
    input_bytes = sum(d.h2d_bytes for d in resident)

    def step_resident():
        return hp.run_device(resident, [g[0] for g in groups], to_host=False)["count"]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        k_step = step_resident()
    # ---- timed: device-resident ("value") ------------------------------------------------------------------------------
    lib.pv_profile_reset()
    lib.pv_profile_enable(1)
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = lib.pv_launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        k_step = step_resident()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.pv_launch_count() - launches0
    lib.pv_profile_enable(0)
    prof = capi.profile_collect()
    clk = clocks.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = bp * world * args.steps / (ms_max / 1e3) / 1e6

    # ---- timed: end to end through the public API with host buffers ("e2e") ----------------------------------------------
    for _ in range(max(args.warmup, 3)):     # the caching allocators (device + pinned staging) settle after two passes
        hp.run_host(batch, rank * n_regions)
    barrier()
    w0 = time.perf_counter()
    e0.record()
    d2h = 0
    for _ in range(args.steps):
        pred = hp.run_host(batch, rank * n_regions)
        d2h = sum(getattr(pred, f).nbytes for f in ("region", "position", "depth", "frequency", "allele", "allele_len", "probs", "genotype"))
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    t = torch.tensor([ms_e2e], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = bp * world * args.steps / (float(t.item()) / 1e3) / 1e6

    # ---- the same end-to-end call with the qualities travelling as PREDICATES (pv_pack_quals_pred: identical summaries
    # and candidates for these thresholds, the qualities themselves stay on the host). Reported beside `e2e`, which
    # stays on the lossless wire forms. ----------------------------------------------------------------------------------
    e2e_qp = None
    if os.environ.get("PV_BENCH_QUALS_PRED", "1") == "1":
        import copy
        bq = copy.copy(batch)
        bq.pack_quals_pred(thr.min_snp_baseq, thr.min_indel_baseq, threads=gen_threads, pinned=True)
        if bq.quals_patch is not None:
            qp_bytes = input_bytes + bq.quals_patch.nbytes + bq.read_qpatch_off.nbytes - (
                min((batch.quals.size * batch.qual_bits + 7) // 8, batch.quals_packed.nbytes) if batch.quals_packed is not None else batch.quals.nbytes)
            bq.quals_packed, bq.qual_bits = None, 0
            # 0.5 GB per step left on the wire: the kernels now set the pace, so the bases go back to the 2-bit form
            # (its expansion kernel is 6x cheaper than the reference prediction, the extra 0.6 GB hides behind the
            # kernels) and the groups stay large to the end
            # ... as long as the ranks do not saturate the host's aggregate H2D bandwidth (~178 GB/s on this box: beyond 4 GPUs
            # the bytes per step decide again, and the reference-predicted bases stay)
            if os.environ.get("PV_BENCH_QP_BASES2", "1" if world <= 4 else "0") == "1" and bq.bases_patch is not None:
                bq.pack_bases2(gen_threads, pinned=True)
                if bq.bases2 is not None:
                    qp_bytes += bq.bases2.nbytes + bq.base_exceptions.nbytes - bq.bases_patch.nbytes - bq.read_patch_off.nbytes
                    bq.bases_patch, bq.read_patch_off = None, None
            hp_q = pipeline.HotPath(model, thr, device, group_regions=int(os.environ.get("PV_BENCH_QP_GROUP", "160")), taper=False)
            hp_lossless, hp = hp, hp_q
            for _ in range(max(args.warmup, 3)):
                pred_q = hp.run_host(bq, rank * n_regions)
            same = (len(pred_q) == len(pred) and np.array_equal(pred_q.position, pred.position)
                    and np.array_equal(pred_q.allele, pred.allele) and np.array_equal(pred_q.genotype, pred.genotype))
            barrier()
            e0.record()
            for _ in range(args.steps):
                hp.run_host(bq, rank * n_regions)
            e1.record()
            barrier()
            tq = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(tq, op=dist.ReduceOp.MAX)
            e2e_qp = {"value": round(bp * world * args.steps / (float(tq.item()) / 1e3) / 1e6, 2), "unit": "Mbp/s",
                      "h2d_bytes_per_step": int(qp_bytes), "ms_per_step": round(float(tq.item()) / args.steps, 2),
                      "same_candidates_and_genotypes_as_e2e": bool(same),
                      "groups_of_regions": hp.group_regions, "bases": "2-bit + exception list" if bq.bases2 is not None else "reference-predicted",
                      "wire": "qualities as threshold predicates: fill byte %d + %d patch entries (min_snp_baseq %g, min_indel_baseq %g); "
                              "summaries bit-identical, qualities not recoverable" % (bq.quals_fill, bq.quals_patch.size, thr.min_snp_baseq, thr.min_indel_baseq)}
            hp = hp_lossless
        del bq

    # ---- roofline of the dominant kernel (CUDA events recorded inside the timed region, per family) -----------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tc_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    fam_ms = {f: v[0] for f, v in prof.items()}
    dominant = max(fam_ms, key=fam_ms.get) if fam_ms else None
    k_per_step = k_step
    roof = None
    if dominant:
        ms_d, n_d = prof[dominant]
        per_launch_ms = ms_d / max(1, n_d)
        if dominant.startswith("lstm") or dominant.startswith("gru"):
            per_win = {"lstm_decoder_steps": LSTM_DEC_STEP_FLOP_PER_WINDOW, "lstm_encoder_steps": 2 * 2 * 1024 * (256 + 26),
                       "lstm_mlp_head": 2 * (16896 * 512 + 4 * 512 * 512 + 512 * 3) / 6.0}.get(dominant, 0)
            # launches of this family per step = 33 (or 6) per chunk of <= 8192 windows; windows per launch = K / chunks
            flops_total = per_win * k_per_step * args.steps * (33 if "steps" in dominant else 6)
            achieved = flops_total / (ms_d / 1e3) / 1e12
            roof = {"kernel": dominant, "bound": "tensor", "achieved": round(achieved, 2), "peak": tc_peak, "unit": "TFLOP/s",
                    "frac": round(achieved / tc_peak, 4), "traffic": None, "avg_launch_ms": round(per_launch_ms, 4),
                    "launches": n_d, "peak_source": peak_src + ", bf16 sustained"}
        else:
            alg = batch.algorithmic_bytes(k_per_step) * args.steps
            achieved = alg / (ms_d / 1e3) / 1e9
            roof = {"kernel": dominant, "bound": "hbm", "achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s",
                    "frac": round(achieved / hbm_peak, 4), "traffic": None, "avg_launch_ms": round(per_launch_ms, 4),
                    "launches": n_d, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg / max(1, n_d))}
            # DRAM bytes of one launch of this kernel from the committed `ncu --set full` capture of this command, scaled to
            # this run's regions per launch: v8 = with the min_qual promise (no quality loads), v7 = without it
            allq = (batch.min_qual > 0 and batch.min_qual >= thr.min_snp_baseq and batch.min_qual >= thr.min_indel_baseq
                    and os.environ.get("PV_NO_ALLQ", "0") != "1")
            tfile = "r1_k1_v8_bench_traffic.json" if allq else "r1_k1_v7_bench_traffic.json"
            roof["quality_loads"] = ("skipped: the batch's min_qual (%d) clears both thresholds, so 1 B/base of the algorithmic "
                                     "bytes is never read" % batch.min_qual) if allq else "1 B/base"
            if allq:
                # the same rate on the bytes this path actually has to read (SURVEY 8d's formula minus 1 B per read base)
                real_bases = int(batch.read_len.astype(np.int64).sum())
                alg_q = (batch.algorithmic_bytes(k_per_step) - real_bases) * args.steps
                roof["achieved_without_quality_bytes"] = round(alg_q / (ms_d / 1e3) / 1e9, 1)
                roof["frac_without_quality_bytes"] = round(alg_q / (ms_d / 1e3) / 1e9 / hbm_peak, 4)
            try:
                with open(os.path.join(ROOT, "profiles", tfile)) as f:
                    tr = json.load(f)
                if tr.get("kernel") == dominant:
                    regions_per_launch = n_regions * args.steps / max(1, n_d)
                    roof["traffic"] = int(tr["dram_bytes_per_launch"] * regions_per_launch / tr["regions_per_launch"])
                    roof["traffic_source"] = "profiles/%s (dram__bytes_read.sum + dram__bytes_write.sum, bytes per launch)" % tfile
            except Exception:
                pass
    # the summary chain's own HBM roofline is always reported next to it
    sum_ms = sum(fam_ms.get(f, 0.0) for f in ("sum_cigar_prefix", "sum_pileup_tile", "sum_site_alleles", "sum_key_sort", "sum_emit_windows"))
    alg = batch.algorithmic_bytes(k_per_step) * args.steps
    tile_ms = fam_ms.get("sum_pileup_tile", 0.0)
    roof_summary = {"kernel": "sum_pileup_tile", "bound": "hbm", "achieved": round(alg / max(tile_ms, 1e-9) * 1e3 / 1e9, 1),
                    "peak": hbm_peak, "unit": "GB/s", "frac": round(alg / max(tile_ms, 1e-9) * 1e3 / 1e9 / hbm_peak, 4),
                    "algorithmic_bytes_per_step": int(alg / args.steps), "chain_ms_per_step": round(sum_ms / args.steps, 3)}

    # ... and the model's tensor-core roofline (161.33 MFLOP per window, elementwise work excluded)
    inf_ms = sum(fam_ms.get(f, 0.0) for f in ("lstm_input_prep", "lstm_encoder_steps", "lstm_decoder_steps", "lstm_mlp_head"))
    inf_tf = 161.33e6 * k_per_step * args.steps / max(inf_ms, 1e-9) * 1e3 / 1e12
    roof_inference = {"kernels": "lstm_input_prep + lstm_encoder_steps + lstm_decoder_steps + lstm_mlp_head", "bound": "tensor",
                      "achieved": round(inf_tf, 1), "peak": tc_peak, "unit": "TFLOP/s", "frac": round(inf_tf / tc_peak, 4),
                      "windows_per_step": int(k_per_step), "ms_per_step": round(inf_ms / args.steps, 3),
                      "windows_per_s": round(k_per_step * args.steps / max(inf_ms, 1e-9) * 1e3), "peak_source": peak_src + ", bf16 sustained"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline beside it (rank 0, N=1 only) -----------------------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from concurrent.futures import ProcessPoolExecutor
        cores = os.cpu_count() or 1
        n_cpu = args.cpu_regions or min(2 * cores, 64)
        with ProcessPoolExecutor(max_workers=cores) as pool:
            c = cpu_reference_step(n_cpu, cores, pool=pool)
        cpu = {"value": round(c["mbps"], 4), "unit": "Mbp/s", "cores": cores, "kind": "reference" if c["use_ref"] else "port",
               "sample": "%d of the %d regions (100 kbp, %gx): unmodified reference C++ summary on all cores + eager fp32 "
                         "PyTorch LSTM model (torch-module re-declaration of the reference TransducerGRU), batch 512" % (
                             n_cpu, n_regions, COVERAGE),
               "summary_mbps": round(c["summary_mbps"], 3), "infer_windows_per_s": round(c["infer_wps"], 1)}

    line = {"metric": METRIC, "value": round(value, 2), "unit": "Mbp/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_max / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16 (tensor-core GEMMs, fp32 accumulate/state); summary int32/u8, fp64 thresholds",
            "data": "synthetic (seeded reads/contig, random-init weights torch.manual_seed(0))",
            "config": {"workload": "chr20-scale synthetic %g Mbp per GPU, %gx %s preset, summary+LSTM on B200" % (args.mbp, COVERAGE, PRESET_NAME),
                       "regions_per_gpu": n_regions, "region_bp": REGION_BP, "reads": batch.n_reads,
                       "read_bases": int(batch.read_len.astype(np.int64).sum()), "candidates_per_step_rank0": int(k_per_step), "min_qual": int(batch.min_qual),
                       "l2": "inputs (%.2f GB) larger than L2 (126 MB), no flush needed" % (input_bytes / 1e9),
                       "groups_of_regions": hp.group_regions, "resident_groups_of_regions": res_group, "synth_seconds": round(gen_s, 1),
                       "host_format": "packed SoA batch, bases %s, qualities %s, CIGAR %s (lossless, expanded on the device)" % (
                           "reference-predicted + patch list" if batch.bases_patch is not None else (
                               "2-bit + exception list" if batch.bases2 is not None else ("4-bit (BAM nt16)" if batch.bases4 is not None else "u8")),
                           "%d-bit packed" % batch.qual_bits if batch.quals_packed is not None else "u8",
                           "8-bit codes + escapes" if batch.cigar8 is not None else ("u16" if batch.cigar16 is not None else "u32 (BAM)"))},
            "e2e": {"value": round(e2e_value, 2), "unit": "Mbp/s", "h2d_bytes_per_step": int(input_bytes),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": round(ms_e2e / args.steps, 2)},
            "e2e_quals_pred": e2e_qp,
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "roofline_summary": roof_summary, "roofline_inference": roof_inference,
            "kernel_ms_per_step": {f: round(v / args.steps, 3) for f, v in fam_ms.items() if v > 0},
            "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)
