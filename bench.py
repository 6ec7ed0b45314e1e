#!/usr/bin/env python
"""bench.py -- Mbp/s of pileup-summary + TransducerGRU inference (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]                  our arm (one process per GPU under torchrun)
    python bench.py --impl reference [--gpus N --steps K --warmup W]     the reference's CPU path on the host cores

A step = one pass of the hot path (summary kernels -> int16 windows -> LSTM model) over one batch of synthetic
regions. The default workload is BASELINE.json configs[1] (chr20-scale 64 Mbp, 50x ONT R9 Guppy5 SUP preset); under
torchrun every rank owns its own 64 Mbp block of regions (weak scaling, no data-path collective). The same line also
carries, under "extras", the other BASELINE configs the box can run in a bounded time:
    config3   HiFi preset, 64 Mbp at 35x, the regions of ONE contig sharded over the N ranks (strong scaling)
    config4   ONT R10 Q20 preset at 40x, whole-genome scale, reads generated per region ON the device and streamed
              (3.1 Gbp over 8 GPUs = 387.5 Mbp per GPU; with fewer GPUs the same per-GPU share is run and said so)
    config5   TransducerGRU inference-only sweep (windows x 100 positions, hidden 128) next to the LSTM model
    bam_ingest  SURVEY 8f row 1: a synthetic 30x ONT BAM + FASTA written here, decoded ON the device (BGZF inflate, record
              parsing, get_reads clipping -> PvReadBatch in HBM) and taken through the summary kernels: BAM -> candidates
Flags select any of them as the main workload instead: --preset/--coverage/--mbp/--scaling.

Three numbers per workload:
    value      inputs already resident in HBM when the timed region starts
    e2e        host buffers in, host results out through HotPath.run_host, inputs = the PLAIN PvReadBatch arrays (one byte
               per base, BAM u32 CIGAR words, per-read / per-region headers, reference) in page-locked host memory -- the
               analogue of the CPU arm's pre-built type_read lists. Every upload is inside the timed region.
    e2e_wire   the same call on the compact wire forms (reference-predicted bases, 8-bit CIGAR); packing them is host
               work OUTSIDE the timed region, so its seconds are reported with it and folded into value_including_pack.
"""
from __future__ import annotations

import argparse
import hashlib
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
PRESET_NAME = {"ont_r9": "ONT R9 Guppy5 SUP", "ont_r10": "ONT R10 Q20", "hifi": "HiFi"}
REGION_BP = 100000
LSTM_FLOP_PER_WINDOW = 2 * 80664064
LSTM_DEC_STEP_FLOP_PER_WINDOW = 2 * 2 * 1024 * 768   # one decoder step launch: 2 directions x [1024 x (256+512)] MACs
SUMMARY_FAMILIES = ("sum_cigar_prefix", "sum_pileup_tile", "sum_site_alleles", "sum_key_sort", "sum_emit_windows")
LSTM_FAMILIES = ("lstm_input_prep", "lstm_encoder_steps", "lstm_decoder_steps", "lstm_mlp_head")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--preset", default=os.environ.get("PV_BENCH_PRESET", "ont_r9"), choices=sorted(PRESET_NAME))
    ap.add_argument("--coverage", type=float, default=float(os.environ.get("PV_BENCH_COVERAGE", "50")))
    ap.add_argument("--mbp", type=float, default=float(os.environ.get("PV_BENCH_MBP", "64")),
                    help="Mbp of contig per GPU per step (weak) or in total (strong); default: the chr20-scale 64")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every rank owns its own --mbp block; strong: ONE --mbp contig, regions sharded over the ranks")
    ap.add_argument("--cpu-regions", type=int, default=0, help="regions in the CPU-baseline sample (0 = 2 per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the config3 / config4 / config5 measurements")
    return ap.parse_args()


def workload_name(preset, coverage, mbp, scaling):
    if scaling == "weak":
        return "chr20-scale synthetic %g Mbp per GPU, %gx %s preset, summary+LSTM on B200" % (mbp, coverage, PRESET_NAME[preset])
    return "synthetic %g Mbp contig, %gx %s preset, regions sharded over the GPUs, summary+LSTM on B200" % (
        mbp, coverage, PRESET_NAME[preset])


def shared_config(args):
    """The `config` object both arms print (the reference arm runs a bounded sample of this workload)."""
    return {"workload": workload_name(args.preset, args.coverage, args.mbp, args.scaling), "preset": args.preset,
            "coverage": args.coverage, "mbp": args.mbp, "region_bp": REGION_BP, "scaling": args.scaling}


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
def _ref_worker(conn, preset, coverage, use_ref):
    """One host core of the reference's stage 1: owns the regions the parent assigns to it (interval i -> worker
    i % workers, ImageGenerationUI.py:211). 'prep' builds the inputs (synthetic reads -> type_read lists: NOT timed,
    BAM decoding is excluded on both sides); 'run' is the timed part: the UNMODIFIED reference C++ on every owned region."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as O
    from pepper_thesis_b200 import synth
    thr = synth.PROFILES[preset].thresholds
    work, images = [], None
    while True:
        msg = conn.recv()
        if msg[0] == "prep":
            work = []
            for region in msg[1]:
                b = synth.generate(preset, 10 ** 9, coverage, seed=1, first_region=region, num_regions=1, threads=1)
                work.append((b, O.ref_build_reads(b, 0) if use_ref else None))
            conn.send(("ready",))
        elif msg[0] == "run":
            t0 = time.perf_counter()
            outs = [O.ref_run(b, 0, thr, rs) if use_ref else O.port_summary(b, 0, thr) for b, rs in work]
            busy = time.perf_counter() - t0
            conn.send(("done", busy, sum(len(o["position"]) for o in outs), sum(b.candidate_bp for b, _ in work)))
            parts = [np.asarray(o["images"], dtype=np.int16).reshape(-1, 33, 26) for o in outs if len(o["position"])]
            images = np.concatenate(parts) if parts else np.zeros((0, 33, 26), np.int16)
        elif msg[0] == "images":
            conn.send(images)
        else:
            return


class CpuReference:
    """The reference's CPU path on all host cores: stage 1 = its own region_summary.cpp (oracle/_ref; the C port only when
    that is not built), one worker process per core; stage 2 = the TransducerGRU in eager fp32 PyTorch with all threads,
    batch 512 (predict_distributed_cpu.py:102-147; onnxruntime is absent), on the windows stage 1 produced."""

    def __init__(self, preset, coverage, cores):
        import multiprocessing as mp
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pyoracle as O
        import torch
        import model_port as MP
        self.use_ref = O.have_ref()
        if self.use_ref:
            O.ref_module()                      # loaded in the parent too: the workers are forks of this process
        self.cores = cores
        ctx = mp.get_context("fork")
        self.workers = []
        for _ in range(cores):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_ref_worker, args=(b, preset, coverage, self.use_ref), daemon=True)
            p.start()
            self.workers.append((p, a))
        torch.set_num_threads(cores)
        self.torch = torch
        self.model = MP.TorchVariantModule(MP.variant_state_dict(0)).eval()

    def step(self, n_regions, first_region=0):
        own = [[first_region + r for r in range(w, n_regions, self.cores)] for w in range(self.cores)]
        for (_, c), regs in zip(self.workers, own):
            c.send(("prep", regs))
        for _, c in self.workers:
            c.recv()
        t0 = time.perf_counter()                            # ---- timed: stage 1 on all cores (wall) ----
        for _, c in self.workers:
            c.send(("run",))
        res = [c.recv() for _, c in self.workers]
        t_summary = time.perf_counter() - t0
        k = sum(r[2] for r in res)
        bp = sum(r[3] for r in res)
        parts = []
        for _, c in self.workers:
            c.send(("images",))
            parts.append(c.recv())
        win = np.concatenate(parts) if parts else np.zeros((0, 33, 26), np.int16)
        x = self.torch.from_numpy(win.astype(np.float32))
        with self.torch.no_grad():
            t1 = time.perf_counter()                        # ---- timed: stage 2, every window of the step ----
            for i in range(0, x.shape[0], 512):
                self.model(x[i:i + 512])
            t_infer = time.perf_counter() - t1
        return dict(bp=bp, candidates=k, t_summary=t_summary, t_infer=t_infer, busy=sum(r[1] for r in res),
                    mbps=bp / (t_summary + t_infer) / 1e6, summary_mbps=bp / t_summary / 1e6,
                    infer_wps=(k / t_infer) if t_infer > 0 else 0.0, n_regions=n_regions)

    def close(self):
        for p, c in self.workers:
            try:
                c.send(("stop",))
            except Exception:
                pass
        for p, _ in self.workers:
            p.join(timeout=5)

    def describe(self, n_regions, coverage):
        return ("%d regions x 100 kbp at %gx per step, region i on worker i %% %d: stage 1 = %s, wall time over all workers; "
                "stage 2 = eager fp32 PyTorch TransducerGRU (nn.LSTM / nn.Linear re-declaration of the reference module; the "
                "reference's onnxruntime is absent), %d threads, batch 512, every window of the step" % (
                    n_regions, coverage, self.cores, "unmodified reference region_summary.cpp (oracle/_ref)" if self.use_ref
                    else "C port (oracle/region_summary_port.c)", self.cores))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_regions = args.cpu_regions or min(2 * cores, 64)
    ref = CpuReference(args.preset, args.coverage, cores)
    try:
        for w in range(max(0, args.warmup)):
            ref.step(n_regions, first_region=w * n_regions)
        t0 = time.perf_counter()
        steps = [ref.step(n_regions, first_region=(args.warmup + s) * n_regions) for s in range(max(1, args.steps))]
        wall = time.perf_counter() - t0
    finally:
        ref.close()
    timed_s = sum(s["t_summary"] + s["t_infer"] for s in steps)
    mbps = sum(s["bp"] for s in steps) / timed_s / 1e6
    kind = "reference" if ref.use_ref else "port"
    cfg = shared_config(args)
    line = {"impl": "reference", "metric": METRIC, "value": round(mbps, 4), "unit": "Mbp/s", "n_gpus": args.gpus,
            "steps": len(steps), "warmup": max(0, args.warmup), "ms_per_step": round(1e3 * timed_s / len(steps), 1),
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "fp32 model; summary int32, fp64 thresholds",
            "data": "synthetic (seeded reads/contig, random-init weights torch.manual_seed(0))",
            "config": cfg,
            "cpu_baseline": {"value": round(mbps, 4), "unit": "Mbp/s", "cores": cores, "kind": kind,
                             "sample": ref.describe(n_regions, args.coverage),
                             "regions_per_step": n_regions, "candidates_per_step": int(np.mean([s["candidates"] for s in steps])),
                             "summary_mbps": round(float(np.mean([s["summary_mbps"] for s in steps])), 3),
                             "infer_windows_per_s": round(float(np.mean([s["infer_wps"] for s in steps])), 1),
                             "wall_s_including_input_preparation": round(wall, 1)},
            "e2e": {"value": round(mbps, 4), "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---- our arm --------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def _max_over_ranks(ctx, ms):
    t = ctx.torch.tensor([ms], dtype=ctx.torch.float64, device=ctx.device)
    if ctx.world > 1:
        ctx.dist.all_reduce(t, op=ctx.dist.ReduceOp.MAX)
    return float(t.item())


def _sum_over_ranks(ctx, v):
    t = ctx.torch.tensor([float(v)], dtype=ctx.torch.float64, device=ctx.device)
    if ctx.world > 1:
        ctx.dist.all_reduce(t, op=ctx.dist.ReduceOp.SUM)
    return float(t.item())


def _barrier(ctx):
    ctx.torch.cuda.synchronize()
    if ctx.world > 1:
        ctx.dist.barrier()
    ctx.torch.cuda.synchronize()


def _timed(ctx, fn, steps):
    """`steps` calls of fn between two CUDA events on the current stream, barrier + synchronize on both sides."""
    torch = ctx.torch
    _barrier(ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = None
    for _ in range(steps):
        out = fn()
    e1.record()
    _barrier(ctx)
    return e0.elapsed_time(e1), out


def measure_workload(ctx, preset, coverage, mbp, scaling, steps, warmup, main):
    """One workload on this rank's share of the regions. Returns a dict of measurements (all ranks compute it)."""
    from pepper_thesis_b200 import capi, device as dev, pipeline, synth
    torch, lib = ctx.torch, ctx.lib
    total_regions = max(1, int(round(mbp * 1e6 / REGION_BP)))
    if scaling == "weak":
        n_regions, first = total_regions, ctx.rank * total_regions
        contig_len = total_regions * ctx.world * REGION_BP + 1000
    else:
        lo, hi = pipeline.shard_regions(total_regions, ctx.rank, ctx.world)     # contiguous blocks of ONE contig
        n_regions, first = hi - lo, lo
        contig_len = total_regions * REGION_BP + 1000
    t0 = time.time()
    if n_regions <= 0:
        raise ValueError("fewer regions than ranks")
    batch = synth.generate(preset, contig_len, coverage, seed=1, first_region=first, num_regions=n_regions,
                           threads=ctx.gen_threads)
    # the batch's smallest base quality is metadata of the batch (PvReadBatch.min_qual, tracked by the ingest while it
    # decodes; here one scan). When it clears both quality thresholds no kernel reads a quality: the quality array is
    # neither loaded by the tile kernel nor uploaded by the host path.
    batch.scan_min_qual(ctx.gen_threads)
    thr = synth.PROFILES[preset].thresholds
    skip_q = dev.quals_not_needed(batch.min_qual, thr) and os.environ.get("PV_BENCH_UPLOAD_QUALS", "0") != "1"
    batch.pin_plain(with_quals=not skip_q)
    gen_s = time.time() - t0
    bp = batch.candidate_bp
    bp_all = _sum_over_ranks(ctx, bp)
    hp = pipeline.HotPath(ctx.model, thr, ctx.device, group_regions=int(os.environ.get("PV_BENCH_HOST_GROUP", "128")),
                          skip_quals=skip_q, infer_batch=int(os.environ.get("PV_BENCH_INFER_BATCH", "32768")))
    out = {"preset": preset, "coverage": coverage, "mbp": mbp, "scaling": scaling, "regions_rank0": n_regions,
           "reads_rank0": batch.n_reads, "read_bases_rank0": int(batch.read_len.astype(np.int64).sum()),
           "min_qual": int(batch.min_qual), "synth_seconds": round(gen_s, 1), "bp_all_ranks": int(bp_all)}

    # ---- device-resident ("value"): inputs in HBM before the timed region. Groups are larger than the host-path groups:
    # there is no upload to overlap, and the small kernels of the chain are launch-latency bound.
    res_group = int(os.environ.get("PV_BENCH_RESIDENT_GROUP", "160"))
    groups = [(r0, min(n_regions, r0 + res_group)) for r0 in range(0, n_regions, res_group)]
    resident = [dev.DeviceBatch(batch.region_range_view(*g), ctx.device, non_blocking=False) for g in groups]
    torch.cuda.synchronize()
    out["resident_bytes_rank0"] = int(sum(d.h2d_bytes for d in resident))

    def step_resident(dbs=resident):
        return hp.run_device(dbs, [g[0] for g in groups], to_host=False)["count"] if dbs else 0

    for _ in range(max(warmup, 3)):
        k_step = step_resident()
    lib.pv_profile_reset()
    lib.pv_profile_enable(1)
    clocks = ClockSampler(ctx.local) if main else None
    if clocks:
        clocks.start()
    launches0 = lib.pv_launch_count()
    ms, k_step = _timed(ctx, step_resident, steps)
    out["gpu_launches"] = int(lib.pv_launch_count() - launches0)
    lib.pv_profile_enable(0)
    out["prof"] = capi.profile_collect()
    if clocks:
        out["clocks"] = clocks.stop()
    ms_max = _max_over_ranks(ctx, ms)
    out["ms_per_step"] = ms_max / steps
    out["value"] = bp_all * steps / (ms_max / 1e3) / 1e6
    out["candidates_rank0"] = int(k_step)
    out["batch"] = batch

    # ---- the same pass WITHOUT the min_qual promise: every quality is loaded and tested (the general path of the tile
    # kernel), so the number does not depend on the synthetic qualities clearing the thresholds
    if main and batch.min_qual > 0 and not skip_q_forced_off():
        plain = [d.with_min_qual(0) for d in resident]
        for _ in range(2):
            step_resident(plain)
        lib.pv_profile_reset()
        lib.pv_profile_enable(1)
        n_g = max(2, min(steps, 5))
        ms_g, _ = _timed(ctx, lambda: step_resident(plain), n_g)
        lib.pv_profile_enable(0)
        prof_g = capi.profile_collect()
        ms_g = _max_over_ranks(ctx, ms_g)
        tile_ms = prof_g.get("sum_pileup_tile", (0.0, 0))[0] / n_g
        alg = batch.algorithmic_bytes(int(k_step))
        out["general_path"] = {"what": "device-resident pass with min_qual withheld: the tile kernel loads and tests every quality",
                               "value": round(bp_all * n_g / (ms_g / 1e3) / 1e6, 2), "unit": "Mbp/s", "steps": n_g,
                               "ms_per_step": round(ms_g / n_g, 3), "pileup_tile_ms_per_step": round(tile_ms, 3),
                               "pileup_tile_achieved_gbs": round(alg / max(tile_ms, 1e-9) * 1e3 / 1e9, 1)}
        del plain
    del resident
    torch.cuda.empty_cache()

    # ---- end to end through the public API, PLAIN host arrays in page-locked memory ("e2e") ----------------------------
    # Two ways of moving the same plain arrays, both entirely inside the timed region: uploaded as they are (PCIe-bound), or
    # squeezed group by group into 2-bit bases / 16-bit CIGAR by host threads while the group before is on the wire
    # (HotPath(pack_inline=True), csrc/host_pack.cpp) and expanded again on the device. The faster one is the line's `e2e`
    # (which one wins depends on the host cores per GPU); both are reported.
    pred_box = {}

    # measured on a 16-core box (tools/probe_pack_inline.py): groups of 32-48 regions, two uploads in flight, CIGAR words
    # left plain (packing them costs more host memory bandwidth than the wire time it saves) -> 45-46 ms per 64 Mbp
    hp_pack = pipeline.HotPath(ctx.model, thr, ctx.device, group_regions=int(os.environ.get("PV_BENCH_PACK_GROUP", "48")),
                               skip_quals=skip_q, infer_batch=hp.infer_batch, pack_inline=True, host_ahead=2)

    def measure_host(pack_inline):
        h = hp_pack if pack_inline else hp

        def step_host():
            pred_box["p"] = h.run_host(batch, first)
            return pred_box["p"]
        for _ in range(max(warmup, 3)):      # the caching allocators (device + pinned staging) settle after two passes
            h.run_host(batch, first)
        ms_h, pred_h = _timed(ctx, step_host, steps)
        return _max_over_ranks(ctx, ms_h), pred_h, int(h.last_h2d_bytes)

    quals_txt = ("NOT uploaded: min_qual %d clears both thresholds (%g, %g), so no kernel reads one (PV_BENCH_UPLOAD_QUALS=1 "
                 "uploads them)" % (batch.min_qual, thr.min_snp_baseq, thr.min_indel_baseq) if skip_q else "u8")
    ms_e, pred, h2d_e = measure_host(False)
    e2e_groups = hp.group_regions
    modes = {"plain_upload": {"value": round(bp_all * steps / (ms_e / 1e3) / 1e6, 2), "ms_per_step": round(ms_e / steps, 2),
                              "h2d_bytes_per_step": h2d_e}}
    fmt = "plain PvReadBatch arrays in page-locked memory: bases u8, CIGAR u32 (BAM words), per-read / per-region headers, " \
          "reference; qualities %s" % quals_txt
    if os.environ.get("PV_BENCH_PACK_INLINE", "1") == "1":
        ms_p, pred_p, h2d_p = measure_host(True)
        same_p = (len(pred_p) == len(pred) and np.array_equal(pred_p.position, pred.position)
                  and np.array_equal(pred_p.allele, pred.allele) and np.array_equal(pred_p.genotype, pred.genotype)
                  and np.array_equal(pred_p.probs, pred.probs))
        modes["packed_inline"] = {"value": round(bp_all * steps / (ms_p / 1e3) / 1e6, 2), "ms_per_step": round(ms_p / steps, 2),
                                  "h2d_bytes_per_step": h2d_p, "pack_threads": hp_pack.pack_threads,
                                  "groups_of_regions": hp_pack.group_regions,
                                  "identical_results_to_plain_upload": bool(same_p)}
        if ms_p < ms_e and same_p:
            ms_e, pred, h2d_e = ms_p, pred_p, h2d_p
            e2e_groups = hp_pack.group_regions
            fmt = ("the same plain PvReadBatch arrays in page-locked memory (bases u8, CIGAR u32, headers, reference); INSIDE the "
                   "timed region %d host threads squeeze each group's bases into 2 bits + an exception list (pv_pack_group) in "
                   "pinned staging while the group before is on the wire, the device expands them (pv_unpack_bases2); "
                   "qualities %s" % (hp_pack.pack_threads, quals_txt))
    del hp_pack
    d2h = sum(getattr(pred, f).nbytes for f in ("region", "position", "depth", "frequency", "allele", "allele_len", "probs", "genotype"))
    out["e2e"] = {"value": round(bp_all * steps / (ms_e / 1e3) / 1e6, 2), "unit": "Mbp/s",
                  "h2d_bytes_per_step": int(h2d_e), "d2h_bytes_per_step": int(d2h),
                  "ms_per_step": round(ms_e / steps, 2), "host_format": fmt, "groups_of_regions": e2e_groups,
                  "modes": modes}

    # ---- the same call on the compact wire forms; packing is host work outside the timed region and is reported -----------
    if main and os.environ.get("PV_BENCH_WIRE", "1") == "1" and n_regions > 0:
        import copy
        bw = copy.copy(batch)
        t0 = time.time()
        bw.pack_wire(threads=ctx.gen_threads, pinned=True)
        pack_s = time.time() - t0
        hp_w = pipeline.HotPath(ctx.model, thr, ctx.device, group_regions=int(os.environ.get("PV_BENCH_WIRE_GROUP", "160")),
                                taper=False, skip_quals=skip_q)
        for _ in range(max(warmup, 3)):
            pw = hp_w.run_host(bw, first)
        same = (len(pw) == len(pred) and np.array_equal(pw.position, pred.position) and np.array_equal(pw.allele, pred.allele)
                and np.array_equal(pw.genotype, pred.genotype))
        ms_w, _ = _timed(ctx, lambda: hp_w.run_host(bw, first), steps)
        ms_w = _max_over_ranks(ctx, ms_w)
        pack_max = _max_over_ranks(ctx, pack_s)
        out["e2e_wire"] = {"value": round(bp_all * steps / (ms_w / 1e3) / 1e6, 2), "unit": "Mbp/s",
                           "h2d_bytes_per_step": int(hp_w.last_h2d_bytes), "ms_per_step": round(ms_w / steps, 2),
                           "pack_seconds_untimed": round(pack_max, 2), "pack_threads": ctx.gen_threads,
                           "value_including_pack": round(bp_all / (ms_w / steps / 1e3 + pack_max) / 1e6, 2),
                           "same_candidates_and_genotypes_as_e2e": bool(same), "groups_of_regions": hp_w.group_regions,
                           "host_format": "bases %s, CIGAR %s, qualities %s (packed once by pv_pack_* on the host)" % (
                               "reference-predicted + patch list" if bw.bases_patch is not None else (
                                   "2-bit + exception list" if bw.bases2 is not None else "4-bit (BAM nt16)"),
                               "8-bit codes + escapes" if bw.cigar8 is not None else ("u16" if bw.cigar16 is not None else "u32"),
                               "not uploaded (min_qual promise)" if skip_q else "%d-bit packed" % bw.qual_bits)}
        del bw, hp_w
    torch.cuda.empty_cache()
    return out


def skip_q_forced_off():
    return os.environ.get("PV_NO_ALLQ", "0") == "1"


def rooflines(res, steps, peaks):
    """roofline of the dominant kernel + the summary chain's and the model's own, from the CUDA events of the timed region."""
    batch, prof, k = res["batch"], res["prof"], res["candidates_rank0"]
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tc_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    fam_ms = {f: v[0] for f, v in prof.items()}
    dominant = max(fam_ms, key=fam_ms.get) if fam_ms else None
    roof = None
    allq = res["min_qual"] > 0 and not skip_q_forced_off() and res.get("allq", True)
    if dominant:
        ms_d, n_d = prof[dominant]
        per_launch_ms = ms_d / max(1, n_d)
        if dominant.startswith("lstm") or dominant.startswith("gru"):
            per_win = {"lstm_decoder_steps": LSTM_DEC_STEP_FLOP_PER_WINDOW, "lstm_encoder_steps": 2 * 2 * 1024 * (256 + 26),
                       "lstm_mlp_head": 2 * (16896 * 512 + 4 * 512 * 512 + 512 * 3) / 6.0}.get(dominant, 0)
            flops_total = per_win * k * steps * (33 if "steps" in dominant else 6)
            achieved = flops_total / (ms_d / 1e3) / 1e12
            roof = {"kernel": dominant, "bound": "tensor", "achieved": round(achieved, 2), "peak": tc_peak, "unit": "TFLOP/s",
                    "frac": round(achieved / tc_peak, 4), "traffic": None, "avg_launch_ms": round(per_launch_ms, 4),
                    "launches": n_d, "peak_source": peak_src + ", bf16 sustained"}
        else:
            alg = batch.algorithmic_bytes(k) * steps
            achieved = alg / (ms_d / 1e3) / 1e9
            roof = {"kernel": dominant, "bound": "hbm", "achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s",
                    "frac": round(achieved / hbm_peak, 4), "traffic": None, "avg_launch_ms": round(per_launch_ms, 4),
                    "launches": n_d, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg / max(1, n_d))}
    alg = batch.algorithmic_bytes(k) * steps
    tile_ms, tile_n = prof.get("sum_pileup_tile", (0.0, 0))
    sum_ms = sum(fam_ms.get(f, 0.0) for f in SUMMARY_FAMILIES)
    roof_summary = {"kernel": "sum_pileup_tile", "bound": "hbm", "achieved": round(alg / max(tile_ms, 1e-9) * 1e3 / 1e9, 1),
                    "peak": hbm_peak, "unit": "GB/s", "frac": round(alg / max(tile_ms, 1e-9) * 1e3 / 1e9 / hbm_peak, 4),
                    "algorithmic_bytes_per_step": int(alg / steps), "avg_launch_ms": round(tile_ms / max(1, tile_n), 4),
                    "launches": tile_n, "chain_ms_per_step": round(sum_ms / steps, 3), "traffic": None, "peak_source": peak_src,
                    "quality_loads": ("skipped: the batch's min_qual (%d) clears both thresholds, so 1 B/base of the algorithmic "
                                      "bytes is never read" % res["min_qual"]) if allq else "1 B/base"}
    if allq:
        real_bases = res["read_bases_rank0"]
        alg_q = (batch.algorithmic_bytes(k) - real_bases) * steps
        roof_summary["achieved_without_quality_bytes"] = round(alg_q / max(tile_ms, 1e-9) * 1e3 / 1e9, 1)
        roof_summary["frac_without_quality_bytes"] = round(alg_q / max(tile_ms, 1e-9) * 1e3 / 1e9 / hbm_peak, 4)
    # DRAM bytes of one launch of the tile kernel from the committed `ncu --set full` capture of this command, scaled to
    # this run's regions per launch -- only while the capture is of the CURRENT kernel source (else null: stale)
    try:
        with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
            tr = json.load(f)
        with open(os.path.join(ROOT, "pepper-thesis_b200", "csrc", "summary.cu"), "rb") as f:
            sha = hashlib.sha256(f.read()).hexdigest()[:16]
        if tr.get("summary_cu_sha16") == sha and tile_n:
            regions_per_launch = res["regions_rank0"] * steps / tile_n
            traffic = int(tr["dram_bytes_per_launch"] * regions_per_launch / tr["regions_per_launch"])
            roof_summary["traffic"] = traffic
            roof_summary["traffic_source"] = "profiles/k1_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum per launch, %s)" % tr.get("capture", "")
            if roof and roof["kernel"] == "sum_pileup_tile":
                roof["traffic"] = traffic
                roof["traffic_source"] = roof_summary["traffic_source"]
    except Exception:
        pass
    inf_ms = sum(fam_ms.get(f, 0.0) for f in LSTM_FAMILIES)
    inf_tf = 161.33e6 * k * steps / max(inf_ms, 1e-9) * 1e3 / 1e12
    roof_inference = {"kernels": " + ".join(LSTM_FAMILIES), "bound": "tensor",
                      "achieved": round(inf_tf, 1), "peak": tc_peak, "unit": "TFLOP/s", "frac": round(inf_tf / tc_peak, 4),
                      "windows_per_step": int(k), "ms_per_step": round(inf_ms / steps, 3),
                      "windows_per_s": round(k * steps / max(inf_ms, 1e-9) * 1e3), "peak_source": peak_src + ", bf16 sustained"}
    return roof, roof_summary, roof_inference, {f: round(v / steps, 3) for f, v in fam_ms.items() if v > 0}


def extras_config5(ctx):
    """BASELINE config 5 in brief: polisher TransducerGRU (biGRU x2, hidden 128, [B,100,10]) and the variant LSTM model,
    inference only, device-resident, rank 0. The full sweep against torch/cuDNN is tools/sweep_models.py."""
    from pepper_thesis_b200 import models
    torch = ctx.torch
    rows = []
    gru = models.PolisherTransducerGRU(1, 10, 1, 128, 5, True)
    gru.load_state_dict(models.random_polisher_state_dict(0))
    for n in (256, 1024, 4096, 16384):
        row = {"windows": n}
        x = torch.randint(0, 31, (n, 100, 10), dtype=torch.uint8, device=ctx.device)
        h = torch.zeros((n, 2, 128), dtype=torch.float32, device=ctx.device)
        for _ in range(3):
            gru.forward(x, h)
        ms, _ = _timed_local(torch, lambda: gru.forward(x, h), 10)
        row["gru_ms"] = round(ms / 10, 3)
        row["gru_tflops"] = round(80.44e6 * n / (ms / 10 / 1e3) / 1e12, 1)
        w = torch.randint(-30, 31, (n, 33, 26), dtype=torch.int16, device=ctx.device)
        for _ in range(3):
            ctx.model.infer_windows(w, wrap_int8=False)
        ms, _ = _timed_local(torch, lambda: ctx.model.infer_windows(w, wrap_int8=False), 10)
        row["lstm_ms"] = round(ms / 10, 3)
        row["lstm_tflops"] = round(161.33e6 * n / (ms / 10 / 1e3) / 1e12, 1)
        rows.append(row)
    return {"what": "TransducerGRU inference only, device-resident inputs, bf16 tensor-core GEMMs with fp32 state (config 5); "
                    "gru = polisher biGRU x2 hidden 128 on [B,100,10], lstm = variant biLSTM x2 + MLP on [B,33,26]",
            "rows": rows}


def _timed_local(torch, fn, iters):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = None
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from pepper_thesis_b200 import capi, models
    from pepper_thesis_b200 import nativebuild as build

    ctx = Ctx()
    ctx.torch, ctx.dist = torch, dist
    ctx.world = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.rank = int(os.environ.get("RANK", "0"))
    ctx.local = int(os.environ.get("LOCAL_RANK", "0"))
    if ctx.world > 1:
        # rank 0 prints ONE JSON line on stdout: NCCL's version banner appears at NCCL_DEBUG=VERSION / WARN / INFO, so an
        # inherited setting is dropped for this process (NCCL_DEBUG_FILE keeps it available to whoever wants the log)
        if not os.environ.get("NCCL_DEBUG_FILE"):
            os.environ.pop("NCCL_DEBUG", None)
        dist.init_process_group("nccl", device_id=torch.device("cuda", ctx.local))
    torch.cuda.set_device(ctx.local)
    ctx.device = torch.device("cuda", ctx.local)
    if ctx.rank == 0:
        build.build_all()
    if ctx.world > 1:
        dist.barrier()
    ctx.lib = capi.load()                                # raises when the CUDA library is missing: no fallback
    ctx.gen_threads = max(1, min(64, (os.cpu_count() or 1) // max(1, ctx.world)))        # ranks share the host cores
    ctx.model = models.TransducerGRU(26, 1, 256, 28, 3, True)
    ctx.model.load_state_dict(models.random_variant_state_dict(0))

    res = measure_workload(ctx, args.preset, args.coverage, args.mbp, args.scaling, args.steps, args.warmup, main=True)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    roof, roof_summary, roof_inference, kernel_ms = rooflines(res, args.steps, peaks)
    del res["batch"]

    extras = {}
    if not args.no_extras:
        ex_steps, ex_warm = max(2, min(args.steps, 5)), 3
        try:
            if not (args.preset == "hifi" and args.scaling == "strong"):
                r3 = measure_workload(ctx, "hifi", 35.0, 64.0, "strong", ex_steps, ex_warm, main=False)
                _, rs3, ri3, km3 = rooflines(r3, ex_steps, peaks)
                extras["config3"] = {"workload": workload_name("hifi", 35.0, 64.0, "strong"), "n_gpus": ctx.world, "scaling": "strong",
                                     "value": round(r3["value"], 2), "unit": "Mbp/s", "steps": ex_steps, "ms_per_step": round(r3["ms_per_step"], 3),
                                     "e2e": r3["e2e"], "regions_rank0": r3["regions_rank0"], "candidates_rank0": r3["candidates_rank0"],
                                     "pileup_tile_frac_of_hbm_peak": rs3["frac"], "inference_frac_of_tensor_peak": ri3["frac"],
                                     "kernel_ms_per_step": km3}
                del r3
        except Exception as e:      # an extra must never take the main line down
            extras["config3"] = {"error": repr(e)[:300]}
        try:
            from pepper_thesis_b200 import stream_bench
            extras["config4"] = stream_bench.run_config4(ctx, ex_steps)
        except ImportError:
            extras["config4"] = {"unavailable": "device-side read generator not built"}
        except Exception as e:
            extras["config4"] = {"error": repr(e)[:300]}
        if ctx.world > 1:
            dist.barrier()
        if ctx.rank == 0:
            try:
                extras["config5"] = extras_config5(ctx)
            except Exception as e:
                extras["config5"] = {"error": repr(e)[:300]}
            # BAM + FASTA -> candidates with the BAM decoded on the device (a synthetic 30x ONT BAM written here)
            try:
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import bench_ingest_gpu
                extras["bam_ingest"] = bench_ingest_gpu.run(float(os.environ.get("PV_BENCH_INGEST_MBP", "16")), 30.0)
            except Exception as e:
                extras["bam_ingest"] = {"error": repr(e)[:300]}

    if ctx.rank != 0:
        if ctx.world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- CPU baseline beside it (rank 0, N=1 only) -----------------------------------------------------------------------
    cpu = None
    if ctx.world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_cpu = args.cpu_regions or min(2 * cores, 64)
        ref = CpuReference(args.preset, args.coverage, cores)
        try:
            ref.step(min(n_cpu, cores))
            c = ref.step(n_cpu, first_region=cores)
        finally:
            ref.close()
        cpu = {"value": round(c["mbps"], 4), "unit": "Mbp/s", "cores": cores, "kind": "reference" if ref.use_ref else "port",
               "sample": ref.describe(n_cpu, args.coverage) + "; one warm-up step, one timed step",
               "summary_mbps": round(c["summary_mbps"], 3), "infer_windows_per_s": round(c["infer_wps"], 1)}

    cfg = shared_config(args)
    line = {"metric": METRIC, "value": round(res["value"], 2), "unit": "Mbp/s", "n_gpus": ctx.world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(res["ms_per_step"], 3), "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None,
            "dtype": "bf16 (tensor-core GEMMs, fp32 accumulate/state); summary int32/u8, fp64 thresholds",
            "data": "synthetic (seeded reads/contig, random-init weights torch.manual_seed(0))",
            "config": cfg,
            "detail": {"regions_per_gpu": res["regions_rank0"], "reads": res["reads_rank0"], "read_bases": res["read_bases_rank0"],
                       "candidates_per_step_rank0": res["candidates_rank0"], "min_qual": res["min_qual"],
                       "l2": "inputs (%.2f GB resident) larger than L2 (126 MB), no flush needed" % (res["resident_bytes_rank0"] / 1e9),
                       "resident_groups_of_regions": int(os.environ.get("PV_BENCH_RESIDENT_GROUP", "160")),
                       "synth_seconds": res["synth_seconds"]},
            "e2e": res["e2e"], "e2e_wire": res.get("e2e_wire"), "general_path": res.get("general_path"),
            "gpu_launches": res["gpu_launches"], "clocks": res.get("clocks"), "roofline": roof, "roofline_summary": roof_summary,
            "roofline_inference": roof_inference, "kernel_ms_per_step": kernel_ms, "extras": extras, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if ctx.world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)
