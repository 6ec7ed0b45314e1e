"""BASELINE.json configs[3] for bench.py: whole-genome-scale synthetic 3.1 Gbp at 40x, ONT R10 Q20 preset, on 8 B200.

124 G read bases do not fit any host, so every block of regions is GENERATED ON THE DEVICE (synth_device.generate, the
bit-identical twin of the seeded host generator), run through summary + inference, and dropped: nothing is materialised.
Each rank owns a contiguous 1/8 of the genome's 31 000 regions (387.5 Mbp); with fewer than 8 ranks the same per-rank
share is run and the line says so. The metric excludes input synthesis: the hot path of every block is
bracketed by CUDA events on its stream and the event times are summed; the wall time including generation is reported
beside it.
"""
from __future__ import annotations

import time

import torch

from . import pipeline, synth, synth_device

GENOME_BP = 3_100_000_000
REGION_BP = 100000
RANKS_FULL = 8
BLOCK_REGIONS = 640            # 64 Mbp per block: four resident groups of 160 regions
GROUP_REGIONS = 160


def run_config4(ctx, steps=1, share_regions=None):
    preset, coverage = "ont_r10", 40.0
    total_regions = GENOME_BP // REGION_BP                                  # 31 000
    per_rank = share_regions or (total_regions + RANKS_FULL - 1) // RANKS_FULL   # 3875
    first = ctx.rank * per_rank
    contig_len = total_regions * REGION_BP + 1000
    thr = synth.PROFILES[preset].thresholds
    hp = pipeline.HotPath(ctx.model, thr, ctx.device, group_regions=GROUP_REGIONS)
    dev = ctx.device
    torch.cuda.synchronize(dev)
    if ctx.world > 1:
        ctx.dist.barrier()
    t_wall0 = time.perf_counter()
    hot_ms, gen_ms, cands, bp, bases = 0.0, 0.0, 0, 0, 0
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    warm = True
    r0 = 0
    while r0 < per_rank:
        n = min(BLOCK_REGIONS, per_rank - r0)
        e[0].record()
        groups = [synth_device.generate(preset, contig_len, coverage, seed=1, first_region=first + r0 + g0,
                                        num_regions=min(GROUP_REGIONS, n - g0), device=dev) for g0 in range(0, n, GROUP_REGIONS)]
        if warm:                                         # first block once untimed: allocator pools, workspaces, kernels' first launch
            hp.run_device(groups, [first + r0 + g0 for g0 in range(0, n, GROUP_REGIONS)], to_host=False)
            warm = False
        e[1].record()
        out = hp.run_device(groups, [first + r0 + g0 for g0 in range(0, n, GROUP_REGIONS)], to_host=False)
        e[2].record()
        torch.cuda.synchronize(dev)
        gen_ms += e[0].elapsed_time(e[1])
        hot_ms += e[1].elapsed_time(e[2])
        cands += int(out["count"])
        bp += sum(g.candidate_bp for g in groups)
        bases += sum(g.read_bases for g in groups)
        del groups, out
        r0 += n
    wall = time.perf_counter() - t_wall0
    t = torch.tensor([hot_ms, wall * 1e3], dtype=torch.float64, device=dev)
    s = torch.tensor([float(bp), float(cands), float(bases)], dtype=torch.float64, device=dev)
    if ctx.world > 1:
        ctx.dist.all_reduce(t, op=ctx.dist.ReduceOp.MAX)
        ctx.dist.all_reduce(s, op=ctx.dist.ReduceOp.SUM)
    hot_max, wall_max = float(t[0]), float(t[1])
    bp_all = float(s[0])
    return {"workload": "whole-genome-scale synthetic 3.1 Gbp at 40x, ONT R10 Q20 preset: %d of the 31000 regions per GPU, generated on "
                        "the device block by block (64 Mbp), never materialised" % per_rank,
            "n_gpus": ctx.world, "scaling": "weak (per-GPU share of the 8-GPU run)" if ctx.world != RANKS_FULL else "the full genome over 8 GPUs",
            "genome_fraction_covered": round(bp_all / GENOME_BP, 4),
            "value": round(bp_all / (hot_max / 1e3) / 1e6, 2), "unit": "Mbp/s", "steps": 1,
            "hot_path_seconds": round(hot_max / 1e3, 3), "generation_seconds_rank0": round(gen_ms / 1e3, 3),
            "wall_seconds_including_generation": round(wall_max / 1e3, 3),
            "value_including_generation": round(bp_all / (wall_max / 1e3) / 1e6, 2),
            "read_bases_all_ranks": int(s[2]), "candidates_all_ranks": int(s[1]),
            "timing": "CUDA events around summary + inference of every block, summed; MAX over ranks; input synthesis excluded (SURVEY 8d)"}
