"""CPU, world_size 2 over gloo: region sharding + host-side gather of result records. The
per-rank compute is injected (the C oracle port) because no GPU is present; the product code under test is the
sharding / gather / merge logic of pepper_thesis_b200.pipeline."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    import pyoracle as O
    from pepper_thesis_b200 import pipeline, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    thr = synth.PROFILES["hifi"].thresholds
    n_regions = 5
    lo, hi = pipeline.shard_regions(n_regions, rank, world)
    b = synth.generate("hifi", n_regions * 100000, 6.0, seed=31, first_region=lo, num_regions=hi - lo)

    def compute(batch, offset):
        parts = []
        for r in range(batch.n_regions):
            o = O.port_summary(batch, r, thr)
            k = len(o["position"])
            al = np.zeros((k, 64), np.uint8); aln = np.zeros(k, np.uint8)
            for i, a in enumerate(o["alleles"]):
                al[i, :len(a)] = np.frombuffer(a, np.uint8); aln[i] = len(a)
            parts.append(pipeline.Predictions(np.full(k, offset + r, np.int32), o["position"].astype(np.int64), o["depth"],
                                              o["frequency"], al, aln, np.zeros((k, 3), np.float32), np.zeros(k, np.uint8)))
        return pipeline.Predictions.concat(parts)

    local = compute(b, lo)
    merged = pipeline.gather_to_rank0(local)
    if rank == 0:
        full = synth.generate("hifi", n_regions * 100000, 6.0, seed=31)
        want = compute(full, 0)
        ok = (np.array_equal(merged.region, want.region) and np.array_equal(merged.position, want.position)
              and merged.alleles() == want.alleles() and len(want) > 20)
        q.put(bool(ok))
    else:
        assert merged is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo_shard_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=180)
    for p in procs:
        p.join(60)
    assert ok and all(p.exitcode == 0 for p in procs)


def test_shard_regions_partition():
    from pepper_thesis_b200.pipeline import shard_regions
    for n in (1, 7, 640, 641):
        for w in (1, 2, 4, 8):
            blocks = [shard_regions(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [b[1] - b[0] for b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _cv_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from pepper_thesis_b200 import call_variant as CV
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # interval i belongs to rank i % world (ImageGenerationUI.py:211); sites 10000 and 30000 sit on interval boundaries and
    # are reported by both neighbours
    def cand(contig, pos, ref, alt, tag):
        return (contig, pos, pos + 1, ref, [alt], [0, 1], 30, [12], 0.9, [0.05, 0.9, 0.05], [0.9], False, tag)[:12]
    sites = {0: [("chr1", 500, "A", "C"), ("chr1", 10000, "G", "T")], 1: [("chr1", 10000, "G", "T"), ("chr1", 10000, "G", "GA"), ("chr1", 15000, "C", "A")],
             2: [("chr1", 20500, "T", "G"), ("chr1", 30000, "A", "AT")], 3: [("chr1", 30000, "A", "AT"), ("chr0", 7, "C", "G")]}
    variant, phasing, contigs = {}, {}, []
    for i in range(4):
        if i % world != rank:
            continue
        for c, pos, ref, alt in sites[i]:
            variant.setdefault((c, pos), []).append(cand(c, pos, ref, alt, i))
            phasing.setdefault((c, pos), []).append(cand(c, pos, ref, alt, i)[:10])
            if c not in contigs:
                contigs.append(c)
    res = CV.gather_candidates((contigs, phasing, variant, dict(intervals=2, candidates=len(variant))), rank, world)
    if rank == 0:
        contigs, p, v, st = res
        ok = (contigs == ["chr0", "chr1"] and list(v) == [("chr0", 7), ("chr1", 500), ("chr1", 10000), ("chr1", 15000), ("chr1", 20500), ("chr1", 30000)]
              and [(c[3], c[4][0]) for c in v[("chr1", 10000)]] == [("G", "T"), ("G", "GA")] and len(v[("chr1", 30000)]) == 1
              and list(p) == list(v) and st["intervals"] == 4)
        q.put(bool(ok))
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo_candidate_gather():
    """call_variant's gather of the selected candidates: sites sorted, one entry per (ref, alt) across the ranks."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_cv_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=180)
    for p in procs:
        p.join(60)
    assert ok and all(p.exitcode == 0 for p in procs)
