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
