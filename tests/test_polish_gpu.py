"""Polisher pileup summary: CUDA kernels against the UNMODIFIED reference SummaryGenerator compiled
in oracle/_ref/pv_ref_polisher (summary_generator.cpp), on hand-built, fuzzed and synthetic regions; chunking against a
restatement of AlignmentSummarizer.chunk_images."""
import importlib.util
import os

import numpy as np
import pytest

import helpers as H
from pepper_thesis_b200 import polish, synth
from pepper_thesis_b200.read_batch import select_regions

pytestmark = pytest.mark.gpu


def _ref_mod():
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    f = [x for x in os.listdir(d) if x.startswith("pv_ref_polisher")]
    spec = importlib.util.spec_from_file_location("pv_ref_polisher", os.path.join(d, f[0]))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m


def ref_summary(b, r):
    m = _ref_mod()
    ro, rl = int(b.region_ref_off[r]), int(b.region_ref_len[r])
    return m.polisher_summary(b.read_pos, b.read_base_off, b.read_len, b.read_cigar_off, b.read_n_ops, b.read_flags, b.read_mapq,
                              b.bases, b.cigar, int(b.region_read_begin[r]), int(b.region_read_begin[r + 1]),
                              bytes(b.ref[ro:ro + rl]).decode(), int(b.region_ref_start[r]), int(b.region_ref_end[r]))


def check(b):
    s = polish.PolishSummary(b)
    for r in range(b.n_regions):
        img, gp = ref_summary(b, r)
        gi, gg = s.region(r)
        assert np.array_equal(gg.cpu().numpy(), gp), "genomic_pos region %d" % r
        got = gi.cpu().numpy()
        bad = np.argwhere(got != img)
        assert bad.size == 0, "image region %d differs first at %s: got %d want %d" % (r, bad[0].tolist(), got[tuple(bad[0])], img[tuple(bad[0])])
    return s


def test_hand_built():
    ref = "ACGT" * 10
    reads = [H.Read(0, "ACGTACGTAC", [(0, 10)]),
             H.Read(2, "GTTTACG", [(0, 2), (1, 2), (0, 3)], rev=True),             # insert TT behind position 3
             H.Read(2, "GTTTTACG", [(0, 2), (1, 3), (0, 3)]),                       # longer insert at the same anchor
             H.Read(4, "ACAC", [(0, 2), (2, 3), (0, 2)]),                            # deletion of 6..8: coverage charged to 6 only
             H.Read(4, "ACAC", [(0, 2), (2, 3), (0, 2)], rev=True),
             H.Read(5, "NNAC", [(4, 2), (0, 2)]),                                    # soft clip, then match at 5
             H.Read(30, "ACGTACGTACGTACG", [(0, 15)]),                               # runs past the region end
             H.Read(1, "CG", [(0, 2)], mapq=0)]                                      # mapq 0: skipped
    b = H.one_region(ref, reads)
    s = check(b)
    assert s.n_rows == 40 + 3


def test_deletion_only_columns_overflow_uint8():
    """Positions covered only by deletions have coverage 0 -> count * 254 does not fit uint8: the reference's
    double -> uint8_t conversion (x86) is reproduced."""
    ref = "ACGT" * 10
    reads = [H.Read(0, "ACGTAC" + "ACGT", [(0, 6), (2, 4), (0, 4)], rev=(i % 2 == 1)) for i in range(5)]
    check(H.one_region(ref, reads))


@pytest.mark.parametrize("seed", range(12))
def test_fuzz(seed):
    check(H.fuzz_region(seed, consistent=True))


@pytest.mark.parametrize("profile,cov", [("ont_r9", 30.0), ("hifi", 20.0)])
def test_synthetic_regions_and_chunks(profile, cov):
    b = synth.generate(profile, 250000, cov, seed=9)
    s = check(b)
    images, positions, ids, regs = s.chunks()
    # restatement of chunk_images (AlignmentSummarizer.py:19-56) on the oracle output
    k = 0
    for r in range(b.n_regions):
        img, gp = ref_summary(b, r)
        n = len(gp)
        start, end, cid = 0, min(n, 1000), 0
        while True:
            want_i = np.zeros((1000, 10), np.uint8); want_p = np.full((1000, 2), -1, np.int64)
            want_i[:end - start] = img[start:end]; want_p[:end - start] = gp[start:end]
            assert regs[k] == r and ids[k] == cid
            assert np.array_equal(images[k].cpu().numpy(), want_i) and np.array_equal(positions[k].cpu().numpy(), want_p)
            k += 1; cid += 1
            if end == n:
                break
            start = end - 50; end = min(n, start + 1000)
    assert k == images.shape[0] > b.n_regions * 90


def test_dropin_summary_generator():
    ref = "ACGT" * 50
    reads = [H.Read(0, ref[:150], [(0, 150)], rev=(i % 2 == 0)) for i in range(4)] + \
            [H.Read(10, ref[10:60] + "GG" + ref[60:120], [(0, 50), (1, 2), (0, 60)])]
    g = polish.SummaryGenerator(ref, "c", 0, 199)
    g.generate_summary(reads, 0, 199)
    b = H.one_region(ref, reads)
    img, gp = ref_summary(b, 0)
    assert g.image == img.tolist() and g.genomic_pos == [tuple(x) for x in gp.tolist()]
    assert (59, 1) in g.genomic_pos and (59, 2) in g.genomic_pos


def test_polisher_end_to_end():
    """reads -> polisher summary -> 1000/50 chunks -> model M-B chunk loop (predict_distributed_gpu.py:63-96): the GPU
    chain against the reference generator + the fp32 model restatement."""
    import torch
    import model_port as MP
    from pepper_thesis_b200 import models
    b = synth.generate("ont_r9", 100000, 20.0, seed=4)
    s = polish.PolishSummary(b)
    images, positions, ids, regs = s.chunks()
    sd = models.random_polisher_state_dict(1)
    m = models.PolisherTransducerGRU().load_state_dict(sd)
    acc, labels = m.predict_chunks(images, 100, 50)
    n = min(12, images.shape[0])
    racc, rlab = MP.polisher_predict_chunks(sd, images[:n].cpu(), 100, 50)
    assert (acc[:n].cpu() - racc).abs().max() < 2e-2
    lab = labels[:n].cpu().numpy()
    want = racc.numpy().argmax(-1)
    srt = np.sort(racc.numpy(), -1)
    clear = (srt[..., -1] - srt[..., -2]) > 2e-2
    assert (lab[clear] == want[clear]).all() and clear.mean() > 0.5
    assert int(positions[0, 0, 0]) == int(b.region_ref_start[0])
