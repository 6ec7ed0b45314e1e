"""The variant module's legacy SummaryGenerator / ImageSummary (legacy_summary.py over csrc/polish_summary.cu) against the
UNMODIFIED reference file compiled as oracle/_ref/pv_ref_legacy (pepper_variant/modules/cpp/summary_generator.cpp):
image, genomic_pos, ref_image, longest_insert_count and the chunks of chunk_image, on hand-built, fuzzed and synthetic
regions -- reads with mapping quality 0 included (this generator counts them, the polisher's does not)."""
import importlib.util
import os

import numpy as np
import pytest

import helpers as H
from pepper_thesis_b200 import legacy_summary, synth

pytestmark = pytest.mark.gpu


def _ref_mod():
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    f = [x for x in os.listdir(d) if x.startswith("pv_ref_legacy")]
    spec = importlib.util.spec_from_file_location("pv_ref_legacy", os.path.join(d, f[0]))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m


class _Op:
    def __init__(self, w):
        self.cigar_op, self.cigar_len = int(w & 15), int(w >> 4)


class _Flags:
    def __init__(self, rev):
        self.is_reverse = bool(rev)


class _Read:
    """type_read duck type rebuilt from a packed batch"""

    def __init__(self, b, r):
        bo, n = int(b.read_base_off[r]), int(b.read_len[r])
        co, k = int(b.read_cigar_off[r]), int(b.read_n_ops[r])
        self.pos = int(b.read_pos[r])
        self.sequence = bytes(b.bases[bo:bo + n]).decode("latin-1")
        self.base_qualities = b.quals[bo:bo + n].tolist()
        self.cigar_tuples = [_Op(w) for w in b.cigar[co:co + k]]
        self.flags = _Flags(b.read_flags[r] & 1)
        self.mapping_quality = int(b.read_mapq[r])


def check(b, r=0, chunk=(1000, 50)):
    ro, rl = int(b.region_ref_off[r]), int(b.region_ref_len[r])
    ref = bytes(b.ref[ro:ro + rl]).decode()
    start, end = int(b.region_ref_start[r]), int(b.region_ref_end[r])
    want = _ref_mod().legacy_summary(b.read_pos, b.read_base_off, b.read_len, b.read_cigar_off, b.read_n_ops, b.read_flags,
                                     b.read_mapq, b.bases, b.quals, b.cigar, int(b.region_read_begin[r]),
                                     int(b.region_read_begin[r + 1]), ref, start, end, chunk[0], chunk[1])
    g = legacy_summary.SummaryGenerator(ref, "c", start, end)
    g.generate_summary([_Read(b, i) for i in range(int(b.region_read_begin[r]), int(b.region_read_begin[r + 1]))], start, end)
    assert g.genomic_pos == [tuple(x) for x in want["genomic_pos"].tolist()]
    got = np.asarray(g.image, np.uint8).reshape(-1, 10)
    bad = np.argwhere(got != want["image"])
    assert bad.size == 0, "image differs first at %s" % bad[0].tolist()
    assert g.ref_image == list(want["ref_image"])
    assert g.longest_insert_count == dict(want["longest_insert_count"])
    s = g.chunk_image(chunk[0], chunk[1], 10)
    assert s.chunk_ids == list(want["chunk_ids"])
    assert s.images == [[list(row) for row in c] for c in want["chunk_images"]]
    assert s.positions == [[tuple(p) for p in c] for c in want["chunk_positions"]]
    assert s.refs == [list(c) for c in want["chunk_refs"]] and s.labels == [list(c) for c in want["chunk_labels"]]
    images, positions, ids = g.chunks_device(chunk[0], chunk[1])
    assert images.cpu().numpy().tolist() == s.images and list(ids) == s.chunk_ids
    assert [[tuple(p) for p in c] for c in positions.cpu().numpy().tolist()] == s.positions
    return g


def test_product_matches_the_committed_golden_vectors():
    """tests/golden/legacy_summary.json (written by the compiled reference, tests/golden/make_legacy_golden.py)"""
    import json
    import legacy_cases as LC
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "legacy_summary.json")))
    for name, (b, chunk) in LC.cases().items():
        ref = bytes(b.ref[int(b.region_ref_off[0]):int(b.region_ref_off[0]) + int(b.region_ref_len[0])]).decode()
        start, end = int(b.region_ref_start[0]), int(b.region_ref_end[0])
        g = legacy_summary.SummaryGenerator(ref, "c", start, end)
        g.generate_summary([_Read(b, i) for i in range(b.n_reads)], start, end)
        w = gold[name]
        assert g.image == w["image"] and [list(p) for p in g.genomic_pos] == w["genomic_pos"] and g.ref_image == w["ref_image"]
        assert {str(k): v for k, v in g.longest_insert_count.items()} == w["longest_insert_count"]
        s = g.chunk_image(chunk[0], chunk[1], 10)
        assert s.images == w["chunk_images"] and [[list(p) for p in c] for c in s.positions] == w["chunk_positions"]
        assert s.refs == w["chunk_refs"] and s.labels == w["chunk_labels"] and s.chunk_ids == w["chunk_ids"]


def test_hand_built_counts_mapq0_reads():
    ref = "ACGTNacgt" + "ACGT" * 8
    reads = [H.Read(0, "ACGTACGTAC", [(0, 10)]),
             H.Read(2, "GTTTACG", [(0, 2), (1, 2), (0, 3)], rev=True),
             H.Read(2, "GTTTTACG", [(0, 2), (1, 3), (0, 3)]),
             H.Read(4, "ACAC", [(0, 2), (2, 3), (0, 2)]),
             H.Read(4, "ACAC", [(0, 2), (2, 3), (0, 2)], rev=True),
             H.Read(5, "NNAC", [(4, 2), (0, 2)]),
             H.Read(30, "ACGTACGTACGTACG", [(0, 15)]),
             H.Read(1, "CG", [(0, 2)], mapq=0)]                                      # mapq 0: COUNTED by this generator
    g = check(H.one_region(ref, reads), chunk=(16, 5))
    # rows: positions 0-3, three insert rows behind position 3 ('*' -> 0), N -> 0, lower case counts like upper case
    assert g.ref_image[:12] == [1, 2, 3, 4, 0, 0, 0, 0, 1, 2, 3, 4]
    # the polisher's generator drops the mapq-0 read, this one does not: position 1 has one more forward C
    from pepper_thesis_b200 import polish
    p = polish.SummaryGenerator(ref, "c", 0, len(ref) - 1)
    p.generate_summary(reads, 0, len(ref) - 1)
    assert p.image != g.image and p.genomic_pos == g.genomic_pos


@pytest.mark.parametrize("seed", range(10))
def test_fuzz(seed):
    check(H.fuzz_region(seed, consistent=True), chunk=(64, 9))


@pytest.mark.parametrize("profile,cov", [("ont_r9", 20.0), ("hifi", 12.0)])
def test_synthetic_region(profile, cov):
    b = synth.generate(profile, 130000, cov, seed=11)
    check(b, r=0)
