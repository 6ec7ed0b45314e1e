"""BAM/FASTA ingest (libpv_ingest.so) against files written by the pure-Python writer in tests/bamio.py, an independent
Python BAM decoder, the restatement of BAM_handler::get_reads in oracle/bam_port.py and -- round 2 --
the UNMODIFIED reference bam_handler.cpp / fasta_handler.cpp compiled over a minimal htslib stand-in
(oracle/_ref/pv_ref_bam, oracle/ref_shim_bam.cpp, oracle/hts_mini/) plus digests of its output (tests/golden/bam_get_reads.json)."""
import os
import struct

import numpy as np
import pytest

import hashlib
import json
import sys

import bam_port
import bamio
from pepper_thesis_b200 import ingest, synth
from pepper_thesis_b200.summarizer import reservoir_downsample

CONTIG_LEN = 120000


def _records_from_synth(seed=3, cov=12.0):
    """Whole (un-clipped) synthetic ONT reads of one contig -> BAM records with edge cases sprinkled in."""
    b = synth.generate("ont_r9", CONTIG_LEN, cov, seed=seed, region_size=CONTIG_LEN, margin=0)
    ref = bytes(b.ref[:CONTIG_LEN]).decode()
    rng = np.random.RandomState(seed)
    recs = []
    for i in range(b.n_reads):
        bo, n = int(b.read_base_off[i]), int(b.read_len[i])
        co, k = int(b.read_cigar_off[i]), int(b.read_n_ops[i])
        seq = bytes(b.bases[bo:bo + n]).decode()
        qual = bytes(b.quals[bo:bo + n])
        cigar = [(int(c) & 15, int(c) >> 4) for c in b.cigar[co:co + k]]
        flag = 0x10 if b.read_flags[i] & 1 else 0
        mapq = 60
        tags = b""
        u = rng.randint(0, 100)
        if u < 8:                                   # leading soft clip (+ hard clip in front of it)
            s = int(rng.randint(1, 40))
            seq = "ACGT" * (s // 4 + 1)
            seq = seq[:s] + bytes(b.bases[bo:bo + n]).decode()
            qual = bytes([7] * s) + qual
            cigar = [(5, 11), (4, s)] + cigar
        elif u < 16:                                # trailing soft clip
            s = int(rng.randint(1, 40))
            seq = seq + "T" * s
            qual = qual + bytes([9] * s)
            cigar = cigar + [(4, s), (5, 3)]
        elif u < 19: flag |= 0x100                  # secondary
        elif u < 23: flag |= 0x800                  # supplementary
        elif u < 25: flag |= 0x400                  # duplicate
        elif u < 27: flag |= 0x200                  # QC fail
        elif u < 29: flag |= 0x4                    # unmapped but placed
        elif u < 35: mapq = int(rng.randint(0, 20))
        if u % 3 == 0:
            t = "cCsSiI"[u % 6]
            fmt = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I"}[t]
            tags += b"NMi" + struct.pack("<i", 5) + b"RGZgrp1\0" + b"HP" + t.encode() + struct.pack(fmt, 1 + u % 2)
        elif u % 3 == 1:
            tags += b"XAA" + b"q" + b"ZBBs" + struct.pack("<I", 3) + struct.pack("<3h", 1, 2, 3) + b"XFf" + struct.pack("<f", 1.5)
        recs.append(dict(tid=0, pos=int(b.read_pos[i]), mapq=mapq, flag=flag, name="read%05d" % i, cigar=cigar, seq=seq,
                         qual=qual, tags=tags))
    # a read whose CIGAR has more than 65535 ops (CG tag), one with SEQ '*', one using N (ref skip) and =/X ops
    n_long = 70000
    lseq = "".join("ACGT"[(j * 7 + j // 3) % 4] for j in range(n_long))
    recs.append(dict(tid=0, pos=20000, mapq=50, flag=0, name="longcigar", seq=lseq, qual=bytes([20 + j % 10 for j in range(n_long)]),
                     cigar=[(7 if j % 2 == 0 else 8, 1) for j in range(n_long)], tags=b"HPC\x02"))
    recs.append(dict(tid=0, pos=30000, mapq=50, flag=0, name="noseq", seq="", qual=b"", cigar=[(0, 500)], tags=b""))
    recs.append(dict(tid=0, pos=40100, mapq=50, flag=0x10, name="refskip", seq="A" * 300, qual=bytes([30] * 300),
                     cigar=[(0, 100), (3, 5000), (0, 100), (1, 20), (0, 50), (2, 10), (0, 30)], tags=b""))
    recs.append(dict(tid=1, pos=100, mapq=60, flag=0, name="other_contig", seq="ACGT" * 25, qual=bytes([30] * 100), cigar=[(0, 100)], tags=b""))
    recs.sort(key=lambda r: (r["tid"], r["pos"]))
    return ref, recs


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("ingest")
    ref, recs = _records_from_synth()
    bam = str(d / "t.bam")
    fa = str(d / "t.fa")
    header = ("@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:chrS\tLN:%d\n@SQ\tSN:chrT\tLN:5000\n"
              "@RG\tID:grp1\tSM:HG002\tPL:ONT\n@RG\tID:grp2\tPL:ONT\tSM:HG003:x\n@RG\tID:grp3\tSM:HG002\n" % CONTIG_LEN)
    bamio.write_bam(bam, [("chrS", CONTIG_LEN), ("chrT", 5000)], recs, header_text=header, block=0x8000)
    bamio.write_fasta(fa, [("chrS", ref.lower()[:500] + ref[500:]), ("chrT", "ACGTN" * 1000)], width=70)
    text, refs, decoded = bamio.read_bam(bam)
    assert refs == [("chrS", CONTIG_LEN), ("chrT", 5000)] and len(decoded) == len(recs)
    return dict(bam=bam, fa=fa, ref=ref, recs=decoded)


_HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(_HERE, "golden", "bam_get_reads.json")
SPANS = [(0, 5000), (19900, 20100), (39000, 47000), (61234, 71234), (CONTIG_LEN - 3000, CONTIG_LEN + 100), (50000, 50001), (7, 8),
         (29000, 31000), (0, CONTIG_LEN)]
FILTERS = [(False, 0, 1), (True, 10, 1), (True, 0, 15)]


def ref_bam_module():
    """oracle/_ref/pv_ref_bam (the compiled reference) or None where it was never built."""
    d = os.path.join(os.path.dirname(_HERE), "oracle", "_ref")
    if not (os.path.isdir(d) and any(f.startswith("pv_ref_bam") for f in os.listdir(d))):
        return None
    if d not in sys.path:
        sys.path.insert(0, d)
    import pv_ref_bam
    return pv_ref_bam


def _ref_reads(handler, contig, a, b, supp, mapq, baseq):
    """Reads of the compiled reference as the dicts bam_port returns. The record with SEQ '*' is left out: the reference
    indexes seq / qual of that record unchecked (bam_handler.cpp:210-213: reads past the record, undefined)."""
    out = []
    for r in handler.get_reads(contig, a, b, supp, mapq, baseq):
        if r["query_name"] == "noseq":
            continue
        r = dict(r)
        r["sequence"] = r["sequence"].decode()
        r["cigar_tuples"] = [tuple(c) for c in r["cigar_tuples"]]
        out.append(r)
    return out


def _digest(reads):
    h = hashlib.sha256()
    for w in reads:
        h.update(repr((w["query_name"], w["pos"], w["pos_end"], w["sequence"], list(w["base_qualities"]), [tuple(c) for c in w["cigar_tuples"]],
                       bool(w["is_reverse"]), w["mapping_quality"], w["hp_tag"])).encode())
    return "%d:%s" % (len(reads), h.hexdigest()[:24])


def _assert_reads_equal(got: ingest.IngestedReads, want, lo=0):
    b = got.batch
    assert b.n_reads - lo >= 0
    for j, w in enumerate(want):
        i = lo + j
        n = int(b.read_len[i]); bo = int(b.read_base_off[i]); co = int(b.read_cigar_off[i]); k = int(b.read_n_ops[i])
        assert bo % 16 == 0
        assert got.query_names[i] == w["query_name"], (i, got.query_names[i], w["query_name"])
        assert int(b.read_pos[i]) == w["pos"] and int(got.pos_end[i]) == w["pos_end"], w["query_name"]
        assert bytes(b.bases[bo:bo + n]).decode() == w["sequence"], w["query_name"]
        assert list(b.quals[bo:bo + n]) == list(w["base_qualities"]), w["query_name"]
        assert [(int(c) & 15, int(c) >> 4) for c in b.cigar[co:co + k]] == w["cigar_tuples"], w["query_name"]
        assert bool(b.read_flags[i] & 1) == w["is_reverse"] and int(b.read_mapq[i]) == w["mapping_quality"]
        assert int(got.hp_tag[i]) == w["hp_tag"], w["query_name"]


def test_header_and_samples(files):
    bam = ingest.BAMHandler(files["bam"])
    assert bam.get_chromosome_sequence_names() == ["chrS", "chrT"]
    assert bam.get_chromosome_sequence_names_with_length() == [("chrS", CONTIG_LEN), ("chrT", 5000)]
    assert bam.get_sample_names() == {"HG002", "HG003"}


@pytest.mark.parametrize("span", [(0, 5000), (19900, 20100), (39000, 47000), (61234, 71234), (CONTIG_LEN - 3000, CONTIG_LEN + 100),
                                  (50000, 50001), (7, 8)])
@pytest.mark.parametrize("supp,min_mapq", [(False, 0), (True, 10)])
def test_get_reads_matches_port(files, span, supp, min_mapq):
    bam = ingest.BAMHandler(files["bam"])
    got = bam.get_reads_packed("chrS", span[0], span[1], supp, min_mapq, 1)
    want = bam_port.get_reads(files["recs"], 0, span[0], span[1], supp, min_mapq, 1)
    assert got.batch.n_reads == len(want)
    _assert_reads_equal(got, want)
    if span == (19900, 20100):
        assert any(w["query_name"] == "longcigar" for w in want)
    if span == (39000, 47000):
        names = [w["query_name"] for w in want]
        assert "refskip" in names and "noseq" not in names


def _packed_as_dicts(got):
    b = got.batch
    out = []
    for i in range(b.n_reads):
        n = int(b.read_len[i]); bo = int(b.read_base_off[i]); co = int(b.read_cigar_off[i]); k = int(b.read_n_ops[i])
        out.append(dict(query_name=got.query_names[i], pos=int(b.read_pos[i]), pos_end=int(got.pos_end[i]),
                        sequence=bytes(b.bases[bo:bo + n]).decode(), base_qualities=[int(q) for q in b.quals[bo:bo + n]],
                        cigar_tuples=[(int(c) & 15, int(c) >> 4) for c in b.cigar[co:co + k]], is_reverse=bool(b.read_flags[i] & 1),
                        mapping_quality=int(b.read_mapq[i]), hp_tag=int(got.hp_tag[i])))
    return out


def test_get_reads_matches_compiled_reference(files):
    """PINNED: the product's get_reads against the reference's own bam_handler.cpp (compiled over oracle/hts_mini, which
    finds the records by a linear scan, so the product's BAI query is checked too), the restatement against the same, and
    all three against the committed digests."""
    golden = json.load(open(GOLDEN))
    mod = ref_bam_module()
    ref = mod.BAM_handler(files["bam"]) if mod else None
    bam = ingest.BAMHandler(files["bam"])
    n_checked = 0
    for span in SPANS:
        for supp, mapq, baseq in FILTERS:
            key = "%d-%d supp=%d mapq=%d baseq=%d" % (span[0], span[1], supp, mapq, baseq)
            port = bam_port.get_reads(files["recs"], 0, span[0], span[1], supp, mapq, baseq)
            got = bam.get_reads_packed("chrS", span[0], span[1], supp, mapq, baseq)
            assert _digest(port) == golden["get_reads"][key], key
            assert _digest(_packed_as_dicts(got)) == golden["get_reads"][key], key
            if ref is not None:
                want = _ref_reads(ref, "chrS", span[0], span[1], supp, mapq, baseq)
                assert _digest(want) == golden["get_reads"][key], key
                _assert_reads_equal(got, want)
                n_checked += len(want)
    assert ref is None or n_checked > 500
    if ref is not None:
        assert ref.get_chromosome_sequence_names() == bam.get_chromosome_sequence_names()
        assert set(mod.BAM_handler(files["bam"]).get_sample_names()) == bam.get_sample_names()
        assert [tuple(x) for x in mod.BAM_handler(files["bam"]).get_chromosome_sequence_names_with_length()] == \
            bam.get_chromosome_sequence_names_with_length()
        other = _ref_reads(ref, "chrT", 0, 5000, False, 0, 0)
        assert [r["query_name"] for r in other] == ["other_contig"]


def test_bad_indicies_match_compiled_reference(files):
    """type_read.bad_indicies of the drop-in module (bam_handler.cpp:216-222, :307) against the compiled reference."""
    mod = ref_bam_module()
    if mod is None:
        pytest.skip("oracle/_ref/pv_ref_bam not built")
    from pepper_thesis_b200.build import PEPPER_VARIANT as PV
    ours = PV.BAM_handler(files["bam"]).get_reads("chrS", 39000, 47000, True, 0, 15)
    want = _ref_reads(mod.BAM_handler(files["bam"]), "chrS", 39000, 47000, True, 0, 15)
    assert len(ours) == len(want) > 5
    for r, w in zip(ours, want):
        assert r.query_name == w["query_name"] and r.bad_indicies == list(w["bad_indicies"])
        assert r.flags.is_supplementary == w["is_supplementary"]


def test_fasta_matches_compiled_reference(files):
    mod = ref_bam_module()
    golden = json.load(open(GOLDEN))
    fa = ingest.FASTAHandler(files["fa"])
    ref = mod.FASTA_handler(files["fa"]) if mod else None
    for c, a, b in [("chrS", 0, 1), ("chrS", 0, 70), ("chrS", 69, 71), ("chrS", 1, 1000), ("chrS", 433, 567), ("chrS", CONTIG_LEN - 10, CONTIG_LEN),
                    ("chrS", CONTIG_LEN - 10, CONTIG_LEN + 50), ("chrT", 3, 12), ("chrT", 4990, 5600), ("chrS", 0, CONTIG_LEN)]:
        got = fa.get_reference_sequence(c, a, b)
        key = "%s:%d-%d" % (c, a, b)
        assert hashlib.sha256(got.encode()).hexdigest()[:24] == golden["fasta"][key], key
        if ref is not None:
            assert got == ref.get_reference_sequence(c, a, b).decode(), key
    if ref is not None:
        assert ref.get_chromosome_names() == fa.get_chromosome_names()
        assert ref.get_chromosome_sequence_length("chrS") == fa.get_chromosome_sequence_length("chrS")
        assert ref.get_chromosome_sequence_length("zz") == fa.get_chromosome_sequence_length("zz")


def test_get_reads_other_contig_and_missing(files):
    bam = ingest.BAMHandler(files["bam"])
    got = bam.get_reads_packed("chrT", 0, 5000, False)
    assert got.query_names == ["other_contig"]
    assert bam.get_reads_packed("nope", 0, 5000, False).batch.n_reads == 0
    assert bam.get_reads_packed("chrS", 90000, 90000, False).batch.n_reads == 0


def test_fasta(files):
    fa = ingest.FASTAHandler(files["fa"])
    ref = files["ref"]
    assert fa.get_chromosome_names() == ["chrS", "chrT"]
    assert fa.get_chromosome_sequence_length("chrS") == CONTIG_LEN and fa.get_chromosome_sequence_length("zz") == -1
    for a, b in [(0, 1), (0, 70), (69, 71), (1, 1000), (433, 567), (CONTIG_LEN - 10, CONTIG_LEN), (CONTIG_LEN - 10, CONTIG_LEN + 50), (140, 140)]:
        assert fa.get_reference_sequence("chrS", a, b) == ref[a:b], (a, b)          # upper-cased (fasta_handler.cpp:49)
    assert fa.get_reference_sequence("chrT", 3, 12) == ("ACGTN" * 4)[3:12]
    with pytest.raises(RuntimeError):
        fa.get_reference_sequence("zz", 0, 10)


def test_open_errors(tmp_path):
    with pytest.raises(RuntimeError, match="INVALID BAM FILE"):
        ingest.BAMHandler(str(tmp_path / "missing.bam"))
    p = tmp_path / "noidx.bam"
    bamio.write_bam(str(p), [("c", 100)], [])
    os.remove(str(p) + ".bai")
    with pytest.raises(RuntimeError, match="INDEX"):
        ingest.BAMHandler(str(p))
    with pytest.raises(RuntimeError, match="INVALID FASTA"):
        ingest.FASTAHandler(str(tmp_path / "missing.fa"))


def test_ingest_regions_matches_per_interval_port(files):
    bam, fa = ingest.BAMHandler(files["bam"]), ingest.FASTAHandler(files["fa"])
    starts = [0, 20000, 40000, 60000, 100000]
    ends = [20000, 40000, 60000, 80000, CONTIG_LEN - 1]
    got = ingest.ingest_regions(bam, fa, "chrS", starts, ends, include_supplementary=False, min_mapq=5, min_baseq=1, threads=3)
    b = got.batch
    assert b.n_regions == 5
    ref = files["ref"]
    for r, (s, e) in enumerate(zip(starts, ends)):
        rs, re_ = max(0, s - 100), e + 100
        assert (int(b.region_ref_start[r]), int(b.region_ref_end[r]), int(b.region_cand_start[r]), int(b.region_cand_end[r])) == (rs, re_, s, e)
        want_ref = ref[rs:re_ + 1]
        want_ref += "N" * (re_ + 1 - rs - len(want_ref))                                # past the contig end
        ro, rl = int(b.region_ref_off[r]), int(b.region_ref_len[r])
        assert bytes(b.ref[ro:ro + rl]).decode() == want_ref
        want = bam_port.get_reads(files["recs"], 0, rs, re_, False, 5, 1)
        lo, hi = int(b.region_read_begin[r]), int(b.region_read_begin[r + 1])
        assert hi - lo == len(want)
        _assert_reads_equal(got, want, lo)
    from pepper_thesis_b200 import capi
    assert capi.load().pv_batch_validate(__import__("ctypes").byref(b.as_struct())) == 0
    # the ingest hands the smallest base quality along (PvReadBatch.min_qual; validate above checked the promise)
    true_min = min(int(b.quals[int(o):int(o) + int(n)].min()) for o, n in zip(b.read_base_off, b.read_len) if n)
    assert b.min_qual == true_min and half_is_lower_bound(bam, fa, true_min)


def half_is_lower_bound(bam, fa, true_min):
    half = ingest.ingest_regions(bam, fa, "chrS", [10000, 50000], [19999, 59999], downsample_rate=0.5).batch
    got = min(int(half.quals[int(o):int(o) + int(n)].min()) for o, n in zip(half.read_base_off, half.read_len) if n)
    return 0 <= half.min_qual <= got


def test_downsampling_matches_reference_reservoir(files):
    # reservoir on indices == reservoir on the read list (AlignmentSummarizer.py:191-208)
    for n, rate in [(50, 0.5), (6000, 1.0), (6000, 0.3), (10, 1.0)]:
        idx = ingest.reservoir_indices(n, rate)
        want = reservoir_downsample(list(range(n)), rate)
        assert (list(range(n)) if idx is None else idx.tolist()) == want
    bam, fa = ingest.BAMHandler(files["bam"]), ingest.FASTAHandler(files["fa"])
    full = ingest.ingest_regions(bam, fa, "chrS", [10000, 50000], [19999, 59999])
    half = ingest.ingest_regions(bam, fa, "chrS", [10000, 50000], [19999, 59999], downsample_rate=0.5)
    fb, hb = full.batch, half.batch
    for r in range(2):
        n = int(fb.region_read_begin[r + 1] - fb.region_read_begin[r])
        keep = ingest.reservoir_indices(n, 0.5)
        lo = int(hb.region_read_begin[r])
        assert int(hb.region_read_begin[r + 1]) - lo == len(keep)
        names = [full.query_names[int(fb.region_read_begin[r]) + int(k)] for k in keep]
        assert half.query_names[lo:lo + len(keep)] == names
    assert hb.n_bases % 16 == 0 and np.all(hb.read_base_off % 16 == 0)


def test_exports():
    lib = ingest.load()
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "pepper_ingest.h")).read()
    import re
    declared = set(re.findall(r"\b(pv_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(ingest.EXPORTS)
    for s in declared:
        assert hasattr(lib, s), s


def _oracle_reads(files, a, b, supp, mapq, baseq):
    """get_reads of the compiled reference where it was built (else the restatement, which the CPU tests pin to it)."""
    mod = ref_bam_module()
    if mod is None:
        return bam_port.get_reads(files["recs"], 0, a, b, supp, mapq, baseq)
    return _ref_reads(mod.BAM_handler(files["bam"]), "chrS", a, b, supp, mapq, baseq)


@pytest.mark.gpu
def test_bam_to_candidates_gpu(files):
    """Config-1 shape: BAM + FASTA -> ingest -> CUDA summary, against the reference oracle fed by the get_reads port."""
    import helpers as H
    import pyoracle as O
    from pepper_thesis_b200 import capi
    from pepper_thesis_b200.read_batch import Region, pack_regions
    bam, fa = ingest.BAMHandler(files["bam"]), ingest.FASTAHandler(files["fa"])
    starts, ends = [0, 30000, 60000], [30000, 60000, 90000]
    got = ingest.ingest_regions(bam, fa, "chrS", starts, ends, min_mapq=55)   # drops the N-op read: UB in the reference
    thr = synth.PROFILES["ont_r9"].thresholds
    out = capi.summary_regions_host(got.batch, thr)
    d = out.trimmed()
    total = 0
    for r, (s, e) in enumerate(zip(starts, ends)):
        rs, re_ = max(0, s - 100), e + 100
        reads = [H.Read(w["pos"], w["sequence"], w["cigar_tuples"], rev=w["is_reverse"], q=w["base_qualities"], mapq=w["mapping_quality"])
                 for w in _oracle_reads(files, rs, re_, False, 55, 0)]
        ob = pack_regions([Region("chrS", rs, re_, files["ref"][rs:re_ + 1], s, e, reads)])
        want = O.ref_summary(ob, 0, thr) if O.have_ref() else O.port_summary(ob, 0, thr)
        m = d["region"] == r
        assert np.array_equal(d["position"][m], want["position"]), "region %d positions" % r
        assert np.array_equal(d["images"][m].astype(np.int32), np.asarray(want["images"]).astype(np.int32))
        total += int(m.sum())
    assert total > 50


def test_dropin_module_handlers(files):
    """PEPPER_VARIANT.BAM_handler / FASTA_handler (pybind_api.h:223-246): same calls, type_read objects out."""
    from pepper_thesis_b200.build import PEPPER_VARIANT as PV
    bam = PV.BAM_handler(files["bam"])
    fa = PV.FASTA_handler(files["fa"])
    assert bam.get_chromosome_sequence_names() == ["chrS", "chrT"]
    sq = bam.get_chromosome_sequence_names_with_length()
    assert [(s.sequence_name, s.sequence_length) for s in sq] == [("chrS", CONTIG_LEN), ("chrT", 5000)]
    assert bam.get_sample_names() == {"HG002", "HG003"}
    reads = bam.get_reads("chrS", 39000, 47000, True, 0, 15)
    want = bam_port.get_reads(files["recs"], 0, 39000, 47000, True, 0, 15)
    assert len(reads) == len(want) > 5
    for r, w in zip(reads, want):
        assert (r.query_name, r.pos, r.pos_end, r.sequence, r.mapping_quality, r.hp_tag, r.flags.is_reverse) == \
               (w["query_name"], w["pos"], w["pos_end"], w["sequence"], w["mapping_quality"], w["hp_tag"], w["is_reverse"])
        assert r.base_qualities == list(w["base_qualities"])
        assert [(c.cigar_op, c.cigar_len) for c in r.cigar_tuples] == w["cigar_tuples"]
        bad = [j for j, (b, q) in enumerate(zip(w["sequence"], w["base_qualities"])) if q < 15 or b not in "ACGT"] + [len(w["sequence"]) + 1]
        assert r.bad_indicies == bad
    assert any(r.flags.is_supplementary for r in reads)
    assert fa.get_reference_sequence("chrS", 100, 700) == files["ref"][100:700]
    assert fa.get_chromosome_sequence_length("chrS") == CONTIG_LEN and fa.get_chromosome_names() == ["chrS", "chrT"]
    with pytest.raises(RuntimeError):
        PV.BAM_handler(files["bam"] + ".missing")


@pytest.mark.gpu
def test_create_summary_from_bam_gpu(files):
    """The reference's own call sequence (AlignmentSummarizer.create_summary) on BAM + FASTA through the drop-in module."""
    import types
    import pyoracle as O
    import helpers as H
    from pepper_thesis_b200.build import PEPPER_VARIANT as PV
    from pepper_thesis_b200.read_batch import Region, pack_regions
    from pepper_thesis_b200.summarizer import AlignmentSummarizer
    thr = synth.PROFILES["ont_r9"].thresholds
    opt = types.SimpleNamespace(include_supplementary=False, min_mapq=55, min_snp_baseq=thr.min_snp_baseq, min_indel_baseq=thr.min_indel_baseq,
                                snp_frequency=thr.snp_freq, insert_frequency=thr.insert_freq, delete_frequency=thr.delete_freq,
                                min_coverage_threshold=thr.min_coverage, snp_candidate_frequency_threshold=thr.snp_candidate_freq,
                                indel_candidate_frequency_threshold=thr.indel_candidate_freq, candidate_support_threshold=thr.candidate_support,
                                skip_indels=False, downsample_rate=1.0, train_mode=False)
    s, e = 40000, 70000
    cands = AlignmentSummarizer(PV.BAM_handler(files["bam"]), PV.FASTA_handler(files["fa"]), "chrS", s, e).create_summary(opt)
    rs, re_ = s - 100, e + 100
    reads = [H.Read(w["pos"], w["sequence"], w["cigar_tuples"], rev=w["is_reverse"], q=w["base_qualities"], mapq=w["mapping_quality"])
             for w in _oracle_reads(files, rs, re_, False, 55, 1)]
    ob = pack_regions([Region("chrS", rs, re_, files["ref"][rs:re_ + 1], s, e, reads)])
    want = O.ref_summary(ob, 0, thr) if O.have_ref() else O.port_summary(ob, 0, thr)
    assert [c.position for c in cands] == list(want["position"]) and len(cands) > 20
    assert np.array_equal(np.asarray([c.image_matrix for c in cands], np.int32), np.asarray(want["images"]).astype(np.int32))
