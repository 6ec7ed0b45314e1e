"""BAM decoded on the device (pepper_thesis_b200.ingest_gpu, csrc/ingest_gpu.cu) against the host ingest (which
tests/test_ingest.py pins to the compiled reference bam_handler.cpp): every array of the packed batch bit-identical, then
BAM -> candidates through the summary kernels from the device-born batch."""
import numpy as np
import pytest

import test_ingest as TI
from test_ingest import files  # noqa: F401  (fixture)
from pepper_thesis_b200 import ingest, synth

pytestmark = pytest.mark.gpu

CASES = [
    ([0, 20000, 40000, 60000, 100000], [20000, 40000, 60000, 80000, TI.CONTIG_LEN - 1], dict(min_mapq=5)),
    ([0], [TI.CONTIG_LEN - 1], dict()),
    ([39000], [47000], dict(include_supplementary=True, min_mapq=10)),
    ([0, 30000, 60000], [30000, 60000, 90000], dict(min_mapq=55)),
    ([19950, 20050, 20060], [20050, 20060, 29000], dict()),                   # tiny spans: one record feeds three regions
    ([50000], [50001], dict()),
    ([10000, 50000], [19999, 59999], dict(downsample_rate=0.5)),
]


def _compare(got, want):
    gb, wb = got.batch.to_host(), want.batch
    assert gb.n_reads == wb.n_reads and gb.n_regions == wb.n_regions
    for name in ("read_pos", "read_base_off", "read_len", "read_cigar_off", "read_n_ops", "read_flags", "read_mapq", "bases", "quals",
                 "cigar", "region_ref_start", "region_ref_end", "region_cand_start", "region_cand_end", "region_ref_off", "region_ref_len",
                 "region_read_begin", "ref"):
        a, b = getattr(gb, name), getattr(wb, name)
        assert a.shape == b.shape and np.array_equal(a, b), name
    assert np.array_equal(got.pos_end.cpu().numpy(), want.pos_end)
    assert np.array_equal(got.hp_tag.cpu().numpy(), want.hp_tag)
    assert np.array_equal(got.bam_flag.cpu().numpy().view(np.uint16), want.bam_flag)
    assert got.query_names == want.query_names
    assert gb.min_qual == wb.min_qual
    assert np.array_equal(gb.region_contig_len, wb.region_contig_len)


@pytest.mark.parametrize("case", range(len(CASES)))
def test_gpu_ingest_matches_host_ingest(files, case):
    from pepper_thesis_b200 import ingest_gpu
    starts, ends, kw = CASES[case]
    bam, fa = ingest.BAMHandler(files["bam"]), ingest.FASTAHandler(files["fa"])
    want = ingest.ingest_regions(bam, fa, "chrS", starts, ends, **kw)
    got = ingest_gpu.ingest_regions_gpu(bam, fa, "chrS", starts, ends, **kw)
    _compare(got, want)
    assert got.stats["records"] >= want.batch.n_reads * (0 if kw.get("downsample_rate") else 1) // 3


def test_gpu_ingest_other_contig_and_empty(files):
    from pepper_thesis_b200 import ingest_gpu
    bam, fa = ingest.BAMHandler(files["bam"]), ingest.FASTAHandler(files["fa"])
    got = ingest_gpu.ingest_regions_gpu(bam, fa, "chrT", [0], [4000])
    want = ingest.ingest_regions(bam, fa, "chrT", [0], [4000])
    _compare(got, want)
    assert got.query_names == ["other_contig"]


def test_gpu_ingest_refuses_corrupt_blocks(files, tmp_path):
    """A flipped payload byte must be caught on the device (decoder error or CRC-32 mismatch), not decoded into reads."""
    from pepper_thesis_b200 import capi, ingest_gpu
    raw = bytearray(open(files["bam"], "rb").read())
    raw[len(raw) // 2] ^= 0x40
    p = str(tmp_path / "bad.bam")
    open(p, "wb").write(bytes(raw))
    open(p + ".bai", "wb").write(open(files["bam"] + ".bai", "rb").read())
    bam, fa = ingest.BAMHandler(p), ingest.FASTAHandler(files["fa"])
    with pytest.raises((capi.PvError, RuntimeError)):
        ingest_gpu.ingest_regions_gpu(bam, fa, "chrS", [0], [TI.CONTIG_LEN - 1])


def test_bam_to_candidates_from_device_born_batch(files):
    """BAM + FASTA -> device decode -> summary kernels, equal to the host-ingested batch through the same kernels and to the
    reference oracle fed by the compiled reference's get_reads."""
    import torch
    import helpers as H
    import pyoracle as O
    from pepper_thesis_b200 import capi, device as dev, ingest_gpu
    from pepper_thesis_b200.read_batch import Region, pack_regions
    bam, fa = ingest.BAMHandler(files["bam"]), ingest.FASTAHandler(files["fa"])
    starts, ends = [0, 30000, 60000], [30000, 60000, 90000]
    thr = synth.PROFILES["ont_r9"].thresholds
    got = ingest_gpu.ingest_regions_gpu(bam, fa, "chrS", starts, ends, min_mapq=55)
    db = got.batch
    ws = dev.SummaryWorkspace.for_batch(db, 8192)
    dev.summary_regions(db, thr, ws)
    torch.cuda.synchronize()
    k = int(ws.count.item())
    assert ws.status() == 0 and k > 50
    pos, reg, img = ws.position[:k].cpu().numpy(), ws.region[:k].cpu().numpy(), ws.windows[:k].cpu().numpy()
    host = ingest.ingest_regions(bam, fa, "chrS", starts, ends, min_mapq=55)
    d = capi.summary_regions_host(host.batch, thr).trimmed()
    assert np.array_equal(pos, d["position"]) and np.array_equal(reg, d["region"]) and np.array_equal(img, d["images"])
    for r, (s, e) in enumerate(zip(starts, ends)):
        rs, re_ = max(0, s - 100), e + 100
        reads = [H.Read(w["pos"], w["sequence"], w["cigar_tuples"], rev=w["is_reverse"], q=w["base_qualities"], mapq=w["mapping_quality"])
                 for w in TI._oracle_reads(files, rs, re_, False, 55, 0)]
        ob = pack_regions([Region("chrS", rs, re_, files["ref"][rs:re_ + 1], s, e, reads)])
        want = O.ref_summary(ob, 0, thr) if O.have_ref() else O.port_summary(ob, 0, thr)
        m = reg == r
        assert np.array_equal(pos[m], want["position"]) and np.array_equal(img[m].astype(np.int32), np.asarray(want["images"]).astype(np.int32))


def test_polisher_summary_from_device_born_batch(files):
    """The polisher's path from a BAM (pepper/modules/python/AlignmentSummarizer.py:296-350: get_reads(chr, start, end) with
    no safe bases and mapq 0, reference [start, end + 1), SummaryGenerator, 1000 / 50 chunks) with the BAM decoded on the
    device: image, positions and chunks equal the same kernels fed by the host ingest (pinned to the compiled reference)."""
    import torch
    from pepper_thesis_b200 import ingest_gpu, polish
    bam, fa = ingest.BAMHandler(files["bam"]), ingest.FASTAHandler(files["fa"])
    starts, ends = [1000, 40000, 90000], [21000, 47500, TI.CONTIG_LEN - 1]
    got = ingest_gpu.ingest_regions_gpu(bam, fa, "chrS", starts, ends, safe_bases=0, min_mapq=0)
    want = ingest.ingest_regions(bam, fa, "chrS", starts, ends, safe_bases=0, min_mapq=0)
    a, b = polish.PolishSummary(got.batch), polish.PolishSummary(want.batch)
    assert a.n_rows == b.n_rows > 20000 and np.array_equal(a.region_rows, b.region_rows)
    assert torch.equal(a.image, b.image) and torch.equal(a.genomic_pos, b.genomic_pos)
    ca, cb = a.chunks(), b.chunks()
    assert torch.equal(ca[0], cb[0]) and torch.equal(ca[1], cb[1]) and np.array_equal(ca[2], cb[2]) and np.array_equal(ca[3], cb[3])


def _bam_with_patched_record(tmp_path, tag, patch):
    """A small BAM whose SECOND record is `patch`ed after encoding (a writer that lies about a length field)."""
    import struct
    import bamio
    rng = np.random.default_rng(4)
    recs = []
    for i in range(6):
        n = 400
        seq = "".join("ACGT"[int(x)] for x in rng.integers(0, 4, n))
        recs.append(dict(tid=0, pos=100 + 300 * i, mapq=60, flag=0, name="q%d" % i, cigar=[(0, 200), (1, 5), (0, 195)], seq=seq, qual=[30] * n))
    real = bamio.encode_record

    def enc(rec):
        b = bytearray(real(rec))
        if rec["name"] == "q1":
            patch(b, struct)
        return bytes(b)
    bamio.encode_record = enc
    try:
        path = str(tmp_path / ("bad_%s.bam" % tag))
        bamio.write_bam(path, [("chrS", 5000)], recs)
    finally:
        bamio.encode_record = real
    fa = str(tmp_path / ("bad_%s.fa" % tag))
    bamio.write_fasta(fa, [("chrS", "ACGT" * 1250)])
    return path, fa


@pytest.mark.parametrize("tag", ["n_cigar", "l_seq", "block_small", "block_large", "l_name"])
def test_lying_length_fields_are_refused_on_the_device(tmp_path, tag):
    """Records whose length fields point outside the record (valid BGZF, valid CRC: only the record parser can notice): the
    device decode reports them through its status words and the call raises (the same record / clip helpers run under ASan on
    the CPU in tests/test_bam_core_cpu.py)."""
    from pepper_thesis_b200 import capi, ingest_gpu

    def patch(b, struct):
        if tag == "n_cigar":
            struct.pack_into("<H", b, 4 + 12, 60000)             # 60000 CIGAR ops in a 600-byte record
        elif tag == "l_seq":
            struct.pack_into("<i", b, 4 + 16, 1 << 20)           # a megabase of sequence
        elif tag == "block_small":
            struct.pack_into("<i", b, 0, 10)                     # shorter than the fixed part
        elif tag == "block_large":
            struct.pack_into("<i", b, 0, 1 << 28)                # runs past the stream
        elif tag == "l_name":
            b[4 + 8] = 255                                       # the name swallows the CIGAR
    path, fa = _bam_with_patched_record(tmp_path, tag, patch)
    bam, fah = ingest.BAMHandler(path), ingest.FASTAHandler(fa)
    try:
        got = ingest_gpu.ingest_regions_gpu(bam, fah, "chrS", [0], [4000])
    except (capi.PvError, RuntimeError):
        return
    # a patched field can still describe a record that parses (l_name = 255 inside a long record): then the device must agree
    # with the host ingest on what it means
    want = ingest.ingest_regions(bam, fah, "chrS", [0], [4000])
    _compare(got, want)


@pytest.mark.parametrize("preset,cov", [("hifi", 12.0), ("ont_r10", 20.0)])
def test_other_presets_through_a_written_bam(tmp_path, preset, cov):
    """HiFi-like (15 kbp reads, few CIGAR ops) and R10-like reads of the synthetic generator written as a BAM (level-1
    DEFLATE, 64 KiB blocks: the layout of a real file) and decoded both ways: every array of the batch identical, and
    BAM -> candidates equal to the summary of the generator's own batch (the BAM round trip loses nothing)."""
    import os
    import sys
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import bamio
    import fast_bam
    from pepper_thesis_b200 import capi, device as dev, ingest_gpu
    L = 600000
    b = synth.generate(preset, L, cov, seed=9, region_size=L, margin=0)
    bam, fa = str(tmp_path / "p.bam"), str(tmp_path / "p.fa")
    fast_bam.write_bam_from_batch(bam, b, "chrS", L)
    bamio.write_fasta(fa, [("chrS", bytes(b.ref[:L]).decode())])
    bh, fh = ingest.BAMHandler(bam), ingest.FASTAHandler(fa)
    starts = list(range(0, L, 100000)); ends = [min(L - 1, s + 100000) for s in starts]
    want = ingest.ingest_regions(bh, fh, "chrS", starts, ends, min_mapq=1)
    got = ingest_gpu.ingest_regions_gpu(bh, fh, "chrS", starts, ends, min_mapq=1)
    _compare(got, want)
    thr = synth.PROFILES[preset].thresholds
    ws = dev.SummaryWorkspace.for_batch(got.batch, 16384)
    dev.summary_regions(got.batch, thr, ws)
    k = int(ws.count.item())
    d = capi.summary_regions_host(want.batch, thr).trimmed()
    assert ws.status() == 0 and k == len(d["position"]) > 100
    assert np.array_equal(ws.position[:k].cpu().numpy(), d["position"]) and np.array_equal(ws.windows[:k].cpu().numpy(), d["images"])
