"""Stage-3 VCF writer (pepper_thesis_b200/vcf_writer.py) against
  * golden vectors from the UNMODIFIED reference VcfWriter.py (run through oracle/ref_stage3.py with a recording stand-in
    for pysam; tests/golden/make_stage3_golden.py): which record goes to which of the five files with which fields,
  * the restatement of the site logic (oracle/vcf_writer_port.py, VcfWriter.py:48-221), itself checked against the same
    vectors, on more random candidate sets,
and the files it writes (BGZF framing, header, tabix index). CPU only. What stays unpinned: the text pysam/htslib would
render for a float field (pysam is absent); fields are compared numerically."""
import gzip
import json
import os
import struct

import numpy as np
import pytest

import stage3_worlds as W
import vcf_writer_port as port
from pepper_thesis_b200 import vcf_writer as vw


_random_sites = W.random_sites


def _golden(seed):
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stage3_vcf_seed%d.json" % seed)) as f:
        return json.load(f)


def _norm(r):
    """A record dict (port or golden) -> comparable plain values; the file list as a sorted list."""
    d = W.plain({k: v for k, v in r.items() if k != "id"})
    d["files"] = sorted(d["files"]); d["alleles"] = list(d["alleles"])
    return d


@pytest.mark.parametrize("seed,opt", W.VCF_CASES)
def test_port_records_match_reference_golden(seed, opt):
    """The restatement against the records the UNMODIFIED VcfWriter.py handed to (a recording stand-in of) pysam."""
    g = _golden(seed)
    want = [_norm(r) for r in g["records"]]
    got = [_norm(r) for r in port.records(W.random_sites(seed, 800), vw.VcfOptions(*opt))]
    assert len(got) == len(want) == g["counts"][0]
    for a, b in zip(got, want):
        assert a == b, (a, b)


def test_live_reference_reproduces_vcf_golden():
    import ref_stage3 as R
    if not R.available():
        pytest.skip("no /root/reference here")
    for seed, opt in W.VCF_CASES:
        g = _golden(seed)
        counts, recs, header = R.ref_vcf(W.random_sites(seed, 800), opt, [("chr1", "A" * 50), ("chr2", "C" * 40), ("chrUn", "G" * 7)])
        assert list(counts) == g["counts"] and [_norm(r) for r in recs] == [_norm(r) for r in g["records"]]
        assert W.plain(header) == g["header"]


@pytest.mark.parametrize("seed,opt", W.VCF_CASES)
def test_files_match_reference_golden(tmp_path, seed, opt):
    """The product writer's five .vcf.gz files against the unmodified reference's records: same records in the same
    files, same header lines (FILTER / FORMAT ids, contigs, sample)."""
    g = _golden(seed)
    w = vw.VCFWriter(["chr1", "chr2"], [("chr1", 50), ("chr2", 40), ("chrUn", 7)], "HG002", str(tmp_path) + "/", "FULL", "PEPPER", "VC")
    counts = w.write_vcf_records(W.random_sites(seed, 800), vw.VcfOptions(*opt))
    w.close()
    assert list(counts) == g["counts"]
    for k in vw.VCFWriter.FILES:
        head, rows = _read_vcf(w.paths[k])
        want = [r for r in g["records"] if k in r["files"]]
        assert len(rows) == len(want)
        for key, items in g["header"]["meta"]:
            d = dict(items)
            assert any(h.startswith("##%s=<ID=%s," % (key, d["ID"])) for h in head), (key, d["ID"])
        for name, length in g["header"]["contigs"]:
            assert "##contig=<ID=%s,length=%d>" % (name, length) in head
        assert head[-1].split("\t")[-1] == g["header"]["samples"][0]
        for row, r in zip(rows, want):
            assert row[0] == r["contig"] and int(row[1]) == r["start"] + 1 and row[2] == "." and row[3] == r["alleles"][0]
            assert row[4].split(",") == list(r["alleles"][1:]) and int(row[5]) == r["qual"] and row[6] == r["filter"]
            gt, ap, gq, dp, ad, vaf, rep = row[9].split(":")
            assert [int(x) for x in gt.split("/")] == r["GT"] and float(gq) == r["GQ"] and int(dp) == r["DP"] and rep == r["REP"]
            assert [int(x) for x in ad.split(",")] == r["AD"]
            assert np.allclose([float(x) for x in vaf.split(",")], r["VAF"], rtol=1e-5, atol=0)
            assert np.allclose([float(x) for x in ap.split(",")], r["AP"], rtol=1e-5, atol=0)


@pytest.mark.parametrize("seed", range(5))
def test_merge_site_matches_port(seed):
    for allowed in (1, 2, 4):
        for key, cands in _random_sites(seed).items():
            want = port.candidate_list_to_variant(cands, allowed)
            got = vw.merge_site(cands, allowed)
            assert got[:8] == want[:8], key
            assert float(got[8]) == float(want[8]) and [float(x) for x in got[9]] == [float(x) for x in want[9]] and got[10] == want[10]


def _read_vcf(path):
    with gzip.open(path, "rt") as f:
        lines = f.read().splitlines()
    head = [l for l in lines if l.startswith("#")]
    return head, [l.split("\t") for l in lines if not l.startswith("#")]


@pytest.mark.parametrize("seed,opts", [(11, vw.VcfOptions()), (12, vw.VcfOptions(2, 5, 30, 1, 12)), (13, vw.VcfOptions(1, 90, 90, 90, 90))])
def test_files_match_port_records(tmp_path, seed, opts):
    sites = _random_sites(seed, 2500)                          # > 64 KiB of text: several BGZF blocks
    w = vw.VCFWriter(["chr1", "chr2"], [("chr1", 5_000_000), ("chr2", 100_000), ("chrUn", 7)], "HG002", str(tmp_path) + "/", "FULL",
                     "PEPPER", "VC")
    counts = w.write_vcf_records(sites, opts)
    w.close()
    want = port.records(sites, opts)
    by_file = {k: [r for r in want if k in r["files"]] for k in vw.VCFWriter.FILES}
    assert counts == tuple(len(by_file[k]) for k in vw.VCFWriter.FILES)
    assert counts[0] == counts[1] + counts[2] and counts[2] == counts[3] + counts[4] and counts[0] > 1000
    for k in vw.VCFWriter.FILES:
        head, rows = _read_vcf(w.paths[k])
        assert head[0] == "##fileformat=VCFv4.2" and head[-1].split("\t")[-1] == "HG002"
        assert "##contig=<ID=chrUn,length=7>" in head and sum(h.startswith("##FORMAT=<ID=") for h in head) == 7
        assert len(rows) == len(by_file[k])
        for row, r in zip(rows, by_file[k]):
            assert row[0] == r["contig"] and int(row[1]) == r["start"] + 1 and row[2] == "." and row[3] == r["alleles"][0]
            assert row[4].split(",") == list(r["alleles"][1:]) and int(row[5]) == r["qual"] and row[6] == r["filter"] and row[7] == "."
            assert row[8] == "GT:AP:GQ:DP:AD:VAF:REP"
            gt, ap, gq, dp, ad, vaf, rep = row[9].split(":")
            assert [int(x) for x in gt.split("/")] == r["GT"] and float(gq) == r["GQ"] and int(dp) == r["DP"] and rep == r["REP"]
            assert [int(x) for x in ad.split(",")] == r["AD"]
            assert np.allclose([float(x) for x in vaf.split(",")], r["VAF"], rtol=1e-5, atol=0)
            assert np.allclose([float(x) for x in ap.split(",")], [float(x) for x in r["AP"]], rtol=1e-5, atol=0)
    # BGZF framing: every member carries the BC field with its own size, the file ends with the 28-byte EOF marker
    raw = open(w.paths["full"], "rb").read()
    off, blocks = 0, 0
    while off < len(raw):
        assert raw[off:off + 4] == b"\x1f\x8b\x08\x04" and raw[off + 12:off + 14] == b"BC"
        off += struct.unpack_from("<H", raw, off + 16)[0] + 1
        blocks += 1
    assert off == len(raw) and blocks >= 3
    assert raw[-28:] == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def test_fasta_path_constructor(tmp_path):
    import bamio
    fa = str(tmp_path / "r.fa")
    bamio.write_fasta(fa, [("chrS", "ACGT" * 100), ("chrT", "A" * 33)], width=60)
    w = vw.VCFWriter(["chrS"], fa, "s1", str(tmp_path) + "/", "a", "b", "c")
    assert w.contigs == ["chrS", "chrT"]
    probs = np.array([0.01, 0.04, 0.95], np.float32)
    n = w.write_vcf_records({("chrS", 10): [("chrS", 10, 11, "G", ["T"], [1, 1], 30, [28], probs[2], probs, [probs[2]], False)]}, vw.VcfOptions())
    w.close()
    assert n == (1, 0, 1, 1, 0)                                  # QUAL 13 <= snp_q_cutoff 20 -> re-genotyped
    head, rows = _read_vcf(w.paths["full"])
    assert "##contig=<ID=chrS,length=400>" in head and "##contig=<ID=chrT,length=33>" in head
    assert rows == [["chrS", "11", ".", "G", "T", "13", "PASS", ".", "GT:AP:GQ:DP:AD:VAF:REP", "1/1:0.95:13:30:28:0.933:0"]]
    assert _read_vcf(w.paths["variant_calling_snp"])[1] == rows and _read_vcf(w.paths["pepper"])[1] == []


def _bgzf_read(path):
    """-> (decompressed bytes, {compressed block offset: offset of its data in the decompressed stream})"""
    import zlib
    raw = open(path, "rb").read()
    out, starts, off = bytearray(), {}, 0
    while off < len(raw):
        size = struct.unpack_from("<H", raw, off + 16)[0] + 1
        starts[off] = len(out)
        out += zlib.decompress(raw[off + 18:off + size - 8], -15)
        off += size
    return bytes(out), starts


def test_tabix_index_finds_every_record(tmp_path):
    """close() leaves a .tbi beside every file (the reference calls pysam.tabix_index): format per the tabix specification
    (magic, VCF preset columns, names, bins with chunks, linear index); a region query through it -- bins of the region,
    chunk offsets resolved to text -- returns exactly the records that overlap the region."""
    sites = W.random_sites(21, 3000)
    w = vw.VCFWriter(["chr1", "chr2"], [("chr1", 5_000_000), ("chr2", 100_000)], "S", str(tmp_path) + "/", "F", "P", "V")
    w.write_vcf_records(sites, vw.VcfOptions())
    w.close()
    for k in vw.VCFWriter.FILES:
        text, starts = _bgzf_read(w.paths[k])
        idx, _ = _bgzf_read(w.paths[k] + ".tbi")
        magic, n_ref, fmt, c_seq, c_beg, c_end, meta, skip, l_nm = struct.unpack_from("<4s8i", idx, 0)
        assert (magic, n_ref, fmt, c_seq, c_beg, c_end, meta, skip) == (b"TBI\1", 2, 2, 1, 2, 0, ord("#"), 0)
        names = idx[36:36 + l_nm].split(b"\0")[:-1]
        assert names == [b"chr1", b"chr2"]
        p = 36 + l_nm
        rows = [l.split("\t") for l in text.decode().splitlines() if not l.startswith("#")]
        for tid in range(n_ref):
            n_bin = struct.unpack_from("<i", idx, p)[0]; p += 4
            bins = {}
            for _ in range(n_bin):
                b, n_chunk = struct.unpack_from("<Ii", idx, p); p += 8
                bins[b] = [struct.unpack_from("<QQ", idx, p + 16 * i) for i in range(n_chunk)]; p += 16 * n_chunk
            n_intv = struct.unpack_from("<i", idx, p)[0]; p += 4
            lin = struct.unpack_from("<%dQ" % n_intv, idx, p); p += 8 * n_intv
            mine = [r for r in rows if r[0] == names[tid].decode()]
            if not mine:
                assert n_bin == 0
                continue
            assert 37450 in bins and bins[37450][1][0] == len(mine)
            lo = int(mine[len(mine) // 3][1]) - 1
            hi = lo + 20000
            want = [r for r in mine if int(r[1]) - 1 < hi and int(r[1]) - 1 + len(r[3]) > lo]
            got = []
            for b, chunks in bins.items():
                if b == 37450:
                    continue
                for v0, v1 in chunks:
                    t0 = starts[v0 >> 16] + (v0 & 0xffff)
                    t1 = starts[v1 >> 16] + (v1 & 0xffff) if (v1 >> 16) in starts else len(text)
                    for line in text[t0:t1].decode().splitlines():
                        r = line.split("\t")
                        assert vw._reg2bin(int(r[1]) - 1, int(r[1]) - 1 + len(r[3])) == b      # every record sits in its bin's chunks
                        if int(r[1]) - 1 < hi and int(r[1]) - 1 + len(r[3]) > lo:
                            got.append(r)
            assert sorted(got) == sorted(want) and len(want) > 0
            # linear index: the first record at or behind each window starts no earlier than the window's offset
            first = int(mine[0][1]) - 1
            assert lin[first >> 14] == bins[vw._reg2bin(first, first + len(mine[0][3]))][0][0]
