"""Stage-3 VCF writer (pepper_thesis_b200/vcf_writer.py) against the restatement of the reference's site logic
(oracle/vcf_writer_port.py, VcfWriter.py:48-221) on random candidate sets, and the files it writes (BGZF, header, five
outputs). CPU only; pysam is absent, so byte-level parity with the reference's files is unpinned (stated in both files)."""
import gzip
import struct

import numpy as np
import pytest

import vcf_writer_port as port
from pepper_thesis_b200 import vcf_writer as vw


def _random_sites(seed, n_sites=300):
    rng = np.random.default_rng(seed)
    alpha = "ACGT"
    sites = {}
    pos = 1000
    for _ in range(n_sites):
        pos += int(rng.integers(0, 40))                         # 0: two keys can share a start only across contigs
        contig = "chr%d" % (1 + int(rng.random() < 0.2))
        cands = []
        for _ in range(int(rng.integers(1, 7))):
            kind = rng.random()
            ref_base = alpha[int(rng.integers(0, 4))]
            if kind < 0.5:
                ref, alts = ref_base, [alpha[int(rng.integers(0, 4))]]
            elif kind < 0.75:
                ref, alts = ref_base, [ref_base + "".join(alpha[int(x)] for x in rng.integers(0, 4, int(rng.integers(1, 6))))]
            else:
                ref, alts = ref_base + "".join(alpha[int(x)] for x in rng.integers(0, 4, int(rng.integers(1, 6)))), [ref_base]
            probs = rng.dirichlet([0.6, 0.6, 0.6]).astype(np.float32)
            if rng.random() < 0.1:
                probs = np.array([0.25, 0.375, 0.375], np.float32)      # tie: first maximum wins
            if rng.random() < 0.05:
                probs = np.array([0.0, 0.0, 1.0], np.float32)           # 1 - p == 0: QUAL saturates at 90
            g = int(np.argmax(probs))
            gt = [[0, 0], [0, 1], [1, 1]][g]
            depth = int(rng.integers(1, 120))
            cands.append((contig, pos, pos + len(ref), ref, alts, gt, depth, [int(rng.integers(0, depth + 1))], probs[g], probs,
                          [max(probs[1], probs[2])], bool(rng.random() < 0.3)))
        sites[(contig, pos)] = cands
    return sites


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
