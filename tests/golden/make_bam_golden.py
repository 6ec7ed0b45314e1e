"""Writes tests/golden/bam_get_reads.json: digests of what the UNMODIFIED reference bam_handler.cpp / fasta_handler.cpp
(oracle/_ref/pv_ref_bam = those files compiled over oracle/hts_mini) return on the seeded test BAM / FASTA of
tests/test_ingest.py. Run in the build container (needs /root/reference for `make -C oracle ref`):
    python tests/golden/make_bam_golden.py"""
import hashlib
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import bamio                      # noqa: E402
import test_ingest as T           # noqa: E402


def main():
    mod = T.ref_bam_module()
    assert mod is not None, "build oracle/_ref first (make -C oracle ref)"
    d = tempfile.mkdtemp()
    ref, recs = T._records_from_synth()
    bam, fa = os.path.join(d, "t.bam"), os.path.join(d, "t.fa")
    bamio.write_bam(bam, [("chrS", T.CONTIG_LEN), ("chrT", 5000)], recs, header_text="@HD\tVN:1.6\n", block=0x8000)
    bamio.write_fasta(fa, [("chrS", ref.lower()[:500] + ref[500:]), ("chrT", "ACGTN" * 1000)], width=70)
    h = mod.BAM_handler(bam)
    out = {"get_reads": {}, "fasta": {}, "source": "reference bam_handler.cpp / fasta_handler.cpp via oracle/_ref/pv_ref_bam"}
    for span in T.SPANS:
        for supp, mapq, baseq in T.FILTERS:
            key = "%d-%d supp=%d mapq=%d baseq=%d" % (span[0], span[1], supp, mapq, baseq)
            out["get_reads"][key] = T._digest(T._ref_reads(h, "chrS", span[0], span[1], supp, mapq, baseq))
    f = mod.FASTA_handler(fa)
    for c, a, b in [("chrS", 0, 1), ("chrS", 0, 70), ("chrS", 69, 71), ("chrS", 1, 1000), ("chrS", 433, 567), ("chrS", T.CONTIG_LEN - 10, T.CONTIG_LEN),
                    ("chrS", T.CONTIG_LEN - 10, T.CONTIG_LEN + 50), ("chrT", 3, 12), ("chrT", 4990, 5600), ("chrS", 0, T.CONTIG_LEN)]:
        out["fasta"]["%s:%d-%d" % (c, a, b)] = hashlib.sha256(f.get_reference_sequence(c, a, b)).hexdigest()[:24]
    json.dump(out, open(os.path.join(HERE, "bam_get_reads.json"), "w"), indent=1)
    print("wrote", len(out["get_reads"]), "get_reads digests,", len(out["fasta"]), "fasta digests")


if __name__ == "__main__":
    main()
