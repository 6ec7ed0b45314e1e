"""Writes tests/golden/stage3_*.json from the UNMODIFIED reference stage-3 modules (CandidateFinder.py, VcfWriter.py)
run in this container through oracle/ref_stage3.py on the seeded inputs of tests/stage3_worlds.py.

    python tests/golden/make_stage3_golden.py        (needs /root/reference; the GPU box only reads the JSON files)
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import ref_stage3 as R  # noqa: E402
import stage3_worlds as W  # noqa: E402


def dict_to_list(d):
    return [[list(k), W.plain(v)] for k, v in sorted(d.items())]


def main():
    assert R.available(), "needs the reference tree"
    for seed, opt in W.FILTER_CASES:
        _, _, cands, contig, _ = W.world(seed, lower=(seed == 1))
        contigs, phasing, variant = R.ref_find_candidates(cands, [("ctg", contig)], opt)
        out = {"seed": seed, "options": list(opt), "contigs": contigs, "phasing": dict_to_list(phasing), "variant": dict_to_list(variant),
               "source": "CandidateFinder.find_candidates (unmodified, /root/reference/pepper_variant/modules/python/CandidateFinder.py:532-581)"}
        with open(os.path.join(HERE, "stage3_filter_seed%d.json" % seed), "w") as f:
            json.dump(out, f)
    vcf_contigs = [("chr1", "A" * 50), ("chr2", "C" * 40), ("chrUn", "G" * 7)]
    for seed, opt in W.VCF_CASES:
        sites = W.random_sites(seed, 800)
        counts, recs, header = R.ref_vcf(sites, opt, vcf_contigs)
        out = {"seed": seed, "options": list(opt), "counts": list(counts), "records": W.plain(recs), "header": W.plain(header),
               "source": "VCFWriter.write_vcf_records (unmodified, /root/reference/pepper_variant/modules/python/VcfWriter.py:48-221)"}
        with open(os.path.join(HERE, "stage3_vcf_seed%d.json" % seed), "w") as f:
            json.dump(out, f)
    print("ok")


if __name__ == "__main__":
    main()
