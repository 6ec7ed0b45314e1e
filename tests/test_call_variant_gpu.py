"""BAM + FASTA -> VCF through call_variant (BAM decoded on the device, summary + model kernels, stage-3 filter, VCF
writer): the selected candidates and the written records equal the same chain fed by the HOST ingest (whose reads
tests/test_ingest.py pins to the compiled reference) with one batch for all intervals, for several groupings of the
intervals and for a two-rank split of them merged by hand (interval i -> rank i % world, ImageGenerationUI.py:211)."""
import gzip
import os

import numpy as np
import pytest

import test_ingest as TI
from test_ingest import files  # noqa: F401  (fixture)
from pepper_thesis_b200 import candidate_filter, ingest, models, pipeline, synth

pytestmark = pytest.mark.gpu


def _hot(group=40):
    model = models.TransducerGRU().load_state_dict(models.random_variant_state_dict(0))
    return pipeline.HotPath(model, synth.PROFILES["ont_r9"].thresholds, "cuda", group_regions=group, wrap_int8=True)


def _records(path):
    with gzip.open(path, "rt") as f:
        return [ln.rstrip("\n") for ln in f if not ln.startswith("#")]


def _host_chain(files, opt, regions):
    from pepper_thesis_b200 import call_variant as CV
    bam, fa = ingest.BAMHandler(files["bam"]), ingest.FASTAHandler(files["fa"])
    phasing, variant = {}, {}
    for name in dict.fromkeys(n for n, _ in regions):
        ivs = [(s, e) for n, s, e in CV.contig_intervals(fa, [r for r in regions if r[0] == name], opt.region_size)]
        got = ingest.ingest_regions(bam, fa, name, [s for s, _ in ivs], [e for _, e in ivs], min_mapq=opt.min_mapq)
        pred = _hot().run_host(got.batch)
        _, p, v = candidate_filter.find_candidates(pred, got.batch, opt.filter)
        phasing.update(p); variant.update(v)
    return phasing, variant


def _same(a, b):
    assert sorted(a) == sorted(b)
    for k in a:
        assert len(a[k]) == len(b[k]), k
        for x, y in zip(a[k], b[k]):
            for u, w in zip(x, y):
                if isinstance(u, (float, np.floating, np.ndarray)) or (isinstance(u, list) and u and isinstance(u[0], (float, np.floating))):
                    assert np.allclose(np.asarray(u, np.float64), np.asarray(w, np.float64), atol=1e-6), (k, u, w)
                else:
                    assert u == w, (k, u, w)


@pytest.mark.parametrize("group_mbp", [0.02, 0.05, 16.0])
def test_bam_to_vcf_equals_host_ingest_chain(files, tmp_path, group_mbp):
    from pepper_thesis_b200 import call_variant as CV
    opt = CV.CallOptions(region_size=10000, min_mapq=5, group_mbp=group_mbp)
    regions = [("chrS", None), ("chrT", None)]
    contigs, phasing, variant, stats = CV.call_candidates(files["bam"], files["fa"], _hot(), regions, opt)
    want_p, want_v = _host_chain(files, opt, regions)
    assert stats["intervals"] == 12 + 1 and stats["candidates"] > 200 and len(variant) > 20
    _same(phasing, want_p)
    _same(variant, want_v)
    opt.predictions_hdf = str(tmp_path / "predictions.hdf")          # the reference's stage-2 file beside the run (datastore.py)
    counts, stats2, paths = CV.call_variant(files["bam"], files["fa"], _hot(), str(tmp_path / "out"), "HG002", regions, opt)
    assert counts[0] == len(_records(paths["full"])) > 0 and counts[0] == counts[1] + counts[2]
    from pepper_thesis_b200 import hdf5_lite
    rd = hdf5_lite.Reader(opt.predictions_hdf)
    batches = rd.keys("predictions")
    n_pred = sum(rd["predictions/%s/positions" % b].shape[0] for b in batches)
    assert n_pred == stats2["candidates"] and rd["predictions/%s/base_prediction" % batches[0]].shape[1] == 3
    assert set(rd["predictions/%s/contigs" % batches[0]].tolist()) == {b"chrS"}
    # stage 3 from that FILE + the FASTA (the reference's FindCandidates route) selects the candidates of the in-memory route
    from pepper_thesis_b200 import datastore
    c3, p3, v3 = datastore.find_candidates_hdf5(opt.predictions_hdf, files["fa"], opt.filter)
    _same(p3, want_p)
    _same(v3, want_v)
    assert all(os.path.exists(p + ".tbi") for p in paths.values())
    # the same records as the writer produces from the host chain's candidates
    from pepper_thesis_b200.vcf_writer import VCFWriter, VcfOptions
    out2 = str(tmp_path / "host") + "/"
    os.makedirs(out2)
    w = VCFWriter(sorted({k[0] for k in want_v}), files["fa"], "HG002", out2, "PEPPER_VARIANT_FULL", "PEPPER_VARIANT_OUTPUT_PEPPER",
                  "PEPPER_VARIANT_OUTPUT_VARIANT_CALLING")
    w.write_vcf_records(want_v, VcfOptions())
    w.close()
    for k in paths:
        assert _records(paths[k]) == _records(w.paths[k]), k


def test_rank_split_merges_to_the_single_rank_result(files):
    """Interval i belongs to rank i % world: the two ranks' candidates together are the single-rank result."""
    from pepper_thesis_b200 import call_variant as CV
    opt = CV.CallOptions(region_size=10000, min_mapq=5, group_mbp=0.03)
    regions = [("chrS", (5000, 95000))]
    _, p_all, v_all, st = CV.call_candidates(files["bam"], files["fa"], _hot(), regions, opt)
    merged_p, merged_v, n = {}, {}, 0
    for rank in range(2):
        _, p, v, s = _one_rank(files, regions, opt, rank)
        n += s["intervals"]
        for k, lst in p.items():
            merged_p.setdefault(k, []).extend(lst)
        for k, lst in v.items():
            merged_v.setdefault(k, []).extend(lst)
    assert n == st["intervals"] == 9
    def dedup(d):
        out = {}
        for k, lst in d.items():
            seen, keep = [], []
            for c in lst:
                if (c[3], c[4][0]) not in seen:
                    seen.append((c[3], c[4][0])); keep.append(c)
            out[k] = keep
        return out
    _same(dedup(merged_p), p_all)
    _same(dedup(merged_v), v_all)


def _one_rank(files, regions, opt, rank):
    """call_candidates' share of one rank without the process group: the same interval deal, no gather."""
    from pepper_thesis_b200 import call_variant as CV
    fa = ingest.FASTAHandler(files["fa"])
    ivs = [iv for i, iv in enumerate(CV.contig_intervals(fa, regions, opt.region_size)) if i % 2 == rank]
    phasing, variant, n = {}, {}, 0
    for name, s, e in ivs:
        _, p, v, st = CV.call_candidates(files["bam"], files["fa"], _hot(), [(name, (s, e))], opt)
        n += st["intervals"]
        phasing.update({k: phasing.get(k, []) + lst for k, lst in p.items()})
        variant.update({k: variant.get(k, []) + lst for k, lst in v.items()})
    return None, phasing, variant, dict(intervals=n)
