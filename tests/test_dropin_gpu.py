"""GPU: the pybind drop-in module and the AlignmentSummarizer mirror read like the reference's own call sites
(AlignmentSummarizer.py:220-238) and give the reference's results."""
import pickle
import types

import numpy as np
import pytest

import helpers as H
import pyoracle as O
from pepper_thesis_b200 import synth

pytestmark = pytest.mark.gpu


def _pv():
    from pepper_thesis_b200.build import PEPPER_VARIANT
    return PEPPER_VARIANT


def _reads_for(mod, batch, r):
    """type_read objects of module `mod` (ours or the reference oracle's) from a packed batch."""
    out = []
    for i in range(int(batch.region_read_begin[r]), int(batch.region_read_begin[r + 1])):
        rd = mod.type_read()
        bo, n = int(batch.read_base_off[i]), int(batch.read_len[i])
        rd.pos = int(batch.read_pos[i]); rd.pos_end = rd.pos
        rd.sequence = batch.bases[bo:bo + n].tobytes().decode("latin-1")
        rd.base_qualities = batch.quals[bo:bo + n].astype(int).tolist()
        co, k = int(batch.read_cigar_off[i]), int(batch.read_n_ops[i])
        rd.cigar_tuples = [mod.CigarOp(int(w & 15), int(w >> 4)) for w in batch.cigar[co:co + k]]
        f = mod.type_read_flags(); f.is_reverse = bool(batch.read_flags[i] & 1); rd.flags = f
        rd.mapping_quality = int(batch.read_mapq[i]); rd.hp_tag = 0; rd.read_id = i
        out.append(rd)
    return out


def _call(mod, batch, r, thr):
    fo, fl = int(batch.region_ref_off[r]), int(batch.region_ref_len[r])
    g = mod.RegionalSummaryGenerator("chr", int(batch.region_ref_start[r]), int(batch.region_ref_end[r]),
                                     batch.ref[fo:fo + fl].tobytes().decode("latin-1"))
    reads = _reads_for(mod, batch, r)
    g.generate_max_insert_summary(reads)
    return g.generate_summary(reads, *thr.as_list9(), thr.skip_indels, int(batch.region_cand_start[r]),
                              int(batch.region_cand_end[r]), 32, 26, False)


def _as_dict(cands):
    return dict(position=np.array([c.position for c in cands], np.int64), depth=np.array([c.depth for c in cands]),
                frequency=np.array([c.candidate_frequency[0] for c in cands]),
                alleles=[c.candidates[0].encode("latin-1") for c in cands],
                images=np.array([c.image_matrix for c in cands], np.int32).reshape(len(cands), 33, 26))


@pytest.mark.parametrize("name", ["toy", "del", "ins", "clamp"])
def test_generator_object_api_kat(name):
    b = H.KATS[name]()
    ours = _call(_pv(), b, 0, H.R9)
    assert all(c.contig == "chr" and c.base_label == 0 and c.type_label == 0 and len(c.candidates) == 1 for c in ours)
    H.assert_same(O.port_summary(b, 0, H.R9), _as_dict(ours), name)


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
def test_generator_matches_reference_objects():
    """Same Python call sequence against the reference's own pybind classes (compiled in oracle/_ref)."""
    b = synth.generate("ont_r9", 120000, 12.0, seed=4, num_regions=1)
    thr = synth.PROFILES["ont_r9"].thresholds
    ours, ref = _call(_pv(), b, 0, thr), _call(O.ref_module(), b, 0, thr)
    assert len(ours) == len(ref) > 10
    H.assert_same(_as_dict(ref), _as_dict(ours), "object api")
    c = pickle.loads(pickle.dumps(ours[0]))
    assert (c.position, c.depth, c.candidates, c.image_matrix) == (ours[0].position, ours[0].depth, ours[0].candidates, ours[0].image_matrix)


def test_alignment_summarizer_mirror():
    from pepper_thesis_b200.summarizer import AlignmentSummarizer
    b = synth.generate("ont_r10", 300000, 10.0, seed=6, first_region=1, num_regions=1)
    thr = synth.PROFILES["ont_r10"].thresholds
    pv = _pv()
    reads = _reads_for(pv, b, 0)
    rs, re_ = int(b.region_ref_start[0]), int(b.region_ref_end[0])

    class Bam:
        def get_reads(self, chrom, start, stop, supp, mapq, baseq):
            assert (start, stop) == (rs, re_)
            return reads

    class Fasta:
        def get_reference_sequence(self, chrom, start, stop):
            assert (start, stop) == (rs, re_ + 1)
            return b.ref.tobytes().decode("latin-1")

    opt = types.SimpleNamespace(train_mode=False, include_supplementary=False, min_mapq=1, min_snp_baseq=thr.min_snp_baseq,
                                min_indel_baseq=thr.min_indel_baseq, snp_frequency=thr.snp_freq, insert_frequency=thr.insert_freq,
                                delete_frequency=thr.delete_freq, min_coverage_threshold=thr.min_coverage,
                                snp_candidate_frequency_threshold=thr.snp_candidate_freq,
                                indel_candidate_frequency_threshold=thr.indel_candidate_freq,
                                candidate_support_threshold=thr.candidate_support, skip_indels=False, downsample_rate=1.0)
    out = AlignmentSummarizer(Bam(), Fasta(), "chr", int(b.region_cand_start[0]), int(b.region_cand_end[0])).create_summary(opt, None, 0)
    H.assert_same(O.port_summary(b, 0, thr), _as_dict(out), "summarizer")


def test_out_of_scope_names_raise():
    pv = _pv()
    with pytest.raises(RuntimeError, match="outside the B200 hot path"):
        pv.RegionalSummaryGeneratorHP("c", 0, 3, "ACGT")
    # the legacy per-position generator is built (legacy_summary.py): the module forwards the name
    lg = pv.SummaryGenerator("ACGT", "c", 0, 3)
    lg.generate_summary([H.Read(0, "ACGT", [(0, 4)])], 0, 3)
    assert lg.genomic_pos == [(0, 0), (1, 0), (2, 0), (3, 0)] and lg.ref_image == [1, 2, 3, 4] and lg.image[0][4] == 254
    assert pv.ImageSummary().chunk_ids == []
    g = pv.RegionalSummaryGenerator("c", 0, 3, "ACGT")
    with pytest.raises(RuntimeError):
        g.generate_summary([], 1, 1, .1, .1, .1, 1, .1, .1, 1, False, 0, 3, 32, 26, True)
