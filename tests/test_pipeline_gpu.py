"""GPU: summary -> inference pipeline (device-resident and host-buffer entry points) against the oracles, and
size-independent properties at larger sizes."""
import numpy as np
import pytest
import torch

import helpers as H
import model_port as MP
import pyoracle as O
from pepper_thesis_b200 import device as dev, models, pipeline, synth

pytestmark = pytest.mark.gpu


def _hot_path(profile, wrap=False, group=3):
    sd = models.random_variant_state_dict(0)
    model = models.TransducerGRU().load_state_dict(sd)
    return pipeline.HotPath(model, synth.PROFILES[profile].thresholds, "cuda", group_regions=group, wrap_int8=wrap), sd


def test_pipeline_matches_oracles_end_to_end():
    b = synth.generate("ont_r9", 520000, 25.0, seed=21)          # 6 regions, groups of 3 -> two uploads
    hp, sd = _hot_path("ont_r9")
    pred = hp.run_host(b)
    thr = synth.PROFILES["ont_r9"].thresholds
    pos, alle, imgs = [], [], []
    for r in range(b.n_regions):
        o = O.ref_summary(b, r, thr) if O.have_ref() else O.port_summary(b, r, thr)
        pos.append(np.asarray(o["position"])); alle += list(o["alleles"]); imgs.append(np.asarray(o["images"]))
    pos = np.concatenate(pos); imgs = np.concatenate(imgs)
    assert np.array_equal(pred.position, pos) and pred.alleles() == alle
    want = MP.variant_forward(sd, torch.from_numpy(imgs.astype(np.float32))).numpy()
    assert np.abs(pred.probs - want).max() < 1e-2
    assert (pred.genotype == pred.probs.argmax(-1)).all()
    assert (np.diff(pred.region) >= 0).all()


def test_grouping_and_sharding_do_not_change_results():
    """Property at a larger size: any grouping of regions, and any rank sharding + merge, gives identical records."""
    b = synth.generate("hifi", 1300000, 20.0, seed=22)            # 13 regions
    hp1, _ = _hot_path("hifi", group=13)
    hp2, _ = _hot_path("hifi", group=4)
    a, c = hp1.run_host(b), hp2.run_host(b)
    assert np.array_equal(a.position, c.position) and a.alleles() == c.alleles() and np.array_equal(a.probs, c.probs)
    parts = []
    for rank in range(3):
        lo, hi = pipeline.shard_regions(b.n_regions, rank, 3)
        parts.append(hp2.run_host(b.region_range_view(lo, hi), region_offset=lo))
    m = pipeline.merge_results(parts[::-1])
    assert np.array_equal(m.region, a.region) and np.array_equal(m.position, a.position) and m.alleles() == a.alleles()
    assert np.abs(m.probs - a.probs).max() < 1e-6


def test_int8_wrap_mode_matches_reference_pipeline():
    b = H.kat_clamp()                                             # 300x coverage: features reach -300
    hp, sd = _hot_path("ont_r9", wrap=True, group=1)
    pred = hp.run_host(b)
    o = O.port_summary(b, 0, H.R9)
    x = o["images"].astype(np.int8).astype(np.float32)            # DataStore.py:68 + dataloader_predict.py:90
    want = MP.variant_forward(sd, torch.from_numpy(x)).numpy()
    assert np.abs(pred.probs - want).max() < 1e-2


def test_device_resident_api_and_counts():
    b = synth.generate("ont_r10", 200000, 40.0, seed=23)
    hp, _ = _hot_path("ont_r10")
    db = dev.DeviceBatch(b)
    out = hp.run_device(db, to_host=False)
    assert out["probs"].is_cuda and out["probs"].shape == (out["count"], 3)
    s = out["probs"].sum(1)
    assert torch.allclose(s, torch.ones_like(s), atol=1e-5)
    pred = hp.run_device(db)
    assert len(pred) == out["count"] and (np.diff(pred.position[pred.region == 0]) >= 0).all()


def test_kernel_bound_host_mode_and_quality_predicates_do_not_change_results():
    """The host path's kernel-bound schedule (no taper, ramped first groups, whole-wave inference passes with the remainder
    carried over) and the quality-predicate wire form give the records of the default schedule on the plain batch."""
    b = synth.generate("hifi", 5000000, 8.0, seed=24)             # 50 regions
    thr = synth.PROFILES["hifi"].thresholds
    hp1, _ = _hot_path("hifi", group=50)
    a = hp1.run_host(b)
    sd = models.random_variant_state_dict(0)
    model = models.TransducerGRU().load_state_dict(sd)
    hp2 = pipeline.HotPath(model, thr, "cuda", group_regions=16, wrap_int8=False, taper=False)
    hp2._wave = 256                                               # small quantum so passes really split and carry over
    q = synth.generate("hifi", 5000000, 8.0, seed=24)
    q.pack_wire(quals_pred=(thr.min_snp_baseq, thr.min_indel_baseq))
    assert q.quals_patch is not None and q.quals_packed is None
    c = hp2.run_host(q)
    assert len(a) > 1000 and len(a) == len(c)
    assert np.array_equal(a.region, c.region) and np.array_equal(a.position, c.position) and a.alleles() == c.alleles()
    assert np.array_equal(a.depth, c.depth) and np.array_equal(a.frequency, c.frequency)
    assert np.abs(a.probs - c.probs).max() < 1e-6 and np.array_equal(a.genotype, c.genotype)
    # a batch packed for other thresholds is refused
    other = synth.Thresholds(*([thr.min_snp_baseq + 1.0] + thr.as_list9()[1:] + [thr.skip_indels]))
    hp3 = pipeline.HotPath(model, other, "cuda", group_regions=16, wrap_int8=False)
    from pepper_thesis_b200 import capi
    with pytest.raises(capi.PvError):
        hp3.run_host(q)


def test_full_size_properties_64mbp_50x():
    """BASELINE.json configs[1] at its full size (64 Mbp, 50x ONT R9: 640 regions, ~3 G read bases) through properties that do not
    need the oracle on every base: (1) seven regions drawn across the contig are bit-exact against the compiled reference /
    port, window for window; (2) groups of 160 regions and groups of 37 regions give the same records (a checksum of every
    output array); (3) a second run gives the same bytes (the kernels' atomics are order-free); (4) records are ordered by
    (region, position) and every genotype carries the largest probability."""
    import hashlib
    b = synth.generate("ont_r9", 64000000, 50.0, seed=7, threads=16)
    assert b.n_regions == 640
    thr = synth.PROFILES["ont_r9"].thresholds
    hp160, _ = _hot_path("ont_r9", wrap=True, group=160)
    hp37, _ = _hot_path("ont_r9", wrap=True, group=37)

    def digest(p):
        h = hashlib.sha256()
        for a in (p.region, p.position, p.depth, p.frequency, p.allele, p.allele_len, p.genotype):
            h.update(np.ascontiguousarray(a).tobytes())
        return h.hexdigest()
    a = hp160.run_host(b)
    c = hp37.run_host(b)
    a2 = hp160.run_host(b)
    assert len(a) > 60000
    assert digest(a) == digest(c) == digest(a2)
    assert np.array_equal(a.probs, a2.probs) and np.abs(a.probs - c.probs).max() < 1e-6
    key = a.region.astype(np.int64) * (1 << 32) + a.position
    assert (np.diff(key) >= 0).all()
    # the genotype is taken on the logits: its probability is the largest one (two logits closer than float32's softmax can
    # tell apart give equal probabilities, where numpy's argmax would name the first)
    assert np.array_equal(a.probs[np.arange(len(a)), a.genotype], a.probs.max(-1))
    assert (a.genotype != a.probs.argmax(-1)).mean() < 1e-3
    # windows and records of sampled regions against the oracle
    rng = np.random.default_rng(3)
    for r in sorted(set([0, 639] + [int(x) for x in rng.integers(1, 639, 5)])):
        o = O.ref_summary(b, r, thr) if O.have_ref() else O.port_summary(b, r, thr)
        m = a.region == r
        assert np.array_equal(a.position[m], np.asarray(o["position"])), r
        assert [x for x, k in zip(a.alleles(), m) if k] == list(o["alleles"]), r
        assert np.array_equal(a.depth[m], np.asarray(o["depth"])) and np.array_equal(a.frequency[m], np.asarray(o["frequency"])), r
        db = dev.DeviceBatch(b.region_range_view(r, r + 1))
        ws = dev.SummaryWorkspace.for_batch(db, 8192)
        dev.summary_regions(db, thr, ws)
        k = int(ws.count.item())
        assert k == int(m.sum())
        assert np.array_equal(ws.windows[:k].cpu().numpy().astype(np.int32), np.asarray(o["images"]).astype(np.int32)), r


def test_inline_packing_of_plain_batches_does_not_change_results():
    """HotPath(pack_inline=True): the plain arrays are squeezed into 2-bit bases / 16-bit CIGAR group by group on host
    threads (pv_pack_group) and expanded on the device -- same records, bit-identical probabilities, fewer bytes on the
    wire; a batch with non-ACGT bases and one with an op too long for 16 bits take the same path."""
    b = synth.generate("ont_r9", 2100000, 14.0, seed=27)          # 21 regions
    b.scan_min_qual()
    hp, _ = _hot_path("ont_r9", group=4)
    a = hp.run_host(b)
    plain_bytes = hp.last_h2d_bytes
    hp.pack_inline, hp.pack_cigar = True, True                    # bases -> 2 bits, CIGAR words -> 16 bits
    for _ in range(2):                                            # second pass reuses the staging ring
        c = hp.run_host(b)
    assert hp.last_h2d_bytes < 0.35 * plain_bytes
    assert len(a) > 500 and len(a) == len(c)
    assert np.array_equal(a.region, c.region) and np.array_equal(a.position, c.position) and a.alleles() == c.alleles()
    assert np.array_equal(a.depth, c.depth) and np.array_equal(a.frequency, c.frequency)
    assert np.array_equal(a.probs, c.probs) and np.array_equal(a.genotype, c.genotype)
    # exceptions (lower case, N); CIGAR words left plain (the default)
    q = synth.generate("ont_r9", 600000, 10.0, seed=28)
    q.bases = q.bases.copy(); q.cigar = q.cigar.copy()
    real = np.nonzero(q.bases)[0]
    q.bases[real[::97]] = ord("N"); q.bases[real[5::389]] = ord("c")
    hp2, _ = _hot_path("ont_r9", group=2)
    w = hp2.run_host(q)
    hp2.pack_inline = True
    v = hp2.run_host(q)
    assert len(w) > 50 and np.array_equal(w.position, v.position) and w.alleles() == v.alleles() and np.array_equal(w.probs, v.probs)


def test_image_file_of_a_summary_has_the_reference_schema(tmp_path):
    """datastore.write_summary_from_workspace: the candidates of a group as ``summaries/<name>`` of the reference's image file
    (DataStore.py:54-71), windows wrapped to int8 like np.int8 does."""
    from pepper_thesis_b200 import datastore, hdf5_lite
    b = H.kat_clamp()                                             # 300x coverage: window values reach -300
    hp, _ = _hot_path("ont_r9", group=1)
    ws, k = hp.summarize(dev.DeviceBatch(b))
    assert k > 0
    p = str(tmp_path / "images.hdf")
    with datastore.DataStore(p, "w") as ds:
        datastore.write_summary_from_workspace(ds, "c_0_100", ws, k, ["c"])
    r = hdf5_lite.Reader(p)
    g = "summaries/c_0_100/"
    win = ws.windows[:k].cpu().numpy()
    assert win.min() < -128                                        # the wrap is exercised
    assert r[g + "images"].dtype == np.int8 and np.array_equal(r[g + "images"], win.astype(np.int8))
    assert np.array_equal(r[g + "positions"], ws.position[:k].cpu().numpy().astype(np.int32))
    o = O.port_summary(b, 0, H.R9)
    assert [x[0].encode("latin-1") for x in r[g + "candidates"].tolist()] == list(o["alleles"])
    assert r[g + "contigs"].tolist() == [b"c"] * k


def test_stage2_from_an_image_file_equals_the_tensor_hand_off(tmp_path):
    """datastore.predict_hdf5: image file in, prediction file out (the reference's file-based stage 2) gives the probabilities
    of the in-memory path with the int8 wrap switched on."""
    from pepper_thesis_b200 import datastore, hdf5_lite
    b = synth.generate("ont_r9", 320000, 20.0, seed=31)           # 3 regions
    hp, sd = _hot_path("ont_r9", wrap=True, group=3)
    want = hp.run_host(b)
    p, q = str(tmp_path / "images.hdf"), str(tmp_path / "pred.hdf")
    with datastore.DataStore(p, "w") as ds:
        for r in range(b.n_regions):
            ws, k = hp.summarize(dev.DeviceBatch(b.region_range_view(r, r + 1)))
            datastore.write_summary_from_workspace(ds, "chr_%d" % r, ws, k, ["chr"])
    n = datastore.predict_hdf5(p, q, hp.model, pass_windows=500)
    rd = hdf5_lite.Reader(q)
    batches = sorted(rd.keys("predictions"), key=lambda s: int(s.split("_")[1]))
    probs = np.concatenate([rd["predictions/%s/base_prediction" % x] for x in batches])
    pos = np.concatenate([rd["predictions/%s/positions" % x] for x in batches])
    assert n == len(want) == probs.shape[0] > 300 and len(batches) >= 3
    assert np.array_equal(pos, want.position.astype(np.int32))
    assert np.abs(probs - want.probs.astype(np.float64)).max() < 1e-6
