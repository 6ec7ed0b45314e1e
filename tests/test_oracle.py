"""CPU: the C restatement (oracle/region_summary_port.c) against the unmodified reference (oracle/_ref) and
against the committed golden vectors."""
import json
import os

import numpy as np
import pytest

import helpers as H
import pyoracle as O
from pepper_thesis_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
needs_ref = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (reference tree absent)")


@pytest.mark.parametrize("name", sorted(H.KATS))
def test_port_matches_golden_kat(name):
    """Golden vectors generated from the compiled reference by tests/golden/make_golden.py."""
    with open(os.path.join(GOLD, "kat_%s.json" % name)) as f:
        g = json.load(f)
    b = H.KATS[name]()
    p = O.port_summary(b, 0, H.R9)
    assert [int(x) for x in p["position"]] == g["position"]
    assert [a.decode("latin-1") for a in p["alleles"]] == g["alleles"]
    assert [int(x) for x in p["depth"]] == g["depth"]
    assert [int(x) for x in p["frequency"]] == g["frequency"]
    assert p["images"].astype(int).tolist() == g["images"]


@needs_ref
@pytest.mark.parametrize("name", sorted(H.KATS))
def test_port_matches_reference_kat(name):
    b = H.KATS[name]()
    H.assert_same(O.ref_summary(b, 0, H.R9), O.port_summary(b, 0, H.R9), name)


@needs_ref
@pytest.mark.parametrize("seed", range(40))
def test_port_matches_reference_fuzz(seed):
    b = H.fuzz_region(seed)
    thr = H.fuzz_thresholds(seed)
    H.assert_same(O.ref_summary(b, 0, thr), O.port_summary(b, 0, thr), "fuzz %d" % seed)


@needs_ref
@pytest.mark.parametrize("profile,cov", [("ont_r9", 30.0), ("ont_r10", 40.0), ("hifi", 35.0)])
def test_port_matches_reference_synthetic(profile, cov):
    b = synth.generate(profile, 250000, cov, seed=3, first_region=1, num_regions=1)
    thr = synth.PROFILES[profile].thresholds
    a = O.ref_summary(b, 0, thr)
    assert len(a["position"]) > 20
    H.assert_same(a, O.port_summary(b, 0, thr), profile)


def test_survey_kat_values():
    """The centre rows quoted in SURVEY.md section 8a."""
    p = O.port_summary(H.kat_toy(), 0, H.R9)
    assert p["alleles"] == [b"1T"] and int(p["depth"][0]) == 6 and int(p["frequency"][0]) == 3
    assert p["images"][0, 16].tolist() == [1, 4, 0, 0, -3, 2, 0, 0, -1, 0, 0, 2, 0, 0, 0, -3, 1, 0, 0, -2, 0, 0, 1, 0, 0, 0]
    p = O.port_summary(H.kat_clamp(), 0, H.R9)
    assert p["alleles"] == [b"1A", b"1T"] and p["depth"].tolist() == [125, 125] and p["frequency"].tolist() == [100, 100]
    assert p["images"][0, 16, :12].tolist() == [3, 1, 0, 0, -300, 100, 0, 0, 100, 0, -100, -100]
    assert p["images"][0, 17, :12].tolist() == [4, 0, 0, 0, -300, 0, 0, 0, 0, 0, 0, -125]
    p = O.port_summary(H.kat_ins(), 0, H.R9)
    assert p["alleles"] == [b"2GTT"]
    assert p["images"][0, 16].tolist() == [3, 0, 3, 0, -1, 0, 2, 0, 0, 0, -3, 0, 2, 0, 0, -2, 0, 1, 0, 0, 0, -3, 0, 1, 0, 0]
    p = O.port_summary(H.kat_del(), 0, H.R9)
    assert p["alleles"] == [b"3GTAC"] and int(p["depth"][0]) == 8 and int(p["frequency"][0]) == 4
