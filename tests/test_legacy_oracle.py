"""The oracle of the legacy SummaryGenerator (oracle/_ref/pv_ref_legacy = the unmodified reference file, compiled) against the
committed golden vectors it wrote (tests/golden/legacy_summary.json, tests/golden/make_legacy_golden.py). No GPU."""
import json
import os

import pytest

import legacy_cases as LC

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "legacy_summary.json")))


@pytest.mark.skipif(LC.ref_mod() is None, reason="oracle/_ref/pv_ref_legacy not built")
@pytest.mark.parametrize("name", sorted(GOLD))
def test_compiled_reference_reproduces_its_golden(name):
    b, chunk = LC.cases()[name]
    assert LC.as_golden(LC.run_reference(LC.ref_mod(), b, chunk)) == GOLD[name]


def test_golden_shapes():
    g = GOLD["hand_built"]
    assert len(g["image"]) == len(g["genomic_pos"]) == len(g["ref_image"]) == 41 + 3      # 41 positions, 3 insert rows
    assert g["ref_image"][:12] == [1, 2, 3, 4, 0, 0, 0, 0, 1, 2, 3, 4]
    assert all(len(c) == 16 for c in g["chunk_images"]) and g["chunk_positions"][-1][-1] == [-1, -1]
