"""CPU: the C-ABI library loads, exports every symbol include/pepper_b200.h declares, validates batches on the host and
fails loudly (no fallback) when no GPU is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import helpers as H
from pepper_thesis_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header():
    lib = capi.load()
    hdr = open(os.path.join(ROOT, "include", "pepper_b200.h")).read()
    declared = set(re.findall(r"\b(pv_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for s in declared:
        assert hasattr(lib, s), "missing export " + s
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    assert lib.pv_version().startswith(b"pepper_b200")


def test_batch_validate():
    lib = capi.load()
    b = H.kat_ins()
    assert lib.pv_batch_validate(C.byref(b.as_struct())) == 0
    bad = H.kat_ins()
    bad.read_base_off = bad.read_base_off + 8          # not 16-byte aligned
    assert lib.pv_batch_validate(C.byref(bad.as_struct())) == -1
    assert b"16-byte" in lib.pv_last_error()
    bad = H.kat_ins()
    bad.region_ref_len = bad.region_ref_len - 50
    assert lib.pv_batch_validate(C.byref(bad.as_struct())) == -1


def test_no_cpu_fallback_without_gpu():
    lib = capi.load()
    if lib.pv_device_count() > 0:
        pytest.skip("a GPU is present")
    b = H.kat_toy()
    with pytest.raises(capi.PvError) as e:
        capi.summary_regions_host(b, H.R9)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    from pepper_thesis_b200 import models
    m = models.TransducerGRU().load_state_dict(models.random_variant_state_dict(0))
    with pytest.raises(Exception):
        import torch
        m(torch.zeros(1, 33, 26), False)


def test_pack_regions_rejects_bad_input():
    from pepper_thesis_b200.read_batch import Region, pack_regions
    r = H.Read(0, "ACGT", [(0, 4)], q=[1, 2, 3, 300])
    with pytest.raises(ValueError):
        pack_regions([Region("c", 0, 3, "ACGT", 0, 3, [r])])
    with pytest.raises(ValueError):
        pack_regions([Region("c", 0, 9, "ACGT", 0, 9, [])])


def test_region_range_view_equals_repack():
    from pepper_thesis_b200.read_batch import select_regions
    import pyoracle as O
    b = synth.generate("hifi", 350000, 8.0, seed=9)
    thr = synth.PROFILES["hifi"].thresholds
    v = b.region_range_view(1, 3)
    s = select_regions(b, [1, 2])
    for r in range(2):
        H.assert_same(O.port_summary(v, r, thr), O.port_summary(s, r, thr), "view %d" % r)
        H.assert_same(O.port_summary(v, r, thr), O.port_summary(b, r + 1, thr), "orig %d" % r)


def test_algorithmic_bytes_formula():
    b = H.kat_toy()
    assert b.algorithmic_bytes(1) == 2 * 480 + 4 * 6 + 32 * 6 + 80 + (33 * 26 * 2 + 32)
