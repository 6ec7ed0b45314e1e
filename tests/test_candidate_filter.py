"""Stage 3 (candidate filter): the device kernel + host tuple builder against
  * golden vectors written by the UNMODIFIED reference module (CandidateFinder.find_candidates / small_chunk_stitch run
    through oracle/ref_stage3.py's h5py / PEPPER_VARIANT stand-ins; tests/golden/make_stage3_golden.py) -- parity pinned,
  * the restatement of CandidateFinder.py:391-600 in oracle/candidate_finder_port.py (itself checked against the same
    golden vectors and, where /root/reference exists, against the live reference)."""
import json
import os

import numpy as np
import pytest

import candidate_finder_port as CP
import stage3_worlds as W
from pepper_thesis_b200 import candidate_filter as CF

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _world(seed, lower=False, **kw):
    batch, pred, cands, contig, contig_len = W.world(seed, lower=lower, **kw)
    return batch, pred, cands, W.fetcher(contig), contig_len


def _golden(seed):
    with open(os.path.join(GOLDEN, "stage3_filter_seed%d.json" % seed)) as f:
        return json.load(f)


def _as_golden(result):
    contigs, phasing, variant = result
    return {"contigs": list(contigs), "phasing": [[list(k), W.plain(v)] for k, v in sorted(phasing.items())],
            "variant": [[list(k), W.plain(v)] for k, v in sorted(variant.items())]}


def _assert_equals_golden(result, g, what):
    got = _as_golden(result)
    assert got["contigs"] == g["contigs"], what
    for part in ("phasing", "variant"):
        assert [k for k, _ in got[part]] == [k for k, _ in g[part]], (what, part)
        for (k, a), (_, b) in zip(got[part], g[part]):
            assert a == b, (what, part, k, a, b)


@pytest.mark.parametrize("seed,opt", W.FILTER_CASES)
def test_port_matches_reference_golden(seed, opt):
    """The restatement against the golden vectors of the unmodified reference (CPU, everywhere)."""
    _, _, cands, fetch, _ = _world(seed, lower=(seed == 1))
    margin, deepv = CP.stitch(cands, fetch, CF.FilterOptions(*opt))
    _assert_equals_golden(CP.find_candidates(margin, deepv), _golden(seed), "port seed %d" % seed)


def test_live_reference_reproduces_golden():
    """Where the reference tree exists: the unmodified CandidateFinder.py, run now, gives the committed vectors."""
    import ref_stage3 as R
    if not R.available():
        pytest.skip("no /root/reference here")
    for seed, opt in W.FILTER_CASES:
        _, _, cands, contig, _ = W.world(seed, lower=(seed == 1))
        _assert_equals_golden(R.ref_find_candidates(cands, [("ctg", contig)], opt), _golden(seed), "live seed %d" % seed)
    # the per-batch function on its own, against the restatement's lists
    _, _, cands, contig, _ = W.world(2)
    m_ref, d_ref = R.ref_stitch(cands, [("ctg", contig)], W.FILTER_CASES[2][1])
    m, d = CP.stitch(cands, W.fetcher(contig), CF.FilterOptions(*W.FILTER_CASES[2][1]))
    assert W.plain(m_ref) == W.plain(m) and W.plain(d_ref) == W.plain(d) and len(d) > 50


@pytest.mark.gpu
@pytest.mark.parametrize("seed,opt", W.FILTER_CASES)
def test_filter_matches_reference_golden(seed, opt):
    """candidate_filter.cu + find_candidates against the unmodified reference's output on the same candidates."""
    batch, pred, _, _, contig_len = _world(seed, lower=(seed == 1))
    got = CF.find_candidates(pred, batch, CF.FilterOptions(*opt), contig_len=[contig_len] * batch.n_regions)
    g = _golden(seed)
    _assert_equals_golden(got, g, "device seed %d" % seed)
    if seed in (0, 1, 2):
        assert len(g["variant"]) > 20




def test_port_repeat_annotation():
    assert CP.repeat_annotation("AAAAACGT", 1) == [5, 5, 5, 5, 5, 1, 1, 1]
    assert CP.repeat_annotation("ACAAAT", 1) == [1, 1, 3, 3, 3, 1]
    assert CP.repeat_annotation("", 1) == []


def _same_tuple(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        if isinstance(x, np.ndarray) or isinstance(y, np.ndarray):
            assert np.array_equal(np.asarray(x), np.asarray(y))
        elif isinstance(x, (list, tuple)):
            assert [float(v) if isinstance(v, (np.floating, float)) else v for v in x] == \
                   [float(v) if isinstance(v, (np.floating, float)) else v for v in y]
        elif isinstance(x, (np.floating, float)):
            assert float(x) == float(y)
        else:
            assert x == y, (x, y)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,opts", [
    (0, CF.FilterOptions()),
    (1, CF.FilterOptions(0.3, 0.5, 0.25, 0.4, 0.2, 0.6, 0.0, 0.0)),
    (2, CF.FilterOptions(0.6, 0.7, 0.6, 0.7, 0.6, 0.7, 0.35, 0.25)),         # frequency rules in play (incl. the delete quirk)
    (3, CF.FilterOptions(1.1, 1.1, 1.1, 1.1, 1.1, 1.1, 0.0, 0.0)),           # nothing passes by probability
])
def test_filter_matches_port(seed, opts):
    batch, pred, cands, fetch, contig_len = _world(seed, lower=(seed == 1))
    got = CF.find_candidates(pred, batch, opts, contig_len=[contig_len] * batch.n_regions)
    margin, deepv = CP.stitch(cands, fetch, opts)
    want = CP.find_candidates(margin, deepv)
    assert got[0] == want[0]
    for gd, wd in ((got[1], want[1]), (got[2], want[2])):
        assert sorted(gd.keys()) == sorted(wd.keys())
        for k in wd:
            assert len(gd[k]) == len(wd[k]), k
            for a, b in zip(gd[k], wd[k]):
                _same_tuple(a, b)
    if seed in (0, 1, 2):
        assert len(want[2]) > 20
    if seed == 2:
        assert any(c[3] != c[3][:1] or True for v in want[2].values() for c in v)


@pytest.mark.gpu
def test_filter_flags_device_entry():
    """pv_candidate_filter with device pointers == the host-array entry point."""
    import ctypes as C
    import torch
    from pepper_thesis_b200 import capi
    batch, pred, _, _, contig_len = _world(5)
    opts = CF.FilterOptions()
    want = CF.filter_flags(pred, batch, opts, [contig_len] * batch.n_regions)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    t = [dev(pred.position), dev(pred.region), dev(pred.depth), dev(pred.frequency), dev(pred.allele), dev(pred.allele_len), dev(pred.probs),
         dev(batch.region_ref_start), dev(batch.region_ref_off), dev(batch.region_ref_len),
         dev(np.full(batch.n_regions, contig_len, np.int64)), dev(batch.ref)]
    flags = torch.zeros(len(pred), dtype=torch.uint8, device="cuda")
    o = opts.as_struct()
    capi.check(capi.load().pv_candidate_filter(len(pred), *[x.data_ptr() for x in t], C.byref(o), flags.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert np.array_equal(flags.cpu().numpy(), want)
    assert (want & CF.BAD_REF).any() and (want & CF.IN_REPEAT).any() and (want & CF.VARIANT).any()
