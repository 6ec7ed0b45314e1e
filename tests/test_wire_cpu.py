"""Host side of the reference-predicted bases wire form (pv_pack_bases_ref): the patch lists decode back to the reads with
a pure-Python restatement of the prediction rule documented in include/pepper_b200.h (no GPU)."""
import bisect

import numpy as np
import pytest

import helpers as H
from pepper_thesis_b200 import synth


def _predict(b, r, g):
    ro, rl = int(b.region_ref_off[g]), int(b.region_ref_len[g])
    rp = int(b.read_pos[r]) - int(b.region_ref_start[g])
    co, n, L = int(b.read_cigar_off[r]), int(b.read_n_ops[r]), int(b.read_len[r])
    out = np.full(L, ord("A"), np.uint8)
    ri = 0
    for k in range(n):
        w = int(b.cigar[co + k]); op, ln = w & 15, w >> 4
        m = op in (0, 7, 8)
        if m or op in (1, 4):
            for j in range(min(ln, max(0, L - ri))):
                p = rp + j
                if m and 0 <= p < rl:
                    out[ri + j] = b.ref[ro + p]
            ri += ln
        if m or op in (2, 3):
            rp += ln
    return out


def _decode(b, r, pred):
    cur = 0
    for e in b.bases_patch[int(b.read_patch_off[r]):int(b.read_patch_off[r + 1])]:
        s = int(e) & 255
        if s == 255:
            cur += 255
        else:
            pred[cur + s] = int(e) >> 8
            cur += s + 1
    return pred


@pytest.mark.parametrize("seed", range(6))
def test_patch_lists_decode_fuzz(seed):
    b = H.fuzz_region(seed)
    b.pack_bases_ref(threads=3)
    starts = list(b.region_read_begin)
    assert int(b.read_patch_off[-1]) == b.bases_patch.size
    for r in range(b.n_reads):
        g = bisect.bisect_right(starts, r) - 1
        o, L = int(b.read_base_off[r]), int(b.read_len[r])
        assert np.array_equal(_decode(b, r, _predict(b, r, g)), b.bases[o:o + L]), "read %d" % r


def test_patch_lists_decode_synthetic_and_views():
    b = synth.generate("ont_r9", 250000, 6.0, seed=3)
    rng = np.random.default_rng(0)
    pos = rng.integers(0, b.n_bases, 300)
    b.bases[pos] = rng.choice(np.frombuffer(b"NnacgtRY=\xff", np.uint8), 300)
    b.pack_bases_ref(threads=4)
    assert b.bases_patch.nbytes + b.read_patch_off.nbytes < b.n_bases // 8        # < 1 bit per base at ONT error rates
    starts = list(b.region_read_begin)
    for r in list(range(0, b.n_reads, 7)) + [b.n_reads - 1]:
        g = bisect.bisect_right(starts, r) - 1
        o, L = int(b.read_base_off[r]), int(b.read_len[r])
        assert np.array_equal(_decode(b, r, _predict(b, r, g)), b.bases[o:o + L]), "read %d" % r
    # a region view re-bases the patch offsets and slices the entries
    v = b.region_range_view(1, 3)
    assert int(v.read_patch_off[0]) == 0 and int(v.read_patch_off[-1]) == v.bases_patch.size
    vs = list(v.region_read_begin)
    for r in (0, v.n_reads // 2, v.n_reads - 1):
        g = bisect.bisect_right(vs, r) - 1
        o, L = int(v.read_base_off[r]), int(v.read_len[r])
        assert np.array_equal(_decode(v, r, _predict(v, r, g)), v.bases[o:o + L])
