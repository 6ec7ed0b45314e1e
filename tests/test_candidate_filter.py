"""Stage 3 (candidate filter): the device kernel + host tuple builder against the restatement of
CandidateFinder.py:391-600 in oracle/candidate_finder_port.py."""
import numpy as np
import pytest

import candidate_finder_port as CP
from pepper_thesis_b200 import candidate_filter as CF
from pepper_thesis_b200.pipeline import Predictions
from pepper_thesis_b200.read_batch import ReadBatch


def test_port_repeat_annotation():
    assert CP.repeat_annotation("AAAAACGT", 1) == [5, 5, 5, 5, 5, 1, 1, 1]
    assert CP.repeat_annotation("ACAAAT", 1) == [1, 1, 3, 3, 3, 1]
    assert CP.repeat_annotation("", 1) == []


def _world(seed, n_regions=3, L=700, K=600, lower=False):
    """Random reference (homopolymer-rich, some N / lower-case), regions with margins, random candidates and probs."""
    rng = np.random.RandomState(seed)
    contig_len = n_regions * L
    seq = []
    while len(seq) < contig_len:
        b = "ACGT"[rng.randint(4)] if rng.rand() > 0.01 else "N"
        seq.extend(b * int(rng.choice([1, 1, 1, 2, 3, 5, 6, 9])))
    contig = "".join(seq[:contig_len])
    if lower:
        contig = "".join(c.lower() if rng.rand() < 0.3 else c for c in contig)
    starts = [r * L for r in range(n_regions)]
    ends = [min(contig_len - 1, (r + 1) * L) for r in range(n_regions)]
    rs = [max(0, s - 100) for s in starts]
    re_ = [e + 100 for e in ends]
    refs, off = [], [0]
    for a, b in zip(rs, re_):
        piece = contig[a:b + 1]
        piece += "N" * (b + 1 - a - len(piece))                  # what the ingest pads past the contig end
        refs.append(piece); off.append(off[-1] + len(piece))
    z = np.zeros(0, np.int64)
    batch = ReadBatch(read_pos=z, read_base_off=z, read_len=np.zeros(0, np.int32), read_cigar_off=z, read_n_ops=np.zeros(0, np.int32),
                      read_flags=np.zeros(0, np.uint8), read_mapq=np.zeros(0, np.uint8), bases=np.zeros(0, np.uint8),
                      quals=np.zeros(0, np.uint8), cigar=np.zeros(0, np.uint32),
                      region_ref_start=np.array(rs, np.int64), region_ref_end=np.array(re_, np.int64),
                      region_cand_start=np.array(starts, np.int64), region_cand_end=np.array(ends, np.int64),
                      region_ref_off=np.array(off[:-1], np.int64), region_ref_len=np.array([len(x) for x in refs], np.int64),
                      region_read_begin=np.zeros(n_regions + 1, np.int64), ref=np.frombuffer("".join(refs).encode(), np.uint8).copy(),
                      contigs=["ctg"] * n_regions)
    region = np.sort(rng.randint(0, n_regions, K)).astype(np.int32)
    position = np.array([rng.randint(starts[r], ends[r] + 1) for r in region], np.int64)
    position[:4] = [0, 3, 9, 12][:min(4, K)]; region[:4] = 0                 # contig start: short downstream context
    position[-3:] = [contig_len - 1, contig_len - 4, contig_len - 11]; region[-3:] = n_regions - 1
    order = np.lexsort((position, region)); region, position = region[order], position[order]
    depth = rng.randint(3, 126, K).astype(np.int32)
    freq = np.minimum(depth, rng.randint(1, 126, K)).astype(np.int32)
    allele = np.zeros((K, 64), np.uint8); alen = np.zeros(K, np.uint8)
    strs = []
    for i in range(K):
        t = rng.choice([1, 1, 2, 3])
        n = 1 if t == 1 else rng.randint(2, 12)
        bases = "".join(rng.choice(list("ACGT") if rng.rand() > 0.05 else list("ACGTN")) for _ in range(n))
        s = str(t) + bases
        strs.append(s); allele[i, :len(s)] = np.frombuffer(s.encode(), np.uint8); alen[i] = len(s)
    raw = rng.rand(K, 3).astype(np.float32) ** 3
    raw[rng.rand(K) < 0.1] = [0.25, 0.25, 0.25]                              # ties: np.argmax takes the first
    probs = (raw / raw.sum(1, keepdims=True)).astype(np.float32)
    probs[rng.rand(K) < 0.05, 1] = np.float32(0.1)                          # exactly at a threshold (float32 0.1 > double 0.1)
    pred = Predictions(region, position, depth, freq, allele, alen, probs, probs.argmax(1).astype(np.uint8))
    fetch = lambda c, a, b: contig[max(0, a):max(0, b)].upper()             # FASTA_handler semantics (clips, upper-cases)
    cands = [("ctg", int(position[i]), int(depth[i]), [strs[i]], [int(freq[i])], probs[i]) for i in range(K)]
    return batch, pred, cands, fetch, contig_len


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
