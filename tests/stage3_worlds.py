"""Seeded inputs of the stage-3 tests (candidate filter, VCF writer): shared by the tests, by the golden-vector script
(tests/golden/make_stage3_golden.py) and by the live comparison with the unmodified reference modules."""
import numpy as np

from pepper_thesis_b200.pipeline import Predictions
from pepper_thesis_b200.read_batch import ReadBatch


def world(seed, n_regions=3, L=700, K=600, lower=False):
    """Random reference (homopolymer-rich, some N / lower-case), regions with margins, random candidates and probs.
    -> (batch, pred, cands, contig, contig_len); cands = (contig name, position, depth, [allele], [frequency], probs)."""
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
    cands = [("ctg", int(position[i]), int(depth[i]), [strs[i]], [int(freq[i])], probs[i]) for i in range(K)]
    return batch, pred, cands, contig, contig_len


def fetcher(contig):
    """FASTA_handler.get_reference_sequence semantics over one in-memory contig (clips at the ends, upper-cases)."""
    return lambda c, a, b: contig[max(0, a):max(0, b)].upper()


FILTER_CASES = [(0, (0.1, 0.1, 0.1, 0.15, 0.1, 0.1, 0.0, 0.0)), (1, (0.3, 0.5, 0.25, 0.4, 0.2, 0.6, 0.0, 0.0)),
                (2, (0.6, 0.7, 0.6, 0.7, 0.6, 0.7, 0.35, 0.25)),             # frequency rules in play (incl. the delete quirk)
                (3, (1.1, 1.1, 1.1, 1.1, 1.1, 1.1, 0.0, 0.0))]               # nothing passes by probability


def random_sites(seed, n_sites=300):
    """(contig, position) -> list of the 12-tuples find_candidates emits (random, incl. ties and saturating qualities)."""
    rng = np.random.default_rng(seed)
    alpha = "ACGT"
    sites = {}
    pos = 1000
    for _ in range(n_sites):
        pos += int(rng.integers(0, 40))                         # 0: two keys can share a start only across contigs
        contig = "chr%d" % (1 + int(rng.random() < 0.2))
        cands = []
        for _ in range(int(rng.integers(1, 7))):
            kind = rng.random()
            ref_base = alpha[int(rng.integers(0, 4))]
            if kind < 0.5:
                ref, alts = ref_base, [alpha[int(rng.integers(0, 4))]]
            elif kind < 0.75:
                ref, alts = ref_base, [ref_base + "".join(alpha[int(x)] for x in rng.integers(0, 4, int(rng.integers(1, 6))))]
            else:
                ref, alts = ref_base + "".join(alpha[int(x)] for x in rng.integers(0, 4, int(rng.integers(1, 6)))), [ref_base]
            probs = rng.dirichlet([0.6, 0.6, 0.6]).astype(np.float32)
            if rng.random() < 0.1:
                probs = np.array([0.25, 0.375, 0.375], np.float32)      # tie: first maximum wins
            if rng.random() < 0.05:
                probs = np.array([0.0, 0.0, 1.0], np.float32)           # 1 - p == 0: QUAL saturates at 90
            g = int(np.argmax(probs))
            gt = [[0, 0], [0, 1], [1, 1]][g]
            depth = int(rng.integers(1, 120))
            cands.append((contig, pos, pos + len(ref), ref, alts, gt, depth, [int(rng.integers(0, depth + 1))], probs[g], probs,
                          [max(probs[1], probs[2])], bool(rng.random() < 0.3)))
        sites[(contig, pos)] = cands
    return sites


VCF_CASES = [(11, (4, 20, 15, 20, 10)), (12, (2, 5, 30, 1, 12)), (13, (1, 90, 90, 90, 90))]


def plain(x):
    """numpy scalars / arrays / tuples -> plain Python (JSON-able, float32 values as exact doubles)."""
    if isinstance(x, dict):
        return {str(k): plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [plain(v) for v in x]
    if isinstance(x, np.ndarray):
        return [plain(v) for v in x.tolist()]
    if isinstance(x, (np.floating,)):
        return float(x)
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, (np.bool_,)):
        return bool(x)
    return x
