"""Shared test helpers: duck-typed reads, known-answer cases, fuzz generator, comparisons."""
import numpy as np

from pepper_thesis_b200.read_batch import Region, pack_regions
from pepper_thesis_b200.synth import Thresholds

R9 = Thresholds(1, 1, 0.10, 0.15, 0.15, 3, 0.10, 0.10, 2, False)


class Flags:
    def __init__(self, rev):
        self.is_reverse = rev


class Read:
    """Duck type of the reference's type_read (pybind_api.h:208-221)."""

    def __init__(self, pos, seq, cigar, rev=False, q=30, mapq=60):
        self.pos = pos
        self.sequence = seq
        self.base_qualities = [q] * len(seq) if isinstance(q, int) else list(q)
        self.cigar_tuples = list(cigar)
        self.flags = Flags(rev)
        self.mapping_quality = mapq


def one_region(ref, reads, ref_start=0, ref_end=None, cand=None, contig="c"):
    ref_end = ref_start + len(ref) - 1 if ref_end is None else ref_end
    cs, ce = cand if cand else (ref_start, ref_end)
    return pack_regions([Region(contig, ref_start, ref_end, ref, cs, ce, reads)])


# ---- SURVEY.md section 8a known-answer cases -------------------------------------------------------------------
def kat_toy():
    ref = "ACGT" * 20
    reads = []
    for i in range(6):
        s = list(ref[:80])
        if i < 3:
            s[40] = "T"
        reads.append(Read(0, "".join(s), [(0, 80)], rev=(i % 2 == 1)))
    return one_region(ref, reads)


def kat_clamp():
    ref = "ACGT" * 25
    reads = []
    for i in range(300):
        s = list(ref)
        if i < 100:
            s[50] = "A"
        elif i < 200:
            s[50] = "T"
        reads.append(Read(0, "".join(s), [(0, 100)]))
    return one_region(ref, reads)


def kat_del():
    ref = "ACGT" * 25
    reads = []
    for i in range(8):
        if i % 2 == 0:
            reads.append(Read(0, ref[:51] + ref[54:], [(0, 51), (2, 3), (0, 46)], rev=(i % 4 == 2)))
        else:
            reads.append(Read(0, ref, [(0, 100)], rev=(i % 4 == 3)))
    return one_region(ref, reads)


def kat_ins():
    ref = "ACGT" * 25
    reads = []
    for i in range(6):
        if i < 3:
            reads.append(Read(0, ref[:51] + "TT" + ref[51:], [(0, 51), (1, 2), (0, 49)], rev=(i % 2 == 1)))
        else:
            reads.append(Read(0, ref, [(0, 100)], rev=(i % 2 == 1)))
    return one_region(ref, reads)


def kat_refskip():
    """REF_SKIP falls through into SOFT_CLIP (region_summary.cpp:556-561): read_index also advances by 10.
    The read carries 10 spare bases (trailing soft clip) so the reference never indexes past its sequence."""
    ref = "ACGT" * 25
    reads = [Read(0, ref[:40] + ref[50:] + "ACGTACGTAC", [(0, 40), (3, 10), (0, 50), (4, 10)]) for _ in range(6)]
    return one_region(ref, reads)


KATS = {"toy": kat_toy, "clamp": kat_clamp, "del": kat_del, "ins": kat_ins, "refskip": kat_refskip}


# ---- fuzz ---------------------------------------------------------------------------------------------------------
def fuzz_region(seed, L=400, n_reads=40, weird=True, ref_n=False, ref_start=1000, consistent=False):
    """Random reads exercising every CIGAR op, odd bytes, low qualities, mapq 0, reads crossing the region ends.
    The reference's undefined behaviours are avoided (read index stays inside the sequence under the
    reference's own stepping, reads end in a match, no candidate on a non-ACGT reference base unless ref_n).
    ``consistent``: N / P ops consume no read bases (SAM semantics, what the polisher's generator assumes) instead of the
    variant generator's fall-through stepping."""
    rng = np.random.default_rng(seed)
    alpha = np.frombuffer(b"ACGT", np.uint8)
    ref = alpha[rng.integers(0, 4, L + 80)].copy()
    if ref_n:
        for _ in range(3):
            ref[rng.integers(0, L)] = ord("N")
        ref[rng.integers(0, L)] = ord("a")
    ref_b = ref.tobytes()
    region_ref = ref_b[40:40 + L]          # region occupies [ref_start, ref_start+L-1]; contig has 40 bp around it
    reads = []
    hot = rng.integers(20, L - 20, 6)      # positions where many reads agree on a variant
    for _ in range(n_reads):
        start = int(rng.integers(-30, L - 10))        # region-relative start (may be left of the region)
        ops, seq, quals = [], bytearray(), []
        x = start
        target = int(rng.integers(30, L + 60))
        first = True
        while x < start + target and x < L + 35:
            r = rng.random()
            if first or r < 0.62:
                op = int(rng.choice([0, 0, 0, 7, 8])) if weird else 0
                ln = int(rng.integers(1, 40))
                for i in range(ln):
                    c = ref[40 + x + i] if 0 <= 40 + x + i < len(ref) else ord("A")
                    if rng.random() < 0.06:
                        c = int(alpha[rng.integers(0, 4)])
                    for h in hot:
                        if x + i == h and rng.random() < 0.5:
                            c = int(alpha[(int(np.searchsorted(alpha, ref[40 + h])) + 1 + h % 3) % 4])
                    if weird and rng.random() < 0.02:
                        c = int(rng.choice(np.frombuffer(b"NacgtIDMR=*", np.uint8)))
                    seq.append(c); quals.append(int(rng.choice([0, 1, 2, 7, 20, 30, 40])) if rng.random() < 0.3 else 30)
                x += ln
            elif r < 0.75:
                op = 1; ln = int(rng.choice([1, 1, 2, 3, 5, 58, 59, 60, 61]) if rng.random() < 0.9 else rng.integers(1, 80))
                hotins = any(x - 1 == h for h in hot)
                for i in range(ln):
                    seq.append(int(alpha[(i + (0 if hotins else rng.integers(0, 4))) % 4]))
                    quals.append(int(rng.choice([0, 1, 5, 30])))
            elif r < 0.88:
                op = 2; ln = int(rng.choice([1, 1, 2, 3, 5, 58, 59, 60, 61]) if rng.random() < 0.9 else rng.integers(1, 80))
                x += ln
            elif weird and r < 0.92:
                op = 4; ln = int(rng.integers(1, 6))
                for i in range(ln):
                    seq.append(ord("A")); quals.append(30)
            elif weird and r < 0.95:
                op = int(rng.choice([3, 6])); ln = int(rng.integers(1, 8))
                x += ln
                for i in range(0 if consistent else ln):  # the variant reference also advances the read here
                    seq.append(ord("C")); quals.append(30)
            elif weird and r < 0.97:
                op = 5; ln = int(rng.integers(1, 5))
            else:
                continue
            first = False
            if ops and ops[-1][0] == op and op in (0, 7, 8):
                ops[-1] = (op, ops[-1][1] + ln)
            else:
                ops.append((op, ln))
        # end in a match so the reference never reads base_qualities[read_index] past the end (:506)
        ops.append((0, 3))
        for i in range(3):
            c = ref[40 + x + i] if 0 <= 40 + x + i < len(ref) else ord("A")
            seq.append(int(c)); quals.append(30)
        mapq = 0 if rng.random() < 0.05 else 60
        reads.append(Read(ref_start + start, bytes(seq).decode("latin-1"), ops, rev=bool(rng.integers(0, 2)), q=quals, mapq=mapq))
    cand = (ref_start + int(rng.integers(0, 30)), ref_start + L - 1 - int(rng.integers(0, 30)))
    return pack_regions([Region("fz", ref_start, ref_start + L - 1, region_ref, cand[0], cand[1], reads)])


def fuzz_thresholds(seed):
    rng = np.random.default_rng(seed + 7919)
    return Thresholds(float(rng.choice([0, 1, 2, 7.5, 10])), float(rng.choice([0, 1, 5.5, 10])),
                      float(rng.choice([0.05, 0.10, 0.2])), float(rng.choice([0.05, 0.15])),
                      float(rng.choice([0.05, 0.15])), float(rng.choice([1, 3])),
                      float(rng.choice([0.0, 0.10])), float(rng.choice([0.0, 0.10])), float(rng.choice([1, 2])),
                      bool(rng.random() < 0.15))


def assert_same(a, b, what=""):
    """a, b: dicts with position/depth/frequency/alleles/images."""
    assert len(a["position"]) == len(b["position"]), "%s: %d vs %d candidates" % (what, len(a["position"]), len(b["position"]))
    assert np.array_equal(np.asarray(a["position"]), np.asarray(b["position"])), what + ": positions"
    assert list(a["alleles"]) == list(b["alleles"]), what + ": alleles"
    assert np.array_equal(np.asarray(a["depth"]), np.asarray(b["depth"])), what + ": depth"
    assert np.array_equal(np.asarray(a["frequency"]), np.asarray(b["frequency"])), what + ": frequency"
    ia, ib = np.asarray(a["images"]).astype(np.int32), np.asarray(b["images"]).astype(np.int32)
    if ia.size or ib.size:
        bad = np.argwhere(ia != ib)
        assert bad.size == 0, "%s: windows differ first at %s: %s vs %s" % (
            what, bad[0].tolist(), ia[tuple(bad[0])], ib[tuple(bad[0])])
