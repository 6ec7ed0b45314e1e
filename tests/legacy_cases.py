"""Hand-built and fuzzed regions shared by the legacy-SummaryGenerator golden script and tests."""
import importlib.util
import os

import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ref_mod():
    d = os.path.join(ROOT, "oracle", "_ref")
    f = [x for x in os.listdir(d) if x.startswith("pv_ref_legacy")] if os.path.isdir(d) else []
    if not f:
        return None
    spec = importlib.util.spec_from_file_location("pv_ref_legacy", os.path.join(d, f[0]))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m


def hand_built():
    ref = "ACGTNacgt" + "ACGT" * 8
    reads = [H.Read(0, "ACGTACGTAC", [(0, 10)]),
             H.Read(2, "GTTTACG", [(0, 2), (1, 2), (0, 3)], rev=True),
             H.Read(2, "GTTTTACG", [(0, 2), (1, 3), (0, 3)]),
             H.Read(4, "ACAC", [(0, 2), (2, 3), (0, 2)]),
             H.Read(4, "ACAC", [(0, 2), (2, 3), (0, 2)], rev=True),
             H.Read(5, "NNAC", [(4, 2), (0, 2)]),
             H.Read(30, "ACGTACGTACGTACG", [(0, 15)]),
             H.Read(1, "CG", [(0, 2)], mapq=0)]                      # mapq 0: COUNTED by this generator
    return ref, reads


def cases():
    """name -> (batch, (chunk_size, chunk_overlap))"""
    ref, reads = hand_built()
    out = {"hand_built": (H.one_region(ref, reads), (16, 5))}
    for seed in (3, 8):
        out["fuzz%d" % seed] = (H.fuzz_region(seed, consistent=True), (64, 9))
    return out


def run_reference(m, b, chunk, r=0):
    ro, rl = int(b.region_ref_off[r]), int(b.region_ref_len[r])
    return m.legacy_summary(b.read_pos, b.read_base_off, b.read_len, b.read_cigar_off, b.read_n_ops, b.read_flags,
                            b.read_mapq, b.bases, b.quals, b.cigar, int(b.region_read_begin[r]),
                            int(b.region_read_begin[r + 1]), bytes(b.ref[ro:ro + rl]).decode(), int(b.region_ref_start[r]),
                            int(b.region_ref_end[r]), chunk[0], chunk[1])


def as_golden(d):
    return dict(image=d["image"].tolist(), genomic_pos=d["genomic_pos"].tolist(), ref_image=[int(x) for x in d["ref_image"]],
                longest_insert_count={str(k): int(v) for k, v in dict(d["longest_insert_count"]).items()},
                chunk_images=[[list(map(int, row)) for row in c] for c in d["chunk_images"]],
                chunk_positions=[[list(map(int, p)) for p in c] for c in d["chunk_positions"]],
                chunk_refs=[list(map(int, c)) for c in d["chunk_refs"]], chunk_labels=[list(map(int, c)) for c in d["chunk_labels"]],
                chunk_ids=[int(x) for x in d["chunk_ids"]])
