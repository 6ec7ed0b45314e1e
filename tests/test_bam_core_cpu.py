"""The __host__ __device__ core of the GPU ingest (pepper-thesis_b200/csrc/bam_core.cuh) run on the CPU through a small
harness (tests/native/bam_core_host.cpp, built here with g++): its DEFLATE decoder against zlib, its CRC-32 against
zlib's, and its record parsing + get_reads clipping against the CPU ingest (which tests/test_ingest.py pins to the compiled
reference bam_handler.cpp). The GPU kernels (ingest_gpu.cu) call exactly these record / clip functions, one thread or warp per
record; the device's DEFLATE decoder is the warp-cooperative one of inflate_warp.cuh (tests/test_inflate_gpu.py), of which
inflate_block here is the single-thread reference form."""
import ctypes as C
import gzip
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

import bamio
from pepper_thesis_b200 import ingest
import test_ingest as TI

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def core(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("bamcore") / "libbamcore_host.so")
    subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-I", os.path.join(ROOT, "pepper-thesis_b200", "csrc"),
                    os.path.join(HERE, "native", "bam_core_host.cpp"), "-o", so], check=True)
    lib = C.CDLL(so)
    lib.pvt_inflate.argtypes = [C.c_char_p, C.c_int64, C.c_void_p, C.c_int64]
    lib.pvt_crc32.argtypes = [C.c_char_p, C.c_int64]; lib.pvt_crc32.restype = C.c_uint32
    lib.pvt_get_reads.restype = C.c_int64
    lib.pvt_get_reads.argtypes = [C.c_char_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_int32] + [C.c_void_p] * 12
    return lib


def _deflate(data, level, strategy=zlib.Z_DEFAULT_STRATEGY):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    return c.compress(data) + c.flush()


@pytest.mark.parametrize("size", [0, 1, 2, 17, 255, 4096, 65280])
def test_inflate_matches_zlib(core, size):
    rng = np.random.default_rng(size)
    kinds = [bytes(size), bytes(rng.integers(0, 256, size, dtype=np.uint8)), bytes(rng.integers(5, 30, size, dtype=np.uint8)),
             (b"ACGTTGCA" * (size // 8 + 1))[:size], bytes(np.repeat(rng.integers(0, 256, size // 50 + 1, dtype=np.uint8), 50)[:size]),
             bytes((rng.integers(0, 4, size, dtype=np.uint8) * 17 + rng.integers(0, 2, size, dtype=np.uint8)).astype(np.uint8))]
    for data in kinds:
        for level, strategy in [(0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY),
                                (9, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)]:
            comp = _deflate(data, level, strategy)
            out = np.zeros(size + 1, np.uint8)
            assert core.pvt_inflate(comp, len(comp), out.ctypes.data, size) == 0, (level, strategy, size)
            assert bytes(out[:size]) == data
            if size > 1:
                assert core.pvt_inflate(comp, len(comp), out.ctypes.data, size - 1) != 0
        assert core.pvt_crc32(data, len(data)) == (zlib.crc32(data) & 0xffffffff)


def test_inflate_survives_corruption(core):
    rng = np.random.default_rng(2)
    data = bytes(rng.integers(5, 30, 60000, dtype=np.uint8))
    comp = _deflate(data, 6)
    out = np.zeros(len(data) + 1, np.uint8)
    for _ in range(200):
        b2 = bytearray(comp)
        for _k in range(int(rng.integers(1, 4))):
            b2[int(rng.integers(0, len(b2)))] ^= 1 << int(rng.integers(0, 8))
        cut = bytes(b2[:int(rng.integers(1, len(b2)))]) if rng.integers(0, 2) else bytes(b2)
        rc = core.pvt_inflate(cut, len(cut), out.ctypes.data, len(data))
        assert rc != 0 or zlib.crc32(bytes(out[:len(data)])) != zlib.crc32(data) or bytes(out[:len(data)]) == data


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("bamcore_files")
    ref, recs = TI._records_from_synth()
    bam = str(d / "t.bam")
    bamio.write_bam(bam, [("chrS", TI.CONTIG_LEN), ("chrT", 5000)], recs, header_text="@HD\tVN:1.6\n", block=0x8000)
    U = gzip.decompress(open(bam, "rb").read())
    l_text, = struct.unpack_from("<I", U, 4)
    o = 8 + l_text
    n_ref, = struct.unpack_from("<I", U, o); o += 4
    for _ in range(n_ref):
        l, = struct.unpack_from("<I", U, o); o += 4 + l + 4
    return dict(bam=bam, U=U, first=o)


@pytest.mark.parametrize("span", TI.SPANS)
@pytest.mark.parametrize("supp,min_mapq", [(False, 0), (True, 10)])
def test_clip_matches_cpu_ingest(core, files, span, supp, min_mapq):
    U = files["U"]
    cap_r, cap_b, cap_o = 4096, 4 << 20, 1 << 20
    a = dict(pos=np.zeros(cap_r, np.int64), pos_end=np.zeros(cap_r, np.int64), len=np.zeros(cap_r, np.int32), n_ops=np.zeros(cap_r, np.int32),
             hp=np.zeros(cap_r, np.int32), rev=np.zeros(cap_r, np.uint8), mapq=np.zeros(cap_r, np.uint8), bases=np.zeros(cap_b, np.uint8),
             quals=np.zeros(cap_b, np.uint8), cigar=np.zeros(cap_o, np.uint32), nb=np.zeros(1, np.int64), no=np.zeros(1, np.int64))
    n = core.pvt_get_reads(U, len(U), files["first"], None, 0, 0, span[0], span[1], int(supp), min_mapq,
                           *[a[k].ctypes.data for k in ("pos", "pos_end", "len", "n_ops", "hp", "rev", "mapq", "bases", "quals", "cigar", "nb", "no")])
    want = ingest.BAMHandler(files["bam"]).get_reads_packed("chrS", span[0], span[1], supp, min_mapq, 1)
    b = want.batch
    assert n == b.n_reads
    bo = co = 0
    for i in range(n):
        L, K = int(b.read_len[i]), int(b.read_n_ops[i])
        assert (int(a["pos"][i]), int(a["pos_end"][i]), int(a["len"][i]), int(a["n_ops"][i])) == (int(b.read_pos[i]), int(want.pos_end[i]), L, K)
        assert int(a["hp"][i]) == int(want.hp_tag[i]) and int(a["rev"][i]) == int(b.read_flags[i] & 1) and int(a["mapq"][i]) == int(b.read_mapq[i])
        wb = int(b.read_base_off[i]); wc = int(b.read_cigar_off[i])
        assert np.array_equal(a["bases"][bo:bo + L], b.bases[wb:wb + L]) and np.array_equal(a["quals"][bo:bo + L], b.quals[wb:wb + L])
        assert np.array_equal(a["cigar"][co:co + K], b.cigar[wc:wc + K])
        bo += L; co += K


@pytest.mark.parametrize("span", [(0, TI.CONTIG_LEN), (39000, 47000), (61234, 71234), (0, 5000), (TI.CONTIG_LEN - 3000, TI.CONTIG_LEN + 100)])
def test_host_plan_blocks_and_chain_segments(core, files, span):
    """pv_bam_plan*: the planned BGZF blocks inflate (with the device decoder, run here on the CPU) to a stream in which every
    chain segment starts at a record and chases exactly to its end, and the records found that way cut to the same reads as
    the CPU ingest."""
    ilib = ingest.load()
    bam = ingest.BAMHandler(files["bam"])
    plan = C.c_void_p()
    ingest._check(ilib.pv_bam_plan(bam._h, b"chrS", span[0], span[1], C.byref(plan)))
    nbytes = ilib.pv_bam_plan_comp_bytes(plan)
    comp = np.zeros(nbytes, np.uint8)
    ingest._check(ilib.pv_bam_plan_load(plan, comp.ctypes.data, 3))
    nb, ns, ub = ilib.pv_bam_plan_n_blocks(plan), ilib.pv_bam_plan_n_segments(plan), ilib.pv_bam_plan_inflated_bytes(plan)
    from pepper_thesis_b200.ingest_gpu import _BLOCK_DT
    blocks = np.zeros(nb, _BLOCK_DT)
    seg = np.zeros((2, ns), np.int64)
    ingest._check(ilib.pv_bam_plan_tables(plan, blocks.ctypes.data, seg[0].ctypes.data, seg[1].ctypes.data))
    ilib.pv_bam_plan_free(plan)
    assert nb > 0 and ns > 0 and int(blocks["isize"].sum()) == ub
    U = np.zeros(ub + 1, np.uint8)
    for b in blocks:
        payload = comp[int(b["c_off"]):int(b["c_off"]) + int(b["c_len"])].tobytes()
        assert core.pvt_inflate(payload, len(payload), U[int(b["u_off"]):].ctypes.data, int(b["isize"])) == 0
        assert core.pvt_crc32(U[int(b["u_off"]):int(b["u_off"]) + int(b["isize"])].tobytes(), int(b["isize"])) == int(b["crc"])
    Ub = U[:ub].tobytes()
    recs = []
    for a, z in zip(seg[0], seg[1]):
        off = int(a)
        while off < z:
            bs, = struct.unpack_from("<I", Ub, off)
            assert bs >= 32
            recs.append(off)
            off += 4 + bs
        assert off == z
    recs = np.asarray(recs, np.int64)
    # the records of the plan, cut by the core, are the CPU ingest's reads of the span
    cap_r, cap_b, cap_o = 4096, 4 << 20, 1 << 20
    a = [np.zeros(cap_r, np.int64), np.zeros(cap_r, np.int64), np.zeros(cap_r, np.int32), np.zeros(cap_r, np.int32), np.zeros(cap_r, np.int32),
         np.zeros(cap_r, np.uint8), np.zeros(cap_r, np.uint8), np.zeros(cap_b, np.uint8), np.zeros(cap_b, np.uint8), np.zeros(cap_o, np.uint32),
         np.zeros(1, np.int64), np.zeros(1, np.int64)]
    n = core.pvt_get_reads(Ub, len(Ub), 0, recs.ctypes.data, len(recs), 0, span[0], span[1], 0, 0, *[x.ctypes.data for x in a])
    want = bam.get_reads_packed("chrS", span[0], span[1], False, 0, 1)
    assert n == want.batch.n_reads and np.array_equal(a[0][:n], want.batch.read_pos)


def test_record_helpers_under_sanitizers(tmp_path):
    """A valid record stream corrupted a few bytes at a time (length fields above all) through the chain / parse / clip / base
    reads the device kernels perform, in a heap buffer of exactly the stream's size, built with AddressSanitizer +
    UndefinedBehaviorSanitizer: no read leaves the stream, whatever the fields claim."""
    rng = np.random.default_rng(8)
    recs = []
    for i in range(40):
        n = int(rng.integers(50, 3000))
        seq = "".join("ACGTN"[int(x)] for x in rng.choice(5, n, p=[0.24, 0.24, 0.24, 0.24, 0.04]))
        cig, left = [], n
        if rng.random() < 0.3:
            s = int(rng.integers(1, 20)); cig.append((4, s)); left -= s
        while left > 0:
            op = int(rng.choice([0, 1, 2, 3, 7, 8], p=[0.5, 0.15, 0.15, 0.05, 0.1, 0.05]))
            l = int(rng.integers(1, 60))
            if op in (0, 1, 7, 8):
                l = min(l, left); left -= l
            cig.append((op, l))
        tags = b"HPi" + struct.pack("<i", int(rng.integers(0, 3))) if i % 3 == 0 else (b"XZZ" + b"abc\0" + b"HPC\x01" if i % 3 == 1 else b"")
        recs.append(dict(tid=0, pos=100 + 150 * i, mapq=int(rng.integers(0, 61)), flag=int(rng.choice([0, 16, 2048])), name="read%d" % i,
                         cigar=cig, seq=seq, qual=list(rng.integers(0, 60, n)), tags=tags))
    stream = b"".join(bamio.encode_record(r) for r in recs)
    p = tmp_path / "stream.bin"
    p.write_bytes(stream)
    exe = str(tmp_path / "bam_core_fuzz")
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-I",
                    os.path.join(ROOT, "pepper-thesis_b200", "csrc"), os.path.join(HERE, "native", "bam_core_fuzz.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe, str(p), "0", "30000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    assert "reads cut" in r.stdout
