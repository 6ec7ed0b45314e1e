"""The ingest library's own DEFLATE decoder (csrc/fast_inflate.h, exported as pv_inflate_raw) against zlib: every block
type and strategy, sizes around the BGZF limits, wrong sizes, corrupt and truncated streams (must be refused or decode to
something the CRC check of the BGZF reader rejects -- never crash). CPU only."""
import zlib

import numpy as np
import pytest

from pepper_thesis_b200 import ingest


def _deflate(data, level, strategy=zlib.Z_DEFAULT_STRATEGY):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    return c.compress(data) + c.flush()


def _inflate(comp, n):
    out = np.zeros(n + 1, np.uint8)
    rc = ingest.load().pv_inflate_raw(comp, len(comp), out.ctypes.data, n)
    return rc, bytes(out[:n])


@pytest.mark.parametrize("size", [0, 1, 2, 17, 255, 4096, 65280, 200000])
def test_matches_zlib(size):
    rng = np.random.default_rng(size)
    kinds = [bytes(size), bytes(rng.integers(0, 256, size, dtype=np.uint8)), bytes(rng.integers(5, 30, size, dtype=np.uint8)),
             (b"ACGTTGCA" * (size // 8 + 1))[:size], bytes(np.repeat(rng.integers(0, 256, size // 50 + 1, dtype=np.uint8), 50)[:size]),
             bytes((rng.integers(0, 4, size, dtype=np.uint8) * 17 + rng.integers(0, 2, size, dtype=np.uint8)).astype(np.uint8))]
    for data in kinds:
        for level, strategy in [(0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY),
                                (9, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)]:
            comp = _deflate(data, level, strategy)
            rc, got = _inflate(comp, len(data))
            assert rc == 0 and got == data, (level, strategy, size)
            if size > 1:
                assert _inflate(comp, len(data) - 1)[0] != 0          # exact size or nothing


def test_multi_block_streams_and_sync_flushes():
    rng = np.random.default_rng(1)
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    parts, comp = [], b""
    for i in range(6):
        p = bytes(rng.integers(0, 40, 3000 + 500 * i, dtype=np.uint8))
        parts.append(p)
        comp += c.compress(p) + c.flush(zlib.Z_SYNC_FLUSH if i % 2 else zlib.Z_FULL_FLUSH)   # empty stored blocks in between
    comp += c.flush()
    data = b"".join(parts)
    rc, got = _inflate(comp, len(data))
    assert rc == 0 and got == data


def test_corrupt_and_truncated_streams_are_survived():
    rng = np.random.default_rng(2)
    data = bytes(rng.integers(5, 30, 60000, dtype=np.uint8))
    comp = _deflate(data, 6)
    refused = 0
    for _ in range(300):
        b2 = bytearray(comp)
        pos = int(rng.integers(0, len(b2)))
        b2[pos] ^= 1 << int(rng.integers(0, 8))
        rc, got = _inflate(bytes(b2), len(data))
        refused += rc != 0
        assert rc != 0 or zlib.crc32(got) != zlib.crc32(data) or got == data
        if pos < len(comp) - 1:
            assert _inflate(bytes(comp[:pos]), len(data))[0] != 0
    assert refused > 20


def test_reader_uses_it_and_zlib_fallback_agrees(tmp_path):
    import bamio
    from pepper_thesis_b200 import synth
    lib = ingest.load()
    b = synth.generate("ont_r9", 200000, 8.0, seed=3, region_size=200000, margin=0)
    recs = []
    for i in range(b.n_reads):
        bo, n = int(b.read_base_off[i]), int(b.read_len[i]); co, k = int(b.read_cigar_off[i]), int(b.read_n_ops[i])
        recs.append(dict(tid=0, pos=int(b.read_pos[i]), mapq=60, flag=0, name="r%d" % i, cigar=[(int(c) & 15, int(c) >> 4) for c in b.cigar[co:co + k]],
                         seq=bytes(b.bases[bo:bo + n]).decode(), qual=bytes(b.quals[bo:bo + n]), tags=b""))
    recs.sort(key=lambda r: r["pos"])
    bam, fa = str(tmp_path / "t.bam"), str(tmp_path / "t.fa")
    bamio.write_bam(bam, [("chrS", 200000)], recs)
    bamio.write_fasta(fa, [("chrS", bytes(b.ref[:200000]).decode())])
    n0 = lib.pv_ingest_fast_blocks()
    got = ingest.ingest_regions(ingest.BAMHandler(bam), ingest.FASTAHandler(fa), "chrS", [0, 100000], [100000, 199999], threads=2).batch
    assert lib.pv_ingest_fast_blocks() > n0                          # the blocks went through the own decoder (CRC-confirmed)
    assert got.n_reads >= b.n_reads and int(got.read_len.sum()) > 0.9 * int(b.read_len.sum())
