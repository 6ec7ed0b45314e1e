"""The device's DEFLATE decoder (csrc/inflate_warp.cuh, one warp per BGZF block) through the C-ABI call
pv_bam_inflate_blocks, against zlib: raw DEFLATE streams of every block type (stored / fixed / dynamic), compression
levels and strategies, alphabets that need codes longer than the decoder's 10-bit table, self-overlapping matches of every
short period, unaligned payload offsets, empty and one-byte payloads, and corrupted streams (refused, never decoded into
wrong bytes silently: the CRC-32 of the trailer is part of the call)."""
import ctypes as C
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BLOCK_DT = np.dtype([("c_off", np.int64), ("c_len", np.int32), ("isize", np.int32), ("u_off", np.int64), ("crc", np.uint32), ("_pad", np.uint32)])


def deflate(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, mem=8):
    co = zlib.compressobj(level, zlib.DEFLATED, -15, mem, strategy)
    return co.compress(data) + co.flush()


def payloads():
    rng = np.random.default_rng(11)
    out = []

    def add(name, data, **kw):
        out.append((name, bytes(data), kw))
    add("empty", b"")
    add("one", b"A")
    add("two", b"AB", level=9)
    for lvl in (0, 1, 6, 9):
        add("random_l%d" % lvl, rng.integers(0, 256, 65280, dtype=np.uint8), level=lvl)
        add("qual_l%d" % lvl, rng.choice(np.arange(33, 75, dtype=np.uint8), 65280, p=np.r_[np.full(41, 0.5 / 41), 0.5]), level=lvl)
        add("bases_l%d" % lvl, rng.choice(np.array([0x11, 0x12, 0x14, 0x18, 0x21, 0x22, 0x24, 0x28, 0x41, 0x42, 0x44, 0x48, 0x81, 0x82, 0x84, 0x88], np.uint8), 60000), level=lvl)
    add("stored_multi", rng.integers(0, 256, 65536, dtype=np.uint8), level=0)
    add("fixed_text", (b"the quick brown fox jumps over the lazy dog " * 900)[:40000], strategy=zlib.Z_FIXED)
    add("fixed_random", rng.integers(0, 256, 5000, dtype=np.uint8), strategy=zlib.Z_FIXED)
    add("huffman_only", rng.choice(np.arange(0, 200, dtype=np.uint8), 50000, p=np.r_[np.full(199, 0.3 / 199), 0.7]), strategy=zlib.Z_HUFFMAN_ONLY)
    add("rle", np.repeat(rng.integers(0, 256, 700, dtype=np.uint8), rng.integers(1, 300, 700))[:65536], strategy=zlib.Z_RLE)
    add("zeros", np.zeros(65536, np.uint8), level=9)
    for period in list(range(1, 41)) + [63, 64, 65, 257, 258, 259, 1000]:
        pat = rng.integers(0, 256, period, dtype=np.uint8)
        add("period_%d" % period, np.tile(pat, 3000 // period + 3)[:3000 + period], level=9)
    # geometric alphabet over all 256 byte values: code lengths run past 10 bits (the canonical walk)
    p = 0.5 ** (np.arange(256) / 12.0); p /= p.sum()
    add("long_codes", rng.choice(np.arange(256, dtype=np.uint8), 65000, p=p), level=6)
    p = 0.5 ** (np.arange(256) / 4.0); p /= p.sum()
    add("very_long_codes", rng.choice(np.arange(256, dtype=np.uint8), 65000, p=p), level=6)
    # far matches (distance codes with many extra bits) and match-dense data
    blk = rng.integers(0, 256, 2000, dtype=np.uint8)
    add("far_matches", np.concatenate([blk, rng.integers(0, 256, 30000, dtype=np.uint8), blk, blk[:500], blk[100:1900]]), level=9)
    words = [bytes(rng.integers(97, 123, int(rng.integers(2, 12)), dtype=np.uint8)) for _ in range(300)]
    add("wordy", b" ".join(words[int(i)] for i in rng.integers(0, 300, 9000))[:65000], level=6)
    add("wordy_mem1", b" ".join(words[int(i)] for i in rng.integers(0, 300, 9000))[:65000], level=6, mem=1)   # small hash: many deflate blocks
    return out


def run_blocks(comp, table, total, verify=True):
    import torch
    from pepper_thesis_b200 import capi
    lib = capi.load()
    c = torch.from_numpy(np.frombuffer(bytes(comp) + b"\0" * 0, np.uint8).copy() if len(comp) else np.zeros(4, np.uint8)).cuda()
    t = torch.from_numpy(table.view(np.uint8).reshape(-1).copy()).cuda()
    U = torch.full((total + 64,), 0xEE, dtype=torch.uint8, device="cuda")
    bad = torch.zeros(2, dtype=torch.int32, device="cuda")        # bad blocks, the kernel's ticket
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    capi.check(lib.pv_bam_inflate_blocks(C.c_void_p(c.data_ptr()), len(comp), C.c_void_p(t.data_ptr()), len(table), C.c_void_p(U.data_ptr()), total,
                                         int(verify), C.c_void_p(bad.data_ptr()), st))
    torch.cuda.synchronize()
    return U.cpu().numpy(), int(bad[0].item())


def pack(items, gap_rng=None):
    comp, table, u = bytearray(), np.zeros(len(items), BLOCK_DT), 0
    for i, (name, data, kw) in enumerate(items):
        if gap_rng is not None:
            comp += bytes(int(gap_rng.integers(0, 4)))             # every payload alignment
        c = deflate(data, **kw)
        table[i] = (len(comp), len(c), len(data), u, zlib.crc32(data) & 0xffffffff, 0)
        comp += c
        u += len(data)
    return comp, table, u


def test_streams_match_zlib():
    items = payloads()
    comp, table, total = pack(items, np.random.default_rng(3))
    U, bad = run_blocks(comp, table, total)
    for (name, data, kw), row in zip(items, table):
        got = U[row["u_off"]:row["u_off"] + row["isize"]].tobytes()
        assert got == data, name
    assert bad == 0
    assert np.all(U[total:] == 0xEE)                              # nothing written behind the stream


def test_without_crc_and_tail_of_buffer():
    """The last payload ends with the buffer (no slack behind it), sizes that leave a partial last word."""
    rng = np.random.default_rng(5)
    for n in (1, 2, 3, 5, 31, 32, 33, 1000):
        data = bytes(rng.integers(65, 70, n, dtype=np.uint8))
        comp, table, total = pack([("x", data, dict(level=6))])
        U, bad = run_blocks(comp, table, total, verify=False)
        assert bad == 0 and U[:n].tobytes() == data


def test_corrupt_streams_are_refused():
    rng = np.random.default_rng(7)
    items = [(n, d, kw) for n, d, kw in payloads() if len(d) > 2000][:24]
    comp, table, total = pack(items)
    comp = bytearray(comp)
    want_bad = 0
    for (name, data, kw), row in zip(items, table):               # one flipped bit in every payload
        comp[int(row["c_off"]) + int(rng.integers(0, max(1, row["c_len"] - 1)))] ^= 1 << int(rng.integers(0, 8))
        try:                                                      # a flip can leave the bytes as they were (a distance inside a run)
            same = zlib.decompressobj(-15).decompress(bytes(comp[row["c_off"]:row["c_off"] + row["c_len"]])) == data
        except zlib.error:
            same = False
        want_bad += not same
    U, bad = run_blocks(comp, table, total)
    assert bad == want_bad and want_bad >= len(table) - 3
    # wrong sizes: one byte more / less than the stream holds
    comp, table, total = pack(items[:6])
    table["isize"][0::2] += 1
    table["isize"][1::2] -= 1
    table["u_off"] = np.arange(len(table)) * 70000
    U, bad = run_blocks(comp, table, 70000 * len(table), verify=False)
    assert bad == len(table)
    # truncated payloads
    comp, table, total = pack(items[:6])
    table["c_len"] //= 2
    U, bad = run_blocks(comp, table, total, verify=False)
    assert bad == len(table)


def test_many_blocks_random_mix():
    """A few thousand blocks of mixed content, as a BAM region would bring them."""
    rng = np.random.default_rng(13)
    items = []
    for i in range(1500):
        n = int(rng.integers(1, 65281))
        kind = i % 3
        if kind == 0:
            d = rng.choice(np.arange(33, 60, dtype=np.uint8), n)
        elif kind == 1:
            d = np.repeat(rng.integers(0, 256, n // 7 + 1, dtype=np.uint8), 7)[:n]
        else:
            d = np.concatenate([rng.integers(0, 16, n // 2, dtype=np.uint8) * 17, rng.choice(np.arange(35, 50, dtype=np.uint8), n - n // 2)])
        items.append((str(i), bytes(d), dict(level=int(rng.integers(1, 7)))))
    comp, table, total = pack(items, rng)
    U, bad = run_blocks(comp, table, total)
    assert bad == 0
    want = b"".join(d for _, d, _ in items)
    assert U[:total].tobytes() == want
