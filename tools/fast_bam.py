"""Bench tool: writes a packed synthetic ReadBatch as BAM + BAI + FASTA quickly (numpy base packing, BGZF blocks compressed
on a thread pool). Same layouts as tests/bamio.py (SAMv1 4.1 / 4.2 / 5.2), which the tests use; this one exists so that a
multi-Mbp BAM can be produced inside a benchmark run."""
import os
import struct
import sys
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bamio  # noqa: E402

_CODE = np.zeros(256, np.uint8)
for _i, _c in enumerate(bamio.NT16):
    _CODE[ord(_c)] = _i


def _block(data, level):
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    c = co.compress(data) + co.flush()
    hdr = struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, len(c) + 25)
    return hdr + c + struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data))


def write_bam_from_batch(path, b, contig, contig_len, level=1, threads=0, block=0xff00):
    """b: ReadBatch of ONE whole-contig region (synth.generate(..., region_size=L, margin=0)); reads sorted by position."""
    order = np.argsort(b.read_pos, kind="stable")
    text = ("@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:%s\tLN:%d\n" % (contig, contig_len)).encode()
    nb = contig.encode() + b"\0"
    parts = [b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", 1) + struct.pack("<i", len(nb)) + nb + struct.pack("<i", contig_len)]
    hdr_len = len(parts[0])
    hdr_pad = (-hdr_len) % block                                  # the header gets its own blocks, like bamio.write_bam
    rec_u, rec_pos, rec_end = [], [], []
    u = 0
    for i in order:
        bo, n = int(b.read_base_off[i]), int(b.read_len[i]); co, k = int(b.read_cigar_off[i]), int(b.read_n_ops[i])
        cig = b.cigar[co:co + k]
        ops, lens = cig & 15, (cig >> 4).astype(np.int64)
        rlen = int(lens[(ops == 0) | (ops == 2) | (ops == 3) | (ops == 7) | (ops == 8)].sum())
        pos = int(b.read_pos[i]); end = pos + max(1, rlen)
        codes = _CODE[b.bases[bo:bo + n]]
        if n & 1:
            codes = np.concatenate([codes, np.zeros(1, np.uint8)])
        packed = ((codes[0::2] << 4) | codes[1::2]).tobytes()
        name = b"r%d\0" % i
        words = cig.astype("<u4").tobytes()
        tags = b""
        n_cig = k
        if k > 65535:
            tags = b"CGBI" + struct.pack("<I", k) + words
            words = struct.pack("<II", (n << 4) | 4, (rlen << 4) | 3); n_cig = 2
        body = struct.pack("<iiBBHHHiiii", 0, pos, len(name), int(b.read_mapq[i]), bamio.reg2bin(pos, end), n_cig,
                           0x10 if b.read_flags[i] & 1 else 0, n, -1, -1, 0) + name + words + packed + b.quals[bo:bo + n].tobytes() + tags
        rec = struct.pack("<i", len(body)) + body
        rec_u.append(u); rec_pos.append(pos); rec_end.append(end)
        parts.append(rec); u += len(rec)
    stream = b"".join(parts[1:])
    chunks = [parts[0][j:j + block] for j in range(0, hdr_len, block)] + [stream[j:j + block] for j in range(0, len(stream), block)]
    n_hdr_blocks = (hdr_len + block - 1) // block
    with ThreadPoolExecutor(threads or (os.cpu_count() or 1)) as ex:
        comp = list(ex.map(lambda d: _block(d, level), chunks, chunksize=64))
    coff = np.concatenate([[0], np.cumsum([len(c) for c in comp])]).astype(np.int64)
    with open(path, "wb") as f:
        for c in comp:
            f.write(c)
        f.write(bamio.BGZF_EOF)

    def voff(x):                                                  # offset in the record stream -> virtual offset
        blk = n_hdr_blocks + x // block
        if blk >= len(comp):
            return int(coff[len(comp)]) << 16
        return (int(coff[blk]) << 16) | (x % block)
    bins, linear = {}, {}
    for j in range(len(rec_u)):
        beg_v, end_v = voff(rec_u[j]), voff(rec_u[j] + (len(parts[1 + j])))
        ch = bins.setdefault(bamio.reg2bin(rec_pos[j], rec_end[j]), [])
        if ch and ch[-1][1] == beg_v:
            ch[-1][1] = end_v
        else:
            ch.append([beg_v, end_v])
        for win in range(rec_pos[j] >> 14, ((rec_end[j] - 1) >> 14) + 1):
            if win not in linear:
                linear[win] = beg_v
    with open(path + ".bai", "wb") as f:
        f.write(b"BAI\1" + struct.pack("<i", 1) + struct.pack("<i", len(bins)))
        for bn in sorted(bins):
            f.write(struct.pack("<Ii", bn, len(bins[bn])))
            for beg_v, end_v in bins[bn]:
                f.write(struct.pack("<QQ", beg_v, end_v))
        n_intv = (max(linear) + 1) if linear else 0
        f.write(struct.pack("<i", n_intv))
        lin = [linear.get(w) for w in range(n_intv)]
        nxt = None
        for w in range(n_intv - 1, -1, -1):
            if lin[w] is None:
                lin[w] = nxt if nxt is not None else 0
            else:
                nxt = lin[w]
        for v in lin:
            f.write(struct.pack("<Q", v))
    return dict(records=len(rec_u), inflated_bytes=hdr_len + len(stream), file_bytes=int(coff[-1]) + len(bamio.BGZF_EOF))
