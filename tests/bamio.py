"""Test infrastructure: a small pure-Python BAM + BAI + FASTA/.fai WRITER and an independent BAM record READER
(gzip module), so the C++ ingest (libpv_ingest.so) is checked against files and a decoder it shares no code with.
Layouts: SAMv1 sections 4.1 (BGZF), 4.2 (BAM), 5.2 (BAI)."""
import gzip
import struct
import zlib

NT16 = "=ACMGRSVTWYHKDBN"
_NT16_CODE = {c: i for i, c in enumerate(NT16)}


def reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14: return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17: return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20: return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23: return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26: return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def ref_len_of(cigar):
    return sum(l for op, l in cigar if op in (0, 2, 3, 7, 8))


def encode_record(rec):
    """rec: dict(tid, pos, mapq, flag, name, cigar [(op, len)], seq str, qual bytes/list, tags bytes)."""
    name = rec["name"].encode() + b"\0"
    cigar = rec["cigar"]
    seq = rec["seq"]
    l_seq = len(seq)
    tags = rec.get("tags", b"")
    cig_words = [(l << 4) | op for op, l in cigar]
    if len(cig_words) > 65535:                       # SAMv1 4.2.2: real CIGAR in the CG tag
        tags = tags + b"CGBI" + struct.pack("<I", len(cig_words)) + struct.pack("<%dI" % len(cig_words), *cig_words)
        cig_words = [(l_seq << 4) | 4, (ref_len_of(cigar) << 4) | 3]
    end = rec["pos"] + max(1, ref_len_of(cigar))
    packed = bytearray((l_seq + 1) // 2)
    for i, c in enumerate(seq):
        packed[i >> 1] |= _NT16_CODE[c] << (0 if i & 1 else 4)
    qual = bytes(rec["qual"]) if l_seq else b""
    body = struct.pack("<iiBBHHHiiii", rec["tid"], rec["pos"], len(name), rec["mapq"], reg2bin(rec["pos"], end),
                       len(cig_words), rec["flag"], l_seq, -1, -1, 0)
    body += name + struct.pack("<%dI" % len(cig_words), *cig_words) + bytes(packed) + qual + tags
    return struct.pack("<i", len(body)) + body


def _bgzf_block(data, level=6):
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    cdata = co.compress(data) + co.flush()
    bsize = len(cdata) + 25
    hdr = struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize)
    return hdr + cdata + struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data))


BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


class BgzfWriter:
    def __init__(self, f, block=0xff00, level=6):
        self.f, self.block, self.level = f, block, level
        self.buf = bytearray()
        self.coff = 0

    def tell(self):
        return (self.coff << 16) | len(self.buf)

    def write(self, data):
        data = memoryview(data)
        while len(data):
            room = self.block - len(self.buf)
            self.buf += data[:room]
            data = data[room:]
            if len(self.buf) >= self.block:
                self.flush()

    def flush(self):
        if self.buf:
            blk = _bgzf_block(bytes(self.buf), self.level)
            self.f.write(blk)
            self.coff += len(blk)
            self.buf = bytearray()

    def close(self):
        self.flush()
        self.f.write(BGZF_EOF)


def write_bam(path, refs, records, header_text=None, block=0xff00):
    """refs: [(name, length)]; records sorted by (tid, pos). Writes path and path + '.bai'."""
    if header_text is None:
        header_text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % r for r in refs)
    index = [dict(bins={}, linear={}) for _ in refs]
    with open(path, "wb") as f:
        w = BgzfWriter(f, block)
        text = header_text.encode()
        w.write(b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(refs)))
        for name, ln in refs:
            nb = name.encode() + b"\0"
            w.write(struct.pack("<i", len(nb)) + nb + struct.pack("<i", ln))
        w.flush()
        for rec in records:
            beg_v = w.tell()
            w.write(encode_record(rec))
            end_v = w.tell()
            if rec["tid"] < 0:
                continue
            idx = index[rec["tid"]]
            end = rec["pos"] + max(1, ref_len_of(rec["cigar"]))
            b = reg2bin(rec["pos"], end)
            chunks = idx["bins"].setdefault(b, [])
            if chunks and chunks[-1][1] == beg_v:
                chunks[-1][1] = end_v
            else:
                chunks.append([beg_v, end_v])
            for win in range(rec["pos"] >> 14, ((end - 1) >> 14) + 1):
                if win not in idx["linear"]:
                    idx["linear"][win] = beg_v
        w.close()
    with open(path + ".bai", "wb") as f:
        f.write(b"BAI\1" + struct.pack("<i", len(refs)))
        for idx in index:
            f.write(struct.pack("<i", len(idx["bins"])))
            for b in sorted(idx["bins"]):
                f.write(struct.pack("<Ii", b, len(idx["bins"][b])))
                for beg_v, end_v in idx["bins"][b]:
                    f.write(struct.pack("<QQ", beg_v, end_v))
            n_intv = (max(idx["linear"]) + 1) if idx["linear"] else 0
            f.write(struct.pack("<i", n_intv))
            last = 0
            lin = []
            for win in range(n_intv):                      # htslib back-fills empty windows with the next offset;
                lin.append(idx["linear"].get(win))           # readers only need a lower bound, so carry the previous
            nxt = None
            for win in range(n_intv - 1, -1, -1):
                if lin[win] is None:
                    lin[win] = nxt if nxt is not None else 0
                else:
                    nxt = lin[win]
            for v in lin:
                f.write(struct.pack("<Q", v))


def write_fasta(path, seqs, width=60):
    """seqs: [(name, sequence str)]; writes path and path + '.fai'."""
    fai = []
    with open(path, "wb") as f:
        for name, s in seqs:
            f.write((">%s test sequence\n" % name).encode())
            off = f.tell()
            for i in range(0, len(s), width):
                f.write(s[i:i + width].encode() + b"\n")
            fai.append("%s\t%d\t%d\t%d\t%d\n" % (name, len(s), off, width, width + 1))
    with open(path + ".fai", "w") as f:
        f.write("".join(fai))


def read_bam(path):
    """Independent decoder: (header_text, refs, records) with records as dicts like encode_record takes (+ 'aux')."""
    with gzip.open(path, "rb") as g:
        data = g.read()
    assert data[:4] == b"BAM\1"
    l_text, = struct.unpack_from("<i", data, 4)
    text = data[8:8 + l_text].decode()
    p = 8 + l_text
    n_ref, = struct.unpack_from("<i", data, p); p += 4
    refs = []
    for _ in range(n_ref):
        l, = struct.unpack_from("<i", data, p); p += 4
        name = data[p:p + l - 1].decode(); p += l
        ln, = struct.unpack_from("<i", data, p); p += 4
        refs.append((name, ln))
    recs = []
    while p < len(data):
        bs, = struct.unpack_from("<i", data, p); p += 4
        r = data[p:p + bs]; p += bs
        tid, pos, l_name, mapq, _bin, n_cig, flag, l_seq, _nt, _np, _tl = struct.unpack_from("<iiBBHHHiiii", r, 0)
        o = 32
        name = r[o:o + l_name - 1].decode(); o += l_name
        cig = struct.unpack_from("<%dI" % n_cig, r, o); o += 4 * n_cig
        packed = r[o:o + (l_seq + 1) // 2]; o += (l_seq + 1) // 2
        qual = r[o:o + l_seq]; o += l_seq
        aux = r[o:]
        seq = "".join(NT16[(packed[i >> 1] >> (0 if i & 1 else 4)) & 15] for i in range(l_seq))
        cigar = [(c & 15, c >> 4) for c in cig]
        if n_cig == 2 and cigar[0] == (4, l_seq) and cigar[1][0] == 3:
            k = aux.find(b"CGBI")
            if k >= 0:
                n, = struct.unpack_from("<I", aux, k + 4)
                cigar = [(c & 15, c >> 4) for c in struct.unpack_from("<%dI" % n, aux, k + 8)]
        recs.append(dict(tid=tid, pos=pos, mapq=mapq, flag=flag, name=name, cigar=cigar, seq=seq, qual=bytes(qual), aux=bytes(aux)))
    return text, refs, recs
