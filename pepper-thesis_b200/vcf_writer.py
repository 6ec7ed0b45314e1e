"""Stage-3 VCF output without pysam: the reference's ``VCFWriter``
(/root/reference/pepper_variant/modules/python/VcfWriter.py:12-289) over the candidate tuples
:func:`candidate_filter.find_candidates` returns.

Same constructor and ``write_vcf_records`` contract as the reference (five files: full, pepper, variant-calling and its
SNP / INDEL split; the five counts come back), same per-site merge (multi-allelic sites, reference-allele
normalisation, genotype / quality rules, q cut-offs). What differs:
  * the files are written by this module as BGZF-compressed VCF 4.2 text (``.vcf.gz``, readable by bgzip / bcftools /
    pysam); pysam and htslib are not needed;
  * no tabix index is written (the reference calls ``pysam.tabix_index`` on close, :41-45) -- run ``tabix -p vcf`` on the
    files where an index is needed;
  * floats are printed with ``%g`` (htslib prints 6 significant digits as well, but byte-level identity with pysam's
    output is not pinned: pysam is absent from this image, see oracle/vcf_writer_port.py).
Host-side Python on purpose: the reference's writer is Python, a site costs microseconds, and nothing here touches the GPU.
"""
from __future__ import annotations

import math
import os
import struct
import zlib
from dataclasses import dataclass
from typing import Dict, Iterable, List, Sequence, Tuple


@dataclass
class VcfOptions:
    """The option fields ``write_vcf_records`` reads; defaults = ONT R9 Guppy5 SUP preset (SetParameters.py:40-65)."""
    allowed_multiallelics: int = 4
    snp_q_cutoff: float = 20
    indel_q_cutoff: float = 15
    snp_q_cutoff_in_lc: float = 20
    indel_q_cutoff_in_lc: float = 10


class BgzfTextWriter:
    """Minimal BGZF writer (SAM spec 4.1): gzip members of <= 64 KiB with the ``BC`` extra field + the EOF marker."""
    BLOCK = 0xff00

    def __init__(self, path: str, level: int = 6):
        self._f = open(path, "wb")
        self._buf = bytearray()
        self._level = level
        self.upos = 0                                            # uncompressed bytes accepted so far
        self.block_coff = []                                     # compressed file offset of every data block written

    def write(self, text: str):
        data = text.encode()
        self._buf += data
        self.upos += len(data)
        while len(self._buf) >= self.BLOCK:
            self._block(bytes(self._buf[:self.BLOCK]))
            del self._buf[:self.BLOCK]

    def _block(self, data: bytes):
        c = zlib.compressobj(self._level, zlib.DEFLATED, -15)
        body = c.compress(data) + c.flush()
        bsize = len(body) + 25                                   # total block size - 1
        if data:
            self.block_coff.append(self._f.tell())
        self._f.write(struct.pack("<4BI2BH2BHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize))
        self._f.write(body)
        self._f.write(struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data)))

    def virtual_offset(self, upos: int) -> int:
        """BGZF virtual file offset (compressed block start << 16 | offset inside the block) of uncompressed byte ``upos``;
        valid once the file is closed. Data blocks are cut every BLOCK bytes of text."""
        if not self.block_coff:
            return 0
        blk = min(upos // self.BLOCK, len(self.block_coff) - 1)  # the position right behind the file's last byte stays in its block
        return (self.block_coff[blk] << 16) | (upos - blk * self.BLOCK)

    def close(self):
        if self._f is None:
            return
        if self._buf:
            self._block(bytes(self._buf))
            self._buf.clear()
        self._block(b"")                                         # EOF marker: an empty block
        self._f.close()
        self._f = None


def _reg2bin(beg: int, end: int) -> int:
    """UCSC binning scheme of tabix / BAI (SAMv1 5.3): 0-based half-open [beg, end), 16 kbp smallest bins, 5 levels."""
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def write_tabix_index(path: str, names: Sequence[str], records: Sequence[tuple]):
    """``<vcf.gz>.tbi`` as ``pysam.tabix_index(..., preset="vcf")`` leaves it beside every output (VcfWriter.py:41-45): the
    tabix format of the htslib specification -- BGZF-compressed; header (format 2 = VCF, columns 1/2/0, meta '#'), the
    sequence names, per sequence the bins with their chunks of virtual offsets and the 16 kbp linear index.
    ``records``: (sequence index, 0-based begin, end, virtual offset of the line, virtual offset behind it), in file order."""
    n_ref = len(names)
    bins = [dict() for _ in range(n_ref)]
    lin = [[] for _ in range(n_ref)]
    span = [[None, None, 0] for _ in range(n_ref)]             # first / last virtual offset, record count (the meta pseudo-bin)
    for tid, beg, end, v0, v1 in records:
        end = max(end, beg + 1)
        chunks = bins[tid].setdefault(_reg2bin(beg, end), [])
        if chunks and chunks[-1][1] == v0:
            chunks[-1][1] = v1                                   # adjacent records of one bin share a chunk
        else:
            chunks.append([v0, v1])
        w0, w1 = beg >> 14, (end - 1) >> 14
        if len(lin[tid]) <= w1:
            lin[tid].extend([0] * (w1 + 1 - len(lin[tid])))
        for w in range(w0, w1 + 1):
            if lin[tid][w] == 0:
                lin[tid][w] = v0
        span[tid][0] = v0 if span[tid][0] is None else span[tid][0]
        span[tid][1] = v1
        span[tid][2] += 1
    out = bytearray()
    nm = b"".join(n.encode() + b"\0" for n in names)
    out += struct.pack("<4s8i", b"TBI\1", n_ref, 2, 1, 2, 0, ord("#"), 0, len(nm)) + nm
    for tid in range(n_ref):
        b = bins[tid]
        out += struct.pack("<i", len(b) + (1 if span[tid][2] else 0))
        for k in sorted(b):
            out += struct.pack("<Ii", k, len(b[k]))
            for v0, v1 in b[k]:
                out += struct.pack("<QQ", v0, v1)
        if span[tid][2]:                                         # htslib's pseudo-bin 37450: file span and record counts
            out += struct.pack("<Ii4Q", 37450, 2, span[tid][0], span[tid][1], span[tid][2], 0)
        last = 0
        filled = []
        for v in lin[tid]:                                       # windows without a record carry the offset of the window before
            last = v if v else last
            filled.append(v if v else last)
        out += struct.pack("<i", len(filled)) + b"".join(struct.pack("<Q", v) for v in filled)
    w = BgzfTextWriter(path + ".tbi")
    w._buf += bytes(out)
    w.upos += len(out)
    while len(w._buf) >= w.BLOCK:
        w._block(bytes(w._buf[:w.BLOCK]))
        del w._buf[:w.BLOCK]
    w.close()


def _phred(p_correct: float) -> int:
    return max(1, int(-10 * math.log10(max(0.000000001, 1.0 - p_correct))))          # VcfWriter.py:157


def merge_site(candidates: Sequence[tuple], allowed_multiallelics: int):
    """The selected candidates of one position -> one site (``candidate_list_to_variant``, VcfWriter.py:48-139).

    Returns ``(contig, start, end, ref, alts, gt, depth, supports, gt_qual, non_alt_predictions, in_repeat)``.
    Candidates are ranked by (genotype, genotype probability) descending and cut to ``allowed_multiallelics``; every
    allele is extended to the longest reference allele of the site; the site's genotype lists the (1-based) alleles
    predicted het once and hom-alt twice and is kept only when that makes one or two entries."""
    ranked = sorted(candidates, key=lambda c: (c[5], c[8]), reverse=True)
    if len(ranked) > allowed_multiallelics:
        ranked = ranked[:allowed_multiallelics]
    longest = ""
    for c in ranked:
        if len(c[3]) > len(longest):
            longest = c[3]
    alts, supports, non_alt, called = [], [], [], []
    gt_qual, depth, in_repeat = -1.0, None, False
    head = None
    for k, c in enumerate(ranked, start=1):
        ref_allele = c[3]
        tail = longest[len(ref_allele):] if len(ref_allele) < len(longest) else ""   # == longest[-missing:]
        probs = c[9]
        best = max(range(len(probs)), key=lambda j: (probs[j], -j))                  # first maximum, like np.argmax
        if best != 0:
            gt_qual = probs[best] if gt_qual < 0 else min(gt_qual, probs[best])
        elif gt_qual < 0:
            gt_qual = max(probs[1], probs[2])
        if head is None:
            head = (c[0], c[1], c[1] + len(ref_allele + tail), ref_allele + tail)
            depth = c[6]
        depth = min(depth, c[6])
        in_repeat = in_repeat or bool(c[11])
        alts.append(c[4][0] + tail)
        supports.append(c[7][0])
        non_alt.extend(c[10])
        called += [k] * (1 if best == 1 else 2 if best == 2 else 0)
    # het alleles first, then the second copies of the hom-alt ones (genotype_hp1 + genotype_hp2, :122-131)
    first, second = [], []
    seen = set()
    for k in called:
        (second if k in seen else first).append(k)
        seen.add(k)
    gt = first + second
    if len(gt) == 1:
        gt = [0, gt[0]]
    elif len(gt) != 2:
        gt = [0, 0]
    if head is None:
        head = ("", 0, 0, "")
        depth = 0
    return head[0], head[1], head[2], head[3], alts, gt, depth, supports, gt_qual, non_alt, in_repeat


def _num(x) -> str:
    if isinstance(x, (int,)) or (hasattr(x, "dtype") and getattr(x.dtype, "kind", "") in "iu"):
        return str(int(x))
    return "%g" % float(x)


class VCFWriter:
    """Drop-in for ``VCFWriter(all_contigs, reference_file_path, sample_name, output_dir, filename_full, filename_pepper,
    filename_variant_calling)`` (VcfWriter.py:13-32). ``reference_file_path`` may be a FASTA path (read through the
    repo's own ``FASTAHandler``: ``.fai`` lookups, no htslib) or a list of ``(contig, length)`` pairs."""

    FILES = ("full", "pepper", "variant_calling", "variant_calling_snp", "variant_calling_indel")

    def __init__(self, all_contigs, reference_file_path, sample_name, output_dir, filename_full, filename_pepper,
                 filename_variant_calling):
        if isinstance(reference_file_path, (str, os.PathLike)):
            from .ingest import FASTAHandler
            fa = FASTAHandler(str(reference_file_path))
            contigs = [(n, fa.get_chromosome_sequence_length(n)) for n in fa.get_chromosome_names()]
        else:
            contigs = [(str(n), int(l)) for n, l in reference_file_path]
        self.contigs = [n for n, _ in contigs]                       # every contig of the FASTA, as the reference (:15-17)
        self.sample_name = sample_name
        self.output_dir = output_dir
        names = dict(full=filename_full, pepper=filename_pepper, variant_calling=filename_variant_calling,
                     variant_calling_snp=filename_variant_calling + "_SNPs", variant_calling_indel=filename_variant_calling + "_INDEL")
        self.paths = {k: output_dir + v + ".vcf.gz" for k, v in names.items()}         # string concatenation, :21-25
        header = self.header_text(sample_name, contigs)
        self._out = {k: BgzfTextWriter(p) for k, p in self.paths.items()}
        self._index = {k: [] for k in self.paths}                  # per file: (contig, begin, end, text offset, text offset behind)
        for w in self._out.values():
            w.write(header)

    # the meta lines of get_vcf_header (:223-289), in its order (FORMAT/GT is declared twice there; once here)
    @staticmethod
    def header_text(sample_name: str, contigs: Iterable[Tuple[str, int]]) -> str:
        lines = ["##fileformat=VCFv4.2",
                 '##FILTER=<ID=PASS,Description="All filters passed">',
                 '##FILTER=<ID=refCall,Description="Call is homozygous">',
                 '##FILTER=<ID=lowGQ,Description="Low genotype quality">',
                 '##FILTER=<ID=lowQUAL,Description="Low variant call quality">',
                 '##FILTER=<ID=conflictPos,Description="Overlapping record">',
                 '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">',
                 '##FORMAT=<ID=DP,Number=1,Type=Integer,Description="Depth">',
                 '##FORMAT=<ID=AD,Number=A,Type=Integer,Description="Allele depth">',
                 '##FORMAT=<ID=VAF,Number=A,Type=Float,Description="Variant allele fractions.">',
                 '##FORMAT=<ID=AP,Number=A,Type=Float,Description="Maximum variant allele probability for each allele.">',
                 '##FORMAT=<ID=GQ,Number=1,Type=Float,Description="Genotype Quality">',
                 '##FORMAT=<ID=REP,Number=1,Type=String,Description="If set to 1 then variant site is considered to be ina LowCompexity repeat region">']
        lines += ["##contig=<ID=%s,length=%d>" % (n, l) for n, l in contigs]
        lines.append("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + sample_name)
        return "\n".join(lines) + "\n"

    @staticmethod
    def record_line(contig, start, ref, alts, qual, flt, gt, ap, gq, dp, ad, vaf, rep) -> str:
        """One VCF line. FORMAT order = GT first, then the keyword order of the reference's new_record call (:190-201)."""
        sample = ":".join(["/".join(str(g) for g in gt), ",".join(_num(v) for v in ap), _num(gq), str(int(dp)),
                           ",".join(str(int(v)) for v in ad), ",".join(_num(v) for v in vaf), rep])
        return "\t".join([str(contig), str(int(start) + 1), ".", ref, ",".join(alts), str(int(qual)), flt, ".",
                          "GT:AP:GQ:DP:AD:VAF:REP", sample]) + "\n"

    def write_vcf_records(self, variants_list: Dict[Tuple[str, int], List[tuple]], options) -> Tuple[int, int, int, int, int]:
        """VcfWriter.py:141-221. Returns (all, pepper, variant calling, variant calling SNPs, variant calling INDELs)."""
        n = dict.fromkeys(self.FILES, 0)
        previous_start = -1
        for key in sorted(variants_list):
            contig, start, end, ref, alts, gt, depth, supports, gt_qual, non_alt, in_repeat = \
                merge_site(variants_list[key], options.allowed_multiallelics)
            if not alts or start == previous_start:               # :150-153 (the position test ignores the contig, as there)
                continue
            previous_start = start
            qual = _phred(gt_qual)
            snp = max(len(ref), max(len(a) for a in alts)) == 1
            if snp:
                cutoff = options.snp_q_cutoff_in_lc if in_repeat else options.snp_q_cutoff
            else:
                cutoff = options.indel_q_cutoff_in_lc if in_repeat else options.indel_q_cutoff
            ref_call = gt == [0, 0]
            regenotype = ref_call or qual <= cutoff
            vaf = [round(ad / max(1, depth), 3) for ad in supports]
            line = self.record_line(contig, start, ref, alts, qual, "refCall" if ref_call else "PASS", gt, non_alt, qual,
                                    depth, supports, vaf, "1" if in_repeat else "0")
            targets = ["full"] + ((["variant_calling_snp" if snp else "variant_calling_indel", "variant_calling"])
                                  if regenotype else ["pepper"])
            for t in targets:
                u0 = self._out[t].upos
                self._out[t].write(line)
                self._index[t].append((str(contig), int(start), int(start) + len(ref), u0, self._out[t].upos))
                n[t] += 1
        return n["full"], n["pepper"], n["variant_calling"], n["variant_calling_snp"], n["variant_calling_indel"]

    def close(self):
        """Closes the five files and writes a tabix index beside each (the reference: pysam.tabix_index, VcfWriter.py:41-45)."""
        if getattr(self, "_closed", False):
            return
        self._closed = True
        for k, w in self._out.items():
            w.close()
            tid = {n: i for i, n in enumerate(self.contigs)}
            names = list(self.contigs)
            recs = []
            for contig, beg, end, u0, u1 in self._index[k]:
                if contig not in tid:                              # a contig the FASTA does not list: indexed behind the others
                    tid[contig] = len(names); names.append(contig)
                recs.append((tid[contig], beg, end, w.virtual_offset(u0), w.virtual_offset(u1)))
            write_tabix_index(self.paths[k], names, recs)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
