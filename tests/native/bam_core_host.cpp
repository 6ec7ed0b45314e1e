// TEST HARNESS ONLY (built by tests/test_bam_core_cpu.py with g++): runs the __host__ __device__ core of the GPU ingest
// (pepper-thesis_b200/csrc/bam_core.cuh) on the CPU, record by record, so that its DEFLATE decoder and its get_reads
// clipping can be checked against zlib and against the CPU ingest / the compiled reference without a GPU.
#include <cstdint>
#include <cstring>
#include <vector>
#include "bam_core.cuh"

using namespace bamcore;

extern "C" int pvt_inflate(const uint8_t* in, int64_t n_in, uint8_t* out, int64_t n_out) {
    std::vector<InflateState> s(1);
    return inflate_block(in, n_in, out, n_out, s[0]) ? 0 : -1;
}

extern "C" uint32_t pvt_crc32(const uint8_t* p, int64_t n) {
    uint32_t t[256];
    crc32_table(t);
    return crc32_bytes(t, p, n);
}

// all records from first_off on (or, with rec_offs, exactly those n_rec records): reads of `tid` cut to [start, stop] ->
// packed arrays (caller-sized); returns the read count
extern "C" int64_t pvt_get_reads(const uint8_t* U, int64_t u_size, int64_t first_off, const int64_t* rec_offs, int64_t n_rec, int32_t tid, int64_t start, int64_t stop,
                                 int32_t supp, int32_t min_mapq, int64_t* pos, int64_t* pos_end, int32_t* len, int32_t* n_ops,
                                 int32_t* hp, uint8_t* rev, uint8_t* mapq, uint8_t* bases, uint8_t* quals, uint32_t* cigar,
                                 int64_t* n_bases_out, int64_t* n_ops_out) {
    int64_t off = first_off, n = 0, nb = 0, no = 0;
    for (int64_t k = 0; rec_offs ? k < n_rec : off + 4 <= u_size; k++) {
        const RecHdr h = parse_record(U, rec_offs ? rec_offs[k] : off, u_size);
        if (!h.ok) return -1;
        off = h.rec_end;
        if (h.tid != tid) continue;
        const int64_t endpos = record_endpos(U, h);
        if (h.pos >= stop || endpos <= start) continue;
        if (!record_passes(h, supp, min_mapq)) continue;
        const Clip c = clip_walk<true>(U, h, start, stop, cigar + no);
        if (c.bad || c.n_bases == 0) continue;
        if (c.split) return -2;
        pos[n] = c.pos_start; pos_end[n] = c.pos_end; len[n] = (int32_t)c.n_bases; n_ops[n] = c.n_ops;
        hp[n] = parse_hp(U, h.aux_off, h.rec_end); rev[n] = (h.flag & 0x10) ? 1 : 0; mapq[n] = (uint8_t)h.mapq;
        for (int64_t i = 0; i < c.n_bases; i++) { bases[nb + i] = record_base(U, h, c.idx0 + i); quals[nb + i] = U[h.qual_off + c.idx0 + i]; }
        nb += c.n_bases; no += c.n_ops; n++;
    }
    *n_bases_out = nb; *n_ops_out = no;
    return n;
}
