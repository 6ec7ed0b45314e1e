// TEST INFRASTRUCTURE ONLY -- never imported by the product path.
//
// oracle/_ref/pv_ref_polisher: the *unmodified* reference SummaryGenerator of the polisher
// (/root/reference/pepper/modules/src/pileup_summary/summary_generator.cpp), compiled where it lies. Its header pulls
// in dataio/bam_handler.h, which includes htslib headers that are absent here: oracle/hts_stub/ supplies empty
// stand-ins (three opaque pointer types), nothing of BAM_handler is compiled or linked.
// One packed entry point drives generate_summary (summary_generator.cpp:370-392) from the SoA batch arrays.
#include <vector>
#include <map>
#include <set>
#include <string>
#include <iostream>
#include <cstdint>
#include <cstring>

#include "pepper/modules/src/pileup_summary/summary_generator.cpp"   // the reference, verbatim (-I /root/reference)

#include <pybind11/pybind11.h>
#include <pybind11/numpy.h>
namespace py = pybind11;

// reads of one region from packed arrays -> (image uint8 [rows][10], genomic_pos int64 [rows][2])
static py::tuple polisher_summary(py::array_t<int64_t> read_pos, py::array_t<int64_t> base_off, py::array_t<int32_t> read_len,
                                  py::array_t<int64_t> cigar_off, py::array_t<int32_t> n_ops, py::array_t<uint8_t> flags,
                                  py::array_t<uint8_t> mapq, py::array_t<uint8_t> bases, py::array_t<uint32_t> cigar,
                                  int64_t r_begin, int64_t r_end, std::string ref, long long start, long long end) {
    std::vector<type_read> reads;
    for (int64_t r = r_begin; r < r_end; r++) {
        type_read t;
        t.pos = read_pos.at(r); t.pos_end = 0; t.mapping_quality = mapq.at(r); t.read_id = 0; t.hp_tag = 0;
        memset(&t.flags, 0, sizeof(t.flags));
        t.flags.is_reverse = flags.at(r) & 1;
        t.sequence.assign((const char*)bases.data() + base_off.at(r), (size_t)read_len.at(r));
        for (int k = 0; k < n_ops.at(r); k++) { const uint32_t w = cigar.at(cigar_off.at(r) + k); t.cigar_tuples.emplace_back((int)(w & 15u), (int)(w >> 4)); }
        reads.push_back(t);
    }
    SummaryGenerator g(ref, "c", start, end);
    g.generate_summary(reads, start, end);
    const size_t n = g.image.size();
    py::array_t<uint8_t> img({n, (size_t)10});
    py::array_t<int64_t> gp({g.genomic_pos.size(), (size_t)2});
    auto im = img.mutable_unchecked<2>(); auto gm = gp.mutable_unchecked<2>();
    for (size_t i = 0; i < n; i++) for (int j = 0; j < 10; j++) im(i, j) = g.image[i][j];
    for (size_t i = 0; i < g.genomic_pos.size(); i++) { gm(i, 0) = g.genomic_pos[i].first; gm(i, 1) = g.genomic_pos[i].second; }
    return py::make_tuple(img, gp);
}

PYBIND11_MODULE(pv_ref_polisher, m) {
    m.doc() = "unmodified reference polisher SummaryGenerator (test oracle)";
    m.def("polisher_summary", &polisher_summary);
}
