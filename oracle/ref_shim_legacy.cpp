// TEST INFRASTRUCTURE ONLY -- never imported by the product path.
//
// oracle/_ref/pv_ref_legacy: the *unmodified* LEGACY SummaryGenerator / ImageSummary of the variant module
// (/root/reference/pepper_variant/modules/cpp/summary_generator.cpp, bound at pybind_api.h:24-43), compiled where it lies.
// Its header pulls in bam_handler.h, which includes htslib headers that are absent here: oracle/hts_stub/ supplies empty
// stand-ins, nothing of BAM_handler is compiled or linked.
// One packed entry point drives generate_summary (:455-487) and chunk_image (:491-536) from the SoA batch arrays.
#include <vector>
#include <map>
#include <set>
#include <string>
#include <iostream>
#include <cstdint>
#include <cstring>

#include "summary_generator.cpp"   // the reference, verbatim (-I /root/reference/pepper_variant/modules/cpp)

#include <pybind11/pybind11.h>
#include <pybind11/numpy.h>
#include <pybind11/stl.h>
namespace py = pybind11;

// reads of one region from packed arrays -> dict(image, genomic_pos, ref_image, longest_insert_count, chunk_*)
static py::dict legacy_summary(py::array_t<int64_t> read_pos, py::array_t<int64_t> base_off, py::array_t<int32_t> read_len,
                               py::array_t<int64_t> cigar_off, py::array_t<int32_t> n_ops, py::array_t<uint8_t> flags,
                               py::array_t<uint8_t> mapq, py::array_t<uint8_t> bases, py::array_t<uint8_t> quals,
                               py::array_t<uint32_t> cigar, int64_t r_begin, int64_t r_end, std::string ref, long long start,
                               long long end, int chunk_size, int chunk_overlap) {
    std::vector<type_read> reads;
    for (int64_t r = r_begin; r < r_end; r++) {
        type_read t;
        t.pos = read_pos.at(r); t.pos_end = 0; t.mapping_quality = mapq.at(r); t.read_id = 0; t.hp_tag = 0;
        memset(&t.flags, 0, sizeof(t.flags));
        t.flags.is_reverse = flags.at(r) & 1;
        t.sequence.assign((const char*)bases.data() + base_off.at(r), (size_t)read_len.at(r));
        // the legacy walk READS a quality per base (its value is never used, :116,:137,:151); one spare entry keeps the read
        // behind a trailing deletion inside the vector
        for (int i = 0; i < read_len.at(r); i++) t.base_qualities.push_back((int)quals.at(base_off.at(r) + i));
        t.base_qualities.push_back(0);
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
    py::dict d;
    d["image"] = img; d["genomic_pos"] = gp; d["ref_image"] = g.ref_image;
    d["longest_insert_count"] = g.longest_insert_count;
    ImageSummary s = g.chunk_image(chunk_size, chunk_overlap, 10);
    d["chunk_images"] = s.images; d["chunk_positions"] = s.positions; d["chunk_refs"] = s.refs; d["chunk_labels"] = s.labels;
    d["chunk_ids"] = s.chunk_ids;
    return d;
}

PYBIND11_MODULE(pv_ref_legacy, m) {
    m.doc() = "unmodified reference legacy SummaryGenerator / ImageSummary of the variant module (test oracle)";
    m.def("legacy_summary", &legacy_summary);
}
