// TEST INFRASTRUCTURE ONLY -- never imported by the product path.
//
// oracle/_ref/pv_ref_bam: the *unmodified* reference BAM_handler and FASTA_handler
// (/root/reference/pepper_variant/modules/cpp/bam_handler.cpp, fasta_handler.cpp), compiled where they lie, over
// oracle/hts_mini/ -- a minimal stand-in for the htslib 1.9 calls they make (htslib is an absent third-party
// dependency: pepper/modules/htslib.cmake:9). Everything the reference itself decides -- which records are kept
// (flags, mapq: :137-150), how a read is cut to [start, stop] (:178-306), the aux walk for HP (:313-421), upper-casing
// of reference bases (fasta_handler.cpp:49) -- is the reference's own compiled code.
#include <cerrno>
#include "bam_handler.cpp"     // the reference, verbatim (-I $(REFCPP))
#include "fasta_handler.cpp"   // the reference, verbatim

#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
namespace py = pybind11;

static py::list get_reads(BAM_handler& h, std::string contig, long long start, long long stop, bool supp, int min_mapq, int min_baseq) {
    py::list out;
    for (auto& r : h.get_reads(contig, start, stop, supp, min_mapq, min_baseq)) {
        py::list cig;
        for (auto& c : r.cigar_tuples) cig.append(py::make_tuple(c.operation, c.length));
        py::dict d;
        d["query_name"] = r.query_name; d["pos"] = r.pos; d["pos_end"] = r.pos_end; d["sequence"] = py::bytes(r.sequence);
        d["base_qualities"] = r.base_qualities; d["cigar_tuples"] = cig; d["is_reverse"] = (bool)r.flags.is_reverse;
        d["is_supplementary"] = (bool)r.flags.is_supplementary; d["mapping_quality"] = r.mapping_quality;
        d["hp_tag"] = r.hp_tag; d["bad_indicies"] = r.bad_indicies;
        out.append(d);
    }
    return out;
}

PYBIND11_MODULE(pv_ref_bam, m) {
    m.doc() = "unmodified reference BAM_handler / FASTA_handler over a minimal htslib stand-in (test oracle)";
    py::class_<BAM_handler>(m, "BAM_handler")
        .def(py::init<const std::string&>())
        .def("get_reads", &get_reads)
        .def("get_chromosome_sequence_names", &BAM_handler::get_chromosome_sequence_names)
        .def("get_sample_names", &BAM_handler::get_sample_names)
        .def("get_chromosome_sequence_names_with_length", [](BAM_handler& h) {
            py::list out;
            for (auto& s : h.get_chromosome_sequence_names_with_length()) out.append(py::make_tuple(s.sequence_name, s.sequence_length));
            return out;
        });
    py::class_<FASTA_handler>(m, "FASTA_handler")
        .def(py::init<const std::string&>())
        .def("get_reference_sequence", [](FASTA_handler& f, std::string c, long long a, long long b) { return py::bytes(f.get_reference_sequence(c, a, b)); })
        .def("get_chromosome_sequence_length", &FASTA_handler::get_chromosome_sequence_length)
        .def("get_chromosome_names", &FASTA_handler::get_chromosome_names);
}
