// TEST INFRASTRUCTURE ONLY -- never imported by the product path.
//
// oracle/_ref: the *unmodified* reference implementation of the pileup-summary hot path,
// compiled where it lies under /root/reference (no source is copied into this repo).
//
// This translation unit only supplies what the reference's unity build (pybind_api.cpp:5-12)
// would have supplied from files that need htslib (absent here):
//   * the three AlleleType constants (candidate_finder.h:23-27) used by region_summary.cpp
//   * the STL includes pulled in transitively by bam_handler.h
// and then #includes the reference's region_summary.cpp verbatim. Bindings below expose
//   (a) the reference's own pybind surface for this path (pybind_api.h:55-62,73-101,186-221) and
//   (b) a packed fast path used by tests/bench to drive the same C++ code without paying for
//       Python object construction (reads are converted to the reference's `type_read` first;
//       only RegionalSummaryGenerator::generate_max_insert_summary + generate_summary are timed).
//
// Build: see oracle/Makefile (g++ -O3 -fPIC -pipe = the flags of the reference CMakeLists.txt:5).
#include <vector>
#include <map>
#include <set>
#include <string>
#include <iostream>
#include <iomanip>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <pthread.h>

#include "read.h"       // /root/reference/pepper_variant/modules/cpp/read.h
#include "cigar.h"      // /root/reference/pepper_variant/modules/cpp/cigar.h

namespace AlleleType {  // candidate_finder.h:23-27 (that header needs htslib via bam_handler.h)
    static constexpr int SNP_ALLELE = 1;
    static constexpr int INSERT_ALLELE = 2;
    static constexpr int DELETE_ALLELE = 3;
};

#include "region_summary.cpp"   // the reference, verbatim

#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#include <pybind11/numpy.h>
namespace py = pybind11;

// ---------------------------------------------------------------------------------------------
// Packed fast path. Layout = include/pepper_b200.h PvReadBatch (one region at a time here).
// ---------------------------------------------------------------------------------------------
struct ReadSet {
    vector<type_read> reads;
    long long n_bases = 0;
};

static ReadSet* build_reads(py::array_t<int64_t> read_pos, py::array_t<int64_t> base_off,
                            py::array_t<int32_t> read_len, py::array_t<int64_t> cigar_off,
                            py::array_t<int32_t> n_ops, py::array_t<uint8_t> flags,
                            py::array_t<uint8_t> mapq, py::array_t<uint8_t> bases,
                            py::array_t<uint8_t> quals, py::array_t<uint32_t> cigar,
                            long long begin, long long end) {
    auto rp = read_pos.unchecked<1>(); auto bo = base_off.unchecked<1>();
    auto rl = read_len.unchecked<1>(); auto co = cigar_off.unchecked<1>();
    auto no = n_ops.unchecked<1>(); auto fl = flags.unchecked<1>(); auto mq = mapq.unchecked<1>();
    const uint8_t* b = bases.data(); const uint8_t* q = quals.data(); const uint32_t* c = cigar.data();
    ReadSet* rs = new ReadSet();
    rs->reads.reserve(end - begin);
    for (long long r = begin; r < end; r++) {
        type_read read;
        read.pos = rp(r);
        read.pos_end = rp(r);
        read.flags.is_reverse = (fl(r) & 1) != 0;
        read.mapping_quality = mq(r);
        read.hp_tag = 0;
        read.read_id = (int)(r - begin);
        read.sequence.assign((const char*)b + bo(r), (size_t)rl(r));
        read.base_qualities.resize(rl(r));
        for (int i = 0; i < rl(r); i++) read.base_qualities[i] = q[bo(r) + i];
        read.cigar_tuples.reserve(no(r));
        for (int k = 0; k < no(r); k++) {
            uint32_t v = c[co(r) + k];
            read.cigar_tuples.push_back(CigarOp((int)(v & 15u), (int)(v >> 4)));
        }
        rs->n_bases += rl(r);
        rs->reads.push_back(std::move(read));
    }
    return rs;
}

struct RunArgs {
    ReadSet* rs; string contig; long long ref_start, ref_end; string ref;
    double thr[9]; bool skip_indels; long long cand_start, cand_end; int window, features;
    vector<CandidateImageSummary> out;
};

static void* run_thread(void* p) {
    RunArgs* a = (RunArgs*)p;
    RegionalSummaryGenerator gen(a->contig, a->ref_start, a->ref_end, a->ref);
    gen.generate_max_insert_summary(a->rs->reads);          // AlignmentSummarizer.py:221
    a->out = gen.generate_summary(a->rs->reads, a->thr[0], a->thr[1], a->thr[2], a->thr[3], a->thr[4],
                                  a->thr[5], a->thr[6], a->thr[7], a->thr[8], a->skip_indels,
                                  a->cand_start, a->cand_end, a->window, a->features, false);  // :223-238
    return nullptr;
}

// Runs the reference generator on one region; the reference keeps ~2 MB of VLAs per 100 kbp on the
// stack (region_summary.cpp:586-589,626-628) so it is run on a thread with a 512 MB stack.
static py::dict run_region(ReadSet& rs, const string& contig, long long ref_start, long long ref_end,
                           py::bytes ref, vector<double> thr, bool skip_indels,
                           long long cand_start, long long cand_end, int window, int features) {
    if (thr.size() != 9) throw std::runtime_error("need 9 thresholds");
    RunArgs a;
    a.rs = &rs; a.contig = contig; a.ref_start = ref_start; a.ref_end = ref_end; a.ref = (string)ref;
    for (int i = 0; i < 9; i++) a.thr[i] = thr[i];
    a.skip_indels = skip_indels; a.cand_start = cand_start; a.cand_end = cand_end;
    a.window = window; a.features = features;
    {
        py::gil_scoped_release rel;
        pthread_attr_t attr; pthread_attr_init(&attr);
        pthread_attr_setstacksize(&attr, (size_t)512 << 20);
        pthread_t th;
        if (pthread_create(&th, &attr, run_thread, &a) != 0) throw std::runtime_error("pthread_create failed");
        pthread_join(th, nullptr);
        pthread_attr_destroy(&attr);
    }
    const size_t K = a.out.size();
    const int W = window + 1;
    py::array_t<int64_t> pos(K); py::array_t<int32_t> depth(K); py::array_t<int32_t> freq(K);
    py::array_t<int32_t> img({(py::ssize_t)K, (py::ssize_t)W, (py::ssize_t)features});
    py::list alleles;
    auto pp = pos.mutable_unchecked<1>(); auto dd = depth.mutable_unchecked<1>();
    auto ff = freq.mutable_unchecked<1>(); int32_t* ip = img.mutable_data();
    for (size_t k = 0; k < K; k++) {
        const CandidateImageSummary& c = a.out[k];
        pp(k) = c.position; dd(k) = c.depth;
        ff(k) = c.candidate_frequency.empty() ? -1 : c.candidate_frequency[0];
        alleles.append(py::bytes(c.candidates.empty() ? string("") : c.candidates[0]));
        for (int i = 0; i < W; i++)
            for (int j = 0; j < features; j++) ip[(k * W + i) * features + j] = c.image_matrix[i][j];
    }
    py::dict d;
    d["position"] = pos; d["depth"] = depth; d["frequency"] = freq; d["alleles"] = alleles; d["images"] = img;
    return d;
}

PYBIND11_MODULE(pv_ref_oracle, m) {
    m.doc() = "unmodified reference RegionalSummaryGenerator (test oracle / CPU baseline)";

    py::class_<CigarOp>(m, "CigarOp")                                   // pybind_api.h:186-189
        .def(py::init<>()).def(py::init<int, int>())
        .def_readwrite("cigar_op", &CigarOp::operation)
        .def_readwrite("cigar_len", &CigarOp::length);
    py::class_<type_read_flags>(m, "type_read_flags")                   // pybind_api.h:192-205
        .def(py::init())
        .def_readwrite("is_reverse", &type_read_flags::is_reverse)
        .def_readwrite("is_supplementary", &type_read_flags::is_supplementary);
    py::class_<type_read>(m, "type_read")                               // pybind_api.h:208-221
        .def(py::init<>())
        .def_readwrite("pos", &type_read::pos)
        .def_readwrite("pos_end", &type_read::pos_end)
        .def_readwrite("query_name", &type_read::query_name)
        .def_readwrite("read_id", &type_read::read_id)
        .def_readwrite("flags", &type_read::flags)
        .def_readwrite("hp_tag", &type_read::hp_tag)
        .def_readwrite("sequence", &type_read::sequence)
        .def_readwrite("cigar_tuples", &type_read::cigar_tuples)
        .def_readwrite("mapping_quality", &type_read::mapping_quality)
        .def_readwrite("base_qualities", &type_read::base_qualities);
    py::class_<CandidateImageSummary>(m, "CandidateImageSummary")       // pybind_api.h:73-82
        .def(py::init<>())
        .def_readwrite("contig", &CandidateImageSummary::contig)
        .def_readwrite("position", &CandidateImageSummary::position)
        .def_readwrite("depth", &CandidateImageSummary::depth)
        .def_readwrite("candidates", &CandidateImageSummary::candidates)
        .def_readwrite("candidate_frequency", &CandidateImageSummary::candidate_frequency)
        .def_readwrite("image_matrix", &CandidateImageSummary::image_matrix)
        .def_readwrite("base_label", &CandidateImageSummary::base_label)
        .def_readwrite("type_label", &CandidateImageSummary::type_label);
    py::class_<RegionalSummaryGenerator>(m, "RegionalSummaryGenerator") // pybind_api.h:55-62
        .def(py::init<string &, long long &, long long &, string &>())
        .def_readwrite("max_observed_insert", &RegionalSummaryGenerator::max_observed_insert)
        .def_readwrite("cumulative_observed_insert", &RegionalSummaryGenerator::cumulative_observed_insert)
        .def_readwrite("total_observered_insert_bases", &RegionalSummaryGenerator::total_observered_insert_bases)
        .def("generate_summary", &RegionalSummaryGenerator::generate_summary)
        .def("generate_max_insert_summary", &RegionalSummaryGenerator::generate_max_insert_summary);

    py::class_<ReadSet>(m, "ReadSet")
        .def_readonly("n_bases", &ReadSet::n_bases)
        .def("__len__", [](const ReadSet& r) { return r.reads.size(); });
    m.def("build_reads", &build_reads, py::return_value_policy::take_ownership);
    m.def("run_region", &run_region);
}
