// PEPPER_VARIANT -- drop-in pybind11 module for the hot-path part of the reference's module of the same name
// (/root/reference/pepper_variant/modules/cpp/pybind_api.h:23-278, imported by the reference as
// `from pepper_variant.build import PEPPER_VARIANT`).
//
// Same class names, constructor signatures, attribute names, pickling tuples and value semantics for everything
// AlignmentSummarizer.create_summary touches (pybind_api.h:55-62, 73-101, 133-160, 186-221):
//     RegionalSummaryGenerator, CandidateImageSummary, CandidateImagePrediction, type_read, type_read_flags, CigarOp.
// RegionalSummaryGenerator.generate_summary packs the reads into the SoA batch and calls the C-ABI
// (pv_summary_regions_host in libpepper_b200.so): the work happens in the CUDA kernels, never on the CPU.
// Classes of the reference module that are outside the hot path (htslib I/O, legacy generators, training labels) are
// exported by name and raise on use -- see INTEGRATION.md.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "pepper_b200.h"

namespace py = pybind11;
using std::string;
using std::vector;

namespace pvb {   // own namespace: the test oracle registers the reference's global types of the same names

struct CigarOp {                                   // cigar.h:30-53
    int operation = -1;
    int length = 0;
    CigarOp() {}
    CigarOp(int op, int len) : operation(op), length(len) {}
};

struct type_read_flags {                           // read.h:15-58
    bool is_paired = false, is_proper_pair = false, is_unmapped = false, is_mate_unmapped = false, is_reverse = false,
         is_mate_is_reverse = false, is_read1 = false, is_read2 = false, is_secondary = false, is_qc_failed = false,
         is_duplicate = false, is_supplementary = false;
};

struct type_read {                                 // read.h:60-108
    long long pos = 0, pos_end = 0;
    string query_name;
    type_read_flags flags;
    string sequence;
    vector<CigarOp> cigar_tuples;
    vector<int> bad_indicies;
    int mapping_quality = 0;
    vector<int> base_qualities;
    int read_id = 0;
    int hp_tag = 0;
    void set_read_id(int id) { read_id = id; }
    bool less(const type_read& that) const { return pos == that.pos ? pos_end < that.pos_end : pos < that.pos; }
};

struct CandidateImageSummary {                     // region_summary.h:88-111
    string contig;
    long long position = 0;
    vector<vector<int>> image_matrix;
    vector<string> candidates;
    vector<int> candidate_frequency;
    int depth = 0;
    uint8_t base_label = 0, type_label = 0;
};

struct CandidateImagePrediction {                  // region_summary.h:114-136
    string contig;
    long long position = 0;
    int depth = 0;
    vector<string> candidates;
    vector<int> candidate_frequency;
    vector<float> prediction_base, prediction_type;
};

class RegionalSummaryGenerator {                   // region_summary.h:138-210
    string contig;
    long long ref_start, ref_end;
    string reference_sequence;

public:
    vector<uint64_t> max_observed_insert, cumulative_observed_insert;
    uint64_t total_observered_insert_bases = 0;

    RegionalSummaryGenerator(string contig_, long long region_start, long long region_end, string reference)
        : contig(std::move(contig_)), ref_start(region_start), ref_end(region_end), reference_sequence(std::move(reference)) {
        if (region_end < region_start) throw std::invalid_argument("region_end < region_start");
        max_observed_insert.assign((size_t)(region_end - region_start + 1), 0);          // region_summary.cpp:15-16
        cumulative_observed_insert.assign((size_t)(region_end - region_start + 1), 0);
    }

    // region_summary.cpp:69-96 with GENERATE_INDELS == false: no insert columns, every offset stays 0
    void generate_max_insert_summary(const vector<type_read>&) {}

    void generate_labels(py::object, py::object) {
        throw std::runtime_error("generate_labels (training labels) is outside the B200 hot path");
    }

    vector<CandidateImageSummary> generate_summary(const vector<type_read>& reads, double min_snp_baseq, double min_indel_baseq,
                                                   double snp_freq_threshold, double insert_freq_threshold,
                                                   double delete_freq_threshold, double min_coverage_threshold,
                                                   double snp_candidate_freq_threshold, double indel_candidate_freq_threshold,
                                                   double candidate_support_threshold, bool skip_indels,
                                                   long long candidate_region_start, long long candidate_region_end,
                                                   int candidate_window_size, int feature_size, bool train_mode) {
        if (train_mode) throw std::runtime_error("train_mode=True is outside the B200 hot path (inference only)");
        const int64_t L = ref_end - ref_start + 1;
        if ((int64_t)reference_sequence.size() < L) throw std::invalid_argument("reference_sequence shorter than the region");
        if (reads.size() > 32767) throw std::invalid_argument("more than 32767 reads in a region");

        // ---- pack (read.h AoS -> PvReadBatch SoA) -------------------------------------------------------------------------
        const size_t n = reads.size();
        vector<int64_t> read_pos(n), base_off(n), cigar_off(n);
        vector<int32_t> read_len(n), n_ops(n);
        vector<uint8_t> flags(n), mapq(n), bases, quals;
        vector<uint32_t> cigar;
        for (size_t i = 0; i < n; i++) {
            const type_read& r = reads[i];
            if (r.base_qualities.size() != r.sequence.size()) throw std::invalid_argument("base_qualities and sequence differ in length");
            read_pos[i] = r.pos; base_off[i] = (int64_t)bases.size(); read_len[i] = (int32_t)r.sequence.size();
            cigar_off[i] = (int64_t)cigar.size(); n_ops[i] = (int32_t)r.cigar_tuples.size();
            flags[i] = r.flags.is_reverse ? 1 : 0;
            mapq[i] = (uint8_t)(r.mapping_quality < 0 ? 0 : (r.mapping_quality > 255 ? 255 : r.mapping_quality));
            bases.insert(bases.end(), r.sequence.begin(), r.sequence.end());
            for (int q : r.base_qualities) {
                if (q < 0 || q > 255) throw std::invalid_argument("base quality outside [0, 255]");
                quals.push_back((uint8_t)q);
            }
            while (bases.size() & 15) { bases.push_back(0); quals.push_back(0); }
            for (const CigarOp& c : r.cigar_tuples) {
                if (c.operation < 0 || c.operation > 15 || c.length < 0 || c.length >= (1 << 28)) throw std::invalid_argument("bad CIGAR op");
                cigar.push_back(((uint32_t)c.length << 4) | (uint32_t)c.operation);
            }
        }
        int64_t rs = ref_start, re = ref_end, cs = candidate_region_start, ce = candidate_region_end, ro = 0,
                rl = (int64_t)reference_sequence.size(), rb[2] = {0, (int64_t)n};
        PvReadBatch b;
        b.bases4 = nullptr;
        b.n_reads = (int64_t)n; b.n_bases = (int64_t)bases.size(); b.n_ops = (int64_t)cigar.size(); b.n_ref = rl; b.n_regions = 1;
        b.read_pos = read_pos.data(); b.read_base_off = base_off.data(); b.read_len = read_len.data();
        b.read_cigar_off = cigar_off.data(); b.read_n_ops = n_ops.data(); b.read_flags = flags.data(); b.read_mapq = mapq.data();
        b.bases = bases.data(); b.quals = quals.data(); b.cigar = cigar.data();
        b.region_ref_start = &rs; b.region_ref_end = &re; b.region_cand_start = &cs; b.region_cand_end = &ce;
        b.region_ref_off = &ro; b.region_ref_len = &rl; b.region_read_begin = rb;
        b.ref = (const uint8_t*)reference_sequence.data();
        PvThresholds t;
        t.min_snp_baseq = min_snp_baseq; t.min_indel_baseq = min_indel_baseq; t.snp_freq = snp_freq_threshold;
        t.insert_freq = insert_freq_threshold; t.delete_freq = delete_freq_threshold; t.min_coverage = min_coverage_threshold;
        t.snp_candidate_freq = snp_candidate_freq_threshold; t.indel_candidate_freq = indel_candidate_freq_threshold;
        t.candidate_support = candidate_support_threshold; t.skip_indels = skip_indels ? 1 : 0; t._pad = 0;

        // ---- run on the GPU; grow the capacity when the region has more candidates --------------------------------------
        int64_t cap = L / 16 + 1024, found = 0;
        vector<int16_t> win; vector<int64_t> pos; vector<int32_t> reg, dep, frq; vector<uint8_t> al, aln;
        for (;;) {
            win.resize((size_t)cap * PV_WINDOW * PV_FEATURES); pos.resize(cap); reg.resize(cap); dep.resize(cap); frq.resize(cap);
            al.resize((size_t)cap * PV_ALLELE_BYTES); aln.resize(cap);
            PvCandidates out{cap, win.data(), pos.data(), reg.data(), dep.data(), frq.data(), al.data(), aln.data()};
            int rc;
            {
                py::gil_scoped_release release;
                rc = pv_summary_regions_host(&b, &t, candidate_window_size, feature_size, &out, &found, nullptr);
            }
            if (rc == PV_EOVERFLOW) { cap = found > cap ? found + 16 : cap * 4; continue; }
            if (rc != PV_OK) throw std::runtime_error(string("pepper_b200: ") + pv_last_error());
            break;
        }
        // ---- unpack to the reference's value types ---------------------------------------------------------------------------
        vector<CandidateImageSummary> res((size_t)found);
        for (int64_t k = 0; k < found; k++) {
            CandidateImageSummary& c = res[(size_t)k];
            c.contig = contig; c.position = pos[k]; c.depth = dep[k];
            c.candidates.push_back(string((const char*)&al[(size_t)k * PV_ALLELE_BYTES], (size_t)aln[k]));
            c.candidate_frequency.push_back(frq[k]);
            c.image_matrix.assign(PV_WINDOW, vector<int>(PV_FEATURES));
            const int16_t* w = &win[(size_t)k * PV_WINDOW * PV_FEATURES];
            for (int i = 0; i < PV_WINDOW; i++)
                for (int j = 0; j < PV_FEATURES; j++) c.image_matrix[i][j] = w[i * PV_FEATURES + j];
        }
        return res;
    }
};

}  // namespace pvb
using namespace pvb;

static void out_of_scope(py::module_& m, const char* name, const char* why) {
    string n = name, w = why;
    m.attr(name) = py::cpp_function([n, w](py::args, py::kwargs) -> py::object {
        throw std::runtime_error("PEPPER_VARIANT." + n + " is outside the B200 hot path (" + w + "); see INTEGRATION.md");
    });
}

PYBIND11_MODULE(PEPPER_VARIANT, m) {
    m.doc() = "B200-native drop-in for the hot-path classes of the reference PEPPER_VARIANT module";

    py::class_<CigarOp>(m, "CigarOp")                                   // pybind_api.h:186-189
        .def(py::init<>()).def(py::init<int, int>())
        .def_readwrite("cigar_op", &CigarOp::operation)
        .def_readwrite("cigar_len", &CigarOp::length);

    py::class_<type_read_flags>(m, "type_read_flags")                   // pybind_api.h:192-205
        .def(py::init())
        .def_readwrite("is_paired", &type_read_flags::is_paired)
        .def_readwrite("is_proper_pair", &type_read_flags::is_proper_pair)
        .def_readwrite("is_unmapped", &type_read_flags::is_unmapped)
        .def_readwrite("is_mate_unmapped", &type_read_flags::is_mate_unmapped)
        .def_readwrite("is_reverse", &type_read_flags::is_reverse)
        .def_readwrite("is_mate_is_reverse", &type_read_flags::is_mate_is_reverse)
        .def_readwrite("is_read1", &type_read_flags::is_read1)
        .def_readwrite("is_read2", &type_read_flags::is_read2)
        .def_readwrite("is_secondary", &type_read_flags::is_secondary)
        .def_readwrite("is_qc_failed", &type_read_flags::is_qc_failed)
        .def_readwrite("is_duplicate", &type_read_flags::is_duplicate)
        .def_readwrite("is_supplementary", &type_read_flags::is_supplementary);

    py::class_<type_read>(m, "type_read")                               // pybind_api.h:208-221 (+ a constructor)
        .def(py::init<>())
        .def("set_read_id", &type_read::set_read_id)
        .def("__lt__", &type_read::less, py::is_operator())
        .def_readwrite("pos", &type_read::pos)
        .def_readwrite("pos_end", &type_read::pos_end)
        .def_readwrite("query_name", &type_read::query_name)
        .def_readwrite("read_id", &type_read::read_id)
        .def_readwrite("flags", &type_read::flags)
        .def_readwrite("hp_tag", &type_read::hp_tag)
        .def_readwrite("sequence", &type_read::sequence)
        .def_readwrite("cigar_tuples", &type_read::cigar_tuples)
        .def_readwrite("mapping_quality", &type_read::mapping_quality)
        .def_readwrite("base_qualities", &type_read::base_qualities)
        .def_readwrite("bad_indicies", &type_read::bad_indicies);

    py::class_<RegionalSummaryGenerator>(m, "RegionalSummaryGenerator") // pybind_api.h:55-62
        .def(py::init<string&, long long&, long long&, string&>())
        .def_readwrite("max_observed_insert", &RegionalSummaryGenerator::max_observed_insert)
        .def_readwrite("cumulative_observed_insert", &RegionalSummaryGenerator::cumulative_observed_insert)
        .def_readwrite("total_observered_insert_bases", &RegionalSummaryGenerator::total_observered_insert_bases)
        .def("generate_summary", &RegionalSummaryGenerator::generate_summary)
        .def("generate_labels", &RegionalSummaryGenerator::generate_labels)
        .def("generate_max_insert_summary", &RegionalSummaryGenerator::generate_max_insert_summary);

    py::class_<CandidateImageSummary>(m, "CandidateImageSummary")       // pybind_api.h:73-101
        .def(py::init<>())
        .def(py::init([](string& contig, long long& position, int& depth, vector<string>& candidates, vector<int>& freq,
                         vector<vector<int>>& image, int& base_label, int& type_label) {
            CandidateImageSummary c;
            c.contig = contig; c.position = position; c.depth = depth; c.candidates = candidates; c.candidate_frequency = freq;
            c.image_matrix = image; c.base_label = (uint8_t)base_label; c.type_label = (uint8_t)type_label;
            return c;
        }))
        .def_readwrite("contig", &CandidateImageSummary::contig)
        .def_readwrite("position", &CandidateImageSummary::position)
        .def_readwrite("depth", &CandidateImageSummary::depth)
        .def_readwrite("candidates", &CandidateImageSummary::candidates)
        .def_readwrite("candidate_frequency", &CandidateImageSummary::candidate_frequency)
        .def_readwrite("image_matrix", &CandidateImageSummary::image_matrix)
        .def_readwrite("base_label", &CandidateImageSummary::base_label)
        .def_readwrite("type_label", &CandidateImageSummary::type_label)
        .def(py::pickle(
            [](const CandidateImageSummary& p) {
                return py::make_tuple(p.contig, p.position, p.depth, p.candidates, p.candidate_frequency, p.image_matrix,
                                      p.base_label, p.type_label);
            },
            [](py::tuple t) {
                if (t.size() != 8) throw std::runtime_error("Invalid state!");
                CandidateImageSummary c;
                c.contig = t[0].cast<string>(); c.position = t[1].cast<long long>(); c.depth = t[2].cast<int>();
                c.candidates = t[3].cast<vector<string>>(); c.candidate_frequency = t[4].cast<vector<int>>();
                c.image_matrix = t[5].cast<vector<vector<int>>>(); c.base_label = (uint8_t)t[6].cast<int>();
                c.type_label = (uint8_t)t[7].cast<int>();
                return c;
            }));

    py::class_<CandidateImagePrediction>(m, "CandidateImagePrediction") // pybind_api.h:133-160
        .def(py::init<>())
        .def(py::init([](string& contig, long long& position, int& depth, vector<string>& candidates, vector<int>& freq,
                         vector<float>& pb, vector<float>& pt) {
            CandidateImagePrediction c;
            c.contig = contig; c.position = position; c.depth = depth; c.candidates = candidates; c.candidate_frequency = freq;
            c.prediction_base = pb; c.prediction_type = pt;
            return c;
        }))
        .def_readwrite("contig", &CandidateImagePrediction::contig)
        .def_readwrite("position", &CandidateImagePrediction::position)
        .def_readwrite("depth", &CandidateImagePrediction::depth)
        .def_readwrite("candidates", &CandidateImagePrediction::candidates)
        .def_readwrite("candidate_frequency", &CandidateImagePrediction::candidate_frequency)
        .def_readwrite("prediction_base", &CandidateImagePrediction::prediction_base)
        .def_readwrite("prediction_type", &CandidateImagePrediction::prediction_type)
        .def(py::pickle(
            [](const CandidateImagePrediction& p) {
                return py::make_tuple(p.contig, p.position, p.depth, p.candidates, p.candidate_frequency, p.prediction_base,
                                      p.prediction_type);
            },
            [](py::tuple t) {
                if (t.size() != 7) throw std::runtime_error("Invalid state!");
                CandidateImagePrediction c;
                c.contig = t[0].cast<string>(); c.position = t[1].cast<long long>(); c.depth = t[2].cast<int>();
                c.candidates = t[3].cast<vector<string>>(); c.candidate_frequency = t[4].cast<vector<int>>();
                c.prediction_base = t[5].cast<vector<float>>(); c.prediction_type = t[6].cast<vector<float>>();
                return c;
            }));

    // names of the reference module that are not on the hot path
    out_of_scope(m, "BAM_handler", "htslib BAM reader: 'next' row 1 of SURVEY.md section 8f");
    out_of_scope(m, "FASTA_handler", "htslib FASTA reader: 'next' row 1 of SURVEY.md section 8f");
    out_of_scope(m, "SummaryGenerator", "legacy per-position generator, not called by call_variant");
    out_of_scope(m, "RegionalSummaryGeneratorHP", "haplotype-aware variant, -hp flag only");
    out_of_scope(m, "CandidateFinder", "legacy C++ candidate finder, unreachable from call_variant");
    out_of_scope(m, "CandidateFinderHP", "legacy C++ candidate finder, unreachable from call_variant");
    out_of_scope(m, "type_truth_record", "training labels");
}
