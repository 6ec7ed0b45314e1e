// PEPPER_VARIANT -- drop-in pybind11 module for the hot-path part of the reference's module of the same name
// (/root/reference/pepper_variant/modules/cpp/pybind_api.h:23-278, imported by the reference as
// `from pepper_variant.build import PEPPER_VARIANT`).
//
// Same class names, constructor signatures, attribute names, pickling tuples and value semantics for everything
// AlignmentSummarizer.create_summary touches (pybind_api.h:55-62, 73-101, 133-160, 186-221):
//     RegionalSummaryGenerator, CandidateImageSummary, CandidateImagePrediction, type_read, type_read_flags, CigarOp.
// RegionalSummaryGenerator.generate_summary packs the reads into the SoA batch and calls the C-ABI
// (pv_summary_regions_host in libpepper_b200.so): the work happens in the CUDA kernels, never on the CPU.
// BAM_handler / FASTA_handler are bound to libpv_ingest.so (no htslib). Classes of the reference module that are outside
// the hot path (legacy generators, training labels) are exported by name and raise on use -- see INTEGRATION.md.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "pepper_b200.h"
#include "pepper_ingest.h"
#include <set>

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
        memset(&b, 0, sizeof(b));
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

struct type_sequence {                              // sequence.h (bam_handler.h): name + length of a BAM target
    string sequence_name;
    int sequence_length = 0;
};

// BAM_handler (bam_handler.cpp) over libpv_ingest.so: no htslib. Errors raise instead of exit(EXIT_FAILURE).
class BAM_handler {
    PvBamFile* f = nullptr;
    static void check(int rc) { if (rc != 0) throw std::runtime_error(pv_ingest_last_error()); }

public:
    explicit BAM_handler(const string& path) { check(pv_bam_open(path.c_str(), nullptr, &f)); }
    BAM_handler(const BAM_handler&) = delete;
    ~BAM_handler() { pv_bam_close(f); }

    vector<string> get_chromosome_sequence_names() {
        vector<string> v;
        for (int i = 0; i < pv_bam_n_targets(f); i++) v.push_back(pv_bam_target_name(f, i));
        return v;
    }
    vector<type_sequence> get_chromosome_sequence_names_with_length() {
        vector<type_sequence> v;
        for (int i = 0; i < pv_bam_n_targets(f); i++) { type_sequence s; s.sequence_name = pv_bam_target_name(f, i); s.sequence_length = (int)pv_bam_target_len(f, i); v.push_back(s); }
        return v;
    }
    std::set<string> get_sample_names() {
        const int64_t n = pv_bam_sample_names(f, nullptr, 0);
        string buf((size_t)n + 1, '\0');
        pv_bam_sample_names(f, &buf[0], n + 1);
        std::set<string> out;
        size_t p = 0;
        while (p < (size_t)n) { size_t e = buf.find('\n', p); if (e == string::npos || e > (size_t)n) e = (size_t)n; out.insert(buf.substr(p, e - p)); p = e + 1; }
        return out;
    }
    // bam_handler.cpp:115-444
    vector<type_read> get_reads(const string& chromosome, long long start, long long stop, bool include_supplementary,
                                int min_mapq, int min_baseq) {
        PvIngestOptions o; memset(&o, 0, sizeof(o));
        o.include_supplementary = include_supplementary; o.min_mapq = min_mapq; o.min_baseq = min_baseq; o.threads = 1;
        PvIngestBatch* b = nullptr;
        check(pv_bam_get_reads(f, chromosome.c_str(), start, stop, &o, &b));
        PvReadBatch v;
        pv_ingest_view(b, &v);
        const int32_t* hp = pv_ingest_hp_tags(b);
        const int64_t* pe = pv_ingest_pos_end(b);
        const uint16_t* fl = pv_ingest_bam_flags(b);
        int64_t nbytes = 0;
        const char* names = pv_ingest_query_names(b, &nbytes);
        vector<type_read> out((size_t)v.n_reads);
        for (int64_t i = 0; i < v.n_reads; i++) {
            type_read& r = out[(size_t)i];
            r.query_name = names; names += r.query_name.size() + 1;
            r.pos = v.read_pos[i]; r.pos_end = pe[i];
            const uint8_t* bs = v.bases + v.read_base_off[i];
            const uint8_t* qs = v.quals + v.read_base_off[i];
            const int n = v.read_len[i];
            r.sequence.assign((const char*)bs, (size_t)n);
            r.base_qualities.assign(qs, qs + n);
            for (int j = 0; j < n; j++)                      // bad_indicies, :206-212
                if ((int)qs[j] < min_baseq || (bs[j] != 'A' && bs[j] != 'C' && bs[j] != 'G' && bs[j] != 'T')) r.bad_indicies.push_back(j);
            r.bad_indicies.push_back(n + 1);                 // :307
            const uint32_t* cg = v.cigar + v.read_cigar_off[i];
            for (int k = 0; k < v.read_n_ops[i]; k++) r.cigar_tuples.emplace_back((int)(cg[k] & 15u), (int)(cg[k] >> 4));
            r.mapping_quality = v.read_mapq[i];
            r.hp_tag = hp[i];
            const int g = fl[i];                              // get_read_flags, :72-87
            r.flags.is_paired = g & 0x1; r.flags.is_proper_pair = g & 0x2; r.flags.is_unmapped = g & 0x4;
            r.flags.is_mate_unmapped = g & 0x8; r.flags.is_reverse = g & 0x10; r.flags.is_mate_is_reverse = g & 0x20;
            r.flags.is_read1 = g & 0x40; r.flags.is_read2 = g & 0x80; r.flags.is_secondary = g & 0x100;
            r.flags.is_qc_failed = g & 0x200; r.flags.is_duplicate = g & 0x400; r.flags.is_supplementary = g & 0x800;
        }
        pv_ingest_free(b);
        return out;
    }
};

class FASTA_handler {                              // fasta_handler.cpp
    PvFastaFile* f = nullptr;

public:
    explicit FASTA_handler(const string& path) { if (pv_fasta_open(path.c_str(), &f) != 0) throw std::runtime_error(pv_ingest_last_error()); }
    FASTA_handler(const FASTA_handler&) = delete;
    ~FASTA_handler() { pv_fasta_close(f); }
    string get_reference_sequence(const string& region, long long start, long long stop) {
        string out((size_t)(stop > start ? stop - start : 0), '\0');
        int64_t got = 0;
        if (pv_fasta_fetch(f, region.c_str(), start, stop, out.empty() ? nullptr : &out[0], &got) != 0) throw std::runtime_error(pv_ingest_last_error());
        out.resize((size_t)got);
        return out;
    }
    int get_chromosome_sequence_length(const string& name) { return (int)pv_fasta_seq_len(f, name.c_str()); }
    vector<string> get_chromosome_names() {
        vector<string> v;
        for (int i = 0; i < pv_fasta_n_seq(f); i++) v.push_back(pv_fasta_seq_name(f, i));
        return v;
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
    py::class_<type_sequence>(m, "type_sequence")                       // pybind_api.h:234-237
        .def_readwrite("sequence_length", &type_sequence::sequence_length)
        .def_readwrite("sequence_name", &type_sequence::sequence_name);

    py::class_<BAM_handler>(m, "BAM_handler")                           // pybind_api.h:223-232
        .def(py::init<const string&>())
        .def("get_chromosome_sequence_names", &BAM_handler::get_chromosome_sequence_names)
        .def("get_chromosome_sequence_names_with_length", &BAM_handler::get_chromosome_sequence_names_with_length)
        .def("get_sample_names", &BAM_handler::get_sample_names)
        .def("get_reads", &BAM_handler::get_reads);

    py::class_<FASTA_handler>(m, "FASTA_handler")                       // pybind_api.h:240-246
        .def(py::init<const string&>())
        .def("get_reference_sequence", &FASTA_handler::get_reference_sequence)
        .def("get_chromosome_sequence_length", &FASTA_handler::get_chromosome_sequence_length)
        .def("get_chromosome_names", &FASTA_handler::get_chromosome_names);

    // The legacy per-position generator (pybind_api.h:24-43; not called by call_variant) lives in Python on top of the
    // polisher's summary kernels: pepper_thesis_b200/legacy_summary.py. Forwarded lazily (the import pulls in torch).
    for (const char* name : {"SummaryGenerator", "ImageSummary"}) {
        string n = name;
        m.attr(name) = py::cpp_function([n](py::args a, py::kwargs k) -> py::object {
            return py::module_::import("pepper_thesis_b200.legacy_summary").attr(n.c_str())(*a, **k);
        });
    }
    out_of_scope(m, "RegionalSummaryGeneratorHP", "haplotype-aware variant, -hp flag only");
    out_of_scope(m, "CandidateFinder", "legacy C++ candidate finder, unreachable from call_variant");
    out_of_scope(m, "CandidateFinderHP", "legacy C++ candidate finder, unreachable from call_variant");
    out_of_scope(m, "type_truth_record", "training labels");
}
