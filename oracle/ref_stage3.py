"""TEST INFRASTRUCTURE ONLY -- runs the UNMODIFIED stage-3 modules of the reference
(/root/reference/pepper_variant/modules/python/CandidateFinder.py, VcfWriter.py) in this container.

Both files fail to import here only because of their I/O dependencies (h5py, pysam and the compiled
``pepper_variant.build.PEPPER_VARIANT``). This module injects small in-memory stand-ins for exactly those three before
importing the reference files from where they lie; none of the reference's own logic is touched:

  h5py.File(name)[...]                       -> arrays registered with ``register_predictions`` (dtypes as
                                                DataStorePredict.write_prediction stores them: DataStorePredict.py:49-66)
  PEPPER_VARIANT.FASTA_handler               -> in-memory contigs (fetch clipped to the contig, upper-cased like fasta_handler.cpp:49)
  PEPPER_VARIANT.CandidateImagePrediction    -> a plain record (prediction_base narrowed to float32 like the C++ vector<float>)
  pysam.VariantFile / VariantHeader          -> a recorder that keeps every new_record(**kwargs) and the files it was written to

Used by tests/golden/make_stage3_golden.py (writes the committed golden vectors) and by the CPU tests when
/root/reference is present. Only tests/ may import this file.
"""
import importlib
import os
import sys
import types

import numpy as np

REF_ROOT = "/root/reference"
_H5 = {}          # file name -> {'predictions': {batch_key: {dataset: array}}}
_FASTA = {}       # fasta path -> [(contig, sequence), ...]
_loaded = {}


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "pepper_variant/modules/python/CandidateFinder.py"))


# ---- stand-ins ------------------------------------------------------------------------------------------------------
class _Dataset:
    def __init__(self, a):
        self.a = a

    def __getitem__(self, key):
        return self.a if key == () else self.a[key]


class _Group(dict):
    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        return _Group(v) if isinstance(v, dict) else _Dataset(v)


class _H5File:
    def __init__(self, name, mode="r"):
        self.g = _Group(_H5[name])

    def __enter__(self):
        return self.g

    def __exit__(self, *a):
        return False


class _FastaHandler:
    def __init__(self, path):
        self.contigs = _FASTA[path]

    def get_reference_sequence(self, contig, start, end):
        seq = dict(self.contigs)[contig]
        return seq[max(0, start):max(0, end)].upper()

    def get_chromosome_names(self):
        return [c for c, _ in self.contigs]

    def get_chromosome_sequence_length(self, contig):
        return len(dict(self.contigs)[contig])


class _CandidateImagePrediction:
    def __init__(self, contig, position, depth, candidates, candidate_frequency, prediction_base, prediction_type):
        self.contig, self.position, self.depth = contig, int(position), int(depth)
        self.candidates, self.candidate_frequency = list(candidates), list(candidate_frequency)
        self.prediction_base = [float(np.float32(v)) for v in prediction_base]       # std::vector<float>
        self.prediction_type = list(prediction_type)


class _Header:
    def __init__(self):
        self.meta, self.samples = [], []
        self.contigs = self
        self.contig_list = []

    def add_meta(self, key, items):
        self.meta.append((key, list(items)))

    def add(self, name, length=None):
        self.contig_list.append((name, length))

    def add_sample(self, name):
        self.samples.append(name)


class _VariantFile:
    opened = []

    def __init__(self, path, mode, header=None):
        self.path, self.header, self.records, self.closed = path, header, [], False
        _VariantFile.opened.append(self)

    def new_record(self, **kw):
        return dict(kw)

    def write(self, rec):
        self.records.append(rec)

    def close(self):
        self.closed = True


def _install():
    if "h5py" not in sys.modules or getattr(sys.modules["h5py"], "_pv_standin", False):
        h5 = types.ModuleType("h5py"); h5._pv_standin = True
        h5.File = _H5File
        sys.modules["h5py"] = h5
    if "pysam" not in sys.modules or getattr(sys.modules["pysam"], "_pv_standin", False):
        ps = types.ModuleType("pysam"); ps._pv_standin = True
        ps.VariantFile, ps.VariantHeader = _VariantFile, _Header
        ps.tabix_index = lambda *a, **k: None
        sys.modules["pysam"] = ps
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import pepper_variant                                     # the reference package itself (pure Python at this level)
    build = types.ModuleType("pepper_variant.build")
    pv = types.ModuleType("pepper_variant.build.PEPPER_VARIANT")
    pv.FASTA_handler, pv.CandidateImagePrediction = _FastaHandler, _CandidateImagePrediction
    build.PEPPER_VARIANT = pv
    sys.modules["pepper_variant.build"] = build
    sys.modules["pepper_variant.build.PEPPER_VARIANT"] = pv
    pepper_variant.build = build


def candidate_finder():
    """The imported, unmodified CandidateFinder module."""
    if "cf" not in _loaded:
        _install()
        _loaded["cf"] = importlib.import_module("pepper_variant.modules.python.CandidateFinder")
    return _loaded["cf"]


def vcf_writer():
    if "vw" not in _loaded:
        _install()
        _loaded["vw"] = importlib.import_module("pepper_variant.modules.python.VcfWriter")
    return _loaded["vw"]


# ---- drivers --------------------------------------------------------------------------------------------------------
def register_predictions(file_name, batches):
    """batches: {batch_key: list of (contig, position, depth, [allele], [frequency], probs)} stored with the dtypes of
    DataStorePredict.write_prediction (contigs 'S', positions int32, depths uint8, candidates as str objects (h5py 2.x
    returns vlen strings as str), candidate_frequency uint8, base_prediction float64)."""
    g = {}
    for key, cands in batches.items():
        cand_arr = np.empty((len(cands), 1), dtype=object)
        for i, c in enumerate(cands):
            cand_arr[i, 0] = c[3][0]
        g[key] = {"contigs": np.array([c[0] for c in cands], dtype="S"),
                  "positions": np.array([c[1] for c in cands], dtype=np.int32),
                  "depths": np.array([c[2] for c in cands], dtype=np.uint8),
                  "candidates": cand_arr,
                  "candidate_frequency": np.array([c[4] for c in cands], dtype=np.uint8),
                  "base_prediction": np.array([np.asarray(c[5], dtype=np.float32) for c in cands], dtype=np.float64)}
    _H5[file_name] = {"predictions": g}


def register_fasta(path, contigs):
    _FASTA[path] = list(contigs)


class Options:
    def __init__(self, fasta, threads=1, **kw):
        self.fasta, self.threads = fasta, threads
        for k, v in kw.items():
            setattr(self, k, v)


FILTER_FIELDS = ("snp_p_value", "snp_p_value_in_lc", "insert_p_value", "insert_p_value_in_lc", "delete_p_value",
                 "delete_p_value_in_lc", "report_snp_above_freq", "report_indel_above_freq")


def ref_find_candidates(cands, contigs, opt_values, batch_size=128):
    """Unmodified find_candidates (CandidateFinder.py:532-581, which runs small_chunk_stitch :356-529 in a process pool)
    over in-memory predictions. -> (contigs, phasing dict, variant-calling dict)."""
    cf = candidate_finder()
    register_fasta("mem.fa", contigs)
    batches = {"batch_%d" % (i // batch_size): cands[i:i + batch_size] for i in range(0, len(cands), batch_size)}
    register_predictions("mem.hdf", batches)
    options = Options("mem.fa", threads=1, **dict(zip(FILTER_FIELDS, opt_values)))
    pairs = [("mem.hdf", k) for k in batches]
    return cf.find_candidates(options, "", pairs)


def ref_stitch(cands, contigs, opt_values):
    """Unmodified small_chunk_stitch on one in-memory batch -> (margin list, deepvariant list)."""
    cf = candidate_finder()
    register_fasta("mem.fa", contigs)
    register_predictions("mem1.hdf", {"b": cands})
    return cf.small_chunk_stitch(Options("mem.fa", **dict(zip(FILTER_FIELDS, opt_values))), [("mem1.hdf", "b")])


VCF_FIELDS = ("allowed_multiallelics", "snp_q_cutoff", "indel_q_cutoff", "snp_q_cutoff_in_lc", "indel_q_cutoff_in_lc")


def ref_vcf(sites, opt_values, contigs, sample="HG002"):
    """Unmodified VCFWriter.write_vcf_records (VcfWriter.py:141-221, with candidate_list_to_variant :48-139) into the
    recorder. -> (the five counts, records in the order they were written to the full file, each with the set of files it
    went to, header)."""
    vw = vcf_writer()
    register_fasta("vcf.fa", contigs)
    _VariantFile.opened = []
    w = vw.VCFWriter([c for c, _ in contigs], "vcf.fa", sample, "out/", "FULL", "PEPPER", "VC")
    options = Options("vcf.fa", **dict(zip(VCF_FIELDS, opt_values)))
    counts = w.write_vcf_records(sites, options)
    names = {"out/FULL.vcf.gz": "full", "out/PEPPER.vcf.gz": "pepper", "out/VC.vcf.gz": "variant_calling",
             "out/VC_SNPs.vcf.gz": "variant_calling_snp", "out/VC_INDEL.vcf.gz": "variant_calling_indel"}
    files = {names[f.path]: f for f in _VariantFile.opened}
    recs = []
    for r in files["full"].records:
        where = sorted(k for k, f in files.items() if any(x is r for x in f.records))
        recs.append(dict(r, files=where))
    header = files["full"].header
    return tuple(counts), recs, {"meta": header.meta, "contigs": header.contig_list, "samples": header.samples}
