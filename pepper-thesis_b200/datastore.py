"""The reference's HDF5 hand-off files, written without h5py (hdf5_lite.py): the image file of stage 1
(/root/reference/pepper_variant/modules/python/DataStore.py:54-71: ``summaries/<name>/{contigs 'S', positions int32, depths
uint8, candidates vlen str, candidate_frequency uint8, images int8}``) and the prediction file of stage 2
(DataStorePredict.py:49-66: ``predictions/batch_<n>/{contigs, positions, depths, candidates, candidate_frequency,
base_prediction float64}``). The B200 path itself hands tensors from stage to stage (pipeline.HotPath); these writers exist so
that a run can leave the files the reference's own stage 2 / stage 3 tools read. Same method names and argument order as the
reference classes; the int8 cast of the images (DataStore.py:68) wraps like numpy's."""
from __future__ import annotations

import numpy as np

from . import hdf5_lite


class _Store:
    def __init__(self, filename, mode="w"):
        if mode not in ("w", "x"):
            raise ValueError("only writing is supported (mode 'w'); read the files with h5py or hdf5_lite.Reader")
        self.filename, self.mode = filename, mode
        self.file_handler = None
        self._names = set()

    def __enter__(self):
        self.file_handler = hdf5_lite.Writer(self.filename)
        return self

    def __exit__(self, *args):
        self.close()

    def close(self):
        if self.file_handler is not None:
            self.file_handler.close()
            self.file_handler = None

    def _put(self, path, name, contigs, positions, depths, candidates, candidate_frequency):
        f = self.file_handler
        f["%s/%s/contigs" % (path, name)] = np.array(contigs, dtype="S")
        f["%s/%s/positions" % (path, name)] = np.array(positions, dtype=np.int32)
        f["%s/%s/depths" % (path, name)] = np.array(depths, dtype=np.int64).astype(np.uint8)
        f["%s/%s/candidates" % (path, name)] = hdf5_lite.VlenStr(candidates)
        f["%s/%s/candidate_frequency" % (path, name)] = np.array(candidate_frequency, dtype=np.int64).astype(np.uint8)


class DataStore(_Store):
    """DataStore.py:7-71 (writing side)"""
    _summary_path_ = "summaries"

    def write_summary(self, summary_name, contigs, positions, depths, all_candidates, all_candidate_frequency, all_images,
                      all_base_labels=None, all_type_label=None, train_mode=False):
        if self.file_handler is None:
            self.__enter__()
        if summary_name in self._names:
            return
        self._names.add(summary_name)
        self._put(self._summary_path_, summary_name, contigs, positions, depths, all_candidates, all_candidate_frequency)
        images = np.asarray(all_images)
        self.file_handler["%s/%s/images" % (self._summary_path_, summary_name)] = images.astype(np.int64).astype(np.int8)   # :68
        if train_mode:
            self.file_handler["%s/%s/base_labels" % (self._summary_path_, summary_name)] = np.array(all_base_labels, dtype=np.uint8)
            self.file_handler["%s/%s/type_label" % (self._summary_path_, summary_name)] = np.array(all_type_label, dtype=np.uint8)


class DataStorePredict(_Store):
    """DataStorePredict.py:6-66 (writing side)"""
    _prediction_path_ = "predictions"

    def write_prediction(self, batch_no, contigs, positions, depths, candidates, candidate_frequencies, base_predictions):
        if self.file_handler is None:
            self.__enter__()
        name = "batch_" + str(batch_no)
        if name in self._names:
            return
        self._names.add(name)
        self._put(self._prediction_path_, name, contigs, positions, depths, candidates, candidate_frequencies)
        self.file_handler["%s/%s/base_prediction" % (self._prediction_path_, name)] = np.array(base_predictions, dtype=np.float64)   # :66


def write_prediction_batch(store: DataStorePredict, batch_no: int, pred, contigs_of_region) -> None:
    """One ``pipeline.Predictions`` as one ``predictions/batch_<n>`` group: what predict_distributed_gpu.py:82-93 hands to
    ``write_prediction`` per batch -- contig, position, depth, [candidate allele], [its frequency], the three genotype
    probabilities -- for every candidate."""
    contigs = [contigs_of_region[int(r)] for r in pred.region]
    # depths and frequencies go through the reference's uint8 casts as they are (DataStorePredict.py:63,65: a frequency above 255
    # wraps in the file, exactly as it does there; depths are at most 125, region_summary.cpp:682)
    store.write_prediction(batch_no, contigs, pred.position, pred.depth, [[a.decode("latin-1")] for a in pred.alleles()],
                           [[int(f)] for f in pred.frequency], pred.probs.astype(np.float64))


def write_summary_from_workspace(store: DataStore, summary_name: str, ws, k: int, contigs_of_region) -> None:
    """The first ``k`` candidates of a summary workspace (device.SummaryWorkspace after pv_summary_regions) as one
    ``summaries/<name>`` group: what ImageGenerationUI.py:236-256 hands to ``write_summary`` -- the windows go to disk as int8
    (DataStore.py:68: values beyond [-128, 127] wrap, which the ``wrap_int8`` flag of the inference kernels reproduces)."""
    region = ws.region[:k].cpu().numpy()
    alleles = ws.allele[:k].cpu().numpy()
    lens = ws.allele_len[:k].cpu().numpy()
    store.write_summary(summary_name, [contigs_of_region[int(r)] for r in region], ws.position[:k].cpu().numpy(),
                        ws.depth[:k].cpu().numpy(),
                        [[bytes(alleles[i, :lens[i]]).decode("latin-1")] for i in range(k)],
                        [[int(f)] for f in ws.frequency[:k].cpu().numpy()], ws.windows[:k].cpu().numpy())


def predict_hdf5(image_hdf: str, output_hdf: str, model, device="cuda", pass_windows: int = 32768) -> int:
    """Stage 2 of ``pepper_variant call_variant`` as the reference runs it from files (predict_distributed_gpu.py:48-96 +
    dataloader_predict.py:82-93): every ``summaries/<name>`` group of an image file -- written by the reference's DataStore
    through h5py in the classic layout, or by ``DataStore`` above -- goes through the tcgen05 LSTM model, the genotype
    probabilities leave as ``predictions/batch_<n>`` groups (one per pass instead of one per 512 windows). The int8 values of
    the file are what the network sees, exactly like the reference's ``images.type(torch.FloatTensor)``. Returns the number of
    candidates predicted."""
    import torch
    rd = hdf5_lite.Reader(image_hdf)
    names = rd.keys(DataStore._summary_path_) if DataStore._summary_path_ in rd.keys("/") else []
    total, batch_no = 0, 0
    with DataStorePredict(output_hdf, "w") as out:
        for name in names:
            g = "%s/%s/" % (DataStore._summary_path_, name)
            images = rd[g + "images"]
            contigs, positions, depths = rd[g + "contigs"], rd[g + "positions"], rd[g + "depths"]
            cands, freqs = rd[g + "candidates"], rd[g + "candidate_frequency"]
            for lo in range(0, images.shape[0], pass_windows):
                hi = min(images.shape[0], lo + pass_windows)
                x = torch.from_numpy(np.ascontiguousarray(images[lo:hi]).astype(np.int16)).to(device)
                probs, _ = model.infer_windows(x, wrap_int8=False)          # the file already holds the wrapped values
                out.write_prediction(batch_no, [c.decode("latin-1") for c in contigs[lo:hi].tolist()], positions[lo:hi],
                                     depths[lo:hi], cands[lo:hi].tolist(), freqs[lo:hi].tolist(), probs.cpu().numpy())
                batch_no += 1
                total += hi - lo
    return total


def find_candidates_hdf5(prediction_hdf: str, fasta_path: str, options=None):
    """Stage 3 from the prediction FILE (FindCandidates.py:150-184 -> CandidateFinder.find_candidates, :532-581): every
    ``predictions/batch_<n>`` group of the file and the FASTA go through the device filter (candidate_filter.cu); returns
    ``(contigs, phasing_dict, variant_calling_dict)`` like the reference. The reference context of a contig's candidates is
    fetched once per contig (FASTA_handler.get_reference_sequence) instead of once per candidate."""
    from . import candidate_filter, ingest
    from .pipeline import Predictions
    rd = hdf5_lite.Reader(prediction_hdf)
    path = DataStorePredict._prediction_path_
    batches = sorted(rd.keys(path), key=lambda s: int(s.split("_")[1])) if path in rd.keys("/") else []
    if not batches:
        return [], {}, {}
    col = {k: [] for k in ("contigs", "positions", "depths", "candidates", "candidate_frequency", "base_prediction")}
    for b in batches:
        for k in col:
            col[k].append(rd["%s/%s/%s" % (path, b, k)])
    contigs = [c.decode("latin-1") for a in col["contigs"] for c in a.tolist()]
    pos = np.concatenate(col["positions"]).astype(np.int64)
    depth = np.concatenate(col["depths"]).astype(np.int32)
    freq = np.concatenate([a.reshape(a.shape[0], -1)[:, 0] for a in col["candidate_frequency"]]).astype(np.int32)
    probs = np.concatenate(col["base_prediction"]).astype(np.float32)
    cands = [row[0] if isinstance(row, (list, np.ndarray)) else row for a in col["candidates"] for row in a.tolist()]
    n = len(pos)
    allele = np.zeros((n, 64), np.uint8)
    allele_len = np.zeros(n, np.uint8)
    for i, c in enumerate(cands):
        b = c.encode("latin-1")[:64]
        allele[i, :len(b)] = np.frombuffer(b, np.uint8)
        allele_len[i] = len(b)
    names = list(dict.fromkeys(contigs))
    index = {c: i for i, c in enumerate(names)}
    region = np.array([index[c] for c in contigs], np.int32)
    fasta = ingest.FASTAHandler(fasta_path)

    class _View:
        pass
    v = _View()
    v.n_regions, v.contigs = len(names), names
    starts, offs, lens, clen, refs = [], [], [], [], []
    at = 0
    for i, c in enumerate(names):
        p = pos[region == i]
        total = int(fasta.get_chromosome_sequence_length(c))
        lo, hi = max(0, int(p.min()) - 32), min(total, int(p.max()) + 128)
        seq = fasta.get_reference_sequence(c, lo, hi).encode("latin-1")
        starts.append(lo); offs.append(at); lens.append(len(seq)); clen.append(total); refs.append(np.frombuffer(seq, np.uint8))
        at += len(seq)
    v.region_ref_start, v.region_ref_off = np.array(starts, np.int64), np.array(offs, np.int64)
    v.region_ref_len, v.region_contig_len = np.array(lens, np.int64), np.array(clen, np.int64)
    v.ref = np.ascontiguousarray(np.concatenate(refs))
    pred = Predictions(region, pos, depth, freq, allele, allele_len, probs, probs.argmax(1).astype(np.uint8))
    return candidate_filter.find_candidates(pred, v, options or candidate_filter.FilterOptions())
