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
    store.write_prediction(batch_no, contigs, pred.position, np.minimum(pred.depth, 255), [[a.decode("latin-1")] for a in pred.alleles()],
                           [[int(min(f, 255))] for f in pred.frequency], pred.probs.astype(np.float64))


def write_summary_from_workspace(store: DataStore, summary_name: str, ws, k: int, contigs_of_region) -> None:
    """The first ``k`` candidates of a summary workspace (device.SummaryWorkspace after pv_summary_regions) as one
    ``summaries/<name>`` group: what ImageGenerationUI.py:236-256 hands to ``write_summary`` -- the windows go to disk as int8
    (DataStore.py:68: values beyond [-128, 127] wrap, which the ``wrap_int8`` flag of the inference kernels reproduces)."""
    region = ws.region[:k].cpu().numpy()
    alleles = ws.allele[:k].cpu().numpy()
    lens = ws.allele_len[:k].cpu().numpy()
    store.write_summary(summary_name, [contigs_of_region[int(r)] for r in region], ws.position[:k].cpu().numpy(),
                        np.minimum(ws.depth[:k].cpu().numpy(), 255),
                        [[bytes(alleles[i, :lens[i]]).decode("latin-1")] for i in range(k)],
                        [[int(min(f, 255))] for f in ws.frequency[:k].cpu().numpy()], ws.windows[:k].cpu().numpy())


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
