"""hdf5_lite (the HDF5 writer / reader behind datastore.py). h5py / libhdf5 are absent here, so the anchor
is a file the real library wrote: scipy ships one (a MATLAB 7.3 file). The reader has to parse it; the writer's files have
to come back through the same reader, structure by structure; the header messages the writer emits for a float64 dataset
are compared byte for byte with the ones the library wrote. No GPU."""
import os
import struct

import numpy as np
import pytest

from pepper_thesis_b200 import datastore, hdf5_lite as H


def _genuine():
    try:
        import scipy.io.matlab
    except Exception:
        return None
    p = os.path.join(os.path.dirname(scipy.io.matlab.__file__), "tests", "data", "testhdf5_7.4_GLNX86.mat")
    return p if os.path.exists(p) else None


@pytest.mark.skipif(_genuine() is None, reason="scipy's HDF5 test file is not installed")
def test_reader_parses_a_file_written_by_the_hdf5_library():
    r = H.Reader(_genuine())
    assert r.base == 512 and r.keys("/") == ["testdouble"] and not r.is_group("testdouble")
    v = r["testdouble"]
    assert v.shape == (9, 1) and v.dtype == np.float64
    assert np.allclose(v[:, 0], np.arange(9) * np.pi / 4)


@pytest.mark.skipif(_genuine() is None, reason="scipy's HDF5 test file is not installed")
def test_writer_emits_the_librarys_own_header_messages(tmp_path):
    """dataspace, datatype and fill-value messages of a float64 [9][1] dataset: byte for byte what the library wrote; the
    group structures (root header, heap with its free block, B-tree keys) have the same shape"""
    g = H.Reader(_genuine())
    want = {m[0]: bytes(m[1]) for m in g._messages(g.root["testdouble"])}
    p = str(tmp_path / "a.h5")
    with H.Writer(p) as w:
        w["testdouble"] = (np.arange(9) * np.pi / 4).reshape(9, 1)
    r = H.Reader(p)
    got = {m[0]: bytes(m[1]) for m in r._messages(r.root["testdouble"])}
    for mtype in (H.MSG_DATASPACE, H.MSG_DATATYPE, H.MSG_FILL):
        assert got[mtype] == want[mtype], hex(mtype)
    assert np.array_equal(r["testdouble"], g["testdouble"])
    # root group: one symbol-table message; heap = 8 zero bytes, the names, a free block whose "next" is 1
    for rd in (g, r):
        msgs = rd._messages(rd.root_header)
        assert msgs[0][0] == H.MSG_SYMBOL_TABLE and len(msgs[0][1]) == 16
        bt, hp = struct.unpack_from("<QQ", msgs[0][1], 0)
        seg = rd._heap(hp)
        assert seg[:8] == bytes(8) and seg[8:19] == b"testdouble\0"
        q = rd._at(bt)
        assert rd.d[q:q + 8] == b"TREE\0\0\1\0" and struct.unpack_from("<QQQ", rd.d, q + 24)[::2] == (0, 8)   # keys "", "testdouble"


def test_round_trip_of_every_type_and_of_large_groups(tmp_path):
    rng = np.random.default_rng(0)
    p = str(tmp_path / "b.h5")
    want = {}
    with H.Writer(p) as w:
        for i in range(300):                                   # 300 members: 38 symbol-table nodes under a two-level B-tree
            name = "summaries/chr20_%d_%d/positions" % (i * 1000, i * 1000 + 999)
            want[name] = rng.integers(-2**31, 2**31 - 1, rng.integers(0, 50)).astype(np.int32)
            w[name] = want[name]
        want["t/i8"] = rng.integers(-128, 128, (5, 33, 26)).astype(np.int8)
        want["t/u8"] = rng.integers(0, 256, 77).astype(np.uint8)
        want["t/i64"] = rng.integers(-2**62, 2**62, (3, 2)).astype(np.int64)
        want["t/f64"] = rng.standard_normal((4, 3))
        want["t/f32"] = rng.standard_normal(6).astype(np.float32)
        want["t/S"] = np.array(["chr20", "chrX", ""], dtype="S")
        for k in ("t/i8", "t/u8", "t/i64", "t/f64", "t/f32", "t/S"):
            w[k] = want[k]
        strs = [["1" + "ACGT"[i % 4] * (i % 70) if i % 50 else ""] for i in range(700)]     # ~30 KB of strings: several global heap
        # collections; the empty ones are null heap ids
        w["t/vlen"] = H.VlenStr(strs)
        w["meta/yaml"] = "summaries: !!set {a: null}\n"
        w["t/empty"] = np.zeros((0, 33, 26), np.int8)
    r = H.Reader(p)
    assert os.path.getsize(p) == r.eof and r.keys("/") == ["meta", "summaries", "t"]
    assert len(r.keys("summaries")) == 300 and r.is_group("summaries/chr20_0_999")
    for k, v in want.items():
        got = r[k]
        assert got.dtype == v.dtype and got.shape == v.shape and np.array_equal(got, v), k
    v = r["t/vlen"]
    assert v.shape == (700, 1) and v.tolist() == strs
    assert r["meta/yaml"] == "summaries: !!set {a: null}\n"
    assert r["t/empty"].shape == (0, 33, 26)
    d = r.describe("t/vlen")
    assert d["type_class"] == 9 and d["type_size"] == 16 and d["type_bits"][:2] == (0x01, 0x01)
    assert d["type_props"][:8] == bytes([0x13, 0, 0, 0, 1, 0, 0, 0])          # base type: H5T_C_S1 (1 byte, null terminated, ASCII)
    assert r.describe("t/i8")["type_bits"][0] & 8 and not r.describe("t/u8")["type_bits"][0] & 8
    # every structure starts 8-byte aligned and lies inside the file
    raw = open(p, "rb").read()
    for tag in (b"TREE", b"HEAP", b"SNOD", b"GCOL"):
        at = raw.find(tag)
        assert at > 0 and at % 8 == 0


def test_duplicate_names_and_unsupported_types_are_refused(tmp_path):
    w = H.Writer(str(tmp_path / "c.h5"))
    w["a/b"] = np.zeros(3, np.uint8)
    with pytest.raises(ValueError):
        w["a/b"] = np.zeros(3, np.uint8)
    with pytest.raises(ValueError):
        w["a/b/c"] = np.zeros(3, np.uint8)
    with pytest.raises(TypeError):
        w["a/c"] = np.zeros(3, np.complex64)
    w.close()
    with pytest.raises(H.FormatError):
        open(str(tmp_path / "bad.h5"), "wb").write(b"not hdf5" * 100) and H.Reader(str(tmp_path / "bad.h5"))


def test_datastore_files_have_the_reference_schema(tmp_path):
    """DataStore.py:54-71 / DataStorePredict.py:49-66: names, dtypes, the int8 wrap of the images"""
    n = 12
    contigs = ["chr20"] * n
    positions = list(range(1000, 1000 + n))
    depths = [50 + i for i in range(n)]
    cands = [["1A"], ["2ACGT"], ["3AC"]] * 4
    freqs = [[7]] * n
    images = np.arange(n * 33 * 26).reshape(n, 33, 26) - 5000          # values beyond int8: wrapped like np.int8
    p = str(tmp_path / "images.hdf")
    with datastore.DataStore(p, "w") as ds:
        ds.write_summary("chr20_1000_1012", contigs, positions, depths, cands, freqs, images, None, None, False)
        ds.write_summary("chr20_1000_1012", contigs, positions, depths, cands, freqs, images, None, None, False)   # second call: ignored
    r = H.Reader(p)
    g = "summaries/chr20_1000_1012/"
    assert r.keys("summaries/chr20_1000_1012") == ["candidate_frequency", "candidates", "contigs", "depths", "images", "positions"]
    assert r[g + "contigs"].dtype == np.dtype("S5") and r[g + "contigs"].tolist() == [b"chr20"] * n
    assert r[g + "positions"].dtype == np.int32 and r[g + "positions"].tolist() == positions
    assert r[g + "depths"].dtype == np.uint8 and r[g + "candidate_frequency"].shape == (n, 1)
    assert r[g + "candidates"].tolist() == cands
    assert r[g + "images"].dtype == np.int8 and np.array_equal(r[g + "images"], images.astype(np.int8))
    q = str(tmp_path / "pred.hdf")
    with datastore.DataStorePredict(q, "w") as ds:
        ds.write_prediction(0, contigs, positions, depths, cands, freqs, np.full((n, 3), 1.0 / 3))
        ds.write_prediction(1, contigs[:2], positions[:2], depths[:2], cands[:2], freqs[:2], np.eye(3)[:2])
    r = H.Reader(q)
    assert r.keys("predictions") == ["batch_0", "batch_1"]
    assert r["predictions/batch_1/base_prediction"].dtype == np.float64 and r["predictions/batch_1/base_prediction"].tolist() == [[1, 0, 0], [0, 1, 0]]


def test_predictions_as_the_reference_stage2_file(tmp_path):
    from pepper_thesis_b200.pipeline import Predictions
    n = 40
    rng = np.random.default_rng(3)
    al = np.zeros((n, 64), np.uint8)
    lens = rng.integers(2, 9, n).astype(np.uint8)
    for i in range(n):
        al[i, :lens[i]] = np.frombuffer(("1" + "ACGT"[i % 4] * (int(lens[i]) - 1)).encode(), np.uint8)
    probs = rng.random((n, 3)).astype(np.float32)
    pred = Predictions(rng.integers(0, 2, n).astype(np.int32), np.sort(rng.integers(0, 10**6, n)).astype(np.int64),
                       rng.integers(10, 126, n).astype(np.int32), rng.integers(1, 60, n).astype(np.int32), al, lens, probs,
                       probs.argmax(1).astype(np.uint8))
    p = str(tmp_path / "pred.hdf")
    with datastore.DataStorePredict(p, "w") as ds:
        datastore.write_prediction_batch(ds, 0, pred, ["chr20", "chr21"])
    r = H.Reader(p)
    g = "predictions/batch_0/"
    assert r[g + "contigs"].tolist() == [b"chr20" if x == 0 else b"chr21" for x in pred.region]
    assert np.array_equal(r[g + "positions"], pred.position.astype(np.int32)) and np.array_equal(r[g + "depths"], pred.depth.astype(np.uint8))
    assert [x[0].encode("latin-1") for x in r[g + "candidates"].tolist()] == pred.alleles()
    assert np.array_equal(r[g + "base_prediction"], probs.astype(np.float64))


def test_reader_fails_cleanly_on_corrupted_files(tmp_path):
    """single-byte and truncation damage: every read either succeeds or raises an ordinary exception -- no hang (cyclic free
    lists, B-tree cycles), no crash"""
    p = str(tmp_path / "ok.h5")
    with H.Writer(p) as w:
        for i in range(20):
            w["g/sub%02d/x" % i] = np.arange(i + 1, dtype=np.int32)
        w["g/v"] = H.VlenStr([["1A"], [""], ["2ACGT"]])
        w["s"] = "text"
    raw = bytearray(open(p, "rb").read())
    rng = np.random.default_rng(5)

    def read_all(path):
        r = H.Reader(path)
        def walk(prefix, depth=0):
            for k in r.keys(prefix):
                q = prefix.rstrip("/") + "/" + k
                if r.is_group(q):
                    if depth < 4:
                        walk(q, depth + 1)
                else:
                    r[q]
        walk("/")
    read_all(p)
    survived = 0
    for trial in range(400):
        bad = bytearray(raw)
        if trial % 8 == 7:
            bad = bad[:int(rng.integers(8, len(bad)))]
        else:
            for _ in range(int(rng.integers(1, 4))):
                bad[int(rng.integers(0, len(bad)))] = int(rng.integers(0, 256))
        q = str(tmp_path / "bad.h5")
        open(q, "wb").write(bad)
        try:
            read_all(q)
            survived += 1
        except (H.FormatError, KeyError, ValueError, IndexError, struct.error, UnicodeDecodeError, OverflowError, MemoryError):
            pass
    assert survived < 400


def test_scalars_keep_rank_zero(tmp_path):
    p = str(tmp_path / "s.h5")
    with H.Writer(p) as w:
        w["f"] = np.float64(3.5); w["i"] = np.int32(-7); w["s"] = np.array(b"abc")
    r = H.Reader(p)
    assert r.describe("f")["shape"] == () and float(r["f"]) == 3.5 and int(r["i"]) == -7 and r["s"].item() == b"abc"
