"""GPU parity: the CUDA summary path (through the C-ABI, host-buffer entry point) against the CPU oracles."""
import numpy as np
import pytest

import helpers as H
import pyoracle as O
from pepper_thesis_b200 import capi, synth
from pepper_thesis_b200.read_batch import select_regions

pytestmark = pytest.mark.gpu


def oracle(b, r, thr):
    return O.ref_summary(b, r, thr) if O.have_ref() else O.port_summary(b, r, thr)


def gpu_region_dict(out, r):
    d = out.trimmed()
    m = d["region"] == r
    idx = np.nonzero(m)[0]
    return dict(position=d["position"][m], depth=d["depth"][m], frequency=d["frequency"][m], images=d["images"][m],
                alleles=[d["alleles"][i] for i in idx])


@pytest.mark.parametrize("name", sorted(H.KATS))
def test_kat(name):
    b = H.KATS[name]()
    out, dense = capi.summary_regions_host(b, H.R9, want_dense=True)
    p = O.port_summary(b, 0, H.R9, want_dense=True)
    assert np.array_equal(dense.astype(np.int32), p["dense"]), "dense image"
    H.assert_same(oracle(b, 0, H.R9), gpu_region_dict(out, 0), name)


@pytest.mark.parametrize("seed", range(60))
def test_fuzz(seed):
    b = H.fuzz_region(seed)
    thr = H.fuzz_thresholds(seed)
    out, dense = capi.summary_regions_host(b, thr, want_dense=True)
    p = O.port_summary(b, 0, thr, want_dense=True)
    bad = np.argwhere(dense.astype(np.int32) != p["dense"])
    assert bad.size == 0, "dense image differs first at %s" % bad[0].tolist()
    H.assert_same(oracle(b, 0, thr), gpu_region_dict(out, 0), "fuzz %d" % seed)


@pytest.mark.parametrize("seed", range(60, 70))
def test_fuzz_reference_n(seed):
    """Non-ACGT reference bases: the reference C++ is undefined there (vector[-1]); compare with the port."""
    b = H.fuzz_region(seed, ref_n=True)
    thr = H.fuzz_thresholds(seed)
    out = capi.summary_regions_host(b, thr)
    H.assert_same(O.port_summary(b, 0, thr), gpu_region_dict(out, 0), "fuzz-N %d" % seed)


@pytest.mark.parametrize("profile,cov", [("ont_r9", 30.0), ("ont_r9", 50.0), ("ont_r10", 40.0), ("hifi", 35.0)])
def test_synthetic_regions(profile, cov):
    """Several 100 kbp regions in one batch (incl. the first and last region of the contig)."""
    b = synth.generate(profile, 450000, cov, seed=5)
    thr = synth.PROFILES[profile].thresholds
    out, dense = capi.summary_regions_host(b, thr, want_dense=True)
    off = 0
    for r in range(b.n_regions):
        sub = select_regions(b, [r])
        p = O.port_summary(sub, 0, thr, want_dense=True)
        L = p["dense"].shape[0]
        assert np.array_equal(dense[off:off + L].astype(np.int32), p["dense"]), "dense region %d" % r
        off += L
        H.assert_same(oracle(sub, 0, thr), gpu_region_dict(out, r), "%s region %d" % (profile, r))
    assert out.count > 100


def test_many_fuzz_regions_one_batch():
    from pepper_thesis_b200.read_batch import ReadBatch
    bs = [H.fuzz_region(s, L=300 + 37 * (s % 5)) for s in range(100, 140)]
    import numpy as np
    from pepper_thesis_b200 import read_batch as rb
    # concatenate single-region batches
    parts = {n: [] for n in rb.ARRAY_NAMES}
    b_cur = o_cur = f_cur = r_cur = 0
    rbeg = [0]
    for b in bs:
        for n in ("read_pos", "read_len", "read_n_ops", "read_flags", "read_mapq", "bases", "quals", "cigar", "ref",
                  "region_ref_start", "region_ref_end", "region_cand_start", "region_cand_end", "region_ref_len"):
            parts[n].append(getattr(b, n))
        parts["read_base_off"].append(b.read_base_off + b_cur)
        parts["read_cigar_off"].append(b.read_cigar_off + o_cur)
        parts["region_ref_off"].append(b.region_ref_off + f_cur)
        b_cur += b.n_bases; o_cur += b.n_ops; f_cur += b.ref.shape[0]; r_cur += b.n_reads
        rbeg.append(r_cur)
    arrs = {n: np.ascontiguousarray(np.concatenate(parts[n])) for n in rb.ARRAY_NAMES if n != "region_read_begin"}
    big = ReadBatch(region_read_begin=np.asarray(rbeg, np.int64), contigs=["fz"] * len(bs), **arrs)
    out = capi.summary_regions_host(big, H.R9)
    for r, b in enumerate(bs):
        H.assert_same(oracle(b, 0, H.R9), gpu_region_dict(out, r), "batch region %d" % r)


def test_candidate_overflow_is_reported():
    b = H.kat_refskip()
    with pytest.raises(capi.PvError) as e:
        capi.summary_regions_host(b, H.R9, capacity=3)
    assert e.value.code == capi.PV_EOVERFLOW


def test_empty_region_and_mapq0():
    b = H.one_region("ACGT" * 25, [H.Read(0, "ACGT" * 25, [(0, 100)], mapq=0)])
    out = capi.summary_regions_host(b, H.R9)
    assert out.count == 0
    b = H.one_region("ACGT" * 25, [])
    assert capi.summary_regions_host(b, H.R9).count == 0


def test_bases4_wire_format_equals_byte_format():
    """BAM-native 4-bit bases uploaded and expanded on the device give the same candidates as byte bases."""
    b = synth.generate("ont_r9", 300000, 20.0, seed=8)
    thr = synth.PROFILES["ont_r9"].thresholds
    a = capi.summary_regions_host(b, thr).trimmed()
    b.pack_bases4()
    assert b.bases4.nbytes * 2 == b.bases.nbytes
    c = capi.summary_regions_host(b, thr).trimmed()
    H.assert_same(a, c, "bases4")
    assert np.array_equal(a["region"], c["region"])
    from pepper_thesis_b200 import device as dev
    import torch
    db = dev.DeviceBatch(b.region_range_view(1, 3))
    torch.cuda.synchronize()
    assert np.array_equal(db.t["bases"].cpu().numpy()[:100000], np.where(b.region_range_view(1, 3).bases[:100000] == 0, ord("="), b.region_range_view(1, 3).bases[:100000]))


def test_pack_bases4_rejects_other_bytes():
    b = H.fuzz_region(3)          # contains lower-case / odd bytes
    with pytest.raises(capi.PvError):
        b.pack_bases4()


@pytest.mark.parametrize("profile", ["ont_r9", "hifi"])
def test_compact_wire_formats_equal_plain(profile):
    """Bit-packed qualities + 16-bit CIGAR + 4-bit bases (host wire forms) give the same candidates as the plain arrays,
    through the host C-ABI entry point and through the device-resident pipeline path (views of region ranges)."""
    import torch
    from pepper_thesis_b200 import device as dev
    b = synth.generate(profile, 350000, 25.0, seed=8)
    thr = synth.PROFILES[profile].thresholds
    plain = capi.summary_regions_host(b, thr)
    b.pack_wire()
    assert b.quals_packed is not None and 1 <= b.qual_bits <= 6 and b.quals_packed.nbytes < b.quals.nbytes
    packed = capi.summary_regions_host(b, thr)
    for r in range(b.n_regions):
        H.assert_same(gpu_region_dict(plain, r), gpu_region_dict(packed, r), "wire %s region %d" % (profile, r))
    # a view that starts at a 16- but not 32-aligned base offset
    v = b.region_range_view(1, 3)
    db = dev.DeviceBatch(v)
    torch.cuda.synchronize()
    assert np.array_equal(db.t["quals"].cpu().numpy(), v.quals)
    assert np.array_equal(db.t["cigar"].cpu().numpy().view(np.uint32), v.cigar)
    real = np.zeros(v.n_bases, bool)
    for i in range(v.n_reads):
        real[int(v.read_base_off[i]):int(v.read_base_off[i]) + int(v.read_len[i])] = True
    assert np.array_equal(db.t["bases"].cpu().numpy()[real], v.bases[real])           # padding behind a read is don't-care
    assert v.bases2 is not None or v.bases_patch is not None
    assert db.h2d_bytes < sum(getattr(v, n).nbytes for n in ("bases", "quals", "cigar")) * 0.75


def test_pack_quals_roundtrip_all_widths():
    import torch
    rng = np.random.default_rng(0)
    lib = capi.load()
    for bits in range(1, 8):
        for n in (16, 32, 48, 4096 + 16):
            q = rng.integers(0, 1 << bits, n).astype(np.uint8)
            q[0] = (1 << bits) - 1
            assert lib.pv_qual_bits(q.ctypes.data, n, 3) == bits
            packed = np.zeros((n + 31) // 32 * bits * 4, np.uint8)
            capi.check(lib.pv_pack_quals(q.ctypes.data, n, bits, packed.ctypes.data, 3))
            d = torch.from_numpy(packed).cuda()
            out = torch.zeros(n + 16, dtype=torch.uint8, device="cuda")
            capi.check(lib.pv_unpack_quals(d.data_ptr(), n, bits, out.data_ptr(), None))
            torch.cuda.synchronize()
            assert np.array_equal(out[:n].cpu().numpy(), q) and int(out[n:].sum()) == 0
    big = np.array([4096 << 4], np.uint32)
    with pytest.raises(capi.PvError):
        capi.check(lib.pv_pack_cigar16(big.ctypes.data, 1, np.zeros(1, np.uint16).ctypes.data, 1))


def _assert_device_bases_equal(v):
    import torch
    from pepper_thesis_b200 import device as dev
    db = dev.DeviceBatch(v)
    torch.cuda.synchronize()
    got = db.t["bases"].cpu().numpy()
    for i in range(v.n_reads):
        o, n = int(v.read_base_off[i]), int(v.read_len[i])
        assert np.array_equal(got[o:o + n], v.bases[o:o + n]), "read %d" % i
    return db


def test_bases_ref_roundtrip():
    """Reference-predicted bases (pv_pack_bases_ref / pv_unpack_bases_ref): every read comes back byte for byte, for
    fuzzed CIGARs (soft clips, N, reads hanging over both region ends, q/len edge cases), odd bytes (N, lower case,
    '=', 255), region views (patch offsets re-based) and an empty patch list."""
    big = synth.generate("ont_r9", 350000, 12.0, seed=4)
    rng = np.random.default_rng(2)
    pos = rng.integers(0, big.n_bases, 800)
    big.bases[pos] = rng.choice(np.frombuffer(b"NnacgtRY=\xff", np.uint8), 800)
    batches = [H.fuzz_region(s) for s in range(8)] + [big]
    for batch in batches:
        batch.pack_bases_ref()
        assert batch.bases_patch is not None and int(batch.read_patch_off[-1]) == batch.bases_patch.size
        _assert_device_bases_equal(batch)
    for r0, r1 in ((0, 1), (1, 3), (2, 3)):
        _assert_device_bases_equal(big.region_range_view(r0, r1))
    # at ONT error rates the patch list is far smaller than 2 bits per base, and pack_wire picks it
    big.pack_wire(bases_ref=True)
    assert big.bases_patch is not None and big.bases2 is None
    assert big.bases_patch.nbytes + big.read_patch_off.nbytes < big.n_bases // 8
    # reads that equal their prediction exactly: no patch entries at all
    clean = synth.generate("hifi", 120000, 5.0, seed=6)
    clean.pack_bases_ref()
    _assert_device_bases_equal(clean)
    # and the summary of a batch that travelled this way is the summary of the plain batch
    from pepper_thesis_b200 import device as dev
    import torch
    plain = capi.summary_regions_host(big, H.R9)
    db = dev.DeviceBatch(big)
    ws = dev.SummaryWorkspace.for_batch(db, 1 << 15)
    dev.summary_regions(db, H.R9, ws)
    k = int(ws.count.item())
    want = plain.trimmed()
    assert k == plain.count and k > 0
    assert np.array_equal(ws.position[:k].cpu().numpy(), want["position"])
    assert np.array_equal(ws.windows[:k].cpu().numpy().astype(np.int32), np.asarray(want["images"]).astype(np.int32))


def test_cigar8_roundtrip():
    """8-bit CIGAR codes + escape stream (pv_pack_cigar8 / pv_unpack_cigar8): the device CIGAR equals the original for
    fuzzed op mixes (S, N, =, X, long and empty ops), the synthetic presets and region views."""
    import torch
    from pepper_thesis_b200 import device as dev
    big = synth.generate("ont_r9", 350000, 12.0, seed=4)
    for batch in [H.fuzz_region(s) for s in range(6)] + [big, synth.generate("hifi", 150000, 6.0, seed=2)]:
        batch.pack_cigar8()
        assert batch.cigar8 is not None and int(batch.read_esc_off[-1]) == batch.cigar_esc.size
        for v in ([batch] if batch is not big else [batch, big.region_range_view(1, 3), big.region_range_view(2, 3)]):
            db = dev.DeviceBatch(v)
            torch.cuda.synchronize()
            got = db.t["cigar"].cpu().numpy().view(np.uint32)
            for i in range(v.n_reads):
                o, n = int(v.read_cigar_off[i]), int(v.read_n_ops[i])
                assert np.array_equal(got[o:o + n], v.cigar[o:o + n]), "read %d" % i
    big.pack_wire()
    assert big.cigar8 is not None and big.cigar16 is None and big.bases_patch is not None
    db = _assert_device_bases_equal(big)            # the base prediction walks the CIGAR rebuilt from the codes
    assert db.h2d_bytes < 0.45 * sum(getattr(big, n).nbytes for n in ("bases", "quals", "cigar"))


def test_bases2_exceptions_roundtrip():
    """2-bit bases with every kind of odd byte (N, lower case, '=', 255) in the exception list; views re-base the list."""
    import torch
    from pepper_thesis_b200 import device as dev
    b = H.fuzz_region(3)                       # weird bytes in the reads
    big = synth.generate("ont_r9", 250000, 10.0, seed=2)
    rng = np.random.default_rng(1)
    pos = rng.integers(0, big.n_bases, 500)
    big.bases[pos] = rng.choice(np.frombuffer(b"NnacgtRY=\xff", np.uint8), 500)
    for batch in (b, big):
        plain = capi.summary_regions_host(batch, H.R9)
        batch.pack_bases2()
        assert batch.bases2 is not None and batch.base_exceptions.size > 0
        packed = capi.summary_regions_host(batch, H.R9)
        for r in range(batch.n_regions):
            H.assert_same(gpu_region_dict(plain, r), gpu_region_dict(packed, r), "bases2 region %d" % r)
    v = big.region_range_view(1, 2)
    db = dev.DeviceBatch(v)
    torch.cuda.synchronize()
    got = db.t["bases"].cpu().numpy()
    for i in range(v.n_reads):
        o, n = int(v.read_base_off[i]), int(v.read_len[i])
        assert np.array_equal(got[o:o + n], v.bases[o:o + n])


def _device_summary(batch, thr, cap=1 << 15):
    import torch
    from pepper_thesis_b200 import device as dev
    db = dev.DeviceBatch(batch)
    ws = dev.SummaryWorkspace.for_batch(db, cap)
    dev.summary_regions(db, thr, ws)
    k = int(ws.count.item())
    return db, {"k": k, "position": ws.position[:k].cpu().numpy(), "windows": ws.windows[:k].cpu().numpy().astype(np.int32),
                "depth": ws.depth[:k].cpu().numpy(), "frequency": ws.frequency[:k].cpu().numpy()}


def test_quals_pred_same_summary():
    """Quality predicates (pv_pack_quals_pred / pv_unpack_quals_pred): the batch travels without its qualities and the
    device summary is the summary of the plain batch -- fuzzed CIGARs (chained inserts, inserts behind deletions / clips,
    59..61-base inserts) under thresholds on both sides of the qualities, the synthetic preset under thresholds that fail
    half of the bases, region views, and the device array itself against the decoded patch list."""
    import torch
    from test_wire_cpu import _surrogate_quals
    from pepper_thesis_b200.synth import Thresholds
    for seed in range(10):
        thr = H.fuzz_thresholds(seed)
        if seed % 2:
            thr = Thresholds(*([float([3, 7.5, 20.5, 31, 25][seed // 2]), float([12.25, 20, 5.5, 16, 30.5][seed // 2])] + thr.as_list9()[2:] + [thr.skip_indels]))
        b = H.fuzz_region(seed, n_reads=60)
        want = capi.summary_regions_host(b, thr).trimmed()
        b.pack_quals_pred(thr.min_snp_baseq, thr.min_indel_baseq)
        assert b.quals_patch is not None
        db, got = _device_summary(b, thr)
        assert db.qpatches is not None
        sq = _surrogate_quals(b)
        dq = db.t["quals"].cpu().numpy()
        for i in range(b.n_reads):
            o, n = int(b.read_base_off[i]), int(b.read_len[i])
            assert np.array_equal(dq[o:o + n], sq[o:o + n]), "seed %d read %d" % (seed, i)
        assert got["k"] == len(want["position"]), "seed %d" % seed
        assert np.array_equal(got["position"], want["position"])
        assert np.array_equal(got["windows"], np.asarray(want["images"]).astype(np.int32).reshape(got["windows"].shape)), "seed %d" % seed
    big = synth.generate("ont_r9", 350000, 12.0, seed=4)
    for snp, indel in ((17.0, 18.5), (1.0, 1.0), (10.0, 29.0)):
        thr = Thresholds(*([snp, indel] + H.R9.as_list9()[2:] + [False]))
        plain = capi.summary_regions_host(big, thr).trimmed()
        b = synth.generate("ont_r9", 350000, 12.0, seed=4)
        b.pack_wire(quals_pred=(snp, indel))
        assert b.quals_patch is not None and b.quals_packed is None
        for v, sel in ((b, None), (b.region_range_view(1, 3), (1, 3))):
            db, got = _device_summary(v, thr)
            if sel is None:
                assert got["k"] == len(plain["position"]) and got["k"] > 0
                assert np.array_equal(got["position"], plain["position"])
                assert np.array_equal(got["windows"], np.asarray(plain["images"]).astype(np.int32).reshape(got["windows"].shape))
            else:
                pv = capi.summary_regions_host(big.region_range_view(*sel), thr).trimmed()
                assert got["k"] == len(pv["position"]) and np.array_equal(got["position"], pv["position"])
                assert np.array_equal(got["windows"], np.asarray(pv["images"]).astype(np.int32).reshape(got["windows"].shape))
        if snp == 1.0:
            assert b.quals_patch.size == 0 and db.h2d_bytes < 0.12 * sum(getattr(big, n).nbytes for n in ("bases", "quals", "cigar"))


def test_min_qual_promise_skips_qualities_without_changing_results():
    """PvReadBatch.min_qual: a batch whose smallest base quality clears both thresholds takes the tile kernel's
    no-quality path; candidates and windows equal the plain path (promise withheld) and the CPU oracle. A promise below
    a threshold changes nothing."""
    import pyoracle
    from pepper_thesis_b200.synth import Thresholds
    lib = capi.load()
    for seed in range(6):
        b = H.fuzz_region(seed, n_reads=60)
        b.quals = np.maximum(b.quals, 3).astype(np.uint8)          # padding too: only read bases are promised
        thr = H.fuzz_thresholds(seed)
        thr = Thresholds(*([2.0, 2.5] + thr.as_list9()[2:] + [thr.skip_indels]))
        want = capi.summary_regions_host(b, thr).trimmed()          # min_qual == 0: every quality is loaded and tested
        port = pyoracle.port_summary(b, 0, thr)
        H.assert_same(want, port, "plain vs oracle, seed %d" % seed)
        b.scan_min_qual(threads=2)
        assert b.min_qual == 3
        _, got = _device_summary(b, thr)
        assert got["k"] == len(want["position"]) and np.array_equal(got["position"], want["position"]), "seed %d" % seed
        assert np.array_equal(got["windows"], np.asarray(want["images"]).astype(np.int32).reshape(got["windows"].shape)), "seed %d" % seed
        hot = capi.summary_regions_host(b, thr).trimmed()           # host entry point with the promise
        H.assert_same(hot, want, "host entry, seed %d" % seed)
        thr2 = Thresholds(*([7.5, 2.5] + thr.as_list9()[2:] + [thr.skip_indels]))   # promise (3) below the SNP threshold
        H.assert_same(capi.summary_regions_host(b, thr2).trimmed(), pyoracle.port_summary(b, 0, thr2), "no fast path, seed %d" % seed)
    big = synth.generate("ont_r9", 350000, 12.0, seed=4)
    plain = capi.summary_regions_host(big, H.R9).trimmed()
    big.scan_min_qual()
    assert big.min_qual == 5
    for v in (big, big.region_range_view(1, 3)):
        _, got = _device_summary(v, H.R9)
        pv = plain if v is big else capi.summary_regions_host(synth.generate("ont_r9", 350000, 12.0, seed=4).region_range_view(1, 3), H.R9).trimmed()
        assert got["k"] == len(pv["position"]) and got["k"] > 0 and np.array_equal(got["position"], pv["position"])
        assert np.array_equal(got["windows"], np.asarray(pv["images"]).astype(np.int32).reshape(got["windows"].shape))


def test_quality_array_not_uploaded_when_the_promise_clears_the_thresholds():
    """DeviceBatch(skip_quals=True) / the host entry point: a batch whose min_qual promise clears both thresholds travels
    WITHOUT its quality array (PvReadBatch.quals == NULL on the device) and gives the candidates of the plain batch. A
    read whose CIGAR runs over its own end -- the one case the promise cannot decide -- raises status bit 4 and the host
    entry point repeats the call with the qualities; a promise that does not clear the thresholds keeps the upload."""
    import torch
    from pepper_thesis_b200 import device as dev
    from pepper_thesis_b200.synth import Thresholds
    for seed in range(4):
        b = H.fuzz_region(100 + seed, n_reads=60)
        b.quals = np.maximum(b.quals, 3).astype(np.uint8)
        thr = H.fuzz_thresholds(seed)
        thr = Thresholds(*([2.0, 2.5] + thr.as_list9()[2:] + [thr.skip_indels]))
        want = capi.summary_regions_host(b, thr).trimmed()          # no promise: qualities uploaded and tested
        b.scan_min_qual(threads=2)
        assert dev.quals_not_needed(b.min_qual, thr)
        db = dev.DeviceBatch(b, skip_quals=True)
        assert db.quals_skipped and db.struct.quals in (0, None) and db.h2d_bytes == dev.DeviceBatch(b).h2d_bytes - b.quals.nbytes
        ws = dev.SummaryWorkspace.for_batch(db, 1 << 15)
        dev.summary_regions(db, thr, ws)
        k = int(ws.count.item())
        assert ws.status() == 0
        assert k == len(want["position"]) and np.array_equal(ws.position[:k].cpu().numpy(), want["position"]), "seed %d" % seed
        assert np.array_equal(ws.windows[:k].cpu().numpy().astype(np.int32),
                              np.asarray(want["images"]).astype(np.int32).reshape(k, 33, 26)), "seed %d" % seed
        H.assert_same(capi.summary_regions_host(b, thr).trimmed(), want, "host entry without qualities, seed %d" % seed)
        assert not dev.quals_not_needed(b.min_qual, Thresholds(*([7.5, 2.5] + thr.as_list9()[2:] + [thr.skip_indels])))
    # an insert that runs over the end of its read: 10M 5I with 12 bases. Qualities: the two inserted bases that exist are
    # low, so the plain path fails the insert's quality test where a blind "promise" pass would count it.
    ref = (b"ACGT" * 30)
    reads = [H.Read(5, ref[5:15].decode() + "TT", [(0, 10), (1, 5)], rev=bool(i & 1), q=[30] * 10 + [3, 3]) for i in range(6)]
    reads += [H.Read(0, ref[0:60].decode(), [(0, 60)], rev=bool(i & 1), q=[30] * 60) for i in range(4)]
    b = H.one_region(ref, reads)
    thr = Thresholds(2.0, 20.0, 0.1, 0.15, 0.15, 1.0, 0.1, 0.1, 1.0, False)
    want = capi.summary_regions_host(b, thr).trimmed()
    H.assert_same(want, O.port_summary(b, 0, thr), "plain vs port")
    b.min_qual = 3
    thr_ok = Thresholds(2.0, 2.5, 0.1, 0.15, 0.15, 1.0, 0.1, 0.1, 1.0, False)
    want_ok = O.port_summary(b, 0, thr_ok)
    db = dev.DeviceBatch(b, skip_quals=True)
    ws = dev.SummaryWorkspace.for_batch(db, 4096)
    dev.summary_regions(db, thr_ok, ws)
    torch.cuda.synchronize()
    assert ws.status() & 16, "the kernels must flag the missing quality"
    db.ensure_quals()
    dev.summary_regions(db, thr_ok, ws)
    assert ws.status() == 0 and int(ws.count.item()) == len(want_ok["position"])
    H.assert_same(capi.summary_regions_host(b, thr_ok).trimmed(), want_ok, "host entry retries with qualities")
