"""Host side of the inline transport packing (pv_pack_group, csrc/host_pack.cpp; pipeline._GroupPacker): the 2-bit bases +
exception list and the 16-bit CIGAR of a group equal a numpy restatement of the formats documented in
include/pepper_b200.h (PvReadBatch.bases2 / cigar16) and the older per-array packers (no GPU)."""
import ctypes as C

import numpy as np
import pytest

from pepper_thesis_b200 import capi, pipeline, synth


def _pack_group(bases, cigar, threads, cap=1 << 18, misalign=(17, 3)):
    lib = capi.load()
    n = bases.size
    out2 = np.full(n // 4 + 64, 0xEE, np.uint8)[misalign[0]:misalign[0] + n // 4]
    out16 = np.zeros(cigar.size + 8, np.uint16)[misalign[1]:misalign[1] + cigar.size]
    exc = np.zeros(max(cap, 1), np.uint64)
    ne, fits = C.c_int64(0), C.c_int32(-1)
    capi.check(lib.pv_pack_group(C.c_void_p(bases.ctypes.data), C.c_int64(n), C.c_void_p(out2.ctypes.data), C.c_void_p(exc.ctypes.data),
                                 C.c_int64(cap), C.byref(ne), C.c_void_p(cigar.ctypes.data), C.c_int64(cigar.size),
                                 C.c_void_p(out16.ctypes.data), C.byref(fits), C.c_int32(threads)))
    return out2, exc[:min(ne.value, cap)], ne.value, out16, fits.value


def _numpy_forms(bases, cigar):
    code, ok = np.zeros(256, np.uint8), np.zeros(256, bool)
    for k, ch in enumerate(b"ACGT"):
        code[ch], ok[ch] = k, True
    c = code[bases].reshape(-1, 4)
    p = (c[:, 0] | (c[:, 1] << 2) | (c[:, 2] << 4) | (c[:, 3] << 6)).astype(np.uint8)
    idx = np.nonzero(~ok[bases] & (bases != 0))[0]
    exc = (idx.astype(np.uint64) << np.uint64(8)) | bases[idx].astype(np.uint64)
    return p, exc, cigar.astype(np.uint16), int(not (cigar >> 16).any()) if cigar.size else 1


@pytest.mark.parametrize("n", [0, 4, 124, 128, 132, 1000, 4096 + 12, 300000, 5 * (1 << 20) + 12])
@pytest.mark.parametrize("threads", [1, 3, 8])
def test_pack_group_matches_the_documented_forms(n, threads):
    rng = np.random.default_rng(n + threads)
    bases = rng.choice(np.frombuffer(b"ACGT", np.uint8), n).astype(np.uint8)
    if n:
        k = max(1, n // 50)                       # every byte value occurs among the exceptions, 0 (padding) included
        bases[rng.integers(0, n, k)] = rng.integers(0, 256, k).astype(np.uint8)
    m = n // 16 + 3
    cigar = (rng.integers(0, 4096, m).astype(np.uint32) << 4) | rng.integers(0, 9, m).astype(np.uint32)
    if threads == 3 and n > 1000:
        cigar[m // 2] = (5000 << 4) | 2           # one op too long for 16 bits
    p, e, ne, c16, fits = _pack_group(bases, cigar, threads)
    rp, re_, rc16, rfits = _numpy_forms(bases, cigar)
    assert np.array_equal(p, rp)
    assert ne == re_.size and np.array_equal(e, re_)
    assert fits == rfits and np.array_equal(c16, rc16)


def test_pack_group_counts_when_the_exception_buffer_is_too_small_and_rejects_bad_sizes():
    bases = np.frombuffer(b"ACGN" * 1000, np.uint8).copy()
    _, e, ne, _, _ = _pack_group(bases, np.zeros(0, np.uint32), 2, cap=10)
    assert ne == 1000
    lib = capi.load()
    ne, fits = C.c_int64(0), C.c_int32(0)
    rc = lib.pv_pack_group(C.c_void_p(bases.ctypes.data), C.c_int64(6), C.c_void_p(bases.ctypes.data), None, C.c_int64(0), C.byref(ne),
                           None, C.c_int64(0), None, C.byref(fits), C.c_int32(1))
    assert rc == -1                                              # PV_EINVAL
    with pytest.raises(capi.PvError):
        capi.check(rc)


def test_group_packer_equals_the_per_array_packers_on_views():
    """pipeline._GroupPacker on region views of a synthetic batch = ReadBatch.pack_bases2 / pack_cigar16 of the same view."""
    b = synth.generate("ont_r9", 420000, 12.0, seed=5)            # 5 regions; the generator plants N bases
    packer = pipeline._GroupPacker(threads=3, slots=2)
    for j, g in enumerate([(0, 2), (2, 5), (1, 4)]):
        view, slot = packer.submit(b, g, j).result()
        assert slot == j % 2 and view.bases2 is not None and view.cigar16 is not None
        want = b.region_range_view(*g)
        want.pack_bases2(threads=2)
        want.pack_cigar16(threads=2)
        assert np.array_equal(view.bases2, want.bases2)
        assert np.array_equal(view.base_exceptions, want.base_exceptions)
        assert np.array_equal(view.cigar16, want.cigar16)
        assert np.array_equal(view.bases, want.bases) and view.n_reads == want.n_reads


def test_mix_of_packed_and_plain_groups_follows_the_measured_rates(monkeypatch):
    """_GroupPacker.choose_plain: x = (T_p - T_wp) / (T_p - T_wp + T_wl) of the groups travel plain, dealt out evenly; nothing
    before both rates are known; nothing while the two sides are nearly balanced (x < 0.2)."""
    pk = pipeline._GroupPacker(threads=2, slots=2)
    n_b, other = 240_000_000, 60_000_000
    assert not pk.choose_plain(n_b, other)                        # no measurements yet
    pk.wire_s_per_byte = 0.02e-9                                  # 50 GB/s
    t_wp, t_wl = (n_b // 4 + other) * 0.02e-9, (n_b + other) * 0.02e-9
    for t_p_ms, want in [(2.0, 0.0), (3.0, 0.0), (6.4, None), (15.0, None), (45.0, None)]:
        pk.pack_s_per_base, pk._plain_acc = t_p_ms * 1e-3 / n_b, 0.0
        t_p = t_p_ms * 1e-3
        x = 0.0 if t_p <= t_wp else (t_p - t_wp) / (t_p - t_wp + t_wl)
        picks = [pk.choose_plain(n_b, other) for _ in range(200)]
        if want == 0.0 or x < 0.2:
            assert not any(picks)
        else:
            assert abs(sum(picks) / 200.0 - x) < 0.02
            runs = "".join("1" if p else "0" for p in picks)
            assert "1111" not in runs or x > 0.75                  # spread out, not bunched
    monkeypatch.setenv("PV_PACK_MIX", "0")
    pk.pack_s_per_base = 45e-3 / n_b
    assert not any(pk.choose_plain(n_b, other) for _ in range(20))


def test_default_pack_threads_is_this_ranks_share_of_the_cores(monkeypatch):
    import os
    cores = min(os.cpu_count() or 1, len(os.sched_getaffinity(0)))
    monkeypatch.delenv("PV_PACK_THREADS", raising=False)
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "1")
    assert pipeline.default_pack_threads() == max(1, min(32, cores - 1))
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "4")
    assert pipeline.default_pack_threads() == max(1, min(32, cores // 4 - 1))
    monkeypatch.setenv("PV_PACK_THREADS", "5")
    assert pipeline.default_pack_threads() == 5


def test_pack_group_under_asan_and_ubsan(tmp_path):
    """csrc/host_pack.cpp compiled with -fsanitize=address,undefined and driven by tests/native/host_pack_fuzz.cpp: random sizes
    and thread counts, exact-size misaligned heap buffers (an overrun of the streaming stores or of the scalar head / tail is
    caught), outputs verified element by element."""
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    exe = str(tmp_path / "host_pack_fuzz")
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
                    "-I", os.path.join(root, "include"), "-I", os.path.join(root, "pepper-thesis_b200", "csrc"),
                    os.path.join(here, "native", "host_pack_fuzz.cpp"), os.path.join(root, "pepper-thesis_b200", "csrc", "host_pack.cpp"),
                    "-o", exe, "-lpthread"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "asan harness ok" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
