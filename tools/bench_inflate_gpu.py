"""Device DEFLATE decoder (pv_bam_inflate_blocks) alone: N synthetic BGZF-sized blocks per content kind, timed with CUDA events.
One JSON line per kind. Usage: python tools/bench_inflate_gpu.py [blocks] [kinds,comma]"""
import ctypes as C, json, os, sys, zlib
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pepper_thesis_b200 import capi

BLOCK_DT = np.dtype([("c_off", np.int64), ("c_len", np.int32), ("isize", np.int32), ("u_off", np.int64), ("crc", np.uint32), ("_pad", np.uint32)])
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
kinds = sys.argv[2].split(",") if len(sys.argv) > 2 else ["qual", "bases", "bam", "stored", "zeros", "text"]
rng = np.random.default_rng(1)
NT = np.array([1, 2, 4, 8], np.uint8)


def content(kind, n=0xff00):
    if kind == "qual":
        return rng.integers(33, 75, n, dtype=np.uint8).tobytes(), 1
    if kind == "bases":
        c = NT[rng.integers(0, 4, 2 * n)]
        return ((c[0::2] << 4) | c[1::2]).tobytes(), 1
    if kind == "bam":                                           # a 12 kbp read record: 1/3 packed bases, 2/3 qualities, a little CIGAR
        out = bytearray()
        while len(out) < n:
            c = NT[rng.integers(0, 4, 12000)]
            out += rng.integers(0, 256, 36, dtype=np.uint8).tobytes() + (rng.integers(1, 40, 700, dtype=np.uint32) << 4).tobytes()
            out += ((c[0::2] << 4) | c[1::2]).tobytes() + rng.integers(33, 75, 12000, dtype=np.uint8).tobytes()
        return bytes(out[:n]), 1
    if kind == "stored":
        return rng.integers(0, 256, n, dtype=np.uint8).tobytes(), 0
    if kind == "zeros":
        return bytes(n), 6
    if kind == "text":
        words = [bytes(rng.integers(97, 123, int(rng.integers(2, 12)), dtype=np.uint8)) for _ in range(300)]
        return b" ".join(words[int(i)] for i in rng.integers(0, 300, 12000))[:n], 6
    raise ValueError(kind)


def comp_one(args):
    data, level = args
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    return co.compress(data) + co.flush(), zlib.crc32(data) & 0xffffffff, len(data)


lib = capi.load()
for kind in kinds:
    distinct = [content(kind) for _ in range(64)]
    with ThreadPoolExecutor(8) as ex:
        cs = list(ex.map(comp_one, distinct))
    comp, table, u = bytearray(), np.zeros(N, BLOCK_DT), 0
    offs = []
    for c, crc, n in cs:
        offs.append(len(comp)); comp += c
    # every block reads its own copy of the payload (distinct addresses, as in a file)
    one = bytes(comp)
    reps = (N + 63) // 64
    big = one * reps
    for i in range(N):
        c, crc, n = cs[i % 64]
        table[i] = ((i // 64) * len(one) + offs[i % 64], len(c), n, u, crc, 0)
        u += n
    cd = torch.from_numpy(np.frombuffer(big, np.uint8).copy()).cuda()
    td = torch.from_numpy(table.view(np.uint8).reshape(-1).copy()).cuda()
    U = torch.empty(u + 64, dtype=torch.uint8, device="cuda")
    bad = torch.zeros(2, dtype=torch.int32, device="cuda")        # bad blocks, the kernel's ticket
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    res = {}
    for crc in (1, 0):
        ts = []
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            capi.check(lib.pv_bam_inflate_blocks(C.c_void_p(cd.data_ptr()), len(big), C.c_void_p(td.data_ptr()), N, C.c_void_p(U.data_ptr()), u, crc, C.c_void_p(bad.data_ptr()), st))
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        assert int(bad[0].item()) == 0, (kind, int(bad[0].item()))
        res["ms_crc%d" % crc] = round(min(ts[1:]), 3)
    ms = res["ms_crc1"]
    print(json.dumps(dict(kind=kind, blocks=N, comp_MB=round(len(big) / 1e6, 1), out_MB=round(u / 1e6, 1), ratio=round(u / len(big), 2), **res,
                          out_GBps=round(u / ms / 1e6, 2), comp_GBps=round(len(big) / ms / 1e6, 2))), flush=True)
