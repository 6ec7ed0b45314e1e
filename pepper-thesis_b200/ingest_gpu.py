"""BAM -> packed read batch ON THE DEVICE ("overlapping inflate with GPU work").

The host reads the compressed bytes a BAI query names into page-locked memory and walks the BGZF block headers
(``pv_bam_plan*`` of libpv_ingest.so); the bytes go up once, and kernels of libpepper_b200.so (csrc/ingest_gpu.cu) inflate
every BGZF block (one warp per block, csrc/inflate_warp.cuh), find the record boundaries, apply ``BAM_handler::get_reads``
(/root/reference/pepper_variant/modules/cpp/bam_handler.cpp:115-451: flag / mapq filters, the cut of every read to its
region +- 100 bases, HP tag) and write the ``PvReadBatch`` arrays in HBM -- the summary kernels start from there, no read
byte ever exists on the host in decoded form. Bit-identical to :func:`ingest.ingest_regions` (tests/test_ingest_gpu.py).
There is no CPU fallback: without a device the C-ABI returns PV_ENODEVICE.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np
import torch

from . import capi, ingest
from .device import DeviceReadBatch
from .read_batch import ARRAY_NAMES
from .summarizer import REGION_SAFE_BASES

_BLOCK_DT = np.dtype([("c_off", np.int64), ("c_len", np.int32), ("isize", np.int32), ("u_off", np.int64), ("crc", np.uint32), ("_pad", np.uint32)])


def _ptr(t):
    return C.c_void_p(t.data_ptr() if t is not None and t.numel() else None)


class DeviceIngestedReads:
    """Result of :func:`ingest_regions_gpu`: ``batch`` (a :class:`DeviceReadBatch`) + per-read ``pos_end``, ``hp_tag``,
    ``bam_flag`` device tensors; ``query_names`` are gathered and downloaded on first use."""

    def __init__(self, batch, pos_end, hp_tag, bam_flag, inflated, name_off, name_len, stats):
        self.batch, self.pos_end, self.hp_tag, self.bam_flag = batch, pos_end, hp_tag, bam_flag
        self._inflated, self._name_off, self._name_len = inflated, name_off, name_len
        self.stats = stats
        self._names = None

    @property
    def query_names(self) -> List[str]:
        if self._names is None:
            n = int(self._name_len.numel())
            if n == 0:
                self._names = []
            else:
                sizes = self._name_len.to(torch.int64) + 1
                off = torch.cumsum(sizes, 0) - sizes
                total = int(sizes.sum().item())
                out = torch.empty(total, dtype=torch.uint8, device=self._inflated.device)
                st = C.c_void_p(torch.cuda.current_stream(out.device).cuda_stream)
                capi.check(capi.load().pv_bam_gather_names(_ptr(self._inflated), _ptr(self._name_off), _ptr(self._name_len), _ptr(off), n, _ptr(out), st))
                self._names = [s.decode() for s in out.cpu().numpy().tobytes().split(b"\0")[:-1]]
        return self._names

    def release_stream(self):
        """Drop the inflated BAM bytes (kept only for the query names)."""
        self.query_names
        self._inflated = None


def _status(t, what):
    v = int(t.item())
    if v:
        raise capi.PvError(-1, "%s: device status %d (1 record chain broken, 2 capacity, 4 malformed record, 8/16 internal)" % (what, v))


class _PinnedPool:
    """Page-locked staging buffers, reused across calls (cudaHostAlloc of a GB costs more than decoding it). A buffer goes
    back with the event behind its last upload and is handed out again only after that event."""

    def __init__(self):
        import threading
        self._lock = threading.Lock()
        self._free = []

    def take(self, nbytes):
        nbytes = max(int(nbytes), 256)
        with self._lock:
            best = None
            for i, (buf, ev) in enumerate(self._free):
                if nbytes <= buf.numel() <= max(4 * nbytes, nbytes + (64 << 20)) and (best is None or buf.numel() < self._free[best][0].numel()):
                    best = i
            got = self._free.pop(best) if best is not None else None
        if got is None:
            return torch.empty(nbytes + nbytes // 8, dtype=torch.uint8, pin_memory=True)
        if got[1] is not None:
            got[1].synchronize()
        return got[0]

    def give(self, buf, event):
        with self._lock:
            self._free.append((buf, event))
            if len(self._free) > 6:                               # keep the largest few
                self._free.sort(key=lambda x: -x[0].numel())
                del self._free[6:]


_PINNED = _PinnedPool()


class _HostStage:
    """What the host prepares for one group of intervals: compressed bytes + block / segment tables + the reference fetch,
    all in page-locked memory. Built by :func:`_host_stage` (no CUDA call besides the pinned allocations), consumed by
    :func:`_device_stage`."""
    __slots__ = ("contig", "s", "e", "span_start", "span_stop", "n_spans", "tid", "comp_bytes", "comp_host", "blocks", "seg", "n_blocks",
                 "n_seg", "u_bytes", "clen", "regions", "rlen", "fetched_host", "fetched_len", "lo", "hi", "seconds", "opts")


def _host_stage(bam, fasta, contig, starts, ends, include_supplementary, min_mapq, downsample_rate, threads, safe_bases) -> _HostStage:
    import os
    import time
    ilib = ingest.load()
    t0 = time.perf_counter()
    h = _HostStage()
    h.contig = contig
    h.s = np.ascontiguousarray(starts, np.int64)
    h.e = np.ascontiguousarray(ends, np.int64)
    assert h.s.shape == h.e.shape and h.s.ndim == 1
    h.n_spans = n_spans = int(h.s.shape[0])
    h.span_start = np.maximum(0, h.s - safe_bases)
    h.span_stop = h.e + safe_bases
    if n_spans > 1 and (np.any(np.diff(h.span_start) < 0) or np.any(np.diff(h.span_stop) < 0)):
        raise ValueError("ingest_regions_gpu: intervals must ascend")
    if threads <= 0:
        threads = min(16, os.cpu_count() or 1)
    h.opts = (int(bool(include_supplementary)), int(min_mapq), float(downsample_rate))
    # ---- reference + region fields: ONE fetch of the covering span by FASTA_handler's C side (cut per region on the device),
    # on a thread of its own beside the BAM reads (both release the GIL)
    h.clen = int(fasta.get_chromosome_sequence_length(contig)) if n_spans else 0
    if n_spans and h.clen < 0:
        raise RuntimeError("CHROMOSOME NAME NOT PRESENT IN REFERENCE FASTA FILE: %s" % contig)
    h.regions = {"region_ref_start": h.span_start.astype(np.int64), "region_ref_end": h.span_stop.astype(np.int64),
                 "region_cand_start": h.s.copy(), "region_cand_end": h.e.copy()}
    h.rlen = (h.span_stop + 1 - h.span_start).astype(np.int64) if n_spans else np.zeros(0, np.int64)     # region_end + 1 exclusive (:214-216)
    h.regions["region_ref_len"] = h.rlen
    h.regions["region_ref_off"] = (np.cumsum(h.rlen) - h.rlen).astype(np.int64)
    h.lo = h.hi = 0
    h.fetched_host, h.fetched_len = None, 0
    fetch_thread, fetch_err = None, []
    if n_spans:
        import threading
        h.lo, h.hi = int(h.span_start.min()), int(h.span_stop.max()) + 1
        h.fetched_host = _PINNED.take(h.hi - h.lo)

        def fetch():
            try:
                got_len = C.c_int64(0)
                ingest._check(ilib.pv_fasta_fetch(fasta._h, contig.encode(), h.lo, h.hi, C.c_void_p(h.fetched_host.data_ptr()), C.byref(got_len)))
                h.fetched_len = int(got_len.value)
            except Exception as ex:                               # re-raised on the caller's thread
                fetch_err.append(ex)
        fetch_thread = threading.Thread(target=fetch)
        fetch_thread.start()
    # ---- which bytes, which blocks, which chain entry points
    plan = C.c_void_p()
    try:
        ingest._check(ilib.pv_bam_plan(bam._h, contig.encode(), int(h.span_start.min()) if n_spans else 0, int(h.span_stop.max()) if n_spans else 0, C.byref(plan)))
        h.tid = int(ilib.pv_bam_plan_tid(plan))
        if n_spans and h.tid < 0:
            raise RuntimeError("contig %s is not in the BAM header" % contig)
        h.comp_bytes = int(ilib.pv_bam_plan_comp_bytes(plan))
        h.comp_host = _PINNED.take(h.comp_bytes)
        ingest._check(ilib.pv_bam_plan_load(plan, C.c_void_p(h.comp_host.data_ptr()), int(threads)))
        h.n_blocks, h.n_seg = int(ilib.pv_bam_plan_n_blocks(plan)), int(ilib.pv_bam_plan_n_segments(plan))
        h.u_bytes = int(ilib.pv_bam_plan_inflated_bytes(plan))
        h.blocks = np.zeros(max(h.n_blocks, 1), _BLOCK_DT)
        h.seg = np.zeros((2, max(h.n_seg, 1)), np.int64)
        ingest._check(ilib.pv_bam_plan_tables(plan, h.blocks.ctypes.data, h.seg[0].ctypes.data, h.seg[1].ctypes.data))
    finally:
        if plan:
            ilib.pv_bam_plan_free(plan)
        if fetch_thread is not None:
            fetch_thread.join()
    if fetch_err:
        raise fetch_err[0]
    h.seconds = time.perf_counter() - t0
    return h


_COPY_STREAMS = {}
_PIECE_STREAMS = {}


def _copy_stream(dev):
    key = (dev.type, torch.cuda.current_device())
    copy = _COPY_STREAMS.get(key)
    if copy is None:
        copy = _COPY_STREAMS[key] = torch.cuda.Stream(dev)
    return copy


def _start_upload(h: _HostStage, dev):
    """Queues the whole compressed buffer of a host stage on the copy stream NOW (the streamed generator calls this for group
    i+1 before it queues the kernels of group i, so the transfer runs under them); :func:`_device_stage` then waits for the
    event and inflates with one launch."""
    if dev.index is not None:
        torch.cuda.set_device(dev)
    copy, main = _copy_stream(dev), torch.cuda.current_stream(dev)
    comp_n = (h.comp_bytes + 255) & ~255
    comp = torch.empty(comp_n, dtype=torch.uint8, device=dev)
    copy.wait_stream(main)
    n = min(comp_n, h.comp_host.numel())
    with torch.cuda.stream(copy):
        comp[:n].copy_(h.comp_host[:n], non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(copy)
    _PINNED.give(h.comp_host, ev)
    h.comp_host = None
    return comp, ev


def _device_stage(h: _HostStage, dev, verify_crc=True, uploaded=None) -> DeviceIngestedReads:
    import os
    import time
    lib = capi.load()
    if dev.index is not None:
        torch.cuda.set_device(dev)
    t_plan = time.perf_counter()
    n_spans, n_blocks, n_seg, u_bytes, comp_bytes, tid = h.n_spans, h.n_blocks, h.n_seg, h.u_bytes, h.comp_bytes, h.tid
    include_supplementary, min_mapq, downsample_rate = h.opts
    span_start, span_stop, s, e = h.span_start, h.span_stop, h.s, h.e
    main = torch.cuda.current_stream(dev)
    st = C.c_void_p(main.cuda_stream)
    trace = {} if os.environ.get("PV_INGEST_TRACE") else None

    def mark(name, _last=[t_plan]):
        if trace is not None:
            torch.cuda.synchronize(dev)
            now = time.perf_counter()
            trace[name] = round((now - _last[0]) * 1e3, 3)
            _last[0] = now

    # ---- upload in pieces on a copy stream; the blocks of a piece inflate as soon as its bytes have landed
    copy = _copy_stream(dev)
    key_dev = (dev.type, torch.cuda.current_device())
    blocks_d = torch.from_numpy(h.blocks.view(np.uint8).reshape(-1)).to(dev, non_blocking=True)
    seg_d = torch.from_numpy(h.seg).to(dev, non_blocking=True)
    spans_d = torch.from_numpy(np.stack([span_start, span_stop]) if n_spans else np.zeros((2, 1), np.int64)).to(dev, non_blocking=True)
    U = torch.empty(u_bytes + 64, dtype=torch.uint8, device=dev)
    flags = torch.zeros(4, dtype=torch.int32, device=dev)       # [1] chain status, [2] clip status, [3] min_qual
    bs = _BLOCK_DT.itemsize
    if uploaded is not None:
        # the bytes were queued on the copy stream earlier (under the previous group's kernels): one launch for all blocks
        comp, ev_up = uploaded
        bad_d = torch.zeros((1, 2), dtype=torch.int32, device=dev)
        main.wait_event(ev_up)
        capi.check(lib.pv_bam_inflate_blocks(_ptr(comp), comp_bytes, _ptr(blocks_d), n_blocks, _ptr(U), u_bytes, int(bool(verify_crc)),
                                             _ptr(bad_d), st))
    else:
        comp_n = (comp_bytes + 255) & ~255
        comp = torch.empty(comp_n, dtype=torch.uint8, device=dev)
        # pieces of whole waves of the inflate kernel (32 warps = blocks per SM at a time): one wave, two waves, the rest --
        # only the first piece's upload is exposed, the others land while the piece before them inflates
        wave = 32 * torch.cuda.get_device_properties(dev).multi_processor_count
        first = [0] + [b for b in (wave, 3 * wave) if b < n_blocks - wave // 2] + [n_blocks]
        pieces = len(first) - 1
        bad_d = torch.zeros((pieces, 2), dtype=torch.int32, device=dev)   # per piece: bad blocks, the kernel's ticket
        cut = [0] + [int(h.blocks["c_off"][first[k]]) & ~255 for k in range(1, pieces)] + [min(comp_n, h.comp_host.numel())]
        copy.wait_stream(main)
        # every piece inflates on a stream of its own: the next piece's warps move into the SMs that the tail of the piece
        # before leaves idle (a BGZF block takes ~10 ms, so a launch ends in a tail of about that length)
        side = _PIECE_STREAMS.get(key_dev)
        if side is None:
            side = _PIECE_STREAMS[key_dev] = [torch.cuda.Stream(dev) for _ in range(3)]
        for k in range(pieces):
            with torch.cuda.stream(copy):
                comp[cut[k]:cut[k + 1]].copy_(h.comp_host[cut[k]:cut[k + 1]], non_blocking=True)
            sk = side[k % len(side)]
            sk.wait_stream(main)                                 # the block table, U
            sk.wait_stream(copy)                                 # this piece's bytes
            if first[k + 1] > first[k]:
                capi.check(lib.pv_bam_inflate_blocks(_ptr(comp), comp_bytes, C.c_void_p(blocks_d.data_ptr() + first[k] * bs), first[k + 1] - first[k],
                                                     _ptr(U), u_bytes, int(bool(verify_crc)), C.c_void_p(bad_d.data_ptr() + 8 * k),
                                                     C.c_void_p(sk.cuda_stream)))
        for k in range(min(pieces, len(side))):
            main.wait_stream(side[k])
        ev_copy = torch.cuda.Event()
        ev_copy.record(copy)
        _PINNED.give(h.comp_host, ev_copy)
        h.comp_host = None
    # ---- the regions' reference bytes
    ref_bytes = int(h.rlen.sum())
    ref_t = {}
    for name in ("region_ref_start", "region_ref_end", "region_cand_start", "region_cand_end", "region_ref_off", "region_ref_len"):
        ref_t[name] = torch.from_numpy(np.ascontiguousarray(h.regions[name])).to(dev, non_blocking=True) if n_spans else torch.zeros(1, dtype=torch.int64, device=dev)[:0]
    ref_t["ref"] = torch.empty(ref_bytes, dtype=torch.uint8, device=dev)
    if n_spans:
        fetched = h.fetched_host[:max(h.fetched_len, 1)].to(dev, non_blocking=True)
        capi.check(lib.pv_bam_gather_reference(_ptr(fetched), h.fetched_len, h.lo, _ptr(spans_d[0]), _ptr(ref_t["region_ref_off"]),
                                               _ptr(ref_t["region_ref_len"]), n_spans, int(h.rlen.max()), _ptr(ref_t["ref"]), st))
    if h.fetched_host is not None:
        ev_main = torch.cuda.Event()
        ev_main.record(main)
        _PINNED.give(h.fetched_host, ev_main)
        h.fetched_host = None
    mark("upload+inflate+reference")
    seg_first = torch.zeros(n_seg + 1, dtype=torch.int64, device=dev)
    capi.check(lib.pv_bam_index_records(_ptr(U), u_bytes, _ptr(seg_d[0]), _ptr(seg_d[1]), n_seg, _ptr(seg_first), None, 0, _ptr(flags[1:2]), st))
    n_rec = int(seg_first[n_seg].item())
    bad = int(bad_d[:, 0].sum().item())
    if bad:
        raise capi.PvError(-1, "BGZF: %d block(s) failed to inflate or their CRC-32 does not match" % bad)
    _status(flags[1], "record index (count)")
    rec_off = torch.empty(max(n_rec, 1), dtype=torch.int64, device=dev)
    capi.check(lib.pv_bam_index_records(_ptr(U), u_bytes, _ptr(seg_d[0]), _ptr(seg_d[1]), n_seg, _ptr(seg_first), _ptr(rec_off), n_rec, _ptr(flags[1:2]), st))
    mark("record_index")
    pair_first = torch.zeros(n_rec + 1, dtype=torch.int64, device=dev)
    clip_args = (_ptr(U), u_bytes, _ptr(rec_off), n_rec, tid, _ptr(spans_d[0]), _ptr(spans_d[1]), n_spans, int(bool(include_supplementary)), int(min_mapq))
    capi.check(lib.pv_bam_clip_count(*clip_args, _ptr(pair_first), _ptr(flags[2:3]), st))
    n_pairs = int(pair_first[n_rec].item())
    _status(flags[1], "record index (fill)")
    mark("clip_count")
    ws = torch.empty(int(lib.pv_bam_clip_workspace_bytes(n_pairs)), dtype=torch.uint8, device=dev)
    pairs = torch.empty((max(n_pairs, 1), 8), dtype=torch.int64, device=dev)     # PvBamPair: 64 bytes
    base_off = torch.empty(max(n_pairs, 1), dtype=torch.int64, device=dev)
    cigar_off = torch.empty(max(n_pairs, 1), dtype=torch.int64, device=dev)
    read_begin = torch.zeros(n_spans + 1, dtype=torch.int64, device=dev)
    totals = torch.zeros(2, dtype=torch.int64, device=dev)
    capi.check(lib.pv_bam_clip_layout(*clip_args, _ptr(pair_first), n_pairs, _ptr(ws), ws.numel(), _ptr(pairs), _ptr(base_off), _ptr(cigar_off),
                                      _ptr(read_begin), _ptr(totals), _ptr(flags[2:3]), st))
    rb_host = read_begin.cpu().numpy()
    tot = totals.cpu().numpy()
    mark("clip_layout")
    _status(flags[2], "clip")
    n_reads, n_bases, n_ops = n_pairs, int(tot[0]), int(tot[1])
    # reservoir down-sampling (AlignmentSummarizer.py:191-208) on read indices, before any read is materialised
    keep, changed = [], False
    for r in range(n_spans):
        idx = ingest.reservoir_indices(int(rb_host[r + 1] - rb_host[r]), downsample_rate)
        if idx is None:
            keep.append(np.arange(rb_host[r], rb_host[r + 1], dtype=np.int64))
        else:
            keep.append(idx + rb_host[r]); changed = True
    if changed:
        k = torch.from_numpy(np.concatenate(keep)).to(dev)
        pairs = pairs[:n_pairs].index_select(0, k).contiguous()
        n_reads = int(k.numel())
        padded = (pairs[:, 2] + 15) & ~15
        nops = pairs[:, 1] >> 32                                  # PvBamPair: span (low 32 bits) | n_ops (high 32 bits)
        base_off = (torch.cumsum(padded, 0) - padded).contiguous()
        cigar_off = (torch.cumsum(nops, 0) - nops).contiguous()
        n_bases, n_ops = int(padded.sum().item()), int(nops.sum().item())
        rb_host = np.concatenate([[0], np.cumsum([len(x) for x in keep])]).astype(np.int64)
        read_begin = torch.from_numpy(rb_host).to(dev)
        if n_reads == 0:
            pairs = torch.empty((1, 8), dtype=torch.int64, device=dev)

    def new(n, dt):
        return torch.empty(max(int(n), 1), dtype=dt, device=dev)
    t = {"read_pos": new(n_reads, torch.int64), "read_base_off": base_off, "read_len": new(n_reads, torch.int32),
         "read_cigar_off": cigar_off, "read_n_ops": new(n_reads, torch.int32), "read_flags": new(n_reads, torch.uint8),
         "read_mapq": new(n_reads, torch.uint8), "bases": torch.empty(n_bases, dtype=torch.uint8, device=dev),
         "quals": torch.empty(n_bases, dtype=torch.uint8, device=dev), "cigar": torch.empty(n_ops, dtype=torch.int32, device=dev),
         "region_read_begin": read_begin}
    pos_end, hp = new(n_reads, torch.int64), new(n_reads, torch.int32)
    bam_flag = new(n_reads, torch.int16)
    name_off, name_len = new(n_reads, torch.int64), new(n_reads, torch.int32)
    capi.check(lib.pv_bam_clip_write(_ptr(U), u_bytes, _ptr(pairs), n_reads, _ptr(spans_d[0]), _ptr(spans_d[1]), _ptr(base_off), _ptr(cigar_off),
                                     _ptr(t["read_pos"]), _ptr(pos_end), _ptr(t["read_len"]), _ptr(t["read_n_ops"]), _ptr(t["read_flags"]),
                                     _ptr(t["read_mapq"]), _ptr(hp), _ptr(bam_flag), _ptr(name_off), _ptr(name_len), _ptr(t["bases"]),
                                     _ptr(t["quals"]), _ptr(t["cigar"]), _ptr(flags[3:4]), _ptr(flags[2:3]), st))
    mark("clip_write")
    for name in ("region_ref_start", "region_ref_end", "region_cand_start", "region_cand_end", "region_ref_off", "region_ref_len"):
        t[name] = ref_t[name]
    t["ref"] = ref_t["ref"]
    mq = int(flags[3].item())
    _status(flags[2], "clip (write)")
    for name in ("read_pos", "read_base_off", "read_len", "read_cigar_off", "read_n_ops", "read_flags", "read_mapq"):
        t[name] = t[name][:n_reads]
    batch = DeviceReadBatch(t, h.regions, mq if (n_reads and 0 < mq <= 255) else 0, dev, contigs=[h.contig] * n_spans,
                            region_contig_len=np.full(n_spans, h.clen, np.int64), h2d_bytes=comp_bytes + (h.hi - h.lo))
    assert set(ARRAY_NAMES) <= set(t)
    t_done = time.perf_counter()
    stats = {"compressed_bytes": comp_bytes, "inflated_bytes": u_bytes, "bgzf_blocks": n_blocks, "chain_segments": n_seg,
             "records": n_rec, "reads": n_reads, "host_plan_s": h.seconds, "device_s": t_done - t_plan}
    if trace is not None:
        stats["trace_ms"] = trace
    return DeviceIngestedReads(batch, pos_end[:n_reads], hp[:n_reads], bam_flag[:n_reads], U, name_off[:n_reads], name_len[:n_reads], stats)


def ingest_regions_gpu(bam: ingest.BAMHandler, fasta: ingest.FASTAHandler, contig: str, starts: Sequence[int], ends: Sequence[int],
                       include_supplementary=False, min_mapq=0, min_baseq=0, downsample_rate=1.0, threads=0,
                       safe_bases=REGION_SAFE_BASES, device="cuda", verify_crc=True) -> DeviceIngestedReads:
    """Device twin of :func:`ingest.ingest_regions`: packed batch for intervals ``[starts[i], ends[i]]`` (ascending) of one
    contig, decoded and cut on the GPU. ``min_baseq`` is accepted for signature parity (it only feeds ``bad_indicies`` in
    the reference, which the packed batch does not carry)."""
    h = _host_stage(bam, fasta, contig, starts, ends, include_supplementary, min_mapq, downsample_rate, threads, safe_bases)
    return _device_stage(h, torch.device(device), verify_crc)


def stream_regions_gpu(bam: ingest.BAMHandler, fasta: ingest.FASTAHandler, contig: str, groups, include_supplementary=False, min_mapq=0,
                       downsample_rate=1.0, threads=0, safe_bases=REGION_SAFE_BASES, device="cuda", verify_crc=True):
    """Generator over ``groups`` = [(starts, ends), ...] of one contig: yields one :class:`DeviceIngestedReads` per group, with
    the HOST share of group i+1 (BAI query, threaded read of the compressed bytes into page-locked memory, BGZF header walk,
    FASTA fetch) running on a worker thread while the caller's kernels -- the decode of group i and whatever consumes its
    batch -- occupy the device ("overlapping inflate with GPU work"). The C side releases the GIL."""
    from concurrent.futures import ThreadPoolExecutor
    dev = torch.device(device)
    groups = list(groups)
    if not groups:
        return
    if threads <= 0:
        import os
        threads = max(2, min(8, (os.cpu_count() or 2) // 2))      # the caller's thread needs a core to keep the device fed
    args = (include_supplementary, min_mapq, downsample_rate, threads, safe_bases)
    with ThreadPoolExecutor(1) as ex:
        # three stages in flight: host stage of group i+2 (worker thread) | upload of group i+1 (copy stream) | kernels of group i
        futs = [ex.submit(_host_stage, bam, fasta, contig, g[0], g[1], *args) for g in groups[:2]]
        h = futs.pop(0).result()
        up = _start_upload(h, dev)
        for i in range(len(groups)):
            nxt = None
            if futs:
                h_next = futs.pop(0).result()
                if i + 2 < len(groups):
                    futs.append(ex.submit(_host_stage, bam, fasta, contig, groups[i + 2][0], groups[i + 2][1], *args))
                nxt = (h_next, _start_upload(h_next, dev))         # queued before this group's kernels: runs under them
            yield _device_stage(h, dev, verify_crc, uploaded=up)
            if nxt is not None:
                h, up = nxt
