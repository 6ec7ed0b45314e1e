"""Region-sharded hot path: pileup summary -> candidate windows -> TransducerGRU inference, all on the GPU.

Mirrors what `pepper_variant call_variant` does between reading the BAM and writing predictions
(/root/reference/pepper_variant/modules/python/ImageGenerationUI.py:191-260 make_images +
/root/reference/pepper_variant/modules/python/models/predict_distributed_gpu.py:48-70 run_inference) without the HDF5
round trip: the int16 windows the summary kernels emit are consumed in place by the inference kernels.

Multi-GPU: regions are independent. Rank r of N takes the contiguous block
``shard_regions(n_regions, r, N)``; there is NO data-path collective, only a host-side gather of the small result
records (``gather_to_rank0``), exactly like the reference concatenates per-process outputs.
"""
from __future__ import annotations

import os

from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from . import device as dev
from .read_batch import ReadBatch


def shard_regions(n_regions: int, rank: int, world: int):
    """Contiguous block of regions for ``rank`` (block sizes differ by at most one)."""
    base, rem = divmod(n_regions, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


@dataclass
class Predictions:
    """SoA form of ``list[CandidateImagePrediction]`` (region_summary.h:114-136) for a set of regions."""
    region: np.ndarray        # int32 [K] region index (global)
    position: np.ndarray      # int64 [K]
    depth: np.ndarray         # int32 [K]
    frequency: np.ndarray     # int32 [K]
    allele: np.ndarray        # uint8 [K, 64]
    allele_len: np.ndarray    # uint8 [K]
    probs: np.ndarray         # float32 [K, 3]  prediction_type (softmax)
    genotype: np.ndarray      # uint8 [K] argmax

    def __len__(self):
        return int(self.position.shape[0])

    def alleles(self):
        return [bytes(self.allele[i, :self.allele_len[i]]) for i in range(len(self))]

    @staticmethod
    def empty():
        return Predictions(np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros(0, np.int32), np.zeros(0, np.int32),
                           np.zeros((0, 64), np.uint8), np.zeros(0, np.uint8), np.zeros((0, 3), np.float32),
                           np.zeros(0, np.uint8))

    @staticmethod
    def concat(parts: List["Predictions"]) -> "Predictions":
        parts = [p for p in parts if p is not None]
        if not parts:
            return Predictions.empty()
        return Predictions(*[np.concatenate([getattr(p, f) for p in parts]) for f in
                             ("region", "position", "depth", "frequency", "allele", "allele_len", "probs", "genotype")])

    def sorted(self) -> "Predictions":
        """(region, position) order; within a position the emission order is kept (stable)."""
        order = np.lexsort((self.position, self.region))
        return Predictions(*[getattr(self, f)[order] for f in
                             ("region", "position", "depth", "frequency", "allele", "allele_len", "probs", "genotype")])


def merge_results(parts: List[Predictions]) -> Predictions:
    """Host-side gather step: concatenate per-rank (or per-group) results and restore (region, position) order."""
    return Predictions.concat(parts).sorted()


def gather_to_rank0(local: Predictions, group=None) -> Optional[Predictions]:
    """Gather the per-rank result records on rank 0 (torch.distributed, any backend; host objects, no device collective)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local, gathered, dst=0, group=group)
    return merge_results(gathered) if rank == 0 else None


class _GroupPacker:
    """Packs the groups of a PLAIN host batch into the compact transport forms (2-bit bases + exception list, 16-bit CIGAR;
    ``pv_pack_group``, csrc/host_pack.cpp) on a worker thread, a few groups in front of their upload, into a ring of
    page-locked staging slots. A slot is reused only after the upload that read it has completed (its event)."""

    def __init__(self, threads: int, slots: int, pack_cigar: bool = True):
        from concurrent.futures import ThreadPoolExecutor
        self.threads = threads
        self.pack_cigar = pack_cigar
        # two workers take turns: while one queues its group's upload (Python, ~0.5 ms) the other one is already packing
        # (the C call is threaded itself, admits one job at a time and releases the GIL)
        self.pool = ThreadPoolExecutor(max_workers=4)       # (as many as groups in flight: a plain group must not queue behind waiting packers)
        self.slots = [dict(b2=None, c16=None, exc=None, ev=None) for _ in range(slots)]
        self.lib = None
        # the mix of packed and plain groups (choose_plain): running estimates of the two rates that decide it
        self.pack_s_per_base = None         # seconds of packing per base (this rank's threads, as contended as they are)
        self.wire_s_per_byte = None         # seconds per uploaded byte (this rank's share of the host's H2D bandwidth)
        self._uploads = []                  # (start event, end event, bytes) of uploads not yet measured
        import threading
        self._uploads_lock = threading.Lock()
        self._plain_acc = 0.0
        self.n_plain = self.n_packed = 0

    def note_upload(self, e0, e1, nbytes):
        with self._uploads_lock:
            self._uploads.append((e0, e1, nbytes))

    def _poll_uploads(self):
        with self._uploads_lock:
            pending, self._uploads = self._uploads, []
        keep = []
        for e0, e1, nbytes in pending:
            if e1.query():
                t = e0.elapsed_time(e1) * 1e-3 / max(1, nbytes)
                self.wire_s_per_byte = t if self.wire_s_per_byte is None else 0.7 * self.wire_s_per_byte + 0.3 * t
            else:
                keep.append((e0, e1, nbytes))
        with self._uploads_lock:
            self._uploads = keep + self._uploads

    def choose_plain(self, n_bases: int, other_bytes: int) -> bool:
        """Should this group travel plain? Packing costs host time (T_p), the wire carries a packed group in T_wp and a plain
        one in T_wl. With a share x of the groups plain the packing threads are busy (1 - x) T_p per group and the wire
        (1 - x) T_wp + x T_wl; both finish together at x = (T_p - T_wp) / (T_p - T_wp + T_wl) (0 when the wire is the slower
        one anyway). With 7 threads per GPU that is one group in four, with 3 threads one in two (measured on one B200, ms per
        64 Mbp: 7 threads 53 mixed / 66 all packed / 76 all plain; 3 threads 69 / 124 / 76; with ONE thread the mix loses to
        the plain upload, 93 against 76 -- the caller should not ask for inline packing then, and bench.py times both). The rates are measured
        as the run goes (packing: host clock around the call; wire: events around every upload), so the mix follows the
        cores and the H2D bandwidth this rank really gets. Groups are dealt out like Bresenham's line."""
        self._poll_uploads()
        if self.pack_s_per_base is None or self.wire_s_per_byte is None or os.environ.get("PV_PACK_MIX", "1") != "1":
            return False
        t_p = n_bases * self.pack_s_per_base
        t_wp = (n_bases // 4 + other_bytes) * self.wire_s_per_byte
        t_wl = (n_bases + other_bytes) * self.wire_s_per_byte
        x = 0.0 if t_p <= t_wp else (t_p - t_wp) / (t_p - t_wp + t_wl)
        if x < 0.2:         # nearly balanced already: a plain group then costs more (its 5.8 ms on the wire delay the packed
            return False    # groups behind it) than it saves -- measured with 15 threads: 48.8 against 45.8 ms per 64 Mbp
        self._plain_acc += x
        if self._plain_acc >= 1.0:
            self._plain_acc -= 1.0
            return True
        return False

    @staticmethod
    def _pinned(nbytes):
        # (page-locked whenever there is a device to upload to; the packing itself is host work and is unit-tested without one)
        return torch.empty(max(64, int(nbytes * 1.1) + 4096), dtype=torch.uint8, pin_memory=torch.cuda.is_available())

    def _pack(self, batch: ReadBatch, g, slot_i: int, plain: bool = False):
        import ctypes as C
        import time
        from . import capi
        if self.lib is None:
            self.lib = capi.load()
        view = batch.region_range_view(*g)
        nb, no = view.n_bases, view.n_ops
        if nb == 0 or nb % 16 or no == 0:
            return view, None
        if plain:
            self.n_plain += 1
            return view, None
        self.n_packed += 1
        t0 = time.perf_counter()
        sl = self.slots[slot_i]
        if sl["ev"] is not None:
            sl["ev"].synchronize()
            sl["ev"] = None
        if sl["b2"] is None or sl["b2"].numel() < nb // 4:
            sl["b2"] = self._pinned(nb // 4)
        if sl["c16"] is None or sl["c16"].numel() < no * 2:
            sl["c16"] = self._pinned(no * 2)
        if sl["exc"] is None:
            sl["exc"] = self._pinned(1 << 20)
        n_exc, fits = C.c_int64(0), C.c_int32(0)
        for _ in range(2):
            cap = sl["exc"].numel() // 8
            capi.check(self.lib.pv_pack_group(C.c_void_p(view.bases.ctypes.data), C.c_int64(nb), C.c_void_p(sl["b2"].data_ptr()),
                                              C.c_void_p(sl["exc"].data_ptr()), C.c_int64(cap), C.byref(n_exc),
                                              C.c_void_p(view.cigar.ctypes.data), C.c_int64(no if self.pack_cigar else 0),
                                              C.c_void_p(sl["c16"].data_ptr()), C.byref(fits), C.c_int32(self.threads)))
            if n_exc.value <= cap or n_exc.value * 8 > nb // 4:
                break
            sl["exc"] = self._pinned(n_exc.value * 8)
        if n_exc.value * 8 <= nb // 4:            # else the exception list would cost more than the packing saves: plain bases
            view.bases2 = sl["b2"].numpy()[:nb // 4]
            view.base_exceptions = sl["exc"].numpy()[:n_exc.value * 8].view(np.uint64)
        if fits.value and self.pack_cigar:
            view.cigar16 = sl["c16"].numpy()[:no * 2].view(np.uint16)
        t = (time.perf_counter() - t0) / nb
        self.pack_s_per_base = t if self.pack_s_per_base is None else 0.7 * self.pack_s_per_base + 0.3 * t
        return view, slot_i

    def submit(self, batch, g, j):
        return self.pool.submit(self._pack, batch, g, j % len(self.slots))

    def uploaded(self, slot_i, ev):
        self.slots[slot_i]["ev"] = ev


def default_pack_threads() -> int:
    """Host threads one rank may spend on packing: its share of the cores, less one for the Python thread that drives the
    GPU."""
    cores = os.cpu_count() or 1
    try:
        cores = min(cores, len(os.sched_getaffinity(0)))       # the cores this process may run on
    except (AttributeError, OSError):
        pass
    local = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")) or 1)
    return int(os.environ.get("PV_PACK_THREADS", max(1, min(32, cores // max(1, local) - 1))))


class HotPath:
    """summary + inference for batches of regions on one GPU.

    Regions are summarised in groups of ``group_regions`` (bounded scratch, H2D of the next group overlaps the
    kernels of the current one); the int16 windows of consecutive groups are accumulated in HBM and the model runs
    on full passes of ``infer_batch`` windows, so the tensor-core kernels always see full waves of tiles.
    """

    def __init__(self, model, thresholds, device="cuda", group_regions: int = 40, candidates_per_kbp: float = 8.0,
                 wrap_int8: bool = True, infer_batch: int = 32768, taper: bool = True, skip_quals: bool = True,
                 pack_inline: bool = False, pack_threads: int = 0, pack_cigar: bool = False, host_ahead: int = 0):
        self.model = model
        self.pack_inline = pack_inline      # run_host on a PLAIN batch: every group is squeezed into the 2-bit / 16-bit transport
                                            # forms by host threads while the group before it is on the wire (_GroupPacker);
                                            # pays when the host has cores to spare per GPU (the plain upload is PCIe-bound)
        self.pack_threads = pack_threads or default_pack_threads()
        self.pack_cigar = pack_cigar        # also the CIGAR words (32 -> 16 bits): pays only while the wire, not the host's
                                            # memory bandwidth, sets the pace (measured on 16 cores: 49 ms without, 54-61 with)
        self.host_ahead = host_ahead or int(os.environ.get("PV_HOST_AHEAD", "4"))   # groups uploaded in front of the kernels
        self._packer = None
        self.skip_quals = skip_quals        # run_host: a batch whose min_qual promise clears both quality thresholds is
                                            # uploaded without its quality array (no kernel would read it)
        self.taper = taper                  # run_host: shrink the last groups (upload-bound runs); False when the kernels,
                                            # not the uploads, set the pace (compact wire forms): full groups to the end
        self.thr = thresholds
        self.device = torch.device(device)
        self.group_regions = group_regions
        self.cand_per_kbp = candidates_per_kbp
        self.wrap_int8 = wrap_int8
        self.infer_batch = infer_batch
        self._ws = None
        self._ws_key = None
        self._host_counts = None
        self._acc = None
        self._wave = None
        self.copy_stream = None
        self.last_h2d_bytes = 0

    # ---- summary of one group -----------------------------------------------------------------------------------------
    # Two workspaces alternate: the kernels of group i+1 are already queued when the host waits (on an event, not on the
    # whole stream) for the candidate count of group i, so the GPU never idles on the one host round trip per group.
    def _key(self, db: dev.DeviceBatch):
        h = db.host
        cap = max(4096, int(db.total_positions / 1000.0 * self.cand_per_kbp))
        return (h.n_reads, h.n_ops, h.n_regions, db.total_positions, cap, int(db.region_len.max()) if db.region_len.size else 0)

    def _ensure_workspaces(self, db: dev.DeviceBatch):
        key = self._key(db)
        # the scratch a shape needs is not monotone in the shape (a smaller batch gets smaller tiles, hence more of them):
        # the pool is checked against the bytes this very shape asks for, not against the shape it was built for
        need = dev.SummaryWorkspace.scratch_bytes(key[0], key[1], key[2], key[3], key[4], key[5])
        if self._ws is None or key[4] > self._ws[0].capacity or need > self._ws[0].ws.numel():
            torch.cuda.current_stream(self.device).synchronize()
            grow = tuple(int(max(a, b) * 1.05) + 16 for a, b in zip(self._ws_key or key, key))
            self._ws = [dev.SummaryWorkspace(grow[0], grow[1], grow[2], grow[3], grow[4], self.device, max_region_len=grow[5],
                                             min_scratch_bytes=int(need * 1.05)) for _ in range(2)]
            self._ws_key = grow
            self._host_counts = [torch.zeros(2, dtype=torch.int64).pin_memory() for _ in range(2)]
            return True
        return False

    def _launch_summary(self, db: dev.DeviceBatch, slot: int):
        ws = self._ws[slot]
        dev.summary_regions(db, self.thr, ws)
        hc = self._host_counts[slot]
        hc[0:1].copy_(ws.count, non_blocking=True)
        hc[1:2].copy_(ws.ws[12:16].view(torch.int32), non_blocking=True)       # status word
        # (a blocking event -- the host sleeps instead of spinning while the packing threads use the cores -- was measured with
        # inline packing and lost: 50.8 against 46.8 ms per 64 Mbp, the wake-up comes late)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        return dict(db=db, slot=slot, ev=ev, ws=ws, hc=hc)      # keeps its workspace alive even if the pool is regrown

    def _collect_summary(self, h):
        """Waits for the summary of one group; returns (workspace, K). Re-runs the group with a larger candidate
        capacity when it overflowed (rare)."""
        h["ev"].synchronize()
        k, st = int(h["hc"][0]), int(h["hc"][1])
        ws = h["ws"]
        for _ in range(12):
            if st & 40:
                raise RuntimeError("internal inconsistency in the summary kernels (status %d)" % st)
            if k <= ws.capacity and not (st & 23):
                return ws, k
            if st & 16:
                # the batch travelled without its qualities (min_qual promise) and a read whose CIGAR runs over its own
                # end needed one: upload them and run the group again
                h["db"].ensure_quals()
            if k > ws.capacity or (st & 7):
                # site / event / candidate scratch all grow with the candidate capacity (summary.cu make_plan)
                self.cand_per_kbp *= 2.0 * max(1.0, k / max(1, ws.capacity))
                self._ws = None
                self._ensure_workspaces(h["db"])
                ws = self._ws[h["slot"]]
            dev.summary_regions(h["db"], self.thr, ws)
            k, st = int(ws.count.item()), ws.status()
        raise RuntimeError("summary scratch still overflows after 12 capacity doublings (status %d, %d candidates, capacity %d)"
                           % (st, k, ws.capacity))

    def summarize(self, db: dev.DeviceBatch):
        """Summary kernel chain on one group (synchronous form); returns (workspace, K)."""
        self._ensure_workspaces(db)
        return self._collect_summary(self._launch_summary(db, 0))

    # ---- accumulate windows, infer in full passes -----------------------------------------------------------------------
    class _Run:
        def __init__(self):
            self.meta = []          # per group: dict of small device tensors
            self.probs = []
            self.arg = []
            self.n_acc = 0
            self.total = 0

    def _acc_buffer(self, need, keep=0):
        if self._acc is None or self._acc.shape[0] < need:
            new = torch.empty((int(need * 1.25) + 1024, 33, 26), dtype=torch.int16, device=self.device)
            if keep:
                new[:keep].copy_(self._acc[:keep])
            self._acc = new
        return self._acc

    def _push(self, run, ws, k, region_offset):
        if k == 0:
            return
        acc = self._acc_buffer(run.n_acc + k, keep=run.n_acc)
        acc[run.n_acc:run.n_acc + k].copy_(ws.windows[:k])
        run.meta.append(dict(region=ws.region[:k] + region_offset, position=ws.position[:k].clone(),
                             depth=ws.depth[:k].clone(), frequency=ws.frequency[:k].clone(),
                             allele=ws.allele[:k].clone(), allele_len=ws.allele_len[:k].clone()))
        run.n_acc += k
        run.total += k
        self._drain(run, final=False)

    def _wave_windows(self):
        """Windows that fill whole waves of the recurrent step kernels: a launch works on (windows / 256) x 8 cluster tiles
        (4 column tiles x 2 directions per pair of 128-row tiles) with one 2-CTA cluster per pair of SMs."""
        if self._wave is None:
            import math
            clusters = max(1, torch.cuda.get_device_properties(self.device).multi_processor_count // 2)
            self._wave = 256 * clusters // math.gcd(clusters, 8)
        return self._wave

    def _drain(self, run, final, aligned=False):
        """``aligned``: infer the largest whole number of waves that has accumulated and keep the rest for the next pass
        (a pass of 19.5 k windows is 8.3 waves on 148 SMs and costs 9)."""
        b = self.infer_batch
        if aligned and not final:
            q = self._wave_windows()
            b = min(b // q, run.n_acc // q) * q
            if b <= 0:
                return
        while run.n_acc >= b or (final and run.n_acc > 0):
            m = b if run.n_acc >= b else run.n_acc
            probs, arg = self.model.infer_windows(self._acc[:m], wrap_int8=self.wrap_int8)
            run.probs.append(probs)
            run.arg.append(arg)
            rest = run.n_acc - m
            if rest:
                self._acc[:rest].copy_(self._acc[m:m + rest].clone() if rest > m else self._acc[m:m + rest])
            run.n_acc = rest

    def _finish(self, run, to_host):
        self._drain(run, final=True)
        if run.total == 0:
            return Predictions.empty() if to_host else dict(count=0)
        cat = {k: torch.cat([m[k] for m in run.meta]) for k in run.meta[0]}
        probs, arg = torch.cat(run.probs), torch.cat(run.arg)
        if not to_host:
            return dict(count=run.total, probs=probs, genotype=arg, **cat)
        return Predictions(cat["region"].cpu().numpy(), cat["position"].cpu().numpy(), cat["depth"].cpu().numpy(),
                           cat["frequency"].cpu().numpy(), cat["allele"].cpu().numpy(), cat["allele_len"].cpu().numpy(),
                           probs.cpu().numpy(), arg.cpu().numpy())

    # ---- public entry points --------------------------------------------------------------------------------------------
    def run_device(self, dbs, region_offsets=None, to_host: bool = True):
        """Groups of regions already resident in HBM (one DeviceBatch or a list of them)."""
        if isinstance(dbs, (dev.DeviceBatch, dev.DeviceReadBatch)):
            dbs = [dbs]
        if region_offsets is None:
            region_offsets, o = [], 0
            for d in dbs:
                region_offsets.append(o)
                o += d.host.n_regions
        elif isinstance(region_offsets, int):
            region_offsets = [region_offsets]
        run = HotPath._Run()
        self._acc_buffer(self.infer_batch * 2)
        pending = None
        for i, (d, off) in enumerate(zip(dbs, region_offsets)):
            self._ensure_workspaces(d)            # a regrown pool leaves the pending group's own workspace untouched
            h = self._launch_summary(d, i & 1)
            if pending is not None:
                ws, k = self._collect_summary(pending[0])
                self._push(run, ws, k, pending[1])
            pending = (h, off)
        if pending is not None:
            ws, k = self._collect_summary(pending[0])
            self._push(run, ws, k, pending[1])
        return self._finish(run, to_host)

    def run_host(self, batch: ReadBatch, region_offset: int = 0) -> Predictions:
        """Host buffers in, host results out: the H2D copy of group i+1 (side stream, truly asynchronous when the batch
        lives in pinned memory) overlaps the kernels of group i."""
        n = batch.n_regions
        if n == 0:
            return Predictions.empty()
        if self.copy_stream is None:
            self.copy_stream = torch.cuda.Stream(self.device)
        # The run ends one group's kernels (+ the inference of whatever windows are still waiting) after the last upload
        # ends, so the last groups are made small: a short first group (nothing to overlap with yet), full groups, then a
        # taper g/2, g/4, ... whose windows are inferred at once.
        g = self.group_regions
        taper = [t for t in (g // 2, g // 4, g // 8) if t >= 2] if self.taper else []
        while taper and sum(taper) > n // 2:
            taper.pop(0)
        body = n - sum(taper)
        sizes = [body % g] if body % g else []
        sizes += [g] * (body // g)
        if not self.taper and g >= 16 and n >= 3 * g:
            # kernel-bound run: nothing overlaps the first upload, and every group's upload has to hide behind the
            # kernels of the group before it (an upload takes ~0.65 of its group's kernel time): ramp up g/8, g/4, g/2
            ramp = [g // 8, g // 4, g // 2]
            rest = n - sum(ramp)
            r = rest % g
            if r >= g // 4:
                ramp.append(r)
            else:
                ramp[-1] += r
            sizes = ramp + [g] * (rest // g)
        n_body = len(sizes)
        sizes += taper
        groups, r0 = [], 0
        for sz in sizes:
            groups.append((r0, r0 + sz)); r0 += sz
        main = torch.cuda.current_stream(self.device)

        skip_q = bool(self.skip_quals and dev.quals_not_needed(batch.min_qual, self.thr))
        self.last_h2d_bytes = 0              # bytes this call copies host -> device (counted from the uploaded arrays)
        ahead = self.host_ahead
        # inline packing only of a batch that carries no transport form of its own
        packing = bool(self.pack_inline and batch.bases2 is None and batch.bases4 is None and batch.bases_patch is None
                       and batch.cigar16 is None and batch.cigar8 is None)
        run = HotPath._Run()
        self._acc_buffer(self.infer_batch * 2)
        if packing:
            return self._run_host_packing(batch, groups, n_body, region_offset, run, main, skip_q)

        def upload(j):
            with torch.cuda.stream(self.copy_stream):
                db = dev.DeviceBatch(batch.region_range_view(*groups[j]), self.device, non_blocking=True, defer_unpack=True,
                                     skip_quals=skip_q)
                self.last_h2d_bytes += db.h2d_bytes
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
            return db, ev

        # uploads are queued `ahead` groups in front of the kernels: the copy stream never waits for the host (which
        # blocks on every group's candidate count), so the copies run back to back at PCIe speed
        # building a group's upload costs the host ~0.5 ms: the first group's kernels are launched as soon as two uploads
        # are queued, the look-ahead fills up (two more per group) while the GPU already works
        queue = [upload(j) for j in range(min(2, ahead, len(groups)))]
        nxt = len(queue)
        for i, g in enumerate(groups):
            db, ev = queue.pop(0)
            main.wait_event(ev)
            db.unpack()                           # compact wire forms -> plain arrays, on the compute stream
            self._ensure_workspaces(db)
            h = self._launch_summary(db, 0)
            for _ in range(2):                    # host work while the GPU runs the group
                if nxt < len(groups) and nxt <= i + ahead:
                    queue.append(upload(nxt)); nxt += 1
            ws, k = self._collect_summary(h)
            self._push(run, ws, k, region_offset + g[0])
            # this path is bound by the uploads and the kernels have ~50 % slack: infer every full group's windows at once
            # (smaller passes, a little less efficient) instead of letting full passes pile up behind the last upload;
            # the small groups of the taper share one last pass (a pass costs >= 66 launch latencies however small)
            if i == len(groups) - 1:
                self._drain(run, final=True)
            elif i < n_body:
                self._drain(run, final=False, aligned=True)
            db.record_stream(main)
        return self._finish(run, True)

    def _run_host_packing(self, batch, groups, n_body, region_offset, run, main, skip_q):
        """run_host with inline transport packing: the packing threads set the pace (a group of 48 regions is packed in ~3 ms,
        uploaded in 2.3, summarised and inferred in ~2.5), so everything else is arranged around them. Two worker threads take
        the groups in turn: each packs its group (one packing job at a time inside pv_pack_group) and then queues the group's
        upload on the copy stream ITSELF -- the wire starts the moment the bytes exist, and the other worker is already
        packing the next group meanwhile. The calling thread only consumes finished uploads: expansion + summary kernels,
        inference in whole waves of the step kernels while full groups arrive (one host round trip per group is what the
        calling thread can afford), everything that is waiting with the last full group, so that behind the last packing job
        only the small groups of the taper remain."""
        in_flight = self.host_ahead + 2
        nslots = in_flight + 2
        if self._packer is None or len(self._packer.slots) < nslots:
            self._packer = _GroupPacker(self.pack_threads, nslots, pack_cigar=self.pack_cigar)
        packer = self._packer

        import threading
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        turn = threading.Condition()
        state = {"next": 0, "submitted": 0, "tickets": 0}

        def job(j, plain, ticket):
            torch.cuda.set_device(dev_index)              # the current device is per thread
            if plain:                                      # travels as it is: straight to the wire, past the packing jobs
                view, slot = packer._pack(batch, groups[j], j % nslots, plain=True)
            else:
                with turn:                                 # groups are packed in order, whichever worker holds them
                    turn.wait_for(lambda: state["next"] == ticket)
                try:
                    view, slot = packer._pack(batch, groups[j], j % nslots)
                finally:
                    with turn:
                        state["next"] = ticket + 1
                        turn.notify_all()
            with torch.cuda.stream(self.copy_stream):
                e0 = torch.cuda.Event(enable_timing=True)
                e0.record(self.copy_stream)
                db = dev.DeviceBatch(view, self.device, non_blocking=True, defer_unpack=True, skip_quals=skip_q)
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(self.copy_stream)
            packer.note_upload(e0, ev, db.h2d_bytes)
            if slot is not None:
                packer.uploaded(slot, ev)
            return db, ev

        futs = {}

        rb, bo, co = batch.region_read_begin, batch.read_base_off, batch.read_cigar_off

        def request(upto):
            while state["submitted"] < min(len(groups), upto):
                j = state["submitted"]
                r0, r1 = int(rb[groups[j][0]]), int(rb[groups[j][1]])
                n_b = int(bo[r1] - bo[r0]) if r1 < len(bo) else int(batch.n_bases - bo[r0]) if r0 < len(bo) else 0
                n_o = int(co[r1] - co[r0]) if r1 < len(co) else int(batch.n_ops - co[r0]) if r0 < len(co) else 0
                plain = packer.choose_plain(n_b, 4 * n_o + 100000 * (groups[j][1] - groups[j][0]))
                futs[j] = packer.pool.submit(job, j, plain, state["tickets"])
                if not plain:
                    state["tickets"] += 1
                state["submitted"] += 1
        request(in_flight)
        for i, g in enumerate(groups):
            db, ev = futs.pop(i).result()
            request(i + 1 + in_flight)
            self.last_h2d_bytes += db.h2d_bytes
            main.wait_event(ev)
            db.unpack()
            self._ensure_workspaces(db)
            ws, k = self._collect_summary(self._launch_summary(db, 0))
            self._push(run, ws, k, region_offset + g[0])
            if i == n_body - 1 or i == len(groups) - 1:
                self._drain(run, final=True)
            elif i < n_body:
                self._drain(run, final=False, aligned=True)
            db.record_stream(main)
        return self._finish(run, True)
