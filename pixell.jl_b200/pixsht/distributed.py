"""m-sharded multi-GPU pipeline: one process per GPU, torch.distributed for the plumbing.

  alm2map:  Legendre stage on this rank's m values (all rings) into the rank's own phase buffer  ->  stream-ordered
            barrier  ->  ring FFTs on this rank's contiguous slab of rings, whose row loads fetch every m from the phase
            buffer of the GPU that owns it (peer memory over NVLink)  ->  the rank's rows of the map.
  map2alm:  ring FFTs of the rank's rows, whose row stores put every m into its owner's phase buffer  ->  barrier  ->
            Legendre analysis on the rank's m values from its own buffer.

So the phase transpose between the m-sharded and the ring-sharded stage is fused into the FFT kernels' own row loads and
stores; there is no pack / all-to-all / unpack pass and no second copy of the phase array.  m values are dealt to ranks
in runs of 16 consecutive m, so the remote accesses are 256-byte runs (16-byte remote stores from the Legendre kernels --
the first design -- hit an NVLink transaction-rate ceiling at 8 GPUs: 106 ms instead of 74 ms of Legendre time at C4).
The only collective left is the barrier that orders the two stages (a 1-element all-reduce on the compute stream with
NCCL; dist.barrier() with gloo).

Partition: runs of 16 m are assigned by longest-processing-time on the measured work per m (executed Legendre steps from
the plan's activation tables); rings are split into contiguous slabs, which are contiguous row ranges of the caller's
column-major map.  Nothing else is partitioned; tables are replicated.

The compute stages are the pixsht_stage_* entry points of the C ABI; the peer-visible buffers are pixsht_shared_alloc /
pixsht_shared_open (CUDA IPC; POSIX shared memory in the host-emulation build that the gloo CPU tests use).
"""
import ctypes
import os

import numpy as np
import torch
import torch.distributed as dist

from ._lib import get_lib, Geom, F64, F32


def partition_m(mmax, world):
    """Deal the pairs (m, mmax-m), m = 0..ceil(mmax/2), round-robin over ranks.  Returns a list of int32 arrays."""
    lists = [[] for _ in range(world)]
    lo, hi, k = 0, mmax, 0
    while lo <= hi:
        r = k % world if (k // world) % 2 == 0 else world - 1 - (k % world)   # boustrophedon keeps the counts even
        lists[r].append(lo)
        if hi != lo:
            lists[r].append(hi)
        lo, hi, k = lo + 1, hi - 1, k + 1
    return [np.array(sorted(x), dtype=np.int32) for x in lists]


M_RUN = 16   # consecutive m per assignment unit: 256-byte runs in the phase rows


def partition_m_weighted(weights, world, run=M_RUN):
    """Longest-processing-time assignment of runs of `run` consecutive m values to ranks by work (weights[m]; measured:
    executed Legendre steps per m from the plan's activation tables), equal m counts not required.  Deterministic: every
    rank computes the same lists."""
    weights = np.asarray(weights, dtype=np.float64)
    n = len(weights)
    starts = np.arange(0, n, run)
    wrun = np.add.reduceat(weights, starts)
    order = np.argsort(-wrun, kind="stable")
    load = np.zeros(world)
    lists = [[] for _ in range(world)]
    for b in order:
        r = int(np.argmin(load))
        lists[r].extend(range(int(starts[b]), min(n, int(starts[b]) + run)))
        load[r] += wrun[b]
    return [np.array(sorted(x), dtype=np.int32) for x in lists]


def partition_rings(nrings, world):
    """Contiguous slabs [begin, end) of band rings per rank."""
    edges = [(nrings * r) // world for r in range(world + 1)]
    return [(edges[r], edges[r + 1]) for r in range(world)]


class ShardedSHT:
    """Distributed map2alm / alm2map on one 8-GPU box.  All tensors live on `device` ('cuda:i', or 'cpu' with the
    emulation library in the gloo tests).  dtype = element type of the map slabs (float64 | float32); alm tensors are
    complex128 or complex64 accordingly (Float32 data is widened on the device: the Legendre stage computes in FP64)."""

    MAX_NCOMP = 3

    def __init__(self, band, lmax, mmax=None, group=None, device=None, lib=None, balance="work", dtype=torch.float64):
        self.lib = get_lib() if lib is None else lib
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.band, self.lmax = band, int(lmax)
        self.mmax = int(lmax if mmax is None else mmax)
        self.device = torch.device(device if device is not None else "cuda")
        self.dev_index = self.device.index if self.device.type == "cuda" and self.device.index is not None else 0
        g = Geom(band.nphi, band.nrings_total, band.ring_first, band.nrings, band.nx, int(band.flipx), int(band.flipy), int(getattr(band, "ring_scheme", 0)), band.phi0)
        h = ctypes.c_void_p()
        self.dtype = dtype
        if dtype not in (torch.float64, torch.float32):
            raise TypeError("maps must be float64 or float32")
        self.cdtype = torch.complex128 if dtype == torch.float64 else torch.complex64
        self.lib.check(self.lib.lib.pixsht_plan_create(ctypes.byref(h), ctypes.byref(g), self.lmax, self.mmax,
                                                       F64 if dtype == torch.float64 else F32, self.dev_index))
        self.handle = h
        self.nalm = int(self.lib.lib.pixsht_nalm(self.lmax, self.mmax))
        self.MP = int(self.lib.lib.pixsht_phase_row_len(self.handle))
        self.nrings = band.nrings
        # m partition: by measured work per m when the library can report it (activation tables of both spin families,
        # spin 2 weighted 3x: 12 vs 4 FP64 ops per step), else the closed-form (m, mmax-m) pairing
        if self.world > 1 and balance == "work":
            w0 = (ctypes.c_double * (self.mmax + 1))()
            w2 = (ctypes.c_double * (self.mmax + 1))()
            self.lib.check(self.lib.lib.pixsht_plan_work_per_m(self.handle, 0, w0))
            self.lib.check(self.lib.lib.pixsht_plan_work_per_m(self.handle, 2, w2))
            self.m_lists = partition_m_weighted(np.array(w0[:]) + 3.0 * np.array(w2[:]), self.world)
        else:
            # closed form: cost of an m ~ lmax - m + 1 (no pruning)
            self.m_lists = partition_m_weighted(self.lmax + 1.0 - np.arange(self.mmax + 1), self.world)
        self.ring_ranges = partition_rings(self.nrings, self.world)
        self.my_m = self.m_lists[self.rank]
        self.nm = len(self.my_m)
        self.r0, self.r1 = self.ring_ranges[self.rank]
        self.nloc = self.r1 - self.r0
        self.d_m_list = torch.from_numpy(self.my_m.copy()).to(self.device)
        self.last_ms = {}
        # ---- peer-visible phase buffer: all rings x MAX_NCOMP components x (this rank's m values, padded) complex doubles ----
        L = self.lib.lib
        self.row_lens = [(len(ml) + 7) // 8 * 8 for ml in self.m_lists]
        self.row_len = self.row_lens[self.rank]
        nbytes = self.nrings * self.MAX_NCOMP * max(8, self.row_len) * 16
        own = ctypes.c_void_p()
        hbuf = ctypes.create_string_buffer(64)
        self.lib.check(L.pixsht_shared_alloc(self.dev_index, nbytes, ctypes.byref(own), hbuf))
        self.own_ptr = own.value
        self.peer_ptrs = [None] * self.world
        self.peer_ptrs[self.rank] = self.own_ptr
        self._opened = []
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(hbuf.raw), group=self.group)
            for r in range(self.world):
                if r == self.rank:
                    continue
                p = ctypes.c_void_p()
                self.lib.check(L.pixsht_shared_open(self.dev_index, handles[r], ctypes.byref(p)))
                self.peer_ptrs[r] = p.value
                self._opened.append(p.value)
        self._mtab = None
        self._flag = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.host_pieces = int(os.environ.get("PIXSHT_HOST_PIECES", "3"))   # pieces of the first input / last output of the host-resident pipelines (2 GPUs, C4: 3 -> 353 ms, 8 -> 359 ms)
        # the copy streams also run small pack / unpack kernels: high priority, so that they are not queued behind a Legendre grid
        self._copy_prio = int(os.environ.get("PIXSHT_COPY_PRIO", "-1"))

    # ---- the caller's view of the data ---------------------------------------------------------------------------
    def map_rows(self):
        """Rows [a, b) of the caller's (nx, ny) map that this rank reads / writes."""
        if self.band.flipy:
            return self.nrings - self.r1, self.nrings - self.r0
        return self.r0, self.r1

    def alm_columns(self):
        """(start, stop) index ranges in the triangular alm vector of this rank's m values."""
        L = self.lmax
        return [(int(m) * (2 * L + 1 - int(m)) // 2 + int(m), int(m) * (2 * L + 1 - int(m)) // 2 + L + 1) for m in self.my_m]

    def close(self):
        if getattr(self, "handle", None):
            L = self.lib.lib
            if self.device.type == "cuda":
                torch.cuda.synchronize(self.device)
            if self.world > 1 and dist.is_initialized():
                dist.barrier(group=self.group)      # nobody may still be reading or writing a buffer about to be unmapped
            for p in self._opened:
                L.pixsht_shared_close(ctypes.c_void_p(p))
            self._opened = []
            if self.world > 1 and dist.is_initialized():
                dist.barrier(group=self.group)
            L.pixsht_shared_free(ctypes.c_void_p(self.own_ptr))
            L.pixsht_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.world == 1:
                self.close()
        except Exception:
            pass

    # ---- helpers -----------------------------------------------------------------------------------------------
    def _m_table(self):
        """Device table of 2*(mmax+1) int64 for the FFT stages: per m the address of element (ring 0, component 0, m) in the
        buffer of the rank that owns m, and that buffer's row length."""
        if self._mtab is None:
            tab = np.empty((self.mmax + 1, 2), dtype=np.int64)
            for r, ml in enumerate(self.m_lists):
                tab[ml, 0] = self.peer_ptrs[r] + np.arange(len(ml), dtype=np.int64) * 16
                tab[ml, 1] = self.row_lens[r]
            self._mtab = torch.from_numpy(tab.reshape(-1)).to(self.device)
        return self._mtab

    def _stream_ptr(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream) if self.device.type == "cuda" else ctypes.c_void_p(0)

    @staticmethod
    def _ptrs(tensors):
        return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])

    def _slab_base_ptrs(self, slabs):
        """The FFT stages address full-map rows; a rank only holds its slab, so hand them the address the full map
        WOULD start at.  Only this rank's rows are ever touched."""
        a, _ = self.map_rows()
        out = []
        for s in slabs:
            if s.dtype != self.dtype or not s.is_contiguous() or s.numel() != self.nloc * self.band.nx:
                raise ValueError("map slab must be a contiguous %s tensor of nx * (local rows) elements" % self.dtype)
            out.append(s.data_ptr() - a * self.band.nx * s.element_size())
        return (ctypes.c_void_p * len(out))(*out)

    def _ev(self):
        if self.device.type != "cuda":
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record(torch.cuda.current_stream(self.device))
        return e

    def _barrier(self):
        """Orders the stages across ranks: everything the ranks enqueued before it is complete (and visible in peer
        memory) before anything enqueued after it starts.  Stream-ordered with NCCL (no host synchronisation)."""
        if self.world == 1:
            return
        if self.device.type == "cuda":
            dist.all_reduce(self._flag, group=self.group)
        else:
            dist.barrier(group=self.group)

    # ---- transforms ------------------------------------------------------------------------------------------
    def alm2map(self, d_alms, d_map_slabs):
        """d_alms: ncomp full-length complex alm tensors on the device (only this rank's m columns are read);
        d_map_slabs: ncomp real tensors with this rank's rows (map_rows()) of the column-major map, nx fastest."""
        nc = len(d_alms)
        L = self.lib.lib
        st = self._stream_ptr()
        mtab = self._m_table()
        if self.dtype != torch.float64:
            d_alms = [a.to(torch.complex128) for a in d_alms]
        self._barrier()          # every rank is done reading the previous contents of my phase buffer
        ev = [self._ev()]
        self.lib.check(L.pixsht_stage_alm2phase(self.handle, nc, self._ptrs(d_alms), self.nm, ctypes.c_void_p(self.d_m_list.data_ptr()),
                                                ctypes.c_void_p(self.own_ptr), self.row_len, st))
        ev.append(self._ev())
        self._barrier()          # every rank's m columns are complete
        ev.append(self._ev())
        self.lib.check(L.pixsht_stage_phase2map(self.handle, nc, ctypes.c_void_p(mtab.data_ptr()), self.r0, self.nloc,
                                                self._slab_base_ptrs(d_map_slabs), st))
        ev.append(self._ev())
        self._record("alm2map", ev)

    def map2alm(self, d_map_slabs, d_alms):
        """Inverse pipeline.  d_alms: ncomp full-length complex tensors; this rank's m columns receive the result,
        every other entry is zeroed (so the sum over ranks is the full alm)."""
        nc = len(d_alms)
        L = self.lib.lib
        st = self._stream_ptr()
        mtab = self._m_table()
        self._barrier()          # every rank is done with the previous contents of the phase buffers I am about to write
        ev = [self._ev()]
        self.lib.check(L.pixsht_stage_map2phase(self.handle, nc, self._slab_base_ptrs(d_map_slabs), self.r0, self.nloc,
                                                ctypes.c_void_p(mtab.data_ptr()), st))
        ev.append(self._ev())
        self._barrier()          # all rings of my m columns have arrived
        ev.append(self._ev())
        outs = d_alms if self.dtype == torch.float64 else [torch.empty(a.shape, dtype=torch.complex128, device=a.device) for a in d_alms]
        for t in outs:
            t.zero_()   # the analysis kernels accumulate atomically into pre-zeroed columns
        self.lib.check(L.pixsht_stage_phase2alm(self.handle, nc, ctypes.c_void_p(self.own_ptr), self.row_len, self.nm,
                                                ctypes.c_void_p(self.d_m_list.data_ptr()), self._ptrs(outs), st))
        if outs is not d_alms:
            for a, o in zip(d_alms, outs):
                a.copy_(o)
        ev.append(self._ev())
        self._record("map2alm", ev)

    # ---- host-resident data: the same pipelines with the PCIe copies overlapped, one spin family at a time ----------------
    def _families(self, nc):
        """[(spin0, spin2, components)] in processing order: T first, so that its stages run while Q/U (E/B) are on the wire."""
        return {1: [(1, 0, [0])], 2: [(0, 1, [0, 1])], 3: [(1, 0, [0]), (0, 1, [1, 2])]}[nc]

    def _host_state(self, nc):
        if getattr(self, "_hs", None) is None or self._hs["nc"] < nc:
            cuda = self.device.type == "cuda"
            idx = torch.cat([torch.arange(s, e, device=self.device) for (s, e) in self.alm_columns()]) if self.nm else torch.zeros(0, dtype=torch.long, device=self.device)
            self._hs = {"nc": nc, "idx": idx,
                        "pin": [torch.empty(idx.numel(), dtype=self.cdtype, device=self.device) for _ in range(nc)],
                        "pout": [torch.empty(idx.numel(), dtype=self.cdtype, device=self.device) for _ in range(nc)],
                        "sh": torch.cuda.Stream(self.device, priority=self._copy_prio) if cuda else None,
                        "sd": torch.cuda.Stream(self.device, priority=self._copy_prio) if cuda else None}
        return self._hs

    def _on(self, stream):
        import contextlib
        return torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()

    def _set_families(self, f0, f2):
        self.lib.check(self.lib.lib.pixsht_plan_set_stage_families(self.handle, int(f0), int(f2)))

    def _pieces(self, cuda):
        # Float32 data is widened as a whole before the Legendre stage: no pieces there
        return self.host_pieces if self.dtype == torch.float64 else 1

    def _m_pieces(self, K):
        """Contiguous slices [j0, j1) of this rank's m list with about equal Legendre work, and the offsets of the packed alm
        columns: column j occupies [off[j], off[j+1]) of a packed buffer."""
        L = self.lmax
        length = (L + 1 - self.my_m).astype(np.int64)
        off = np.concatenate([[0], np.cumsum(length)])
        if self.nm == 0:
            return [], off
        K = max(1, min(K, self.nm))
        edges = [int(np.searchsorted(off, off[-1] * k / K)) for k in range(K + 1)]
        edges[0], edges[-1] = 0, self.nm
        edges = np.maximum.accumulate(edges)
        return [(int(a), int(b)) for a, b in zip(edges[:-1], edges[1:]) if b > a], off

    def _ring_pieces(self, K):
        """This rank's rings in K contiguous sub-ranges (ra, rb) with the element slice of the slab that holds their rows."""
        K = max(1, min(K, self.nloc))
        edges = [self.r0 + (self.nloc * k) // K for k in range(K + 1)]
        a, _ = self.map_rows()
        out = []
        for ra, rb in zip(edges[:-1], edges[1:]):
            if rb <= ra:
                continue
            row0 = (self.nrings - rb) if self.band.flipy else ra
            out.append((ra, rb, slice((row0 - a) * self.band.nx, (row0 - a + rb - ra) * self.band.nx)))
        return out

    def alm2map_host(self, h_alm_cols, h_map_slabs, d_alms, d_map_slabs):
        """alm2map from / to (pinned) host memory.  h_alm_cols[c]: this rank's alm columns, packed in the order of
        alm_columns(); h_map_slabs[c]: receives this rank's map rows.  d_alms / d_map_slabs: device work tensors as for
        alm2map().  The copies run on two copy streams.  IQU goes polarisation first: its alm columns arrive in m ranges
        whose synthesis starts at once, T arrives under that work, the Q/U rows leave under the T stages, and only the T rows
        (the smaller output) leave after the last kernel."""
        nc = len(d_alms)
        hs = self._host_state(nc)
        L = self.lib.lib
        cuda = self.device.type == "cuda"
        sc = torch.cuda.current_stream(self.device) if cuda else None
        sh, sd = hs["sh"], hs["sd"]
        if cuda:
            sh.wait_stream(sc); sd.wait_stream(sc)
        fams = list(reversed(self._families(nc)))
        pieces, off = self._m_pieces(self._pieces(cuda))
        ev_in = []
        for k, (_, _, comps) in enumerate(fams):
            segs = pieces if k == 0 else [(0, self.nm)]
            evs = []
            for (j0, j1) in segs:
                seg = slice(int(off[j0]), int(off[j1]))
                with self._on(sh):
                    for c in comps:
                        hs["pin"][c][seg].copy_(h_alm_cols[c][seg], non_blocking=True)
                        d_alms[c].index_copy_(0, hs["idx"][seg], hs["pin"][c][seg])
                    evs.append(sh.record_event() if cuda else None)
            ev_in.append((segs, evs))
        mtab = self._m_table()
        st = self._stream_ptr()
        self._barrier()          # every rank is done reading the previous contents of my phase buffer
        try:
            for k, (f0, f2, comps) in enumerate(fams):
                src = list(d_alms)
                if self.dtype != torch.float64:
                    if cuda:
                        sc.wait_event(ev_in[k][1][-1])
                    for c in comps:
                        src[c] = d_alms[c].to(torch.complex128)
                self._set_families(f0, f2)
                for (j0, j1), ev in zip(*ev_in[k]):
                    if cuda:
                        sc.wait_event(ev)
                    self.lib.check(L.pixsht_stage_alm2phase(self.handle, nc, self._ptrs(src), j1 - j0,
                                                            ctypes.c_void_p(self.d_m_list.data_ptr() + 4 * j0),
                                                            ctypes.c_void_p(self.own_ptr + 16 * j0), self.row_len, st))
                self._barrier()      # every rank's m columns of this family are complete
                self.lib.check(L.pixsht_stage_phase2map(self.handle, nc, ctypes.c_void_p(mtab.data_ptr()), self.r0, self.nloc,
                                                        self._slab_base_ptrs(d_map_slabs), st))
                if cuda:
                    sd.wait_event(sc.record_event())
                with self._on(sd):
                    for c in comps:
                        h_map_slabs[c].copy_(d_map_slabs[c], non_blocking=True)
        finally:
            self._set_families(1, 1)
        if cuda:
            sc.wait_stream(sd)   # a synchronize on the caller's stream covers the copies

    def map2alm_host(self, h_map_slabs, h_alm_cols, d_map_slabs, d_alms):
        """map2alm from / to (pinned) host memory; arguments as for alm2map_host (h_alm_cols receives this rank's columns).
        T first: its rows arrive in ring ranges whose FFTs start at once; Q/U arrive under the T stages; the polarisation
        analysis runs in m ranges whose alm columns leave one by one."""
        nc = len(d_alms)
        hs = self._host_state(nc)
        L = self.lib.lib
        cuda = self.device.type == "cuda"
        sc = torch.cuda.current_stream(self.device) if cuda else None
        sh, sd = hs["sh"], hs["sd"]
        if cuda:
            sh.wait_stream(sc); sd.wait_stream(sc)
        fams = self._families(nc)
        K = self._pieces(cuda)
        ev_in = []
        for k, (_, _, comps) in enumerate(fams):
            segs = self._ring_pieces(K if k == 0 else 1)
            evs = []
            for (_, _, sl) in segs:
                with self._on(sh):
                    for c in comps:
                        d_map_slabs[c][sl].copy_(h_map_slabs[c][sl], non_blocking=True)
                    evs.append(sh.record_event() if cuda else None)
            ev_in.append((segs, evs))
        mtab = self._m_table()
        st = self._stream_ptr()
        mpieces, off = self._m_pieces(K)
        self._barrier()          # every rank is done with the previous contents of the phase buffers I am about to write
        try:
            for k, (f0, f2, comps) in enumerate(fams):
                self._set_families(f0, f2)
                for (ra, rb, _), ev in zip(*ev_in[k]):
                    if cuda:
                        sc.wait_event(ev)
                    self.lib.check(L.pixsht_stage_map2phase(self.handle, nc, self._slab_base_ptrs(d_map_slabs), ra, rb - ra,
                                                            ctypes.c_void_p(mtab.data_ptr()), st))
                self._barrier()      # all rings of my m columns of this family have arrived
                outs = list(d_alms)
                for c in comps:
                    if self.dtype != torch.float64:
                        outs[c] = torch.empty(d_alms[c].shape, dtype=torch.complex128, device=d_alms[c].device)
                    outs[c].zero_()
                last = k == len(fams) - 1
                for (j0, j1) in (mpieces if last else [(0, self.nm)]):
                    self.lib.check(L.pixsht_stage_phase2alm(self.handle, nc, ctypes.c_void_p(self.own_ptr + 16 * j0), self.row_len, j1 - j0,
                                                            ctypes.c_void_p(self.d_m_list.data_ptr() + 4 * j0), self._ptrs(outs), st))
                    seg = slice(int(off[j0]), int(off[j1]))
                    for c in comps:
                        if outs[c] is not d_alms[c]:
                            d_alms[c].copy_(outs[c])
                        hs["pout"][c][seg].copy_(d_alms[c].index_select(0, hs["idx"][seg]))
                    if cuda:
                        sd.wait_event(sc.record_event())
                    with self._on(sd):
                        for c in comps:
                            h_alm_cols[c][seg].copy_(hs["pout"][c][seg], non_blocking=True)
        finally:
            self._set_families(1, 1)
        if cuda:
            sc.wait_stream(sd)

    def _record(self, name, ev):
        self.last_ms[name] = ev

    def stage_ms(self, name):
        """Device milliseconds of the three stages of the last call (Legendre/FFT first, barrier, FFT/Legendre last)."""
        ev = self.last_ms.get(name)
        if not ev or ev[0] is None:
            return None
        ev[-1].synchronize()
        return [ev[i].elapsed_time(ev[i + 1]) for i in range(len(ev) - 1)]
