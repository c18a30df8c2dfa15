"""m-sharded multi-GPU pipeline: one process per GPU, torch.distributed for the plumbing.

  alm2map:  Legendre stage on this rank's m values (all rings)  ->  all-to-all (phase transpose)  ->  ring FFTs on this
            rank's contiguous slab of rings (all m)  ->  the rank's rows of the map.
  map2alm:  the same pipeline backwards.

Partition: m values are dealt out in load-balanced pairs (m, mmax-m) -- the Legendre cost of an m is ~ (lmax-m+1), so a
pair costs the same whichever it is; rings are split into contiguous slabs, which are contiguous row ranges of the
caller's column-major map.  Nothing else is partitioned; tables are replicated.

The compute stages are the pixsht_stage_* entry points of the C ABI; this module only owns the partition, the
pack/unpack of the exchange buffers and the collective (torch.distributed.all_to_all_single: NCCL over NVLink on GPUs,
gloo in the CPU tests, where the library handle is the host-emulation build).
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from ._lib import get_lib, Geom, F64


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


def partition_rings(nrings, world):
    """Contiguous slabs [begin, end) of band rings per rank."""
    edges = [(nrings * r) // world for r in range(world + 1)]
    return [(edges[r], edges[r + 1]) for r in range(world)]


class ShardedSHT:
    """Distributed map2alm / alm2map on one 8-GPU box.  All tensors live on `device` ('cuda:i', or 'cpu' with the
    emulation library in the gloo tests).  Float64 only (the BASELINE multi-GPU configs are Float64)."""

    def __init__(self, band, lmax, mmax=None, group=None, device=None, lib=None):
        self.lib = get_lib() if lib is None else lib
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.band, self.lmax = band, int(lmax)
        self.mmax = int(lmax if mmax is None else mmax)
        self.device = torch.device(device if device is not None else "cuda")
        dev_index = self.device.index if self.device.type == "cuda" and self.device.index is not None else 0
        g = Geom(band.nphi, band.nrings_total, band.ring_first, band.nrings, band.nx, int(band.flipx), int(band.flipy), 0, band.phi0)
        h = ctypes.c_void_p()
        self.lib.check(self.lib.lib.pixsht_plan_create(ctypes.byref(h), ctypes.byref(g), self.lmax, self.mmax, F64, dev_index))
        self.handle = h
        self.nalm = int(self.lib.lib.pixsht_nalm(self.lmax, self.mmax))
        self.nrings = band.nrings
        self.m_lists = partition_m(self.mmax, self.world)
        self.ring_ranges = partition_rings(self.nrings, self.world)
        self.my_m = self.m_lists[self.rank]
        self.nm = len(self.my_m)
        self.r0, self.r1 = self.ring_ranges[self.rank]
        self.nloc = self.r1 - self.r0
        # row of m in the exchanged buffer = position in the concatenation of all ranks' m lists
        m_row = np.empty(self.mmax + 1, dtype=np.int32)
        m_row[np.concatenate(self.m_lists)] = np.arange(self.mmax + 1, dtype=np.int32)
        self.d_m_list = torch.from_numpy(self.my_m.copy()).to(self.device)
        self.d_m_row = torch.from_numpy(m_row).to(self.device)
        self._bufs = {}
        self.last_ms = {}

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
            self.lib.lib.pixsht_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers -----------------------------------------------------------------------------------------------
    def _buf(self, name, n):
        b = self._bufs.get(name)
        if b is None or b.numel() < n:
            b = self._bufs[name] = torch.empty(n, dtype=torch.complex128, device=self.device)
        return b[:n]

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
            if s.dtype != torch.float64 or not s.is_contiguous() or s.numel() != self.nloc * self.band.nx:
                raise ValueError("map slab must be a contiguous float64 tensor of nx * (local rows) elements")
            out.append(s.data_ptr() - a * self.band.nx * 8)
        return (ctypes.c_void_p * len(out))(*out)

    def _ev(self):
        if self.device.type != "cuda":
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record(torch.cuda.current_stream(self.device))
        return e

    def _exchange(self, send, out, in_splits, out_splits):
        if self.world == 1:
            out.copy_(send)
        else:
            dist.all_to_all_single(torch.view_as_real(out), torch.view_as_real(send), [s * 1 for s in out_splits],
                                   [s * 1 for s in in_splits], group=self.group)

    # ---- transforms ------------------------------------------------------------------------------------------
    def alm2map(self, d_alms, d_map_slabs):
        """d_alms: ncomp full-length complex128 alm tensors on the device (only this rank's m columns are read);
        d_map_slabs: ncomp float64 tensors with this rank's rows (map_rows()) of the column-major map, nx fastest."""
        nc = len(d_alms)
        L = self.lib.lib
        st = self._stream_ptr()
        nr = self.nrings
        ev = [self._ev()]
        legbuf = self._buf("leg", self.nm * nc * nr)
        self.lib.check(L.pixsht_stage_alm2phase(self.handle, nc, self._ptrs(d_alms), self.nm, ctypes.c_void_p(self.d_m_list.data_ptr()),
                                                ctypes.c_void_p(legbuf.data_ptr()), st))
        ev.append(self._ev())
        # pack: [mi][c][ring] -> per destination h: [mi][c][ring in slab h]
        v = legbuf.view(self.nm, nc, nr)
        send = self._buf("send", self.nm * nc * nr)
        in_splits, off = [], 0
        for (a, b) in self.ring_ranges:
            n = self.nm * nc * (b - a)
            send[off:off + n].view(self.nm, nc, b - a).copy_(v[:, :, a:b])
            in_splits.append(n)
            off += n
        out_splits = [len(ml) * nc * self.nloc for ml in self.m_lists]
        recv = self._buf("recv", (self.mmax + 1) * nc * self.nloc)
        self._exchange(send, recv, in_splits, out_splits)
        ev.append(self._ev())
        self.lib.check(L.pixsht_stage_phase2map(self.handle, nc, ctypes.c_void_p(recv.data_ptr()), ctypes.c_void_p(self.d_m_row.data_ptr()),
                                                self.r0, self.nloc, self._slab_base_ptrs(d_map_slabs), st))
        ev.append(self._ev())
        self._record("alm2map", ev)

    def map2alm(self, d_map_slabs, d_alms):
        """Inverse pipeline.  d_alms: ncomp full-length complex128 tensors; this rank's m columns receive the result,
        every other entry is zeroed (so the sum over ranks is the full alm)."""
        nc = len(d_alms)
        L = self.lib.lib
        st = self._stream_ptr()
        nr = self.nrings
        ev = [self._ev()]
        fftbuf = self._buf("recv", (self.mmax + 1) * nc * self.nloc)
        self.lib.check(L.pixsht_stage_map2phase(self.handle, nc, self._slab_base_ptrs(d_map_slabs), ctypes.c_void_p(self.d_m_row.data_ptr()),
                                                self.r0, self.nloc, ctypes.c_void_p(fftbuf.data_ptr()), st))
        ev.append(self._ev())
        in_splits = [len(ml) * nc * self.nloc for ml in self.m_lists]           # rows of rank g are a contiguous block
        out_splits = [self.nm * nc * (b - a) for (a, b) in self.ring_ranges]
        recv = self._buf("send", self.nm * nc * nr)
        self._exchange(fftbuf, recv, in_splits, out_splits)
        # unpack: from each source h [mi][c][ring in slab h] -> [mi][c][ring]
        legbuf = self._buf("leg", self.nm * nc * nr)
        v = legbuf.view(self.nm, nc, nr)
        off = 0
        for (a, b), n in zip(self.ring_ranges, out_splits):
            v[:, :, a:b].copy_(recv[off:off + n].view(self.nm, nc, b - a))
            off += n
        ev.append(self._ev())
        for t in d_alms:
            t.zero_()   # the analysis kernels accumulate atomically into pre-zeroed columns
        self.lib.check(L.pixsht_stage_phase2alm(self.handle, nc, ctypes.c_void_p(legbuf.data_ptr()), self.nm,
                                                ctypes.c_void_p(self.d_m_list.data_ptr()), self._ptrs(d_alms), st))
        ev.append(self._ev())
        self._record("map2alm", ev)

    def _record(self, name, ev):
        self.last_ms[name] = ev

    def stage_ms(self, name):
        """Device milliseconds of the three stages of the last call (Legendre/FFT first, exchange, FFT/Legendre last)."""
        ev = self.last_ms.get(name)
        if not ev or ev[0] is None:
            return None
        ev[-1].synchronize()
        return [ev[i].elapsed_time(ev[i + 1]) for i in range(len(ev) - 1)]
