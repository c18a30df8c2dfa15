#!/usr/bin/env python
"""One-process multi-GPU end-to-end probe (profiling aid): the blocking pixsht_execute call on whole HOST arrays allocated through
pixsht_host_alloc, on a pixsht_plan_create_multi plan over every GPU of the box -- what the Julia binding does.  Prints wall time
and the device span of each direction.  usage: tools/multi_e2e.py [C3|C4] [steps]   (env: PROBE_NDEV, PIXSHT_MULTI_PIECES,
PIXSHT_MULTI_SEGS, PIXSHT_HOST_NUMA)"""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pixell.jl_b200")]
import numpy as np
import pixsht
from pixsht.transforms import Plan, get_lib, ALM2MAP, MAP2ALM, HOST

wl = sys.argv[1] if len(sys.argv) > 1 else "C4"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
res, lmax = {"C3": (2.0, 5400), "C4": (1.0, 10800)}[wl]
lib = get_lib()
ndev = int(os.environ.get("PROBE_NDEV", lib.device_count()))
shape, wcs = pixsht.fullsky_geometry(res * pixsht.arcminute)
band = pixsht.sht_band(shape, wcs)
t0 = time.perf_counter()
plan = Plan(band, lmax, devices=list(range(ndev)))
t_plan = time.perf_counter() - t0
nalm, nc, npix = plan.nalm, 3, band.nx * band.nrings


def host(nbytes):
    p = ctypes.c_void_p()
    lib.check(lib.lib.pixsht_host_alloc(ctypes.byref(p), nbytes))
    return p


pa = [host(nalm * 16) for _ in range(nc)]
po = [host(nalm * 16) for _ in range(nc)]
pm = [host(npix * 8) for _ in range(nc)]
rng = np.random.default_rng(7)
for c, p in enumerate(pa):
    a = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_double)), shape=(2 * nalm,))
    a[:] = rng.standard_normal(2 * nalm)
    a[1:2 * (lmax + 1):2] = 0.0            # m = 0: real
    if c > 0:
        a[:4] = 0.0; a[2 * (lmax + 1):2 * (lmax + 1) + 2] = 0.0
A, O, M = [p.value for p in pa], [p.value for p in po], [p.value for p in pm]
best = [1e30, 1e30]; span = [0, 0]
for it in range(steps + 1):
    t0 = time.perf_counter(); plan.execute_ptrs(ALM2MAP, A, M, HOST); t1 = time.perf_counter(); s1 = plan.timings()["compute_span"]
    plan.execute_ptrs(MAP2ALM, O, M, HOST); t2 = time.perf_counter(); s2 = plan.timings()["compute_span"]
    if it:
        if t1 - t0 < best[0]: best[0], span[0] = t1 - t0, s1
        if t2 - t1 < best[1]: best[1], span[1] = t2 - t1, s2
a0 = np.ctypeslib.as_array(ctypes.cast(pa[0], ctypes.POINTER(ctypes.c_double)), shape=(2 * nalm,))
o0 = np.ctypeslib.as_array(ctypes.cast(po[0], ctypes.POINTER(ctypes.c_double)), shape=(2 * nalm,))
rt = float(np.sqrt(np.sum((a0[:20000] - o0[:20000]) ** 2) / np.sum(a0[:20000] ** 2)))
print("%s ndev=%d pieces=%s segs=%s numa=%s: e2e %.1f ms (alm2map %.1f [device span %.1f], map2alm %.1f [%.1f]); plan %.1f s; round trip T m=0 rel %.1e" % (
    wl, ndev, os.environ.get("PIXSHT_MULTI_PIECES", "3"), os.environ.get("PIXSHT_MULTI_SEGS", "4"), os.environ.get("PIXSHT_HOST_NUMA", "interleave"),
    1e3 * (best[0] + best[1]), 1e3 * best[0], span[0], 1e3 * best[1], span[1], t_plan, rt))
plan.close()
for p in pa + po + pm:
    lib.lib.pixsht_host_free(p)
