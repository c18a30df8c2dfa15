#!/usr/bin/env python
"""Where the host-pointer call spends its time (one GPU, pinned arrays): wall time of the call, busy span of the compute stream
(from its first wait to its last kernel), and the device-resident time of the same transform.  wall - span = exposed head / tail
copies; span - device = gaps and launch tails of the pieces.  Usage: e2e_probe.py [C4|C3] [PIXSHT_SPLITS values...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pixell.jl_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch, pixsht
from pixsht.transforms import Plan
from pixsht import _lib
wl = sys.argv[1] if len(sys.argv) > 1 else "C4"
res, lmax = {"C4": (1.0, 10800), "C3": (2.0, 5400)}[wl]
shape, wcs = pixsht.fullsky_geometry(res * pixsht.arcminute)
band = pixsht.sht_band(shape, wcs)
for splits in (sys.argv[2:] or ["8"]):
    os.environ["PIXSHT_SPLITS"] = splits
    plan = Plan(band, lmax)
    nalm = plan.nalm
    g = torch.Generator().manual_seed(1)
    h_alm = [torch.randn(nalm, dtype=torch.complex128, generator=g).pin_memory() for _ in range(3)]
    h_out = [torch.empty(nalm, dtype=torch.complex128).pin_memory() for _ in range(3)]
    h_map = [torch.empty(band.nx * band.nrings, dtype=torch.float64).pin_memory() for _ in range(3)]
    d_alm = [a.cuda() for a in h_alm]; d_map = [torch.empty_like(m, device="cuda") for m in h_map]; d_out = [torch.empty_like(a, device="cuda") for a in h_alm]
    P = lambda ts: [t.data_ptr() for t in ts]
    for rep in range(3):
        plan.execute_ptrs(_lib.ALM2MAP, P(h_alm), P(h_map)); t1 = plan.timings()
        plan.execute_ptrs(_lib.MAP2ALM, P(h_out), P(h_map)); t2 = plan.timings()
        plan.execute_ptrs(_lib.ALM2MAP, P(d_alm), P(d_map), _lib.DEVICE); torch.cuda.synchronize(); u1 = plan.timings()
        plan.execute_ptrs(_lib.MAP2ALM, P(d_out), P(d_map), _lib.DEVICE); torch.cuda.synchronize(); u2 = plan.timings()
    print("%s splits=%s" % (wl, splits))
    for name, t, u in (("alm2map", t1, u1), ("map2alm", t2, u2)):
        print("  %s: host call %.1f ms, compute-stream span %.1f, device-resident call %.1f (legendre %.1f fft %.1f)  -> head/tail %.1f, gaps %.1f"
              % (name, t["total"], t["compute_span"], u["total"], u["legendre"], u["fft"], t["total"] - t["compute_span"], t["compute_span"] - u["total"]))
    plan.close()
