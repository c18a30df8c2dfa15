#!/usr/bin/env python
"""Inner-loop efficiency probe: time the Legendre kernels on the NM lowest m and the CHUNKS equator-most chunks only
(PIXSHT_DBG_NM / PIXSHT_DBG_CHUNKS), where every ring is active over ~the whole l range, and compare with the DFMA peak.
Needs a probe build of the library: PIXSHT_NVCC_EXTRA=-DPIXSHT_PROBES bash pixell.jl_b200/build.sh (never ship that build)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pixell.jl_b200")]
import torch, pixsht
from pixsht.transforms import Plan, get_lib, MAP2ALM, ALM2MAP, DEVICE
nm, nch = int(sys.argv[1]), int(sys.argv[2])
res = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
os.environ["PIXSHT_DBG_NM"] = str(nm); os.environ["PIXSHT_DBG_CHUNKS"] = str(nch)
lib = get_lib()
shape, wcs = pixsht.fullsky_geometry(res * pixsht.arcminute)
band = pixsht.sht_band(shape, wcs); lmax = band.nphi // 2
plan = Plan(band, lmax)
info = plan.info()
dev = torch.device("cuda", 0)
peak, _ = lib.measure_fma_peak(0)
for nc in (1, 2):
    alm = [torch.randn(plan.nalm, dtype=torch.complex128, device=dev) for _ in range(nc)]
    mp = [torch.randn(band.nx * band.nrings, dtype=torch.float64, device=dev) for _ in range(nc)]
    for d, name in ((ALM2MAP, "synth"), (MAP2ALM, "anal")):
        R = info["R0" if nc == 1 else "R2"] if d == ALM2MAP else info["R0a" if nc == 1 else "R2a"]
        pairs = min(nch * 32 * R, info["npairs"])
        steps = sum(lmax - max(m, 0 if nc == 1 else 2) + 1 for m in range(nm))
        flop = 2.0 * (4 if nc == 1 else 12) * steps * pairs
        for _ in range(3):
            plan.execute_ptrs(d, [a.data_ptr() for a in alm], [m.data_ptr() for m in mp], DEVICE)
        t = plan.timings()["legendre"]
        print("%s spin%d R=%d: %.3f ms  %.2f TF = %.1f%% of DFMA peak %.1f" % (name, 0 if nc == 1 else 2, R, t, flop / t / 1e9, 100 * flop / t / 1e9 / peak, peak))
