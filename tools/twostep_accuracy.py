#!/usr/bin/env python
"""Per-ring difference between the two-step and the standard spin-0 kernels (profiling aid): T alm2map at C3 / C4 size with
PIXSHT_TWOSTEP=1 against =0 on the same alm, rel-RMS per ring; and the map2alm of one map both ways, rel-RMS per m."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pixell.jl_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import pixsht
from pixsht.transforms import Plan
from helpers import synth_alm
res, lmax = (2.0, 5400) if (len(sys.argv) < 2 or sys.argv[1] == "C3") else (1.0, 10800)
shape, wcs = pixsht.fullsky_geometry(res * pixsht.arcminute)
band = pixsht.sht_band(shape, wcs)
alm = synth_alm(lmax, lmax, 123)
out = {}
for two in ("0", "1"):
    os.environ["PIXSHT_TWOSTEP"] = two
    p = Plan(band, lmax)
    m = p.alm2map([alm])[0]
    a = p.map2alm([out["0"][0] if two == "1" else m])[0]
    out[two] = (m, a)
    p.close()
m0, m1 = out["0"][0], out["1"][0]
rms = np.sqrt(np.mean(m0 ** 2, axis=0))
d = np.sqrt(np.mean((m1 - m0) ** 2, axis=0)) / np.maximum(rms, 1e-300)
idx = np.argsort(d)[::-1][:12]
print("alm2map two-step vs standard, rel-RMS per ring: whole map %.2e; worst rings:" % (np.sqrt(np.mean((m1 - m0) ** 2) / np.mean(m0 ** 2))))
print("  ", [(int(i), "%.1e" % d[i]) for i in idx])
print("   rings 0..8:", ["%.1e" % v for v in d[:9]])
n = d.size
print("   rings around the equator:", ["%.1e" % v for v in d[n // 2 - 3:n // 2 + 4]])
print("   quantiles 50/90/99/99.9 %:", ["%.1e" % np.quantile(d, q) for q in (0.5, 0.9, 0.99, 0.999)])
a0, a1 = out["0"][1], out["1"][1]
print("map2alm two-step vs standard: rel-RMS %.2e, max abs / rms %.2e" % (np.sqrt(np.sum(np.abs(a1 - a0) ** 2) / np.sum(np.abs(a0) ** 2)), np.max(np.abs(a1 - a0)) / np.sqrt(np.mean(np.abs(a0) ** 2))))
# per m
errs = []
for mm in (0, 1, 2, 10, 100, 1000, lmax // 2, lmax - 10, lmax):
    i0 = mm * (2 * lmax + 1 - mm) // 2 + mm; i1 = i0 + lmax - mm + 1
    errs.append((mm, "%.1e" % (np.sqrt(np.sum(np.abs(a1[i0:i1] - a0[i0:i1]) ** 2) / max(np.sum(np.abs(a0[i0:i1]) ** 2), 1e-300)))))
print("   per m:", errs)
