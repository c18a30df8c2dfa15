#!/usr/bin/env python
"""Per-ring accuracy of alm2map near the pole at a BASELINE size: GPU (FP64 recurrences in x = cos(theta)) against the
long-double oracle on selected rings.  White-noise alm up to lmax is the worst case (no beam / C_l fall-off)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pixell.jl_b200"), os.path.join(ROOT, "tests")]
import numpy as np, pixsht
from pixsht.transforms import Plan
from helpers import synth_alm
from oracle import get_oracle, cc_geometry
res = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
shape, wcs = pixsht.fullsky_geometry(res * pixsht.arcminute)
band = pixsht.sht_band(shape, wcs); lmax = band.nphi // 2
plan = Plan(band, lmax)
theta, _ = cc_geometry(band.nrings_total, band.nphi)
rings = [0, 1, 2, 3, 5, 10, 30, 100, 300, 1000, band.nrings // 4, band.nrings // 2, band.nrings - 2, band.nrings - 1]
orc = get_oracle("ld")
for spin, alms in ((0, [synth_alm(lmax, lmax, 4000)]), (2, [synth_alm(lmax, lmax, 4001, True), synth_alm(lmax, lmax, 4002, True)])):
    maps = plan.alm2map(alms)
    ref = orc.alm2map(np.stack(alms), theta[rings], band.phi0, band.nphi, lmax, spin=spin)   # (ncomp, len(rings), nphi), band orientation
    tot_num = tot_den = 0.0
    for i, r in enumerate(rings):
        row = (band.nrings - 1 - r) if band.flipy else r
        num = den = 0.0
        for c in range(len(alms)):
            g = maps[c][:, row]; g = g[::-1] if band.flipx else g
            num += float(np.sum((g - ref[c, i]) ** 2)); den += float(np.sum(ref[c, i] ** 2))
        print("spin %d ring %5d theta %9.5f deg  rel rms %.2e  (rms value %.3g)" % (spin, r, np.degrees(theta[r]), np.sqrt(num / max(den, 1e-300)), np.sqrt(den / band.nphi / len(alms))))
