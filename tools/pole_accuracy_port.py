#!/usr/bin/env python
"""CPU only: per-ring accuracy near the pole of the libsharp2-style FP64 CPU port (oracle/sht_cpu.c: three-term recurrences in
x = cos(theta), scaled seek, the algorithm libsharp2 runs) against the long-double oracle, at a BASELINE size.  Companion of
tools/pole_accuracy.py (the same rings on the GPU): shows that the loss on the last few rings before a pole belongs to the FP64
recurrence, not to the CUDA engine.  Usage: pole_accuracy_port.py [res_arcmin=1]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pixell.jl_b200"), os.path.join(ROOT, "tests")]
import numpy as np, pixsht
from helpers import synth_alm
from oracle import get_oracle, get_cpu_sht, cc_geometry
res = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
shape, wcs = pixsht.fullsky_geometry(res * pixsht.arcminute)
band = pixsht.sht_band(shape, wcs); lmax = band.nphi // 2
theta, _ = cc_geometry(band.nrings_total, band.nphi)
rings = [0, 1, 2, 3, 5, 10, 30, 100, 300, 1000, band.nrings // 4, band.nrings // 2, band.nrings - 2, band.nrings - 1]
orc, cpu = get_oracle("ld"), get_cpu_sht()
cpu.use_all_cores()
for spin, alms in ((0, [synth_alm(lmax, lmax, 4000)]), (2, [synth_alm(lmax, lmax, 4001, True), synth_alm(lmax, lmax, 4002, True)])):
    ref = orc.alm2map(np.stack(alms), theta[rings], band.phi0, band.nphi, lmax, spin=spin)
    got = cpu.alm2map(np.stack(alms), theta[rings], band.phi0, band.nphi, lmax, spin=spin)
    for i, r in enumerate(rings):
        num = float(np.sum((got[:, i] - ref[:, i]) ** 2)); den = float(np.sum(ref[:, i] ** 2))
        print("spin %d ring %5d theta %9.5f deg  CPU port (FP64) vs long double: rel rms %.2e" % (spin, r, np.degrees(theta[r]), np.sqrt(num / max(den, 1e-300))))
