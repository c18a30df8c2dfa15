#!/usr/bin/env python
"""One-process multi-GPU plan probe (profiling aid): C3 or C4 through pixsht_plan_create_multi on every GPU of the box, data
already sharded on the devices (pixsht_execute_sharded), a few steps.  Run plain, then under
  ncu --metrics gpu__time_duration.sum,nvlrx__bytes.sum,nvltx__bytes.sum -k regex:fft_ ...
to read the NVLink bytes the FFT kernels' fused transpose moves (profiles/r02)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pixell.jl_b200")]
import numpy as np
import torch
import pixsht
from pixsht.transforms import Plan, get_lib, ALM2MAP, MAP2ALM

wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
res, lmax = {"C3": (2.0, 5400), "C4": (1.0, 10800)}[wl]
ndev = int(os.environ.get("PROBE_NDEV", torch.cuda.device_count()))
shape, wcs = pixsht.fullsky_geometry(res * pixsht.arcminute)
band = pixsht.sht_band(shape, wcs)
plan = Plan(band, lmax, devices=list(range(ndev)))
nalm, nc = plan.nalm, 3
sa, sm, so = [], [], []
for (dv, r0, nr, ml) in plan.shards():
    td = torch.device("cuda", dv)
    g = torch.Generator(device=td); g.manual_seed(7)
    sa.append([torch.view_as_complex(torch.randn(nalm, 2, generator=g, device=td, dtype=torch.float64)) for _ in range(nc)])
    sm.append([torch.empty(nr * band.nx, dtype=torch.float64, device=td) for _ in range(nc)])
    so.append([torch.empty(nalm, dtype=torch.complex128, device=td) for _ in range(nc)])
flat = lambda x: [t.data_ptr() for per in x for t in per]
fa, fm, fo = flat(sa), flat(sm), flat(so)
for it in range(steps):
    t0 = time.perf_counter()
    plan.execute_sharded_ptrs(ALM2MAP, nc, fa, fm); a = plan.timings()["compute_span"]
    plan.execute_sharded_ptrs(MAP2ALM, nc, fo, fm); b = plan.timings()["compute_span"]
    print("%s ndev=%d step %d: wall %.2f ms, device spans alm2map %.2f + map2alm %.2f ms" % (wl, ndev, it, 1e3 * (time.perf_counter() - t0), a, b))
plan.close()
