"""Parity tests proper: the nvcc-built sm_100a library, called through the C ABI (ctypes), against the CPU oracle and
the reference's golden alm files.  Need a B200: run with `-m gpu`.

Tolerances (BASELINE.json north_star): relative RMS error <= 1e-10 in Float64 and <= 1e-5 in Float32.
Round trips are compared with the ORACLE's round trip, not with identity (plain CC weights are not exact
at lmax = nphi/2)."""
import ctypes
import math

import numpy as np
import pytest

import pixsht
from pixsht import Enmap, Alm, CarClenshawCurtis, fullsky_geometry, geometry, degree, arcminute
from pixsht.transforms import Plan, map2alm, alm2map, get_lib, PixshtError
from pixsht import _lib
from helpers import (golden_alm, gen_spin0, gen_spin2, oracle_map2alm, oracle_alm2map, rel_rms, synth_alm, band_copy)
from oracle import cc_geometry, get_oracle, nalm, alm_index

pytestmark = pytest.mark.gpu
TOL64, TOL32 = 1e-10, 1e-5


def relmax(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def test_library_is_the_cuda_build():
    lib = get_lib()
    assert "sm_100a" in lib.version() and "EMULATION" not in lib.version()
    assert lib.device_count() >= 1


# ---- the reference's own tests (test/test_transforms.jl:11-77) --------------------------------------------------
def test_golden_spin0_fullsky_sliced_box():
    shape, wcs = fullsky_geometry(10.0 * degree)
    m = Enmap(gen_spin0(shape), wcs)
    assert relmax(map2alm(m, lmax=18).alm, golden_alm("simple_analytic_sht")) < 1e-12
    assert len(map2alm(m).alm) == 190
    assert relmax(map2alm(m[5:-2, 4:-3], lmax=18).alm, golden_alm("simple_analytic_sht_sliced")) < 1e-12
    box = [[10 * degree, -10 * degree], [-5 * degree, 5 * degree]]
    bshape, bwcs = geometry(CarClenshawCurtis, box, 1.0 * degree)
    mb = Enmap(gen_spin0(bshape, 2.5), bwcs)
    assert relmax(map2alm(mb, lmax=100).alm, golden_alm("simple_box_analytic_sht")) < 1e-12


def test_golden_spin2_stack_tuple_iqu():
    shape, wcs = fullsky_geometry(10.0 * degree, dims=(2,))
    qu = Enmap(gen_spin2(shape), wcs)
    ref_e, ref_b = golden_alm("simple_pol_analytic_sht", (0, 1)), golden_alm("simple_pol_analytic_sht", (2, 3))
    e, b = map2alm(qu, lmax=3 * 36)
    assert relmax(e.alm, ref_e) < 1e-12 and relmax(b.alm, ref_b) < 1e-12
    e, b = map2alm((Enmap(qu.data[:, :, 0], wcs), Enmap(qu.data[:, :, 1], wcs)), lmax=3 * 36)
    assert relmax(e.alm, ref_e) < 1e-12 and relmax(b.alm, ref_b) < 1e-12
    shape3, wcs3 = fullsky_geometry(10.0 * degree, dims=(3,))
    d = np.zeros(shape3, order="F")
    d[:, :, 0] = gen_spin0(shape3)
    d[:, :, 1:] = gen_spin2(shape3)
    t, e, b = map2alm(Enmap(d, wcs3), lmax=3 * 36)
    assert relmax(t.alm, golden_alm("simple_analytic_sht_fullalm")) < 1e-12
    assert relmax(e.alm, ref_e) < 1e-12 and relmax(b.alm, ref_b) < 1e-12
    t2, e2, b2 = map2alm(tuple(Enmap(d[:, :, c], wcs3) for c in range(3)), lmax=3 * 36)
    assert np.array_equal(t2.alm, t.alm) or relmax(t2.alm, t.alm) < 1e-14


def test_bad_ncomp_is_an_error_not_a_crash():
    shape, wcs = fullsky_geometry(10.0 * degree, dims=(4,))
    with pytest.raises(ValueError):
        map2alm(Enmap(np.zeros(shape, order="F"), wcs), lmax=18)
    lib = get_lib()
    p = Plan(pixsht.sht_band((36, 19), wcs), 18)
    bad = (ctypes.c_void_p * 4)(1, 1, 1, 1)
    rc = lib.lib.pixsht_execute(p.handle, 0, 4, bad, bad, 0)
    assert rc == _lib.ERR_ARG and b"1 <= ncomp <= 3" in lib.lib.pixsht_last_error()
    p.close()


# ---- BASELINE config C1: 360 x 181 Float64 spin-0 round trip, lmax 180 -----------------------------------------
def test_c1_spin0_f64_both_directions_and_roundtrip():
    shape, wcs = fullsky_geometry(1.0 * degree)
    lmax = 180
    alm = synth_alm(lmax, lmax, 1000)
    ref_map = oracle_alm2map(alm[None], shape, wcs, lmax)[:, :, 0]
    got = alm2map(Alm(lmax, lmax, alm), shape, wcs)
    assert rel_rms(got.data, ref_map) < TOL64
    ref_alm = oracle_map2alm(Enmap(ref_map, wcs), lmax)[0]
    got_alm = map2alm(Enmap(ref_map, wcs), lmax=lmax).alm
    assert rel_rms(got_alm, ref_alm) < TOL64
    # round trip vs the oracle's round trip
    rt = map2alm(got, lmax=lmax).alm
    assert rel_rms(rt, ref_alm) < TOL64
    # white-noise (not band-limited) map
    rng = np.random.default_rng(2001)
    noise = Enmap(np.asfortranarray(rng.standard_normal(shape)), wcs)
    assert rel_rms(map2alm(noise, lmax=lmax).alm, oracle_map2alm(noise, lmax)[0]) < TOL64


@pytest.mark.parametrize("res_deg,lmax,mmax", [(0.5, 360, 360), (0.5, 300, 200), (0.25, 720, 720)])
def test_iqu_f64_vs_oracle(res_deg, lmax, mmax):
    shape, wcs = fullsky_geometry(res_deg * degree, dims=(3,))
    alms = [synth_alm(lmax, mmax, 3000 + c, spin2=c > 0) for c in range(3)]
    ref = np.concatenate([oracle_alm2map(alms[0][None], shape, wcs, lmax, mmax),
                          oracle_alm2map(np.stack(alms[1:]), shape, wcs, lmax, mmax, spin=2)], axis=2)
    got = alm2map(tuple(Alm(lmax, mmax, a) for a in alms), shape, wcs)
    assert isinstance(got, tuple) and len(got) == 3
    for c in range(3):
        assert rel_rms(got[c].data, ref[:, :, c]) < TOL64
    m = Enmap(np.asfortranarray(ref), wcs)
    t, e, b = map2alm(m, lmax=lmax, mmax=mmax)
    rt = oracle_map2alm(Enmap(ref[:, :, 0], wcs), lmax, mmax)[0]
    reb = oracle_map2alm(Enmap(ref[:, :, 1:], wcs), lmax, mmax, spin=2)
    assert rel_rms(t.alm, rt) < TOL64 and rel_rms(e.alm, reb[0]) < TOL64 and rel_rms(b.alm, reb[1]) < TOL64


def test_partial_sky_band_and_unflipped_geometry():
    # a cut-sky band sliced out of a full-sky grid, and a geometry with ascending RA / descending DEC (no flips)
    shape, wcs = fullsky_geometry(1.0 * degree)
    full = Enmap(gen_spin0(shape, 1.5), wcs)
    sub = full[40:300, 30:120]
    lmax = 150
    assert rel_rms(map2alm(sub, lmax=lmax).alm, oracle_map2alm(sub, lmax)[0]) < TOL64
    alm = synth_alm(lmax, lmax, 77)
    got = alm2map(Alm(lmax, lmax, alm), sub.shape, sub.wcs)
    assert rel_rms(got.data, oracle_alm2map(alm[None], sub.shape, sub.wcs, lmax)[:, :, 0]) < TOL64
    flipped = full[::-1, ::-1]
    assert rel_rms(map2alm(flipped, lmax=lmax).alm, oracle_map2alm(flipped, lmax)[0]) < TOL64
    got = alm2map(Alm(lmax, lmax, alm), flipped.shape, flipped.wcs)
    assert rel_rms(got.data, oracle_alm2map(alm[None], flipped.shape, flipped.wcs, lmax)[:, :, 0]) < TOL64


def test_float32_maps():
    shape, wcs = fullsky_geometry(0.5 * degree)
    lmax = 360
    alm = synth_alm(lmax, lmax, 2000)
    ref = oracle_alm2map(alm[None], shape, wcs, lmax)[:, :, 0]
    got = alm2map(Alm(lmax, lmax, alm), shape, wcs, dtype=np.float32)
    assert got.dtype == np.float32 and rel_rms(got.data, ref) < TOL32
    a32 = map2alm(Enmap(np.asfortranarray(ref, dtype=np.float32), wcs), lmax=lmax)
    assert a32.alm.dtype == np.complex64
    assert rel_rms(a32.alm, oracle_map2alm(Enmap(ref, wcs), lmax)[0]) < TOL32


def test_adjointness_and_linearity_midsize():
    """size-independent properties at a size the full oracle would be slow for: <W Y a, m> = <a, Y^T W m>, linearity."""
    shape, wcs = fullsky_geometry(4.0 * arcminute)   # 5400 x 2701 (geometry of BASELINE config C2)
    lmax = 2700
    band = pixsht.sht_band(shape, wcs)
    plan = Plan(band, lmax)
    w, _ = plan.weights()
    rng = np.random.default_rng(9)
    a1, a2 = synth_alm(lmax, lmax, 91), synth_alm(lmax, lmax, 92)
    m1, m2 = plan.alm2map([a1])[0], plan.alm2map([a2])[0]
    m12 = plan.alm2map([a1 + 0.5 * a2])[0]
    assert rel_rms(m12, m1 + 0.5 * m2) < 1e-12
    x = np.asfortranarray(rng.standard_normal(shape))
    ytx = plan.map2alm([x])[0]
    wrow = w[::-1] if band.flipy else w
    lhs = float(np.sum(m1 * x * wrow[None, :]))
    fac = np.full(a1.shape, 2.0)
    fac[:lmax + 1] = 1.0
    rhs = float(np.sum(fac * (np.conj(a1) * ytx).real))
    assert abs(lhs - rhs) < 1e-10 * abs(lhs)
    # sampled rings / sampled m against the oracle at this size
    ref = oracle_alm2map(a1[None], shape, wcs, lmax, ring_stride=337, ring_offset=5)[:, :, 0]
    rows = np.arange(band.nrings)[5::337]
    rows_map = (band.nrings - 1 - rows) if band.flipy else rows
    assert rel_rms(m1[:, rows_map], ref[:, rows_map]) < TOL64
    ref_alm = oracle_map2alm(Enmap(x, wcs), lmax, m_stride=451, m_offset=7)[0]
    sel = np.concatenate([np.arange(alm_index(lmax, m, m), alm_index(lmax, lmax, m) + 1) for m in range(7, lmax + 1, 451)])
    assert rel_rms(ytx[sel], ref_alm[sel]) < TOL64
    plan.close()


def test_sharp_shim_runs_the_reference_call_sequence():
    """The ccall sequence of src/transforms.jl:33-63,88-108 against the libsharp2-compatible symbols."""
    lib = get_lib()
    L = lib.lib
    shape, wcs = fullsky_geometry(10.0 * degree)
    m = Enmap(gen_spin0(shape), wcs)
    band, b = band_copy(m)
    theta, w = cc_geometry(b.nrings_total, b.nphi, b.ring_first, b.nrings)
    n = b.nrings
    nph = (ctypes.c_int * n)(*([b.nphi] * n))
    ofs = (ctypes.c_ssize_t * n)(*[b.nphi * i for i in range(n)])
    stride = (ctypes.c_int * n)(*([1] * n))
    phi0 = (ctypes.c_double * n)(*([b.phi0] * n))
    th = (ctypes.c_double * n)(*theta)
    wg = (ctypes.c_double * n)(*w)
    geom, ainfo = ctypes.c_void_p(), ctypes.c_void_p()
    L.sharp_make_geom_info(n, nph, ofs, stride, phi0, th, wg, ctypes.byref(geom))
    L.sharp_make_triangular_alm_info(18, 18, 1, ctypes.byref(ainfo))
    assert L.sharp_alm_count(ainfo) == 190 and L.sharp_map_size(geom) == 36 * 19
    alm = np.zeros(190, dtype=np.complex128)
    flat = np.ascontiguousarray(band[0]).ravel()
    SHARP_MAP2ALM, SHARP_ALM2MAP, SHARP_DP = 0, 1, 1 << 4
    L.sharp_execute(SHARP_MAP2ALM, 0, (ctypes.c_void_p * 1)(alm.ctypes.data), (ctypes.c_void_p * 1)(flat.ctypes.data), geom, ainfo,
                    SHARP_DP, None, None)
    assert L.pixsht_shim_status() == 0
    assert relmax(alm, golden_alm("simple_analytic_sht")) < 1e-12
    back = np.zeros_like(flat)
    L.sharp_execute(SHARP_ALM2MAP, 0, (ctypes.c_void_p * 1)(alm.ctypes.data), (ctypes.c_void_p * 1)(back.ctypes.data), geom, ainfo,
                    SHARP_DP, None, None)
    th_full, _ = cc_geometry(b.nrings_total, b.nphi)
    ref = get_oracle("ld").alm2map(alm[None], th_full, b.phi0, b.nphi, 18)[0].ravel()
    assert rel_rms(back, ref) < TOL64
    L.sharp_destroy_alm_info(ainfo)
    L.sharp_destroy_geom_info(geom)
