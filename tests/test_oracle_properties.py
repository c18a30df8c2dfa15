"""Oracle self-consistency beyond the golden files (CPU only): lambda generators against brute-force Wigner-d sums,
adjointness <W Y a, m> = <a, Y^T W m> (pins alm2map to the golden-pinned map2alm), exact round trip when
2 lmax <= nrings-1, and agreement of the double build with the long-double build."""
import math
import numpy as np
import pytest

from pixsht import Enmap, fullsky_geometry, degree, sht_band
from oracle import get_oracle, cc_geometry, nalm, alm_index
from helpers import synth_alm, rel_rms


def wigner_d(j, mp, m, beta):
    """Explicit sum (Wikipedia convention) for d^j_{mp,m}(beta)."""
    f = math.factorial
    pref = math.sqrt(f(j + mp) * f(j - mp) * f(j + m) * f(j - m))
    s = 0.0
    for k in range(max(0, m - mp), min(j + m, j - mp) + 1):
        den = f(j + m - k) * f(k) * f(mp - m + k) * f(j - mp - k)
        s += (-1) ** (mp - m + k) / den * math.cos(beta / 2) ** (2 * j - 2 * k + m - mp) * math.sin(beta / 2) ** (2 * k - m + mp)
    return pref * s


@pytest.mark.parametrize("s", [0, 2, -2])
def test_lambda_vs_wigner(s):
    orc = get_oracle("ld")
    lmax = 12
    for theta in (0.0, 0.3, 1.1, math.pi / 2, 2.5, math.pi):
        for m in range(0, lmax + 1):
            lam = orc.lam(lmax, m, s, theta)
            for l in range(max(m, abs(s)), lmax + 1):
                ref = (-1) ** m * math.sqrt((2 * l + 1) / (4 * math.pi)) * wigner_d(l, -m, s, theta)
                assert abs(lam[l] - ref) < 2e-13, (s, theta, m, l, lam[l], ref)


def test_lambda_spin0_is_ylm():
    from scipy.special import sph_harm_y
    orc = get_oracle("ld")
    lmax = 40
    for theta in (0.2, 1.0, 2.9):
        for m in (0, 1, 7, 40):
            lam = orc.lam(lmax, m, 0, theta)
            ref = np.array([sph_harm_y(l, m, theta, 0.0).real if l >= m else 0.0 for l in range(lmax + 1)])
            assert np.max(np.abs(lam - ref)) < 1e-13


def test_lambda_extreme_range():
    """sin^m(theta) ~ 1e-14000 must neither underflow to garbage nor produce NaN; mirror parity must hold."""
    for kind in ("ld", "d"):
        orc = get_oracle(kind)
        lam = orc.lam(4000, 3500, 0, 1e-4)
        assert np.all(lam == 0.0)
        a = orc.lam(3000, 1500, 0, 0.7)
        b = orc.lam(3000, 1500, 0, math.pi - 0.7)
        l = np.arange(3001)
        assert np.all(np.isfinite(a)) and np.max(np.abs(a)) > 0.1
        assert np.max(np.abs(b - (-1.0) ** (l + 1500) * a)) < 1e-9


@pytest.mark.parametrize("spin", [0, 2])
def test_adjoint(spin):
    orc = get_oracle("ld")
    shape, wcs = fullsky_geometry(6.0 * degree)  # 60 x 31
    b = sht_band(shape, wcs)
    theta, w = cc_geometry(b.nrings_total, b.nphi)
    lmax = 30
    nc = 1 if spin == 0 else 2
    rng = np.random.default_rng(3)
    alm = np.stack([synth_alm(lmax, lmax, 10 + c, spin2=spin == 2) for c in range(nc)])
    mp = rng.standard_normal((nc, b.nrings, b.nphi))
    ya = orc.alm2map(alm, theta, b.phi0, b.nphi, lmax, spin=spin)
    ytm = orc.map2alm(mp, theta, w, b.phi0, lmax, spin=spin)
    lhs = np.sum(w[None, :, None] * ya * mp)
    # real-field inner product over m >= 0 storage: m = 0 once, m > 0 twice
    fac = np.full(nalm(lmax), 2.0)
    fac[:lmax + 1] = 1.0
    rhs = np.sum(fac * (alm.conj() * ytm).real)
    assert abs(lhs - rhs) < 1e-11 * abs(lhs)


@pytest.mark.parametrize("spin", [0, 2])
def test_roundtrip_exact_when_bandlimited(spin):
    orc = get_oracle("ld")
    shape, wcs = fullsky_geometry(5.0 * degree)  # 72 x 37
    b = sht_band(shape, wcs)
    theta, w = cc_geometry(b.nrings_total, b.nphi)
    lmax = 18  # 2 lmax <= nrings - 1
    nc = 1 if spin == 0 else 2
    alm = np.stack([synth_alm(lmax, lmax, 20 + c, spin2=spin == 2) for c in range(nc)])
    mp = orc.alm2map(alm, theta, b.phi0, b.nphi, lmax, spin=spin)
    back = orc.map2alm(mp, theta, w, b.phi0, lmax, spin=spin)
    assert rel_rms(back, alm) < 1e-13


def test_double_build_matches_long_double():
    shape, wcs = fullsky_geometry(1.0 * degree)  # 360 x 181  (BASELINE config C1)
    b = sht_band(shape, wcs)
    theta, w = cc_geometry(b.nrings_total, b.nphi)
    lmax = 180
    alm = synth_alm(lmax, lmax, 1000)[None]
    m_ld = get_oracle("ld").alm2map(alm, theta, b.phi0, b.nphi, lmax)
    m_d = get_oracle("d").alm2map(alm, theta, b.phi0, b.nphi, lmax)
    assert rel_rms(m_d, m_ld) < 1e-13
    a_ld = get_oracle("ld").map2alm(m_ld, theta, w, b.phi0, lmax)
    a_d = get_oracle("d").map2alm(m_ld, theta, w, b.phi0, lmax)
    assert rel_rms(a_d, a_ld) < 1e-13


def test_sampling_options():
    orc = get_oracle("ld")
    shape, wcs = fullsky_geometry(6.0 * degree)
    b = sht_band(shape, wcs)
    theta, w = cc_geometry(b.nrings_total, b.nphi)
    lmax = 30
    alm = synth_alm(lmax, lmax, 5)[None]
    full = orc.alm2map(alm, theta, b.phi0, b.nphi, lmax)
    part = orc.alm2map(alm, theta, b.phi0, b.nphi, lmax, ring_stride=4, ring_offset=1)
    assert np.array_equal(part[:, 1::4], full[:, 1::4]) and not part[:, 0::4].any()
    fa = orc.map2alm(full, theta, w, b.phi0, lmax)
    pa = orc.map2alm(full, theta, w, b.phi0, lmax, m_stride=8, m_offset=3)  # direct-DFT branch
    for m in range(lmax + 1):
        i0, i1 = alm_index(lmax, m, m), alm_index(lmax, lmax, m) + 1
        if m % 8 == 3:
            assert rel_rms(pa[0, i0:i1], fa[0, i0:i1]) < 1e-13
        else:
            assert not pa[0, i0:i1].any()


# ---- the libsharp2-style CPU implementation (bench.py's timed CPU baseline) against the long-double checker -----------
@pytest.mark.parametrize("spin", [0, 2])
def test_cpu_baseline_implementation_matches_checker(spin):
    """oracle/sht_cpu.c (ring-pair folding, scaled seek, pruning, half-length real FFT) is an independent algorithm: it
    must agree with the naive long-double checker in both directions, on a full-sky grid and on a partial ring band
    (unpaired rings), and its m sampling must select exactly the requested columns."""
    from oracle import get_cpu_sht
    cpu, orc = get_cpu_sht(), get_oracle("ld")
    nc = 1 if spin == 0 else 2
    rng = np.random.default_rng(11 + spin)
    for nphi, nrt, first, nr, lmax in ((72, 37, 0, 37, 36), (120, 61, 7, 40, 75), (360, 181, 0, 181, 180)):
        theta, w = cc_geometry(nrt, nphi, first, nr)
        alms = np.stack([synth_alm(lmax, lmax, 50 + c, spin2=spin == 2) for c in range(nc)])
        got = cpu.alm2map(alms, theta, 0.3, nphi, lmax, spin=spin)
        ref = orc.alm2map(alms, theta, 0.3, nphi, lmax, spin=spin)
        assert rel_rms(got, ref) < 1e-12
        x = rng.standard_normal((nc, nr, nphi))
        got = cpu.map2alm(x, theta, w, 0.3, lmax, spin=spin)
        ref = orc.map2alm(x, theta, w, 0.3, lmax, spin=spin)
        assert rel_rms(got, ref) < 1e-12
        part = cpu.map2alm(x, theta, w, 0.3, lmax, spin=spin, m_stride=5, m_offset=2)
        sel = np.zeros(nalm(lmax), dtype=bool)
        for m in range(2, lmax + 1, 5):
            sel[alm_index(lmax, m, m):alm_index(lmax, lmax, m) + 1] = True
        assert np.array_equal(part[:, sel], got[:, sel]) and not np.any(part[:, ~sel])
