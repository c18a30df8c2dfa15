"""Pins the CPU oracle -- the naive checker in both precisions and the libsharp2-style CPU implementation that bench.py times
-- against every SHT golden vector the reference ships: test/test_transforms.jl:11-77 with
test/data/simple_*.txt.  CPU only."""
import numpy as np
import pytest

import pixsht
from pixsht import Enmap, CarClenshawCurtis, fullsky_geometry, geometry, degree
from helpers import golden_alm, gen_spin0, gen_spin2, oracle_map2alm, rel_rms

TOL = 1e-12  # the reference itself only asks for isapprox (rtol sqrt(eps)); the oracle does much better


def relmax(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


@pytest.mark.parametrize("kind", ["ld", "d", "cpu"])
def test_spin0_fullsky_lmax18(kind):
    shape, wcs = fullsky_geometry(10.0 * degree)
    assert shape == (36, 19)
    m = Enmap(gen_spin0(shape), wcs)
    alm = oracle_map2alm(m, 18, kind=kind)[0]
    ref = golden_alm("simple_analytic_sht")
    assert alm.shape == ref.shape == (190,)
    assert relmax(alm, ref) < TOL


def test_spin0_default_lmax_count():
    shape, wcs = fullsky_geometry(10.0 * degree)
    assert pixsht.getlmax(wcs) == 18  # -> 190 coefficients (test/test_transforms.jl:21)


@pytest.mark.parametrize("kind", ["ld", "d", "cpu"])
def test_spin0_sliced(kind):
    shape, wcs = fullsky_geometry(10.0 * degree)
    m = Enmap(gen_spin0(shape), wcs)
    sub = m[5:-2, 4:-3]  # Julia m[6:end-2, 5:end-3]
    assert sub.shape == (29, 12)
    alm = oracle_map2alm(sub, 18, kind=kind)[0]
    assert relmax(alm, golden_alm("simple_analytic_sht_sliced")) < TOL


@pytest.mark.parametrize("kind", ["ld", "cpu"])
def test_spin0_box_lmax100(kind):
    box = [[10 * degree, -10 * degree], [-5 * degree, 5 * degree]]
    shape, wcs = geometry(CarClenshawCurtis, box, 1.0 * degree)
    assert shape == (20, 10)
    m = Enmap(gen_spin0(shape, 2.5), wcs)
    alm = oracle_map2alm(m, 100, kind=kind)[0]
    assert relmax(alm, golden_alm("simple_box_analytic_sht")) < TOL


@pytest.mark.parametrize("kind", ["ld", "cpu"])
def test_spin0_aliased_lmax108(kind):
    shape, wcs = fullsky_geometry(10.0 * degree)
    m = Enmap(gen_spin0(shape), wcs)
    alm = oracle_map2alm(m, 108, kind=kind)[0]
    assert relmax(alm, golden_alm("simple_analytic_sht_fullalm")) < TOL


@pytest.mark.parametrize("kind", ["ld", "d", "cpu"])
def test_spin2_lmax108(kind):
    shape, wcs = fullsky_geometry(10.0 * degree, dims=(2,))
    m = Enmap(gen_spin2(shape), wcs)
    eb = oracle_map2alm(m, 108, spin=2, kind=kind)
    assert relmax(eb[0], golden_alm("simple_pol_analytic_sht", (0, 1))) < TOL
    assert relmax(eb[1], golden_alm("simple_pol_analytic_sht", (2, 3))) < TOL
