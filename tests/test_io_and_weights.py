"""FITS map I/O, the IAU/COSMO Stokes-U convention and pixel-area weights: the data formats and callers either side of the transform
path.  CPU tests: the reference's own fixtures (test/test_io.jl:4-14 on test.fits, kept byte for byte as
tests/golden/ref_test.fits; test/test_geometry.jl:287-316 on the python-pixell pixel areas, tests/golden/ref_pixareas.npz) and the
kernel-side U sign on the host-emulation build.  The GPU run of the U sign is in test_gpu_parity.py."""
import os

import numpy as np
import pytest

import pixsht
from pixsht import Enmap, CarClenshawCurtis, fullsky_geometry, geometry, degree, arcminute, read_map, write_map
from pixsht.transforms import PixshtLib, map2alm, alm2map, pixareamap, pixareamap_, ring_pixarea
from helpers import gen_spin2

HERE = os.path.dirname(os.path.abspath(__file__))
FITS = os.path.join(HERE, "golden", "ref_test.fits")
AREAS = np.load(os.path.join(HERE, "golden", "ref_pixareas.npz"))


def _lib():
    """pixsht_ring_pixarea is host arithmetic: any build of the library serves (the nvcc build loads without a GPU)."""
    from pixsht._lib import DEFAULT_LIB
    emu = os.path.join(HERE, "emu", "_build", "libpixsht_emu.so")
    return PixshtLib(DEFAULT_LIB if os.path.exists(DEFAULT_LIB) else emu)


@pytest.mark.parametrize("trim", [True, False])
def test_read_map_reference_fixture(trim):
    # test/test_io.jl:4-14
    imap = read_map(FITS, trim=trim)
    assert imap.shape == (100, 100, 3)
    assert imap.wcs.naxis == 2
    assert list(imap.wcs.cdelt) == [-1, 1]
    assert list(imap.wcs.crval) == [0.5, 0.0]
    assert abs(float(imap.data.sum()) - 14967.2985) < 1e-4
    assert imap.data.flags.f_contiguous and imap.data.dtype == np.float64
    sub = read_map(FITS, sel=(slice(10, 20), slice(20, 40), slice(0, 2)), trim=trim)     # the reference's 11:20, 21:40, 1:2
    assert sub.shape == (10, 20, 2)
    assert np.array_equal(sub.data, imap.data[10:20, 20:40, 0:2])
    assert sub.wcs == imap.wcs          # as in the reference, `sel` does not touch the WCS


def test_write_read_round_trip(tmp_path):
    shape, wcs = geometry(CarClenshawCurtis, np.array([[10.0, -10.0], [-5.0, 5.0]]) * degree, 0.5 * degree)
    rng = np.random.default_rng(3)
    for dt in (np.float64, np.float32):
        m = Enmap(np.asfortranarray(rng.standard_normal(tuple(shape) + (3,)).astype(dt)), wcs)
        f = str(tmp_path / ("m_%s.fits" % np.dtype(dt).name))
        write_map(f, m)
        assert os.path.getsize(f) % 2880 == 0
        r = read_map(f, verbose=False)
        assert r.data.dtype == dt and np.array_equal(r.data, m.data)
        assert r.wcs == m.wcs
    # the reference fixture survives a rewrite
    a = read_map(FITS)
    f = str(tmp_path / "again.fits")
    write_map(f, a)
    b = read_map(f)
    assert np.array_equal(a.data, b.data) and a.wcs == b.wcs


def test_polcconv_iau_is_flipped_on_read_or_deferred(tmp_path):
    # src/enmap.jl:178-196, 209-215: a STOKES axis with POLCCONV = IAU -> U negated at read time; COSMO / no keyword -> untouched
    shape, wcs = fullsky_geometry(10.0 * degree)
    rng = np.random.default_rng(5)
    m = Enmap(np.asfortranarray(rng.standard_normal(tuple(shape) + (3,))), wcs)
    m.polcconv = "IAU"
    f = str(tmp_path / "iau.fits")
    write_map(f, m)
    r = read_map(f, verbose=False)
    assert r.polcconv == "COSMO"
    assert np.array_equal(r.data[:, :, :2], m.data[:, :, :2]) and np.array_equal(r.data[:, :, 2], -m.data[:, :, 2])
    d = read_map(f, verbose=False, defer_polcconv=True)
    assert d.polcconv == "IAU" and np.array_equal(d.data, m.data)
    # selections: U kept by a range, U picked by an integer, U dropped
    s = read_map(f, verbose=False, sel=(slice(None), slice(None), slice(1, 3)))
    assert np.array_equal(s.data[:, :, 0], m.data[:, :, 1]) and np.array_equal(s.data[:, :, 1], -m.data[:, :, 2])
    u = read_map(f, verbose=False, sel=(slice(None), slice(None), 2))
    assert np.array_equal(u.data, -m.data[:, :, 2])
    q = read_map(f, verbose=False, sel=(slice(None), slice(None), 1))
    assert np.array_equal(q.data, m.data[:, :, 1])
    m.polcconv = "COSMO"
    write_map(f, m)
    assert np.array_equal(read_map(f, verbose=False).data, m.data)


def test_pixareamap_reference_vectors():
    # test/test_geometry.jl:287-316
    lib = _lib()
    box = np.array([[10.0, -10.0], [-5.0, 5.0]]) * degree
    boxgeom = geometry(CarClenshawCurtis, box, 5.0 * arcminute)
    fullgeom = fullsky_geometry(np.pi / 180)
    for (shape, wcs), ref in ((fullgeom, AREAS["fullsky"]), (boxgeom, AREAS["box"])):
        pm = pixareamap(shape, wcs, lib=lib)
        assert pm.shape == tuple(shape[:2])
        assert np.sum(np.abs(pm.data[0, :] - ref)) < 100 * np.finfo(float).eps
        assert np.all(pm.data == pm.data[:1, :])
        m = Enmap.zeros(shape, wcs)
        pm2 = pixareamap(m, lib=lib)
        assert np.sum(np.abs(pm2.data[0, :] - ref)) < 100 * np.finfo(float).eps
        m.data[...] = 0.0
        pm3 = pixareamap_(m, lib=lib)
        assert pm3 is m and np.sum(np.abs(m.data[0, :] - ref)) < 100 * np.finfo(float).eps
    # the areas of a full-sky map tile the sphere
    shape, wcs = fullgeom
    assert abs(np.sum(ring_pixarea(shape, wcs, lib=lib)) * shape[0] - 4 * np.pi) < 1e-12


def test_pixarea_bad_geometry_is_an_error():
    import ctypes
    from pixsht._lib import Geom, ERR_ARG
    lib = _lib()
    g = Geom(36, 19, 10, 12, 36, 0, 0, 0, 0.0)     # band runs past the last ring
    out = np.zeros(12)
    assert lib.lib.pixsht_ring_pixarea(ctypes.byref(g), out.ctypes.data_as(ctypes.POINTER(ctypes.c_double))) == ERR_ARG


def test_u_sign_in_the_kernels_emulation_build():
    """pixsht_plan_set_polconv(IAU): map2alm of (Q, -U) tagged IAU == map2alm of (Q, U); alm2map(..., polcconv="IAU") returns -U.
    Host-emulation build of the same fft.cuh (logic only; the device run is test_gpu_parity.py::test_polconv_iau)."""
    import subprocess
    so = os.path.join(HERE, "emu", "_build", "libpixsht_emu.so")
    if not os.path.exists(so):
        subprocess.check_call(["bash", os.path.join(HERE, "emu", "build_emu.sh")])
    emu = PixshtLib(so)
    shape, wcs = fullsky_geometry(10.0 * degree)
    d = gen_spin2(shape)
    rng = np.random.default_rng(11)
    t = np.asfortranarray(rng.standard_normal(shape))
    iqu = Enmap(np.asfortranarray(np.dstack([t, d[:, :, 0], d[:, :, 1]])), wcs)
    iau = Enmap(np.asfortranarray(np.dstack([t, d[:, :, 0], -d[:, :, 1]])), wcs)
    ref = map2alm(iqu, lmax=18, lib=emu)
    got = map2alm(iau, lmax=18, lib=emu, polcconv="IAU")
    iau.polcconv = "IAU"
    tag = map2alm(iau, lmax=18, lib=emu)                      # the tag left by read_map(..., defer_polcconv=True)
    for a, b, c in zip(ref, got, tag):
        assert np.array_equal(a.alm, b.alm) and np.array_equal(a.alm, c.alm)
    qu = map2alm((Enmap(iau.data[:, :, 1], wcs), Enmap(iau.data[:, :, 2], wcs)), lmax=18, lib=emu, polcconv="IAU")
    assert np.array_equal(qu[0].alm, ref[1].alm) and np.array_equal(qu[1].alm, ref[2].alm)
    cos = alm2map(ref, shape, wcs, lib=emu)
    out = alm2map(ref, shape, wcs, lib=emu, polcconv="IAU")
    assert out[2].polcconv == "IAU"
    assert np.array_equal(out[0].data, cos[0].data) and np.array_equal(out[1].data, cos[1].data) and np.array_equal(out[2].data, -cos[2].data)
    # the flag is per call: the cached plan goes back to COSMO
    again = map2alm(iqu, lmax=18, lib=emu)
    assert np.array_equal(again[2].alm, ref[2].alm)
