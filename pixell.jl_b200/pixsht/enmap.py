"""Enmap container and Alm container: host-side mirrors of src/enmap.jl:10-80 and of Healpix.Alm (imported by the
reference at src/Pixell.jl:17).  Only what the SHT path touches: data + wcs, size, WCS-aware slicing/views."""
import numpy as np

from .geometry import slice_geometry


class Enmap:
    """Array + WCS (src/enmap.jl:10-18).  `data` has shape (nx, ny[, ncomp]) with the RA index fastest in memory
    (Fortran order), exactly the reference's column-major Julia array, so raw buffers can be handed to the C ABI."""

    def __init__(self, data, wcs):
        data = np.asarray(data)
        if data.ndim not in (2, 3):
            raise ValueError("Enmap needs (nx, ny) or (nx, ny, ncomp) data")
        self.data = data
        self.wcs = wcs

    @classmethod
    def zeros(cls, shape, wcs, dtype=np.float64):
        return cls(np.zeros(tuple(shape), dtype=dtype, order="F"), wcs)

    @property
    def shape(self):
        return self.data.shape

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def ndim(self):
        return self.data.ndim

    def getwcs(self):
        return self.wcs

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.data, dtype=dtype)

    def _index(self, idx, copy):
        if not isinstance(idx, tuple):
            idx = (idx,)
        idx = idx + (slice(None),) * (self.data.ndim - len(idx))
        ix, iy = idx[0], idx[1]
        sub = self.data[idx]
        if isinstance(ix, (int, np.integer)) or isinstance(iy, (int, np.integer)):
            return sub.copy() if copy else sub  # a spatial axis was dropped: plain array, as in the reference
        _, wcs = slice_geometry(self.data.shape, self.wcs, ix, iy)
        if sub.ndim not in (2, 3):
            return sub.copy() if copy else sub
        return Enmap(np.asfortranarray(sub) if copy else sub, wcs)

    def __getitem__(self, idx):
        """getindex semantics of the reference: slicing copies and updates the WCS (src/enmap.jl:52-78)."""
        return self._index(idx, copy=True)

    def view(self, *idx):
        """view semantics: no copy, WCS updated (src/enmap.jl:40-43)."""
        return self._index(tuple(idx), copy=False)

    def __setitem__(self, idx, v):
        self.data[idx] = v

    def __repr__(self):
        return "Enmap(shape=%s,wcs=%s)" % (self.data.shape, self.wcs)


class Alm:
    """Healpix.Alm mirror: `alm` complex128 vector, triangular m-major, idx0(l,m) = m(2 lmax+1-m)/2 + l, m >= 0."""

    def __init__(self, lmax, mmax=None, alm=None):
        mmax = lmax if mmax is None else mmax
        if mmax > lmax or lmax < 0 or mmax < 0:
            raise ValueError("need 0 <= mmax <= lmax")
        n = (mmax + 1) * (lmax + 1) - mmax * (mmax + 1) // 2
        if alm is None:
            alm = np.zeros(n, dtype=np.complex128)
        alm = np.ascontiguousarray(alm)
        if alm.shape != (n,):
            raise ValueError("alm has %d entries, expected %d" % (alm.size, n))
        self.lmax, self.mmax, self.alm = int(lmax), int(mmax), alm

    def index(self, l, m):
        return m * (2 * self.lmax + 1 - m) // 2 + l

    def __len__(self):
        return self.alm.shape[0]


def alm2cl(a, b=None, lib=None):
    """Healpix.alm2cl mirror (used by the reference's tests, test/test_transforms.jl:104-107):
    C_l = (a_l0 b_l0* + 2 sum_{m>=1} Re(a_lm b_lm*)) / (2l+1).
    With `lib` (a PixshtLib) the sum runs on the GPU through pixsht_alm2cl; without, it is host bookkeeping in numpy."""
    b = a if b is None else b
    if (a.lmax, a.mmax) != (b.lmax, b.mmax):
        raise ValueError("alm geometries differ")
    lmax, mmax = a.lmax, a.mmax
    if lib is not None:
        import ctypes
        x = np.ascontiguousarray(a.alm, dtype=np.complex128)
        y = np.ascontiguousarray(b.alm, dtype=np.complex128)
        out = np.empty(lmax + 1)
        lib.check(lib.lib.pixsht_alm2cl(lmax, mmax, x.ctypes.data, y.ctypes.data, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 0, 0, 0))
        return out
    cl = np.zeros(lmax + 1)
    for m in range(mmax + 1):
        i0 = a.index(m, m)
        seg = (a.alm[i0:i0 + lmax - m + 1] * np.conj(b.alm[i0:i0 + lmax - m + 1])).real
        cl[m:] += seg if m == 0 else 2 * seg
    return cl / (2 * np.arange(lmax + 1) + 1)
