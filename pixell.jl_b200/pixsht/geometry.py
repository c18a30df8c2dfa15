"""Host-side mirror of the Pixell.jl geometry pieces that the SHT path uses.

Mirrors (names, argument meaning, results), with file:line of the reference it follows:
  CarClenshawCurtis / CarFejer1      src/projections/car_proj.jl:7-24
  fullsky_geometry                   src/enmap_geom.jl:47-73
  geometry (bounding box)            src/enmap_geom.jl:77-108
  slice_geometry / sliced_wcs        src/enmap_ops.jl:154-167, src/projections/car_proj.jl:275-278
  pix2sky (scalar) / rewind          src/projections/car_proj.jl:141-152, src/enmap_ops.jl:10-13
  fullringsize / fullringnum / getlmax / first_last_rings_in_fullsky / get_flip_slices
                                     src/transforms.jl:3-30,85

Pixel coordinates passed to pix2sky are 1-based (FITS/WCS and Julia convention); array indices of Enmap are 0-based
Python indices.  Selections are given as Python slices and are translated to the reference's 1-based ranges.
"""
from dataclasses import dataclass, replace
import math

degree = math.pi / 180.0
arcminute = degree / 60.0
radian = 1.0


@dataclass(frozen=True)
class CarClenshawCurtis:
    """Fast CAR WCS with pixels on both poles (src/projections/car_proj.jl:7-12). cdelt/crpix/crval in degrees."""
    cdelt: tuple
    crpix: tuple
    crval: tuple
    unit: float = math.pi / 180.0
    naxis: int = 2


@dataclass(frozen=True)
class CarFejer1:
    """Fejer-1 CAR WCS type (src/projections/car_proj.jl:14-19).  As in the reference it has no SHT path."""
    cdelt: tuple
    crpix: tuple
    crval: tuple
    unit: float = math.pi / 180.0
    naxis: int = 2


def getunit(wcs):
    return wcs.unit


def getcdelt(wcs):
    return wcs.cdelt


def getcrpix(wcs):
    return wcs.crpix


def getcrval(wcs):
    return wcs.crval


def create_car_wcs(W, cdelt, crpix, crval):
    return W((float(cdelt[0]), float(cdelt[1])), (float(crpix[0]), float(crpix[1])),
             (float(crval[0]), float(crval[1])), math.pi / 180.0)


def _julia_round(x):
    """Julia's round(Int, x): ties to even (same as Python's round)."""
    return int(round(x))


def fullsky_geometry(res, W=CarClenshawCurtis, shape=None, dims=()):
    """Full-sky CAR geometry with pixels on the poles (src/enmap_geom.jl:47-73). `res` in radians (number or pair).
    W = CarFejer1 (extension: the reference has the type but no constructor) gives the Fejer-1 grid: ny =
    pi/res rings at colatitudes (k + 1/2) res, none on the poles."""
    if not isinstance(res, (tuple, list)):
        res = (res, res)
    resx, resy = float(res[0]), float(res[1])
    if W is CarFejer1:
        if shape is None:
            shape = (_julia_round(2 * math.pi / resx), _julia_round(math.pi / resy))
        nx, ny = shape
        if not abs(resx * nx - 2 * math.pi) < 1e-8 or not abs(resy * ny - math.pi) < 1e-8:
            raise AssertionError("Resolution does not evenly divide the sky; this is required for SHTs.")
        wcs = create_car_wcs(W, (-360.0 / nx, 180.0 / ny), (math.floor(nx / 2) + 0.5, (ny + 1) / 2), (resy * 90 / math.pi, 0.0))
        return (nx, ny) + tuple(dims), wcs
    if shape is None:
        shape = (_julia_round(2 * math.pi / resx), _julia_round(math.pi / resy + 1))
    nx, ny = shape
    if not abs(resx * nx - 2 * math.pi) < 1e-8:
        raise AssertionError("Horizontal resolution does not evenly divide the sky; this is required for SHTs.")
    if not abs(resy * (ny - 1) - math.pi) < 1e-8:
        raise AssertionError("Vertical resolution does not evenly divide the sky; this is required for SHTs.")
    wcs = create_car_wcs(W, (-360.0 / nx, 180.0 / (ny - 1)), (math.floor(nx / 2) + 0.5, (ny + 1) / 2),
                         (resy * 90 / math.pi, 0.0))
    return (nx, ny) + tuple(dims), wcs


def geometry(W, bbox_coords, res):
    """Bounding-box CAR geometry (src/enmap_geom.jl:77-108).  bbox_coords[row][col]: rows (RA, DEC), cols (from, to),
    in radians; res in radians."""
    if not isinstance(res, (tuple, list)):
        res = (res, res)
    resx, resy = float(res[0]), float(res[1])
    if not abs(2 * math.pi / resx - round(2 * math.pi / resx)) < 1e-8:
        raise AssertionError("Horizontal resolution does not evenly divide the sky; this is required for SHTs.")
    if not abs(2 * math.pi / resy - round(2 * math.pi / resy)) < 1e-8:
        raise AssertionError("Vertical resolution does not evenly divide the sky; this is required for SHTs.")
    pos1 = (float(bbox_coords[0][0]), float(bbox_coords[1][0]))
    pos2 = (float(bbox_coords[0][1]), float(bbox_coords[1][1]))
    dra, ddec = abs(pos1[0] - pos2[0]), abs(pos1[1] - pos2[1])
    shape = (_julia_round(dra / resx), _julia_round(ddec / resy))
    mid = ((pos1[0] + pos2[0]) / 2, (pos1[1] + pos2[1]) / 2)
    crval = (mid[0], 0.0)
    sign = lambda v: (v > 0) - (v < 0)
    cdelt = (abs(resx) * sign(pos2[0] - pos1[0]), abs(resy) * sign(pos2[1] - pos1[1]))
    crpix = (1 - (pos1[0] - crval[0]) / cdelt[0], 1 - (pos1[1] - crval[1]) / cdelt[1])
    wcs = create_car_wcs(W, (math.degrees(cdelt[0]), math.degrees(cdelt[1])), crpix,
                         (math.degrees(crval[0]), math.degrees(crval[1])))
    return shape, wcs


def rewind(angle, period=2 * math.pi, ref_angle=0.0):
    """src/enmap_ops.jl:10-13 (Julia's mod is floored, like Python's %)."""
    half = period / 2
    return ref_angle + ((angle - ref_angle + half) % period) - half


def pix2sky(shape, wcs, ra_pixel, dec_pixel, safe=True):
    """Scalar CAR pix2sky with 1-based pixel coordinates (src/projections/car_proj.jl:141-152). Returns radians."""
    u = getunit(wcs)
    a0, d0 = getcrval(wcs)[0] * u, getcrval(wcs)[1] * u
    da, dd = getcdelt(wcs)[0] * u, getcdelt(wcs)[1] * u
    ia0, id0 = getcrpix(wcs)
    a = a0 + (ra_pixel - ia0) * da
    d = d0 + (dec_pixel - id0) * dd
    if safe:
        return rewind(a), rewind(d)
    return a, d


def _to_julia_range(sel, n):
    """Python slice / int (0-based) on an axis of length n -> (first, step, last) of the equivalent 1-based Julia range."""
    if isinstance(sel, int):
        i = sel + n if sel < 0 else sel
        return i + 1, 1, i + 1
    start, stop, step = sel.indices(n)
    cnt = len(range(start, stop, step))
    if cnt == 0:
        raise ValueError("empty selection")
    return start + 1, step, start + (cnt - 1) * step + 1


def slice_geometry(shape_all, wcs, sel_x=slice(None), sel_y=slice(None)):
    """src/enmap_ops.jl:154-167; sel_x / sel_y are Python slices (negative steps allowed)."""
    other = tuple(shape_all[2:])
    fx, sx, lx = _to_julia_range(sel_x, shape_all[0])
    fy, sy, ly = _to_julia_range(sel_y, shape_all[1])
    starts = (fx - 1 if sx > 0 else fx, fy - 1 if sy > 0 else fy)
    steps = (sx, sy)
    sel_sizes = (lx - fx + sx, ly - fy + sy)
    crpix = getcrpix(wcs)
    cdelt = getcdelt(wcs)
    crpix2 = tuple((crpix[k] - (starts[k] + 0.5)) / steps[k] + 0.5 for k in range(2))
    cdelt2 = tuple(cdelt[k] * steps[k] for k in range(2))
    shape = tuple(sel_sizes[k] // steps[k] for k in range(2))
    return shape + other, replace(wcs, cdelt=cdelt2, crpix=crpix2)


def fullringsize(wcs):
    """Number of pixels in a full CAR ring of this WCS (src/transforms.jl:3-4)."""
    return _julia_round(abs(2 * math.pi / (getunit(wcs) * getcdelt(wcs)[0])))


def fullringnum(wcs):
    """Number of rings of the full-sky version of this WCS (src/transforms.jl:7-8); Fejer-1 grids have no pole rings."""
    n = _julia_round(abs(math.pi / (getunit(wcs) * getcdelt(wcs)[1])))
    return n if isinstance(wcs, CarFejer1) else 1 + n


def getlmax(wcs):
    """src/transforms.jl:85"""
    return fullringsize(wcs) // 2


def first_last_rings_in_fullsky(shape, wcs):
    """1-based (first, last) full-sky ring indices of the map's first/last row (src/transforms.jl:11-22)."""
    dth = abs(getcdelt(wcs)[1] * getunit(wcs))
    d1 = pix2sky(shape, wcs, 1, 1)[1]
    d2 = pix2sky(shape, wcs, 1, shape[1])[1]
    half = 0.5 if isinstance(wcs, CarFejer1) else 0.0      # Fejer-1 rings sit at (k + 1/2) dtheta
    i1 = _julia_round((math.pi / 2 - d1) / dth - half) + 1
    i2 = _julia_round((math.pi / 2 - d2) / dth - half) + 1
    return i1, i2


def get_flip_slices(shape, wcs):
    """libsharp wants ascending colatitude and ascending RA (src/transforms.jl:25-30). Returns Python slices."""
    da, dd = getcdelt(wcs)[0] * getunit(wcs), getcdelt(wcs)[1] * getunit(wcs)
    fx = slice(None) if da >= 0 else slice(None, None, -1)
    fy = slice(None) if dd <= 0 else slice(None, None, -1)
    return fx, fy


@dataclass(frozen=True)
class ShtBand:
    """Everything the native engine needs to know about how a map sits on the full-sky ring grid.
    Host restatement of create_sht_band + make_cc_geom_info (src/transforms.jl:33-82), without copying the map."""
    nphi: int          # full ring size
    nrings_total: int  # rings of the full-sky grid (poles included)
    ring_first: int    # 0-based full-sky index of the band's first ring (ascending theta)
    nrings: int        # rings in the band (= ny)
    nx: int            # columns actually present in the map (<= nphi); band columns nx.. are zero padding
    flipx: bool        # band column i  <-> map column nx-1-i
    flipy: bool        # band ring r    <-> map row ny-1-r
    phi0: float        # RA of band column 0, radians, in [-pi, pi]
    ring_scheme: int = 0   # 0: Clenshaw-Curtis rings (poles included), 1: Fejer-1 rings


def sht_band(shape, wcs):
    fx, fy = get_flip_slices(shape, wcs)
    _, w2 = slice_geometry(tuple(shape), wcs, fx, fy)
    i1, i2 = first_last_rings_in_fullsky(shape, w2)
    if not i1 <= i2:
        raise AssertionError("vertical angle must be increasing")
    phi0 = pix2sky(shape, w2, 1, 2)[0]
    nphi = fullringsize(w2)
    if shape[0] > nphi:
        raise ValueError("map is wider than a full ring")
    if i2 - i1 + 1 != shape[1]:
        raise ValueError("map rows do not align with the full-sky ring grid")
    return ShtBand(nphi=nphi, nrings_total=fullringnum(w2), ring_first=i1 - 1, nrings=shape[1], nx=shape[0],
                   flipx=fx.step == -1, flipy=fy.step == -1, phi0=phi0, ring_scheme=1 if isinstance(wcs, CarFejer1) else 0)
