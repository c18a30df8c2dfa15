"""FITS map I/O either side of the transform path: host-side mirror of read_map / write_map / resolve_polcconv!
(src/enmap.jl:178-237 of the reference, which goes through FITSIO.jl/CFITSIO and WCS.jl/WCSLIB -- neither is in this image, so
the few pieces of the FITS standard the path needs are restated here with numpy: 2880-byte header blocks of 80-character cards,
big-endian image data, the CAR keywords CTYPE/CRPIX/CDELT/CRVAL).

Stokes-U sign: the transforms compute in the COSMO / HEALPix convention.  The reference negates U of a file tagged
POLCCONV = 'IAU' on the host at read time (resolve_polcconv!, src/enmap.jl:178-196).  read_map does the same by default;
with defer_polcconv=True the values stay as stored, the Enmap is tagged `polcconv = "IAU"` and map2alm applies the sign inside
the ring-FFT kernels (pixsht_plan_set_polconv) -- no extra pass over the map.
"""
import numpy as np

from .enmap import Enmap
from .geometry import CarClenshawCurtis, CarFejer1, create_car_wcs

_BLOCK = 2880
_BITPIX = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}
_BITPIX_OF = {"u1": 8, "i2": 16, "i4": 32, "i8": 64, "f4": -32, "f8": -64}


def _parse_value(raw):
    s = raw.strip()
    if s.startswith("'"):
        end = 1
        out = []
        while end < len(s):                      # '' inside a string is an escaped quote
            if s[end] == "'":
                if end + 1 < len(s) and s[end + 1] == "'":
                    out.append("'"); end += 2; continue
                break
            out.append(s[end]); end += 1
        return "".join(out).rstrip()
    s = s.split("/")[0].strip()
    if s in ("T", "F"):
        return s == "T"
    try:
        return int(s)
    except ValueError:
        try:
            return float(s.replace("D", "E"))
        except ValueError:
            return s


def _read_header(f):
    """-> (dict of keyword -> value, list of cards).  Leaves the file position at the first data block."""
    hdr, cards = {}, []
    while True:
        block = f.read(_BLOCK)
        if len(block) < _BLOCK:
            raise ValueError("truncated FITS header")
        for i in range(0, _BLOCK, 80):
            card = block[i:i + 80].decode("ascii", "replace")
            key = card[:8].strip()
            if key == "END":
                return hdr, cards
            cards.append(card)
            if card[8:10] == "= " and key:
                hdr[key] = _parse_value(card[10:])


def _data_bytes(hdr):
    naxis = hdr.get("NAXIS", 0)
    if naxis == 0:
        return 0
    n = 1
    for i in range(1, naxis + 1):
        n *= hdr["NAXIS%d" % i]
    n = abs(hdr["BITPIX"]) // 8 * hdr.get("GCOUNT", 1) * (hdr.get("PCOUNT", 0) + n)
    return (n + _BLOCK - 1) // _BLOCK * _BLOCK


def _wcs_from_header(hdr, W=CarClenshawCurtis):
    if hdr.get("CTYPE1") != "RA---CAR" or hdr.get("CTYPE2") != "DEC--CAR":
        raise AssertionError("read_map needs a CAR map (CTYPE1 = RA---CAR, CTYPE2 = DEC--CAR)")
    for k in (1, 2):
        if str(hdr.get("CUNIT%d" % k, "deg")).strip() not in ("deg", ""):
            raise ValueError("CUNIT%d = %r: only degrees are supported" % (k, hdr.get("CUNIT%d" % k)))
    cdelt = [hdr.get("CDELT%d" % k, hdr.get("CD%d_%d" % (k, k), 1.0)) for k in (1, 2)]
    return create_car_wcs(W, cdelt, (hdr.get("CRPIX1", 0.0), hdr.get("CRPIX2", 0.0)), (hdr.get("CRVAL1", 0.0), hdr.get("CRVAL2", 0.0)))


def resolve_polcconv(data, hdr, sel=(), verbose=True):
    """IAU -> COSMO: negate the U plane (third entry) of every axis whose CTYPE is STOKES, in place (src/enmap.jl:178-196).
    `sel` is the selection the data were read with (0-based slices / ints per axis), so that the plane is found after slicing."""
    naxis = hdr["NAXIS"]
    for i in range(1, naxis + 1):
        if hdr.get("CTYPE%d" % i, "") != "STOKES":
            continue
        n = hdr["NAXIS%d" % i]
        if n < 3:
            continue
        signs = np.ones(n)
        signs[2] = -1
        ax = i - 1
        if len(sel) == naxis:
            s = sel[ax]
            if isinstance(s, (int, np.integer)):
                if signs[s] < 0:
                    data *= -1        # the selection dropped the axis and kept U
                continue
            signs = signs[s]
        if verbose:
            print("convert to IAU: flip U in axis %d" % i)
        shape = [1] * data.ndim
        # axes dropped by integer selections shift the position of this one
        pos = ax - sum(1 for k in range(ax) if len(sel) == naxis and isinstance(sel[k], (int, np.integer)))
        shape[pos] = signs.size
        data *= signs.reshape(shape)
    return data


def read_map(path, hdu=1, sel=(), wcs=None, verbose=True, trim=True, defer_polcconv=False):
    """read_map(path; hdu=1, sel=(), wcs=nothing, verbose=true, trim=true)  (src/enmap.jl:199-228).

    hdu is 1-based as in the reference (1 = primary).  sel: tuple of 0-based Python slices / ints, one per FITS axis (the
    reference's 1-based ranges); as in the reference the WCS is taken from the header and is NOT adjusted for `sel`.
    The data come back as a Fortran-ordered array of shape (NAXIS1, NAXIS2[, NAXIS3]) -- the reference's column-major layout.
    trim=False returns the same CAR container (the reference returns the raw two-axis WCSTransform there)."""
    with open(path, "rb") as f:
        hdr = None
        for _ in range(hdu):
            hdr, _cards = _read_header(f)
            nbytes = _data_bytes(hdr)
            start = f.tell()
            f.seek(start + nbytes)
        f.seek(start)
        naxis = hdr.get("NAXIS", 0)
        if naxis == 0:
            raise ValueError("HDU %d holds no image" % hdu)
        shape = tuple(hdr["NAXIS%d" % i] for i in range(1, naxis + 1))
        dt = np.dtype(_BITPIX[hdr["BITPIX"]])
        raw = np.frombuffer(f.read(int(np.prod(shape)) * dt.itemsize), dtype=dt)
    data = raw.reshape(shape, order="F")
    if sel:
        if len(sel) != naxis:
            raise ValueError("sel needs one entry per FITS axis (%d)" % naxis)
        data = data[tuple(sel)]
    native = np.dtype(dt.str.replace(">", "="))
    if "BSCALE" in hdr or "BZERO" in hdr:
        data = data.astype(np.float64) * hdr.get("BSCALE", 1.0) + hdr.get("BZERO", 0.0)
    data = np.array(data, dtype=native if data.dtype == dt else data.dtype, order="F")
    polcconv = "COSMO"
    if wcs is None:
        if "STOKES" in hdr.values():
            if verbose and "POLCCONV" not in hdr:
                print("STOKES found but POLCCONV not found, assuming IAU")   # the reference's message; its default is in fact COSMO
            polcconv = hdr.get("POLCCONV", "COSMO")
            if polcconv == "IAU" and not defer_polcconv:
                resolve_polcconv(data, hdr, sel, verbose=verbose)
                polcconv = "COSMO"
        wcs = _wcs_from_header(hdr, CarFejer1 if hdr.get("PIXSHTRS") == "FEJER1" else CarClenshawCurtis)
    m = Enmap(data, wcs)
    m.polcconv = polcconv
    return m


def _card(key, value, comment=""):
    if isinstance(value, bool):
        v = "%20s" % ("T" if value else "F")
    elif isinstance(value, (int, np.integer)):
        v = "%20d" % value
    elif isinstance(value, (float, np.floating)):
        v = "%20s" % repr(float(value)).upper().replace("INF", "inf")
    else:
        v = "%-20s" % ("'%-8s'" % str(value).replace("'", "''"))
    c = "%-8s= %s" % (key, v)
    if comment:
        c += " / " + comment
    return ("%-80s" % c)[:80]


def write_map(fname, emap):
    """write_map(fname, emap)  (src/enmap.jl:230-237 writes the data, then the WCS header cards).  Primary HDU only."""
    data = np.asarray(emap.data)
    if data.dtype.kind == "f" and data.dtype.itemsize not in (4, 8):
        data = data.astype(np.float64)
    code = data.dtype.str[1:]
    if code not in _BITPIX_OF:
        raise TypeError("cannot write dtype %s to FITS" % data.dtype)
    wcs = emap.wcs
    cards = [_card("SIMPLE", True, "file does conform to FITS standard"), _card("BITPIX", _BITPIX_OF[code], "number of bits per data pixel"),
             _card("NAXIS", data.ndim, "number of data axes")]
    for i, n in enumerate(data.shape):
        cards.append(_card("NAXIS%d" % (i + 1), int(n), "length of data axis %d" % (i + 1)))
    cards.append(_card("EXTEND", True, "FITS dataset may contain extensions"))
    cards.append(_card("WCSAXES", 2, "Number of coordinate axes"))
    for k in (1, 2):
        cards.append(_card("CRPIX%d" % k, float(wcs.crpix[k - 1]), "Pixel coordinate of reference point"))
    for k in (1, 2):
        cards.append(_card("CDELT%d" % k, float(wcs.cdelt[k - 1]), "[deg] Coordinate increment at reference point"))
    for k in (1, 2):
        cards.append(_card("CUNIT%d" % k, "deg", "Units of coordinate increment and value"))
    cards.append(_card("CTYPE1", "RA---CAR", "Right ascension, plate caree projection"))
    cards.append(_card("CTYPE2", "DEC--CAR", "Declination, plate caree projection"))
    for k in (1, 2):
        cards.append(_card("CRVAL%d" % k, float(wcs.crval[k - 1]), "[deg] Coordinate value at reference point"))
    if isinstance(wcs, CarFejer1):
        cards.append(_card("PIXSHTRS", "FEJER1", "ring scheme (pixsht extension)"))
    pol = getattr(emap, "polcconv", None)
    if data.ndim == 3 and pol in ("IAU", "COSMO") and data.shape[2] == 3:
        cards += [_card("CTYPE3", "STOKES"), _card("CRPIX3", 1.0), _card("CRVAL3", 1.0), _card("CDELT3", 1.0), _card("POLCCONV", pol)]
    cards.append("%-80s" % "END")
    head = "".join(cards).encode("ascii")
    head += b" " * (-len(head) % _BLOCK)
    body = np.asfortranarray(data).astype(">" + code).tobytes(order="F")
    body += b"\0" * (-len(body) % _BLOCK)
    with open(fname, "wb") as f:
        f.write(head)
        f.write(body)
