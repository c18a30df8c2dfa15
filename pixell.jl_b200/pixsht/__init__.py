"""pixsht: B200-native spherical harmonic transforms behind Pixell.jl's map2alm / alm2map API (host-side mirror).

The arithmetic lives in the CUDA shared library pixell.jl_b200/lib/libpixsht.so (C ABI: include/pixsht.h); importing the
geometry/container mirrors needs no GPU, calling a transform does and fails loudly without the library or a device.
"""
from .geometry import (CarClenshawCurtis, CarFejer1, fullsky_geometry, geometry, slice_geometry, pix2sky, rewind,  # noqa
                       fullringsize, fullringnum, getlmax, first_last_rings_in_fullsky, get_flip_slices, sht_band,
                       ShtBand, degree, arcminute, radian, getcdelt, getcrpix, getcrval, getunit)
from .enmap import Enmap, Alm, alm2cl  # noqa: F401
from .fits_io import read_map, write_map, resolve_polcconv  # noqa: F401


def __getattr__(name):
    # transforms import the ctypes binding lazily so that geometry-only users never touch the native library
    if name in ("map2alm", "alm2map", "Plan", "get_lib", "PixshtError", "PixshtLib", "pixareamap", "pixareamap_", "ring_pixarea"):
        from . import transforms
        return getattr(transforms, name)
    raise AttributeError(name)
