/*
 * pixsht_sharp_shim.h -- the subset of libsharp2's C API that Pixell.jl reaches (directly by ccall or through
 * Libsharp.jl 0.2), re-exported by libpixsht.so with identical symbol names and signatures, so that the UNMODIFIED
 * src/transforms.jl of the reference runs on the B200 engine by pointing `Libsharp.libsharp2` at libpixsht.so.
 *
 * Call sites replaced (file:line under /root/reference):
 *   sharp_make_geom_info            src/transforms.jl:55-60   (direct ccall)
 *   sharp_make_triangular_alm_info  src/transforms.jl:94,124,179,212,230,238  (Libsharp.make_triangular_alm_info)
 *   sharp_alm_count / sharp_map_size  src/transforms.jl:96-97,121,125,176,180,210,234
 *   sharp_execute                   src/transforms.jl:185-194 (direct ccall), :101-106,128-132,214-218,240-244 (sharp_execute!)
 *   sharp_destroy_geom_info / sharp_destroy_alm_info   Libsharp.jl finalizers
 *
 * Signatures and constants follow the upstream libsharp2 header (libsharp2/sharp.h, not in /root/reference).
 * Supported: iso-latitude rings of equal length (nph constant, stride 1, ofs[i] = i*nph, phi0 constant) -- exactly what
 * make_cc_geom_info builds -- triangular alm with stride 1, spin 0 and spin 2, SHARP_MAP2ALM / SHARP_ALM2MAP, with or
 * without SHARP_DP.  Anything else (SHARP_ADD, other spins, ragged rings) leaves the outputs untouched and records an
 * error retrievable with pixsht_last_error(); the shim never aborts the process (libsharp2 would).
 */
#ifndef PIXSHT_SHARP_SHIM_H
#define PIXSHT_SHARP_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct sharp_geom_info sharp_geom_info;
typedef struct sharp_alm_info sharp_alm_info;

typedef enum { SHARP_YtW = 0, SHARP_MAP2ALM = SHARP_YtW, SHARP_Y = 1, SHARP_ALM2MAP = SHARP_Y, SHARP_Yt = 2, SHARP_WY = 3,
               SHARP_ALM2MAP_DERIV1 = 4 } sharp_jobtype;
enum { SHARP_DP = 1 << 4, SHARP_ADD = 1 << 5, SHARP_NO_FFT = 1 << 7 };

void sharp_make_geom_info(int nrings, const int *nph, const ptrdiff_t *ofs, const int *stride, const double *phi0,
                          const double *theta, const double *wgt, sharp_geom_info **geom_info);
void sharp_destroy_geom_info(sharp_geom_info *info);
ptrdiff_t sharp_map_size(const sharp_geom_info *info);

void sharp_make_triangular_alm_info(int lmax, int mmax, int stride, sharp_alm_info **alm_info);
void sharp_destroy_alm_info(sharp_alm_info *info);
ptrdiff_t sharp_alm_count(const sharp_alm_info *self);

void sharp_execute(sharp_jobtype type, int spin, void *alm, void *map, const sharp_geom_info *geom_info,
                   const sharp_alm_info *alm_info, int flags, double *time, unsigned long long *opcnt);

/* status of the last shim call on this thread: 0 = ok (extension, not in libsharp2) */
int pixsht_shim_status(void);

#ifdef __cplusplus
}
#endif
#endif
