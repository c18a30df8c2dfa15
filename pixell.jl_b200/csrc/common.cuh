// common.cuh -- shared definitions for the pixsht kernels (sm_100a).
#pragma once
#ifndef PIXSHT_EMU
#include <cuda_runtime.h>
#define PIXSHT_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define PIXSHT_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace pixsht {

// ---- activation / rescaling constants of the scaled recurrences (DESIGN.md "dynamic range") ----------------
// A recurrence value is carried as p * 2^e with e <= 0 a multiple of 64.  While e < 0 the ring is "seeking":
// whenever |p| >= 2^SEEK_THR_LOG2 it is multiplied by 2^-64 and e += 64.  It becomes active (contributes) when
// e reaches 0, i.e. once the true value has grown to >= 2^(SEEK_THR_LOG2-64) = 2^-90.
constexpr int SEEK_THR_LOG2 = -26;
constexpr int SEEK_QUANT = 64;
constexpr int ACT_LOG2 = SEEK_THR_LOG2 - SEEK_QUANT;  // -90
constexpr unsigned SEEK_THR_EXPBITS = (unsigned)(1023 + SEEK_THR_LOG2) << 20;
constexpr int E_DEAD = 1;  // marks a ring slot that never contributes (padding / pruned / zero seed)

constexpr int LEG_NT = 128;   // threads per Legendre block
constexpr int LEG_LC = 256;   // l values staged in shared memory per chunk

// ---- double-double helpers (used only for seeds: O(1) per (m, ring)) -----------------------------------------
struct dd { double hi, lo; };
__host__ __device__ __forceinline__ dd two_sum(double a, double b)
{
    double s = a + b, bb = s - a;
    dd r; r.hi = s; r.lo = (a - (s - bb)) + (b - bb);
    return r;
}
// n * (hi, lo) for an integer-valued double n
__host__ __device__ __forceinline__ dd dd_mul_d(dd a, double n)
{
    double p = a.hi * n;
    double e = fma(a.hi, n, -p);
    dd r; r.hi = p; r.lo = e + a.lo * n;
    return r;
}
__host__ __device__ __forceinline__ dd dd_add(dd a, dd b)
{
    dd s = two_sum(a.hi, b.hi);
    s.lo += a.lo + b.lo;
    return s;
}

// triangular m-major alm index (Healpix.Alm / make_triangular_alm_info(lmax, mmax, 1))
__host__ __device__ __forceinline__ long long alm_index(int lmax, int l, int m)
{
    return (long long)m * (2LL * lmax + 1 - m) / 2 + l;
}

}  // namespace pixsht
