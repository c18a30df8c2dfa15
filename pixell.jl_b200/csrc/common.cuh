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

constexpr int LEG_NT = 32;    // threads per Legendre CTA: one warp = one independent work unit (no block barriers)

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

// ---- warp-private asynchronous staging: TMA bulk copy global -> shared, completion on an mbarrier ------------------
#ifndef PIXSHT_EMU
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// lane-0 side: announce `bytes` and start the bulk copy that will complete on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic-proxy reads of dst are ordered before the async write
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "PIXSHT_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra PIXSHT_DONE_%=;\n"
        "bra PIXSHT_WAIT_%=;\n"
        "PIXSHT_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// hint: bring the 128-byte line at p into L2 (no register, no dependency)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#else
inline void prefetch_l2(const void*) {}
inline void mbar_init(unsigned long long*, unsigned) {}
inline void mbar_init_fence() {}
inline void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long*) { memcpy(dst, src, bytes); }
inline void mbar_wait(unsigned long long*, unsigned) { __syncwarp(); }   // emulation: lane 0's memcpy happens-before the readers
#endif

// triangular m-major alm index (Healpix.Alm / make_triangular_alm_info(lmax, mmax, 1))
__host__ __device__ __forceinline__ long long alm_index(int lmax, int l, int m)
{
    return (long long)m * (2LL * lmax + 1 - m) / 2 + l;
}

}  // namespace pixsht
