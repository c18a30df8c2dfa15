// fft_edge.cuh -- the ring-FFT kernels with their two OUTER super-passes fused into the row I/O (sm_100a, HBM bound).
//
// The plain kernels of fft.cuh spend a ring's time in seven strictly serial phases on the one ring an SM can hold at 1'
// (load, pre-processing, four super-passes, store; measured 18 / 13 / 53 / 16 % of the CTA time at C4), each a full
// shared-memory round trip between two block barriers, and the global-memory phases run at the per-SM latency limit
// (4 x 16 B in flight per thread).  Here the ring touches shared memory three times instead of six:
//
//   synthesis (phase -> map), decimation in frequency on natural-order input
//     A  global -> registers -> shared: a thread loads the Q phase entries k = kk + L j of one radix-Q butterfly of the LAST
//        (largest-stride, L = n / Q) pass together with their mirror entries n - k -- which are exactly the butterfly
//        kk' = L - kk, j' = Q - 1 - j -- rotates them by e^{i m phi0} (two-level shared-memory table instead of a second
//        global row), forms the real-FFT pre-processing pair Z[k], Z[n-k] in registers (one complex product per pair:
//        Z[n-k] = conj(E) + i conj(O) when Z[k] = E + i O), does both butterflies and writes the 2 Q results;
//     B  the inner super-passes in shared memory (fft_passes of fft.cuh on a sub-range);
//     C  shared -> registers -> global: a thread reads one contiguous block of the FIRST super-pass (sub-length 1, no pass
//        twiddles), transforms it and stores the results straight into the caller's row.  Thread t takes the block whose
//        outputs are the samples t, t + M, t + 2M, ... (M = n / points), so every store instruction of a warp writes 32
//        neighbouring sample pairs; the un-permutation is the block address perm[t].
//   analysis (map -> phase) is the exact transpose: A' loads the samples t + M r of a block from the caller's row, does the
//     first super-pass and scatters the block; B' inner passes; C' last radix-Q pass on the butterflies kk and L - kk,
//     post-processing pair, quadrature weight, e^{-i m phi0} and the coalesced phase-row stores.
//
// 2 Q (phase side) or up to 16 (map side) independent 16-byte global accesses are in flight per thread, and the arithmetic of
// the outer passes runs under the memory latency of the other warps.  Eligible plans: even ring length that fits shared memory,
// radices 2..5 at both ends, no aliasing (mmax <= nphi / 2), at least two super-passes; everything else (odd or very long rings,
// large prime factors, lmax beyond the ring's Nyquist mode) keeps the kernels of fft.cuh.  Flips, partial rows, the U sign and
// the m-sharded multi-GPU phase layout go through the same helpers (load_pair / store_pair / phase_elem) as there.
#pragma once
#include "fft.cuh"

namespace pixsht {

// e^{+i m phi0} = A[m >> 7] * B[m & 127], m <= mmax, from two small shared-memory tables (double precision for every T: the
// phase rows are double)
// 512 threads leave 128 registers per thread: the 15 / 16-point super-passes do not spill (the time per ring is flat between
// 384 and 640 threads, profiles/r02/fft_edge_threads.txt: the phases are bound by their serial resource use, not by the warp
// count).  Measured and rejected: loading a thread's next butterfly pair during the arithmetic of the current one (ptxas keeps the
// second set of raw entries in local memory: 2 KB of spill traffic per thread), and two pairs per loop iteration with all 4 Q loads
// issued first (no spills at 128 registers, but 5.01 against 4.82 ms at C4, profiles/r02/fft_edge_twopairs.txt).
#ifndef PIXSHT_EF_MAXTHREADS
#define PIXSHT_EF_MAXTHREADS 512
#endif
constexpr int EF_MAXTHREADS = PIXSHT_EF_MAXTHREADS;
constexpr int EF_ROT_LO = 128;
__host__ __device__ __forceinline__ int ef_rot_entries(int mmax) { return EF_ROT_LO + (mmax >> 7) + 1; }
struct RotTab { const double2* A; const double2* B; };
__device__ __forceinline__ RotTab rot_setup(const FftParams& P, double2* tab)
{
    const int na = (P.mmax >> 7) + 1;
    for (int i = threadIdx.x; i < EF_ROT_LO + na; i += blockDim.x)
        tab[i] = (i < EF_ROT_LO) ? P.phi0tw[i <= P.mmax ? i : 0] : P.phi0tw[(i - EF_ROT_LO) << 7];
    RotTab r; r.B = tab; r.A = tab + EF_ROT_LO;
    return r;
}
__device__ __forceinline__ double2 rot_at(const RotTab& R, int m)
{
    const double2 a = R.A[m >> 7], b = R.B[m & 127];
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__device__ __forceinline__ double2 ef_phase_ld(const FftParams& P, double2* row, int ring, int c, int k)
{
    return (k <= P.mmax) ? *phase_elem(P, row, ring, c, k) : make_double2(0.0, 0.0);
}
// rotated half-spectrum entry X[k] from the raw phase value; k == n (m = nphi/2): the ring carries only the doubled real part
template <class T>
__device__ __forceinline__ cpx<T> ef_X(const FftParams& P, const RotTab& R, double2 a, int k, double sg)
{
    cpx<T> v; v.x = (T)0; v.y = (T)0;
    if (k <= P.mmax) {
        const double2 r = rot_at(R, k);
        const double sx = sg * (a.x * r.x - a.y * r.y), sy = sg * (a.x * r.y + a.y * r.x);
        if (k == P.n) v.x = (T)(2.0 * sx);
        else { v.x = (T)sx; v.y = (T)sy; }
    }
    return v;
}
// phase_m = w e^{-i m phi0} F[m]
template <class T>
__device__ __forceinline__ void ef_phase_st(const FftParams& P, const RotTab& R, double2* row, int ring, int c, int k, cpx<T> f, double w)
{
    if (k <= P.mmax) {
        const double2 r = rot_at(R, k);
        const double fx = (double)f.x, fy = (double)f.y;
        *phase_elem(P, row, ring, c, k) = make_double2(w * (fx * r.x + fy * r.y), w * (fy * r.x - fx * r.y));
    }
}

// synthesis pre-processing of the pair (k, n - k), 0 < k < n:  Z[k] = E + i O,  Z[n-k] = conj E + i conj O  with
// E = X[k] + conj X[n-k],  O = (X[k] - conj X[n-k]) e^{+2 pi i k / nphi}   (e^{2 pi i (n-k)/nphi} = -conj of the twiddle of k)
template <class T>
__device__ __forceinline__ void ef_pre_pair(cpx<T> xa, cpx<T> xb, cpx<T> w, cpx<T>& za, cpx<T>& zb)
{
    const cpx<T> e = cadd(xa, cconj(xb)), o = cmul(csub(xa, cconj(xb)), w);
    za = cadd(e, cmuli<T, +1>(o));
    zb = cadd(cconj(e), cmuli<T, +1>(cconj(o)));
}
// analysis post-processing of the pair:  F[k] = (E - i O) / 2,  F[n-k] = (conj E - i conj O) / 2  with
// E = Z[k] + conj Z[n-k],  O = (Z[k] - conj Z[n-k]) e^{-2 pi i k / nphi}
template <class T>
__device__ __forceinline__ void ef_post_pair(cpx<T> za, cpx<T> zb, cpx<T> w, cpx<T>& fa, cpx<T>& fb)
{
    const cpx<T> e = cadd(za, cconj(zb)), o = cmul(csub(za, cconj(zb)), w);
    const cpx<T> a = cadd(e, cmuli<T, -1>(o)), b = cadd(cconj(e), cmuli<T, -1>(cconj(o)));
    fa.x = (T)0.5 * a.x; fa.y = (T)0.5 * a.y;
    fb.x = (T)0.5 * b.x; fb.y = (T)0.5 * b.y;
}

// ---- phase side: the last (largest-stride) pass, single radix Q, butterflies kk and L - kk together --------------------

// raw phase entries of the butterflies kk and L - kk
template <int Q> struct EfRaw { double2 a[Q], b[Q]; };
template <int Q>
__device__ __forceinline__ void ef_synth_pair_load(const FftParams& P, EfRaw<Q>& r, double2* row, int ring, int c, int L, int kk)
{
#pragma unroll
    for (int j = 0; j < Q; ++j) { r.a[j] = ef_phase_ld(P, row, ring, c, kk + L * j); r.b[j] = ef_phase_ld(P, row, ring, c, L - kk + L * j); }
}
template <class T, int Q>
__device__ __forceinline__ void ef_synth_pair(const FftParams& P, const TwTab<T>& W, const RotTab& R, cpx<T>* buf, const EfRaw<Q>& r,
                                              double sg, int L, int kk)
{
    const int n = P.n, kk2 = L - kk;
    cpx<T> za[Q], zb[Q];
#pragma unroll
    for (int j = 0; j < Q; ++j) {
        const int ka = kk + L * j;   // its mirror n - ka is element Q - 1 - j of butterfly kk2
        const cpx<T> xa = ef_X<T>(P, R, r.a[j], ka, sg), xb = ef_X<T>(P, R, r.b[Q - 1 - j], n - ka, sg);
        ef_pre_pair<T>(xa, xb, twid<T, +1>(W, ka), za[j], zb[Q - 1 - j]);
    }
    // decimation in frequency at sub-length L: the root is n = Q L, W_n^kk = tw[2 kk]
    butterfly_regs<T, +1, Q, true>(za, twid<T, +1>(W, 2 * kk), true);
    butterfly_regs<T, +1, Q, true>(zb, twid<T, +1>(W, 2 * kk2), true);
#pragma unroll
    for (int j = 0; j < Q; ++j) { buf[kk + L * j] = za[j]; buf[kk2 + L * j] = zb[j]; }
}
// the self-mirrored butterflies: kk = 0 (element j pairs with Q - j, element 0 with X[n]) and, for even L, kk = L/2 (j with Q-1-j)
template <class T, int Q, bool ZERO>
__device__ __forceinline__ void ef_synth_self(const FftParams& P, const TwTab<T>& W, const RotTab& R, cpx<T>* buf,
                                              double2* row, int ring, int c, double sg, int L)
{
    const int n = P.n, kk = ZERO ? 0 : L / 2;
    cpx<T> x[Q], z[Q];
#pragma unroll
    for (int j = 0; j < Q; ++j) x[j] = ef_X<T>(P, R, ef_phase_ld(P, row, ring, c, kk + L * j), kk + L * j, sg);
#pragma unroll
    for (int j = 0; j < Q; ++j) {
        const int pj = ZERO ? (Q - j) % Q : Q - 1 - j;
        if (ZERO && j == 0) {
            const cpx<T> xn = ef_X<T>(P, R, ef_phase_ld(P, row, ring, c, n), n, sg);
            z[0].x = x[0].x + xn.x; z[0].y = x[0].x - xn.x;
        } else if (pj >= j) {
            cpx<T> zp;
            ef_pre_pair<T>(x[j], x[pj], twid<T, +1>(W, kk + L * j), z[j], zp);
            if (pj != j) z[pj] = zp;
        }
    }
    butterfly_regs<T, +1, Q, true>(z, twid<T, +1>(W, 2 * kk), !ZERO);
#pragma unroll
    for (int j = 0; j < Q; ++j) buf[kk + L * j] = z[j];
}
template <class T, int Q>
__device__ __forceinline__ void ef_synth_edge(const FftParams& P, const TwTab<T>& W, const RotTab& R, cpx<T>* buf,
                                              double2* row, int ring, int c, double sg)
{
    const int L = P.n / Q, nreg = (L - 1) / 2, nitems = nreg + ((L & 1) ? 1 : 2);
    int i = threadIdx.x;
    for (; i < nreg; i += (int)blockDim.x) {
        EfRaw<Q> r;
        ef_synth_pair_load<Q>(P, r, row, ring, c, L, i + 1);
        ef_synth_pair<T, Q>(P, W, R, buf, r, sg, L, i + 1);
    }
    if (i == nreg) ef_synth_self<T, Q, true>(P, W, R, buf, row, ring, c, sg, L);
    else if (i < nitems) ef_synth_self<T, Q, false>(P, W, R, buf, row, ring, c, sg, L);
}

template <class T, int Q>
__device__ __forceinline__ void ef_anal_pair(const FftParams& P, const TwTab<T>& W, const RotTab& R, const cpx<T>* buf,
                                             double2* row, int ring, int c, double w, int L, int kk)
{
    const int n = P.n, kk2 = L - kk;
    cpx<T> za[Q], zb[Q];
#pragma unroll
    for (int j = 0; j < Q; ++j) { za[j] = buf[kk + L * j]; zb[j] = buf[kk2 + L * j]; }
    butterfly_regs<T, -1, Q, false>(za, twid<T, -1>(W, 2 * kk), true);
    butterfly_regs<T, -1, Q, false>(zb, twid<T, -1>(W, 2 * kk2), true);
#pragma unroll
    for (int j = 0; j < Q; ++j) {
        const int ka = kk + L * j;
        cpx<T> fa, fb;
        ef_post_pair<T>(za[j], zb[Q - 1 - j], twid<T, -1>(W, ka), fa, fb);
        ef_phase_st<T>(P, R, row, ring, c, ka, fa, w);
        ef_phase_st<T>(P, R, row, ring, c, n - ka, fb, w);
    }
}
template <class T, int Q, bool ZERO>
__device__ __forceinline__ void ef_anal_self(const FftParams& P, const TwTab<T>& W, const RotTab& R, const cpx<T>* buf,
                                             double2* row, int ring, int c, double w, int L)
{
    const int n = P.n, kk = ZERO ? 0 : L / 2;
    cpx<T> z[Q];
#pragma unroll
    for (int j = 0; j < Q; ++j) z[j] = buf[kk + L * j];
    butterfly_regs<T, -1, Q, false>(z, twid<T, -1>(W, 2 * kk), !ZERO);
#pragma unroll
    for (int j = 0; j < Q; ++j) {
        const int pj = ZERO ? (Q - j) % Q : Q - 1 - j, ka = kk + L * j;
        if (ZERO && j == 0) {
            cpx<T> f0, fn; f0.x = z[0].x + z[0].y; f0.y = (T)0; fn.x = z[0].x - z[0].y; fn.y = (T)0;
            ef_phase_st<T>(P, R, row, ring, c, 0, f0, w);
            ef_phase_st<T>(P, R, row, ring, c, n, fn, w);
        } else if (pj >= j) {
            cpx<T> fa, fb;
            ef_post_pair<T>(z[j], z[pj], twid<T, -1>(W, ka), fa, fb);
            ef_phase_st<T>(P, R, row, ring, c, ka, fa, w);
            if (pj != j) ef_phase_st<T>(P, R, row, ring, c, n - ka, fb, w);
        }
    }
}
template <class T, int Q>
__device__ __forceinline__ void ef_anal_edge(const FftParams& P, const TwTab<T>& W, const RotTab& R, const cpx<T>* buf,
                                             double2* row, int ring, int c, double w)
{
    const int L = P.n / Q, nreg = (L - 1) / 2, nitems = nreg + ((L & 1) ? 1 : 2);
    for (int i = threadIdx.x; i < nitems; i += blockDim.x) {
        if (i < nreg) ef_anal_pair<T, Q>(P, W, R, buf, row, ring, c, w, L, i + 1);
        else if (i == nreg) ef_anal_self<T, Q, true>(P, W, R, buf, row, ring, c, w, L);
        else ef_anal_self<T, Q, false>(P, W, R, buf, row, ring, c, w, L);
    }
}

// ---- map side: the first super-pass (sub-length 1), Q1 x Q2 points (Q2 == 1: a single radix) --------------------------
// block of item t at buf[perm[t]]; its element j2 Q1 + j1 is sample pair t + M (j2 + Q2 j1) of the ring
template <class T, int Q1, int Q2>
__device__ __forceinline__ void ef_synth_store(const FftParams& P, const cpx<T>* buf, T* orow, bool vec)
{
    constexpr int S = Q1 * Q2;
    const int M = P.n / S;
    cpx<T> one; one.x = (T)1; one.y = (T)0;
    for (int t = threadIdx.x; t < M; t += blockDim.x) {
        const cpx<T>* e = buf + P.perm[t];
        cpx<T> a[Q2][Q1];
#pragma unroll
        for (int j2 = 0; j2 < Q2; ++j2)
#pragma unroll
            for (int j1 = 0; j1 < Q1; ++j1) a[j2][j1] = e[j2 * Q1 + j1];
        if constexpr (Q2 > 1) butterfly2_regs<T, +1, Q1, Q2, true, false>(a, one, one);
        else butterfly_regs<T, +1, Q1, true>(a[0], one, false);
#pragma unroll
        for (int j1 = 0; j1 < Q1; ++j1)
#pragma unroll
            for (int j2 = 0; j2 < Q2; ++j2) store_pair<T>(P, orow, t + M * (j2 + Q2 * j1), a[j2][j1], vec);
    }
}
template <class T, int Q1, int Q2>
__device__ __forceinline__ void ef_anal_load(const FftParams& P, cpx<T>* buf, const T* irow, bool vec)
{
    constexpr int S = Q1 * Q2;
    const int M = P.n / S;
    cpx<T> one; one.x = (T)1; one.y = (T)0;
    for (int t = threadIdx.x; t < M; t += blockDim.x) {
        cpx<T>* e = buf + P.perm[t];
        cpx<T> a[Q2][Q1];
#pragma unroll
        for (int j1 = 0; j1 < Q1; ++j1)
#pragma unroll
            for (int j2 = 0; j2 < Q2; ++j2) a[j2][j1] = load_pair<T>(P, irow, t + M * (j2 + Q2 * j1), vec);
        if constexpr (Q2 > 1) butterfly2_regs<T, -1, Q1, Q2, false, false>(a, one, one);
        else butterfly_regs<T, -1, Q1, false>(a[0], one, false);
#pragma unroll
        for (int j2 = 0; j2 < Q2; ++j2)
#pragma unroll
            for (int j1 = 0; j1 < Q1; ++j1) e[j2 * Q1 + j1] = a[j2][j1];
    }
}
template <class T, int Q1, bool SYNTH>
__device__ __forceinline__ void ef_map_q2(const FftParams& P, int q2, cpx<T>* buf, T* row, bool vec)
{
#define EF_CASE(Q2V) \
    if constexpr (Q1 * Q2V <= FFT_FUSE_MAX) { if (q2 == Q2V) { if constexpr (SYNTH) ef_synth_store<T, Q1, Q2V>(P, buf, row, vec); else ef_anal_load<T, Q1, Q2V>(P, buf, row, vec); return; } }
    EF_CASE(1) EF_CASE(2) EF_CASE(3) EF_CASE(4) EF_CASE(5)
#undef EF_CASE
}
template <class T, bool SYNTH>
__device__ __forceinline__ void ef_map_side(const FftParams& P, cpx<T>* buf, T* row, bool vec)
{
    const int t0 = P.sp_first[0], q1 = P.fac[t0], q2 = (P.sp_count[0] == 2) ? P.fac[t0 + 1] : 1;
    if (q1 == 2) ef_map_q2<T, 2, SYNTH>(P, q2, buf, row, vec);
    else if (q1 == 3) ef_map_q2<T, 3, SYNTH>(P, q2, buf, row, vec);
    else if (q1 == 4) ef_map_q2<T, 4, SYNTH>(P, q2, buf, row, vec);
    else ef_map_q2<T, 5, SYNTH>(P, q2, buf, row, vec);
}

// shared memory of these kernels: ring (n + 1) | root twiddles | 2 x 2 pass tables | rotation table (double2, 16-byte aligned)
template <class T>
__host__ __device__ __forceinline__ size_t ef_rot_offset(int n, int nphi, int pt)
{
    const size_t b = (size_t)(n + 1 + fft_tw_entries(nphi) + 4 * pt) * sizeof(cpx<T>);
    return (b + 15) / 16 * 16;
}

// phase -> map (synthesis).  grid = (rows, ncomp): CTA x handles band rings ring_begin + x, x + gridDim.x, ...
template <class T>
__global__ void __launch_bounds__(EF_MAXTHREADS) fft_phase2map_edge(const FftParams P)
{
    PIXSHT_DYN_SMEM(smem_raw);
    cpx<T>* buf = reinterpret_cast<cpx<T>*>(smem_raw);
    cpx<T>* tabs = buf + P.n + 1;
    const TwTab<T> W = tw_setup<T>(P, tabs);
    cpx<T>* ptabs = tabs + fft_tw_entries(P.nphi);
    const RotTab R = rot_setup(P, reinterpret_cast<double2*>(smem_raw + ef_rot_offset<T>(P.n, P.nphi, P.pt)));
    const int c = P.c_begin + blockIdx.y;
    T* out = reinterpret_cast<T*>(P.maps[c]);
    const bool vec = (P.nx == P.nphi) && P.vec_ok;
    const double sg = ((P.neg_mask >> c) & 1) ? -1.0 : 1.0;
    const int Q = P.fac[P.sp_first[P.nsp - 1]];
    FFT_PROF_DECL;
    __syncthreads();   // the tables

    for (int rl = blockIdx.x; rl < P.ring_count; rl += gridDim.x) {
        const int ring = P.ring_begin + rl;
        double2* row = P.phase + ((long long)rl * P.ncomp + c) * P.MP;
        if (Q == 4) ef_synth_edge<T, 4>(P, W, R, buf, row, ring, c, sg);
        else if (Q == 2) ef_synth_edge<T, 2>(P, W, R, buf, row, ring, c, sg);
        else if (Q == 3) ef_synth_edge<T, 3>(P, W, R, buf, row, ring, c, sg);
        else ef_synth_edge<T, 5>(P, W, R, buf, row, ring, c, sg);
        if (P.nsp > 2) superpass_tables<T>(P, ptabs, P.nsp - 2, superpass_L(P, P.nsp - 2));
        __syncthreads();
        FFT_PROF_MARK(0);
        if (P.prefetch && !P.mtab && rl + (int)gridDim.x < P.ring_count)
            prefetch_row(P.phase + ((long long)(rl + gridDim.x) * P.ncomp + c) * P.MP, (size_t)(P.mmax + 1) * sizeof(double2));
        fft_passes<T, +1, true>(P, W, buf, nullptr, ptabs, 1, P.nsp - 2 FFT_PROF_PASS(4));
        const int rowy = P.flipy ? (P.ny - 1 - ring) : ring;
        ef_map_side<T, true>(P, buf, out + (size_t)rowy * P.nx, vec);
        __syncthreads();   // the buffer is reused by this CTA's next ring
        FFT_PROF_MARK(2);
    }
}

// map -> weighted phase (analysis).  grid as above
template <class T>
__global__ void __launch_bounds__(EF_MAXTHREADS) fft_map2phase_edge(const FftParams P)
{
    PIXSHT_DYN_SMEM(smem_raw);
    cpx<T>* buf = reinterpret_cast<cpx<T>*>(smem_raw);
    cpx<T>* tabs = buf + P.n + 1;
    const TwTab<T> W = tw_setup<T>(P, tabs);
    cpx<T>* ptabs = tabs + fft_tw_entries(P.nphi);
    const RotTab R = rot_setup(P, reinterpret_cast<double2*>(smem_raw + ef_rot_offset<T>(P.n, P.nphi, P.pt)));
    const int c = P.c_begin + blockIdx.y;
    T* in = reinterpret_cast<T*>(P.maps[c]);
    const bool vec = (P.nx == P.nphi) && P.vec_ok;
    const double sg = ((P.neg_mask >> c) & 1) ? -1.0 : 1.0;
    const int Q = P.fac[P.sp_first[P.nsp - 1]];
    FFT_PROF_DECL;
    __syncthreads();   // the tables

    for (int rl = blockIdx.x; rl < P.ring_count; rl += gridDim.x) {
        const int ring = P.ring_begin + rl;
        const int rowy = P.flipy ? (P.ny - 1 - ring) : ring;
        ef_map_side<T, false>(P, buf, in + (size_t)rowy * P.nx, vec);
        if (P.nsp > 2) superpass_tables<T>(P, ptabs, 1, superpass_L(P, 1));
        __syncthreads();
        FFT_PROF_MARK(16);
        if (P.prefetch && rl + (int)gridDim.x < P.ring_count) {
            const int ringn = ring + (int)gridDim.x;
            prefetch_row(in + (size_t)(P.flipy ? (P.ny - 1 - ringn) : ringn) * P.nx, (size_t)P.nx * sizeof(T));
        }
        fft_passes<T, -1, false>(P, W, buf, nullptr, ptabs, 1, P.nsp - 2 FFT_PROF_PASS(20));
        const double w = sg * P.wgt[ring];
        double2* row = P.phase + ((long long)rl * P.ncomp + c) * P.MP;
        if (Q == 4) ef_anal_edge<T, 4>(P, W, R, buf, row, ring, c, w);
        else if (Q == 2) ef_anal_edge<T, 2>(P, W, R, buf, row, ring, c, w);
        else if (Q == 3) ef_anal_edge<T, 3>(P, W, R, buf, row, ring, c, w);
        else ef_anal_edge<T, 5>(P, W, R, buf, row, ring, c, w);
        __syncthreads();   // the buffer is reused by this CTA's next ring
        FFT_PROF_MARK(18);
    }
}

}  // namespace pixsht
