// legendre.cuh -- the Legendre stage (alm <-> phase), hand-written for sm_100a.  This is where ~all the time goes
// ("libsharp2 sharp_execute internals, stage B").  FP64 FMA bound; nothing here is a tensor-core shape.
//
// Work decomposition: one CTA = (one m) x (a chunk of LEG_NT*R north/south ring pairs); one thread = R ring pairs of
// one m (pairs lane, lane+32, ... of its warp's contiguous block, so phase I/O is coalesced along the ring index).
// For each pair the scaled functions p_l = lambda_lm(theta)/gamma_l are generated on the fly by
//     p_{l+1} = (alpha_l x + delta_l) p_l - p_{l-1}                       (2 FP64 ops per l and function)
// with per-(l,m) uniform coefficients staged in shared memory, and
//   synthesis: per-ring register accumulators  sum_l p_l * (gamma_l a_lm)   (2 FMA per l, spin 0; 8 for spin 2),
//              north = even + odd, south = even - odd  (equatorial symmetry);
//   analysis : per-l partial sums over the thread's R rings, a warp-private shared-memory transpose-reduction
//              every G (16 / 8) steps, then one atomicAdd per (l, m, CTA).
// Dynamic range: p carries an exponent e (multiple of 64, <= 0); while e < 0 the ring is "seeking" (recurrence only,
// 2 ops per l) and contributes nothing; see common.cuh.  Rings that can never matter for this m are pruned.
#pragma once
#include "common.cuh"

namespace pixsht {

constexpr int LEG_LCA = 128;  // analysis: l values per chunk
struct __align__(16) red4 { double x, y, z, w; };
template <int SPIN> struct RedT { typedef double2 type; static constexpr int G = 16; };   // G: l-steps per warp-level reduction group
template <> struct RedT<2> { typedef red4 type; static constexpr int G = 8; };
template <int SPIN> constexpr size_t leg_anal_smem()
{
    return sizeof(typename RedT<SPIN>::type) * ((size_t)(LEG_NT / 32) * RedT<SPIN>::G * 33 + (size_t)(LEG_NT / 32) * LEG_LCA) + 2 * sizeof(double) * LEG_LCA;
}

struct LegParams {
    int lmax, mmax;
    int nm;                 // number of m values handled by this launch
    const int* m_list;      // device; nullptr => m = row index
    int npairs, nchunks;    // ring pairs; chunks of LEG_NT*R pairs per m
    const double* x;        // [npairs] cos(theta) of the pair's northern member (|x| as stored; may be <0 for lone south rings)
    const double* lsh_hi; const double* lsh_lo;   // log2 sin(theta/2), double-double
    const double* lch_hi; const double* lch_lo;   // log2 cos(theta/2)
    const int* ringN; const int* ringS;           // band ring index of the north/south member, -1 if absent
    const double* mlim;     // [npairs] prune: the pair is skipped for m > mlim
    const double* lgpref_hi; const double* lgpref_lo;   // [mmax+1] log2 of the seed prefactor for this spin family
    const double* alpha; const double* gamma;     // [nalm] recurrence tables of this spin family
    const double* inv_ll1;  // [lmax+1] 2/(l(l+1))
    const double2* alm_in0; const double2* alm_in1;     // synthesis input  (T | E,B)
    double2* alm_out0; double2* alm_out1;               // analysis output  (T | E,B), pre-zeroed, accumulated atomically
    double2* phase;         // element (c,row,ring) at c*stride_c + row*stride_m + ring
    long long stride_c, stride_m;
};

// ---- seeds -------------------------------------------------------------------------------------------------
// value = sign * 2^(lg) with lg = lgpref[m] + a*log2 cos(theta/2) + b*log2 sin(theta/2), returned as (k, frac) with
// lg = k + frac, k integer-valued.  zero => the seed vanishes (pole).
struct LogVal { double k, frac; bool zero; };

__device__ __forceinline__ LogVal seed_log(const LegParams& P, int m, int pair, int a, int b)
{
    LogVal r; r.zero = false; r.k = 0; r.frac = 0;
    dd acc; acc.hi = P.lgpref_hi[m]; acc.lo = P.lgpref_lo[m];
    if (a > 0) {
        dd c; c.hi = P.lch_hi[pair]; c.lo = P.lch_lo[pair];
        if (!(c.hi > -1e300)) { r.zero = true; return r; }
        acc = dd_add(acc, dd_mul_d(c, (double)a));
    }
    if (b > 0) {
        dd s; s.hi = P.lsh_hi[pair]; s.lo = P.lsh_lo[pair];
        if (!(s.hi > -1e300)) { r.zero = true; return r; }
        acc = dd_add(acc, dd_mul_d(s, (double)b));
    }
    r.k = rint(acc.hi);
    r.frac = (acc.hi - r.k) + acc.lo;
    return r;
}

// exponent offset (multiple of 64, <= 0) such that 2^(k - e) < 2^SEEK_THR_LOG2, or 0 when the value is already active
__device__ __forceinline__ int seed_exponent(double k)
{
    if (k >= (double)ACT_LOG2) return 0;
    const double q = ceil(((double)ACT_LOG2 - k) / (double)SEEK_QUANT);
    return (int)(-(double)SEEK_QUANT * q);
}

__device__ __forceinline__ double seed_value(const LogVal& v, int e, double sign)
{
    if (v.zero) return 0.0;
    const double d = v.k - (double)e;
    if (d < -1000.0) return 0.0;
    return sign * ldexp(exp2(v.frac), (int)d);
}

template <int SPIN, int R>
struct RingState {
    double x[R];
    double p[(SPIN == 0 ? 1 : 2)][R];
    double pp[(SPIN == 0 ? 1 : 2)][R];
    int e[R];
};

template <int SPIN, int R>
__device__ __forceinline__ void init_rings(const LegParams& P, int m, int pair0, int lane, RingState<SPIN, R>& S)
{
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const int pair = pair0 + j * 32 + lane;
        S.x[j] = 0.0; S.e[j] = E_DEAD;
        S.p[0][j] = 0.0; S.pp[0][j] = 0.0;
        if (SPIN != 0) { S.p[SPIN != 0][j] = 0.0; S.pp[SPIN != 0][j] = 0.0; }
        if (pair < P.npairs && (double)m <= P.mlim[pair]) {
            S.x[j] = P.x[pair];
            const double sgn = (m & 1) ? -1.0 : 1.0;
            if (SPIN == 0) {
                // lambda_mm = (-1)^m N_m sin^m(theta), sin(theta) = 2 sin(theta/2) cos(theta/2): the factor 2^m is in lgpref
                LogVal v = seed_log(P, m, pair, m, m);
                if (!v.zero) {
                    const int e = seed_exponent(v.k);
                    S.p[0][j] = seed_value(v, e, sgn);
                    S.e[j] = e;
                }
            } else {
                // l0 = max(m,2):  lambda^+ ~ cos^{|m-2|} sin^{m+2},  lambda^- ~ cos^{m+2} sin^{|m-2|}  (half angles)
                const int am = m >= 2 ? m - 2 : 2 - m;
                LogVal vp = seed_log(P, m, pair, am, m + 2);
                LogVal vm = seed_log(P, m, pair, m + 2, am);
                const double sp = sgn, sm = (m >= 2) ? sgn : 1.0;
                if (!(vp.zero && vm.zero)) {
                    double kmax = vp.zero ? vm.k : (vm.zero ? vp.k : fmax(vp.k, vm.k));
                    const int e = seed_exponent(kmax);
                    S.p[0][j] = seed_value(vp, e, sp);
                    S.p[SPIN != 0][j] = seed_value(vm, e, sm);
                    S.e[j] = e;
                }
            }
        }
    }
}

__device__ __forceinline__ bool over_thr(double v)
{
    return ((unsigned)__double2hiint(v) & 0x7ff00000u) >= SEEK_THR_EXPBITS;
}

// one recurrence step for ring slot j (+ rescale check when CHECK)
template <int SPIN, int R, bool CHECK>
__device__ __forceinline__ void rec_step(RingState<SPIN, R>& S, int j, double alpha, double delta)
{
    if (SPIN == 0) {
        const double u = alpha * S.x[j];
        const double pn = fma(u, S.p[0][j], -S.pp[0][j]);
        S.pp[0][j] = S.p[0][j]; S.p[0][j] = pn;
        if (CHECK) {
            if (S.e[j] < 0 && over_thr(pn)) {
                const double sc = 5.421010862427522e-20;  // 2^-64
                S.p[0][j] *= sc; S.pp[0][j] *= sc; S.e[j] += SEEK_QUANT;
            }
        }
    } else {
        const double up = fma(alpha, S.x[j], delta);
        const double um = fma(alpha, S.x[j], -delta);
        const double pn = fma(up, S.p[0][j], -S.pp[0][j]);
        const double mn = fma(um, S.p[SPIN != 0][j], -S.pp[SPIN != 0][j]);
        S.pp[0][j] = S.p[0][j]; S.p[0][j] = pn;
        S.pp[SPIN != 0][j] = S.p[SPIN != 0][j]; S.p[SPIN != 0][j] = mn;
        if (CHECK) {
            if (S.e[j] < 0 && (over_thr(pn) || over_thr(mn))) {
                const double sc = 5.421010862427522e-20;
                S.p[0][j] *= sc; S.pp[0][j] *= sc; S.p[SPIN != 0][j] *= sc; S.pp[SPIN != 0][j] *= sc; S.e[j] += SEEK_QUANT;
            }
        }
    }
}

template <int SPIN, int R>
__device__ __forceinline__ void ring_flags(const RingState<SPIN, R>& S, bool& any_seek, bool& any_act)
{
    bool s = false, a = false;
#pragma unroll
    for (int j = 0; j < R; ++j) { s |= (S.e[j] < 0); a |= (S.e[j] == 0); }
    any_seek = __any_sync(0xffffffffu, s);
    any_act = __any_sync(0xffffffffu, a);
}

// stage alpha / delta for l = l0+c0 .. l0+c0+LEG_LC-1 (zeros beyond lmax)
template <int SPIN>
__device__ __forceinline__ void stage_coef(const LegParams& P, int m, int l0, int c0, double* sA, double* sD)
{
    const long long base = alm_index(P.lmax, 0, m);
    for (int i = threadIdx.x; i < LEG_LC; i += LEG_NT) {
        const int l = l0 + c0 + i;
        double a = 0.0, d = 0.0;
        if (l <= P.lmax) {
            a = P.alpha[base + l];
            if (SPIN != 0) d = a * (double)m * P.inv_ll1[l];   // delta^+ = -alpha mu^+ = alpha * 2m/(l(l+1))
        }
        sA[i] = a;
        if (SPIN != 0) sD[i] = d;
    }
}

// =============================================================================================================
// synthesis: alm -> phase
// =============================================================================================================
template <int SPIN, int R, int MODE, int PAR>
__device__ __forceinline__ void synth_step(RingState<SPIN, R>& S, double (&acc)[(SPIN == 0 ? 4 : 8)][R], double alpha, double delta,
                                           double2 g0, double2 g1)
{
#pragma unroll
    for (int j = 0; j < R; ++j) {
        if (MODE != 0) {
            double p0 = S.p[0][j];
            double p1 = S.p[SPIN != 0][j];
            if (MODE == 1) { if (S.e[j] != 0) { p0 = 0.0; p1 = 0.0; } }
            if (SPIN == 0) {
                acc[2 * PAR + 0][j] = fma(p0, g0.x, acc[2 * PAR + 0][j]);
                acc[2 * PAR + 1][j] = fma(p0, g0.y, acc[2 * PAR + 1][j]);
            } else {
                // north: S+ += p+ G+, S- += p- G-;  south: T+ += sgn p- G+, T- += sgn p+ G-  (sgn alternates with l)
                acc[0][j] = fma(p0, g0.x, acc[0][j]);
                acc[1][j] = fma(p0, g0.y, acc[1][j]);
                constexpr int A4 = (SPIN == 0 ? 0 : 4);   // keeps indices in range in the (dead) SPIN == 0 instantiation
                acc[2][j] = fma(p1, g1.x, acc[2][j]);
                acc[3][j] = fma(p1, g1.y, acc[3][j]);
                if (PAR == 0) {
                    acc[A4 + 0][j] = fma(p1, g0.x, acc[A4 + 0][j]);
                    acc[A4 + 1][j] = fma(p1, g0.y, acc[A4 + 1][j]);
                    acc[A4 + 2][j] = fma(p0, g1.x, acc[A4 + 2][j]);
                    acc[A4 + 3][j] = fma(p0, g1.y, acc[A4 + 3][j]);
                } else {
                    acc[A4 + 0][j] = fma(-p1, g0.x, acc[A4 + 0][j]);
                    acc[A4 + 1][j] = fma(-p1, g0.y, acc[A4 + 1][j]);
                    acc[A4 + 2][j] = fma(-p0, g1.x, acc[A4 + 2][j]);
                    acc[A4 + 3][j] = fma(-p0, g1.y, acc[A4 + 3][j]);
                }
            }
        }
        rec_step<SPIN, R, (MODE != 2)>(S, j, alpha, delta);
    }
}

template <int SPIN, int R>
__global__ void __launch_bounds__(LEG_NT) leg_synth(const LegParams P)
{
    constexpr int NACC = (SPIN == 0) ? 4 : 8;
    __shared__ double sA[LEG_LC];
    __shared__ double sD[LEG_LC];
    __shared__ double2 sG0[LEG_LC];
    __shared__ double2 sG1[LEG_LC];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = blockIdx.x / P.nchunks, chunk = blockIdx.x % P.nchunks;
    const int m = P.m_list ? P.m_list[row] : row;
    const int l0 = (SPIN == 0) ? m : (m > 2 ? m : 2);
    const int pair0 = chunk * (LEG_NT * R) + warp * (32 * R);

    RingState<SPIN, R> S;
    init_rings<SPIN, R>(P, m, pair0, lane, S);
    double acc[NACC][R];
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
        for (int j = 0; j < R; ++j) acc[a][j] = 0.0;

    bool any_seek, any_act;
    ring_flags<SPIN, R>(S, any_seek, any_act);
    const bool warp_live = any_seek || any_act;
    const int nl = P.lmax - l0 + 1;
    const int block_live = __syncthreads_or(warp_live ? 1 : 0);

    if (block_live && nl > 0) {
        const long long abase = alm_index(P.lmax, 0, m);
        for (int c0 = 0; c0 < nl; c0 += LEG_LC) {
            stage_coef<SPIN>(P, m, l0, c0, sA, sD);
            for (int i = tid; i < LEG_LC; i += LEG_NT) {
                const int l = l0 + c0 + i;
                double2 g0 = make_double2(0.0, 0.0), g1 = make_double2(0.0, 0.0);
                if (l <= P.lmax) {
                    const double g = P.gamma[abase + l];
                    if (SPIN == 0) {
                        const double2 a = P.alm_in0[abase + l];
                        g0 = make_double2(g * a.x, (m == 0) ? 0.0 : g * a.y);
                    } else {
                        // G+- = -gamma (E +- iB)/2
                        const double2 E = P.alm_in0[abase + l], B = P.alm_in1[abase + l];
                        const double h = -0.5 * g;
                        g0 = make_double2(h * (E.x - B.y), h * (E.y + B.x));
                        g1 = make_double2(h * (E.x + B.y), h * (E.y - B.x));
                    }
                }
                sG0[i] = g0;
                if (SPIN != 0) sG1[i] = g1;
            }
            __syncthreads();
            if (warp_live) {
                int cnt = nl - c0; if (cnt > LEG_LC) cnt = LEG_LC;
                // steps are taken in (even, odd) pairs; LEG_LC is even and the staged arrays are zero-padded
                for (int i = 0; i < cnt; i += 2) {
                    const double a0 = sA[i], a1 = sA[i + 1];
                    const double d0 = (SPIN != 0) ? sD[i] : 0.0, d1 = (SPIN != 0) ? sD[i + 1] : 0.0;
                    if (!any_seek) {
                        synth_step<SPIN, R, 2, 0>(S, acc, a0, d0, sG0[i], (SPIN != 0) ? sG1[i] : sG0[i]);
                        synth_step<SPIN, R, 2, 1>(S, acc, a1, d1, sG0[i + 1], (SPIN != 0) ? sG1[i + 1] : sG0[i + 1]);
                    } else {
                        if (!any_act) {
                            synth_step<SPIN, R, 0, 0>(S, acc, a0, d0, sG0[i], sG0[i]);
                            synth_step<SPIN, R, 0, 1>(S, acc, a1, d1, sG0[i], sG0[i]);
                        } else {
                            synth_step<SPIN, R, 1, 0>(S, acc, a0, d0, sG0[i], (SPIN != 0) ? sG1[i] : sG0[i]);
                            synth_step<SPIN, R, 1, 1>(S, acc, a1, d1, sG0[i + 1], (SPIN != 0) ? sG1[i + 1] : sG0[i + 1]);
                        }
                        ring_flags<SPIN, R>(S, any_seek, any_act);
                    }
                }
            }
            __syncthreads();
        }
    }

    // ---- write phase (zeros for pruned / never-activated rings) ----
    double2* ph = P.phase + (long long)row * P.stride_m;
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const int pair = pair0 + j * 32 + lane;
        if (pair >= P.npairs) continue;
        const int rN = P.ringN[pair], rS = P.ringS[pair];
        if (SPIN == 0) {
            const double er = acc[0][j], ei = acc[1][j], orr = acc[2][j], oi = acc[3][j];
            if (rN >= 0) ph[rN] = make_double2(er + orr, ei + oi);
            if (rS >= 0) ph[rS] = make_double2(er - orr, ei - oi);
        } else {
            // q = S+ + S-, u = -i (S+ - S-);  south: base sign (-1)^(l0+m) times the alternating sums
            const double bs = ((l0 + m) & 1) ? -1.0 : 1.0;
            if (rN >= 0) {
                ph[rN] = make_double2(acc[0][j] + acc[2][j], acc[1][j] + acc[3][j]);
                ph[P.stride_c + rN] = make_double2(acc[1][j] - acc[3][j], -(acc[0][j] - acc[2][j]));
            }
            constexpr int A4 = (SPIN == 0 ? 0 : 4);
            if (rS >= 0) {
                ph[rS] = make_double2(bs * (acc[A4 + 0][j] + acc[A4 + 2][j]), bs * (acc[A4 + 1][j] + acc[A4 + 3][j]));
                ph[P.stride_c + rS] = make_double2(bs * (acc[A4 + 1][j] - acc[A4 + 3][j]), -bs * (acc[A4 + 0][j] - acc[A4 + 2][j]));
            }
        }
    }
}

// =============================================================================================================
// analysis: (weighted) phase -> alm
// =============================================================================================================
template <int SPIN, int R, int MODE, int PAR>
__device__ __forceinline__ void anal_step(RingState<SPIN, R>& S, const double (&X)[(SPIN == 0 ? 4 : 8)][R], double alpha, double delta,
                                          double (&part)[(SPIN == 0 ? 2 : 4)])
{
#pragma unroll
    for (int j = 0; j < R; ++j) {
        if (MODE != 0) {
            double p0 = S.p[0][j];
            double p1 = S.p[SPIN != 0][j];
            if (MODE == 1) { if (S.e[j] != 0) { p0 = 0.0; p1 = 0.0; } }
            if (SPIN == 0) {
                // even (l-m): X_N + X_S ; odd: X_N - X_S
                part[0] = fma(p0, X[2 * PAR + 0][j], part[0]);
                part[1] = fma(p0, X[2 * PAR + 1][j], part[1]);
            } else {
                // a+ += p+ Y+_N + sgn p- Y+_S ;  a- += p- Y-_N + sgn p+ Y-_S   (Y_S pre-multiplied by the base sign)
                part[0] = fma(p0, X[0][j], part[0]);
                part[1] = fma(p0, X[1][j], part[1]);
                constexpr int NXX = (SPIN == 0 ? 4 : 8), NPP = (SPIN == 0 ? 2 : 4);
                part[NPP - 2] = fma(p1, X[NXX - 4][j], part[NPP - 2]);
                part[NPP - 1] = fma(p1, X[NXX - 3][j], part[NPP - 1]);
                if (PAR == 0) {
                    part[0] = fma(p1, X[2][j], part[0]);
                    part[1] = fma(p1, X[3][j], part[1]);
                    part[NPP - 2] = fma(p0, X[NXX - 2][j], part[NPP - 2]);
                    part[NPP - 1] = fma(p0, X[NXX - 1][j], part[NPP - 1]);
                } else {
                    part[0] = fma(-p1, X[2][j], part[0]);
                    part[1] = fma(-p1, X[3][j], part[1]);
                    part[NPP - 2] = fma(-p0, X[NXX - 2][j], part[NPP - 2]);
                    part[NPP - 1] = fma(-p0, X[NXX - 1][j], part[NPP - 1]);
                }
            }
        }
        rec_step<SPIN, R, (MODE != 2)>(S, j, alpha, delta);
    }
}

template <int SPIN, int R>
__global__ void __launch_bounds__(LEG_NT) leg_anal(const LegParams P)
{
    constexpr int NX = (SPIN == 0) ? 4 : 8;
    constexpr int NPART = (SPIN == 0) ? 2 : 4;
    constexpr int NW = LEG_NT / 32;
    constexpr int G = RedT<SPIN>::G;
    typedef typename RedT<SPIN>::type red_t;
    PIXSHT_DYN_SMEM(smem_raw);
    red_t* red = reinterpret_cast<red_t*>(smem_raw);              // [NW][G][33]
    red_t* outw = red + (size_t)NW * G * 33;                      // [NW][LEG_LCA]
    double* sA = reinterpret_cast<double*>(outw + (size_t)NW * LEG_LCA);   // [LEG_LCA]
    double* sD = sA + LEG_LCA;                                    // [LEG_LCA]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = blockIdx.x / P.nchunks, chunk = blockIdx.x % P.nchunks;
    const int m = P.m_list ? P.m_list[row] : row;
    const int l0 = (SPIN == 0) ? m : (m > 2 ? m : 2);
    const int pair0 = chunk * (LEG_NT * R) + warp * (32 * R);
    red_t* wred = red + (size_t)warp * G * 33;
    red_t* wout = outw + (size_t)warp * LEG_LCA;

    RingState<SPIN, R> S;
    init_rings<SPIN, R>(P, m, pair0, lane, S);

    // folded inputs
    double X[NX][R];
    const double2* ph = P.phase + (long long)row * P.stride_m;
    const double bs = ((l0 + m) & 1) ? -1.0 : 1.0;
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const int pair = pair0 + j * 32 + lane;
        double2 qN = make_double2(0.0, 0.0), qS = qN, uN = qN, uS = qN;
        if (pair < P.npairs && S.e[j] != E_DEAD) {
            const int rN = P.ringN[pair], rS = P.ringS[pair];
            if (rN >= 0) { qN = ph[rN]; if (SPIN != 0) uN = ph[P.stride_c + rN]; }
            if (rS >= 0) { qS = ph[rS]; if (SPIN != 0) uS = ph[P.stride_c + rS]; }
        }
        if (SPIN == 0) {
            X[0][j] = qN.x + qS.x; X[1][j] = qN.y + qS.y;
            X[2][j] = qN.x - qS.x; X[3][j] = qN.y - qS.y;
        } else {
            // Y+ = Xq + i Xu ; Y- = Xq - i Xu
            X[0][j] = qN.x - uN.y; X[1][j] = qN.y + uN.x;                              // Y+_N
            X[2][j] = bs * (qS.x - uS.y); X[3][j] = bs * (qS.y + uS.x);                // Y+_S
            X[NX - 4][j] = qN.x + uN.y; X[NX - 3][j] = qN.y - uN.x;                    // Y-_N
            X[NX - 2][j] = bs * (qS.x + uS.y); X[NX - 1][j] = bs * (qS.y - uS.x);      // Y-_S
        }
    }

    bool any_seek, any_act;
    ring_flags<SPIN, R>(S, any_seek, any_act);
    const bool warp_live = any_seek || any_act;
    const int nl = P.lmax - l0 + 1;
    const int block_live = __syncthreads_or(warp_live ? 1 : 0);
    if (!block_live || nl <= 0) return;   // whole CTA: nothing to add (outputs are pre-zeroed)

    const long long abase = alm_index(P.lmax, 0, m);
    for (int c0 = 0; c0 < nl; c0 += LEG_LCA) {
        // stage alpha / delta (zeros beyond lmax) and clear this warp's chunk accumulators
        for (int i = tid; i < LEG_LCA; i += LEG_NT) {
            const int l = l0 + c0 + i;
            double a = 0.0, d = 0.0;
            if (l <= P.lmax) {
                a = P.alpha[abase + l];
                if (SPIN != 0) d = a * (double)m * P.inv_ll1[l];
            }
            sA[i] = a; sD[i] = d;
        }
        {
            red_t z; memset(&z, 0, sizeof(z));
            for (int i = lane; i < LEG_LCA; i += 32) wout[i] = z;
        }
        __syncthreads();
        int cnt = nl - c0; if (cnt > LEG_LCA) cnt = LEG_LCA;
        if (warp_live) {
            for (int g0 = 0; g0 < cnt; g0 += G) {
                int wrote_from = G;   // first step of this group whose partials were written
#pragma unroll 1
                for (int s = 0; s < G; s += 2) {
                    const int i = g0 + s;
                    const double a0 = sA[i], a1 = sA[i + 1];
                    const double d0 = sD[i], d1 = sD[i + 1];
                    double part0[NPART], part1[NPART];
#pragma unroll
                    for (int k = 0; k < NPART; ++k) { part0[k] = 0.0; part1[k] = 0.0; }
                    if (!any_act) {
                        anal_step<SPIN, R, 0, 0>(S, X, a0, d0, part0);
                        anal_step<SPIN, R, 0, 1>(S, X, a1, d1, part1);
                        ring_flags<SPIN, R>(S, any_seek, any_act);
                        if (any_act) wrote_from = s + 2;
                        continue;
                    }
                    if (any_seek) {
                        anal_step<SPIN, R, 1, 0>(S, X, a0, d0, part0);
                        anal_step<SPIN, R, 1, 1>(S, X, a1, d1, part1);
                        ring_flags<SPIN, R>(S, any_seek, any_act);
                    } else {
                        anal_step<SPIN, R, 2, 0>(S, X, a0, d0, part0);
                        anal_step<SPIN, R, 2, 1>(S, X, a1, d1, part1);
                    }
                    if (wrote_from == G) wrote_from = s;
                    red_t v0, v1;
                    memcpy(&v0, part0, sizeof(red_t)); memcpy(&v1, part1, sizeof(red_t));
                    wred[s * 33 + lane] = v0;
                    wred[(s + 1) * 33 + lane] = v1;
                }
                if (wrote_from < G) {
                    __syncwarp();
                    // transpose-reduce: lane -> (step lq = lane % G, source slice = lane / G of 32/G slices)
                    constexpr int NSL = 32 / G, SL = 32 / NSL;   // slices, sources per slice (= G)
                    const int lq = lane % G, slice = lane / G;
                    double t[NPART];
#pragma unroll
                    for (int k = 0; k < NPART; ++k) t[k] = 0.0;
                    if (lq >= wrote_from) {
#pragma unroll 4
                        for (int k = 0; k < SL; ++k) {
                            const red_t v = wred[lq * 33 + slice * SL + k];
                            const double* vd = reinterpret_cast<const double*>(&v);
#pragma unroll
                            for (int q = 0; q < NPART; ++q) t[q] += vd[q];
                        }
                    }
#pragma unroll
                    for (int off = G; off < 32; off <<= 1)
#pragma unroll
                        for (int q = 0; q < NPART; ++q) t[q] += __shfl_xor_sync(0xffffffffu, t[q], off);
                    if (slice == 0 && lq >= wrote_from) {
                        red_t o; memcpy(&o, t, sizeof(red_t));
                        wout[g0 + lq] = o;
                    }
                    __syncwarp();
                }
            }
        }
        __syncthreads();
        // combine the warps, scale by gamma_l, accumulate into the global alm
        for (int i = tid; i < cnt; i += LEG_NT) {
            const int l = l0 + c0 + i;
            double t[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const double* v = reinterpret_cast<const double*>(&outw[(size_t)w * LEG_LCA + i]);
#pragma unroll
                for (int k = 0; k < NPART; ++k) t[k] += v[k];
            }
            const double g = P.gamma[abase + l];
            if (SPIN == 0) {
                if (t[0] != 0.0) atomicAdd(&P.alm_out0[abase + l].x, g * t[0]);
                if (t[1] != 0.0 && m != 0) atomicAdd(&P.alm_out0[abase + l].y, g * t[1]);
            } else {
                // E = -(a+ + a-)/2 ; B = i (a+ - a-)/2
                const double h = 0.5 * g;
                const double er = -h * (t[0] + t[2]), ei = -h * (t[1] + t[3]);
                const double br = -h * (t[1] - t[3]), bi = h * (t[0] - t[2]);
                if (er != 0.0) atomicAdd(&P.alm_out0[abase + l].x, er);
                if (ei != 0.0 && m != 0) atomicAdd(&P.alm_out0[abase + l].y, ei);
                if (br != 0.0) atomicAdd(&P.alm_out1[abase + l].x, br);
                if (bi != 0.0 && m != 0) atomicAdd(&P.alm_out1[abase + l].y, bi);
            }
        }
        __syncthreads();
    }
}

}  // namespace pixsht
