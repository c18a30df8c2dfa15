// legendre_2s.cuh -- spin-0 Legendre stage in steps of TWO degrees: 6 instead of 8 FP64 operations per (two l, m, ring pair).
//
// For fixed m the normalised functions obey  x lambda_l = c_{l+1} lambda_{l+1} + c_l lambda_{l-1},  c_l^2 = (l^2-m^2)/(4l^2-1).
// Let l_k = m + 2k and  h_k(x) = lambda_{l_k+1}(x) / x  -- an even function of x, a polynomial in x^2 times sin^m.  Then
//     odd  (l - m odd):   lambda_{l_k+1} = x h_k                                        (by definition)
//     even (l - m even):  lambda_{l_k}   = c_{l_k+1} h_k + c_{l_k} h_{k-1}              (the relation above, divided by x)
//     recurrence in x^2:  c_{l+2} c_{l+3} h_{k+1} = (x^2 - c_{l+1}^2 - c_{l+2}^2) h_k - c_l c_{l+1} h_{k-1},   l = l_k,
//     seed:               h_0 = lambda_mm / c_{m+1} = sqrt(2m+3) lambda_mm.
// So ONE sequence serves both parities:  sum_l a_l lambda_l = sum_k h_k [ (c_{l_k+1} a_{l_k} + c_{l_k+2} a_{l_k+2}) + x a_{l_k+1} ]
// -- the even and the odd part of the sum over l, i.e. exactly the north/south fold of the standard kernels (north = even + odd,
// south = even - odd).  Scaled as in legendre.cuh (h_k = gamma_k p_k, unit lower coefficient):
//     p_{k+1} = (alpha_k x^2 + beta_k) p_k - p_{k-1}                                     2 FP64 ops per k and ring pair
//     synthesis:  acc_e += p_k Ge_k (re, im),  acc_o += p_k Go_k (re, im)                4
//     analysis :  Se_k = sum_rings p_k (X_N + X_S),  So_k = sum_rings p_k x (X_N - X_S)  4
// with Ge_k = gamma_k (c_{l_k+1} a_{l_k} + c_{l_k+2} a_{l_k+2}), Go_k = gamma_k a_{l_k+1} prepared once per call, and on analysis
// a_{l_k} += gamma_k c_{l_k+1} Se_k,  a_{l_k+2} += gamma_k c_{l_k+2} Se_k,  a_{l_k+1} += gamma_k So_k  (all local, no solve).
// This is the scheme of libsharp2's successor (ducc0) for spin 0, restated from the recurrences; nothing here is copied.
//
// Accuracy: the recurrence in x^2 is to the equator what the one in x is to the poles -- the even functions are formed from
// h_k ~ lambda / x, and digits are lost like 1/|x| for |x| -> 0 (measured in long double, lmax 10800: 6e-11 at x = 0, 6e-13 at
// 0.01, <= 1e-13 from 0.03 on, the level of the standard recurrence).  The launch logic (pixsht.cu) therefore runs the chunks that
// contain ring pairs with |cos theta| < TWOSTEP_XMIN through the standard kernels and everything else through these.
//
// Data: per (m, k) tables at index alm_index(lmax, 0, m) + m + k ("pseudo degree" m + k, k <= (lmax - m)/2), so that the activation
// table, the TMA record streams and the work decomposition of legendre.cuh carry over with lmax replaced by m + (lmax - m)/2.
#pragma once
#include "legendre.cuh"

namespace pixsht {

constexpr double TWOSTEP_XMIN = 0.05;
constexpr double TWOSTEP_POLE_DEG = 3.0;   // and the chunks that reach within 3 degrees of a pole: the recurrence in x^2 is a few times less accurate than
                                           // the one in x there, on rings whose error (l^1.5 eps) is the largest of the map already

__host__ __device__ __forceinline__ int twostep_lmax(int lmax, int m) { return m + ((lmax - m) >> 1); }   // last pseudo degree of column m

// ---- plan time: coefficient tables ------------------------------------------------------------------------------------
__device__ __forceinline__ double c2_lm(int l, int m)
{
    const double L = (double)l, M = (double)m;
    return ((L - M) * (L + M)) / (4.0 * L * L - 1.0);
}
// per m (one thread each): (alpha_k, beta_k), (gamma_k c_{l_k+1}, gamma_k c_{l_k+2}) and gamma_k for k = 0 .. (lmax-m)/2
__global__ void k_coef_tables_2s(int lmax, int mmax, double2* __restrict__ ad, double2* __restrict__ wg, double* __restrict__ gam)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m > mmax) return;
    const long long base = alm_index(lmax, 0, m) + m;
    const int K = ((lmax - m) >> 1) + 1;
    double g_km1 = 1.0, g_k = 1.0;
    for (int k = 0; k < K; ++k) {
        const int l = m + 2 * k;
        const double q0 = c2_lm(l, m), q1 = c2_lm(l + 1, m), q2 = c2_lm(l + 2, m), q3 = c2_lm(l + 3, m);
        const double e = sqrt(q0 * q1), f = sqrt(q2 * q3), d = q1 + q2;
        const double g_kp1 = (k == 0) ? 1.0 : (e / f) * g_km1;
        const double a = g_k / (g_kp1 * f);
        ad[base + k] = make_double2(a, -a * d);
        wg[base + k] = make_double2(g_k * sqrt(q1), g_k * sqrt(q2));
        gam[base + k] = g_k;
        g_km1 = g_k; g_k = g_kp1;
    }
    // the rest of the column (pseudo degrees that do not exist): zeros, never read by the kernels
    for (int k = K; k <= lmax - m; ++k) { ad[base + k] = make_double2(0.0, 0.0); wg[base + k] = make_double2(0.0, 0.0); gam[base + k] = 0.0; }
}

// ---- plan time: activation table of the h sequence (same state machine as k_seek_table<0>, recurrence in x^2) -----------
__global__ void __launch_bounds__(128) k_seek_table_2s(const SeekParams P)
{
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y;
    if (pair >= P.npairs) return;
    const int lmaxp = twostep_lmax(P.lmax, m);
    const int act_log2 = P.thr_log2 - SEEK_QUANT;
    const unsigned expbits = (unsigned)(1023 + P.thr_log2) << 20;
    double p0 = 0.0, q0 = 0.0;
    int e = E_DEAD;
    if ((double)m <= P.mlim[pair]) {
        // h_0 = sqrt(2m+3) lambda_mm: the factor is in lgpref (pixsht.cu)
        LogVal v = seed_log(P, m, pair, m, m);
        if (!v.zero) { e = seed_exponent(v.k, act_log2); p0 = seed_value(v, e, (m & 1) ? -1.0 : 1.0); }
    }
    int l = m;   // pseudo degree m + k
    if (e < 0) {
        const double x = P.x[pair], x2 = x * x;
        const double2* ad = P.ad + alm_index(P.lmax, 0, m);
        const double sc = 5.421010862427522e-20;  // 2^-64
        while (e < 0 && l <= lmaxp) {
            const double2 c = ad[l];
            const double pn = fma(fma(c.x, x2, c.y), p0, -q0);
            q0 = p0; p0 = pn;
            if (over_thr(pn, expbits)) { p0 *= sc; q0 *= sc; e += SEEK_QUANT; }
            ++l;
        }
    }
    const size_t k = (size_t)m * P.npairs + pair;
    const bool live = (e == 0) && (l <= lmaxp);
    P.lact[k] = live ? l : L_NEVER;
    reinterpret_cast<double2*>(P.st)[k] = live ? make_double2(p0, q0) : make_double2(0.0, 0.0);
}

// ---- per call: synthesis records { alpha_k, beta_k, Ge.re, Ge.im, Go.re, Go.im } ---------------------------------------------
// grid (x: k in pieces, y: position in the m list / offset from m_begin)
__global__ void k_prep_synth_2s(const int* __restrict__ m_list, int m_begin, int lmax, const double2* __restrict__ ad, const double2* __restrict__ wg,
                                const double* __restrict__ gam, const double2* __restrict__ a0, double* __restrict__ rec)
{
    const int m = m_list ? m_list[blockIdx.y] : (m_begin + (int)blockIdx.y);
    const long long colbase = alm_index(lmax, 0, m);
    const int K = ((lmax - m) >> 1) + 1;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) {
        const int l = m + 2 * k;
        const long long t = colbase + m + k;
        const double2 c = ad[t], w = wg[t];
        const double g = gam[t];
        const double2 al = a0[colbase + l];
        const double2 al1 = (l + 1 <= lmax) ? a0[colbase + l + 1] : make_double2(0.0, 0.0);
        const double2 al2 = (l + 2 <= lmax) ? a0[colbase + l + 2] : make_double2(0.0, 0.0);
        double2* r = reinterpret_cast<double2*>(rec + t * 6);
        r[0] = c;
        r[1] = make_double2(w.x * al.x + w.y * al2.x, m == 0 ? 0.0 : w.x * al.y + w.y * al2.y);   // a_l0 is real
        r[2] = make_double2(g * al1.x, m == 0 ? 0.0 : g * al1.y);
    }
}

// ---- one step of the h sequence for ring slot j at local step parity PAR (x2 = cos^2 theta in S.x) --------------------------
template <int R, int PAR>
__device__ __forceinline__ void rec_step_2s(RingState<0, R>& S, int j, double alpha, double beta)
{
    const double u = fma(alpha, S.x[j], beta);
    if (PAR == 0) S.pp[0][j] = fma(u, S.p[0][j], -S.pp[0][j]);
    else S.p[0][j] = fma(u, S.pp[0][j], -S.p[0][j]);
}

// load_rings of legendre.cuh with the column's own last pseudo degree, and x -> x^2
template <int R>
__device__ __forceinline__ void load_rings_2s(const LegParams& P, int m, int lmaxp, int pair0, int lane, RingState<0, R>& S, int& lmin, int& lmaxact, int& lstart)
{
    int mn = L_NEVER, mx = -1;
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const int pair = pair0 + j * 32 + lane;
        S.x[j] = 0.0; S.la[j] = L_NEVER;
        if (pair < P.npairs) {
            const int la = P.lact[(size_t)m * P.npairs + pair];
            if (la <= lmaxp) { const double x = P.x[pair]; S.la[j] = la; S.x[j] = x * x; mn = la < mn ? la : mn; mx = la > mx ? la : mx; }
        }
    }
    lmin = warp_min(mn); lmaxact = warp_max(mx);
    lstart = m + ((lmin - m) & ~1);
#pragma unroll
    for (int j = 0; j < R; ++j) {
        S.p[0][j] = 0.0; S.pp[0][j] = 0.0;
        if (S.la[j] != L_NEVER) {
            const size_t k = (size_t)m * P.npairs + (pair0 + j * 32 + lane);
            const bool odd = ((S.la[j] - lstart) & 1) != 0;
            const double2 v = reinterpret_cast<const double2*>(P.st)[k];
            S.p[0][j] = odd ? v.y : v.x; S.pp[0][j] = odd ? v.x : v.y;
        }
    }
}

// =============================================================================================================
// synthesis
// =============================================================================================================
template <int R, bool MIXED, int PAR>
__device__ __forceinline__ void synth_step_2s(RingState<0, R>& S, double (&acc)[4][R], const double* rec, int l)
{
    const double2 c = *reinterpret_cast<const double2*>(rec);
    const double2 ge = *reinterpret_cast<const double2*>(rec + 2);
    const double2 go = *reinterpret_cast<const double2*>(rec + 4);
#pragma unroll
    for (int j = 0; j < R; ++j) {
        if (!MIXED || l >= S.la[j]) {
            const double p0 = (PAR == 0) ? S.p[0][j] : S.pp[0][j];
            acc[0][j] = fma(p0, ge.x, acc[0][j]);
            acc[1][j] = fma(p0, ge.y, acc[1][j]);
            acc[2][j] = fma(p0, go.x, acc[2][j]);
            acc[3][j] = fma(p0, go.y, acc[3][j]);
            rec_step_2s<R, PAR>(S, j, c.x, c.y);
        }
    }
}

template <int R>
__global__ void __launch_bounds__(LEG_NT) leg_synth_2s(const LegParams P)
{
    constexpr int ND = 6, STEPS = 64;
    __shared__ __align__(16) double sbuf[2 * STEPS * ND];
    __shared__ __align__(8) unsigned long long sbar[2];
    const int lane = threadIdx.x;
    int row, chunk;
    leg_unit(P, row, chunk);
    const int m = P.m_list ? P.m_list[row] : (P.m_begin + row);
    const int lmaxp = twostep_lmax(P.lmax, m);
    const int pair0 = chunk * (32 * R);

    RingState<0, R> S;
    int lmin, lmaxact, lstart;
    load_rings_2s<R>(P, m, lmaxp, pair0, lane, S, lmin, lmaxact, lstart);
    double acc[4][R];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int j = 0; j < R; ++j) acc[a][j] = 0.0;

    if (lmin <= lmaxp) {
        const int nl = lmaxp - lstart + 1;
        const int nmixed = (lmaxact - lstart + 1) & ~1;
        if (lane == 0) { mbar_init(&sbar[0], 1); mbar_init(&sbar[1], 1); mbar_init_fence(); }
        __syncwarp();
        RecStream<ND, STEPS> rs;
        rs.src = P.rec + (size_t)(alm_index(P.lmax, 0, m) + lstart) * ND; rs.nrec = nl; rs.buf = sbuf; rs.bar = sbar;
        const int nchunk = (nl + STEPS - 1) / STEPS;
        rs.issue(0, lane);
        for (int c = 0; c < nchunk; ++c) {
            if (c + 1 < nchunk) rs.issue(c + 1, lane);
            const double* rec = rs.wait(c);
            const int t0 = c * STEPS;
            int cnt = nl - t0; if (cnt > STEPS) cnt = STEPS;
            int na = nmixed - t0; if (na > cnt) na = cnt;
            int i = 0;
#pragma unroll 1
            for (; i + 2 <= na; i += 2) {
                synth_step_2s<R, true, 0>(S, acc, rec + (size_t)i * ND, lstart + t0 + i);
                synth_step_2s<R, true, 1>(S, acc, rec + (size_t)(i + 1) * ND, lstart + t0 + i + 1);
            }
            if (i < na) { synth_step_2s<R, true, 0>(S, acc, rec + (size_t)i * ND, lstart + t0 + i); ++i; }
#pragma unroll 2
            for (; i + 2 <= cnt; i += 2) {
                synth_step_2s<R, false, 0>(S, acc, rec + (size_t)i * ND, 0);
                synth_step_2s<R, false, 1>(S, acc, rec + (size_t)(i + 1) * ND, 0);
            }
            if (i < cnt) synth_step_2s<R, false, 0>(S, acc, rec + (size_t)i * ND, 0);
            __syncwarp();
        }
    }

    // ---- write phase: north = even + x odd, south = even - x odd (zeros for pruned / never-activated rings) ----
    const int col = P.col_is_row ? row : m;
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const int pair = pair0 + j * 32 + lane;
        if (pair >= P.npairs) continue;
        const int rN = P.ringN[pair], rS = P.ringS[pair];
        const double x = P.x[pair];
        const double er = acc[0][j], ei = acc[1][j], orr = x * acc[2][j], oi = x * acc[3][j];
        if (rN >= 0) *phase_row(P, rN, col) = make_double2(er + orr, ei + oi);
        if (rS >= 0) *phase_row(P, rS, col) = make_double2(er - orr, ei - oi);
    }
}

// =============================================================================================================
// analysis
// =============================================================================================================
template <int R, bool MIXED, int PAR>
__device__ __forceinline__ void anal_step_2s(RingState<0, R>& S, const double (&X)[4][R], const double* rec, double (&part)[4], int l)
{
    const double2 c = *reinterpret_cast<const double2*>(rec);
#pragma unroll
    for (int j = 0; j < R; ++j) {
        if (!MIXED || l >= S.la[j]) {
            const double p0 = (PAR == 0) ? S.p[0][j] : S.pp[0][j];
            part[0] = fma(p0, X[0][j], part[0]);
            part[1] = fma(p0, X[1][j], part[1]);
            part[2] = fma(p0, X[2][j], part[2]);
            part[3] = fma(p0, X[3][j], part[3]);
            rec_step_2s<R, PAR>(S, j, c.x, c.y);
        }
    }
}

template <int R>
__global__ void __launch_bounds__(LEG_NT, 12) leg_anal_2s(const LegParams P)
{
    constexpr int G = 16, NV = 2, STEPS = 32;   // rows (value v = even | odd, step): one row per lane, as in leg_anal<2, R>
    __shared__ __align__(16) double sbuf[2 * STEPS * 2];
    __shared__ __align__(16) double2 red[NV * G * 33];
    __shared__ __align__(8) unsigned long long sbar[2];
    const int lane = threadIdx.x;
    int row, chunk;
    leg_unit(P, row, chunk);
    const int m = P.m_list ? P.m_list[row] : (P.m_begin + row);
    const int lmaxp = twostep_lmax(P.lmax, m);
    const int pair0 = chunk * (32 * R);

    RingState<0, R> S;
    int lmin, lmaxact, lstart;
    load_rings_2s<R>(P, m, lmaxp, pair0, lane, S, lmin, lmaxact, lstart);
    if (lmin > lmaxp) return;   // nothing to add (outputs are pre-zeroed)

    // folded inputs: even part X_N + X_S, odd part x (X_N - X_S)
    double X[4][R];
    const int col = P.col_is_row ? row : m;
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const int pair = pair0 + j * 32 + lane;
        double2 qN = make_double2(0.0, 0.0), qS = qN;
        double x = 0.0;
        if (S.la[j] != L_NEVER) {
            const int rN = P.ringN[pair], rS = P.ringS[pair];
            x = P.x[pair];
            if (rN >= 0) qN = *phase_row(P, rN, col);
            if (rS >= 0) qS = *phase_row(P, rS, col);
        }
        X[0][j] = qN.x + qS.x; X[1][j] = qN.y + qS.y;
        X[2][j] = x * (qN.x - qS.x); X[3][j] = x * (qN.y - qS.y);
    }

    const int nl = lmaxp - lstart + 1;
    const int nmixed = (lmaxact - lstart + 1) & ~1;
    if (lane == 0) { mbar_init(&sbar[0], 1); mbar_init(&sbar[1], 1); mbar_init_fence(); }
    __syncwarp();
    const long long abase = alm_index(P.lmax, 0, m);
    RecStream<2, STEPS> rs;
    rs.src = reinterpret_cast<const double*>(P.ad + abase + lstart); rs.nrec = nl; rs.buf = sbuf; rs.bar = sbar;
    const int nchunk = (nl + STEPS - 1) / STEPS;
    const bool odd_row = lane >= G;
    const double2* wg = reinterpret_cast<const double2*>(P.rec);   // (gamma_k c_{l_k+1}, gamma_k c_{l_k+2}); the analysis has no records
    rs.issue(0, lane);
    for (int c = 0; c < nchunk; ++c) {
        if (c + 1 < nchunk) rs.issue(c + 1, lane);
        const double* rec = rs.wait(c);
        int cnt = nl - c * STEPS; if (cnt > STEPS) cnt = STEPS;
        for (int g0 = 0; g0 < cnt; g0 += G) {
            int gcnt = cnt - g0; if (gcnt > G) gcnt = G;
            const int t0 = c * STEPS + g0;
            // this lane's row of the group: step s = lane % G of the even (lanes < G) or odd sums; its weights are fetched ahead
            const bool mine = (lane % G) < gcnt;
            const int pl = lstart + t0 + (lane % G);            // pseudo degree m + k
            const long long tk = abase + pl;
            const int l = m + 2 * (pl - m);                     // l_k
            double2 w = make_double2(0.0, 0.0);
            if (mine) { if (odd_row) w.x = P.gamma[tk]; else w = wg[tk]; }
            if (t0 >= nmixed && gcnt == G) {
#pragma unroll
                for (int s = 0; s < G; s += 2) {
                    const double* r0 = rec + (size_t)(g0 + s) * 2;
                    double part0[4], part1[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) { part0[k] = 0.0; part1[k] = 0.0; }
                    anal_step_2s<R, false, 0>(S, X, r0, part0, 0);
                    anal_step_2s<R, false, 1>(S, X, r0 + 2, part1, 0);
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        red[(v * G + s) * 33 + lane] = make_double2(part0[2 * v], part0[2 * v + 1]);
                        red[(v * G + s + 1) * 33 + lane] = make_double2(part1[2 * v], part1[2 * v + 1]);
                    }
                }
            } else {
#pragma unroll 1
                for (int s = 0; s < gcnt; s += 2) {
                    const double* r0 = rec + (size_t)(g0 + s) * 2;
                    double part0[4], part1[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) { part0[k] = 0.0; part1[k] = 0.0; }
                    anal_step_2s<R, true, 0>(S, X, r0, part0, lstart + t0 + s);
                    if (s + 1 < gcnt) anal_step_2s<R, true, 1>(S, X, r0 + 2, part1, lstart + t0 + s + 1);
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        red[(v * G + s) * 33 + lane] = make_double2(part0[2 * v], part0[2 * v + 1]);
                        red[(v * G + s + 1) * 33 + lane] = make_double2(part1[2 * v], part1[2 * v + 1]);
                    }
                }
            }
            __syncwarp();
            {
                const double2 t = red_sum<NV, G>(red, lane);
                if (mine) {
                    double2* out = P.alm_out0 + abase;
                    if (odd_row) {
                        // So_k -> a_{l_k+1}
                        if (l + 1 <= P.lmax) {
                            atomicAdd(&out[l + 1].x, w.x * t.x);
                            if (m != 0) atomicAdd(&out[l + 1].y, w.x * t.y);
                        }
                    } else {
                        // Se_k -> a_{l_k} and a_{l_k+2}
                        atomicAdd(&out[l].x, w.x * t.x);
                        if (m != 0) atomicAdd(&out[l].y, w.x * t.y);
                        if (l + 2 <= P.lmax) {
                            atomicAdd(&out[l + 2].x, w.y * t.x);
                            if (m != 0) atomicAdd(&out[l + 2].y, w.y * t.y);
                        }
                    }
                }
                __syncwarp();
            }
        }
        __syncwarp();
    }
}

}  // namespace pixsht
