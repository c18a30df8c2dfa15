// legendre_batch.cuh -- Legendre stage for a BATCH of spin-0 transforms that share one geometry (simulation sweeps: the
// "64-sim batched 4' lmax=2700" case of BASELINE.json configs[4]).  The recurrence of a (m, ring pair) does not depend on
// the data, so NB maps ride on one recurrence: per (l, m, ring pair) 2 + 2 NB FP64 FMA-class ops instead of 4 NB, and every
// accumulate FMA takes the function value from the operand reuse cache.  Same work decomposition, record stream and
// activation table as legendre.cuh (spin 0); the batch index plays the role of the component in the phase rows.
#pragma once
#include "legendre.cuh"

namespace pixsht {

constexpr int LEG_MAXBATCH = 4;
template <int NB> struct BatchRec { static constexpr int ND = 2 + 2 * NB; static constexpr int STEPS = 64; };   // { alpha, 0, (gamma a_b).re, .im ... }
template <int NB> struct BatchRed { static constexpr int G = 32 / NB; };   // l-steps per reduction group: NB double2 values per l, one (map, step) row per lane

struct BatchPtrs { const double2* in[LEG_MAXBATCH]; double2* out[LEG_MAXBATCH]; };

// alm of the NB maps -> records, index range [first, first+count)
template <int NB>
__global__ void k_prep_synth_b(long long first, long long count, int lmax, const double2* __restrict__ ad, const double* __restrict__ gamma,
                               const BatchPtrs A, double* __restrict__ rec)
{
    long long k = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x, end = first + count;
    for (; k < end; k += step) {
        const double g = gamma[k];
        double2* r = reinterpret_cast<double2*>(rec + k * BatchRec<NB>::ND);
        r[0] = make_double2(ad[k].x, 0.0);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const double2 a = A.in[b][k];
            r[1 + b] = make_double2(g * a.x, (k <= lmax) ? 0.0 : g * a.y);   // k <= lmax  <=>  m == 0: a_l0 is real
        }
    }
}

template <int R, int NB, bool MIXED, int PAR>
__device__ __forceinline__ void synth_step_b(RingState<0, R>& S, double (&acc)[4 * NB][R], const double* rec, int l)
{
    const double alpha = rec[0];
    double2 g[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) g[b] = *reinterpret_cast<const double2*>(rec + 2 + 2 * b);
#pragma unroll
    for (int j = 0; j < R; ++j) {
        if (!MIXED || l >= S.la[j]) {
            const double p0 = (PAR == 0) ? S.p[0][j] : S.pp[0][j];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                acc[4 * b + 2 * PAR + 0][j] = fma(p0, g[b].x, acc[4 * b + 2 * PAR + 0][j]);
                acc[4 * b + 2 * PAR + 1][j] = fma(p0, g[b].y, acc[4 * b + 2 * PAR + 1][j]);
            }
            rec_step<0, R, PAR>(S, j, alpha, 0.0);
        }
    }
}

template <int R, int NB>
__global__ void __launch_bounds__(LEG_NT) leg_synth_b(const LegParams P)
{
    constexpr int ND = BatchRec<NB>::ND, STEPS = BatchRec<NB>::STEPS;
    __shared__ __align__(16) double sbuf[2 * STEPS * ND];
    __shared__ __align__(8) unsigned long long sbar[2];
    const int lane = threadIdx.x;
    int row, chunk;
    leg_unit(P, row, chunk);
    const int m = P.m_list ? P.m_list[row] : (P.m_begin + row);
    const int l0 = m;
    const int pair0 = chunk * (32 * R);

    RingState<0, R> S;
    int lmin, lmaxact, lstart;
    load_rings<0, R>(P, m, l0, pair0, lane, S, lmin, lmaxact, lstart);
    double acc[4 * NB][R];
#pragma unroll
    for (int a = 0; a < 4 * NB; ++a)
#pragma unroll
        for (int j = 0; j < R; ++j) acc[a][j] = 0.0;

    if (lmin <= P.lmax) {
        const int nl = P.lmax - lstart + 1;
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
                synth_step_b<R, NB, true, 0>(S, acc, rec + (size_t)i * ND, lstart + t0 + i);
                synth_step_b<R, NB, true, 1>(S, acc, rec + (size_t)(i + 1) * ND, lstart + t0 + i + 1);
            }
            if (i < na) { synth_step_b<R, NB, true, 0>(S, acc, rec + (size_t)i * ND, lstart + t0 + i); ++i; }
#pragma unroll 2
            for (; i + 2 <= cnt; i += 2) {
                synth_step_b<R, NB, false, 0>(S, acc, rec + (size_t)i * ND, 0);
                synth_step_b<R, NB, false, 1>(S, acc, rec + (size_t)(i + 1) * ND, 0);
            }
            if (i < cnt) synth_step_b<R, NB, false, 0>(S, acc, rec + (size_t)i * ND, 0);
            __syncwarp();
        }
    }
    const int col = P.col_is_row ? row : m;
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const int pair = pair0 + j * 32 + lane;
        if (pair >= P.npairs) continue;
        const int rN = P.ringN[pair], rS = P.ringS[pair];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const double er = acc[4 * b + 0][j], ei = acc[4 * b + 1][j], orr = acc[4 * b + 2][j], oi = acc[4 * b + 3][j];
            if (rN >= 0) phase_row(P, rN, col)[(long long)b * P.MP] = make_double2(er + orr, ei + oi);
            if (rS >= 0) phase_row(P, rS, col)[(long long)b * P.MP] = make_double2(er - orr, ei - oi);
        }
    }
}

template <int R, int NB, bool MIXED, int PAR>
__device__ __forceinline__ void anal_step_b(RingState<0, R>& S, const double (&X)[4 * NB][R], const double* rec, double (&part)[2 * NB], int l)
{
    const double alpha = rec[0];
#pragma unroll
    for (int j = 0; j < R; ++j) {
        if (!MIXED || l >= S.la[j]) {
            const double p0 = (PAR == 0) ? S.p[0][j] : S.pp[0][j];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                part[2 * b + 0] = fma(p0, X[4 * b + 2 * PAR + 0][j], part[2 * b + 0]);
                part[2 * b + 1] = fma(p0, X[4 * b + 2 * PAR + 1][j], part[2 * b + 1]);
            }
            rec_step<0, R, PAR>(S, j, alpha, 0.0);
        }
    }
}

template <int R, int NB>
__global__ void __launch_bounds__(LEG_NT) leg_anal_b(const LegParams P, const BatchPtrs A)
{
    constexpr int NPART = 2 * NB, G = BatchRed<NB>::G, NV = NB, STEPS = ANAL_STEPS;
    __shared__ __align__(16) double sbuf[2 * STEPS * 2];
    __shared__ __align__(16) double2 red[NV * G * 33];
    __shared__ __align__(8) unsigned long long sbar[2];
    const int lane = threadIdx.x;
    int row, chunk;
    leg_unit(P, row, chunk);
    const int m = P.m_list ? P.m_list[row] : (P.m_begin + row);
    const int l0 = m;
    const int pair0 = chunk * (32 * R);

    RingState<0, R> S;
    int lmin, lmaxact, lstart;
    load_rings<0, R>(P, m, l0, pair0, lane, S, lmin, lmaxact, lstart);
    if (lmin > P.lmax) return;

    double X[4 * NB][R];
    const int col = P.col_is_row ? row : m;
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const int pair = pair0 + j * 32 + lane;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            double2 qN = make_double2(0.0, 0.0), qS = qN;
            if (S.la[j] != L_NEVER) {
                const int rN = P.ringN[pair], rS = P.ringS[pair];
                if (rN >= 0) qN = phase_row(P, rN, col)[(long long)b * P.MP];
                if (rS >= 0) qS = phase_row(P, rS, col)[(long long)b * P.MP];
            }
            X[4 * b + 0][j] = qN.x + qS.x; X[4 * b + 1][j] = qN.y + qS.y;
            X[4 * b + 2][j] = qN.x - qS.x; X[4 * b + 3][j] = qN.y - qS.y;
        }
    }

    const int nl = P.lmax - lstart + 1;
    const int nmixed = (lmaxact - lstart + 1) & ~1;
    if (lane == 0) { mbar_init(&sbar[0], 1); mbar_init(&sbar[1], 1); mbar_init_fence(); }
    __syncwarp();
    const long long abase = alm_index(P.lmax, 0, m);
    double2* outb = A.out[0];   // the map this lane reduces into (lane = (map, step) in the reduction)
#pragma unroll
    for (int b = 1; b < NB; ++b) if (lane / G == b) outb = A.out[b];
    RecStream<2, STEPS> rs;
    rs.src = reinterpret_cast<const double*>(P.ad + abase + lstart); rs.nrec = nl; rs.buf = sbuf; rs.bar = sbar;
    const int nchunk = (nl + STEPS - 1) / STEPS;
    rs.issue(0, lane);
    for (int c = 0; c < nchunk; ++c) {
        if (c + 1 < nchunk) rs.issue(c + 1, lane);
        const double* rec = rs.wait(c);
        int cnt = nl - c * STEPS; if (cnt > STEPS) cnt = STEPS;
        for (int g0 = 0; g0 < cnt; g0 += G) {
            int gcnt = cnt - g0; if (gcnt > G) gcnt = G;
            const int t0 = c * STEPS + g0;
            const bool mine = (lane % G) < gcnt;
            const long long gk = abase + lstart + t0 + (lane % G);
            const double g = mine ? P.gamma[gk] : 0.0;
            if (t0 >= nmixed && gcnt == G) {
#pragma unroll
                for (int s = 0; s < G; s += 2) {
                    const double* r0 = rec + (size_t)(g0 + s) * 2;
                    double part0[NPART], part1[NPART];
#pragma unroll
                    for (int k = 0; k < NPART; ++k) { part0[k] = 0.0; part1[k] = 0.0; }
                    anal_step_b<R, NB, false, 0>(S, X, r0, part0, 0);
                    anal_step_b<R, NB, false, 1>(S, X, r0 + 2, part1, 0);
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
                    double part0[NPART], part1[NPART];
#pragma unroll
                    for (int k = 0; k < NPART; ++k) { part0[k] = 0.0; part1[k] = 0.0; }
                    anal_step_b<R, NB, true, 0>(S, X, r0, part0, lstart + t0 + s);
                    if (s + 1 < gcnt) anal_step_b<R, NB, true, 1>(S, X, r0 + 2, part1, lstart + t0 + s + 1);
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        red[(v * G + s) * 33 + lane] = make_double2(part0[2 * v], part0[2 * v + 1]);
                        red[(v * G + s + 1) * 33 + lane] = make_double2(part1[2 * v], part1[2 * v + 1]);
                    }
                }
            }
            __syncwarp();
            {
                const double2 t = red_sum<NV, G>(red, lane);   // lane = (map b = lane / G, step lane % G)
                if (mine) {
                    double2* out = outb + gk;
                    atomicAdd(&out->x, g * t.x);
                    if (m != 0) atomicAdd(&out->y, g * t.y);
                }
                __syncwarp();
            }
        }
        __syncwarp();
    }
}

}  // namespace pixsht
