// legendre.cuh -- the Legendre stage (alm <-> phase), hand-written for sm_100a.  This is where ~all the time goes
// ("libsharp2 sharp_execute internals, stage B").  FP64 FMA bound; nothing here is a tensor-core shape.
//
// Work decomposition: one CTA = one warp = one independent work unit = (one m) x (32*R contiguous north/south ring
// pairs); lane i owns pairs base + i + 32 j, j < R (coalesced phase I/O, neighbouring rings => the rings of a warp
// become active within a few per cent of the l range of each other).  There is no block-level barrier anywhere: ncu
// (profiles/r01) showed `barrier` as the top stall when four warps with different amounts of work shared a CTA.
//
// For each pair the scaled functions p_l = lambda_lm(theta)/gamma_l are generated on the fly by
//     p_{l+1} = (alpha_l x + delta_l) p_l - p_{l-1}                       (2 FP64 ops per l and function)
// The per-(l,m) uniform operands (alpha, delta and, for synthesis, the pre-scaled alm) are fixed-size records in
// global memory; each warp streams its m-column of records through a private two-stage shared-memory ring with TMA bulk
// copies (cp.async.bulk, completion on an mbarrier), so staging overlaps the arithmetic and costs no register.
//   synthesis: per-ring register accumulators  sum_l p_l * (gamma_l a_lm)   (2 FMA per l, spin 0; 8 for spin 2),
//              north = even + odd, south = even - odd  (equatorial symmetry);
//   analysis : per-l partial sums over the thread's R rings, a warp-private shared-memory transpose-reduction
//              every 16 steps (one (value, step) row per lane), then one atomicAdd per (l, m, warp).
//
// Dynamic range (DESIGN.md "activation table"): lambda_lm(theta) ~ sin^m(theta) underflows FP64 by thousands of
// decades near the poles.  The scaled-exponent "seek" from l = m up to the first l where the function reaches 2^-90
// depends only on the plan (m, theta), not on the data, so it runs ONCE at plan creation (k_seek_table): for every
// (m, ring pair) it stores l_act and the recurrence state (p, p_{l-1}) there, or L_NEVER when the pair never matters
// for l <= lmax.  The transform kernels start each ring at its l_act in plain FP64: no exponents, no rescale checks,
// no seeding arithmetic in the hot loops.  A warp runs "mixed" steps (per-ring predicate l >= l_act) from the earliest
// l_act of its rings to the latest, then branch-free full steps to lmax.
#pragma once
#include "common.cuh"

namespace pixsht {

// records streamed by the synthesis kernels (written per call by k_prep_synth):
//   spin 0: { alpha, 0, gamma*Re a, gamma*Im a }                                  4 doubles
//   spin 2: { alpha, delta, G+re, G+im, G-re, G-im },  G+- = -gamma (E +- iB)/2   6 doubles
// the analysis kernels stream the static (alpha, delta) table (2 doubles per l).
template <int SPIN> struct SynthRec { static constexpr int ND = (SPIN == 0) ? 4 : 6; static constexpr int STEPS = (SPIN == 0) ? 128 : 64; };
// Analysis: every l-step leaves NV double2 partial sums per lane; groups of G steps are summed over the 32 lanes through a
// warp-private shared-memory transpose (row = (value, step), column = source lane, rows padded to 33 so that row-wise stores
// and column-wise loads are conflict-free).  NV*G rows are shared out over the 32 lanes: LPP lanes per row, one shuffle level
// when LPP == 2.  STEPS = records per TMA stage; spin 2 keeps it small so that 12 one-warp CTAs fit the shared memory of an SM.
constexpr int ANAL_STEPS = 128;
template <int SPIN> struct RedT {
    static constexpr int G = 16, NV = (SPIN == 0) ? 1 : 2;
    static constexpr int STEPS = (SPIN == 0) ? ANAL_STEPS : 32;
};
template <int NV, int G> struct RedMap {
    static constexpr int ROWS = NV * G, LPP = 32 / ROWS, SL = 32 / LPP;
    static_assert(ROWS * LPP == 32 && (LPP == 1 || LPP == 2), "rows must tile the warp");
};
constexpr int L_NEVER = 0x3fffffff;

struct LegParams {
    int lmax, mmax;
    int nm;                 // number of m values handled by this launch
    const int* m_list;      // device; nullptr => m = m_begin + row index
    int m_begin;
    int npairs, nchunks;    // ring pairs; this launch handles nchunks chunks of 32*R pairs per m, starting at chunk_begin
    int chunk_begin;
    const double* x;        // [npairs] cos(theta) of the pair's northern member
    const int* ringN; const int* ringS;           // band ring index of the north/south member, -1 if absent
    const int* lact;        // [(mmax+1) * npairs] first l at which the pair contributes, L_NEVER if none
    const double* st;       // [(mmax+1) * npairs * NS] state at l_act: spin 0 (p, p_prev); spin 2 (p+, p-, p+_prev, p-_prev)
    const double2* ad;      // [nalm] (alpha, delta) table of this spin family
    const double* gamma;    // [nalm]
    const double* rec;      // synthesis: [nalm * SynthRec::ND] records
    double2* alm_out0; double2* alm_out1;               // analysis output  (T | E,B), pre-zeroed, accumulated atomically
    // phase element (ring, component c, column) lives at phase + ring*ring_stride + c*MP + column.  The column of an m is m
    // itself in the single-GPU layout (rows hold all m) and the launch row (position in m_list) in the m-sharded layout,
    // where a rank's buffer holds only its own m values for all rings.  c0 = first component handled by this launch.
    double2* phase; long long ring_stride;
    long long MP; int c0; int col_is_row;
    int order;              // grid order of the work units, see leg_unit
};

// work unit of this CTA: (launch row = position of m in the launch, chunk of 32 R ring pairs).  The grid runs over the chunks from the
// equator (longest units) to the pole in groups of `order` chunk positions; inside a group the units of one m are neighbours.
// order >= nchunks (or 0): m-major -- the chunks of one m run together, stream the same coefficient column through L2 and meet on the
// same alm lines with their atomics; order 1: chunk-major -- neighbouring CTAs work on different m at the same latitude (equal
// length, longest first: measured 3 % faster at lmax 10800, at the price of re-reading the coefficient columns from HBM).
__device__ __forceinline__ void leg_unit(const LegParams& P, int& row, int& chunk)
{
    const int g = (P.order <= 0 || P.order > P.nchunks) ? P.nchunks : P.order;
    const int per = P.nm * g;                       // units per full group
    const int cg = (int)(blockIdx.x / (unsigned)per);
    const int b = (int)blockIdx.x - cg * per;
    const int left = P.nchunks - cg * g;            // chunk positions in this group (the last group may be short)
    const int gg = left < g ? left : g;
    row = b / gg;
    const int pos = cg * g + (b - row * gg);
    chunk = P.chunk_begin + P.nchunks - 1 - pos;
}

__device__ __forceinline__ double2* phase_row(const LegParams& P, int ring, int col)
{
    return P.phase + (long long)ring * P.ring_stride + (long long)P.c0 * P.MP + col;
}

// plan-time inputs of the activation table
struct SeekParams {
    int lmax, mmax, npairs;
    const double* x;
    const double* lsh_hi; const double* lsh_lo;   // log2 sin(theta/2), double-double
    const double* lch_hi; const double* lch_lo;   // log2 cos(theta/2)
    const double* mlim;     // [npairs] prune: the pair is skipped for m > mlim
    const double* lgpref_hi; const double* lgpref_lo;   // [mmax+1] log2 of the seed prefactor for this spin family
    const double2* ad;
    int* lact; double* st;
    int thr_log2;           // rescale threshold 2^thr_log2 of the seek (SEEK_THR_LOG2 unless the plan overrides it): a pair becomes
                            // active once its function has grown to 2^(thr_log2 - 64)
};

// ---- pre-scaling pass: alm -> records (element-wise; ~1 ms at lmax = 10800) ----------------------------------------------
template <int SPIN>
__device__ __forceinline__ void prep_record(long long k, bool m0, const double2* __restrict__ ad, const double* __restrict__ gamma,
                                            const double2* __restrict__ a0, const double2* __restrict__ a1, double* __restrict__ rec)
{
    const double2 c = ad[k];
    const double g = gamma[k];
    double2* r = reinterpret_cast<double2*>(rec + k * SynthRec<SPIN>::ND);
    if (SPIN == 0) {
        const double2 a = a0[k];
        r[0] = make_double2(c.x, 0.0);
        r[1] = make_double2(g * a.x, m0 ? 0.0 : g * a.y);   // a_l0 is real
    } else {
        const double2 E = a0[k], B = a1[k];
        const double h = -0.5 * g;
        r[0] = c;
        r[1] = make_double2(h * (E.x - B.y), h * (E.y + B.x));
        r[SPIN != 0 ? 2 : 0] = make_double2(h * (E.x + B.y), h * (E.y - B.x));
    }
}
// over the alm index range [first, first+count)
template <int SPIN>
__global__ void k_prep_synth(long long first, long long count, int lmax, const double2* __restrict__ ad, const double* __restrict__ gamma,
                             const double2* __restrict__ a0, const double2* __restrict__ a1, double* __restrict__ rec)
{
    long long k = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x, end = first + count;
    for (; k < end; k += step) prep_record<SPIN>(k, k <= lmax, ad, gamma, a0, a1, rec);   // k <= lmax  <=>  m == 0
}
// over the alm columns of the m values m_list[0..nm) (blockIdx.y = position in the list): the m-sharded pipelines prepare
// only the columns a launch owns (and that have arrived)
template <int SPIN>
__global__ void k_prep_synth_rows(const int* __restrict__ m_list, int lmax, const double2* __restrict__ ad, const double* __restrict__ gamma,
                                  const double2* __restrict__ a0, const double2* __restrict__ a1, double* __restrict__ rec)
{
    const int m = m_list[blockIdx.y];
    const long long base = alm_index(lmax, 0, m);
    for (int l = m + blockIdx.x * blockDim.x + threadIdx.x; l <= lmax; l += gridDim.x * blockDim.x)
        prep_record<SPIN>(base + l, m == 0, ad, gamma, a0, a1, rec);
}

// ---- seeds (plan time only) --------------------------------------------------------------------------------
// value = sign * 2^(lg) with lg = lgpref[m] + a*log2 cos(theta/2) + b*log2 sin(theta/2), returned as (k, frac) with
// lg = k + frac, k integer-valued.  zero => the seed vanishes (pole).
struct LogVal { double k, frac; bool zero; };

__device__ __forceinline__ LogVal seed_log(const SeekParams& P, int m, int pair, int a, int b)
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
__device__ __forceinline__ int seed_exponent(double k, int act_log2)
{
    if (k >= (double)act_log2) return 0;
    const double q = ceil(((double)act_log2 - k) / (double)SEEK_QUANT);
    return (int)(-(double)SEEK_QUANT * q);
}

__device__ __forceinline__ double seed_value(const LogVal& v, int e, double sign)
{
    if (v.zero) return 0.0;
    const double d = v.k - (double)e;
    if (d < -1000.0) return 0.0;
    return sign * ldexp(exp2(v.frac), (int)d);
}

__device__ __forceinline__ bool over_thr(double v, unsigned expbits)
{
    return ((unsigned)__double2hiint(v) & 0x7ff00000u) >= expbits;
}

// One thread per (m, ring pair): seed at l0 = max(m, |s|) as mantissa * 2^e (e a multiple of 64, <= 0), run the
// recurrence with rescaling until e reaches 0 (the function has grown to >= 2^-90) and record l and the state there.
template <int SPIN>
__global__ void __launch_bounds__(128) k_seek_table(const SeekParams P)
{
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y;
    if (pair >= P.npairs) return;
    const int l0 = (SPIN == 0) ? m : (m > 2 ? m : 2);
    const int act_log2 = P.thr_log2 - SEEK_QUANT;
    const unsigned expbits = (unsigned)(1023 + P.thr_log2) << 20;
    double p0 = 0.0, p1 = 0.0, q0 = 0.0, q1 = 0.0;   // p: value at l, q: value at l-1  (0: lambda / lambda+, 1: lambda-)
    int e = E_DEAD;
    if ((double)m <= P.mlim[pair] && l0 <= P.lmax) {
        const double sgn = (m & 1) ? -1.0 : 1.0;
        if (SPIN == 0) {
            // lambda_mm = (-1)^m N_m sin^m(theta), sin(theta) = 2 sin(theta/2) cos(theta/2): the factor 2^m is in lgpref
            LogVal v = seed_log(P, m, pair, m, m);
            if (!v.zero) { e = seed_exponent(v.k, act_log2); p0 = seed_value(v, e, sgn); }
        } else {
            // l0 = max(m,2):  lambda^+ ~ cos^{|m-2|} sin^{m+2},  lambda^- ~ cos^{m+2} sin^{|m-2|}  (half angles)
            const int am = m >= 2 ? m - 2 : 2 - m;
            LogVal vp = seed_log(P, m, pair, am, m + 2);
            LogVal vm = seed_log(P, m, pair, m + 2, am);
            const double sp = sgn, sm = (m >= 2) ? sgn : 1.0;
            if (!(vp.zero && vm.zero)) {
                const double kmax = vp.zero ? vm.k : (vm.zero ? vp.k : fmax(vp.k, vm.k));
                e = seed_exponent(kmax, act_log2);
                p0 = seed_value(vp, e, sp);
                p1 = seed_value(vm, e, sm);
            }
        }
    }
    int l = l0;
    if (e < 0) {
        const double x = P.x[pair];
        const double2* ad = P.ad + alm_index(P.lmax, 0, m);
        const double sc = 5.421010862427522e-20;  // 2^-64
        while (e < 0 && l <= P.lmax) {
            const double2 c = ad[l];
            if (SPIN == 0) {
                const double pn = fma(c.x * x, p0, -q0);
                q0 = p0; p0 = pn;
                if (over_thr(pn, expbits)) { p0 *= sc; q0 *= sc; e += SEEK_QUANT; }
            } else {
                const double pn = fma(fma(c.x, x, c.y), p0, -q0);
                const double mn = fma(fma(c.x, x, -c.y), p1, -q1);
                q0 = p0; p0 = pn; q1 = p1; p1 = mn;
                if (over_thr(pn, expbits) || over_thr(mn, expbits)) { p0 *= sc; q0 *= sc; p1 *= sc; q1 *= sc; e += SEEK_QUANT; }
            }
            ++l;
        }
    }
    const size_t k = (size_t)m * P.npairs + pair;
    const bool live = (e == 0) && (l <= P.lmax);
    P.lact[k] = live ? l : L_NEVER;
    if (SPIN == 0) reinterpret_cast<double2*>(P.st)[k] = live ? make_double2(p0, q0) : make_double2(0.0, 0.0);
    else {
        double2* o = reinterpret_cast<double2*>(P.st) + 2 * k;
        o[0] = live ? make_double2(p0, p1) : make_double2(0.0, 0.0);
        o[SPIN != 0] = live ? make_double2(q0, q1) : make_double2(0.0, 0.0);
    }
}

// ---- per-thread ring state of the transform kernels ------------------------------------------------------------
// The two recurrence registers of a function alternate roles instead of being rotated: at an even local step (PAR 0)
// `p` holds the current value and `pp` the previous one, the step overwrites `pp` with the next value; at an odd local
// step (PAR 1) the roles are swapped.  No register moves, and a predicated step is a handful of predicated FMAs.
template <int SPIN, int R>
struct RingState {
    double x[R];
    double p[(SPIN == 0 ? 1 : 2)][R];
    double pp[(SPIN == 0 ? 1 : 2)][R];
    int la[R];
};

__device__ __forceinline__ int warp_min(int v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { const int o = __shfl_xor_sync(0xffffffffu, v, off); v = o < v ? o : v; }
    return v;
}
__device__ __forceinline__ int warp_max(int v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { const int o = __shfl_xor_sync(0xffffffffu, v, off); v = o > v ? o : v; }
    return v;
}

// loads the activation state of the warp's rings; lmin / lmaxact = earliest / latest l_act over the live rings of the
// warp; lstart = first l of the warp's stream (same parity as l0, <= lmin)
template <int SPIN, int R>
__device__ __forceinline__ void load_rings(const LegParams& P, int m, int l0, int pair0, int lane, RingState<SPIN, R>& S, int& lmin,
                                           int& lmaxact, int& lstart)
{
    int mn = L_NEVER, mx = -1;
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const int pair = pair0 + j * 32 + lane;
        S.x[j] = 0.0; S.la[j] = L_NEVER;
        if (pair < P.npairs) {
            const int la = P.lact[(size_t)m * P.npairs + pair];
            if (la <= P.lmax) { S.la[j] = la; S.x[j] = P.x[pair]; mn = la < mn ? la : mn; mx = la > mx ? la : mx; }
        }
    }
    lmin = warp_min(mn); lmaxact = warp_max(mx);
    lstart = l0 + ((lmin - l0) & ~1);
#pragma unroll
    for (int j = 0; j < R; ++j) {
        S.p[0][j] = 0.0; S.pp[0][j] = 0.0;
        if (SPIN != 0) { S.p[SPIN != 0][j] = 0.0; S.pp[SPIN != 0][j] = 0.0; }
        if (S.la[j] != L_NEVER) {
            const size_t k = (size_t)m * P.npairs + (pair0 + j * 32 + lane);
            const bool odd = ((S.la[j] - lstart) & 1) != 0;   // the ring's first step is an odd local step: roles swapped
            if (SPIN == 0) {
                const double2 v = reinterpret_cast<const double2*>(P.st)[k];
                S.p[0][j] = odd ? v.y : v.x; S.pp[0][j] = odd ? v.x : v.y;
            } else {
                const double2 v = reinterpret_cast<const double2*>(P.st)[2 * k], w = reinterpret_cast<const double2*>(P.st)[2 * k + 1];
                S.p[0][j] = odd ? w.x : v.x; S.pp[0][j] = odd ? v.x : w.x;
                S.p[SPIN != 0][j] = odd ? w.y : v.y; S.pp[SPIN != 0][j] = odd ? v.y : w.y;
            }
        }
    }
}

// one recurrence step for ring slot j at local step parity PAR
template <int SPIN, int R, int PAR>
__device__ __forceinline__ void rec_step(RingState<SPIN, R>& S, int j, double alpha, double delta)
{
    if (SPIN == 0) {
        // alpha * (x * p) - p_prev rather than (alpha * x) * p - p_prev: the FMA then has the warp-uniform alpha (served by
        // the operand reuse cache) and only two fresh vector-register operands.  A DFMA with three fresh register operands
        // issues at half rate on sm_100 (tools/dfma_mix.cu, DESIGN.md "operand bandwidth").
        if (PAR == 0) { const double t = S.x[j] * S.p[0][j]; S.pp[0][j] = fma(alpha, t, -S.pp[0][j]); }
        else { const double t = S.x[j] * S.pp[0][j]; S.p[0][j] = fma(alpha, t, -S.p[0][j]); }
    } else {
        const double up = fma(alpha, S.x[j], delta);
        const double um = fma(alpha, S.x[j], -delta);
        if (PAR == 0) {
            S.pp[0][j] = fma(up, S.p[0][j], -S.pp[0][j]);
            S.pp[SPIN != 0][j] = fma(um, S.p[SPIN != 0][j], -S.pp[SPIN != 0][j]);
        } else {
            S.p[0][j] = fma(up, S.pp[0][j], -S.p[0][j]);
            S.p[SPIN != 0][j] = fma(um, S.pp[SPIN != 0][j], -S.p[SPIN != 0][j]);
        }
    }
}

// ---- warp-private record stream: two-stage shared ring filled by TMA bulk copies ----------------------------------
template <int ND, int STEPS>
struct RecStream {
    const double* src;      // first record of this warp's column
    int nrec;               // records in the column
    double* buf;            // [2][STEPS*ND] in shared memory
    unsigned long long* bar;  // [2]
    __device__ __forceinline__ void issue(int chunk, int lane) const
    {
        if (lane == 0) {
            const int first = chunk * STEPS;
            int n = nrec - first; if (n > STEPS) n = STEPS;
            bulk_g2s(buf + (size_t)(chunk & 1) * STEPS * ND, src + (size_t)first * ND, (unsigned)(n * ND * sizeof(double)), &bar[chunk & 1]);
        }
    }
    __device__ __forceinline__ const double* wait(int chunk) const
    {
        mbar_wait(&bar[chunk & 1], (unsigned)((chunk >> 1) & 1));
        return buf + (size_t)(chunk & 1) * STEPS * ND;
    }
};

// =============================================================================================================
// synthesis: alm -> phase
// =============================================================================================================
// MIXED: per-ring predicate l >= l_act (some rings of the warp have not started yet); otherwise branch-free
template <int SPIN, int R, bool MIXED, int PAR>
__device__ __forceinline__ void synth_step(RingState<SPIN, R>& S, double (&acc)[(SPIN == 0 ? 4 : 8)][R], const double* rec, int l)
{
    const double2 c = *reinterpret_cast<const double2*>(rec);
    const double2 g0 = *reinterpret_cast<const double2*>(rec + 2);
    const double2 g1 = *reinterpret_cast<const double2*>(rec + (SPIN == 0 ? 2 : 4));
#pragma unroll
    for (int j = 0; j < R; ++j) {
        if (!MIXED || l >= S.la[j]) {
            const double p0 = (PAR == 0) ? S.p[0][j] : S.pp[0][j];
            const double p1 = (PAR == 0) ? S.p[SPIN != 0][j] : S.pp[SPIN != 0][j];
            if (SPIN == 0) {
                acc[2 * PAR + 0][j] = fma(p0, g0.x, acc[2 * PAR + 0][j]);
                acc[2 * PAR + 1][j] = fma(p0, g0.y, acc[2 * PAR + 1][j]);
            } else {
                // north: S+ += p+ G+, S- += p- G-;  south: T+ += sgn p- G+, T- += sgn p+ G-  (sgn alternates with l)
                constexpr int A4 = (SPIN == 0 ? 0 : 4);   // keeps indices in range in the (dead) SPIN == 0 instantiation
                const double q0 = (PAR == 0) ? p0 : -p0, q1 = (PAR == 0) ? p1 : -p1;
                acc[0][j] = fma(p0, g0.x, acc[0][j]);
                acc[1][j] = fma(p0, g0.y, acc[1][j]);
                acc[A4 + 2][j] = fma(q0, g1.x, acc[A4 + 2][j]);
                acc[A4 + 3][j] = fma(q0, g1.y, acc[A4 + 3][j]);
                acc[2][j] = fma(p1, g1.x, acc[2][j]);
                acc[3][j] = fma(p1, g1.y, acc[3][j]);
                acc[A4 + 0][j] = fma(q1, g0.x, acc[A4 + 0][j]);
                acc[A4 + 1][j] = fma(q1, g0.y, acc[A4 + 1][j]);
            }
            rec_step<SPIN, R, PAR>(S, j, c.x, c.y);
        }
    }
}

template <int SPIN, int R>
__global__ void __launch_bounds__(LEG_NT) leg_synth(const LegParams P)
{
    constexpr int NACC = (SPIN == 0) ? 4 : 8;
    constexpr int ND = SynthRec<SPIN>::ND, STEPS = SynthRec<SPIN>::STEPS;
    __shared__ __align__(16) double sbuf[2 * STEPS * ND];
    __shared__ __align__(8) unsigned long long sbar[2];
    const int lane = threadIdx.x;
    int row, chunk;
    leg_unit(P, row, chunk);
    const int m = P.m_list ? P.m_list[row] : (P.m_begin + row);
    const int l0 = (SPIN == 0) ? m : (m > 2 ? m : 2);
    const int pair0 = chunk * (32 * R);

    RingState<SPIN, R> S;
    int lmin, lmaxact, lstart;
    load_rings<SPIN, R>(P, m, l0, pair0, lane, S, lmin, lmaxact, lstart);
    double acc[NACC][R];
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
        for (int j = 0; j < R; ++j) acc[a][j] = 0.0;

    if (lmin <= P.lmax) {
        // local step t <-> l = lstart + t; lstart keeps the parity of l0 so that step parity = parity of (l - l0)
        const int nl = P.lmax - lstart + 1;
        const int nmixed = (lmaxact - lstart + 1) & ~1;   // steps [0, nmixed) need the per-ring predicate
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
            // STEPS and nmixed are even, so local step i has the parity of (l - l0)
#pragma unroll 1
            for (; i + 2 <= na; i += 2) {
                synth_step<SPIN, R, true, 0>(S, acc, rec + (size_t)i * ND, lstart + t0 + i);
                synth_step<SPIN, R, true, 1>(S, acc, rec + (size_t)(i + 1) * ND, lstart + t0 + i + 1);
            }
            if (i < na) { synth_step<SPIN, R, true, 0>(S, acc, rec + (size_t)i * ND, lstart + t0 + i); ++i; }   // odd tail at lmax
#pragma unroll 2
            for (; i + 2 <= cnt; i += 2) {
                synth_step<SPIN, R, false, 0>(S, acc, rec + (size_t)i * ND, 0);
                synth_step<SPIN, R, false, 1>(S, acc, rec + (size_t)(i + 1) * ND, 0);
            }
            if (i < cnt) synth_step<SPIN, R, false, 0>(S, acc, rec + (size_t)i * ND, 0);   // odd tail: the very last l of the column
            __syncwarp();   // every lane is done with this stage before it is refilled
        }
    }

    // ---- write phase (zeros for pruned / never-activated rings) ----
    const int col = P.col_is_row ? row : m;
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const int pair = pair0 + j * 32 + lane;
        if (pair >= P.npairs) continue;
        const int rN = P.ringN[pair], rS = P.ringS[pair];
        if (SPIN == 0) {
            const double er = acc[0][j], ei = acc[1][j], orr = acc[2][j], oi = acc[3][j];
            if (rN >= 0) *phase_row(P, rN, col) = make_double2(er + orr, ei + oi);
            if (rS >= 0) *phase_row(P, rS, col) = make_double2(er - orr, ei - oi);
        } else {
            // q = S+ + S-, u = -i (S+ - S-);  south: base sign (-1)^(l0+m) times the alternating sums
            const double bs = ((l0 + m) & 1) ? -1.0 : 1.0;
            constexpr int A4 = (SPIN == 0 ? 0 : 4);
            if (rN >= 0) {
                double2* ph = phase_row(P, rN, col);
                ph[0] = make_double2(acc[0][j] + acc[2][j], acc[1][j] + acc[3][j]);
                ph[P.MP] = make_double2(acc[1][j] - acc[3][j], -(acc[0][j] - acc[2][j]));
            }
            if (rS >= 0) {
                double2* ph = phase_row(P, rS, col);
                ph[0] = make_double2(bs * (acc[A4 + 0][j] + acc[A4 + 2][j]), bs * (acc[A4 + 1][j] + acc[A4 + 3][j]));
                ph[P.MP] = make_double2(bs * (acc[A4 + 1][j] - acc[A4 + 3][j]), -bs * (acc[A4 + 0][j] - acc[A4 + 2][j]));
            }
        }
    }
}

// =============================================================================================================
// analysis: (weighted) phase -> alm
// =============================================================================================================
template <int SPIN, int R, bool MIXED, int PAR>
__device__ __forceinline__ void anal_step(RingState<SPIN, R>& S, const double (&X)[(SPIN == 0 ? 4 : 8)][R], const double* rec,
                                          double (&part)[(SPIN == 0 ? 2 : 4)], int l)
{
    const double2 c = *reinterpret_cast<const double2*>(rec);
#pragma unroll
    for (int j = 0; j < R; ++j) {
        if (!MIXED || l >= S.la[j]) {
            const double p0 = (PAR == 0) ? S.p[0][j] : S.pp[0][j];
            const double p1 = (PAR == 0) ? S.p[SPIN != 0][j] : S.pp[SPIN != 0][j];
            if (SPIN == 0) {
                // even (l-m): X_N + X_S ; odd: X_N - X_S
                part[0] = fma(p0, X[2 * PAR + 0][j], part[0]);
                part[1] = fma(p0, X[2 * PAR + 1][j], part[1]);
            } else {
                // a+ += p+ Y+_N + sgn p- Y+_S ;  a- += p- Y-_N + sgn p+ Y-_S   (Y_S pre-multiplied by the base sign)
                // the four products of one function are adjacent so that its value is fetched once (operand reuse cache)
                constexpr int NXX = (SPIN == 0 ? 4 : 8), NPP = (SPIN == 0 ? 2 : 4);
                const double q0 = (PAR == 0) ? p0 : -p0, q1 = (PAR == 0) ? p1 : -p1;
                part[0] = fma(p0, X[0][j], part[0]);
                part[1] = fma(p0, X[1][j], part[1]);
                part[NPP - 2] = fma(q0, X[NXX - 2][j], part[NPP - 2]);
                part[NPP - 1] = fma(q0, X[NXX - 1][j], part[NPP - 1]);
                part[NPP - 2] = fma(p1, X[NXX - 4][j], part[NPP - 2]);
                part[NPP - 1] = fma(p1, X[NXX - 3][j], part[NPP - 1]);
                part[0] = fma(q1, X[2][j], part[0]);
                part[1] = fma(q1, X[3][j], part[1]);
            }
            rec_step<SPIN, R, PAR>(S, j, c.x, c.y);
        }
    }
}

// Sum of row (lane % ROWS) over the 32 source lanes; every lane of the LPP sharing a row returns the full sum.
template <int NV, int G>
__device__ __forceinline__ double2 red_sum(const double2* red, int lane)
{
    using M = RedMap<NV, G>;
    const double2* src = red + (lane % M::ROWS) * 33 + (lane / M::ROWS) * M::SL;
    double2 a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = src[i];
#pragma unroll
    for (int k = 4; k < M::SL; k += 4)
#pragma unroll
        for (int i = 0; i < 4; ++i) { const double2 q = src[k + i]; a[i].x += q.x; a[i].y += q.y; }
    double2 t = make_double2((a[0].x + a[1].x) + (a[2].x + a[3].x), (a[0].y + a[1].y) + (a[2].y + a[3].y));
    if (M::LPP == 2) { t.x += __shfl_xor_sync(0xffffffffu, t.x, 16); t.y += __shfl_xor_sync(0xffffffffu, t.y, 16); }
    return t;
}

template <int SPIN, int R>
__global__ void __launch_bounds__(LEG_NT) leg_anal(const LegParams P)
{
    constexpr int NX = (SPIN == 0) ? 4 : 8;
    constexpr int NPART = (SPIN == 0) ? 2 : 4;
    constexpr int G = RedT<SPIN>::G, NV = RedT<SPIN>::NV;
    constexpr int STEPS = RedT<SPIN>::STEPS;
    using M = RedMap<NV, G>;
    __shared__ __align__(16) double sbuf[2 * STEPS * 2];
    __shared__ __align__(16) double2 red[NV * G * 33];
    __shared__ __align__(8) unsigned long long sbar[2];
    const int lane = threadIdx.x;
    int row, chunk;
    leg_unit(P, row, chunk);
    const int m = P.m_list ? P.m_list[row] : (P.m_begin + row);
    const int l0 = (SPIN == 0) ? m : (m > 2 ? m : 2);
    const int pair0 = chunk * (32 * R);

    RingState<SPIN, R> S;
    int lmin, lmaxact, lstart;
    load_rings<SPIN, R>(P, m, l0, pair0, lane, S, lmin, lmaxact, lstart);
    if (lmin > P.lmax) return;   // nothing to add (outputs are pre-zeroed)

    // folded inputs
    double X[NX][R];
    const double bs = ((l0 + m) & 1) ? -1.0 : 1.0;
    const int col = P.col_is_row ? row : m;
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const int pair = pair0 + j * 32 + lane;
        double2 qN = make_double2(0.0, 0.0), qS = qN, uN = qN, uS = qN;
        if (S.la[j] != L_NEVER) {
            const int rN = P.ringN[pair], rS = P.ringS[pair];
            if (rN >= 0) { const double2* ph = phase_row(P, rN, col); qN = ph[0]; if (SPIN != 0) uN = ph[P.MP]; }
            if (rS >= 0) { const double2* ph = phase_row(P, rS, col); qS = ph[0]; if (SPIN != 0) uS = ph[P.MP]; }
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

    const int nl = P.lmax - lstart + 1;
    const int nmixed = (lmaxact - lstart + 1) & ~1;
    if (lane == 0) { mbar_init(&sbar[0], 1); mbar_init(&sbar[1], 1); mbar_init_fence(); }
    __syncwarp();
    const long long abase = alm_index(P.lmax, 0, m);
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
            // this lane's output element of the group and its normalisation, fetched ahead of the arithmetic that hides the latency
            const bool mine = (lane % G) < gcnt;
            const long long gk = abase + lstart + t0 + (lane % G);
            const double g = mine ? P.gamma[gk] : 0.0;
            if (t0 >= nmixed && gcnt == G) {
                // fast path: every live ring is active for the whole group -> straight-line code
#pragma unroll
                for (int s = 0; s < G; s += 2) {
                    const double* r0 = rec + (size_t)(g0 + s) * 2;
                    double part0[NPART], part1[NPART];
#pragma unroll
                    for (int k = 0; k < NPART; ++k) { part0[k] = 0.0; part1[k] = 0.0; }
                    anal_step<SPIN, R, false, 0>(S, X, r0, part0, 0);
                    anal_step<SPIN, R, false, 1>(S, X, r0 + 2, part1, 0);
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
                    anal_step<SPIN, R, true, 0>(S, X, r0, part0, lstart + t0 + s);
                    if (s + 1 < gcnt) anal_step<SPIN, R, true, 1>(S, X, r0 + 2, part1, lstart + t0 + s + 1);
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
                if (SPIN == 0) {
                    if (mine && lane < M::ROWS) {
                        atomicAdd(&P.alm_out0[gk].x, g * t.x);
                        if (m != 0) atomicAdd(&P.alm_out0[gk].y, g * t.y);
                    }
                } else {
                    // row v = 0 holds a+, v = 1 holds a-:  E = -(a+ + a-)/2 ; B = i (a+ - a-)/2.  The v = 0 lane of a step writes E, the v = 1 lane B.
                    const double ox = __shfl_xor_sync(0xffffffffu, t.x, G), oy = __shfl_xor_sync(0xffffffffu, t.y, G);
                    // g enters last: a product of g alone would be hoisted to the prefetch and stall on it
                    const bool isB = lane >= G;
                    const double re = 0.5 * ((isB ? (t.y - oy) : -(t.x + ox)) * g);
                    const double im = 0.5 * ((isB ? (ox - t.x) : -(t.y + oy)) * g);
                    if (mine) {
                        double2* out = (isB ? P.alm_out1 : P.alm_out0) + gk;
                        atomicAdd(&out->x, re);
                        if (m != 0) atomicAdd(&out->y, im);
                    }
                }
                __syncwarp();
            }
        }
        __syncwarp();   // every lane is done with this stage before it is refilled
    }
}

}  // namespace pixsht
