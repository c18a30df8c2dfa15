// fft.cuh -- the ring-FFT stage (phase <-> map), hand-written for sm_100a.  HBM bound: one read of
// the phase row and one write of the ring (or the reverse) per ring and component.
//
// Phase layout: element (ring, component, m) at ((ring_local * ncomp + c) * MP + m) -- one contiguous row of MP >= mmax+1
// complex doubles per (ring, component), so every global access of this stage is a coalesced row access (the Legendre
// kernels, which are FP64 bound and have the memory system idle, do the strided side of the transpose).
//
// One CTA transforms one ring of one component entirely in shared memory: the real ring of nphi samples is handled
// as a complex FFT of length n = nphi/2 (even/odd packing), in place, mixed radix (2/3/4/5 + generic odd primes <= 64):
//   synthesis (phase -> map): natural-order input, decimation-in-frequency passes, digit-reversed output that is
//                             un-permuted by the coalesced store;
//   analysis  (map -> phase): digit-reversed scatter on load, decimation-in-time passes, natural-order output.
// The two are exact transposes of each other, so one pass geometry and one permutation table serve both.  Odd radices
// come first in the pass order: the small-stride passes then have odd strides (no shared-memory bank conflicts), and the
// even radices run at strides >= 8 elements where consecutive threads touch consecutive addresses.
// Ring lengths outside that fast path (odd nphi: complex FFT of nphi samples, packed = 0; a prime factor > 64: pass_direct;
// nphi/2 samples that do not fit 227 KB) run the same passes on per-CTA work buffers in global memory (GLOBAL = true, one
// CTA per SM looping over the rings) -- libsharp2 takes any ring length, so must this.
// Fused into the same kernels:
//   - the e^{+-i m phi0} rotation and the quadrature weight (libsharp2's ring helper),
//   - m >= nphi/2 aliasing exactly as the direct sum prescribes (golden test at lmax = 3 nphi),
//   - the band bookkeeping of create_sht_band (src/transforms.jl:66-82): x/y flips and zero padding of partial-sky
//     rings are index arithmetic on the caller's array, and the un-flip/slice copy of alm2map (:220-225) disappears.
#pragma once
#include "common.cuh"

namespace pixsht {

// Per-phase cycle counters of the FFT kernels (profiling builds only: -DPIXSHT_FFT_PROF; read by tools/legbench)
#ifdef PIXSHT_FFT_PROF
__device__ unsigned long long g_fft_prof[32];
#define FFT_PROF_DECL long long prof_t_ = clock64()
#define FFT_PROF_MARK(slot) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_fft_prof[slot], (unsigned long long)(t_ - prof_t_)); prof_t_ = t_; } } while (0)
#define FFT_PROF_ARG , long long& prof_t_, int prof_base_
#define FFT_PROF_PASS(b) , prof_t_, b
#define FFT_PROF_MARK_PASS(ss) FFT_PROF_MARK(prof_base_ + (ss))
#else
#define FFT_PROF_DECL
#define FFT_PROF_MARK(slot)
#define FFT_PROF_ARG
#define FFT_PROF_PASS(b)
#define FFT_PROF_MARK_PASS(ss)
#endif

constexpr int FFT_MAXFAC = 24;
constexpr int FFT_MAXRADIX = 64;       // largest radix done in registers / local memory; larger primes take pass_direct
#ifndef PIXSHT_FFT_MAXTHREADS
#define PIXSHT_FFT_MAXTHREADS 640
#endif
constexpr int FFT_MAXTHREADS = PIXSHT_FFT_MAXTHREADS;

struct FftParams {
    int nphi, n;                // ring length, complex FFT length: nphi/2 (even nphi, two real samples per complex), else nphi
    int packed;                 // 1: even nphi, real ring packed into n = nphi/2 complex samples; 0: odd nphi, n = nphi, imaginary part zero
    int pt;                     // entries per pass table (128 lo + hi), FFT_PT unless the work buffer is in global memory
    void* gbuf;                 // nullptr: the ring lives in shared memory.  Else per-CTA work buffers in global memory (L2-resident):
    long long gslot;            //   rings too long for shared memory, or a prime factor > FFT_MAXRADIX; gslot entries per buffer,
    int galt;                   //   galt = 1: two buffers per CTA (pass_direct works out of place)
    int prefetch;               // 1: CTAs loop over several rings; prefetch the next ring's input row into L2 during the passes
    int nfac;
    int fac[FFT_MAXFAC];        // radices in DIT pass order (pass t works on sub-transforms of length L_t = prod_{u<t} fac[u])
    unsigned magic[FFT_MAXFAC]; // floor(2^32 / L_t) + 1: b / L_t == umulhi(b, magic) for b < 2^16
    int nsp;                    // super-passes: consecutive radix passes done together in registers (fft.cuh butterfly2)
    unsigned char sp_first[FFT_MAXFAC], sp_count[FFT_MAXFAC];   // first fine pass and number of fine passes (1 or 2), DIT order
    const double2* tw;          // [nphi] exp(-2 pi i t / nphi)
    const double2* phi0tw;      // [mmax+1] exp(+i m phi0)
    const double* wgt;          // [nrings] quadrature weight per band ring
    const unsigned short* perm; // [n] position of natural index j in the permuted buffer
    int mmax;
    double2* phase;             // this launch's rows: (ringlocal, c, m) at ((ringlocal*ncomp + c)*MP + m)
    const long long* mtab;      // m-sharded layout (multi-GPU), else nullptr: per m the address of element (ring 0, comp 0, m)
                                // in the buffer of the GPU that owns m (possibly peer memory) and that buffer's row length
    long long MP;
    int ncomp, c_begin;         // components per ring in the phase buffer; first component handled by this launch (grid.y of them)
    int ring_begin, ring_count; // band rings handled by this launch; ringlocal = ring - ring_begin
    int nx, ny, flipx, flipy;   // caller's map layout (column-major nx x ny), see pixsht_geom
    void* maps[4];              // up to 3 Stokes components, or a batch of up to 4 spin-0 maps
    int vec_ok;                 // 1: every map pointer of the launch is aligned to two elements, so that a full even-length row can be
                                // read / written as (x[2j], x[2j+1]) pairs in one access; 0: element-wise accesses (odd start offsets of
                                // caller views would otherwise fault on the vector path)
    int edge;                   // 1: kernels of fft_edge.cuh (first and last super-pass fused into the row I/O)
    int neg_mask;               // bit c set: component c of the caller's maps carries the opposite sign (IAU <-> COSMO Stokes U,
                                // src/enmap.jl:178-196 of the reference): negated here in the row I/O, no extra pass over the map
};

template <class T> struct alignas(2 * sizeof(T)) cpx { T x, y; };   // one 8- / 16-byte access in shared and global memory
template <class T> __device__ __forceinline__ cpx<T> cmul(cpx<T> a, cpx<T> b) { cpx<T> r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r; }
template <class T> __device__ __forceinline__ cpx<T> cadd(cpx<T> a, cpx<T> b) { cpx<T> r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <class T> __device__ __forceinline__ cpx<T> csub(cpx<T> a, cpx<T> b) { cpx<T> r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
template <class T> __device__ __forceinline__ cpx<T> cconj(cpx<T> a) { cpx<T> r; r.x = a.x; r.y = -a.y; return r; }
// multiply by +i (SIGN=+1) or -i (SIGN=-1)
template <class T, int SIGN> __device__ __forceinline__ cpx<T> cmuli(cpx<T> a) { cpx<T> r; if (SIGN > 0) { r.x = -a.y; r.y = a.x; } else { r.x = a.y; r.y = -a.x; } return r; }

// Twiddles: exp(SIGN * 2 pi i t / nphi) = A[t >> 7] * B[t & 127] from two small tables held in shared memory behind the ring
// buffer (A: every 128th entry of the global table, B: its first 128 entries).  Reading the full table from global memory
// cost ~4x the algorithmic HBM bytes in L2 traffic (ncu lts__t_bytes, profiles/r01): the 227 KB carve-out leaves no L1.
constexpr int FFT_TWLO_BITS = 7;
constexpr int FFT_TWLO = 1 << FFT_TWLO_BITS;
__host__ __device__ __forceinline__ int fft_tw_entries(int nphi) { return FFT_TWLO + (nphi + FFT_TWLO - 1) / FFT_TWLO; }

template <class T>
struct TwTab { const cpx<T>* A; const cpx<T>* B; };

template <class T>
__device__ __forceinline__ TwTab<T> tw_setup(const FftParams& P, cpx<T>* tab)
{
    const int na = (P.nphi + FFT_TWLO - 1) / FFT_TWLO;
    for (int i = threadIdx.x; i < FFT_TWLO + na; i += blockDim.x) {
        const double2 w = (i < FFT_TWLO) ? P.tw[i < P.nphi ? i : 0] : P.tw[(i - FFT_TWLO) * FFT_TWLO];
        cpx<T> v; v.x = (T)w.x; v.y = (T)w.y;
        tab[i] = v;
    }
    TwTab<T> t; t.B = tab; t.A = tab + FFT_TWLO;
    return t;
}

template <class T, int SIGN>
__device__ __forceinline__ cpx<T> twid(const TwTab<T>& W, int t)
{
    const cpx<T> a = W.A[t >> FFT_TWLO_BITS], b = W.B[t & (FFT_TWLO - 1)];
    cpx<T> r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x;
    if (SIGN > 0) r.y = -r.y;
    return r;
}

// exp(-2 pi i k / S), k < S, for every S <= FFT_CST_MAX: the twiddles between the two radix stages of a fused super-pass.
// Constant memory: after unrolling every index is a compile-time constant, so they become constant-bank operands.
constexpr int FFT_CST_MAX = 25;
#ifndef PIXSHT_EMU
__constant__ double2 c_fft_cst[(FFT_CST_MAX + 1) * FFT_CST_MAX];
#else
static double2 c_fft_cst[(FFT_CST_MAX + 1) * FFT_CST_MAX];
#endif
template <class T, int SIGN>
__device__ __forceinline__ cpx<T> fft_cst(int S, int k)
{
    const double2 w = c_fft_cst[S * FFT_CST_MAX + k];
    cpx<T> r; r.x = (T)w.x; r.y = (T)(SIGN < 0 ? w.y : -w.y);
    return r;
}

// Per-pass twiddle table W_{qL}^{kk}, kk < L, as a product of two conflict-free lookups: lo[kk & 127] * hi[kk >> 7]
// (consecutive threads have consecutive kk).  Built one pass ahead from the global table, double buffered, so that the
// barrier that ends a pass also publishes the next pass's table.
constexpr int FFT_PT = 256;   // 128 lo + up to 128 hi entries
template <class T>
__device__ __forceinline__ void pass_table_build(const FftParams& P, cpx<T>* tab, int L, int tstep)
{
    const int nlo = L < 128 ? L : 128, nhi = (L + 127) >> 7;
    for (int i = threadIdx.x; i < nlo + nhi; i += blockDim.x) {
        const int t = (i < nlo) ? i * tstep : (i - nlo) * 128 * tstep;
        const double2 w = P.tw[t];
        cpx<T> v; v.x = (T)w.x; v.y = (T)w.y;
        tab[(i < nlo) ? i : 128 + (i - nlo)] = v;
    }
}
template <class T, int SIGN>
__device__ __forceinline__ cpx<T> twp(const cpx<T>* tab, int kk, bool two_level)
{
    cpx<T> r = tab[kk & 127];
    if (two_level) r = cmul(tab[128 + (kk >> 7)], r);
    if (SIGN > 0) r.y = -r.y;
    return r;
}

// host side: position of natural index i in the permuted buffer (mixed-radix digit reversal for the DIT pass order fac[])
inline int fft_digit_reverse(const int* fac, int nfac, int n, int i)
{
    int pos = 0, L = n;
    for (int t = nfac - 1; t >= 0; --t) {
        const int q = fac[t];
        L /= q;
        pos += (i % q) * L;
        i /= q;
    }
    return pos;
}

// q-point DFT of a[0..q) in registers, SIGN = -1 forward / +1 inverse
template <class T, int SIGN>
__device__ __forceinline__ void dft2(cpx<T>* a) { const cpx<T> t = a[0]; a[0] = cadd(t, a[1]); a[1] = csub(t, a[1]); }
template <class T, int SIGN>
__device__ __forceinline__ void dft3(cpx<T>* a)
{
    const T c = (T)-0.5, s = (T)(SIGN * 0.86602540378443864676);
    const cpx<T> sum = cadd(a[1], a[2]), dif = csub(a[1], a[2]);
    cpx<T> m1; m1.x = a[0].x + c * sum.x; m1.y = a[0].y + c * sum.y;
    cpx<T> m2; m2.x = -s * dif.y; m2.y = s * dif.x;      // i*s*dif
    a[0] = cadd(a[0], sum); a[1] = cadd(m1, m2); a[2] = csub(m1, m2);
}
template <class T, int SIGN>
__device__ __forceinline__ void dft4(cpx<T>* a)
{
    const cpx<T> s02 = cadd(a[0], a[2]), d02 = csub(a[0], a[2]), s13 = cadd(a[1], a[3]), d13 = cmuli<T, SIGN>(csub(a[1], a[3]));
    a[0] = cadd(s02, s13); a[2] = csub(s02, s13);
    a[1] = cadd(d02, d13); a[3] = csub(d02, d13);
}
template <class T, int SIGN>
__device__ __forceinline__ void dft5(cpx<T>* a)
{
    const T c1 = (T)0.30901699437494742410, c2 = (T)-0.80901699437494742410;
    const T s1 = (T)(SIGN * 0.95105651629515357212), s2 = (T)(SIGN * 0.58778525229247312917);
    const cpx<T> p14 = cadd(a[1], a[4]), m14 = csub(a[1], a[4]), p23 = cadd(a[2], a[3]), m23 = csub(a[2], a[3]);
    cpx<T> r1; r1.x = a[0].x + c1 * p14.x + c2 * p23.x; r1.y = a[0].y + c1 * p14.y + c2 * p23.y;
    cpx<T> r2; r2.x = a[0].x + c2 * p14.x + c1 * p23.x; r2.y = a[0].y + c2 * p14.y + c1 * p23.y;
    cpx<T> i1; i1.x = -(s1 * m14.y + s2 * m23.y); i1.y = s1 * m14.x + s2 * m23.x;   // i*(s1 m14 + s2 m23)
    cpx<T> i2; i2.x = -(s2 * m14.y - s1 * m23.y); i2.y = s2 * m14.x - s1 * m23.x;   // i*(s2 m14 - s1 m23)
    a[0].x = a[0].x + p14.x + p23.x; a[0].y = a[0].y + p14.y + p23.y;
    a[1] = cadd(r1, i1); a[4] = csub(r1, i1);
    a[2] = cadd(r2, i2); a[3] = csub(r2, i2);
}

template <class T, int SIGN, int Q>
__device__ __forceinline__ void dftq(cpx<T>* a)
{
    if constexpr (Q == 2) dft2<T, SIGN>(a);
    else if constexpr (Q == 3) dft3<T, SIGN>(a);
    else if constexpr (Q == 4) dft4<T, SIGN>(a);
    else dft5<T, SIGN>(a);
}
// one butterfly of a radix-Q pass on registers; w1 = the pass twiddle W^kk (ignored when !tw).
// DIF == false: decimation in time (twiddle, then DFT);  DIF == true: the transpose (DFT, then twiddle).
template <class T, int SIGN, int Q, bool DIF>
__device__ __forceinline__ void butterfly_regs(cpx<T>* a, cpx<T> w1, bool tw)
{
    cpx<T> w2, w3, w4;
    w2.x = w3.x = w4.x = (T)1; w2.y = w3.y = w4.y = (T)0;
    if (tw) {
        if (Q > 2) w2 = cmul(w1, w1);   // squaring instead of a second table lookup (shared-memory pipe is the busier one)
        if (Q > 3) w3 = cmul(w1, w2);
        if (Q > 4) w4 = cmul(w2, w2);
    }
    auto WW = [&](int j) { return j == 1 ? w1 : (j == 2 ? w2 : (j == 3 ? w3 : w4)); };
    if (!DIF && tw) {
#pragma unroll
        for (int j = 1; j < Q; ++j) a[j] = cmul(a[j], WW(j));
    }
    dftq<T, SIGN, Q>(a);
    if (DIF && tw) {
#pragma unroll
        for (int j = 1; j < Q; ++j) a[j] = cmul(a[j], WW(j));
    }
}
// the same at element pointer e (stride L); tk = kk, the index into the pass's twiddle table
template <class T, int SIGN, int Q, bool DIF>
__device__ __forceinline__ void butterfly(const cpx<T>* ptab, cpx<T>* e, int L, int tk)
{
    cpx<T> a[Q];
#pragma unroll
    for (int j = 0; j < Q; ++j) a[j] = e[(size_t)j * L];
    cpx<T> w1; w1.x = (T)1; w1.y = (T)0;
    if (tk) w1 = twp<T, SIGN>(ptab, tk, L > 128);
    butterfly_regs<T, SIGN, Q, DIF>(a, w1, tk != 0);
#pragma unroll
    for (int j = 0; j < Q; ++j) e[(size_t)j * L] = a[j];
}

// generic (odd prime) radix, O(q^2)
template <class T, int SIGN, bool DIF>
__device__ void butterfly_generic(const FftParams& P, const TwTab<T>& W, cpx<T>* e, int q, int L, int tk)
{
    cpx<T> a[FFT_MAXRADIX];
    for (int j = 0; j < q; ++j) {
        cpx<T> v = e[(size_t)j * L];
        if (!DIF && tk && j) v = cmul(v, twid<T, SIGN>(W, j * tk));
        a[j] = v;
    }
    const int qstep = P.nphi / q;
    for (int u = 0; u < q; ++u) {
        cpx<T> s = a[0];
        for (int j = 1; j < q; ++j) s = cadd(s, cmul(a[j], twid<T, SIGN>(W, ((j * u) % q) * qstep)));
        if (DIF && tk && u) s = cmul(s, twid<T, SIGN>(W, u * tk));
        e[(size_t)u * L] = s;
    }
}

// Two consecutive radix passes (Q1 at sub-length L, then Q2 at sub-length Q1*L in DIT order) on Q1*Q2 elements held in
// registers: one shared-memory round trip and one barrier instead of two.  Element (j2, j1) sits at e[(j2*Q1 + j1)*L].
//   DIT: x *= WA^{j1};  DFT_Q1 over j1;  y(j2,u1) *= WB^{j2} * W_{Q1Q2}^{j2 u1};  DFT_Q2 over j2;   DIF: the transpose.
// TA / TB: the pass tables for the roots Q1*L and Q1*Q2*L, both indexed by kk.
// TW == false: the super-pass at sub-length 1 (kk == 0 everywhere), no pass twiddles at all
// the fused pair on registers a[j2][j1]; wa / wb = the pass twiddles W^kk of the roots Q1*L and Q1*Q2*L (ignored when !TW)
template <class T, int SIGN, int Q1, int Q2, bool DIF, bool TW>
__device__ __forceinline__ void butterfly2_regs(cpx<T> (*a)[Q1], cpx<T> wa, cpx<T> wb)
{
    constexpr int S = Q1 * Q2;
    // pass twiddles W^kk .. W^{4 kk} of the two roots as scalars (arrays that are only conditionally written end up in
    // local memory); for kk == 0 they are exactly 1
    cpx<T> wa1, wa2, wa3, wa4, wb1, wb2, wb3, wb4;
    wa1.x = wa2.x = wa3.x = wa4.x = wb1.x = wb2.x = wb3.x = wb4.x = (T)1;
    wa1.y = wa2.y = wa3.y = wa4.y = wb1.y = wb2.y = wb3.y = wb4.y = (T)0;
    if (TW) {
        wa1 = wa; wa2 = cmul(wa1, wa1);
        if (Q1 > 3) { wa3 = cmul(wa1, wa2); } if (Q1 > 4) { wa4 = cmul(wa2, wa2); }
        wb1 = wb; wb2 = cmul(wb1, wb1);
        if (Q2 > 3) { wb3 = cmul(wb1, wb2); } if (Q2 > 4) { wb4 = cmul(wb2, wb2); }
    }
    auto WA = [&](int j) { return j == 1 ? wa1 : (j == 2 ? wa2 : (j == 3 ? wa3 : wa4)); };
    auto WB = [&](int j) { return j == 1 ? wb1 : (j == 2 ? wb2 : (j == 3 ? wb3 : wb4)); };
    if (!DIF) {
#pragma unroll
        for (int j2 = 0; j2 < Q2; ++j2) {
            if (TW) {
#pragma unroll
                for (int j1 = 1; j1 < Q1; ++j1) a[j2][j1] = cmul(a[j2][j1], WA(j1));
            }
            dftq<T, SIGN, Q1>(a[j2]);
        }
    } else {
#pragma unroll
        for (int u1 = 0; u1 < Q1; ++u1) {
            cpx<T> col[Q2];
#pragma unroll
            for (int u2 = 0; u2 < Q2; ++u2) col[u2] = a[u2][u1];
            dftq<T, SIGN, Q2>(col);
#pragma unroll
            for (int j2 = 0; j2 < Q2; ++j2) a[j2][u1] = col[j2];
        }
    }
    // the stage between the two DFTs: diagonal in (j2, u1)
#pragma unroll
    for (int j2 = 1; j2 < Q2; ++j2) {
#pragma unroll
        for (int u1 = 0; u1 < Q1; ++u1) {
            if (u1) a[j2][u1] = cmul(a[j2][u1], fft_cst<T, SIGN>(S, (j2 * u1) % S));
            if (TW) a[j2][u1] = cmul(a[j2][u1], WB(j2));
        }
    }
    if (!DIF) {
#pragma unroll
        for (int u1 = 0; u1 < Q1; ++u1) {
            cpx<T> col[Q2];
#pragma unroll
            for (int j2 = 0; j2 < Q2; ++j2) col[j2] = a[j2][u1];
            dftq<T, SIGN, Q2>(col);
#pragma unroll
            for (int u2 = 0; u2 < Q2; ++u2) a[u2][u1] = col[u2];
        }
    } else {
#pragma unroll
        for (int j2 = 0; j2 < Q2; ++j2) {
            dftq<T, SIGN, Q1>(a[j2]);
            if (TW) {
#pragma unroll
                for (int j1 = 1; j1 < Q1; ++j1) a[j2][j1] = cmul(a[j2][j1], WA(j1));
            }
        }
    }
}
template <class T, int SIGN, int Q1, int Q2, bool DIF, bool TW>
__device__ __forceinline__ void butterfly2(const cpx<T>* TA, const cpx<T>* TB, cpx<T>* e, int L, int kk)
{
    cpx<T> a[Q2][Q1];
#pragma unroll
    for (int j2 = 0; j2 < Q2; ++j2)
#pragma unroll
        for (int j1 = 0; j1 < Q1; ++j1) a[j2][j1] = e[(size_t)(j2 * Q1 + j1) * L];
    cpx<T> wa, wb;
    wa.x = wb.x = (T)1; wa.y = wb.y = (T)0;
    if (TW) { wa = twp<T, SIGN>(TA, kk, L > 128); wb = twp<T, SIGN>(TB, kk, L > 128); }
    butterfly2_regs<T, SIGN, Q1, Q2, DIF, TW>(a, wa, wb);
#pragma unroll
    for (int j2 = 0; j2 < Q2; ++j2)
#pragma unroll
        for (int j1 = 0; j1 < Q1; ++j1) e[(size_t)(j2 * Q1 + j1) * L] = a[j2][j1];
}

template <class T, int SIGN, int Q1, int Q2, bool DIF>
__device__ __forceinline__ void pass2_loop(const cpx<T>* TA, const cpx<T>* TB, cpx<T>* buf, int L, int nb, unsigned magic)
{
    if (L == 1) {
        for (int b = threadIdx.x; b < nb; b += blockDim.x) butterfly2<T, SIGN, Q1, Q2, DIF, false>(TA, TB, buf + (size_t)b * (Q1 * Q2), 1, 0);
    } else {
        for (int b = threadIdx.x; b < nb; b += blockDim.x) {
            const int g = (int)__umulhi((unsigned)b, magic);
            const int kk = b - g * L;
            butterfly2<T, SIGN, Q1, Q2, DIF, true>(TA, TB, buf + (size_t)g * (Q1 * Q2) * L + kk, L, kk);
        }
    }
}
constexpr int FFT_FUSE_MAX = 16;   // largest fused super-pass (points held in registers per thread)
template <class T, int SIGN, int Q1, bool DIF>
__device__ __forceinline__ void pass2_dispatch(int q2, const cpx<T>* TA, const cpx<T>* TB, cpx<T>* buf, int L, int nb, unsigned magic)
{
    if constexpr (Q1 * 2 <= FFT_FUSE_MAX) { if (q2 == 2) { pass2_loop<T, SIGN, Q1, 2, DIF>(TA, TB, buf, L, nb, magic); return; } }
    if constexpr (Q1 * 3 <= FFT_FUSE_MAX) { if (q2 == 3) { pass2_loop<T, SIGN, Q1, 3, DIF>(TA, TB, buf, L, nb, magic); return; } }
    if constexpr (Q1 * 4 <= FFT_FUSE_MAX) { if (q2 == 4) { pass2_loop<T, SIGN, Q1, 4, DIF>(TA, TB, buf, L, nb, magic); return; } }
    if constexpr (Q1 * 5 <= FFT_FUSE_MAX) { if (q2 == 5) { pass2_loop<T, SIGN, Q1, 5, DIF>(TA, TB, buf, L, nb, magic); return; } }
}

// all butterflies of one radix-Q pass (Q a template parameter so that the loop body is straight-line code)
template <class T, int SIGN, int Q, bool DIF>
__device__ __forceinline__ void pass_loop(const cpx<T>* ptab, cpx<T>* buf, int L, int nb, unsigned magic)
{
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        const int g = (L == 1) ? b : (int)__umulhi((unsigned)b, magic);
        const int kk = b - g * L;
        butterfly<T, SIGN, Q, DIF>(ptab, buf + (size_t)g * Q * L + kk, L, kk);
    }
}

// tables of one super-pass: slot 0 = root q1*L, slot 1 = root q1*q2*L (pairs only)
template <class T>
__device__ __forceinline__ void superpass_tables(const FftParams& P, cpx<T>* tabs, int sp, int L)
{
    const int t = P.sp_first[sp], q1 = P.fac[t];
    if (q1 > 5) return;   // generic radices take their twiddles from the two-level root table
    pass_table_build<T>(P, tabs, L, P.nphi / (q1 * L));
    if (P.sp_count[sp] == 2) pass_table_build<T>(P, tabs + P.pt, L, P.nphi / (q1 * P.fac[t + 1] * L));
}
// sub-length L at which super-pass sp works (product of the radices before it)
__device__ __forceinline__ int superpass_L(const FftParams& P, int sp)
{
    int L = 1;
    for (int t = 0; t < P.sp_first[sp]; ++t) L *= P.fac[t];
    return L;
}

// One radix-q pass for a large prime q, computed directly (O(n q) per ring) and out of place:
//   DIT: out(g,u,kk) = sum_j in(g,j,kk) W_{qL}^{j kk} W_q^{j u};   DIF: out(g,u,kk) = W_{qL}^{u kk} sum_j in(g,j,kk) W_q^{j u}
// with element (g, j, kk) at (g q + j) L + kk.  One thread per output element.
template <class T, int SIGN, bool DIF>
__device__ void pass_direct(const FftParams& P, const TwTab<T>& W, const cpx<T>* src, cpx<T>* dst, int q, int L, unsigned magic)
{
    const int N = P.nphi;
    const int tstep = N / (q * L), qstep = N / q;
    for (int o = threadIdx.x; o < P.n; o += blockDim.x) {
        const int gu = (L == 1) ? o : (int)__umulhi((unsigned)o, magic);
        const int kk = o - gu * L;
        const int g = gu / q, u = gu - g * q;
        int s = u * qstep + (DIF ? 0 : kk * tstep);   // exponent step per input, < 2 N
        if (s >= N) s -= N;
        const cpx<T>* in = src + (size_t)g * q * L + kk;
        cpx<T> acc = in[0];
        int e = 0;
        for (int j = 1; j < q; ++j) {
            e += s; if (e >= N) e -= N;
            acc = cadd(acc, cmul(in[(size_t)j * L], twid<T, SIGN>(W, e)));
        }
        if (DIF && kk && u) acc = cmul(acc, twid<T, SIGN>(W, u * kk * tstep));
        dst[o] = acc;
    }
}

// mixed-radix passes over buf[0..n), in place except for pass_direct (which alternates between buf and alt).
// DIF == false: digit-reversed input -> natural output (super-passes in order); DIF == true: natural input -> digit-reversed
// output (super-passes in reverse order).  SIGN = -1 forward.  Returns the buffer that holds the result.
// ptabs: 2 x 2 pass tables, double buffered; the tables of the FIRST super-pass were built by the caller before its last barrier.
// Runs the super-passes sp_lo .. sp_lo + sp_cnt - 1 (all of them in the plain kernels; the inner ones in fft_edge.cuh).
template <class T, int SIGN, bool DIF>
__device__ cpx<T>* fft_passes(const FftParams& P, const TwTab<T>& W, cpx<T>* buf, cpx<T>* alt, cpx<T>* ptabs, int sp_lo, int sp_cnt FFT_PROF_ARG)
{
    const int n = P.n;
    for (int ss = 0; ss < sp_cnt; ++ss) {
        const int sp = DIF ? (sp_lo + sp_cnt - 1 - ss) : (sp_lo + ss);
        const int t = P.sp_first[sp], q1 = P.fac[t];
        const int L = superpass_L(P, sp);
        const unsigned magic = P.magic[t];
        const cpx<T>* cur = ptabs + (ss & 1) * 2 * P.pt;
        if (ss + 1 < sp_cnt) {   // next super-pass's tables (read only after the barrier below)
            const int sp2 = DIF ? (sp - 1) : (sp + 1);
            superpass_tables<T>(P, ptabs + ((ss + 1) & 1) * 2 * P.pt, sp2, superpass_L(P, sp2));
        }
        if (P.sp_count[sp] == 2) {
            const int q2 = P.fac[t + 1], nb = n / (q1 * q2);
            if (q1 == 2) pass2_dispatch<T, SIGN, 2, DIF>(q2, cur, cur + P.pt, buf, L, nb, magic);
            else if (q1 == 3) pass2_dispatch<T, SIGN, 3, DIF>(q2, cur, cur + P.pt, buf, L, nb, magic);
            else if (q1 == 4) pass2_dispatch<T, SIGN, 4, DIF>(q2, cur, cur + P.pt, buf, L, nb, magic);
            else pass2_dispatch<T, SIGN, 5, DIF>(q2, cur, cur + P.pt, buf, L, nb, magic);
        } else {
            const int nb = n / q1;
            if (q1 == 4) pass_loop<T, SIGN, 4, DIF>(cur, buf, L, nb, magic);
            else if (q1 == 3) pass_loop<T, SIGN, 3, DIF>(cur, buf, L, nb, magic);
            else if (q1 == 5) pass_loop<T, SIGN, 5, DIF>(cur, buf, L, nb, magic);
            else if (q1 == 2) pass_loop<T, SIGN, 2, DIF>(cur, buf, L, nb, magic);
            else if (q1 <= FFT_MAXRADIX) {
                const int tstep = P.nphi / (q1 * L);   // W_{qL}^{a} = tw[a * tstep]
                for (int b = threadIdx.x; b < nb; b += blockDim.x) {
                    const int g = (L == 1) ? b : (int)__umulhi((unsigned)b, magic);
                    const int kk = b - g * L;
                    butterfly_generic<T, SIGN, DIF>(P, W, buf + (size_t)g * q1 * L + kk, q1, L, kk * tstep);
                }
            } else {
                pass_direct<T, SIGN, DIF>(P, W, buf, alt, q1, L, magic);
                cpx<T>* sw = buf; buf = alt; alt = sw;
            }
        }
        __syncthreads();
        FFT_PROF_MARK_PASS(ss);
    }
    return buf;
}

// work buffers of this CTA: shared memory (ring + tables; GLOBAL == false keeps the pointers provably shared so that the passes
// compile to LDS/STS), or a global-memory slot with the tables alone in shared memory
template <class T, bool GLOBAL>
__device__ __forceinline__ void fft_buffers(const FftParams& P, unsigned char* smem_raw, cpx<T>*& buf, cpx<T>*& alt, cpx<T>*& tabs)
{
    cpx<T>* sm = reinterpret_cast<cpx<T>*>(smem_raw);
    alt = nullptr;
    if (GLOBAL) {
        const long long slot = (long long)blockIdx.y * gridDim.x + blockIdx.x;
        buf = reinterpret_cast<cpx<T>*>(P.gbuf) + slot * P.gslot * (P.galt ? 2 : 1);
        if (P.galt) alt = buf + P.gslot;
        tabs = sm;
    } else {
        buf = sm;
        tabs = sm + P.n + 1;
    }
}

// phase element (band ring, component c, m): in this launch's local rows, or -- m-sharded -- in the buffer of m's owner
__device__ __forceinline__ double2* phase_elem(const FftParams& P, double2* localrow, int ring, int c, int m)
{
    if (!P.mtab) return localrow + m;
    const long long base = P.mtab[2 * m], mp = P.mtab[2 * m + 1];
    return reinterpret_cast<double2*>(base) + ((long long)ring * P.ncomp + c) * mp;
}

// aliased half-spectrum entry X[k], 0 <= k <= n: sum over m == +-k (mod nphi) of the rotated phases (general mmax)
template <class T>
__device__ __forceinline__ cpx<T> load_X(const FftParams& P, double2* row, int ring, int c, int k, double sg)
{
    double sx = 0.0, sy = 0.0;
    for (int m = k; m <= P.mmax; m += P.nphi) {
        const double2 a = *phase_elem(P, row, ring, c, m), r = P.phi0tw[m];
        sx += a.x * r.x - a.y * r.y; sy += a.x * r.y + a.y * r.x;
    }
    for (int m = P.nphi - k; m <= P.mmax; m += P.nphi) {
        const double2 a = *phase_elem(P, row, ring, c, m), r = P.phi0tw[m];
        sx += a.x * r.x - a.y * r.y; sy -= a.x * r.y + a.y * r.x;
    }
    cpx<T> v; v.x = (T)(sg * sx); v.y = (T)(sg * sy);
    return v;
}

// ring samples (x[2j], x[2j+1]) of the caller's row as one complex: a single 2-element vector access when the row is a full,
// even-length ring (the band bookkeeping of create_sht_band otherwise: flips and zero padding by index arithmetic)
template <class T> struct Pair2 { T x, y; };
template <class T>
__device__ __forceinline__ cpx<T> load_pair(const FftParams& P, const T* irow, int j, bool vec)
{
    cpx<T> z;
    if (vec) {
        const int i = P.flipx ? (P.nx - 2 - 2 * j) : 2 * j;
        const cpx<T> v = *reinterpret_cast<const cpx<T>*>(irow + i);
        if (P.flipx) { z.x = v.y; z.y = v.x; } else z = v;
    } else {
        const int i0 = 2 * j, i1 = 2 * j + 1;
        z.x = (i0 < P.nx) ? irow[P.flipx ? (P.nx - 1 - i0) : i0] : (T)0;
        z.y = (i1 < P.nx) ? irow[P.flipx ? (P.nx - 1 - i1) : i1] : (T)0;
    }
    return z;
}
template <class T>
__device__ __forceinline__ void store_pair(const FftParams& P, T* orow, int j, cpx<T> z, bool vec)
{
    if (vec) {
        const int i = P.flipx ? (P.nx - 2 - 2 * j) : 2 * j;
        cpx<T> v; if (P.flipx) { v.x = z.y; v.y = z.x; } else v = z;
        *reinterpret_cast<cpx<T>*>(orow + i) = v;
    } else {
        const int i0 = 2 * j, i1 = 2 * j + 1;
        if (i0 < P.nx) orow[P.flipx ? (P.nx - 1 - i0) : i0] = z.x;
        if (i1 < P.nx) orow[P.flipx ? (P.nx - 1 - i1) : i1] = z.y;
    }
}
// L2 prefetch of `bytes` starting at p, one 128-byte line per thread and iteration
__device__ __forceinline__ void prefetch_row(const void* p, size_t bytes)
{
    const char* c = reinterpret_cast<const char*>(p);
    for (size_t off = (size_t)threadIdx.x * 128; off < bytes; off += (size_t)blockDim.x * 128) prefetch_l2(c + off);
}
constexpr int FFT_IO_UNROLL = 4;   // independent global accesses in flight per thread in the load / store loops

// phase -> map  (synthesis).  grid = (rows, ncomp): CTA x handles band rings ring_begin + x, x + gridDim.x, ...
template <class T, bool GLOBAL>
__global__ void __launch_bounds__(FFT_MAXTHREADS) fft_phase2map(const FftParams P)
{
    PIXSHT_DYN_SMEM(smem_raw);
    cpx<T>*buf, *alt, *tabs;
    fft_buffers<T, GLOBAL>(P, smem_raw, buf, alt, tabs);   // buf: n + 1 entries
    const TwTab<T> W = tw_setup<T>(P, tabs);
    cpx<T>* ptabs = tabs + fft_tw_entries(P.nphi);   // 2 x 2 pass tables
    const int c = P.c_begin + blockIdx.y;
    const int n = P.n, N = P.nphi;
    T* out = reinterpret_cast<T*>(P.maps[c]);
    const bool vec = (P.nx == P.nphi) && P.packed && P.vec_ok;
    const double sg = ((P.neg_mask >> c) & 1) ? -1.0 : 1.0;
    FFT_PROF_DECL;

    for (int rl = blockIdx.x; rl < P.ring_count; rl += gridDim.x) {
        const int ring = P.ring_begin + rl;
        double2* row = P.phase + ((long long)rl * P.ncomp + c) * P.MP;
        if (P.packed) {
            // X[k], k = 0..n, natural order (coalesced row read)
            if (P.mmax <= n) {
                for (int k0 = threadIdx.x; k0 <= n; k0 += FFT_IO_UNROLL * blockDim.x) {
                    double2 a[FFT_IO_UNROLL], r[FFT_IO_UNROLL];
#pragma unroll
                    for (int u = 0; u < FFT_IO_UNROLL; ++u) {
                        const int k = k0 + u * blockDim.x;
                        a[u] = make_double2(0.0, 0.0); r[u] = a[u];
                        if (k <= P.mmax) { a[u] = *phase_elem(P, row, ring, c, k); r[u] = P.phi0tw[k]; r[u].x *= sg; r[u].y *= sg; }
                    }
#pragma unroll
                    for (int u = 0; u < FFT_IO_UNROLL; ++u) {
                        const int k = k0 + u * blockDim.x;
                        if (k > n) break;
                        const double sx = a[u].x * r[u].x - a[u].y * r[u].y, sy = a[u].x * r[u].y + a[u].y * r[u].x;
                        cpx<T> v;
                        if (k == n) { v.x = (T)(2.0 * sx); v.y = (T)0; }          // m = nphi/2: the ring carries only the (doubled) real part
                        else { v.x = (T)sx; v.y = (T)sy; }
                        buf[k] = v;
                    }
                }
            } else {
                for (int k = threadIdx.x; k <= n; k += blockDim.x) buf[k] = load_X<T>(P, row, ring, c, k, sg);
            }
            __syncthreads();
            FFT_PROF_MARK(0);

            // pre-processing in place: Z[k] = (X[k] + conj X[n-k]) + i (X[k] - conj X[n-k]) e^{+2 pi i k/nphi}
            for (int k = threadIdx.x; k <= n / 2; k += blockDim.x) {
                if (k == 0) {
                    const cpx<T> x0 = buf[0], xn = buf[n];
                    cpx<T> z; z.x = x0.x + xn.x; z.y = x0.x - xn.x;
                    buf[0] = z;
                } else {
                    const cpx<T> xa = buf[k], xb = buf[n - k];
                    const cpx<T> wa = twid<T, +1>(W, k);
                    const cpx<T> ea = cadd(xa, cconj(xb)), oa = cmul(csub(xa, cconj(xb)), wa);
                    if (k != n - k) {
                        const cpx<T> wb = twid<T, +1>(W, n - k);
                        const cpx<T> eb = cadd(xb, cconj(xa)), ob = cmul(csub(xb, cconj(xa)), wb);
                        buf[n - k] = cadd(eb, cmuli<T, +1>(ob));
                    }
                    buf[k] = cadd(ea, cmuli<T, +1>(oa));
                }
            }
        } else {
            // odd ring length: the full Hermitian spectrum Z[k] = X[k], Z[N-k] = conj X[k] of the real ring
            for (int k = threadIdx.x; k <= N / 2; k += blockDim.x) {
                cpx<T> x = load_X<T>(P, row, ring, c, k, sg);
                if (k == 0) { x.y = (T)0; buf[0] = x; }
                else { buf[k] = x; buf[N - k] = cconj(x); }
            }
        }
        superpass_tables<T>(P, ptabs, P.nsp - 1, superpass_L(P, P.nsp - 1));
        __syncthreads();
        FFT_PROF_MARK(1);
        if (P.prefetch && !P.mtab && rl + (int)gridDim.x < P.ring_count)
            prefetch_row(P.phase + ((long long)(rl + gridDim.x) * P.ncomp + c) * P.MP, (size_t)(P.mmax + 1) * sizeof(double2));
        const cpx<T>* res = fft_passes<T, +1, true>(P, W, buf, alt, ptabs, 0, P.nsp FFT_PROF_PASS(4));
        if (!GLOBAL) res = buf;   // no out-of-place pass without the global buffers: keeps the pointer provably shared

        // store into the caller's array (flips / partial rings by index arithmetic): packed x[2j] = Re z[j], x[2j+1] = Im z[j]
        const int rowy = P.flipy ? (P.ny - 1 - ring) : ring;
        T* orow = out + (size_t)rowy * P.nx;
        if (P.packed) {
            for (int j0 = threadIdx.x; j0 < n; j0 += FFT_IO_UNROLL * blockDim.x) {
                int pj[FFT_IO_UNROLL];
#pragma unroll
                for (int u = 0; u < FFT_IO_UNROLL; ++u) { const int j = j0 + u * blockDim.x; pj[u] = (j < n) ? (int)P.perm[j] : 0; }
#pragma unroll
                for (int u = 0; u < FFT_IO_UNROLL; ++u) {
                    const int j = j0 + u * blockDim.x;
                    if (j < n) store_pair<T>(P, orow, j, res[pj[u]], vec);
                }
            }
        } else {
            for (int j = threadIdx.x; j < P.nx; j += blockDim.x) orow[P.flipx ? (P.nx - 1 - j) : j] = res[P.perm[j]].x;
        }
        __syncthreads();   // the buffer is reused by this CTA's next ring
        FFT_PROF_MARK(2);
    }
}

// map -> weighted phase  (analysis).  grid as above
template <class T, bool GLOBAL>
__global__ void __launch_bounds__(FFT_MAXTHREADS) fft_map2phase(const FftParams P)
{
    PIXSHT_DYN_SMEM(smem_raw);
    cpx<T>*buf, *alt, *tabs;
    fft_buffers<T, GLOBAL>(P, smem_raw, buf, alt, tabs);   // buf: n + 1 entries
    const TwTab<T> W = tw_setup<T>(P, tabs);
    cpx<T>* ptabs = tabs + fft_tw_entries(P.nphi);   // 2 x 2 pass tables
    const int c = P.c_begin + blockIdx.y;
    const int n = P.n, N = P.nphi;
    const T* in = reinterpret_cast<const T*>(P.maps[c]);
    const bool vec = (P.nx == P.nphi) && P.packed && P.vec_ok;
    const double sg = ((P.neg_mask >> c) & 1) ? -1.0 : 1.0;
    FFT_PROF_DECL;

    for (int rl = blockIdx.x; rl < P.ring_count; rl += gridDim.x) {
        const int ring = P.ring_begin + rl;
        const int rowy = P.flipy ? (P.ny - 1 - ring) : ring;
        const T* irow = in + (size_t)rowy * P.nx;
        if (P.packed) {
            for (int j0 = threadIdx.x; j0 < n; j0 += FFT_IO_UNROLL * blockDim.x) {
                cpx<T> z[FFT_IO_UNROLL]; int pj[FFT_IO_UNROLL];
#pragma unroll
                for (int u = 0; u < FFT_IO_UNROLL; ++u) {
                    const int j = j0 + u * blockDim.x;
                    pj[u] = 0; z[u].x = (T)0; z[u].y = (T)0;
                    if (j < n) { pj[u] = (int)P.perm[j]; z[u] = load_pair<T>(P, irow, j, vec); }
                }
#pragma unroll
                for (int u = 0; u < FFT_IO_UNROLL; ++u) if ((int)(j0 + u * blockDim.x) < n) buf[pj[u]] = z[u];
            }
        } else {
            for (int j = threadIdx.x; j < N; j += blockDim.x) {
                cpx<T> z; z.y = (T)0;
                z.x = (j < P.nx) ? irow[P.flipx ? (P.nx - 1 - j) : j] : (T)0;
                buf[P.perm[j]] = z;
            }
        }
        superpass_tables<T>(P, ptabs, 0, 1);
        __syncthreads();
        FFT_PROF_MARK(16);
        if (P.prefetch && rl + (int)gridDim.x < P.ring_count) {
            const int ringn = ring + (int)gridDim.x;
            prefetch_row(in + (size_t)(P.flipy ? (P.ny - 1 - ringn) : ringn) * P.nx, (size_t)P.nx * sizeof(T));
        }
        cpx<T>* res = fft_passes<T, -1, false>(P, W, buf, alt, ptabs, 0, P.nsp FFT_PROF_PASS(20));
        if (!GLOBAL) res = buf;   // no out-of-place pass without the global buffers: keeps the pointer provably shared

        if (P.packed) {
            // post-processing in place: F[k] = ((Z[k] + conj Z[n-k]) - i e^{-2 pi i k/N} (Z[k] - conj Z[n-k])) / 2,  k = 0..n
            for (int k = threadIdx.x; k <= n / 2; k += blockDim.x) {
                if (k == 0) {
                    const cpx<T> z0 = res[0];
                    cpx<T> f0, fn; f0.x = z0.x + z0.y; f0.y = 0; fn.x = z0.x - z0.y; fn.y = 0;
                    res[0] = f0; res[n] = fn;
                } else {
                    const cpx<T> za = res[k], zb = res[n - k];
                    const cpx<T> ea = cadd(za, cconj(zb)), oa = cmul(csub(za, cconj(zb)), twid<T, -1>(W, k));
                    const cpx<T> fa = cadd(ea, cmuli<T, -1>(oa));
                    cpx<T> r; r.x = (T)0.5 * fa.x; r.y = (T)0.5 * fa.y;
                    if (k != n - k) {
                        const cpx<T> eb = cadd(zb, cconj(za)), ob = cmul(csub(zb, cconj(za)), twid<T, -1>(W, n - k));
                        const cpx<T> fb = cadd(eb, cmuli<T, -1>(ob));
                        cpx<T> rb; rb.x = (T)0.5 * fb.x; rb.y = (T)0.5 * fb.y;
                        res[n - k] = rb;
                    }
                    res[k] = r;
                }
            }
            __syncthreads();
            FFT_PROF_MARK(17);
        }

        // phase_m = w * e^{-i m phi0} * F[m mod N]  (packed: conjugate symmetric upper half); coalesced row write
        const double w = sg * P.wgt[ring];
        double2* row = P.phase + ((long long)rl * P.ncomp + c) * P.MP;
        for (int m0 = threadIdx.x; m0 <= P.mmax; m0 += FFT_IO_UNROLL * blockDim.x) {
            double2 r[FFT_IO_UNROLL];
#pragma unroll
            for (int u = 0; u < FFT_IO_UNROLL; ++u) { const int m = m0 + u * blockDim.x; r[u] = (m <= P.mmax) ? P.phi0tw[m] : make_double2(0.0, 0.0); }
#pragma unroll
            for (int u = 0; u < FFT_IO_UNROLL; ++u) {
                const int m = m0 + u * blockDim.x;
                if (m > P.mmax) break;
                const int kk = (m < N) ? m : (m % N);
                const cpx<T> f = (kk <= n) ? res[kk] : cconj(res[N - kk]);
                const double fx = (double)f.x, fy = (double)f.y;
                *phase_elem(P, row, ring, c, m) = make_double2(w * (fx * r[u].x + fy * r[u].y), w * (fy * r[u].x - fx * r[u].y));
            }
        }
        __syncthreads();   // the buffer is reused by this CTA's next ring
        FFT_PROF_MARK(18);
    }
}

}  // namespace pixsht
