// fft.cuh -- the ring-FFT stage (phase <-> map), hand-written for sm_100a.  HBM bound: one read of
// the phase row set and one write of the ring (or the reverse) per ring.
//
// One CTA transforms one ring of one component entirely in shared memory: the real ring of nphi samples is handled
// as a complex FFT of length n = nphi/2 (even/odd packing) done in place as a mixed-radix decimation-in-time
// transform (radices 4/2/3/5 + generic odd primes) on digit-reversed input.  Fused into the same kernel:
//   - the e^{+-i m phi0} rotation and the quadrature weight (libsharp2's ring helper),
//   - m >= nphi/2 aliasing exactly as the direct sum prescribes (golden test at lmax = 3 nphi),
//   - the band bookkeeping of create_sht_band (src/transforms.jl:66-82): x/y flips and zero padding of partial-sky
//     rings are index arithmetic on the caller's array, and the un-flip/slice copy of alm2map (:220-225) disappears.
#pragma once
#include "common.cuh"

namespace pixsht {

constexpr int FFT_MAXFAC = 24;
constexpr int FFT_MAXRADIX = 64;

struct FftParams {
    int nphi, n;                // ring length, complex FFT length (nphi/2)
    int nfac;
    int fac[FFT_MAXFAC];        // radices in pass order (pass t works on sub-transforms of length L_t = prod_{u<t} fac[u])
    const double2* tw;          // [nphi] exp(-2 pi i t / nphi)
    const double2* phi0tw;      // [mmax+1] exp(+i m phi0)
    const double* wgt;          // [nrings] quadrature weight per band ring
    int mmax;
    const int* m_row;           // phase row of m (nullptr: row = m)
    double2* phase;             // element (c,row,ringlocal) at c*stride_c + row*stride_m + ringlocal
    long long stride_c, stride_m;
    int ring_begin, ring_count; // band rings handled by this launch; ringlocal = ring - ring_begin
    int nx, ny, flipx, flipy;   // caller's map layout (column-major nx x ny), see pixsht_geom
    void* maps[3];
};

template <class T> struct cpx { T x, y; };
template <class T> __device__ __forceinline__ cpx<T> cmul(cpx<T> a, cpx<T> b) { cpx<T> r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r; }
template <class T> __device__ __forceinline__ cpx<T> cadd(cpx<T> a, cpx<T> b) { cpx<T> r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <class T> __device__ __forceinline__ cpx<T> csub(cpx<T> a, cpx<T> b) { cpx<T> r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
template <class T> __device__ __forceinline__ cpx<T> cconj(cpx<T> a) { cpx<T> r; r.x = a.x; r.y = -a.y; return r; }
// multiply by +i (SIGN=+1) or -i (SIGN=-1)
template <class T, int SIGN> __device__ __forceinline__ cpx<T> cmuli(cpx<T> a) { cpx<T> r; if (SIGN > 0) { r.x = -a.y; r.y = a.x; } else { r.x = a.y; r.y = -a.x; } return r; }

// exp(SIGN * 2 pi i t / nphi) from the forward table
template <class T, int SIGN>
__device__ __forceinline__ cpx<T> twid(const FftParams& P, long long t)
{
    const double2 w = P.tw[t];
    cpx<T> r; r.x = (T)w.x; r.y = (T)(SIGN < 0 ? w.y : -w.y);
    return r;
}

__device__ __forceinline__ int digit_reverse(const FftParams& P, int i)
{
    int pos = 0, L = P.n;
    for (int t = P.nfac - 1; t >= 0; --t) {
        const int q = P.fac[t];
        L /= q;
        pos += (i % q) * L;
        i /= q;
    }
    return pos;
}

// in-place mixed-radix DIT passes over buf[0..n) (digit-reversed input, natural-order output). SIGN=-1 forward.
template <class T, int SIGN>
__device__ void fft_passes(const FftParams& P, cpx<T>* buf)
{
    const int n = P.n;
    int L = 1;
    for (int t = 0; t < P.nfac; ++t) {
        const int q = P.fac[t];
        const int nb = n / q;
        const long long tstep = P.nphi / (q * L);   // W_{qL}^{a} = tw[a * tstep]
        for (int b = threadIdx.x; b < nb; b += blockDim.x) {
            const int kk = b % L, g = b / L;
            cpx<T>* e = buf + (size_t)g * q * L + kk;
            if (q == 2) {
                cpx<T> a0 = e[0], a1 = e[L];
                if (kk) a1 = cmul(a1, twid<T, SIGN>(P, (long long)kk * tstep));
                e[0] = cadd(a0, a1); e[L] = csub(a0, a1);
            } else if (q == 4) {
                cpx<T> a0 = e[0], a1 = e[L], a2 = e[2 * L], a3 = e[3 * L];
                if (kk) {
                    a1 = cmul(a1, twid<T, SIGN>(P, (long long)kk * tstep));
                    a2 = cmul(a2, twid<T, SIGN>(P, 2LL * kk * tstep));
                    a3 = cmul(a3, twid<T, SIGN>(P, 3LL * kk * tstep));
                }
                const cpx<T> s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = cmuli<T, SIGN>(csub(a1, a3));
                e[0] = cadd(s02, s13); e[2 * L] = csub(s02, s13);
                e[L] = cadd(d02, d13); e[3 * L] = csub(d02, d13);
            } else if (q == 3) {
                cpx<T> a0 = e[0], a1 = e[L], a2 = e[2 * L];
                if (kk) {
                    a1 = cmul(a1, twid<T, SIGN>(P, (long long)kk * tstep));
                    a2 = cmul(a2, twid<T, SIGN>(P, 2LL * kk * tstep));
                }
                const T c = (T)-0.5, s = (T)(SIGN * 0.86602540378443864676);
                const cpx<T> sum = cadd(a1, a2), dif = csub(a1, a2);
                cpx<T> m1; m1.x = a0.x + c * sum.x; m1.y = a0.y + c * sum.y;
                cpx<T> m2; m2.x = -s * dif.y; m2.y = s * dif.x;      // i*s*dif
                e[0] = cadd(a0, sum); e[L] = cadd(m1, m2); e[2 * L] = csub(m1, m2);
            } else if (q == 5) {
                cpx<T> a0 = e[0], a1 = e[L], a2 = e[2 * L], a3 = e[3 * L], a4 = e[4 * L];
                if (kk) {
                    a1 = cmul(a1, twid<T, SIGN>(P, (long long)kk * tstep));
                    a2 = cmul(a2, twid<T, SIGN>(P, 2LL * kk * tstep));
                    a3 = cmul(a3, twid<T, SIGN>(P, 3LL * kk * tstep));
                    a4 = cmul(a4, twid<T, SIGN>(P, 4LL * kk * tstep));
                }
                const T c1 = (T)0.30901699437494742410, c2 = (T)-0.80901699437494742410;
                const T s1 = (T)(SIGN * 0.95105651629515357212), s2 = (T)(SIGN * 0.58778525229247312917);
                const cpx<T> p14 = cadd(a1, a4), m14 = csub(a1, a4), p23 = cadd(a2, a3), m23 = csub(a2, a3);
                cpx<T> r1; r1.x = a0.x + c1 * p14.x + c2 * p23.x; r1.y = a0.y + c1 * p14.y + c2 * p23.y;
                cpx<T> r2; r2.x = a0.x + c2 * p14.x + c1 * p23.x; r2.y = a0.y + c2 * p14.y + c1 * p23.y;
                cpx<T> i1; i1.x = -(s1 * m14.y + s2 * m23.y); i1.y = s1 * m14.x + s2 * m23.x;   // i*(s1 m14 + s2 m23)
                cpx<T> i2; i2.x = -(s2 * m14.y - s1 * m23.y); i2.y = s2 * m14.x - s1 * m23.x;   // i*(s2 m14 - s1 m23)
                e[0].x = a0.x + p14.x + p23.x; e[0].y = a0.y + p14.y + p23.y;
                e[L] = cadd(r1, i1); e[4 * L] = csub(r1, i1);
                e[2 * L] = cadd(r2, i2); e[3 * L] = csub(r2, i2);
            } else {
                // generic (odd prime) radix, O(q^2)
                cpx<T> a[FFT_MAXRADIX];
                for (int j = 0; j < q; ++j) {
                    cpx<T> v = e[(size_t)j * L];
                    if (kk && j) v = cmul(v, twid<T, SIGN>(P, (long long)j * kk * tstep));
                    a[j] = v;
                }
                const long long qstep = P.nphi / q;
                for (int u = 0; u < q; ++u) {
                    cpx<T> s = a[0];
                    for (int j = 1; j < q; ++j) s = cadd(s, cmul(a[j], twid<T, SIGN>(P, (long long)((j * u) % q) * qstep)));
                    e[(size_t)u * L] = s;
                }
            }
        }
        L *= q;
        __syncthreads();
    }
}

// aliased half-spectrum entry X[k], 0 <= k <= n, of ring `rl` (local index): sum over m == +-k (mod nphi) of the rotated phases
template <class T>
__device__ __forceinline__ cpx<T> load_X(const FftParams& P, const double2* ph, int k)
{
    double sx = 0.0, sy = 0.0;
    for (int m = k; m <= P.mmax; m += P.nphi) {
        const int row = P.m_row ? P.m_row[m] : m;
        const double2 a = ph[(long long)row * P.stride_m], r = P.phi0tw[m];
        sx += a.x * r.x - a.y * r.y; sy += a.x * r.y + a.y * r.x;
    }
    for (int m = P.nphi - k; m <= P.mmax; m += P.nphi) {
        const int row = P.m_row ? P.m_row[m] : m;
        const double2 a = ph[(long long)row * P.stride_m], r = P.phi0tw[m];
        sx += a.x * r.x - a.y * r.y; sy -= a.x * r.y + a.y * r.x;
    }
    cpx<T> v; v.x = (T)sx; v.y = (T)sy;
    return v;
}

// phase -> map  (synthesis).  grid = (ring_count, ncomp)
template <class T>
__global__ void fft_phase2map(const FftParams P)
{
    PIXSHT_DYN_SMEM(smem_raw);
    cpx<T>* buf = reinterpret_cast<cpx<T>*>(smem_raw);
    const int rl = blockIdx.x, c = blockIdx.y, ring = P.ring_begin + rl;
    const int n = P.n;
    const double2* ph = P.phase + (long long)c * P.stride_c + rl;

    // pre-processing: Z[k] = (X[k] + conj X[n-k]) + i (X[k] - conj X[n-k]) e^{+2 pi i k/nphi}, stored digit-reversed
    for (int k = threadIdx.x; k <= n / 2; k += blockDim.x) {
        if (k == 0) {
            const cpx<T> x0 = load_X<T>(P, ph, 0), xn = load_X<T>(P, ph, n);
            cpx<T> z; z.x = x0.x + xn.x; z.y = x0.x - xn.x;
            buf[digit_reverse(P, 0)] = z;
        } else {
            const cpx<T> xa = load_X<T>(P, ph, k), xb = load_X<T>(P, ph, n - k);
            const cpx<T> wa = twid<T, +1>(P, k);
            const cpx<T> ea = cadd(xa, cconj(xb)), oa = cmul(csub(xa, cconj(xb)), wa);
            buf[digit_reverse(P, k)] = cadd(ea, cmuli<T, +1>(oa));
            if (k != n - k) {
                const cpx<T> wb = twid<T, +1>(P, n - k);
                const cpx<T> eb = cadd(xb, cconj(xa)), ob = cmul(csub(xb, cconj(xa)), wb);
                buf[digit_reverse(P, n - k)] = cadd(eb, cmuli<T, +1>(ob));
            }
        }
    }
    __syncthreads();
    fft_passes<T, +1>(P, buf);

    // store x[2j] = Re z[j], x[2j+1] = Im z[j] into the caller's array (flips / partial rings by index arithmetic)
    T* out = reinterpret_cast<T*>(P.maps[c]);
    const int rowy = P.flipy ? (P.ny - 1 - ring) : ring;
    T* orow = out + (size_t)rowy * P.nx;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const cpx<T> z = buf[j];
        const int i0 = 2 * j, i1 = 2 * j + 1;
        if (i0 < P.nx) orow[P.flipx ? (P.nx - 1 - i0) : i0] = z.x;
        if (i1 < P.nx) orow[P.flipx ? (P.nx - 1 - i1) : i1] = z.y;
    }
}

// map -> weighted phase  (analysis).  grid = (ring_count, ncomp)
template <class T>
__global__ void fft_map2phase(const FftParams P)
{
    PIXSHT_DYN_SMEM(smem_raw);
    cpx<T>* buf = reinterpret_cast<cpx<T>*>(smem_raw);   // n + 1 entries
    const int rl = blockIdx.x, c = blockIdx.y, ring = P.ring_begin + rl;
    const int n = P.n, N = P.nphi;
    const T* in = reinterpret_cast<const T*>(P.maps[c]);
    const int rowy = P.flipy ? (P.ny - 1 - ring) : ring;
    const T* irow = in + (size_t)rowy * P.nx;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const int i0 = 2 * j, i1 = 2 * j + 1;
        cpx<T> z;
        z.x = (i0 < P.nx) ? irow[P.flipx ? (P.nx - 1 - i0) : i0] : (T)0;
        z.y = (i1 < P.nx) ? irow[P.flipx ? (P.nx - 1 - i1) : i1] : (T)0;
        buf[digit_reverse(P, j)] = z;
    }
    __syncthreads();
    fft_passes<T, -1>(P, buf);

    // post-processing in place: F[k] = ((Z[k] + conj Z[n-k]) - i e^{-2 pi i k/N} (Z[k] - conj Z[n-k])) / 2,  k = 0..n
    for (int k = threadIdx.x; k <= n / 2; k += blockDim.x) {
        if (k == 0) {
            const cpx<T> z0 = buf[0];
            cpx<T> f0, fn; f0.x = z0.x + z0.y; f0.y = 0; fn.x = z0.x - z0.y; fn.y = 0;
            buf[0] = f0; buf[n] = fn;
        } else {
            const cpx<T> za = buf[k], zb = buf[n - k];
            const cpx<T> ea = cadd(za, cconj(zb)), oa = cmul(csub(za, cconj(zb)), twid<T, -1>(P, k));
            const cpx<T> fa = cadd(ea, cmuli<T, -1>(oa));
            cpx<T> r; r.x = (T)0.5 * fa.x; r.y = (T)0.5 * fa.y;
            if (k != n - k) {
                const cpx<T> eb = cadd(zb, cconj(za)), ob = cmul(csub(zb, cconj(za)), twid<T, -1>(P, n - k));
                const cpx<T> fb = cadd(eb, cmuli<T, -1>(ob));
                cpx<T> rb; rb.x = (T)0.5 * fb.x; rb.y = (T)0.5 * fb.y;
                buf[n - k] = rb;
            }
            buf[k] = r;
        }
    }
    __syncthreads();

    // phase_m = w * e^{-i m phi0} * F[m mod N]  (conjugate symmetric upper half)
    const double w = P.wgt[ring];
    double2* ph = P.phase + (long long)c * P.stride_c + rl;
    for (int m = threadIdx.x; m <= P.mmax; m += blockDim.x) {
        const int kk = m % N;
        cpx<T> f = (kk <= n) ? buf[kk] : cconj(buf[N - kk]);
        const double2 r = P.phi0tw[m];
        const double fx = (double)f.x, fy = (double)f.y;
        const int row = P.m_row ? P.m_row[m] : m;
        ph[(long long)row * P.stride_m] = make_double2(w * (fx * r.x + fy * r.y), w * (fy * r.x - fx * r.y));
    }
}

}  // namespace pixsht
