/*
 * oracle/sht_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle; never on the product path).
 *
 * A from-scratch CPU restatement of the arithmetic that simonsobs/Pixell.jl (v0.2.9) delegates to
 * libsharp2 (via Libsharp.jl 0.2 / libsharp2_jll; source NOT in /root/reference) and to
 * FastTransforms.jl (Clenshaw-Curtis weights), for the `map2alm` / `alm2map` hot path:
 *
 *   - ring weights            : src/transforms.jl:44-45  (clenshawcurtisweights(chebyshevjacobimoments1(N,0,0)) * 2pi/nphi)
 *   - ring colatitudes        : src/transforms.jl:46     (theta_k = range(0, pi, length=N)[k])
 *   - per-ring layout         : src/transforms.jl:49-53  (nph = const, stride 1, offsets = nph*k, phi0 const)
 *   - spin-0 analysis/synth.  : src/transforms.jl:101-106, 214-218 (sharp_execute!(SHARP_MAP2ALM / SHARP_ALM2MAP, 0, ...))
 *   - spin-2 analysis/synth.  : src/transforms.jl:128-132, 185-194, 240-244 (spin = 2, [E,B] <-> [Q,U])
 *   - alm layout              : triangular, m-major, idx(l,m) = m(2 lmax + 1 - m)/2 + l  (make_triangular_alm_info(lmax,mmax,1),
 *                               src/transforms.jl:94; Healpix.Alm)
 *
 * Mathematical spec: SURVEY.md Appendix A (libsharp2 / HEALPix conventions: orthonormal Y_lm with Condon-Shortley phase,
 * sY_lm = (-1)^m sqrt((2l+1)/4pi) d^l_{-m,s}(theta) e^{i m phi}, E = -(a+ + a-)/2, B = i(a+ - a-)/2).
 *
 * This oracle is deliberately NOT the algorithm of the CUDA engine: no north/south folding, no pruning, plain
 * normalised three-term recurrences, extended-range arithmetic by an explicit power-of-two exponent, and (in the
 * ORC_LONG build) 80-bit long double throughout.  It is pinned against the reference's own golden alm files
 * (test/data/simple_*.txt) by tests/test_oracle_golden.py.
 *
 * Build: see oracle/Makefile.  Two shared objects from this one file:
 *   liborc_ld.so (-DORC_LONG, REAL = long double)  -> the parity checker
 *   liborc_d.so  (REAL = double, OpenMP)           -> the timed CPU baseline ("port", not libsharp2)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef ORC_LONG
typedef long double REAL;
#define R_SQRT sqrtl
#define R_LDEXP ldexpl
#define R_FABS fabsl
#define R_SIN sinl
#define R_COS cosl
#define R_PI 3.14159265358979323846264338327950288L
#else
typedef double REAL;
#define R_SQRT sqrt
#define R_LDEXP ldexp
#define R_FABS fabs
#define R_SIN sin
#define R_COS cos
#define R_PI 3.14159265358979323846264338327950288
#endif

#define LD_PI 3.14159265358979323846264338327950288L

typedef struct { REAL re, im; } cplx;

/* ------------------------------------------------------------------------------------------------------------
 * Clenshaw-Curtis weights for n nodes x_k = cos(pi k/(n-1)) (both end points included), integrating on [-1,1].
 * Restates FastTransforms.clenshawcurtisweights(chebyshevjacobimoments1(Float64, n, 0, 0)) as used at
 * src/transforms.jl:44-45; closed form of SURVEY.md A.2.
 * ---------------------------------------------------------------------------------------------------------- */
void orc_cc_weights(int n, double *w)
{
    if (n == 1) { w[0] = 2.0; return; }
    int nn = n - 1;
    for (int k = 0; k < n; ++k) {
        long double s = 0.0L;
        for (int q = 1; q <= nn / 2; ++q) {
            long double b = (2 * q == nn) ? 1.0L : 2.0L;
            /* cos(2 q k pi / nn) with exact integer argument reduction */
            long long t = ((long long)2 * q * k) % (2LL * nn);
            s += b / (4.0L * q * q - 1.0L) * cosl(LD_PI * (long double)t / (long double)nn);
        }
        long double g = (k == 0 || k == nn) ? 1.0L : 2.0L;
        w[k] = (double)(g / nn * (1.0L - s));
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * lambda generators.  out arrays are indexed by l (0..lmax); entries below the starting degree are 0.
 *   s = 0      : lambda_lm(theta) = Y_lm(theta, 0)
 *   s = +-2    : (-1)^m sqrt((2l+1)/4pi) d^l_{-m,s}(theta)
 * Three-term recurrence in l, values carried as v * 2^e to survive sin^m(theta) ~ 1e-38000.
 * ---------------------------------------------------------------------------------------------------------- */
static long double log2_factorial(int n) { return lgammal((long double)n + 1.0L) / logl(2.0L); }

/* A_l of lambda_{l+1} = A_l (x - mu_l) lambda_l - (A_l / A_{l-1}) lambda_{l-1} */
static REAL coefA(int l, int m, int s)
{
    REAL l1 = (REAL)(l + 1);
    REAL num = (REAL)(2 * l + 1) * (REAL)(2 * l + 3);
    REAL den = (l1 * l1 - (REAL)m * (REAL)m) * (l1 * l1 - (REAL)s * (REAL)s);
    return l1 * R_SQRT(num / den);
}

/* fills out[l], l = 0..lmax */
void orc_lambda(int lmax, int m, int s, double theta_in, REAL *out)
{
    for (int l = 0; l <= lmax; ++l) out[l] = 0;
    int as = s < 0 ? -s : s;
    int l0 = m > as ? m : as;
    if (l0 > lmax) return;

    long double theta = (long double)theta_in;
    long double ch = cosl(0.5L * theta), sh = sinl(0.5L * theta);
    /* exact mirror handling so that theta and pi - theta give bit-mirrored half-angle values is not needed here */
    REAL x = (REAL)cosl(theta);

    /* seed: lambda_{l0} = (-1)^m sqrt((2 l0+1)/4pi) * sqrt((2 l0)!/((l0+q)!(l0-q)!)) cos^a(theta/2) sin^b(theta/2) * sgn
     * with (a, b, q, sgn):
     *   m >= |s|: l0 = m, q = |s|, (a,b) = (m - s, m + s), sgn = +1
     *   m <  |s|: l0 = |s|, q = m, s>0: (a,b) = (|s| - m, |s| + m), sgn = +1;  s<0: (a,b) = (|s| + m, |s| - m), sgn = (-1)^m
     * (s = 0 reduces to the usual lambda_mm = (-1)^m sqrt((2m+1)!!/(4pi (2m)!!)) sin^m theta.) */
    int a, b, q; int sgn = (m & 1) ? -1 : 1;
    if (m >= as) { q = as; a = m - s; b = m + s; }
    else { q = m; if (s > 0) { a = as - m; b = as + m; } else { a = as + m; b = as - m; if (m & 1) sgn = -sgn; } }
    long double lg = 0.5L * (log2l((2.0L * l0 + 1.0L) / (4.0L * LD_PI)) + log2_factorial(2 * l0) - log2_factorial(l0 + q) -
                             log2_factorial(l0 - q));
    if (a > 0) { if (ch <= 0.0L) return; lg += a * log2l(ch); }
    if (b > 0) { if (sh <= 0.0L) return; lg += b * log2l(sh); }
    /* value = 2^lg ; split into e (multiple of 256, <= 0 unless value is large) and mantissa part */
    long e = 0;
    if (lg < -200.0L) { e = -256L * (long)ceill((-lg - 200.0L) / 256.0L); }
    REAL v = (REAL)exp2l(lg - (long double)e) * (REAL)sgn;
    REAL vp = 0;
    const REAL big = R_LDEXP((REAL)1, 256), small = R_LDEXP((REAL)1, -256);
#ifdef ORC_LONG
    const long elim = -16000;
#else
    const long elim = -1000;
#endif
    REAL Aprev = 1;
    for (int l = l0; l <= lmax; ++l) {
        out[l] = (e == 0) ? v : (e >= elim ? R_LDEXP(v, (int)e) : (REAL)0);
        if (l == lmax) break;
        REAL A = coefA(l, m, s);
        REAL mu = (s == 0 || l == 0) ? (REAL)0 : -(REAL)m * (REAL)s / ((REAL)l * (REAL)(l + 1));
        REAL vn = A * (x - mu) * v - (l > l0 ? (A / Aprev) * vp : (REAL)0);
        vp = v; v = vn; Aprev = A;
        if (e < 0 && R_FABS(v) > big) { v *= small; vp *= small; e += 256; }
    }
}

/* double-typed convenience wrapper for tests */
void orc_lambda_d(int lmax, int m, int s, double theta, double *out)
{
    REAL *t = (REAL *)malloc(sizeof(REAL) * (size_t)(lmax + 1));
    orc_lambda(lmax, m, s, theta, t);
    for (int l = 0; l <= lmax; ++l) out[l] = (double)t[l];
    free(t);
}

/* ------------------------------------------------------------------------------------------------------------
 * small generic complex FFT (recursive mixed radix, O(p^2) on prime factors) in REAL precision.
 * sign = -1: forward  X[k] = sum_p x[p] e^{-2 pi i k p / n};  sign = +1: backward (unnormalised).
 * ---------------------------------------------------------------------------------------------------------- */
static void twiddle(long long num, long long n, int sign, cplx *w)
{
    num %= n; if (num < 0) num += n;
    long double ang = 2.0L * LD_PI * (long double)num / (long double)n;
    w->re = (REAL)cosl(ang); w->im = (REAL)(sign * sinl(ang));
}

static void fft_rec(int n, int stride, const cplx *in, cplx *out, cplx *scratch, int sign, const cplx *tw, int ntw)
{
    if (n == 1) { out[0] = in[0]; return; }
    int p = n;
    for (int f = 2; (long long)f * f <= n; ++f) if (n % f == 0) { p = f; break; }
    int mlen = n / p;
    /* p sub-transforms of length mlen over decimated input */
    for (int j = 0; j < p; ++j) fft_rec(mlen, stride * p, in + (size_t)j * stride, scratch + (size_t)j * mlen, out, sign, tw, ntw);
    /* combine: X[k + t mlen] = sum_j W_p^{jt} W_n^{jk} Y_j[k] */
    int tstep = ntw / n; /* tw has ntw entries of e^{sign 2 pi i q / ntw} */
    for (int k = 0; k < mlen; ++k) {
        for (int t = 0; t < p; ++t) {
            REAL sr = 0, si = 0;
            for (int j = 0; j < p; ++j) {
                long long idx = ((long long)j * (k + (long long)t * mlen)) % n;
                cplx w = tw[idx * tstep];
                cplx y = scratch[(size_t)j * mlen + k];
                sr += w.re * y.re - w.im * y.im;
                si += w.re * y.im + w.im * y.re;
            }
            out[k + (size_t)t * mlen].re = sr; out[k + (size_t)t * mlen].im = si;
        }
    }
}

typedef struct { int n; int sign; cplx *tw; cplx *buf_a; cplx *buf_b; } fftplan;

static void fftplan_init(fftplan *P, int n, int sign)
{
    P->n = n; P->sign = sign;
    P->tw = (cplx *)malloc(sizeof(cplx) * (size_t)n);
    for (int q = 0; q < n; ++q) twiddle(q, n, sign, &P->tw[q]);
    P->buf_a = (cplx *)malloc(sizeof(cplx) * (size_t)n);
    P->buf_b = (cplx *)malloc(sizeof(cplx) * (size_t)n);
}
static void fftplan_free(fftplan *P) { free(P->tw); free(P->buf_a); free(P->buf_b); }
/* in -> out (both length n, distinct) */
static void fft_exec(fftplan *P, const cplx *in, cplx *out)
{
    /* fft_rec needs: out, scratch distinct from in; recursion ping-pongs out/scratch */
    fft_rec(P->n, 1, in, out, P->buf_a, P->sign, P->tw, P->n);
}

/* ------------------------------------------------------------------------------------------------------------
 * Transforms at the libsharp2 boundary: rings in ascending theta, each ring nphi samples at
 * phi_p = phi0 + 2 pi p / nphi, maps[c][ring*nphi + p]; alms[c] interleaved (re,im) doubles, triangular m-major.
 * spin = 0: 1 map, 1 alm.  spin = 2: maps = {Q,U}, alms = {E,B}.
 * Sampling: only m with (m % m_stride == m_offset) are computed in map2alm (others left untouched);
 *           only rings with (ring % ring_stride == ring_offset) are computed in alm2map (others untouched).
 * Returns 0 on success.
 * ---------------------------------------------------------------------------------------------------------- */
static size_t alm_index(int lmax, int l, int m) { return (size_t)m * (size_t)(2 * lmax + 1 - m) / 2 + (size_t)l; }

int orc_map2alm(int spin, int nrings, const double *theta, const double *wgt, double phi0, int nphi, int lmax, int mmax,
                const double *const *maps, double *const *alms, int m_stride, int m_offset)
{
    if (!(spin == 0 || spin == 2) || mmax > lmax || m_stride < 1) return 1;
    const int nc = spin == 0 ? 1 : 2;
    int nsel = 0;
    for (int m = 0; m <= mmax; ++m) if (m % m_stride == m_offset) ++nsel;
    if (nsel == 0) return 0;
    int *msel = (int *)malloc(sizeof(int) * (size_t)nsel);
    { int i = 0; for (int m = 0; m <= mmax; ++m) if (m % m_stride == m_offset) msel[i++] = m; }

    /* stage A: phase[c][isel][ring] = w_ring * sum_p map e^{-i m phi_p} */
    cplx *phase = (cplx *)malloc(sizeof(cplx) * (size_t)nc * nsel * nrings);
    const int use_fft = ((double)nsel * 4.0 > 6.0 * log2((double)nphi + 1.0));
    /* e^{-i m phi0} */
    cplx *ph0 = (cplx *)malloc(sizeof(cplx) * (size_t)nsel);
    for (int i = 0; i < nsel; ++i) {
        long double ang = fmodl((long double)msel[i] * (long double)phi0, 2.0L * LD_PI);
        ph0[i].re = (REAL)cosl(ang); ph0[i].im = (REAL)(-sinl(ang));
    }
#pragma omp parallel
    {
        fftplan P; cplx *in = NULL, *out = NULL; cplx *tw = NULL;
        if (use_fft) { fftplan_init(&P, nphi, -1); in = (cplx *)malloc(sizeof(cplx) * (size_t)nphi); out = (cplx *)malloc(sizeof(cplx) * (size_t)nphi); }
        else { tw = (cplx *)malloc(sizeof(cplx) * (size_t)nphi); for (int q = 0; q < nphi; ++q) twiddle(q, nphi, -1, &tw[q]); }
#pragma omp for schedule(dynamic, 4) collapse(2)
        for (int c = 0; c < nc; ++c)
            for (int r = 0; r < nrings; ++r) {
                const double *ring = maps[c] + (size_t)r * nphi;
                if (use_fft) {
                    for (int p = 0; p < nphi; ++p) { in[p].re = (REAL)ring[p]; in[p].im = 0; }
                    fft_exec(&P, in, out);
                }
                for (int i = 0; i < nsel; ++i) {
                    int m = msel[i];
                    cplx F;
                    if (use_fft) F = out[m % nphi];
                    else {
                        REAL sr = 0, si = 0; int mm = m % nphi; long long idx = 0;
                        for (int p = 0; p < nphi; ++p) { sr += (REAL)ring[p] * tw[idx].re; si += (REAL)ring[p] * tw[idx].im; idx += mm; if (idx >= nphi) idx -= nphi; }
                        F.re = sr; F.im = si;
                    }
                    REAL w = (REAL)wgt[r];
                    cplx v; v.re = w * (F.re * ph0[i].re - F.im * ph0[i].im); v.im = w * (F.re * ph0[i].im + F.im * ph0[i].re);
                    phase[((size_t)c * nsel + i) * nrings + r] = v;
                }
            }
        if (use_fft) { fftplan_free(&P); free(in); free(out); } else free(tw);
    }

    /* stage B: Legendre sums, independent per m */
#pragma omp parallel
    {
        REAL *lam = (REAL *)malloc(sizeof(REAL) * (size_t)(lmax + 1) * 2);
        REAL *lamm = lam + (lmax + 1);
        cplx *acc = (cplx *)malloc(sizeof(cplx) * (size_t)(lmax + 1) * 2);
#pragma omp for schedule(dynamic, 1)
        for (int i = 0; i < nsel; ++i) {
            int m = msel[i];
            for (int l = 0; l < 2 * (lmax + 1); ++l) { acc[l].re = 0; acc[l].im = 0; }
            for (int r = 0; r < nrings; ++r) {
                if (spin == 0) {
                    orc_lambda(lmax, m, 0, theta[r], lam);
                    cplx ph = phase[(size_t)i * nrings + r];
                    for (int l = m; l <= lmax; ++l) { acc[l].re += lam[l] * ph.re; acc[l].im += lam[l] * ph.im; }
                } else {
                    orc_lambda(lmax, m, +2, theta[r], lam);
                    orc_lambda(lmax, m, -2, theta[r], lamm);
                    cplx q = phase[((size_t)0 * nsel + i) * nrings + r], u = phase[((size_t)1 * nsel + i) * nrings + r];
                    /* Y+ = Xq + i Xu ; Y- = Xq - i Xu */
                    cplx yp = { q.re - u.im, q.im + u.re }, ym = { q.re + u.im, q.im - u.re };
                    int l0 = m > 2 ? m : 2;
                    for (int l = l0; l <= lmax; ++l) {
                        acc[l].re += lam[l] * yp.re; acc[l].im += lam[l] * yp.im;                                  /* a+ */
                        acc[lmax + 1 + l].re += lamm[l] * ym.re; acc[lmax + 1 + l].im += lamm[l] * ym.im;          /* a- */
                    }
                }
            }
            if (spin == 0) {
                for (int l = m; l <= lmax; ++l) { size_t k = alm_index(lmax, l, m); alms[0][2 * k] = (double)acc[l].re; alms[0][2 * k + 1] = (double)acc[l].im; }
            } else {
                for (int l = m; l <= lmax; ++l) {
                    size_t k = alm_index(lmax, l, m);
                    cplx ap = acc[l], am = acc[lmax + 1 + l];
                    /* E = -(a+ + a-)/2 ; B = i (a+ - a-)/2 */
                    alms[0][2 * k] = (double)(-(ap.re + am.re) / 2); alms[0][2 * k + 1] = (double)(-(ap.im + am.im) / 2);
                    alms[1][2 * k] = (double)(-(ap.im - am.im) / 2); alms[1][2 * k + 1] = (double)((ap.re - am.re) / 2);
                }
            }
        }
        free(lam); free(acc);
    }
    free(phase); free(ph0); free(msel);
    return 0;
}

int orc_alm2map(int spin, int nrings, const double *theta, double phi0, int nphi, int lmax, int mmax,
                const double *const *alms, double *const *maps, int ring_stride, int ring_offset)
{
    if (!(spin == 0 || spin == 2) || mmax > lmax || ring_stride < 1) return 1;
    const int nc = spin == 0 ? 1 : 2;
    int nsel = 0;
    for (int r = 0; r < nrings; ++r) if (r % ring_stride == ring_offset) ++nsel;
    if (nsel == 0) return 0;
    int *rsel = (int *)malloc(sizeof(int) * (size_t)nsel);
    { int i = 0; for (int r = 0; r < nrings; ++r) if (r % ring_stride == ring_offset) rsel[i++] = r; }

    /* stage B: phase[c][m][isel] */
    cplx *phase = (cplx *)calloc((size_t)nc * (mmax + 1) * nsel, sizeof(cplx));
#pragma omp parallel
    {
        REAL *lam = (REAL *)malloc(sizeof(REAL) * (size_t)(lmax + 1) * 2);
        REAL *lamm = lam + (lmax + 1);
#pragma omp for schedule(dynamic, 1)
        for (int m = 0; m <= mmax; ++m) {
            for (int i = 0; i < nsel; ++i) {
                double th = theta[rsel[i]];
                if (spin == 0) {
                    orc_lambda(lmax, m, 0, th, lam);
                    REAL sr = 0, si = 0;
                    for (int l = m; l <= lmax; ++l) { size_t k = alm_index(lmax, l, m); sr += lam[l] * (REAL)alms[0][2 * k]; si += lam[l] * (REAL)alms[0][2 * k + 1]; }
                    phase[(size_t)m * nsel + i].re = sr; phase[(size_t)m * nsel + i].im = si;
                } else {
                    orc_lambda(lmax, m, +2, th, lam);
                    orc_lambda(lmax, m, -2, th, lamm);
                    /* Z+ = -sum (E + iB) lam+ ; Z- = -sum (E - iB) lam- ; q = (Z+ + Z-)/2 ; u = -i (Z+ - Z-)/2 */
                    REAL zpr = 0, zpi = 0, zmr = 0, zmi = 0;
                    int l0 = m > 2 ? m : 2;
                    for (int l = l0; l <= lmax; ++l) {
                        size_t k = alm_index(lmax, l, m);
                        REAL er = (REAL)alms[0][2 * k], ei = (REAL)alms[0][2 * k + 1], br = (REAL)alms[1][2 * k], bi = (REAL)alms[1][2 * k + 1];
                        zpr -= (er - bi) * lam[l]; zpi -= (ei + br) * lam[l];
                        zmr -= (er + bi) * lamm[l]; zmi -= (ei - br) * lamm[l];
                    }
                    cplx q = { (zpr + zmr) / 2, (zpi + zmi) / 2 };
                    cplx u = { (zpi - zmi) / 2, -(zpr - zmr) / 2 };
                    phase[((size_t)0 * (mmax + 1) + m) * nsel + i] = q;
                    phase[((size_t)1 * (mmax + 1) + m) * nsel + i] = u;
                }
            }
        }
        free(lam);
    }

    /* stage A: ring[p] = Re sum_m c_m (m==0 ? 1 : 2) e^{i m phi_p}; aliasing handled by folding m mod nphi */
    cplx *ph0 = (cplx *)malloc(sizeof(cplx) * (size_t)(mmax + 1));
    for (int m = 0; m <= mmax; ++m) {
        long double ang = fmodl((long double)m * (long double)phi0, 2.0L * LD_PI);
        ph0[m].re = (REAL)cosl(ang); ph0[m].im = (REAL)sinl(ang);
    }
#pragma omp parallel
    {
        fftplan P; fftplan_init(&P, nphi, +1);
        cplx *in = (cplx *)malloc(sizeof(cplx) * (size_t)nphi), *out = (cplx *)malloc(sizeof(cplx) * (size_t)nphi);
#pragma omp for schedule(dynamic, 4) collapse(2)
        for (int c = 0; c < nc; ++c)
            for (int i = 0; i < nsel; ++i) {
                for (int p = 0; p < nphi; ++p) { in[p].re = 0; in[p].im = 0; }
                for (int m = 0; m <= mmax; ++m) {
                    cplx a = phase[((size_t)c * (mmax + 1) + m) * nsel + i];
                    cplx v = { a.re * ph0[m].re - a.im * ph0[m].im, a.re * ph0[m].im + a.im * ph0[m].re };
                    if (m == 0) { in[0].re += v.re; }
                    else {
                        int kp = m % nphi, kn = (nphi - kp) % nphi;
                        in[kp].re += v.re; in[kp].im += v.im;
                        in[kn].re += v.re; in[kn].im -= v.im;
                    }
                }
                fft_exec(&P, in, out);
                double *ring = maps[c] + (size_t)rsel[i] * nphi;
                for (int p = 0; p < nphi; ++p) ring[p] = (double)out[p].re;
            }
        fftplan_free(&P); free(in); free(out);
    }
    free(phase); free(ph0); free(rsel);
    return 0;
}

int orc_real_bits(void) { return (int)(sizeof(REAL) * 8); }
int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
