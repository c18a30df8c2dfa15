/*
 * oracle/sht_cpu.c -- TEST INFRASTRUCTURE / TIMED CPU BASELINE ONLY (never on the product path).
 *
 * A CPU implementation of the map2alm / alm2map hot path with the ALGORITHM of libsharp2, the library that
 * simonsobs/Pixell.jl (v0.2.9) calls at src/transforms.jl:101-106,128-132,185-194,214-218,240-244 (its source is not under
 * /root/reference; SURVEY.md A.5 lists what is restated here): north/south ring pairing by equatorial symmetry, lambda_lm
 * generated on the fly per m by the three-term recurrence with a scaled-exponent seek from l = m, m-limit pruning
 * (m > lmax sin(theta) + max(100, 0.01 lmax)), ring weights folded into the phase, l-dependent normalisation folded into
 * the alm, OpenMP over m, the ring loop written for the compiler's SIMD vectoriser, real ring FFTs through a half-length
 * complex mixed-radix FFT.  Double precision throughout, as SHARP_DP.
 *
 * Why it exists: oracle/sht_oracle.c is a deliberately naive checker (no folding, no pruning, long double) and is ~50x
 * slower than a production CPU SHT; timing IT as "the CPU reference" would flatter the GPU numbers.  This file is the
 * honest CPU arm of bench.py (cpu_baseline.kind = "port": libsharp2-style restatement, not libsharp2 itself) and a second,
 * algorithmically independent check of the CUDA engine (tests/test_oracle_properties.py compares it with the checker).
 *
 * Conventions as in sht_oracle.c: rings ascending in theta, ring r = nphi samples at phi0 + 2 pi p/nphi stored at
 * maps[c][r*nphi + p]; alms[c] interleaved (re,im), triangular m-major; spin 0: 1 map/1 alm; spin 2: {Q,U} / {E,B}.
 * Sampling: only m with m % m_stride == m_offset take part (both directions; synthesis then yields the map of those m
 * alone, analysis leaves the other alm untouched) -- bench.py times a bounded sample and extrapolates the Legendre part.
 * times[0] = seconds in the Legendre stage, times[1] = seconds in the FFT stage.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI_L 3.14159265358979323846264338327950288L
typedef struct { double re, im; } cpx;

static double now_s(void)
{
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return 0.0;
#endif
}
static size_t alm_idx(int lmax, int l, int m) { return (size_t)m * (size_t)(2 * lmax + 1 - m) / 2 + (size_t)l; }

/* ---------------------------------------------------------------------------------------------------------------
 * complex FFT: Stockham autosort, mixed radix (any factor; O(p^2) butterflies), out of place ping-pong
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct { int n, nfac, fac[32]; cpx *tw; } fft_plan;   /* tw[t] = exp(-2 pi i t / n) */

static void fft_plan_init(fft_plan *P, int n)
{
    P->n = n; P->nfac = 0;
    int r = n;
    while (r % 4 == 0) { P->fac[P->nfac++] = 4; r /= 4; }
    for (int p = 2; r > 1; ++p) while (r % p == 0) { P->fac[P->nfac++] = p; r /= p; }
    P->tw = (cpx *)malloc(sizeof(cpx) * (size_t)n);
    for (int t = 0; t < n; ++t) { long double a = 2.0L * PI_L * t / n; P->tw[t].re = (double)cosl(a); P->tw[t].im = (double)(-sinl(a)); }
}
static void fft_plan_free(fft_plan *P) { free(P->tw); }

/* x -> result in x (scratch y); sign = -1 forward, +1 backward (unnormalised) */
static void fft_run(const fft_plan *P, cpx *x, cpx *y, int sign)
{
    const int n = P->n;
    int nn = n, s = 1;
    cpx *a = x, *b = y;
    for (int f = 0; f < P->nfac; ++f) {
        const int p = P->fac[f], m = nn / p;
        const int tstep = n / nn;   /* exp(-2 pi i j/nn) = tw[j * tstep] */
        cpx wp[32 * 32];            /* W_p^{r u} */
        if (p != 2 && p != 4)
            for (int u = 0; u < p; ++u) for (int r = 0; r < p; ++r) { cpx w = P->tw[(size_t)((r * u) % p) * (n / p)]; wp[u * p + r] = (cpx){ w.re, sign < 0 ? w.im : -w.im }; }
        for (int j = 0; j < m; ++j) {
            cpx wj[32];             /* output twiddles exp(sign 2 pi i j u / nn) */
            for (int u = 1; u < p; ++u) { cpx w = P->tw[(size_t)((long long)j * u % nn) * tstep]; wj[u] = (cpx){ w.re, sign < 0 ? w.im : -w.im }; }
            const cpx *aj = a + (size_t)s * j;
            cpx *bj = b + (size_t)s * p * j;
            if (p == 4) {
                for (int q = 0; q < s; ++q) {
                    const cpx i0 = aj[q], i1 = aj[q + (size_t)s * m], i2 = aj[q + (size_t)s * 2 * m], i3 = aj[q + (size_t)s * 3 * m];
                    const cpx s02 = { i0.re + i2.re, i0.im + i2.im }, d02 = { i0.re - i2.re, i0.im - i2.im };
                    const cpx s13 = { i1.re + i3.re, i1.im + i3.im }, d13 = { i1.re - i3.re, i1.im - i3.im };
                    const cpx jd = sign < 0 ? (cpx){ d13.im, -d13.re } : (cpx){ -d13.im, d13.re };   /* -+ i d13 */
                    const cpx o1 = { d02.re + jd.re, d02.im + jd.im }, o2 = { s02.re - s13.re, s02.im - s13.im }, o3 = { d02.re - jd.re, d02.im - jd.im };
                    bj[q] = (cpx){ s02.re + s13.re, s02.im + s13.im };
                    bj[q + s] = (cpx){ o1.re * wj[1].re - o1.im * wj[1].im, o1.re * wj[1].im + o1.im * wj[1].re };
                    bj[q + 2 * s] = (cpx){ o2.re * wj[2].re - o2.im * wj[2].im, o2.re * wj[2].im + o2.im * wj[2].re };
                    bj[q + 3 * s] = (cpx){ o3.re * wj[3].re - o3.im * wj[3].im, o3.re * wj[3].im + o3.im * wj[3].re };
                }
            } else if (p == 2) {
                for (int q = 0; q < s; ++q) {
                    const cpx i0 = aj[q], i1 = aj[q + (size_t)s * m];
                    const cpx o1 = { i0.re - i1.re, i0.im - i1.im };
                    bj[q] = (cpx){ i0.re + i1.re, i0.im + i1.im };
                    bj[q + s] = (cpx){ o1.re * wj[1].re - o1.im * wj[1].im, o1.re * wj[1].im + o1.im * wj[1].re };
                }
            } else {
                for (int q = 0; q < s; ++q) {
                    cpx in[32];
                    for (int r = 0; r < p; ++r) in[r] = aj[q + (size_t)s * r * m];
                    for (int u = 0; u < p; ++u) {
                        double sr = in[0].re, si = in[0].im;
                        for (int r = 1; r < p; ++r) { const cpx w = wp[u * p + r]; sr += in[r].re * w.re - in[r].im * w.im; si += in[r].re * w.im + in[r].im * w.re; }
                        if (u == 0) bj[q] = (cpx){ sr, si };
                        else bj[q + (size_t)u * s] = (cpx){ sr * wj[u].re - si * wj[u].im, sr * wj[u].im + si * wj[u].re };
                    }
                }
            }
        }
        cpx *t = a; a = b; b = t;
        nn = m; s *= p;
    }
    if (a != x) memcpy(x, a, sizeof(cpx) * (size_t)n);
}

/* ---------------------------------------------------------------------------------------------------------------
 * geometry: ring pairs, recurrence coefficients, seeds
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct { int np; int *rn, *rs; double *x, *mlim; long double *lsh, *lch; } pairs_t;

static void pairs_build(pairs_t *P, int nr, const double *theta, int lmax)
{
    /* theta ascending: walk in from both ends; rings whose cosines cancel form a pair */
    P->rn = (int *)malloc(sizeof(int) * nr); P->rs = (int *)malloc(sizeof(int) * nr);
    P->x = (double *)malloc(sizeof(double) * nr); P->mlim = (double *)malloc(sizeof(double) * nr);
    P->lsh = (long double *)malloc(sizeof(long double) * nr); P->lch = (long double *)malloc(sizeof(long double) * nr);
    int lo = 0, hi = nr - 1, np = 0;
    const double ofs = fmax(100.0, 0.01 * lmax) + 4.0;
    while (lo <= hi) {
        long double tn, xa = cosl((long double)theta[lo]), xb = cosl((long double)theta[hi]);
        if (lo < hi && fabsl(xa + xb) < 1e-12L && xa > 0) { P->rn[np] = lo; P->rs[np] = hi; tn = theta[lo]; ++lo; --hi; }
        else if (xa >= 0 && (lo == hi || xa + xb > 0 || xb >= 0)) { P->rn[np] = lo; P->rs[np] = -1; tn = theta[lo]; ++lo; }
        else { P->rn[np] = -1; P->rs[np] = hi; tn = PI_L - (long double)theta[hi]; --hi; }
        P->x[np] = (double)cosl(tn);
        long double sh = sinl(0.5L * tn), ch = cosl(0.5L * tn);
        P->lsh[np] = sh > 0 ? log2l(sh) : -HUGE_VALL; P->lch[np] = ch > 0 ? log2l(ch) : -HUGE_VALL;
        P->mlim[np] = (double)lmax * (double)sinl(tn) + ofs;
        ++np;
    }
    P->np = np;
}
static void pairs_free(pairs_t *P) { free(P->rn); free(P->rs); free(P->x); free(P->mlim); free(P->lsh); free(P->lch); }

static double coefA(int l, int m, int s)
{
    const double l1 = l + 1.0;
    return l1 * sqrt((2.0 * l + 1.0) * (2.0 * l + 3.0) / (((l1 - m) * (l1 + m)) * ((l1 - s) * (l1 + s))));
}
/* lambda_l = gamma_l p_l,  p_{l+1} = (alpha_l x + delta_l) p_l - p_{l-1}  (delta -> -delta for the s = -2 family) */
static void coefs(int lmax, int m, int s, double *alpha, double *delta, double *gamma)
{
    const int l0 = m > s ? m : s;
    if (l0 > lmax) return;
    const double ms = (double)m * s;
    double Aprev = coefA(l0, m, s), g_lm1 = 1.0, g_l = 1.0;
    gamma[l0] = 1.0; alpha[l0] = Aprev; delta[l0] = l0 > 0 ? Aprev * ms / ((double)l0 * (l0 + 1.0)) : 0.0;
    for (int l = l0 + 1; l <= lmax; ++l) {
        const double A = coefA(l, m, s), g_lp1 = (A / Aprev) * g_lm1, a = A * g_l / g_lp1;
        gamma[l] = g_l; alpha[l] = a; delta[l] = a * ms / ((double)l * (l + 1.0));
        g_lm1 = g_l; g_l = g_lp1; Aprev = A;
    }
}
/* log2 of the seed prefactors for all m (running sums) */
static void prefactors(int mmax, long double *lg0, long double *lg2)
{
    long double acc = 0, cacc = 0;
    for (int m = 0; m <= mmax; ++m) {
        if (m > 0) acc += log2l((2.0L * m - 1.0L) / (2.0L * m));
        lg0[m] = 0.5L * (log2l((2.0L * m + 1.0L) / (4.0L * PI_L)) + acc) + (long double)m;
        if (m < 2) lg2[m] = 0.5L * (log2l(5.0L / (4.0L * PI_L)) + log2l(m == 0 ? 6.0L : 4.0L));
        else { if (m > 2) cacc += log2l((2.0L * m) * (2.0L * m - 1.0L) / ((m + 2.0L) * (m - 2.0L))); lg2[m] = 0.5L * (log2l((2.0L * m + 1.0L) / (4.0L * PI_L)) + cacc); }
    }
}

#define ACT_LOG2 (-90)
#define THR 1.4901161193847656e-08 /* 2^-26 */
#define RESC 5.421010862427522e-20 /* 2^-64 */
#define L_NEVER 0x3fffffff

/* scaled-exponent seek of one (m, pair): first l at which the function reaches 2^-90 and the state (p, p_prev) there */
static int seek(int spin, int lmax, int m, const pairs_t *P, int k, long double lgpref, const double *alpha, const double *delta,
                double *p0, double *q0, double *p1, double *q1)
{
    *p0 = *q0 = *p1 = *q1 = 0.0;
    const int l0 = spin == 0 ? m : (m > 2 ? m : 2);
    if ((double)m > P->mlim[k] || l0 > lmax) return L_NEVER;
    const double sgn = (m & 1) ? -1.0 : 1.0;
    long double la, lb; int za = 0, zb = 0;
    if (spin == 0) {
        if (m > 0 && !(P->lsh[k] > -1e300L && P->lch[k] > -1e300L)) return L_NEVER;
        la = lgpref + (m > 0 ? m * (P->lsh[k] + P->lch[k]) : 0.0L); lb = la; zb = 1;
    } else {
        const int am = m >= 2 ? m - 2 : 2 - m;
        za = (am > 0 && !(P->lch[k] > -1e300L)) || !(P->lsh[k] > -1e300L);
        zb = !(P->lch[k] > -1e300L) || (am > 0 && !(P->lsh[k] > -1e300L));
        la = za ? 0 : lgpref + (am > 0 ? am * P->lch[k] : 0.0L) + (m + 2) * P->lsh[k];
        lb = zb ? 0 : lgpref + (m + 2) * P->lch[k] + (am > 0 ? am * P->lsh[k] : 0.0L);
        if (za && zb) return L_NEVER;
    }
    long double kmax = (spin == 0) ? la : (za ? lb : (zb ? la : fmaxl(la, lb)));
    int e = 0;
    if (kmax < ACT_LOG2) e = -64 * (int)ceill((ACT_LOG2 - kmax) / 64.0L);
    if (spin == 0) *p0 = (la - e < -1000) ? 0.0 : sgn * (double)exp2l(la - e);
    else {
        if (!za) *p0 = (la - e < -1000) ? 0.0 : sgn * (double)exp2l(la - e);
        if (!zb) *p1 = (lb - e < -1000) ? 0.0 : ((m >= 2) ? sgn : 1.0) * (double)exp2l(lb - e);
    }
    int l = l0;
    const double x = P->x[k];
    while (e < 0 && l <= lmax) {
        if (spin == 0) {
            const double pn = alpha[l] * x * (*p0) - *q0; *q0 = *p0; *p0 = pn;
            if (fabs(pn) >= THR) { *p0 *= RESC; *q0 *= RESC; e += 64; }
        } else {
            const double pn = (alpha[l] * x + delta[l]) * (*p0) - *q0, mn = (alpha[l] * x - delta[l]) * (*p1) - *q1;
            *q0 = *p0; *p0 = pn; *q1 = *p1; *p1 = mn;
            if (fabs(pn) >= THR || fabs(mn) >= THR) { *p0 *= RESC; *q0 *= RESC; *p1 *= RESC; *q1 *= RESC; e += 64; }
        }
        ++l;
    }
    return (e == 0 && l <= lmax) ? l : L_NEVER;
}

/* ---------------------------------------------------------------------------------------------------------------
 * Legendre stage for one m over blocks of NB ring pairs (the inner ring loops are what the SIMD vectoriser works on)
 * phase layout: ph[c][(size_t)isel * nrings + ring]
 * ------------------------------------------------------------------------------------------------------------- */
#define NB 32

static void legendre_m(int dir_synth, int spin, int lmax, int m, const pairs_t *P, long double lgpref, int nrings, int isel, int nsel,
                       const double *const *alms, double *const *alm_out, cpx *const *ph, double *alpha, double *delta, double *gamma)
{
    const int s = spin, l0 = s == 0 ? m : (m > 2 ? m : 2);
    if (l0 > lmax) return;
    coefs(lmax, m, s, alpha, delta, gamma);
    const size_t ab = alm_idx(lmax, 0, m);
    (void)nsel;
    /* analysis accumulators over all blocks */
    double *a0r = NULL, *a0i = NULL, *a1r = NULL, *a1i = NULL;
    if (!dir_synth) {
        a0r = (double *)calloc((size_t)(lmax + 1) * 4, sizeof(double)); a0i = a0r + (lmax + 1); a1r = a0i + (lmax + 1); a1i = a1r + (lmax + 1);
    }
    const double bs = ((l0 + m) & 1) ? -1.0 : 1.0;
    for (int kb = 0; kb < P->np; kb += NB) {
        const int nb = P->np - kb < NB ? P->np - kb : NB;
        double x[NB], p0[NB], q0[NB], p1[NB], q1[NB]; int la[NB]; int lmin = L_NEVER;
        for (int r = 0; r < NB; ++r) { x[r] = 0; p0[r] = q0[r] = p1[r] = q1[r] = 0; la[r] = L_NEVER; }
        for (int r = 0; r < nb; ++r) {
            la[r] = seek(spin, lmax, m, P, kb + r, lgpref, alpha, delta, &p0[r], &q0[r], &p1[r], &q1[r]);
            x[r] = P->x[kb + r];
            if (la[r] < lmin) lmin = la[r];
        }
        if (lmin > lmax) {
            if (dir_synth) for (int r = 0; r < nb; ++r) for (int c = 0; c < (s == 0 ? 1 : 2); ++c) {
                if (P->rn[kb + r] >= 0) ph[c][(size_t)isel * nrings + P->rn[kb + r]] = (cpx){0, 0};
                if (P->rs[kb + r] >= 0) ph[c][(size_t)isel * nrings + P->rs[kb + r]] = (cpx){0, 0};
            }
            continue;
        }
        if (s == 0) {
            double er[NB], ei[NB], orr[NB], oi[NB], XeR[NB], XeI[NB], XoR[NB], XoI[NB];
            for (int r = 0; r < NB; ++r) { er[r] = ei[r] = orr[r] = oi[r] = 0; XeR[r] = XeI[r] = XoR[r] = XoI[r] = 0; }
            if (!dir_synth) for (int r = 0; r < nb; ++r) {
                cpx n = {0, 0}, so = {0, 0};
                if (la[r] != L_NEVER) { if (P->rn[kb + r] >= 0) n = ph[0][(size_t)isel * nrings + P->rn[kb + r]]; if (P->rs[kb + r] >= 0) so = ph[0][(size_t)isel * nrings + P->rs[kb + r]]; }
                XeR[r] = n.re + so.re; XeI[r] = n.im + so.im; XoR[r] = n.re - so.re; XoI[r] = n.im - so.im;
            }
            for (int l = lmin; l <= lmax; ++l) {
                const double al = alpha[l]; const int odd = (l - m) & 1;
                if (dir_synth) {
                    const double gr = gamma[l] * alms[0][2 * (ab + l)], gi = (m == 0) ? 0.0 : gamma[l] * alms[0][2 * (ab + l) + 1];
                    double *ar = odd ? orr : er, *ai = odd ? oi : ei;
#pragma omp simd
                    for (int r = 0; r < NB; ++r) {
                        const double on = (l >= la[r]) ? 1.0 : 0.0, pc = p0[r] * on;
                        ar[r] += pc * gr; ai[r] += pc * gi;
                        const double pn = al * x[r] * p0[r] - q0[r];
                        q0[r] = on != 0.0 ? p0[r] : q0[r]; p0[r] = on != 0.0 ? pn : p0[r];
                    }
                } else {
                    const double *Xr = odd ? XoR : XeR, *Xi = odd ? XoI : XeI;
                    double sr = 0, si = 0;
#pragma omp simd reduction(+ : sr, si)
                    for (int r = 0; r < NB; ++r) {
                        const double on = (l >= la[r]) ? 1.0 : 0.0, pc = p0[r] * on;
                        sr += pc * Xr[r]; si += pc * Xi[r];
                        const double pn = al * x[r] * p0[r] - q0[r];
                        q0[r] = on != 0.0 ? p0[r] : q0[r]; p0[r] = on != 0.0 ? pn : p0[r];
                    }
                    a0r[l] += sr; a0i[l] += si;
                }
            }
            if (dir_synth) for (int r = 0; r < nb; ++r) {
                if (P->rn[kb + r] >= 0) ph[0][(size_t)isel * nrings + P->rn[kb + r]] = (cpx){ er[r] + orr[r], ei[r] + oi[r] };
                if (P->rs[kb + r] >= 0) ph[0][(size_t)isel * nrings + P->rs[kb + r]] = (cpx){ er[r] - orr[r], ei[r] - oi[r] };
            }
        } else {
            /* acc: north S+ , S- ; south T+, T-  (re, im each) ;  X: Y+_N, Y+_S, Y-_N, Y-_S */
            double acc[8][NB], X[8][NB];
            for (int a = 0; a < 8; ++a) for (int r = 0; r < NB; ++r) { acc[a][r] = 0; X[a][r] = 0; }
            if (!dir_synth) for (int r = 0; r < nb; ++r) {
                cpx qn = {0, 0}, qs = {0, 0}, un = {0, 0}, us = {0, 0};
                if (la[r] != L_NEVER) {
                    if (P->rn[kb + r] >= 0) { qn = ph[0][(size_t)isel * nrings + P->rn[kb + r]]; un = ph[1][(size_t)isel * nrings + P->rn[kb + r]]; }
                    if (P->rs[kb + r] >= 0) { qs = ph[0][(size_t)isel * nrings + P->rs[kb + r]]; us = ph[1][(size_t)isel * nrings + P->rs[kb + r]]; }
                }
                X[0][r] = qn.re - un.im; X[1][r] = qn.im + un.re; X[2][r] = bs * (qs.re - us.im); X[3][r] = bs * (qs.im + us.re);
                X[4][r] = qn.re + un.im; X[5][r] = qn.im - un.re; X[6][r] = bs * (qs.re + us.im); X[7][r] = bs * (qs.im - us.re);
            }
            const int lstart = l0 + ((lmin - l0) & ~1);
            for (int l = lstart; l <= lmax; ++l) {
                const double al = alpha[l], de = delta[l], sg = ((l - l0) & 1) ? -1.0 : 1.0;
                if (dir_synth) {
                    const double h = -0.5 * gamma[l];
                    const double Er = alms[0][2 * (ab + l)], Ei = (m == 0) ? 0.0 : alms[0][2 * (ab + l) + 1];
                    const double Br = alms[1][2 * (ab + l)], Bi = (m == 0) ? 0.0 : alms[1][2 * (ab + l) + 1];
                    const double gpr = h * (Er - Bi), gpi = h * (Ei + Br), gmr = h * (Er + Bi), gmi = h * (Ei - Br);
#pragma omp simd
                    for (int r = 0; r < NB; ++r) {
                        const double on = (l >= la[r]) ? 1.0 : 0.0, a = p0[r] * on, b = p1[r] * on;
                        acc[0][r] += a * gpr; acc[1][r] += a * gpi; acc[2][r] += b * gmr; acc[3][r] += b * gmi;
                        acc[4][r] += sg * b * gpr; acc[5][r] += sg * b * gpi; acc[6][r] += sg * a * gmr; acc[7][r] += sg * a * gmi;
                        const double pn = (al * x[r] + de) * p0[r] - q0[r], mn = (al * x[r] - de) * p1[r] - q1[r];
                        q0[r] = on != 0.0 ? p0[r] : q0[r]; p0[r] = on != 0.0 ? pn : p0[r];
                        q1[r] = on != 0.0 ? p1[r] : q1[r]; p1[r] = on != 0.0 ? mn : p1[r];
                    }
                } else {
                    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma omp simd reduction(+ : s0, s1, s2, s3)
                    for (int r = 0; r < NB; ++r) {
                        const double on = (l >= la[r]) ? 1.0 : 0.0, a = p0[r] * on, b = p1[r] * on;
                        s0 += a * X[0][r] + sg * b * X[2][r]; s1 += a * X[1][r] + sg * b * X[3][r];
                        s2 += b * X[4][r] + sg * a * X[6][r]; s3 += b * X[5][r] + sg * a * X[7][r];
                        const double pn = (al * x[r] + de) * p0[r] - q0[r], mn = (al * x[r] - de) * p1[r] - q1[r];
                        q0[r] = on != 0.0 ? p0[r] : q0[r]; p0[r] = on != 0.0 ? pn : p0[r];
                        q1[r] = on != 0.0 ? p1[r] : q1[r]; p1[r] = on != 0.0 ? mn : p1[r];
                    }
                    a0r[l] += s0; a0i[l] += s1; a1r[l] += s2; a1i[l] += s3;
                }
            }
            if (dir_synth) for (int r = 0; r < nb; ++r) {
                const int rn = P->rn[kb + r], rs = P->rs[kb + r];
                if (rn >= 0) {
                    ph[0][(size_t)isel * nrings + rn] = (cpx){ acc[0][r] + acc[2][r], acc[1][r] + acc[3][r] };
                    ph[1][(size_t)isel * nrings + rn] = (cpx){ acc[1][r] - acc[3][r], -(acc[0][r] - acc[2][r]) };
                }
                if (rs >= 0) {
                    ph[0][(size_t)isel * nrings + rs] = (cpx){ bs * (acc[4][r] + acc[6][r]), bs * (acc[5][r] + acc[7][r]) };
                    ph[1][(size_t)isel * nrings + rs] = (cpx){ bs * (acc[5][r] - acc[7][r]), -bs * (acc[4][r] - acc[6][r]) };
                }
            }
        }
    }
    if (!dir_synth) {
        for (int l = l0; l <= lmax; ++l) {
            const size_t k = ab + l; const double g = gamma[l];
            if (s == 0) { alm_out[0][2 * k] = g * a0r[l]; alm_out[0][2 * k + 1] = (m == 0) ? 0.0 : g * a0i[l]; }
            else {
                const double h = 0.5 * g;
                alm_out[0][2 * k] = -h * (a0r[l] + a1r[l]); alm_out[0][2 * k + 1] = (m == 0) ? 0.0 : -h * (a0i[l] + a1i[l]);
                alm_out[1][2 * k] = -h * (a0i[l] - a1i[l]); alm_out[1][2 * k + 1] = (m == 0) ? 0.0 : h * (a0r[l] - a1r[l]);
            }
        }
        for (int l = m; l < l0 && l <= lmax; ++l) for (int c = 0; c < 2; ++c) { alm_out[c][2 * (ab + l)] = 0; alm_out[c][2 * (ab + l) + 1] = 0; }
        free(a0r);
    }
}

/* ---------------------------------------------------------------------------------------------------------------
 * ring FFT stage (half-length complex FFT of the even/odd packed ring), selected m only in the phase array
 * ------------------------------------------------------------------------------------------------------------- */
static void ring_synth(const fft_plan *F, int nphi, int mmax, const int *msel, int nsel, const cpx *ph0tw, const cpx *phrow, size_t stride,
                       double *ring, cpx *buf, cpx *scr, cpx *X, const cpx *hw)
{
    const int n = nphi / 2;
    for (int k = 0; k <= n; ++k) X[k] = (cpx){0, 0};
    for (int i = 0; i < nsel; ++i) {
        const int m = msel[i]; if (m > mmax) continue;
        const cpx a = phrow[(size_t)i * stride], r = ph0tw[i];
        const double sx = a.re * r.re - a.im * r.im, sy = a.re * r.im + a.im * r.re;
        const int k = m % nphi;
        if (k <= n) { X[k].re += sx; X[k].im += sy; }
        if (nphi - k <= n) { X[nphi - k].re += sx; X[nphi - k].im -= sy; }     /* conjugate image (k >= n, incl. Nyquist) */
    }
    /* Z[k] = (X[k] + conj X[n-k]) + i e^{+2 pi i k/nphi} (X[k] - conj X[n-k]) */
    for (int k = 0; k < n; ++k) {
        const cpx xa = X[k], xb = X[n - k];
        const double c = hw[k].re, s = -hw[k].im;         /* e^{+2 pi i k / nphi} */
        const double er = xa.re + xb.re, ei = xa.im - xb.im, dr = xa.re - xb.re, di = xa.im + xb.im;
        const double orr = dr * c - di * s, oi = dr * s + di * c;
        buf[k].re = er - oi; buf[k].im = ei + orr;
    }
    fft_run(F, buf, scr, +1);
    for (int j = 0; j < n; ++j) { ring[2 * j] = buf[j].re; ring[2 * j + 1] = buf[j].im; }
}

static void ring_anal(const fft_plan *F, int nphi, const int *msel, int nsel, const cpx *ph0tw, double w, const double *ring, cpx *phrow,
                      size_t stride, cpx *buf, cpx *scr, cpx *X, const cpx *hw)
{
    const int n = nphi / 2;
    for (int j = 0; j < n; ++j) { buf[j].re = ring[2 * j]; buf[j].im = ring[2 * j + 1]; }
    fft_run(F, buf, scr, -1);
    X[0] = (cpx){ buf[0].re + buf[0].im, 0 }; X[n] = (cpx){ buf[0].re - buf[0].im, 0 };
    for (int k = 1; k < n; ++k) {
        const cpx za = buf[k], zb = buf[n - k];
        const double er = za.re + zb.re, ei = za.im - zb.im, dr = za.re - zb.re, di = za.im + zb.im;
        const double c = hw[k].re, s = hw[k].im;          /* e^{-2 pi i k / nphi} */
        const double orr = dr * c - di * s, oi = dr * s + di * c;
        X[k].re = 0.5 * (er + oi); X[k].im = 0.5 * (ei - orr);
    }
    for (int i = 0; i < nsel; ++i) {
        const int k = msel[i] % nphi;
        cpx f = k <= n ? X[k] : (cpx){ X[nphi - k].re, -X[nphi - k].im };
        const cpx r = ph0tw[i];
        phrow[(size_t)i * stride] = (cpx){ w * (f.re * r.re + f.im * r.im), w * (f.im * r.re - f.re * r.im) };
    }
}

static int select_m(int mmax, int m_stride, int m_offset, int **out)
{
    int n = 0;
    for (int m = 0; m <= mmax; ++m) if (m % m_stride == m_offset) ++n;
    *out = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int i = 0;
    for (int m = 0; m <= mmax; ++m) if (m % m_stride == m_offset) (*out)[i++] = m;
    return n;
}

int cpu_alm2map(int spin, int nrings, const double *theta, double phi0, int nphi, int lmax, int mmax, const double *const *alms,
                double *const *maps, int m_stride, int m_offset, double *times)
{
    if (!(spin == 0 || spin == 2) || mmax > lmax || m_stride < 1 || (nphi & 1)) return 1;
    const int nc = spin == 0 ? 1 : 2;
    int *msel; const int nsel = select_m(mmax, m_stride, m_offset, &msel);
    pairs_t P; pairs_build(&P, nrings, theta, lmax);
    long double *lg0 = (long double *)malloc(sizeof(long double) * (size_t)(mmax + 1) * 2), *lg2 = lg0 + (mmax + 1);
    prefactors(mmax, lg0, lg2);
    cpx *ph[2] = { NULL, NULL };
    for (int c = 0; c < nc; ++c) ph[c] = (cpx *)malloc(sizeof(cpx) * (size_t)(nsel > 0 ? nsel : 1) * nrings);
    double t0 = now_s();
#pragma omp parallel
    {
        double *alpha = (double *)malloc(sizeof(double) * (size_t)(lmax + 1) * 3), *delta = alpha + (lmax + 1), *gamma = delta + (lmax + 1);
#pragma omp for schedule(dynamic, 1)
        for (int i = 0; i < nsel; ++i) {
            const int m = msel[i];
            for (int l = 0; l <= lmax; ++l) { alpha[l] = delta[l] = gamma[l] = 0; }
            for (int c = 0; c < nc; ++c) for (int r = 0; r < nrings; ++r) ph[c][(size_t)i * nrings + r] = (cpx){0, 0};
            legendre_m(1, spin, lmax, m, &P, spin == 0 ? lg0[m] : lg2[m], nrings, i, nsel, alms, NULL, ph, alpha, delta, gamma);
        }
        free(alpha);
    }
    double t1 = now_s();
    const double tleg = t1 - t0;
    cpx *ph0tw = (cpx *)malloc(sizeof(cpx) * (size_t)(nsel > 0 ? nsel : 1));
    for (int i = 0; i < nsel; ++i) { long double a = fmodl((long double)msel[i] * (long double)phi0, 2.0L * PI_L); ph0tw[i] = (cpx){ (double)cosl(a), (double)sinl(a) }; }
    fft_plan F; fft_plan_init(&F, nphi / 2);
    cpx *hw = (cpx *)malloc(sizeof(cpx) * (size_t)(nphi / 2 + 1));
    for (int k = 0; k <= nphi / 2; ++k) { long double a = 2.0L * PI_L * k / nphi; hw[k] = (cpx){ (double)cosl(a), (double)(-sinl(a)) }; }
    t1 = now_s();
#pragma omp parallel
    {
        cpx *buf = (cpx *)malloc(sizeof(cpx) * (size_t)(nphi + 2) * 2), *scr = buf + nphi / 2, *X = scr + nphi / 2;
#pragma omp for schedule(dynamic, 8) collapse(2)
        for (int c = 0; c < nc; ++c)
            for (int r = 0; r < nrings; ++r)
                ring_synth(&F, nphi, mmax, msel, nsel, ph0tw, ph[c] + r, (size_t)nrings, maps[c] + (size_t)r * nphi, buf, scr, X, hw);
        free(buf);
    }
    double t2 = now_s();
    if (times) { times[0] = tleg; times[1] = t2 - t1; }
    fft_plan_free(&F); free(ph0tw); free(hw);
    for (int c = 0; c < nc; ++c) free(ph[c]);
    free(lg0); pairs_free(&P); free(msel);
    return 0;
}

int cpu_map2alm(int spin, int nrings, const double *theta, const double *wgt, double phi0, int nphi, int lmax, int mmax,
                const double *const *maps, double *const *alms, int m_stride, int m_offset, double *times)
{
    if (!(spin == 0 || spin == 2) || mmax > lmax || m_stride < 1 || (nphi & 1)) return 1;
    const int nc = spin == 0 ? 1 : 2;
    int *msel; const int nsel = select_m(mmax, m_stride, m_offset, &msel);
    pairs_t P; pairs_build(&P, nrings, theta, lmax);
    long double *lg0 = (long double *)malloc(sizeof(long double) * (size_t)(mmax + 1) * 2), *lg2 = lg0 + (mmax + 1);
    prefactors(mmax, lg0, lg2);
    cpx *ph[2] = { NULL, NULL };
    for (int c = 0; c < nc; ++c) ph[c] = (cpx *)malloc(sizeof(cpx) * (size_t)(nsel > 0 ? nsel : 1) * nrings);
    cpx *ph0tw = (cpx *)malloc(sizeof(cpx) * (size_t)(nsel > 0 ? nsel : 1));
    for (int i = 0; i < nsel; ++i) { long double a = fmodl((long double)msel[i] * (long double)phi0, 2.0L * PI_L); ph0tw[i] = (cpx){ (double)cosl(a), (double)sinl(a) }; }
    cpx *hw = (cpx *)malloc(sizeof(cpx) * (size_t)(nphi / 2 + 1));
    for (int k = 0; k <= nphi / 2; ++k) { long double a = 2.0L * PI_L * k / nphi; hw[k] = (cpx){ (double)cosl(a), (double)(-sinl(a)) }; }
    fft_plan F; fft_plan_init(&F, nphi / 2);
    double t0 = now_s();
#pragma omp parallel
    {
        cpx *buf = (cpx *)malloc(sizeof(cpx) * (size_t)(nphi + 2) * 2), *scr = buf + nphi / 2, *X = scr + nphi / 2;
#pragma omp for schedule(dynamic, 8) collapse(2)
        for (int c = 0; c < nc; ++c)
            for (int r = 0; r < nrings; ++r)
                ring_anal(&F, nphi, msel, nsel, ph0tw, wgt[r], maps[c] + (size_t)r * nphi, ph[c] + r, (size_t)nrings, buf, scr, X, hw);
        free(buf);
    }
    double t1 = now_s();
#pragma omp parallel
    {
        double *alpha = (double *)malloc(sizeof(double) * (size_t)(lmax + 1) * 3), *delta = alpha + (lmax + 1), *gamma = delta + (lmax + 1);
#pragma omp for schedule(dynamic, 1)
        for (int i = 0; i < nsel; ++i) {
            const int m = msel[i];
            for (int l = 0; l <= lmax; ++l) { alpha[l] = delta[l] = gamma[l] = 0; }
            legendre_m(0, spin, lmax, m, &P, spin == 0 ? lg0[m] : lg2[m], nrings, i, nsel, NULL, alms, ph, alpha, delta, gamma);
        }
        free(alpha);
    }
    double t2 = now_s();
    if (times) { times[0] = t2 - t1; times[1] = t1 - t0; }
    fft_plan_free(&F); free(hw); free(ph0tw);
    for (int c = 0; c < nc; ++c) free(ph[c]);
    free(lg0); pairs_free(&P); free(msel);
    return 0;
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the timed baseline must use the cores it claims */
/* Host FP64 FMA peak (GFLOP/s, 1 FMA = 2 flop) on `nthreads` threads: register-resident independent FMA chains, written for the
 * compiler's vectoriser like the Legendre loops above, so that it measures the SIMD width this build was compiled for
 * (-march=native).  bench.py quotes the port's own rate against it (cpu_baseline.frac_of_host_peak). */
double cpu_fma_peak_gflops(int nthreads, double seconds)
{
    enum { NV = 64 };          /* 64 doubles = 8 AVX-512 / 16 AVX2 registers of independent chains */
    double total = 0.0, tmax = 0.0;
    if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads) reduction(+ : total) reduction(max : tmax)
#endif
    {
        double a[NV] __attribute__((aligned(64))), sink = 0.0;
        for (int i = 0; i < NV; ++i) a[i] = 1.0 + 1e-9 * i;
        const double mlt = 1.0 - 1e-12, add = 1e-13;
        double t0 = now_s(), t1 = t0;
        long iters = 0;
        do {
            for (int r = 0; r < 4096; ++r)
#pragma GCC ivdep
                for (int i = 0; i < NV; ++i) a[i] = a[i] * mlt + add;
            iters += 4096;
            t1 = now_s();
        } while (t1 - t0 < seconds);
        for (int i = 0; i < NV; ++i) sink += a[i];
        total += 2.0 * (double)NV * (double)iters + (sink == 12345.678 ? 1.0 : 0.0);
        tmax = t1 - t0;
    }
    return tmax > 0.0 ? total / tmax * 1e-9 : 0.0;
}

void cpu_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int cpu_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
