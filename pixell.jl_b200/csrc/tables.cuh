// tables.cuh -- device-side precompute: quadrature weights, recurrence coefficient tables, FFT twiddles.
// Replaces the per-call host setup of the reference (src/transforms.jl:44-46: FastTransforms CC weights; libsharp2's
// per-m sharp_Ylmgen_prepare) by one-off kernels at plan creation.
#pragma once
#include "common.cuh"

namespace pixsht {

// Clenshaw-Curtis ring weights w_k = c_k * 2pi/nphi for band rings k = ring_first .. ring_first+nrings-1 of the
// N-ring full-sky grid (closed form of SURVEY.md A.2; end points from the exact 1/(n^2-1+n%2)).
__global__ void k_cc_weights(int N, int nphi, int ring_first, int nrings, double* __restrict__ w)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrings) return;
    const int k = ring_first + i;
    const int n = N - 1;
    const double scale = 2.0 * 3.14159265358979323846 / (double)nphi;
    if (N == 1) { w[i] = 2.0 * scale; return; }
    if (k == 0 || k == n) { w[i] = scale / ((double)n * (double)n - 1.0 + (double)(n & 1)); return; }
    // sum small terms first (q descending) for accuracy
    double s = 0.0;
    for (int q = n / 2; q >= 1; --q) {
        const double b = (2 * q == n) ? 1.0 : 2.0;
        const long long t = ((long long)2 * q * k) % (2LL * n);  // exact argument reduction of 2 q k pi / n
        s += b / (4.0 * (double)q * (double)q - 1.0) * cospi((double)t / (double)n);
    }
    w[i] = (2.0 / (double)n) * (1.0 - s) * scale;
}

// Fejer's first rule on the N interior nodes theta_k = pi (k + 1/2)/N (no ring on the poles): ring weights
// w_k = f_k * 2pi/nphi,  f_k = (2/N) [1 - 2 sum_{q=1}^{floor(N/2)} cos(2 q theta_k)/(4 q^2 - 1)]  (exact for polynomials in
// cos(theta) of degree < N).  The reference declares the CarFejer1 type but has no SHT path for it.
__global__ void k_fejer1_weights(int N, int nphi, int ring_first, int nrings, double* __restrict__ w)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrings) return;
    const int k = ring_first + i;
    const double scale = 2.0 * 3.14159265358979323846 / (double)nphi;
    double s = 0.0;
    for (int q = N / 2; q >= 1; --q) {
        const long long t = ((long long)q * (2 * k + 1)) % (2LL * N);   // 2 q theta_k = pi q (2k+1)/N, reduced mod 2 pi
        s += cospi((double)t / (double)N) / (4.0 * (double)q * (double)q - 1.0);
    }
    w[i] = (2.0 / (double)N) * (1.0 - 2.0 * s) * scale;
}

// A_l of lambda_{l+1} = A_l (x - mu_l) lambda_l - (A_l/A_{l-1}) lambda_{l-1}   (normalised d-functions)
__device__ __forceinline__ double coefA(int l, int m, int s)
{
    const double l1 = (double)(l + 1);
    const double num = (double)(2 * l + 1) * (double)(2 * l + 3);
    const double den = ((l1 - (double)m) * (l1 + (double)m)) * ((l1 - (double)s) * (l1 + (double)s));
    return l1 * sqrt(num / den);
}

// Per m (one thread each): (alpha_l, delta_l) and gamma_l, l = l0..lmax, l0 = max(m, s), stored at alm_index(lmax, l, m).
// lambda_l = gamma_l p_l with p_{l+1} = (alpha_l x + delta_l) p_l - p_{l-1}  (unit lower coefficient => 2 FMA per step):
//   gamma_{l0} = gamma_{l0+1} = 1, gamma_{l+1} = (A_l/A_{l-1}) gamma_{l-1}, alpha_l = A_l gamma_l / gamma_{l+1},
//   delta_l = -alpha_l mu_l with mu_l = -m s/(l(l+1)) for d^l_{-m,s}  (the s = -2 sequence uses -delta_l).
__global__ void k_coef_tables(int lmax, int mmax, int s, double2* __restrict__ ad, double* __restrict__ gamma)
{
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m > mmax) return;
    const int l0 = m > s ? m : s;
    const long long base = alm_index(lmax, 0, m);
    for (int l = m; l < l0 && l <= lmax; ++l) { ad[base + l] = make_double2(0.0, 0.0); gamma[base + l] = 0.0; }
    if (l0 > lmax) return;
    const double ms = (double)m * (double)s;
    double Aprev = coefA(l0, m, s);
    double g_lm1 = 1.0, g_l = 1.0;  // gamma_{l-1}, gamma_l while stepping; start at l = l0+1
    gamma[base + l0] = 1.0;
    ad[base + l0] = make_double2(Aprev, l0 > 0 ? Aprev * ms / ((double)l0 * (double)(l0 + 1)) : 0.0);   // gamma_{l0}/gamma_{l0+1} = 1
    for (int l = l0 + 1; l <= lmax; ++l) {
        const double A = coefA(l, m, s);
        const double g_lp1 = (A / Aprev) * g_lm1;
        const double a = A * g_l / g_lp1;
        gamma[base + l] = g_l;
        ad[base + l] = make_double2(a, a * ms / ((double)l * (double)(l + 1)));
        g_lm1 = g_l; g_l = g_lp1; Aprev = A;
    }
}

// tw[t] = exp(-2 pi i t / n), t = 0..n-1
__global__ void k_twiddles(int n, double2* __restrict__ tw)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double s, c;
    sincospi(2.0 * (double)t / (double)n, &s, &c);
    tw[t] = make_double2(c, -s);
}

}  // namespace pixsht
