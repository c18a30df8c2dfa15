// tools/rf_probe.cu -- register-file operand-bandwidth probe for the FP64 pipe of sm_100a (profiling aid, not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/rf_probe tools/rf_probe.cu && tools/_bin/rf_probe
// Question: what does a DFMA cost as a function of how many of its three sources are fresh vector-register reads, and
// does the operand-reuse cache (.reuse) survive when several warps share a scheduler?  Each kernel issues the same number
// of independent DFMAs (16 accumulators x R values); they differ only in where the second and third source come from.
//   UR   : acc[a][j] += p[j] * g[a], g warp-uniform and loop-invariant -> ptxas keeps g in a uniform register
//   REUSE: same, but g is a per-thread value (identical in all lanes): vector register, consecutive DFMAs share it (.reuse)
//   FRESH: acc[a][j] += p[j] * q[(a + j) % 4 ...]: no two consecutive DFMAs share a source in the same slot
//   ANAL : analysis-style chains  part[k] += p[j] * X[k][j]  (k < NP chains, three fresh sources, chain length R)
// Run with one-warp CTAs at 1, 2, 3, 4, 5 warps per scheduler (4 .. 20 CTAs per SM), as the Legendre kernels are launched.
#include <cuda_runtime.h>
#include <cstdio>

constexpr int R = 4;

template <int MODE>
__global__ void __launch_bounds__(MODE == 0 ? 256 : 32) k_acc(double* out, int iters, const double* __restrict__ in)
{
    double p[R], g[4], q[4], acc[4][R];
#pragma unroll
    for (int j = 0; j < R; ++j) p[j] = in[j] + threadIdx.x * 1e-9;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        g[a] = (MODE == 0) ? in[8 + a] : in[8 + a + (threadIdx.x >> 6)];   // >> 6 is 0 for every lane, but per-thread to the compiler
        q[a] = in[12 + a] + threadIdx.x * 1e-9;
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int j = 0; j < R; ++j) acc[a][j] = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    if (MODE <= 1) acc[a][j] = fma(p[j], g[a], acc[a][j]);
                    else acc[a][j] = fma(p[j], q[(a + j + u) & 3], acc[a][j]);
                }
    }
    double s = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int j = 0; j < R; ++j) s += acc[a][j];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// analysis-style: per step NP chains, each summing RR products p[j] * X[k][j]; results folded into a running sum (one DADD per
// chain and step stands in for the store of the partial sums)
template <int RR, int NP>
__global__ void __launch_bounds__(32) k_anal(double* out, int iters, const double* __restrict__ in)
{
    double p[RR], X[NP][RR], tot[NP];
#pragma unroll
    for (int j = 0; j < RR; ++j) {
        p[j] = in[j % 8] + threadIdx.x * 1e-9;
#pragma unroll
        for (int k = 0; k < NP; ++k) X[k][j] = in[(j + k) % 16] * 1e-3 + threadIdx.x * 1e-9;
    }
#pragma unroll
    for (int k = 0; k < NP; ++k) tot[k] = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            double part[NP];
#pragma unroll
            for (int k = 0; k < NP; ++k) part[k] = 0;
#pragma unroll
            for (int j = 0; j < RR; ++j)
#pragma unroll
                for (int k = 0; k < NP; ++k) part[k] = fma(p[j], X[k][j], part[k]);
#pragma unroll
            for (int k = 0; k < NP; ++k) tot[k] += part[k];
#pragma unroll
            for (int j = 0; j < RR; ++j) p[j] = p[j] * 1.0000001;   // keeps p loop-variant (one DMUL per ring and step, as x * p)
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < NP; ++k) s += tot[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static double timeit(F launch)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return best;
}

int main()
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    double *out, *in; cudaMalloc(&out, sizeof(double) * sms * 64 * 1024); cudaMalloc(&in, 32 * 8);
    double h[32]; for (int i = 0; i < 32; ++i) h[i] = 0.3 + 0.01 * i;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    printf("%s, %d SMs; TFLOP/s counts DFMA only (1 DFMA = 2 flop)\n", prop.name, sms);
    const int per_sm[] = {4, 8, 12, 16, 20};
    for (int c : per_sm) {
        const int blocks = sms * c;
        const double thr = (double)blocks * 32;
        const int it = 4096;
        printf("one-warp CTAs, %2d per SM (%d warps per scheduler)\n", c, c / 4);
        double ms;
        ms = timeit([&] { k_acc<0><<<blocks, 32>>>(out, it, in); });
        printf("   acc += p*g   g in a uniform register         : %6.2f TF\n", 2.0 * 16 * R * it * thr / ms / 1e9);
        ms = timeit([&] { k_acc<1><<<blocks, 32>>>(out, it, in); });
        printf("   acc += p*g   g in a vector register (.reuse) : %6.2f TF\n", 2.0 * 16 * R * it * thr / ms / 1e9);
        ms = timeit([&] { k_acc<2><<<blocks, 32>>>(out, it, in); });
        printf("   acc += p*q   three fresh vector sources      : %6.2f TF\n", 2.0 * 16 * R * it * thr / ms / 1e9);
        ms = timeit([&] { k_anal<8, 2><<<blocks, 32>>>(out, it, in); });
        printf("   analysis chains R=8 NP=2 (+1 DMUL,  per ring) : %6.2f TF DFMA, %6.2f TF all FP64\n", 2.0 * 4 * 16 * it * thr / ms / 1e9,
               2.0 * 4 * (16 + 8 + 2) * it * thr / ms / 1e9);
        ms = timeit([&] { k_anal<8, 4><<<blocks, 32>>>(out, it, in); });
        printf("   analysis chains R=8 NP=4                      : %6.2f TF DFMA, %6.2f TF all FP64\n", 2.0 * 4 * 32 * it * thr / ms / 1e9,
               2.0 * 4 * (32 + 8 + 4) * it * thr / ms / 1e9);
        ms = timeit([&] { k_anal<4, 4><<<blocks, 32>>>(out, it, in); });
        printf("   analysis chains R=4 NP=4                      : %6.2f TF DFMA, %6.2f TF all FP64\n", 2.0 * 4 * 16 * it * thr / ms / 1e9,
               2.0 * 4 * (16 + 4 + 4) * it * thr / ms / 1e9);
        ms = timeit([&] { k_anal<4, 8><<<blocks, 32>>>(out, it, in); });
        printf("   analysis chains R=4 NP=8                      : %6.2f TF DFMA, %6.2f TF all FP64\n", 2.0 * 4 * 32 * it * thr / ms / 1e9,
               2.0 * 4 * (32 + 4 + 8) * it * thr / ms / 1e9);
    }
    return 0;
}
