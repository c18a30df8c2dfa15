// tools/dfma_mix.cu -- FP64 pipe micro-benchmarks that bracket the Legendre inner loop (profiling aid, not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dfma_mix tools/dfma_mix.cu && /tmp/dfma_mix
// K0: register-resident FMA chains with constant multiplier/addend (what pixsht_measure_fma_peak measures)
// K1: accumulate pattern  acc[a][j] += p[j] * g[a]            (three distinct register operands per DFMA)
// K2: the spin-0 synthesis step without memory: u = alpha*x[j]; pn = fma(u, p, -pp); acc += p*g   (R = 4)
// K3: K2 with the per-step operands read from shared memory (LDS.64 + LDS.128, warp-uniform address)
// Each is run with 32-thread CTAs (the Legendre launch shape) at several residencies.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define R 4

__global__ void k0(double* out, int iters, double a, double b)
{
    double v[8];
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fma(v[i], a, b);
    double s = 0; for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k1(double* out, int iters, const double* in)
{
    double p[R], g[4], acc[4][R];
    for (int j = 0; j < R; ++j) p[j] = in[j] + threadIdx.x * 1e-9;
    for (int a = 0; a < 4; ++a) g[a] = in[8 + a];
    for (int a = 0; a < 4; ++a) for (int j = 0; j < R; ++j) acc[a][j] = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int j = 0; j < R; ++j) acc[a][j] = fma(p[j], g[a], acc[a][j]);
    }
    double s = 0; for (int a = 0; a < 4; ++a) for (int j = 0; j < R; ++j) s += acc[a][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// MODE 0: operands in registers; 1: LDS.64 + LDS.128 per step (6 registers written); 2: two LDS.128 per step (8 registers);
//      3: one LDS.128 per step (4 registers, alpha in a register); 4: no loads but 6 integer adds per step (6 ALU register writes)
// FORM 0: u = alpha*x; pn = fma(u, p, -pp)   (three vector-register operands in the recurrence FMA)
// FORM 1: t = x*p;     pn = fma(alpha, t, -pp)   (the warp-uniform alpha can come from the operand reuse cache)
// FORM 2: FORM 1, statements ordered ring by ring
template <int RR, int MODE, int FORM = 0>
__global__ void k23(double* out, int iters, const double* in)
{
    __shared__ __align__(16) double rec[128 * 4];
    for (int i = threadIdx.x; i < 128 * 4; i += blockDim.x) rec[i] = in[i % 16] * 1e-3 + ((i % 4 == 0) ? 1.0 : 0.0);
    __syncthreads();
    double x[RR], p[RR], pp[RR], acc[4][RR];
    for (int j = 0; j < RR; ++j) { x[j] = in[j % 8] * 1e-3 + 0.5; p[j] = 1e-3 * (threadIdx.x + 1); pp[j] = 0; }
    for (int a = 0; a < 4; ++a) for (int j = 0; j < RR; ++j) acc[a][j] = 0;
    double alpha0 = in[4], g0x = in[5], g0y = in[6];
    int ia[6] = {1, 2, 3, 4, 5, 6};
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int i = 0; i < 128; i += 2) {
            double al = alpha0, gx = g0x, gy = g0y, al2 = alpha0, gx2 = g0y, gy2 = g0x;
            if (MODE == 1) {
                al = rec[i * 4]; al2 = rec[i * 4 + 4];
                const double2 g = *reinterpret_cast<const double2*>(rec + i * 4 + 2);
                const double2 g2 = *reinterpret_cast<const double2*>(rec + i * 4 + 6);
                gx = g.x; gy = g.y; gx2 = g2.x; gy2 = g2.y;
            } else if (MODE == 2) {
                const double2 c = *reinterpret_cast<const double2*>(rec + i * 4);
                const double2 g = *reinterpret_cast<const double2*>(rec + i * 4 + 2);
                const double2 c2 = *reinterpret_cast<const double2*>(rec + i * 4 + 4);
                const double2 g2 = *reinterpret_cast<const double2*>(rec + i * 4 + 6);
                al = c.x + c.y; gx = g.x; gy = g.y; al2 = c2.x + c2.y; gx2 = g2.x; gy2 = g2.y;
            } else if (MODE == 3) {
                const double2 g = *reinterpret_cast<const double2*>(rec + i * 4 + 2);
                const double2 g2 = *reinterpret_cast<const double2*>(rec + i * 4 + 6);
                gx = g.x; gy = g.y; gx2 = g2.x; gy2 = g2.y;
            } else if (MODE == 4) {
#pragma unroll
                for (int q = 0; q < 6; ++q) { ia[q] += ia[(q + 1) % 6] ^ i; }
#pragma unroll
                for (int q = 0; q < 6; ++q) { ia[q] += ia[(q + 2) % 6] | i; }
            }
            if (FORM == 0) {
#pragma unroll
                for (int j = 0; j < RR; ++j) {
                    acc[0][j] = fma(p[j], gx, acc[0][j]);
                    acc[1][j] = fma(p[j], gy, acc[1][j]);
                    pp[j] = fma(al * x[j], p[j], -pp[j]);
                }
#pragma unroll
                for (int j = 0; j < RR; ++j) {
                    acc[2][j] = fma(pp[j], gx2, acc[2][j]);
                    acc[3][j] = fma(pp[j], gy2, acc[3][j]);
                    p[j] = fma(al2 * x[j], pp[j], -p[j]);
                }
            } else if (FORM == 1) {
                double t[RR];
#pragma unroll
                for (int j = 0; j < RR; ++j) acc[0][j] = fma(p[j], gx, acc[0][j]);
#pragma unroll
                for (int j = 0; j < RR; ++j) acc[1][j] = fma(p[j], gy, acc[1][j]);
#pragma unroll
                for (int j = 0; j < RR; ++j) t[j] = x[j] * p[j];
#pragma unroll
                for (int j = 0; j < RR; ++j) pp[j] = fma(al, t[j], -pp[j]);
#pragma unroll
                for (int j = 0; j < RR; ++j) acc[2][j] = fma(pp[j], gx2, acc[2][j]);
#pragma unroll
                for (int j = 0; j < RR; ++j) acc[3][j] = fma(pp[j], gy2, acc[3][j]);
#pragma unroll
                for (int j = 0; j < RR; ++j) t[j] = x[j] * pp[j];
#pragma unroll
                for (int j = 0; j < RR; ++j) p[j] = fma(al2, t[j], -p[j]);
            } else if (FORM == 3) {
                // ring-major block that reads p[j] once (operand reuse cache), then the recurrence FMAs with alpha reused
                double t[RR];
#pragma unroll
                for (int j = 0; j < RR; ++j) {
                    t[j] = p[j] * x[j];
                    acc[0][j] = fma(p[j], gx, acc[0][j]);
                    acc[1][j] = fma(p[j], gy, acc[1][j]);
                }
#pragma unroll
                for (int j = 0; j < RR; ++j) pp[j] = fma(al, t[j], -pp[j]);
#pragma unroll
                for (int j = 0; j < RR; ++j) {
                    t[j] = pp[j] * x[j];
                    acc[2][j] = fma(pp[j], gx2, acc[2][j]);
                    acc[3][j] = fma(pp[j], gy2, acc[3][j]);
                }
#pragma unroll
                for (int j = 0; j < RR; ++j) p[j] = fma(al2, t[j], -p[j]);
            } else {
#pragma unroll
                for (int j = 0; j < RR; ++j) {
                    acc[0][j] = fma(p[j], gx, acc[0][j]);
                    acc[1][j] = fma(p[j], gy, acc[1][j]);
                    pp[j] = fma(al, p[j] * x[j], -pp[j]);
                }
#pragma unroll
                for (int j = 0; j < RR; ++j) {
                    acc[2][j] = fma(pp[j], gx2, acc[2][j]);
                    acc[3][j] = fma(pp[j], gy2, acc[3][j]);
                    p[j] = fma(al2, pp[j] * x[j], -p[j]);
                }
            }
        }
    }
    double s = 0; for (int a = 0; a < 4; ++a) for (int j = 0; j < RR; ++j) s += acc[a][j] + p[j];
    for (int q = 0; q < 6; ++q) s += ia[q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
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
    return best;
}

int main()
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    double *out, *in; cudaMalloc(&out, sizeof(double) * sms * 64 * 1024); cudaMalloc(&in, 16 * 8);
    double h[16]; for (int i = 0; i < 16; ++i) h[i] = 0.3 + 0.01 * i; h[4] = 1.9;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    printf("%s, %d SMs\n", prop.name, sms);
    const int shapes[][2] = {{32, 12}, {32, 20}, {128, 8}};   // threads per CTA, CTAs per SM
    for (auto& sh : shapes) {
        const int nt = sh[0], blocks = sms * sh[1];
        const double thr = (double)blocks * nt;
        printf("threads %3d x %2d CTA/SM\n", nt, sh[1]);
        {
            const int it = 2048; double ms = timeit([&] { k0<<<blocks, nt>>>(out, it, 1.0000001, 1e-9); });
            printf("   K0 const-operand chains                    : %6.2f TF\n", 2.0 * 64 * it * thr / ms / 1e9);
        }
        {
            const int it = 2048; double ms = timeit([&] { k1<<<blocks, nt>>>(out, it, in); });
            printf("   K1 acc += p*g                              : %6.2f TF\n", 2.0 * 64 * it * thr / ms / 1e9);
        }
        const int it = 32;
#define RUN(RR, MODE, FORM, label) { double ms = timeit([&] { k23<RR, MODE, FORM><<<blocks, nt>>>(out, it, in); }); \
            printf("   synth step R=%d form %d %-26s: %6.2f TF (16 FP64/step counted; only 12 executed when alpha is loop-invariant, form 0 'registers')\n", RR, FORM, label, 2.0 * 4 * RR * 128 * it * thr / ms / 1e9); }
        RUN(4, 0, 0, "registers") RUN(4, 1, 0, "LDS.64+LDS.128") RUN(4, 1, 1, "LDS.64+LDS.128") RUN(4, 1, 2, "LDS.64+LDS.128")
        RUN(8, 1, 0, "LDS.64+LDS.128") RUN(8, 1, 1, "LDS.64+LDS.128") RUN(8, 1, 2, "LDS.64+LDS.128")
        RUN(2, 1, 0, "LDS.64+LDS.128") RUN(2, 1, 1, "LDS.64+LDS.128")
        RUN(4, 1, 3, "LDS.64+LDS.128") RUN(8, 1, 3, "LDS.64+LDS.128")
        RUN(4, 0, 1, "registers (all 16 executed)") RUN(8, 0, 1, "registers (all 16 executed)")
    }
    return 0;
}
