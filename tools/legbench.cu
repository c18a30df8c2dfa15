// tools/legbench.cu -- kernel-level timing harness (profiling aid, not product code): loads a libpixsht build by path, creates a
// full-sky Clenshaw-Curtis plan, runs alm2map + map2alm on device-resident IQU data and prints the per-kernel CUDA-event times of
// pixsht_get_timings plus checksums of the outputs, so that kernel variants (different .so builds) can be compared in one GPU call
// without Python/torch start-up.
//   nvcc -O2 -o tools/_bin/legbench tools/legbench.cu -ldl
//   tools/_bin/legbench <lib.so> <res_arcmin> <lmax> <reps> [ncomp]
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "../include/pixsht.h"

__device__ __forceinline__ double hash01(unsigned long long k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return (double)(k >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}
__global__ void k_fill_alm(double2* a, long long n, int lmax, unsigned long long seed, int spin2)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double re = hash01(seed + 2 * i), im = hash01(seed + 2 * i + 1);
        if (i <= lmax) { im = 0.0; if (spin2 && i < 2) re = 0.0; }           // m = 0 column: real; l < 2 vanish for spin 2
        else if (spin2) {
            // l < 2 entries of the m = 1 column
            if (i == lmax + 1) { re = 0.0; im = 0.0; }
        }
        a[i] = make_double2(re, im);
    }
}
__global__ void k_sum(const double* v, long long n, double* out)
{
    double s = 0.0, s2 = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) { s += v[i] * ((i % 7) + 1); s2 += v[i] * v[i]; }
    atomicAdd(out, s); atomicAdd(out + 1, s2);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)

int main(int argc, char** argv)
{
    if (argc < 5) { fprintf(stderr, "usage: legbench lib.so res_arcmin lmax reps [ncomp]\n"); return 1; }
    const char* path = argv[1];
    const double res = atof(argv[2]);
    const int lmax = atoi(argv[3]), reps = atoi(argv[4]);
    const int nc = argc > 5 ? atoi(argv[5]) : 3;
    void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 1; }
#define SYM(name) auto p_##name = (decltype(&name))dlsym(h, #name); if (!p_##name) { fprintf(stderr, "missing %s\n", #name); return 1; }
    SYM(pixsht_plan_create) SYM(pixsht_execute) SYM(pixsht_get_timings) SYM(pixsht_plan_destroy) SYM(pixsht_nalm) SYM(pixsht_last_error) SYM(pixsht_plan_info)
    pixsht_geom g;
    memset(&g, 0, sizeof(g));
    g.nphi = (int)llround(360.0 * 60.0 / res); g.nrings_total = g.nphi / 2 + 1; g.ring_first = 0; g.nrings = g.nrings_total; g.nx = g.nphi;
    g.flipx = 0; g.flipy = 0; g.ring_scheme = PIXSHT_RINGS_CC; g.phi0 = 0.0;
    pixsht_plan* P = nullptr;
    if (p_pixsht_plan_create(&P, &g, lmax, lmax, PIXSHT_F64, 0)) { fprintf(stderr, "plan: %s\n", p_pixsht_last_error()); return 1; }
    const long long nalm = p_pixsht_nalm(lmax, lmax), npix = (long long)g.nx * g.nrings;
    void *alm[3], *out[3], *map[3];
    for (int c = 0; c < nc; ++c) {
        CK(cudaMalloc(&alm[c], nalm * 16)); CK(cudaMalloc(&out[c], nalm * 16)); CK(cudaMalloc(&map[c], npix * 8));
        const int spin2 = (nc == 2) || (nc == 3 && c > 0);
        k_fill_alm<<<1024, 256>>>((double2*)alm[c], nalm, lmax, 0x9e3779b97f4a7c15ULL * (c + 1), spin2);
    }
    CK(cudaDeviceSynchronize());
    double best[2][8];
    for (int d = 0; d < 2; ++d) for (int k = 0; k < 8; ++k) best[d][k] = 1e30;
    for (int r = 0; r < reps + 1; ++r) {
        double t[8];
        if (p_pixsht_execute(P, PIXSHT_ALM2MAP, nc, alm, map, PIXSHT_DEVICE)) { fprintf(stderr, "alm2map: %s\n", p_pixsht_last_error()); return 1; }
        p_pixsht_get_timings(P, t);
        if (r) for (int k = 0; k < 8; ++k) best[0][k] = fmin(best[0][k], t[k]);
        if (p_pixsht_execute(P, PIXSHT_MAP2ALM, nc, out, map, PIXSHT_DEVICE)) { fprintf(stderr, "map2alm: %s\n", p_pixsht_last_error()); return 1; }
        p_pixsht_get_timings(P, t);
        if (r) for (int k = 0; k < 8; ++k) best[1][k] = fmin(best[1][k], t[k]);
    }
    int32_t info[16]; p_pixsht_plan_info(P, info);
    double* d_s; CK(cudaMalloc(&d_s, 4 * sizeof(double))); CK(cudaMemset(d_s, 0, 4 * sizeof(double)));
    double hs[4];
    printf("%s res=%g lmax=%d nc=%d R=[%d %d %d %d]\n", path, res, lmax, nc, info[10], info[11], info[12], info[13]);
    printf("  alm2map: leg %.3f (spin0 %.3f spin2 %.3f) fft %.3f call %.3f ms\n", best[0][1], best[0][6], best[0][7], best[0][2], best[0][4]);
    printf("  map2alm: leg %.3f (spin0 %.3f spin2 %.3f) fft %.3f call %.3f ms\n", best[1][1], best[1][6], best[1][7], best[1][2], best[1][4]);
    printf("  total %.3f ms\n", best[0][4] + best[1][4]);
    for (int c = 0; c < nc; ++c) {
        CK(cudaMemset(d_s, 0, 4 * sizeof(double)));
        k_sum<<<512, 256>>>((const double*)map[c], npix, d_s);
        k_sum<<<512, 256>>>((const double*)out[c], 2 * nalm, d_s + 2);
        CK(cudaMemcpy(hs, d_s, sizeof(hs), cudaMemcpyDeviceToHost));
        printf("  comp %d: map sum %.15e sq %.15e | alm sum %.15e sq %.15e\n", c, hs[0], hs[1], hs[2], hs[3]);
    }
    if (auto prof = (int (*)(unsigned long long*))dlsym(h, "pixsht_debug_fft_prof")) {
        // -DPIXSHT_FFT_PROF builds: share of each phase in the FFT kernels' CTA time (cycle sums of thread 0 over all CTAs and calls)
        unsigned long long c[32]; prof(c);
        const char* names[2][4] = {{"load", "pre", "store", ""}, {"load", "post", "store", ""}};
        for (int d = 0; d < 2; ++d) {
            double tot = 0; for (int k = 0; k < 16; ++k) tot += (double)c[16 * d + k];
            printf("  %s phases:", d ? "fft_map2phase" : "fft_phase2map");
            for (int k = 0; k < 3; ++k) printf(" %s %.1f%%", names[d][k], 100.0 * c[16 * d + k] / tot);
            for (int k = 4; k < 12; ++k) if (c[16 * d + k]) printf(" pass%d %.1f%%", k - 4, 100.0 * c[16 * d + k] / tot);
            printf("  (cycles per ring-call sum %.3e)\n", tot);
        }
    }
    p_pixsht_plan_destroy(P);
    return 0;
}
