// tests/emu/cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
//
// A minimal host emulation of the CUDA execution model (one OS thread per CUDA thread, blocks run one after another),
// just big enough to execute the kernels of pixell.jl_b200/csrc on the CPU of the GPU-less build container so that
// their *logic* (indexing, parity bookkeeping, rescaling state machine, FFT passes) can be debugged before GPU time is
// spent.  It is compiled only into tests/emu/_build/libpixsht_emu.so by tests/emu/build_emu.sh and loaded only by
// tests/test_emu_*.py.  The product library (pixell.jl_b200/lib/libpixsht.so) is built by nvcc, contains no CPU path,
// and never sees this header.
#pragma once
#include <pthread.h>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define PIXSHT_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __grid_constant__
#define __align__(x) __attribute__((aligned(x)))

struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint3_ { unsigned x, y, z; };
struct double2 { double x, y; };
struct alignas(16) float2 { float x, y; };
struct alignas(32) double4 { double x, y, z, w; };
static inline double2 make_double2(double a, double b) { double2 r; r.x = a; r.y = b; return r; }
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
static inline double4 make_double4(double a, double b, double c, double d) { double4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }

namespace emu {
struct Block {
    pthread_barrier_t bar;
    std::vector<pthread_barrier_t> wbar;
    std::vector<uint64_t> slots;  // 32 per warp
    int or_acc = 0;
    std::mutex mu;
    unsigned char* dyn_smem = nullptr;
};
inline thread_local uint3_ t_threadIdx, t_blockIdx;
inline dim3 g_blockDim, g_gridDim;
inline Block* g_block = nullptr;
}  // namespace emu
#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)
#define warpSize 32

static inline void __syncthreads() { pthread_barrier_wait(&emu::g_block->bar); }
static inline int emu_warp() { return (int)(threadIdx.x / 32); }
static inline int emu_lane() { return (int)(threadIdx.x % 32); }
static inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&emu::g_block->wbar[emu_warp()]); }
static inline int __syncthreads_or(int p)
{
    __syncthreads();
    if (threadIdx.x == 0) emu::g_block->or_acc = 0;
    __syncthreads();
    if (p) { std::lock_guard<std::mutex> g(emu::g_block->mu); emu::g_block->or_acc = 1; }
    __syncthreads();
    int r = emu::g_block->or_acc;
    __syncthreads();
    return r;
}
static inline uint64_t emu_xchg(uint64_t v, int src_lane)
{
    uint64_t* s = &emu::g_block->slots[(size_t)emu_warp() * 32];
    s[emu_lane()] = v;
    __syncwarp();
    uint64_t r = s[src_lane & 31];
    __syncwarp();
    return r;
}
static inline int __any_sync(unsigned, int p)
{
    uint64_t* s = &emu::g_block->slots[(size_t)emu_warp() * 32];
    s[emu_lane()] = p ? 1 : 0;
    __syncwarp();
    int r = 0;
    for (int i = 0; i < 32; ++i) r |= (int)s[i];
    __syncwarp();
    return r;
}
static inline int __all_sync(unsigned m, int p) { return !__any_sync(m, !p); }
static inline unsigned __ballot_sync(unsigned, int p)
{
    uint64_t* s = &emu::g_block->slots[(size_t)emu_warp() * 32];
    s[emu_lane()] = p ? 1 : 0;
    __syncwarp();
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= (unsigned)s[i] << i;
    __syncwarp();
    return r;
}
static inline double __shfl_xor_sync(unsigned, double v, int lanemask)
{
    uint64_t u; memcpy(&u, &v, 8);
    u = emu_xchg(u, emu_lane() ^ lanemask);
    memcpy(&v, &u, 8);
    return v;
}
static inline int __shfl_xor_sync(unsigned, int v, int lanemask) { return (int)(uint32_t)emu_xchg((uint64_t)(uint32_t)v, emu_lane() ^ lanemask); }
static inline int __shfl_sync(unsigned, int v, int src) { return (int)emu_xchg((uint64_t)(uint32_t)v, src); }
static inline double atomicAdd(double* addr, double v)
{
    static std::mutex mu;
    std::lock_guard<std::mutex> g(mu);
    double old = *addr; *addr = old + v; return old;
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline unsigned long long atomicAdd(unsigned long long* addr, unsigned long long v)
{
    static std::mutex mu;
    std::lock_guard<std::mutex> g(mu);
    unsigned long long old = *addr; *addr = old + v; return old;
}
static inline int __double2hiint(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(u >> 32); }
static inline int __double2loint(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(u & 0xffffffffu); }
static inline double __hiloint2double(int hi, int lo)
{
    uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double d; memcpy(&d, &u, 8); return d;
}
static inline void sincospi(double x, double* s, double* c)
{
    long double a = 3.14159265358979323846264338327950288L * (long double)x;
    *s = (double)sinl(a); *c = (double)cosl(a);
}
static inline double cospi(double x) { return (double)cosl(3.14159265358979323846264338327950288L * (long double)x); }
static inline double __ldg(const double* p) { return *p; }
static inline double2 __ldg(const double2* p) { return *p; }
static inline int __ldg(const int* p) { return *p; }
using std::max;
using std::min;

// ---------------------------------------------------------------- runtime shim
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef struct emuEvent { double t; }* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorUnknown = 999 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaStreamNonBlocking = 1 };
struct cudaDeviceProp { int multiProcessorCount; size_t sharedMemPerBlockOptin; size_t sharedMemPerMultiprocessor = 233472; char name[64]; int major, minor; };
static inline const char* cudaGetErrorString(cudaError_t) { return "emu error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int)
{
    p->multiProcessorCount = 4; p->sharedMemPerBlockOptin = 227 * 1024; strcpy(p->name, "host-emulation"); p->major = 10; p->minor = 0;
    return cudaSuccess;
}
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? cudaSuccess : cudaErrorUnknown; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memmove(d, s, n); return cudaSuccess; }
template <class S> static inline cudaError_t cudaMemcpyToSymbol(S& sym, const void* src, size_t n) { memcpy((void*)&sym, src, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emuEvent{0}; return cudaSuccess; }
enum { cudaEventDefault = 0, cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new emuEvent{0}; return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

namespace emu {
template <class K, class... A>
void launch(K kernel, dim3 grid, dim3 block, size_t smem, A... args)
{
    if (block.x % 32 != 0 || block.y != 1 || block.z != 1) { fprintf(stderr, "emu: block must be 1-D, multiple of 32\n"); abort(); }
    g_blockDim = block; g_gridDim = grid;
    Block B;
    const unsigned nt = block.x, nw = nt / 32;
    pthread_barrier_init(&B.bar, nullptr, nt);
    B.wbar.resize(nw);
    for (auto& w : B.wbar) pthread_barrier_init(&w, nullptr, 32);
    B.slots.assign((size_t)nw * 32, 0);
    std::vector<unsigned char> dyn(smem + 64);
    B.dyn_smem = (unsigned char*)(((uintptr_t)dyn.data() + 31) & ~(uintptr_t)31);
    g_block = &B;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                std::vector<std::thread> th;
                th.reserve(nt);
                for (unsigned t = 0; t < nt; ++t)
                    th.emplace_back([=]() {
                        t_threadIdx = {t, 0, 0};
                        t_blockIdx = {bx, by, bz};
                        kernel(args...);
                    });
                for (auto& x : th) x.join();
            }
    pthread_barrier_destroy(&B.bar);
    for (auto& w : B.wbar) pthread_barrier_destroy(&w);
    g_block = nullptr;
}
}  // namespace emu
#define PIXSHT_LAUNCH(kernel, grid, block, smem, stream, ...) emu::launch(kernel, dim3(grid), dim3(block), (size_t)(smem), __VA_ARGS__)
#define PIXSHT_DYN_SMEM(name) unsigned char* name = emu::g_block->dyn_smem
