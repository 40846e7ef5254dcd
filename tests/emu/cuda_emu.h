// TEST INFRASTRUCTURE — a minimal single-process stand-in for the CUDA runtime and the SIMT built-ins, so that
// the SAME sources as the product (gpu-pathtracer_b200/csrc/*.cu, *.cuh) can be compiled by g++ with
// -DB200PT_EMULATE into tests/emu/libb200pt_emu.so.  The build container has no GPU; this lets the CPU test
// suite execute the wavefront state machine, the batching / polling host loop and the tile sharding, and compare
// them bit-for-bit with the CPU oracle.  It is NOT a fallback: the package (gpu_pathtracer_b200/_lib.py) only
// ever loads csrc/libb200pt.so and fails loudly without it; nothing outside tests/ references this library.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <omp.h>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __grid_constant__
#define __launch_bounds__(...)
#define __restrict__
#define __shared__ static thread_local
#define __align__(x)

struct float4 { float x, y, z, w; };
struct int4 { int x, y, z, w; };
struct int2 { int x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }

struct emu_idx { unsigned x, y, z; };
extern thread_local emu_idx threadIdx, blockIdx;
extern emu_idx blockDim, gridDim;

static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicMax(unsigned* p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline int atomicMin(int* p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
// warp built-ins with a ONE-lane warp: every emulated thread is its own warp (lane 0)
template <class T> static inline T __shfl_down_sync(unsigned, T v, int) { return (T)0; }
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
static inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline void __syncthreads() {}

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorEmu = 1 };
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
struct cudaDeviceProp { int multiProcessorCount; };
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->multiProcessorCount = 4; return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n); return *p ? cudaSuccess : cudaErrorEmu; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n); return *p ? cudaSuccess : cudaErrorEmu; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, int) { *s = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, int) { *e = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
template <class K> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, K, int, size_t) { *n = 2; return cudaSuccess; }

// kernel<<<grid, block>>>(args): every (block, thread) pair runs the kernel body once; blocks in parallel
template <class F> static inline void emu_launch(F&& body, unsigned grid, unsigned block) {
    gridDim = {grid, 1, 1}; blockDim = {block, 1, 1};
#pragma omp parallel for schedule(dynamic, 1)
    for (long b = 0; b < (long)grid; ++b) {
        blockIdx = {(unsigned)b, 0, 0};
        for (unsigned t = 0; t < block; ++t) { threadIdx = {t, 0, 0}; body(); }
    }
}
#define PT_LAUNCH(kernel, grid, block, smem, stream, ...) emu_launch([&]() { kernel(__VA_ARGS__); }, (unsigned)(grid), (unsigned)(block))
// kernels written CTA-at-a-time (k_wave.cuh: phases separated by block barriers): the body is called ONCE per block and
// loops over the block's threads itself, phase by phase
#define PT_LAUNCH_CTA(kernel, grid, block, smem, stream, ...) emu_launch([&]() { kernel(__VA_ARGS__); }, (unsigned)(grid), 1u)
