// TEST INFRASTRUCTURE — the four CUDA runtime calls of gpu-pathtracer_b200/host/adapter_harness.cpp as plain host
// functions, so that the reference-signature adapter (BeginRender / Render / EndRender) can be linked against the CPU
// emulation build of the product (tests/emu/libb200pt_emu.so, where "device" memory is host memory) and the drop-in
// boundary is exercised on GPU-less machines too (tests/test_adapter_emu.py).  Never linked into the product.
#include <cuda_runtime.h>
#include <cstdlib>
#include <cstring>

extern "C" cudaError_t CUDARTAPI cudaMalloc(void** p, size_t n) { *p = std::malloc(n); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
extern "C" cudaError_t CUDARTAPI cudaFree(void* p) { std::free(p); return cudaSuccess; }
extern "C" cudaError_t CUDARTAPI cudaMemcpy(void* dst, const void* src, size_t n, enum cudaMemcpyKind) { std::memcpy(dst, src, n); return cudaSuccess; }
extern "C" cudaError_t CUDARTAPI cudaGetLastError(void) { return cudaSuccess; }
