// TEST INFRASTRUCTURE — host shim that lets g++ compile the reference's own src/pathtracer.cu kernel bodies
// as ordinary C++.  Force-included (-include) before the reference translation unit.
// It defines only what the CUDA compiler would otherwise provide; it contains no path-tracing arithmetic.
#pragma once
#define __DEVICE_LAUNCH_PARAMETERS_H__ 1     // skip CUDA's extern-const builtin declarations
#define THRUST_DEVICE_SYSTEM THRUST_DEVICE_SYSTEM_CPP
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
struct shim_idx3 { unsigned x, y, z; };
extern thread_local shim_idx3 threadIdx, blockIdx;
extern shim_idx3 blockDim, gridDim;
#define __powf powf
// CUDA's host-side definition of rsqrtf (crt/math_functions.hpp) is 1.0f / sqrtf(x)
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline float atomicAdd(float* p, float v) { float old; 
#pragma omp atomic capture
  { old = *p; *p += v; }
  return old; }
static inline int atomicAdd(int* p, int v) { int old;
#pragma omp atomic capture
  { old = *p; *p += v; }
  return old; }
