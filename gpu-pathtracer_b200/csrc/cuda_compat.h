// cuda_compat.h — the one place that includes the CUDA runtime and defines the kernel-launch macro.
// (B200PT_EMULATE is defined only by tests/emu/Makefile, which compiles these same sources for the CPU test
// suite of the GPU-less build container; the product library is always built by nvcc for sm_100a.)
#pragma once
#ifdef B200PT_EMULATE
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#define PT_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
// kernels whose body plays a whole CTA per call under emulation (phases separated by block barriers: k_wave.cuh)
#define PT_LAUNCH_CTA(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif
