// TEST INFRASTRUCTURE — storage for the emulated SIMT built-ins (see cuda_emu.h).
#include "cuda_emu.h"
thread_local emu_idx threadIdx, blockIdx;
emu_idx blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
