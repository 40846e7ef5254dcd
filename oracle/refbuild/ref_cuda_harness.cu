// TEST INFRASTRUCTURE — headless C-ABI driver around the reference's OWN CUDA integrator
// (src/pathtracer.cu BeginRender/Render/EndRender, compiled for sm_100a from /root/reference by
// oracle/build_ref.sh).  This is the oracle of record for GPU parity and the GPU reference timing.
// Our code; links against the reference translation units, contains none of their source.
#include "scene.h"
#include "pathtracer.h"
#include "common_view.h"
#include <cuda_runtime.h>

extern float3 *dev_image, *dev_color;   // src/pathtracer.cu:10

static Scene*  g_scene = nullptr;
static Camera  g_cam;
static unsigned g_w = 0, g_h = 0;
static float3* g_out = nullptr;

extern "C" int refcuda_begin(const b200pt_scene_view* v, unsigned w, unsigned h, float eps) {
    if (g_scene) return -1;
    g_scene = new Scene();
    scene_from_view(*g_scene, &g_cam, v);
    g_w = w; g_h = h;
    BeginRender(*g_scene, w, h, eps);
    // dev_image/dev_color are not initialised by BeginRender (src/pathtracer.cu:2664-2666); zero them so the
    // "stale colour on NaN" quirk (:1019) starts from a defined state.
    cudaMemset(dev_image, 0, sizeof(float3) * (size_t)w * h);
    cudaMemset(dev_color, 0, sizeof(float3) * (size_t)w * h);
    cudaMalloc(&g_out, sizeof(float3) * (size_t)w * h);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -2;
}

// Render(iter = first .. first+n-1, reset = reset_first && iter == first); returns device ms of the loop.
extern "C" int refcuda_render(unsigned first_iter, unsigned n, int reset_first, float* out_host, float* ms_out) {
    if (!g_scene) return -1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (unsigned it = first_iter; it < first_iter + n; ++it)
        Render(*g_scene, g_w, g_h, &g_cam, it, reset_first && it == first_iter, g_out);
    cudaEventRecord(e1);
    cudaError_t err = cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    if (ms_out) *ms_out = ms;
    if (out_host) cudaMemcpy(out_host, g_out, sizeof(float3) * (size_t)g_w * g_h, cudaMemcpyDeviceToHost);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (err == cudaSuccess) err = cudaGetLastError();
    return err == cudaSuccess ? 0 : -2;
}

extern "C" int refcuda_set_camera(const void* cam104) { memcpy((void*)&g_cam, cam104, sizeof(Camera)); return 0; }

extern "C" int refcuda_get_accum(float* host) {
    return cudaMemcpy(host, dev_image, sizeof(float3) * (size_t)g_w * g_h, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}
extern "C" int refcuda_get_color(float* host) {
    return cudaMemcpy(host, dev_color, sizeof(float3) * (size_t)g_w * g_h, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}
extern "C" int refcuda_end() {
    if (!g_scene) return -1;
    EndRender();
    cudaFree(g_out); g_out = nullptr;
    delete g_scene; g_scene = nullptr;
    return 0;
}

#ifdef REFDBG
// probe build only (oracle/build_ref_debug.sh): choose the pixel whose per-bounce state the kernels print
extern "C" void refdbg_set(int pixel);
extern "C" int refcuda_debug_pixel(int pixel) { refdbg_set(pixel); fflush(stdout); return 0; }
#endif
