// TEST INFRASTRUCTURE — headless driver that calls the reference-signature BeginRender / Render / EndRender of
// pathtracer_adapter.cpp with a reference `Scene` rebuilt from the flat scene view (oracle/refbuild/common_view.h),
// i.e. exactly what the reference's main.cpp does (main.cpp:300, :139, :246) minus the GL window.  Built by
// oracle/build_ref.sh into oracle/_ref/libadapter.so because it needs the reference's headers.
#include "scene.h"
#include "pathtracer.h"
#include "common_view.h"
#include <cuda_runtime.h>

struct b200pt_ctx;
struct b200pt_multi;
b200pt_ctx* B200ptContext();
b200pt_multi* B200ptMultiContext();

static Scene* g_scene = nullptr;
static Camera g_cam;
static unsigned g_w = 0, g_h = 0;
static float3* g_out = nullptr;

extern "C" int adapter_begin(const b200pt_scene_view* v, unsigned w, unsigned h, float eps) {
    if (g_scene) return -1;
    g_scene = new Scene();
    scene_from_view(*g_scene, &g_cam, v);
    g_w = w; g_h = h;
    BeginRender(*g_scene, w, h, eps);
    return cudaMalloc(&g_out, sizeof(float3) * (size_t)w * h) == cudaSuccess ? 0 : -2;
}
extern "C" int adapter_render(unsigned first_iter, unsigned n, int reset_first, float* out_host) {
    if (!g_scene) return -1;
    for (unsigned it = first_iter; it < first_iter + n; ++it)
        Render(*g_scene, g_w, g_h, &g_cam, it, reset_first && it == first_iter, g_out);
    if (out_host) cudaMemcpy(out_host, g_out, sizeof(float3) * (size_t)g_w * g_h, cudaMemcpyDeviceToHost);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
extern "C" int adapter_get_accum(float* host) {
    if (B200ptMultiContext()) return b200pt_multi_get_accum(B200ptMultiContext(), host);      // B200PT_GPUS > 1
    return b200pt_get_accum(B200ptContext(), host, 0);
}
extern "C" int adapter_end() {
    if (!g_scene) return -1;
    EndRender();
    cudaFree(g_out); g_out = nullptr;
    delete g_scene; g_scene = nullptr;
    return 0;
}
