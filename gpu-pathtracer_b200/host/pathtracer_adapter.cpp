// pathtracer_adapter.cpp — the reference's three C++ entry points (src/pathtracer.h:10-12) implemented on the
// C ABI of include/b200pt.h.  A maintainer of brickray/gpu-pathtracer drops this file into src/, removes
// pathtracer.cu from the executable's sources and links libb200pt.so; main.cpp, parsescene.cpp,
// bvh.cpp and scene.h stay untouched.  It is compiled against the REFERENCE's own headers (scene.h,
// pathtracer.h) and contains none of their code.
//
//   BeginRender(scene, w, h, ep)                    src/pathtracer.cu:2568  ->  b200pt_create
//   Render(scene, w, h, camera, iter, reset, out)   src/pathtracer.cu:2705  ->  b200pt_render(spp = 1, device output)
//   EndRender()                                     src/pathtracer.cu:2697  ->  b200pt_destroy
//
// Same contract as the reference: one global render context per process, `output` is a caller-owned DEVICE
// pointer to w*h float3 (the mapped GL pixel buffer of main.cpp:136), `iter` is 1-based and part of the RNG
// seed, `reset` zeroes the accumulation first, the camera is re-read on every call.  Errors print and abort
// like HANDLE_ERROR (src/common.h:29-39).
//
// B200PT_GPUS=N (N > 1) in the environment: the same three calls drive N GPUs of the box — b200pt_create_multi makes one
// tile-sharded context per device, every Render is followed by ONE NCCL reduce of the accumulation framebuffers onto
// device 0 inside the library, and `output` (device-0 memory) receives the full tonemapped image, bit-identical to N = 1.
#include "scene.h"
#include "pathtracer.h"
#include "b200pt.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

static_assert(sizeof(Camera) == B200PT_SIZEOF_CAMERA, "Camera layout changed");
static_assert(sizeof(Primitive) == B200PT_SIZEOF_PRIMITIVE, "Primitive layout changed");
static_assert(sizeof(LinearBVHNode) == B200PT_SIZEOF_BVHNODE, "LinearBVHNode layout changed");
static_assert(sizeof(Material) == B200PT_SIZEOF_MATERIAL, "Material layout changed");
static_assert(sizeof(Medium) == B200PT_SIZEOF_MEDIUM, "Medium layout changed");
static_assert(sizeof(Area) == B200PT_SIZEOF_AREA, "Area layout changed");
static_assert(sizeof(Infinite) == B200PT_SIZEOF_INFINITE, "Infinite layout changed");

static b200pt_ctx* g_ctx = nullptr;
static b200pt_multi* g_multi = nullptr;         // B200PT_GPUS > 1

static void die(const char* what) {
    fprintf(stderr, "b200pt %s failed: %s\n", what, b200pt_last_error());
    abort();
}

void BeginRender(Scene& scene, unsigned width, unsigned height, float ep) {
    if (g_ctx || g_multi) EndRender();
    std::vector<b200pt_texture> tex(scene.textures.size());
    for (size_t i = 0; i < tex.size(); ++i) {
        tex[i].texels = scene.textures[i].data.data();
        tex[i].width = scene.textures[i].width;
        tex[i].height = scene.textures[i].height;
    }
    b200pt_scene_view v = {};
    v.camera = scene.camera;
    v.prims = scene.bvh.prims.data();              v.n_prims = (int32_t)scene.bvh.prims.size();
    v.nodes = scene.bvh.linear_root;               v.n_nodes = scene.bvh.total_nodes;
    v.materials = scene.materials.data();          v.n_materials = (int32_t)scene.materials.size();
    v.mediums = scene.mediums.data();              v.n_mediums = (int32_t)scene.mediums.size();
    v.lights = scene.lights.data();                v.n_lights = (int32_t)scene.lights.size();
    v.infinite = &scene.infinite;
    v.light_distribution = scene.lightDistribution.data();
    v.n_light_distribution = (int32_t)scene.lightDistribution.size();
    v.textures = tex.empty() ? nullptr : tex.data(); v.n_textures = (int32_t)tex.size();
    v.integrator_type = (int32_t)scene.integrator.type;
    v.max_depth = scene.integrator.maxDepth;
    const char* gpus = getenv("B200PT_GPUS");
    if (gpus && atoi(gpus) > 1) {
        if (b200pt_create_multi(&v, width, height, ep, atoi(gpus), nullptr, &g_multi) != B200PT_OK) die("BeginRender (multi-GPU)");
        return;
    }
    const char* dev = getenv("B200PT_DEVICE");
    if (b200pt_create(&v, width, height, ep, dev ? atoi(dev) : 0, nullptr, &g_ctx) != B200PT_OK) die("BeginRender");
}

void Render(Scene& scene, unsigned width, unsigned height, Camera* camera, unsigned iter, bool reset, float3* output) {
    (void)scene; (void)width; (void)height;       // fixed at BeginRender, exactly like the reference's device copies
    if (g_multi) {
        if (b200pt_multi_render(g_multi, camera, iter, 1, reset ? 1 : 0, (float*)output, 1) != B200PT_OK) die("Render (multi-GPU)");
        return;
    }
    if (!g_ctx) { fprintf(stderr, "Render called before BeginRender\n"); abort(); }
    if (b200pt_render(g_ctx, camera, iter, 1, reset ? 1 : 0, (float*)output, 1) != B200PT_OK) die("Render");
}

void EndRender() {
    if (g_ctx) { b200pt_destroy(g_ctx); g_ctx = nullptr; }
    if (g_multi) { b200pt_multi_destroy(g_multi); g_multi = nullptr; }
}

// Optional batched extension for headless callers: `spp` consecutive Render calls in one launch sequence.
void RenderBatch(Scene& scene, Camera* camera, unsigned first_iter, unsigned spp, bool reset, float3* output) {
    (void)scene;
    if (g_multi) {
        if (b200pt_multi_render(g_multi, camera, first_iter, spp, reset ? 1 : 0, (float*)output, 1) != B200PT_OK) die("RenderBatch (multi-GPU)");
        return;
    }
    if (!g_ctx) { fprintf(stderr, "RenderBatch called before BeginRender\n"); abort(); }
    if (b200pt_render(g_ctx, camera, first_iter, spp, reset ? 1 : 0, (float*)output, 1) != B200PT_OK) die("RenderBatch");
}

b200pt_ctx* B200ptContext() { return g_ctx; }
b200pt_multi* B200ptMultiContext() { return g_multi; }
