// TEST INFRASTRUCTURE — C-ABI driver around the reference's OWN kernel bodies compiled for the host
// (src/pathtracer.cu through host_shim.h; see oracle/build_ref.sh).  Used (i) in this container, where no
// GPU exists, to pin the CPU restatement oracle/pt_oracle.cpp bit-for-bit, (ii) as the "reference" CPU
// baseline of bench.py, (iii) for the reference's scene preparation (Scene::Init = BVH build + light CDF,
// Camera ctor) when generating fixtures.  Our code; it #includes the staged reference translation unit.
#include "pathtracer_host.cu"      // staged + cut copy of the reference's src/pathtracer.cu (never in the repo)
#include "common_view.h"
#include <omp.h>
#include <unistd.h>

thread_local shim_idx3 threadIdx, blockIdx;
shim_idx3 blockDim = {32, 4, 1}, gridDim = {1, 1, 1};

// draw sequencing helpers used by the patched multi-draw argument lists (device order = left to right)
namespace {
Scene*  g_scene = nullptr;
Camera  g_cam;
unsigned g_w = 0, g_h = 0;
std::vector<float3> g_acc, g_color, g_out;
Infinite g_inf;
}

extern "C" int refhost_begin(const b200pt_scene_view* v, unsigned w, unsigned h, float eps) {
    if (g_scene) return -1;
    g_scene = new Scene();
    scene_from_view(*g_scene, &g_cam, v);
    g_w = w; g_h = h;
    g_acc.assign((size_t)w * h, make_float3(0, 0, 0));
    g_color.assign((size_t)w * h, make_float3(0, 0, 0));
    g_out.assign((size_t)w * h, make_float3(0, 0, 0));
    Scene& s = *g_scene;
    // what InitRender stores into the __device__ globals (src/pathtracer.cu:2533-2566)
    kernel_camera = &g_cam;
    kernel_linear = s.bvh.linear_root;
    kernel_primitives = s.bvh.prims.data();
    kernel_materials = s.materials.data();
    kernel_bssrdfs = nullptr;
    kernel_mediums = s.mediums.data();
    kernel_lights = s.lights.data();
    g_inf = s.infinite;
    kernel_infinite = &g_inf;
    // what BeginRender uploads for textures (src/pathtracer.cu:2646-2661): per-texture texel pointers + (w, h) pairs
    static std::vector<uchar4*> tex_ptrs; static std::vector<int> tex_sizes;
    tex_ptrs.clear(); tex_sizes.clear();
    for (size_t i = 0; i < s.textures.size(); ++i) {
        tex_ptrs.push_back(s.textures[i].data.data());
        tex_sizes.push_back(s.textures[i].width); tex_sizes.push_back(s.textures[i].height);
    }
    kernel_textures = tex_ptrs.empty() ? nullptr : tex_ptrs.data();
    kernel_texture_size = tex_sizes.empty() ? nullptr : tex_sizes.data();
    kernel_light_distribution = s.lightDistribution.data();
    kernel_light_size = (int)s.lights.size();
    kernel_light_distribution_size = (int)s.lightDistribution.size();
    kernel_acc_image = g_acc.data();
    kernel_color = g_color.data();
    kernel_epsilon = eps;
    blockDim = {32, 4, 1};
    gridDim = {w / 32, h / 4, 1};
    return 0;
}

extern "C" int refhost_set_camera(const void* cam104) { memcpy((void*)&g_cam, cam104, sizeof(Camera)); return 0; }

// One Render() call per iteration (src/pathtracer.cu:2705-2750) with the launch replaced by loops.
extern "C" int refhost_render(unsigned first_iter, unsigned n, int reset_first, float* out_host, int nthreads) {
    if (!g_scene) return -1;
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    IntegratorType type = g_scene->integrator.type;
    int maxDepth = g_scene->integrator.maxDepth;
    const int nblocks = (int)(gridDim.x * gridDim.y);
    for (unsigned it = first_iter; it < first_iter + n; ++it) {
        bool reset = reset_first && it == first_iter;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
        for (int b = 0; b < nblocks; ++b) {
            blockIdx = {b % gridDim.x, b / gridDim.x, 0};
            for (unsigned ty = 0; ty < blockDim.y; ++ty)
                for (unsigned tx = 0; tx < blockDim.x; ++tx) {
                    threadIdx = {tx, ty, 0};
                    if (type == IT_PT) Path((int)it, maxDepth);
                    else if (type == IT_VPT) Volpath((int)it, maxDepth);
                }
        }
#pragma omp parallel for schedule(static) num_threads(nthreads)
        for (int b = 0; b < nblocks; ++b) {
            blockIdx = {b % gridDim.x, b / gridDim.x, 0};
            for (unsigned ty = 0; ty < blockDim.y; ++ty)
                for (unsigned tx = 0; tx < blockDim.x; ++tx) {
                    threadIdx = {tx, ty, 0};
                    Output((int)it, g_out.data(), reset, g_cam.filmic, type);
                }
        }
    }
    if (out_host) memcpy(out_host, g_out.data(), sizeof(float3) * g_out.size());
    return 0;
}
extern "C" int refhost_get_accum(float* host) { memcpy(host, g_acc.data(), sizeof(float3) * g_acc.size()); return 0; }
extern "C" int refhost_get_color(float* host) { memcpy(host, g_color.data(), sizeof(float3) * g_color.size()); return 0; }
extern "C" int refhost_end() {
    if (!g_scene) return -1;
    free(g_scene->bvh.linear_root);
    delete g_scene; g_scene = nullptr;
    return 0;
}

// ---- function-level known-answer entry points ------------------------------------------------
// ray8 = {o.xyz, d.xyz, tmin, tmax}
static Ray make_ray(const float* r8) {
    Ray r; r.o = make_float3(r8[0], r8[1], r8[2]); r.d = make_float3(r8[3], r8[4], r8[5]);
    r.tmin = r8[6]; r.tmax = r8[7]; r.medium = nullptr; return r;
}
// isect16 = pos3 nor3 uv2 dpdu3 matIdx bssrdf lightIdx mediumInside mediumOutside (ints as raw bits)
static void put_isect(const Intersection& is, float* o) {
    memcpy(o, &is, sizeof(Intersection));
}
extern "C" int refhost_bbox_intersect(const float* box6, const float* ray8) {
    BBox b(make_float3(box6[0], box6[1], box6[2]), make_float3(box6[3], box6[4], box6[5]));
    Ray r = make_ray(ray8);
    return b.Intersect(r) ? 1 : 0;
}
extern "C" int refhost_prim_intersect(const void* prim176, const float* ray8, float* t_out, float* isect16) {
    Primitive p; memcpy((void*)&p, prim176, sizeof(Primitive));
    Ray r = make_ray(ray8);
    Intersection is; memset((void*)&is, 0, sizeof(is)); is.lightIdx = -1;
    bool hit = false;
    if (p.type == GT_TRIANGLE) hit = p.triangle.Intersect(r, &is);
    else if (p.type == GT_SPHERE) hit = p.sphere.Intersect(r, &is);
    else if (p.type == GT_LINES) hit = p.line.Intersect(r, &is);
    *t_out = r.tmax;
    put_isect(is, isect16);
    return hit ? 1 : 0;
}
// closest / any hit over the scene loaded by refhost_begin (src/pathtracer.cu:214, :257)
extern "C" int refhost_intersect(const float* ray8, float* t_out, float* isect16) {
    Ray r = make_ray(ray8);
    Intersection is; memset((void*)&is, 0, sizeof(is)); is.lightIdx = -1;
    bool hit = Intersect(r, &is);
    *t_out = r.tmax; put_isect(is, isect16);
    return hit ? 1 : 0;
}
extern "C" int refhost_intersect_p(const float* ray8) { Ray r = make_ray(ray8); return IntersectP(r) ? 1 : 0; }
extern "C" void refhost_sample_bsdf(const void* mat72, const float* in3, const float* nor3, const float* uv2,
                                    const float* dpdu3, const float* u3, float* out3, float* fr3, float* pdf) {
    Material m; memcpy((void*)&m, mat72, sizeof(Material));
    float3 out = make_float3(0, 0, 0), fr = make_float3(0, 0, 0); float p = 0;
    SampleBSDF(m, make_float3(in3[0], in3[1], in3[2]), make_float3(nor3[0], nor3[1], nor3[2]), make_float2(uv2[0], uv2[1]),
               make_float3(dpdu3[0], dpdu3[1], dpdu3[2]), make_float3(u3[0], u3[1], u3[2]), out, fr, p);
    out3[0] = out.x; out3[1] = out.y; out3[2] = out.z; fr3[0] = fr.x; fr3[1] = fr.y; fr3[2] = fr.z; *pdf = p;
}
extern "C" void refhost_fr(const void* mat72, const float* in3, const float* out3, const float* nor3, const float* uv2,
                           const float* dpdu3, float* fr3, float* pdf) {
    Material m; memcpy((void*)&m, mat72, sizeof(Material));
    float3 fr = make_float3(0, 0, 0); float p = 0;
    Fr(m, make_float3(in3[0], in3[1], in3[2]), make_float3(out3[0], out3[1], out3[2]), make_float3(nor3[0], nor3[1], nor3[2]),
       make_float2(uv2[0], uv2[1]), make_float3(dpdu3[0], dpdu3[1], dpdu3[2]), fr, p);
    fr3[0] = fr.x; fr3[1] = fr.y; fr3[2] = fr.z; *pdf = p;
}
extern "C" void refhost_camera_ray(const void* cam104, float x, float y, float ax, float ay, float* o3, float* d3) {
    Camera c; memcpy((void*)&c, cam104, sizeof(Camera));
    Ray r = c.GeneratePrimaryRay(x, y, make_float2(ax, ay));
    o3[0] = r.o.x; o3[1] = r.o.y; o3[2] = r.o.z; d3[0] = r.d.x; d3[1] = r.d.y; d3[2] = r.d.z;
}
// the first n uniform draws of the stream of (pixel, iter) (src/pathtracer.cu:888-889)
extern "C" void refhost_rng(unsigned pixel, unsigned iter, int n, float* out) {
    thrust::default_random_engine rng(WangHash(pixel) + WangHash(iter));
    thrust::uniform_real_distribution<float> uniform(0.0f, 1.0f);
    for (int i = 0; i < n; ++i) out[i] = uniform(rng);
}
extern "C" void refhost_area_sample(const void* area192, const float* pos3, const float* u2, float eps,
                                    float* rad3, float* ray8, float* nor3, float* pdf) {
    Area a; memcpy((void*)&a, area192, sizeof(Area));
    float3 pos = make_float3(pos3[0], pos3[1], pos3[2]); float2 u = make_float2(u2[0], u2[1]);
    float3 rad, nor; Ray r; float p;
    a.SampleLight(pos, u, rad, r, nor, p, eps);
    rad3[0] = rad.x; rad3[1] = rad.y; rad3[2] = rad.z; nor3[0] = nor.x; nor3[1] = nor.y; nor3[2] = nor.z; *pdf = p;
    ray8[0] = r.o.x; ray8[1] = r.o.y; ray8[2] = r.o.z; ray8[3] = r.d.x; ray8[4] = r.d.y; ray8[5] = r.d.z; ray8[6] = r.tmin; ray8[7] = r.tmax;
}
extern "C" void refhost_infinite_le(const void* inf72, const float* d3, float* rad3) {
    Infinite inf; memcpy((void*)&inf, inf72, sizeof(Infinite));
    float3 r = inf.Le(make_float3(d3[0], d3[1], d3[2]));
    rad3[0] = r.x; rad3[1] = r.y; rad3[2] = r.z;
}
extern "C" void refhost_tonemap(const float* in3, int filmic, float* out3) {
    float3 c = make_float3(in3[0], in3[1], in3[2]);
    if (filmic) FilmicTonemapping(c); else GammaCorrection(c);
    out3[0] = c.x; out3[1] = c.y; out3[2] = c.z;
}

// ---- scene preparation through the reference's own code ---------------------------------------------------
// Camera ctor + Lookat exactly as the reference app does it: parsescene.cpp:162-176 fills config.camera via
// Lookat, then main.cpp:269 constructs Camera(pos,u,v,w,res,0.1,fov,aperture,focal,filmic,medium).
extern "C" void refhost_camera_make(void* cam104, const float* pos3, const float* lookat3, const float* up3,
                                    float resx, float resy, float distance, float fov, float aperture, float focal,
                                    int filmic, int environment, int medium) {
    Camera tmp;
    memset((void*)&tmp, 0, sizeof(Camera));
    tmp.Lookat(make_float3(pos3[0], pos3[1], pos3[2]), make_float3(lookat3[0], lookat3[1], lookat3[2]),
               make_float3(up3[0], up3[1], up3[2]));
    Camera* c = new Camera(tmp.position, tmp.u, tmp.v, tmp.w, make_float2(resx, resy), distance, fov, aperture, focal,
                           filmic != 0, medium);
    c->environment = environment != 0;
    // padding bytes of the 104-B struct are not defined by the ctor: normalise them to zero
    unsigned char raw[sizeof(Camera)]; memset(raw, 0, sizeof(raw));
    Camera* z = (Camera*)raw;
    *z = *c;   // member-wise copy keeps our zeroed padding only if the compiler copies member-wise; fix below
    memcpy(cam104, raw, sizeof(Camera));
    unsigned char* o = (unsigned char*)cam104;
    // bytes 74,75 (after the two bools at 72,73) are padding
    o[74] = 0; o[75] = 0;
    delete c;
}
// Scene::Init (src/scene.h:50-82): LoadOrBuildBVH -> build/split/flatten (src/bvh.cpp) + infinite.Init + light CDF.
// Outputs: prims_out (n x 176, leaf order), nodes_out (capacity 2n+1), lightdist_out (n_lights+2), box6, infinite72 (updated centre/radius).
extern "C" int refhost_scene_init(const void* prims_in, int n_prims, const void* lights, int n_lights,
                                  void* infinite72_inout, void* prims_out, void* nodes_out, int* n_nodes,
                                  float* lightdist_out, int* n_lightdist, float* box6) {
    Scene s;
    fill_vec(s.primitives, prims_in, n_prims);
    fill_vec(s.lights, lights, n_lights);
    if (infinite72_inout) memcpy((void*)&s.infinite, infinite72_inout, sizeof(Infinite));
    else { memset((void*)&s.infinite, 0, sizeof(Infinite)); s.infinite.isvalid = false; }
    char tmpl[] = "/tmp/refhost_bvh_XXXXXX";
    char* dir = mkdtemp(tmpl);
    if (!dir) return -1;
    std::string file = std::string(dir) + "/scene.json";     // bvh.cache lands next to it (src/bvh.cpp:189-191)
    Camera cam; memset((void*)&cam, 0, sizeof(cam));
    s.Init(&cam, file);
    unlink((std::string(dir) + "/bvh.cache").c_str());
    rmdir(dir);
    if ((int)s.bvh.prims.size() != n_prims) return -2;
    memcpy(prims_out, (const void*)s.bvh.prims.data(), sizeof(Primitive) * (size_t)n_prims);
    memcpy(nodes_out, (const void*)s.bvh.linear_root, sizeof(LinearBVHNode) * (size_t)s.bvh.total_nodes);
    *n_nodes = s.bvh.total_nodes;
    for (size_t i = 0; i < s.lightDistribution.size(); ++i) lightdist_out[i] = s.lightDistribution[i];
    *n_lightdist = (int)s.lightDistribution.size();
    box6[0] = s.bvh.root_box.fmin.x; box6[1] = s.bvh.root_box.fmin.y; box6[2] = s.bvh.root_box.fmin.z;
    box6[3] = s.bvh.root_box.fmax.x; box6[4] = s.bvh.root_box.fmax.y; box6[5] = s.bvh.root_box.fmax.z;
    if (infinite72_inout) memcpy(infinite72_inout, (const void*)&s.infinite, sizeof(Infinite));
    return 0;
}
// BVH::LoadOrBuildBVH (src/bvh.cpp:189-217) on a caller-chosen scene path: reads <dir>/bvh.cache if it exists, else
// builds and writes it.  Pins b200pt_bvh_cache_* (file written here loads there and vice versa).
extern "C" int refhost_bvh_load_or_build(const void* prims_in, int n_prims, const char* scene_file, void* prims_out,
                                         int prims_capacity, void* nodes_out, int nodes_capacity, int* n_prims_out,
                                         int* n_nodes, float* box6) {
    std::vector<Primitive> prims;
    fill_vec(prims, prims_in, n_prims);
    BVH bvh;
    bvh.LoadOrBuildBVH(prims, std::string(scene_file));
    if ((int)bvh.prims.size() > prims_capacity || bvh.total_nodes > nodes_capacity) return -2;
    memcpy(prims_out, (const void*)bvh.prims.data(), sizeof(Primitive) * bvh.prims.size());
    memcpy(nodes_out, (const void*)bvh.linear_root, sizeof(LinearBVHNode) * (size_t)bvh.total_nodes);
    *n_prims_out = (int)bvh.prims.size();
    *n_nodes = bvh.total_nodes;
    box6[0] = bvh.root_box.fmin.x; box6[1] = bvh.root_box.fmin.y; box6[2] = bvh.root_box.fmin.z;
    box6[3] = bvh.root_box.fmax.x; box6[4] = bvh.root_box.fmax.y; box6[5] = bvh.root_box.fmax.z;
    return 0;
}
extern "C" int refhost_sizeof(int which) {
    switch (which) {
        case 0: return sizeof(Camera); case 1: return sizeof(Primitive); case 2: return sizeof(LinearBVHNode);
        case 3: return sizeof(Material); case 4: return sizeof(Medium); case 5: return sizeof(Area);
        case 6: return sizeof(Infinite); case 7: return sizeof(Intersection); case 8: return sizeof(Ray);
    }
    return -1;
}
