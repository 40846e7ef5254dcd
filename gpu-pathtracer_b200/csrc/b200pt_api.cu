// b200pt_api.cu — C ABI (include/b200pt.h) of the wavefront path tracer: context creation (scene upload and
// re-layout, replaces BeginRender src/pathtracer.cu:2568-2695), the render loop (replaces Render :2705-2750)
// and teardown (EndRender :2697).  One context = one GPU = one CUDA stream; no global state.
#include "b200pt.h"
#include "ref_layouts.h"
#include "k_shade.cuh"
#include "k_trace.cuh"
#include "k_wave.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <thread>
#ifndef B200PT_EMULATE
#include <dlfcn.h>
#endif

using namespace pt;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
int b200pt_internal_fail(int code, const char* msg) { return fail(code, msg); }   // for bvh_build.cu
#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return fail(B200PT_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
    } while (0)

// One LANE = an independent wavefront (own stream, path pool, ray queue, counters, sample planes) over an
// interleaved subset of the context's screen tiles.  A context runs kLanes of them side by side on the same GPU:
// while one lane's traversal kernel (instruction-issue bound) drains, the other lane's shading kernel
// (latency bound) fills the machine, and the launch gaps of one chain hide behind the kernels of the other.
struct Lane {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_poll[2] = {nullptr, nullptr}, ev_done = nullptr;
    ShardMap map{};
    Pool pool{};
    RayQueue q{};
    Counters* counters = nullptr;
    Counters* h_counters = nullptr;       // pinned, 2 polling slots
    float4* samples = nullptr; size_t samples_cap = 0;   // in float4
    unsigned long long* h_init = nullptr; // pinned: initial {next_sample, done_samples} of a batch
    int pool_cap = 0;                     // slots allocated; pool.n = slots the current batch uses (<= pool_cap)
    uint16_t* sort_keys = nullptr; uint32_t *sort_hist = nullptr, *sort_offs = nullptr, *sort_out = nullptr;   // ray sorting (tree kernel)
    std::vector<void*> pool_allocs;       // cudaMalloc'ed pool/queue planes (re-allocated by the "pool" option)
    ShadeArgs sa; TraceArgs ta;
    int chunk = 0; bool done = false, exact_pending = false, nearly_done = false;
};

struct b200pt_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;        // lane 0's stream: uploads, resolve of lane 0, output copies
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr;
    SceneDev sc{};
    ShardMap map{};                       // the whole context's share of the image (rank-level shard)
    std::vector<Lane> lanes;
    std::vector<void*> allocs;            // everything else cudaMalloc'ed (freed in destroy)
    float *acc = nullptr, *color = nullptr, *out = nullptr;
    uint32_t width = 0, height = 0;
    int num_sms = 148, trace_blocks = 0;
    uint32_t stage_nodes = 0, stage_prims = 0;
    size_t stage_top_bytes = 20 * 1024;    // top-of-tree nodes staged per CTA when the scene does not fit
    int refill_below = 24;
    bool lambert_only = false;             // every referenced material is lambertian -> specialised shade kernel
    uint32_t mats_used = 0;                // MaterialTypes referenced by primitives (picks the k_shade instantiation)
    float4* leaves = nullptr; int n_leaves = 0;   // primitive groups (scenes with <= 256 primitives)
    bool wide = false;                     // tree kernel walks four-child nodes (WNode4)
    bool sort_rays = false;                // the tree kernel's queue is sorted by origin cell / direction octant before each trace step
    bool bin_materials = false;            // k_shade sorts its tile's records by material (scenes with more than one BSDF)
    int n_nodes4 = 0;
    bool small_scene = false;              // use k_trace_small
    bool fused = false;                    // CTA-local wavefront (k_wave_small): one persistent launch per batch and lane
    int wave_blocks = 0;
    uint32_t small_prim_bytes = 0;
    size_t max_batch_bytes = (size_t)2 << 30;
    int steps_per_poll = 8;
    int pool_total = 1 << 22;              // path slots ALLOCATED over all lanes; a batch uses as many as pay for it (batch_pool_slots)
    bool pool_explicit = false;            // set by the "pool" option: every batch uses exactly that many
    double stats[5] = {0, 0, 0, 0, 0};
    double total_ms = 0;
    bool vol = false;
    bool het = false;                      // `vpt` with a heterogeneous medium: the coroutine shade stage (k_het.cuh)
    int last_filmic = 1;
    // captured frame (CUDA graph) for the one-Render-per-frame usage: `pt`, spp == 1, every lane single-pass
    FrameParams* d_frame = nullptr;        // device copy read by the captured kernels
    FrameParams* h_frame = nullptr;        // pinned staging, refreshed before every graph launch
    bool use_graph = true;
#ifndef B200PT_EMULATE
    cudaGraphExec_t graph_exec = nullptr;
#endif
    double graph_launches = 0, graph_steps = 0;
    unsigned long long rays_seen = 0;      // device ray counters at the end of the previous render call
    // multi-GPU: NCCL communicator of this rank (opaque ncclComm_t) and, on the reduce root, the full-image sum
    void* comm = nullptr; int comm_rank = 0, comm_size = 1;
    float* reduced = nullptr;
    uint32_t reduced_iter = 0;
};

template <class T> static int dev_alloc(b200pt_ctx* c, T** p, size_t n, bool zero = false) {
    void* q = nullptr;
    size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e != cudaSuccess) return fail(B200PT_ENOMEM, std::string("cudaMalloc ") + std::to_string(bytes) + " B: " + cudaGetErrorString(e));
    if (zero) cudaMemsetAsync(q, 0, bytes, c->stream);
    c->allocs.push_back(q);
    *p = (T*)q;
    return 0;
}
template <class T> static int dev_upload(b200pt_ctx* c, T** p, const T* src, size_t n) {
    int rc = dev_alloc(c, p, n);
    if (rc) return rc;
    if (n) CK(cudaMemcpyAsync(*p, src, n * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    return 0;
}

// ---- device-side preparation: per-triangle constants with the SAME device arithmetic the reference uses -----
// normalize(dpdv) of Triangle::Intersect's epilogue (src/mesh.h:69-83) and Triangle::GetSurfaceArea (:39).
__global__ void k_prepare_shade(WShade* shade, const WPrim* prims, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    WShade& s = shade[i];
    if (s.type != 0) return;
    const WPrim& p = prims[i];
    f3 e1 = mk3(p.q0.w, p.q1.x, p.q1.y), e2 = mk3(p.q1.z, p.q1.w, p.q2.x);
    f2 duv1 = mk2(s.uv2[0], s.uv2[1]) - mk2(s.uv1[0], s.uv1[1]);
    f2 duv2 = mk2(s.uv3[0], s.uv3[1]) - mk2(s.uv1[0], s.uv1[1]);
    float det = duv1.x * duv2.y - duv1.y * duv2.x;
    f3 dpdu, dpdv;
    if (fabs((double)det) < 1e-8) {
        f3 nn = normalize(cross(e1, e2));
        make_coordinate(nn, dpdu, dpdv);
    } else {
        float invDet = 1 / det;
        dpdu = (duv2.y * e1 - duv1.y * e2) * invDet;
        dpdv = (-duv2.x * e1 + duv1.x * e2) * invDet;
    }
    f3 nd = normalize(dpdv);
    s.ndpdv[0] = nd.x; s.ndpdv[1] = nd.y; s.ndpdv[2] = nd.z;
}
__global__ void k_prepare_lights(WLight* lights, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    WLight& L = lights[i];
    f3 e1 = ld3(L.v2) - ld3(L.v1);
    f3 e2 = ld3(L.v3) - ld3(L.v1);
    L.area = length(cross(e1, e2)) * 0.5f;
}

static int build_scene(b200pt_ctx* c, const b200pt_scene_view* v) {
    if (v->integrator_type != B200PT_IT_PT && v->integrator_type != B200PT_IT_VPT)
        return fail(B200PT_EUNSUPPORTED, "only the `pt` and `vpt` integrators are on the hot path");
    if (v->n_prims <= 0 || v->n_nodes <= 0 || !v->prims || !v->nodes || !v->camera || !v->materials || !v->light_distribution)
        return fail(B200PT_EINVAL, "scene view is missing primitives / nodes / camera / materials / light distribution");
    if (v->n_textures > 0 && !v->textures) return fail(B200PT_EINVAL, "n_textures > 0 without a texture table");
    if (v->n_mediums > 254) return fail(B200PT_EINVAL, "too many media");
    const RefPrimitive* prims = (const RefPrimitive*)v->prims;
    const RefLinearBVHNode* nodes = (const RefLinearBVHNode*)v->nodes;
    const RefMaterial* mats = (const RefMaterial*)v->materials;
    for (int i = 0; i < v->n_materials; ++i)
        if (mats[i].textureIdx < -1 || mats[i].textureIdx >= v->n_textures) return fail(B200PT_EINVAL, "material texture index out of range");

    // -- primitives: intersection records + shading records
    std::vector<WPrim> wp(v->n_prims);
    std::vector<WShade> ws(v->n_prims);
    for (int i = 0; i < v->n_prims; ++i) {
        const RefPrimitive& p = prims[i];
        WPrim& q = wp[i]; WShade& s = ws[i];
        std::memset(&q, 0, sizeof(q)); std::memset(&s, 0, sizeof(s));
        if (p.type == REF_GT_TRIANGLE) {
            const RefTriangle& t = p.u.triangle;
            float e1[3], e2[3];
            for (int k = 0; k < 3; ++k) { e1[k] = t.v2.v[k] - t.v1.v[k]; e2[k] = t.v3.v[k] - t.v1.v[k]; }
            q.q0 = make_float4(t.v1.v[0], t.v1.v[1], t.v1.v[2], e1[0]);
            q.q1 = make_float4(e1[1], e1[2], e2[0], e2[1]);
            q.q2 = make_float4(e2[2], 0.f, 0.f, 0.f);
            std::memcpy(s.n1, t.v1.n, 12); std::memcpy(s.n2, t.v2.n, 12); std::memcpy(s.n3, t.v3.n, 12);
            std::memcpy(s.uv1, t.v1.uv, 8); std::memcpy(s.uv2, t.v2.uv, 8); std::memcpy(s.uv3, t.v3.uv, 8);
            s.matIdx = t.matIdx; s.lightIdx = t.lightIdx; s.mediumInside = t.mediumInside; s.mediumOutside = t.mediumOutside;
            s.type = 0;
            if (t.matIdx >= v->n_materials || t.matIdx < -1 || (t.matIdx < 0 && v->integrator_type == B200PT_IT_PT))
                return fail(B200PT_EINVAL, "triangle " + std::to_string(i) + " has material index " + std::to_string(t.matIdx) +
                                               " (a material-less medium boundary needs the vpt integrator)");
            if (t.lightIdx >= v->n_lights) return fail(B200PT_EINVAL, "triangle light index out of range");
        } else if (p.type == REF_GT_SPHERE) {
            const RefSphere& sp = p.u.sphere;
            q.q0 = make_float4(sp.origin[0], sp.origin[1], sp.origin[2], sp.radius);
            int one = 1; float onef; std::memcpy(&onef, &one, 4);
            q.q2 = make_float4(0.f, onef, 0.f, 0.f);
            std::memcpy(s.n1, sp.origin, 12); s.n2[0] = sp.radius;
            s.matIdx = sp.matIdx; s.lightIdx = -1; s.mediumInside = sp.mediumInside; s.mediumOutside = sp.mediumOutside;
            s.type = 1;
            if (sp.matIdx >= v->n_materials || sp.matIdx < -1 || (sp.matIdx < 0 && v->integrator_type == B200PT_IT_PT))
                return fail(B200PT_EINVAL, "sphere material index out of range");
        } else if (p.type == REF_GT_LINES) {
            // Line::Intersect leaves the hit's medium fields untouched (src/line.h:74-83), which Volpath would then read
            if (v->integrator_type != B200PT_IT_PT) return fail(B200PT_EUNSUPPORTED, "line primitives are only defined for the `pt` integrator");
            const RefLine& ln = p.u.line;
            int two = 2; float twof; std::memcpy(&twof, &two, 4);
            q.q0 = make_float4(ln.p0[0], ln.p0[1], ln.p0[2], ln.p1[0]);
            q.q1 = make_float4(ln.p1[1], ln.p1[2], ln.width0, ln.width1);
            q.q2 = make_float4(0.f, twof, 0.f, 0.f);
            s.matIdx = ln.matIdx; s.lightIdx = -1; s.mediumInside = -1; s.mediumOutside = -1;
            s.type = 2;
            if (ln.matIdx < 0 || ln.matIdx >= v->n_materials) return fail(B200PT_EINVAL, "line material index out of range");
        } else {
            return fail(B200PT_EINVAL, "unknown primitive type");
        }
        if (s.mediumInside >= v->n_mediums || s.mediumOutside >= v->n_mediums || s.mediumInside < -1 || s.mediumOutside < -1)
            return fail(B200PT_EINVAL, "medium index out of range");
        if (s.lightIdx < -1) return fail(B200PT_EINVAL, "light index out of range");
    }
    // -- nodes: reference DFS layout (left child = i+1, right = second_child_offset) -> two-child records
    // inner nodes are renumbered breadth-first, so that the first K records are the top of the tree (the part
    // k_trace stages into shared memory) and siblings/cousins share cache lines
    std::vector<int> inner_id(v->n_nodes, -1);
    int n_inner = 0;
    {
        std::vector<int> bfs; bfs.reserve(v->n_nodes);
        if (!nodes[0].is_leaf) bfs.push_back(0);
        for (size_t h = 0; h < bfs.size(); ++h) {
            const int i = bfs[h];
            inner_id[i] = n_inner++;
            const int l = i + 1, r = nodes[i].second_child_offset;
            if (l >= v->n_nodes || r <= 0 || r >= v->n_nodes) return fail(B200PT_EINVAL, "BVH node child index out of range");
            if (!nodes[l].is_leaf) bfs.push_back(l);
            if (!nodes[r].is_leaf) bfs.push_back(r);
            if ((int)bfs.size() > v->n_nodes) return fail(B200PT_EINVAL, "BVH node links form a cycle");
        }
    }
    auto mark_leaf = [&](const RefLinearBVHNode& n) -> int {
        if (n.start < 0 || n.end >= v->n_prims || n.end < n.start) return -1;
        int one = 1; float onef; std::memcpy(&onef, &one, 4);
        wp[n.end].q2.z = onef;                                   // last-in-leaf flag
        return 0;
    };
    std::vector<WNode> wn(std::max(n_inner, 1));
    std::memset(wn.data(), 0, wn.size() * sizeof(WNode));
    for (int i = 0; i < v->n_nodes; ++i) {
        const RefLinearBVHNode& n = nodes[i];
        if (n.is_leaf) { if (mark_leaf(n)) return fail(B200PT_EINVAL, "leaf node with an invalid primitive range"); continue; }
        if (inner_id[i] < 0) continue;                              // not reachable from the root
        int l = i + 1, r = n.second_child_offset;
        if (l >= v->n_nodes || r <= 0 || r >= v->n_nodes) return fail(B200PT_EINVAL, "BVH node child index out of range");
        const RefLinearBVHNode& L = nodes[l]; const RefLinearBVHNode& R = nodes[r];
        WNode& w = wn[inner_id[i]];
        w.q0 = make_float4(L.fmin[0], L.fmin[1], L.fmin[2], L.fmax[0]);
        w.q1 = make_float4(L.fmax[1], L.fmax[2], R.fmin[0], R.fmin[1]);
        w.q2 = make_float4(R.fmin[2], R.fmax[0], R.fmax[1], R.fmax[2]);
        w.link.x = L.is_leaf ? ~L.start : inner_id[l];
        w.link.y = R.is_leaf ? ~R.start : inner_id[r];
        w.link.z = 0; w.link.w = 0;
    }
    // -- four-child nodes (opt-in, B200PT_WIDE=1): collapse every other level of the reference tree.  A slot holds a
    // (grand)child's own box from the reference's array; an inner child is replaced by its two children while slots are
    // free, largest box first.  Bit-identical images, but measured SLOWER than the two-child records on every scene
    // (C4 97 vs 101 Msamples/s, C3 405 vs 432, hair 356 vs 368: profiles/r02m_wide.txt) — the traversal is bound by
    // instruction issue, not by dependent node fetches, and four exact slab tests per visit test ~1.4x the boxes.
    std::vector<WNode4> wn4;
    const char* wide_env = getenv("B200PT_WIDE");
    if (n_inner > 0 && wide_env && atoi(wide_env) != 0) {
        auto area_of = [&](int i) {
            const double dx = (double)nodes[i].fmax[0] - nodes[i].fmin[0], dy = (double)nodes[i].fmax[1] - nodes[i].fmin[1], dz = (double)nodes[i].fmax[2] - nodes[i].fmin[2];
            return dx * dy + dy * dz + dz * dx;
        };
        std::vector<int> q4{0};                         // reference indices of the inner nodes that become WNode4 records, BFS
        std::vector<int> id4(v->n_nodes, -1);
        id4[0] = 0;
        for (size_t hq = 0; hq < q4.size(); ++hq) {
            const int i = q4[hq];
            int kids[4] = {i + 1, nodes[i].second_child_offset, -1, -1};
            int nk = 2;
            while (nk < 4) {
                int pick = -1; double best = -1.0;
                for (int k = 0; k < nk; ++k)
                    if (!nodes[kids[k]].is_leaf) { const double ar = area_of(kids[k]); if (ar > best) { best = ar; pick = k; } }
                if (pick < 0) break;
                const int e = kids[pick];
                kids[pick] = e + 1; kids[nk++] = nodes[e].second_child_offset;
            }
            WNode4 w;
            std::memset(&w, 0, sizeof(w));
            float* mn[3] = {&w.minx.x, &w.miny.x, &w.minz.x}; float* mx[3] = {&w.maxx.x, &w.maxy.x, &w.maxz.x};
            int* lk = &w.link.x;
            for (int k = 0; k < 4; ++k) {
                if (k >= nk) { lk[k] = kEmptyChild; continue; }
                const RefLinearBVHNode& ch = nodes[kids[k]];
                for (int ax = 0; ax < 3; ++ax) { mn[ax][k] = ch.fmin[ax]; mx[ax][k] = ch.fmax[ax]; }
                if (ch.is_leaf) lk[k] = ~ch.start;
                else { id4[kids[k]] = (int)q4.size(); lk[k] = (int)q4.size(); q4.push_back(kids[k]); }
            }
            wn4.push_back(w);
        }
        // (links were assigned in BFS order, records were appended in BFS order: record j describes q4[j])
    }
    c->n_nodes4 = (int)wn4.size();

    // boxes a ray must pass to reach an emitter (mis_ray_may_reach_emitter, wavefront.cuh): the BVH leaves and, for a small
    // scene, the primitive groups that hold a triangle with a light index — the union, so the list is valid whichever
    // traversal kernel the context ends up using
    std::vector<float4> emit;
    auto add_emit_box = [&](const float* mn, const float* mx) {
        const float4 q0 = make_float4(mn[0], mn[1], mn[2], mx[0]), q1 = make_float4(mx[1], mx[2], 0.f, 0.f);
        for (size_t i = 0; i + 1 < emit.size(); i += 2)
            if (!std::memcmp(&emit[i], &q0, 16) && !std::memcmp(&emit[i + 1], &q1, 8)) return;
        emit.push_back(q0); emit.push_back(q1);
    };
    for (int i = 0; i < v->n_nodes; ++i) {
        if (!nodes[i].is_leaf) continue;
        bool has = false;
        for (int k = nodes[i].start; k <= nodes[i].end && k < v->n_prims; ++k) has = has || ws[k].lightIdx >= 0;
        if (has) add_emit_box(nodes[i].fmin, nodes[i].fmax);
    }
    // primitive groups for k_trace_small (<= 256 primitives): greedy agglomeration of tight primitive boxes under the
    // cost model  cost(group) = C_BOX + P(ray hits box) * C_PRIM * |group|,  P ~ surface area of the box / root's
    if (v->n_prims <= 256) {
        struct Grp { float mn[3], mx[3]; int n; int p[4]; };
        std::vector<Grp> g(v->n_prims);
        for (int i = 0; i < v->n_prims; ++i) {
            const RefPrimitive& p = prims[i];
            Grp& G = g[i]; G.n = 1; G.p[0] = i;
            if (p.type == REF_GT_TRIANGLE) {
                const float* vv[3] = {p.u.triangle.v1.v, p.u.triangle.v2.v, p.u.triangle.v3.v};
                for (int k = 0; k < 3; ++k) {
                    G.mn[k] = std::min(vv[0][k], std::min(vv[1][k], vv[2][k]));
                    G.mx[k] = std::max(vv[0][k], std::max(vv[1][k], vv[2][k]));
                }
            } else if (p.type == REF_GT_LINES) {           // Line::GetBBox, src/line.h:16
                const RefLine& ln = p.u.line;
                const float mw = ln.width0 > ln.width1 ? ln.width0 : ln.width1;
                for (int k = 0; k < 3; ++k) {
                    G.mn[k] = std::nextafter(std::min(ln.p0[k], ln.p1[k]) - mw, -INFINITY);
                    G.mx[k] = std::nextafter(std::max(ln.p0[k], ln.p1[k]) + mw, INFINITY);
                }
            } else {
                for (int k = 0; k < 3; ++k) { G.mn[k] = p.u.sphere.origin[k] - p.u.sphere.radius; G.mx[k] = p.u.sphere.origin[k] + p.u.sphere.radius; }
                // one ulp of slack per side: centre -/+ radius is itself rounded
                for (int k = 0; k < 3; ++k) { G.mn[k] = std::nextafter(G.mn[k], -INFINITY); G.mx[k] = std::nextafter(G.mx[k], INFINITY); }
            }
        }
        auto area = [](const float* mn, const float* mx) {
            double dx = (double)mx[0] - mn[0], dy = (double)mx[1] - mn[1], dz = (double)mx[2] - mn[2];
            return 2.0 * (dx * dy + dy * dz + dz * dx);
        };
        const double root_area = std::max(area(nodes[0].fmin, nodes[0].fmax), 1e-30);
        const double C_BOX = 22.0, C_PRIM = 186.0;
        auto cost = [&](const float* mn, const float* mx, int n) { return C_BOX + std::min(1.0, area(mn, mx) / root_area) * C_PRIM * n; };
        for (;;) {
            double best = 0.0; int bi = -1, bj = -1;
            const bool forced = g.size() > 64;
            for (size_t i = 0; i < g.size(); ++i)
                for (size_t j = i + 1; j < g.size(); ++j) {
                    if (g[i].n + g[j].n > 4) continue;
                    float mn[3], mx[3];
                    for (int k = 0; k < 3; ++k) { mn[k] = std::min(g[i].mn[k], g[j].mn[k]); mx[k] = std::max(g[i].mx[k], g[j].mx[k]); }
                    const double delta = cost(mn, mx, g[i].n + g[j].n) - cost(g[i].mn, g[i].mx, g[i].n) - cost(g[j].mn, g[j].mx, g[j].n);
                    if (bi < 0 || delta < best) { best = delta; bi = (int)i; bj = (int)j; }
                }
            if (bi < 0 || (!forced && best >= 0.0)) break;
            Grp& A = g[bi]; const Grp& B = g[bj];
            for (int k = 0; k < 3; ++k) { A.mn[k] = std::min(A.mn[k], B.mn[k]); A.mx[k] = std::max(A.mx[k], B.mx[k]); }
            for (int k = 0; k < B.n; ++k) A.p[A.n++] = B.p[k];
            g.erase(g.begin() + bj);
        }
        if (g.size() <= 64) {
            std::vector<float4> wl;
            for (const Grp& G : g) {
                int idx[4] = {0xffff, 0xffff, 0xffff, 0xffff};
                for (int k = 0; k < G.n; ++k) idx[k] = G.p[k];
                std::sort(idx, idx + G.n);
                const uint32_t u0 = (uint32_t)idx[0] | ((uint32_t)idx[1] << 16), u1 = (uint32_t)idx[2] | ((uint32_t)idx[3] << 16);
                float f0, f1; std::memcpy(&f0, &u0, 4); std::memcpy(&f1, &u1, 4);
                wl.push_back(make_float4(G.mn[0], G.mn[1], G.mn[2], G.mx[0]));
                wl.push_back(make_float4(G.mx[1], G.mx[2], f0, f1));
            }
            for (const Grp& G : g) {
                bool has = false;
                for (int k = 0; k < G.n; ++k) has = has || ws[G.p[k]].lightIdx >= 0;
                if (has) add_emit_box(G.mn, G.mx);
            }
            c->n_leaves = (int)g.size();
            int rc2 = dev_upload(c, &c->leaves, wl.data(), wl.size());
            if (rc2) return rc2;
        }
    }
    SceneDev& sc = c->sc;
    std::memcpy(sc.root_min, nodes[0].fmin, 12); std::memcpy(sc.root_max, nodes[0].fmax, 12);
    sc.root_leaf_count = nodes[0].is_leaf ? (nodes[0].end - nodes[0].start + 1) : 0;
    sc.n_nodes = n_inner; sc.n_prims = v->n_prims;
    WNode* d_nodes; WPrim* d_prims; WShade* d_shade;
    int rc;
    if ((rc = dev_upload(c, &d_nodes, wn.data(), wn.size()))) return rc;
    if ((rc = dev_upload(c, &d_prims, wp.data(), wp.size()))) return rc;
    if ((rc = dev_upload(c, &d_shade, ws.data(), ws.size()))) return rc;
    sc.nodes = d_nodes; sc.prims = d_prims; sc.shade = d_shade;
    sc.nodes4 = nullptr;
    if (!wn4.empty()) {
        WNode4* d_nodes4;
        if ((rc = dev_upload(c, &d_nodes4, wn4.data(), wn4.size()))) return rc;
        sc.nodes4 = d_nodes4;
    }
    c->wide = sc.nodes4 != nullptr;
    PT_LAUNCH(k_prepare_shade, (v->n_prims + 255) / 256, 256, 0, c->stream, d_shade, d_prims, v->n_prims);

    // -- lights
    std::vector<WLight> wl(std::max(v->n_lights, 1));
    std::memset(wl.data(), 0, wl.size() * sizeof(WLight));
    const RefArea* areas = (const RefArea*)v->lights;
    for (int i = 0; i < v->n_lights; ++i) {
        const RefTriangle& t = areas[i].triangle;
        std::memcpy(wl[i].v1, t.v1.v, 12); std::memcpy(wl[i].v2, t.v2.v, 12); std::memcpy(wl[i].v3, t.v3.v, 12);
        std::memcpy(wl[i].n1, t.v1.n, 12); std::memcpy(wl[i].n2, t.v2.n, 12); std::memcpy(wl[i].n3, t.v3.n, 12);
        std::memcpy(wl[i].radiance, areas[i].radiance, 12);
    }
    WLight* d_lights;
    if ((rc = dev_upload(c, &d_lights, wl.data(), wl.size()))) return rc;
    if (v->n_lights) PT_LAUNCH(k_prepare_lights, (v->n_lights + 127) / 128, 128, 0, c->stream, d_lights, v->n_lights);
    sc.lights = d_lights; sc.n_lights = v->n_lights;

    // -- materials (72-B reference records are used as they are), media, light CDF
    Material* d_mats;
    if ((rc = dev_upload(c, &d_mats, (const Material*)v->materials, (size_t)v->n_materials))) return rc;
    sc.mats = d_mats; sc.n_mats = v->n_materials;
    // material set actually referenced by primitives (scene files often define materials they do not use)
    c->mats_used = 0u;
    for (int i = 0; i < v->n_prims; ++i) {
        const int m = ws[i].matIdx;
        if (m >= 0) c->mats_used |= 1u << (mats[m].type & 31);
    }
    c->lambert_only = (c->mats_used & ~kMatsLambertOnly) == 0u;
    {   // shade-stage sort key per primitive (k_shade's material binning)
        std::vector<unsigned char> key((size_t)std::max(v->n_prims, 1));
        for (int i = 0; i < v->n_prims; ++i) {
            const int m = ws[i].matIdx;
            key[i] = (unsigned char)((m < 0 ? 7 : std::min(mats[m].type & 31, 6)) + (ws[i].lightIdx >= 0 ? 8 : 0));
        }
        unsigned char* d_key;
        if ((rc = dev_upload(c, &d_key, key.data(), key.size()))) return rc;
        sc.prim_key = d_key;
    }
    // measured (profiles/r02n_bin_chunk.txt): six BSDFs in one scene 102 -> 221 Msamples/s (`pt`), 120 -> 214 (`vpt`) in the
    // CTA-local kernel; with two BSDFs (C5: lambertian walls + one glass sphere) the sort costs 5 %; the HBM-pool k_shade
    // of C3 / the hair scene is bound by its pool round trip, not by divergence, and does not move (426 vs 425, 368 vs 368)
    int n_types = 0;
    for (uint32_t m = c->mats_used; m; m &= m - 1u) ++n_types;
    c->bin_materials = n_types >= 3;
    {   // MIS-ray culling: needs at least one emitter, no environment light (a ray that escapes then carries radiance), a short list
        bool on = !emit.empty() && (int)emit.size() / 2 <= kMaxEmitBoxes && !(v->infinite && ((const RefInfinite*)v->infinite)->isvalid);
        if (const char* env = getenv("B200PT_CULL_MIS")) on = on && atoi(env) != 0;
        sc.emit_boxes = nullptr; sc.n_emit_boxes = 0;
        if (on) {
            float4* d_emit;
            if ((rc = dev_upload(c, &d_emit, emit.data(), emit.size()))) return rc;
            sc.emit_boxes = d_emit; sc.n_emit_boxes = (int)emit.size() / 2;
        }
    }
    if (const char* env = getenv("B200PT_BIN_MATERIALS")) c->bin_materials = atoi(env) != 0;
    std::vector<WMedium> wm(std::max(v->n_mediums, 1));
    std::memset(wm.data(), 0, wm.size() * sizeof(WMedium));
    const RefMedium* med = (const RefMedium*)v->mediums;
    std::vector<WHetero> wh(std::max(v->n_mediums, 1));
    std::memset(wh.data(), 0, wh.size() * sizeof(WHetero));
    bool any_het = false;
    for (int i = 0; i < v->n_mediums; ++i) {
        std::memcpy(wm[i].sigmaA, med[i].sigmaA, 12); std::memcpy(wm[i].sigmaS, med[i].sigmaS, 12); std::memcpy(wm[i].sigmaT, med[i].sigmaT, 12);
        wm[i].g = med[i].g; wm[i].type = 0;
        if (med[i].type == REF_MEDIUM_HOMOGENEOUS) continue;
        if (med[i].type != REF_MEDIUM_HETEROGENEOUS) return fail(B200PT_EINVAL, "unknown medium type");
        // Heterogeneous (src/medium.h:52): the density grid is deep-copied to the device (src/pathtracer.cu:2612-2622
        // copies it and delete[]s the caller's array; this library never touches caller memory)
        if (med[i].nx <= 0 || med[i].ny <= 0 || med[i].nz <= 0 || !med[i].density) return fail(B200PT_EINVAL, "heterogeneous medium without a density grid");
        if (med[i].evalTransmittanceType < 0 || med[i].evalTransmittanceType > 2) return fail(B200PT_EINVAL, "evalTransmittanceType must be 0, 1 or 2");
        float* d_density;
        if ((rc = dev_upload(c, &d_density, med[i].density, (size_t)med[i].nx * med[i].ny * med[i].nz))) return rc;
        wm[i].type = 1;
        wh[i].density = d_density; wh[i].nx = med[i].nx; wh[i].ny = med[i].ny; wh[i].nz = med[i].nz;
        wh[i].invMaxDensity = med[i].invMaxDensity;
        std::memcpy(wh[i].p0, med[i].p0, 12); std::memcpy(wh[i].p1, med[i].p1, 12);
        wh[i].iterMax = med[i].iterMax; wh[i].evalTransmittanceType = med[i].evalTransmittanceType;
        any_het = true;
    }
    sc.het = nullptr;
    // (`pt` never looks at media — Path, src/pathtracer.cu:880-1021 — so only `vpt` needs the sequential kernel)
    if (any_het && v->integrator_type == B200PT_IT_VPT) {
        WHetero* d_het;
        if ((rc = dev_upload(c, &d_het, wh.data(), wh.size()))) return rc;
        sc.het = d_het;
        c->het = true;
    }
    WMedium* d_med;
    if ((rc = dev_upload(c, &d_med, wm.data(), wm.size()))) return rc;
    sc.mediums = d_med; sc.n_mediums = v->n_mediums;
    float* d_cdf;
    // one trailing pad entry: the reference's scan reads cdf[n] in its last iteration (src/pathtracer.cu:175)
    std::vector<float> cdf(v->light_distribution, v->light_distribution + v->n_light_distribution);
    cdf.push_back(cdf.empty() ? 0.f : cdf.back());
    if ((rc = dev_upload(c, &d_cdf, cdf.data(), cdf.size()))) return rc;
    sc.cdf = d_cdf; sc.n_cdf = v->n_light_distribution;

    // -- textures: all uchar4 texels back to back + {first texel, w, h} per texture (src/pathtracer.cu:2646-2661)
    sc.texels = nullptr; sc.tex_info = nullptr;
    if (v->n_textures > 0) {
        std::vector<int4> info(v->n_textures);
        std::vector<unsigned char> texels;
        for (int i = 0; i < v->n_textures; ++i) {
            const b200pt_texture& t = v->textures[i];
            if (!t.texels || t.width <= 0 || t.height <= 0) return fail(B200PT_EINVAL, "texture without texels");
            info[i].x = (int)(texels.size() / 4); info[i].y = t.width; info[i].z = t.height; info[i].w = 0;
            const unsigned char* src = (const unsigned char*)t.texels;
            texels.insert(texels.end(), src, src + 4 * (size_t)t.width * t.height);
        }
        unsigned char* d_tex; int4* d_info;
        if ((rc = dev_upload(c, &d_tex, texels.data(), texels.size()))) return rc;
        if ((rc = dev_upload(c, &d_info, info.data(), info.size()))) return rc;
        CK(cudaStreamSynchronize(c->stream));
        sc.texels = d_tex; sc.tex_info = d_info;
    }

    // -- infinite light
    std::memset(&sc.inf, 0, sizeof(sc.inf));
    if (v->infinite) {
        const RefInfinite* inf = (const RefInfinite*)v->infinite;
        if (inf->isvalid) {
            if (!inf->data || inf->width <= 0 || inf->height <= 0) return fail(B200PT_EINVAL, "infinite light without texels");
            float* d_tex;
            if ((rc = dev_upload(c, &d_tex, inf->data, (size_t)3 * inf->width * inf->height))) return rc;
            sc.inf.data = d_tex; sc.inf.width = inf->width; sc.inf.height = inf->height;
            std::memcpy(sc.inf.center, inf->center, 12); sc.inf.radius = inf->radius;
            std::memcpy(sc.inf.u, inf->u, 12); std::memcpy(sc.inf.v, inf->v, 12); std::memcpy(sc.inf.w, inf->w, 12);
            sc.inf.isvalid = 1;
        }
    }
    // consistency of the CDF with the light list (idx == n_lights selects the infinite light, :931)
    int expect = 1 + v->n_lights + (sc.inf.isvalid ? 1 : 0);
    if (v->n_light_distribution != expect)
        return fail(B200PT_EINVAL, "light distribution has " + std::to_string(v->n_light_distribution) + " entries, expected " + std::to_string(expect));
    sc.integrator = v->integrator_type; sc.max_depth = v->max_depth;
    // (maxDepth 0: the reference's `for (bounces = 0; bounces < maxDepth; ...)` never runs and the image is black; a
    // wavefront slot always traces its primary ray, so that degenerate setting is rejected instead of approximated)
    if (v->max_depth < 1 || v->max_depth > 127) return fail(B200PT_EINVAL, "maxDepth must be in [1, 127]");
    c->vol = v->integrator_type == B200PT_IT_VPT;

    // staging of the acceleration structure into shared memory (TMA bulk copy): only when small
    // (TMA bulk copy): everything when the scene is small, else the top of the breadth-first node array
    size_t nb = c->wide ? (size_t)c->n_nodes4 * sizeof(WNode4) : (size_t)n_inner * sizeof(WNode), pb = (size_t)v->n_prims * sizeof(WPrim);
    // (k_trace also keeps 24 KB of traversal stack in shared memory; 20 KB of structure keeps 5 CTAs per SM resident)
    if (const char* env = getenv("B200PT_STAGE_BYTES")) c->stage_top_bytes = (size_t)std::max(0, atoi(env));
    c->stage_top_bytes = std::min<size_t>(c->stage_top_bytes, 200 * 1024 - kTraceStackBytes) & ~(size_t)63;   // 227 KB per CTA on sm_100
    if (nb + pb <= c->stage_top_bytes) { c->stage_nodes = (uint32_t)nb; c->stage_prims = (uint32_t)pb; }
    else { c->stage_nodes = (uint32_t)std::min<size_t>(nb, c->stage_top_bytes); c->stage_prims = 0; }
    c->small_prim_bytes = (uint32_t)pb;
    c->small_scene = c->leaves != nullptr && pb + 64 * 32 <= 40 * 1024;
    CK(cudaStreamSynchronize(c->stream));     // host staging vectors go out of scope
    return 0;
}

// k_trace instantiation for this context (volumetric or not, two- or four-child nodes)
template <class F> static void with_trace_kernel(const b200pt_ctx* c, F&& f) {
    if (c->vol) { if (c->wide) f(k_trace<true, true>); else f(k_trace<true, false>); }
    else { if (c->wide) f(k_trace<false, true>); else f(k_trace<false, false>); }
}
// k_wave_small instantiation for this context (volumetric or not, referenced material set, heterogeneous media)
template <class F> static void with_wave_kernel(const b200pt_ctx* c, F&& f) {
    const bool ldc = (c->mats_used & ~kMatsLDC) == 0u;
    if (c->het) {
        if (c->lambert_only) f(k_wave_small<true, kMatsLambertOnly, true>, true);
        else f(k_wave_small<true, kMatsAll, true>, true);
    } else if (c->vol) {
        if (c->lambert_only) f(k_wave_small<true, kMatsLambertOnly, false>, true);
        else if (ldc) f(k_wave_small<true, kMatsLDC, false>, true);
        else f(k_wave_small<true, kMatsAll, false>, true);
    } else {
        if (c->lambert_only) f(k_wave_small<false, kMatsLambertOnly, false>, false);
        else if (ldc) f(k_wave_small<false, kMatsLDC, false>, false);
        else f(k_wave_small<false, kMatsAll, false>, false);
    }
}
static int wave_threads(const b200pt_ctx* c) {
    if (c->het) return WaveThreads<true>::value;
    const bool all_mats = !c->lambert_only && (c->mats_used & ~kMatsLDC) != 0u;       // the instantiation with_wave_kernel picks
    if (c->lambert_only && !c->vol) return WaveThreads<false, kMatsLambertOnly, false>::value;
    return all_mats ? WaveThreads<false, kMatsAll>::value : WaveThreads<false>::value;
}
static size_t wave_smem(const b200pt_ctx* c) {
    return c->vol ? wave_smem_bytes<true>(c->small_prim_bytes, c->n_leaves, wave_threads(c)) : wave_smem_bytes<false>(c->small_prim_bytes, c->n_leaves, wave_threads(c));
}
static void free_pool(Lane& L) {
    for (void* p : L.pool_allocs) cudaFree(p);
    L.pool_allocs.clear();
    L.pool = Pool{}; L.q.entries = nullptr; L.pool_cap = 0;
    L.sort_keys = nullptr; L.sort_hist = L.sort_offs = L.sort_out = nullptr;
}
static int alloc_pool(b200pt_ctx* c, Lane& L, int n) {
    free_pool(L);
    Pool& p = L.pool;
    float4** arrs[] = {&p.o_rng, &p.d_flags, &p.beta_s, &p.li_t, &p.shd, &p.misd, &p.ldl, &p.misf, &p.beta_old, &p.hit0, &p.hit1, &p.vis, &p.aux, &p.pend_o, &p.carry};
    n = (n + 255) & ~255;
    auto grab = [&](void** q, size_t bytes) -> int {
        cudaError_t e = cudaMalloc(q, bytes);
        if (e != cudaSuccess) return fail(B200PT_ENOMEM, std::string("cudaMalloc ") + std::to_string(bytes) + " B: " + cudaGetErrorString(e));
        cudaMemsetAsync(*q, 0, bytes, L.stream);
        L.pool_allocs.push_back(*q);
        return 0;
    };
    for (auto a : arrs) { int rc = grab((void**)a, (size_t)n * sizeof(float4)); if (rc) return rc; }
    p.n = n; L.pool_cap = n;
    int rc = grab((void**)&L.q.entries, (size_t)3 * n * sizeof(uint32_t));     // at most three rays per slot and step
    if (rc || !c->sort_rays) return rc;
    if ((rc = grab((void**)&L.sort_keys, (size_t)3 * n * sizeof(uint16_t)))) return rc;
    if ((rc = grab((void**)&L.sort_out, (size_t)3 * n * sizeof(uint32_t)))) return rc;
    if ((rc = grab((void**)&L.sort_hist, (size_t)kRaySortBins * sizeof(uint32_t)))) return rc;
    return grab((void**)&L.sort_offs, (size_t)kRaySortBins * sizeof(uint32_t));
}
// Pool slots per lane: the context total split evenly, never more than 8 slots per pixel of the lane.
static int lane_pool_size(const b200pt_ctx* c, const Lane& L) {
    int pool = std::max(1024, c->pool_total / (int)c->lanes.size());
    // a lane whose pixels almost fit gets one slot per pixel: a 1-spp Render call then runs in exact-step mode
    if (L.map.n_local_pixels > pool && (double)L.map.n_local_pixels <= 1.1 * pool) pool = L.map.n_local_pixels;
    if ((size_t)pool > (size_t)L.map.n_local_pixels * 8) pool = std::max(1024, L.map.n_local_pixels * 8);
    return pool;
}

// The ONE place sample planes are (re)allocated.  A captured frame graph has the old pointer baked into its kernel
// arguments, so it is destroyed (after a device sync) whenever the planes move.
static int ensure_samples(b200pt_ctx* c, Lane& L, size_t need) {
    if (need <= L.samples_cap && L.samples) return 0;
    need = std::max<size_t>(need, 1);
#ifndef B200PT_EMULATE
    if (c->graph_exec) { CK(cudaDeviceSynchronize()); cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
#endif
    if (L.samples) { CK(cudaStreamSynchronize(L.stream)); cudaFree(L.samples); L.samples = nullptr; L.samples_cap = 0; }
    cudaError_t e = cudaMalloc((void**)&L.samples, need * sizeof(float4));
    if (e != cudaSuccess) return fail(B200PT_ENOMEM, std::string("cudaMalloc of the sample planes: ") + cudaGetErrorString(e));
    L.samples_cap = need;
    return 0;
}

// `owner_of(k)` decides which tiles this map holds: rank-level shards use shard_owner(k, n); the lanes of a context
// split the rank's tiles once more among themselves (j-th tile of the rank -> lane j % n_lanes).
static int fill_map(b200pt_ctx* c, ShardMap& m, uint32_t width, uint32_t height, int shard, int n_shards, int tile_w, int tile_h,
                    int lane = 0, int n_lanes = 1) {
    m.width = (int)width; m.height = (int)height;
    m.shard = shard * n_lanes + lane; m.n_shards = n_shards * n_lanes; m.tile_w = tile_w; m.tile_h = tile_h;
    m.tiles_x = (int)width / tile_w; m.tiles_y = (int)height / tile_h;
    m.tiles = nullptr;
    const int n_tiles = m.tiles_x * m.tiles_y;
    if (m.n_shards == 1) { m.n_local_tiles = n_tiles; m.n_local_pixels = (int)(width * height); return 0; }
    if (m.tiles_x > 0xffff || m.tiles_y > 0xffff) return fail(B200PT_EINVAL, "too many tiles");
    std::vector<uint32_t> mine;
    int j = 0;
    for (int k = 0; k < n_tiles; ++k) {
        if (shard_owner(k, n_shards) != shard) continue;
        if (j++ % n_lanes == lane) mine.push_back((uint32_t)(k % m.tiles_x) | ((uint32_t)(k / m.tiles_x) << 16));
    }
    m.n_local_tiles = (int)mine.size();
    m.n_local_pixels = m.n_local_tiles * tile_w * tile_h;
    uint32_t* d = nullptr;
    int rc = dev_upload(c, &d, mine.data(), mine.size());
    if (rc) return rc;
    CK(cudaStreamSynchronize(c->stream));
    m.tiles = d;
    return 0;
}

extern "C" int b200pt_create(const b200pt_scene_view* scene, uint32_t width, uint32_t height, float epsilon,
                             int device, const b200pt_shard* shard, b200pt_ctx** out_ctx) {
    if (!scene || !out_ctx) return fail(B200PT_EINVAL, "null argument");
    if (width == 0 || height == 0 || width % 32 || height % 4)
        return fail(B200PT_EINVAL, "width must be a multiple of 32 and height a multiple of 4 (reference launch grid, src/pathtracer.cu:2707-2709)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(B200PT_ECUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " — this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(B200PT_EINVAL, "device index out of range");
    CK(cudaSetDevice(device));
    b200pt_ctx* c = new b200pt_ctx();
    c->device = device; c->width = width; c->height = height;
    auto bail = [&](int rc) { std::string keep = g_err; b200pt_destroy(c); g_err = keep; return rc; };

    // rank-level shard of the image
    int sh = 0, nsh = 1, tw = 32, th = 32;
    if (shard && shard->n_shards > 1) {
        if (shard->shard < 0 || shard->shard >= shard->n_shards || shard->tile_w <= 0 || shard->tile_h <= 0 ||
            width % shard->tile_w || height % shard->tile_h)
            return bail(fail(B200PT_EINVAL, "bad shard description (tiles must divide the image)"));
        sh = shard->shard; nsh = shard->n_shards; tw = shard->tile_w; th = shard->tile_h;
    }

    // lane 0's stream carries the scene upload
    c->lanes.resize(1);
    if (cudaStreamCreateWithFlags(&c->lanes[0].stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(B200PT_ECUDA, "stream create failed"));
    c->stream = c->lanes[0].stream;
    cudaEventCreate(&c->ev0); cudaEventCreate(&c->ev1);
    cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->num_sms = prop.multiProcessorCount;
    c->sc.eps = epsilon;
    int rc = build_scene(c, scene);
    if (rc) return bail(rc);
    if ((rc = fill_map(c, c->map, width, height, sh, nsh, tw, th))) return bail(rc);

    // lanes: local tile j of this context goes to lane j % n_lanes, i.e. lane k is shard (sh + nsh * k) of nsh * n_lanes.
    // Measured (profiles/r01q_lanes.txt): 3 lanes are best when the flat small-scene kernel traces (C2 +5 %), 2 otherwise.
    // scenes whose primitives fit in shared memory run the CTA-local wavefront (k_wave.cuh); B200PT_FUSED=0 keeps the
    // global wavefront (k_shade + k_trace_small over the HBM pool) for A/B runs
    c->fused = c->small_scene;
    // ray sorting before the tree kernel (k_trace.cuh): opt-in, measured slower (profiles/r03c_sort_rays.txt)
    c->sort_rays = false;
    if (const char* env = getenv("B200PT_SORT_RAYS")) c->sort_rays = !c->small_scene && atoi(env) != 0;
    if (const char* env = getenv("B200PT_FUSED")) c->fused = c->fused && atoi(env) != 0;
    int n_lanes = c->fused ? 1 : (c->small_scene ? 3 : 2);
    if (const char* env = getenv("B200PT_LANES")) n_lanes = std::max(1, std::min(8, atoi(env)));
    if (width % tw || height % th || c->map.n_local_tiles < n_lanes) n_lanes = 1;
    c->lanes.resize(n_lanes);
    for (int k = 0; k < n_lanes; ++k) {
        Lane& L = c->lanes[k];
        if (k && cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(B200PT_ECUDA, "stream create failed"));
        cudaEventCreateWithFlags(&L.ev_poll[0], cudaEventDisableTiming); cudaEventCreateWithFlags(&L.ev_poll[1], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&L.ev_done, cudaEventDisableTiming);
        if (n_lanes == 1) L.map = c->map;
        else if ((rc = fill_map(c, L.map, width, height, sh, nsh, tw, th, k, n_lanes))) return bail(rc);
    }

    size_t npix = (size_t)width * height;
    if ((rc = dev_alloc(c, &c->acc, 3 * npix, true))) return bail(rc);
    if ((rc = dev_alloc(c, &c->color, 3 * npix, true))) return bail(rc);
    if ((rc = dev_alloc(c, &c->out, 3 * npix, true))) return bail(rc);
    if ((rc = dev_alloc(c, &c->d_frame, 1, true))) return bail(rc);
    if (cudaMallocHost((void**)&c->h_frame, sizeof(FrameParams)) != cudaSuccess) return bail(fail(B200PT_ENOMEM, "pinned alloc failed"));
    if (const char* env = getenv("B200PT_GRAPH")) c->use_graph = atoi(env) != 0;
    for (Lane& L : c->lanes) {
        if ((rc = dev_alloc(c, &L.counters, 1, true))) return bail(rc);
        if ((rc = dev_alloc(c, &L.q.ctl, 1, true))) return bail(rc);
        if (cudaMallocHost((void**)&L.h_counters, 2 * sizeof(Counters)) != cudaSuccess) return bail(fail(B200PT_ENOMEM, "pinned alloc failed"));
        if (cudaMallocHost((void**)&L.h_init, 2 * sizeof(unsigned long long)) != cudaSuccess) return bail(fail(B200PT_ENOMEM, "pinned alloc failed"));
        std::memset(L.h_counters, 0, 2 * sizeof(Counters));
        if ((rc = alloc_pool(c, L, lane_pool_size(c, L)))) return bail(rc);
    }

    // persistent traversal grid: resident CTAs per SM x SM count
    int per_sm = 0;
    size_t smem = (size_t)c->stage_nodes + c->stage_prims + kTraceStackBytes;
    if (c->small_scene) {
        smem = (size_t)c->small_prim_bytes + (size_t)c->n_leaves * 32;
        if (c->vol) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace_small<true>, kTraceThreads, smem);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace_small<false>, kTraceThreads, smem);
    } else with_trace_kernel(c, [&](auto kernel) { cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kTraceThreads, smem); });
#ifndef B200PT_EMULATE
    if (smem > 48 * 1024) {      // above the default dynamic shared-memory limit: opt in, then ask again
        cudaError_t e = cudaSuccess;
        with_trace_kernel(c, [&](auto kernel) { e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); });
        if (e != cudaSuccess) return bail(fail(B200PT_ECUDA, std::string("shared-memory opt-in failed: ") + cudaGetErrorString(e)));
        with_trace_kernel(c, [&](auto kernel) { cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kTraceThreads, smem); });
    }
#endif
    if (per_sm <= 0) per_sm = 1;
    c->trace_blocks = c->num_sms * per_sm;
    if (c->fused) {
        const size_t wsmem = wave_smem(c);
        int wave_per_sm = 0;
        cudaError_t we = cudaSuccess;
        with_wave_kernel(c, [&](auto kernel, bool) {
#ifndef B200PT_EMULATE
            we = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem);
#endif
            if (we == cudaSuccess) we = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&wave_per_sm, kernel, wave_threads(c), wsmem);
        });
        if (we != cudaSuccess || wave_per_sm <= 0) { c->fused = false; cudaGetLastError(); }      // does not fit: global wavefront
        else {
            if (const char* env = getenv("B200PT_WAVE_CTAS")) wave_per_sm = std::max(1, std::min(wave_per_sm, atoi(env)));
            c->wave_blocks = c->num_sms * wave_per_sm;
        }
    }
#ifdef B200PT_PROBE
    if (const char* env = getenv("B200PT_PROBE")) {           // "pixel,iter" (diagnostic build only)
        int pr[2] = {-1, -1};
        sscanf(env, "%d,%d", &pr[0], &pr[1]);
        cudaMemcpyToSymbol(g_probe, pr, sizeof(pr));
    }
#endif
    if (cudaDeviceSynchronize() != cudaSuccess) return bail(fail(B200PT_ECUDA, std::string("scene upload failed: ") + cudaGetErrorString(cudaGetLastError())));
    *out_ctx = c;
    return B200PT_OK;
}

extern "C" int b200pt_get_info(b200pt_ctx* c, const char* name, int64_t* out_value) {
    if (!c || !name || !out_value) return fail(B200PT_EINVAL, "null argument");
    const std::string n(name);
    if (n == "lanes") *out_value = (int64_t)c->lanes.size();
    else if (n == "pool_per_lane") *out_value = c->lanes.empty() ? 0 : (int64_t)c->lanes[0].pool_cap;
    else if (n == "groups") *out_value = c->small_scene ? c->n_leaves : 0;
    else if (n == "small_kernel") *out_value = c->small_scene ? 1 : 0;
    else if (n == "lambert_only") *out_value = c->lambert_only ? 1 : 0;
    else if (n == "fused") *out_value = c->fused ? 1 : 0;
    else if (n == "bin_materials") *out_value = c->bin_materials ? 1 : 0;
    else if (n == "emit_boxes") *out_value = c->sc.n_emit_boxes;
    else if (n == "sort_rays") *out_value = c->sort_rays ? 1 : 0;
    else if (n == "wide") *out_value = c->wide ? 1 : 0;
    else if (n == "nodes4") *out_value = c->n_nodes4;
    else if (n == "wave_blocks") *out_value = c->wave_blocks;
    else return fail(B200PT_EINVAL, "unknown info " + n);
    return 0;
}

extern "C" int b200pt_set_option(b200pt_ctx* c, const char* name, int64_t value) {
    if (!c || !name) return fail(B200PT_EINVAL, "null argument");
    std::string n(name);
#ifndef B200PT_EMULATE
    if (c->graph_exec) { cudaSetDevice(c->device); cudaDeviceSynchronize(); cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }   // kernel arguments are baked in
#endif
    if (n == "graph") { c->use_graph = value != 0; return 0; }
    if (n == "pool") {
        if (value < 256 || value > (1 << 26)) return fail(B200PT_EINVAL, "pool must be in [256, 2^26]");
        CK(cudaSetDevice(c->device));
        CK(cudaDeviceSynchronize());
        c->pool_total = (int)value; c->pool_explicit = true;
        for (Lane& L : c->lanes) {
            int rc = alloc_pool(c, L, std::max(256, (int)value / (int)c->lanes.size()));
            if (rc) return rc;
        }
        return 0;
    }
    if (n == "reserve_iters") {
        // Pre-size every lane's sample planes for batches of up to `value` iterations, so that no later b200pt_render has
        // to re-allocate them inside the call (cudaFree + cudaMalloc there cost 1-50 ms each, occasionally far more).
        if (value < 1 || value > (1 << 20)) return fail(B200PT_EINVAL, "reserve_iters out of range");
        CK(cudaSetDevice(c->device));
        // b200pt_render never puts more iterations into one batch than fit max_batch_bytes (over all lanes)
        const size_t per_iter = (size_t)std::max(1, c->map.n_local_pixels) * sizeof(float4);
        const size_t iters = std::min<size_t>((size_t)value, std::max<size_t>(1, c->max_batch_bytes / per_iter));
        for (Lane& L : c->lanes) {
            const size_t need = iters * (size_t)L.map.n_local_pixels;
            int rc = ensure_samples(c, L, need);
            if (rc) return rc;
        }
        return 0;
    }
    if (n == "steps_per_poll") { if (value < 1 || value > 1024) return fail(B200PT_EINVAL, "steps_per_poll out of range"); c->steps_per_poll = (int)value; return 0; }
    if (n == "max_batch_bytes") { if (value < (1 << 20)) return fail(B200PT_EINVAL, "max_batch_bytes too small"); c->max_batch_bytes = (size_t)value; return 0; }
    if (n == "trace_ctas_per_sm") { if (value < 1 || value > 32) return fail(B200PT_EINVAL, "trace_ctas_per_sm out of range"); c->trace_blocks = c->num_sms * (int)value; return 0; }
    if (n == "refill_below") { if (value < 1 || value > 32) return fail(B200PT_EINVAL, "refill_below must be in [1, 32]"); c->refill_below = (int)value; return 0; }
    if (n == "small_kernel") { c->small_scene = value != 0 && c->leaves != nullptr; if (!c->small_scene) c->fused = false; return 0; }
    if (n == "bin_materials") { c->bin_materials = value != 0; return 0; }
    if (n == "fused") { c->fused = value != 0 && c->small_scene && c->wave_blocks > 0; return 0; }
    if (n == "wave_ctas_per_sm") { if (value < 1 || value > 8) return fail(B200PT_EINVAL, "wave_ctas_per_sm out of range"); c->wave_blocks = c->num_sms * (int)value; return 0; }
    if (n == "stage_smem") { if (!value) { c->stage_nodes = c->stage_prims = 0; c->small_scene = false; c->fused = false; } return 0; }
    return fail(B200PT_EINVAL, "unknown option " + n);
}

static void launch_shade(b200pt_ctx* c, const Lane& L, const ShadeArgs& sa) {
    const int blocks = L.pool.n / 128;
    const bool ldc = (c->mats_used & ~kMatsLDC) == 0u;
    if (c->het) {            // heterogeneous media: the coroutine stage (k_het.cuh)
        if (c->lambert_only) PT_LAUNCH((k_het_shade<kMatsLambertOnly>), blocks, 128, 0, L.stream, sa);
        else PT_LAUNCH((k_het_shade<kMatsAll>), blocks, 128, 0, L.stream, sa);
        return;
    }
    if (c->vol) {
        if (c->lambert_only) PT_LAUNCH((k_shade<true, kMatsLambertOnly>), blocks, 128, 0, L.stream, sa);
        else if (ldc) PT_LAUNCH((k_shade<true, kMatsLDC>), blocks, 128, 0, L.stream, sa);
        else PT_LAUNCH((k_shade<true, kMatsAll>), blocks, 128, 0, L.stream, sa);
    } else {
        if (c->lambert_only) PT_LAUNCH((k_shade<false, kMatsLambertOnly>), blocks, 128, 0, L.stream, sa);
        else if (ldc) PT_LAUNCH((k_shade<false, kMatsLDC>), blocks, 128, 0, L.stream, sa);
        else PT_LAUNCH((k_shade<false, kMatsAll>), blocks, 128, 0, L.stream, sa);
    }
}
static void launch_wave(b200pt_ctx* c, const Lane& L, const WaveArgs& wa) {
    const size_t smem = wave_smem(c);
    with_wave_kernel(c, [&](auto kernel, bool) { PT_LAUNCH_CTA(kernel, c->wave_blocks, wave_threads(c), smem, L.stream, wa); });
}
static void launch_trace(b200pt_ctx* c, const Lane& L, const TraceArgs& ta) {
    if (c->small_scene) {
        const size_t smem = (size_t)c->small_prim_bytes + (size_t)c->n_leaves * 32;
        if (c->vol) PT_LAUNCH(k_trace_small<true>, c->trace_blocks, kTraceThreads, smem, L.stream, ta);
        else PT_LAUNCH(k_trace_small<false>, c->trace_blocks, kTraceThreads, smem, L.stream, ta);
        return;
    }
    const size_t smem = (size_t)c->stage_nodes + c->stage_prims + kTraceStackBytes;
    if (c->sort_rays && L.sort_out) {
        RaySortArgs sa;
        sa.sc = c->sc; sa.pool = L.pool; sa.q = L.q; sa.parity = ta.parity; sa.sec_tmax = ta.sec_tmax;
        sa.keys = L.sort_keys; sa.hist = L.sort_hist; sa.offs = L.sort_offs; sa.sorted = L.sort_out;
        PT_LAUNCH(k_ray_hist, c->num_sms * 8, 256, 0, L.stream, sa);
        PT_LAUNCH(k_ray_scan, 1, 1024, 0, L.stream, sa);
        PT_LAUNCH(k_ray_scatter, c->num_sms * 8, 256, 0, L.stream, sa);
        TraceArgs t2 = ta;
        t2.q.entries = L.sort_out;
        with_trace_kernel(c, [&](auto kernel) { PT_LAUNCH(kernel, c->trace_blocks, kTraceThreads, smem, L.stream, t2); });
        return;
    }
    with_trace_kernel(c, [&](auto kernel) { PT_LAUNCH(kernel, c->trace_blocks, kTraceThreads, smem, L.stream, ta); });
}
static void fill_args(b200pt_ctx* c, Lane& L, const Camera& cam, const BatchParams& bp) {
    ShadeArgs& sa = L.sa; TraceArgs& ta = L.ta;
    sa.sc = c->sc; sa.pool = L.pool; sa.counters = L.counters; sa.samples = L.samples; sa.q = L.q; sa.parity = 0; sa.cam = cam; sa.map = L.map; sa.batch = bp; sa.frame = nullptr; sa.drain_hint = 0; sa.bin_materials = c->bin_materials ? 1 : 0;
    ta.sc = c->sc; ta.pool = L.pool; ta.q = L.q; ta.counters = L.counters; ta.parity = 0; ta.refill_below = c->refill_below;
    ta.stage_bytes_nodes = c->stage_nodes; ta.stage_bytes_prims = c->stage_prims; ta.small_prim_bytes = c->small_prim_bytes; ta.leaves = c->leaves; ta.n_leaves = c->n_leaves; ta.sec_tmax = c->het ? 1 : 0;
    ta.mis_anyhit = (!c->vol && !c->het && c->sc.n_lights == 0 && c->sc.inf.isvalid && !getenv("B200PT_NO_MIS_ANYHIT")) ? 1 : 0;
}

// How many of a lane's slots a batch of `need` samples uses.  Every wavefront step has a fixed cost (two launches, and a
// kernel cannot end before its longest ray has), so more slots = fewer, fuller steps — until the batch has too few samples
// per slot to keep them regenerating and the ramp / drain steps take over.  Measured (profiles/r03f_pool_sweep.txt,
// r03f_pool_adaptive.txt): C3 at 64 spp (28 M samples) with 0.26 / 0.5 / 1 / 2 / 4 / 8 / 16 M slots -> 402 / 448 / 469 / 486 /
// 501 / 493 / 476 Msamples/s; hair at 32 spp best at 4 samples per slot, C4 at 8 spp flat from 4 to 8 — so a batch uses
// need / 5 slots, at least 2^19 per lane.  All samples in flight at once (one pass, max_depth + 1 steps, no regeneration) is
// kept for the one-Render-per-frame usage, where it is the only choice: at 8 spp it loses 6 % against two passes.
static int batch_pool_slots(const b200pt_ctx* c, const Lane& L, size_t need) {
    if (c->pool_explicit) return L.pool_cap;
    size_t n;
    if (need <= (size_t)L.map.n_local_pixels + (size_t)L.map.n_local_pixels / 8) n = need;          // one iteration: exact-step frame
    else n = std::max<size_t>(need / 5, (size_t)1 << 19);
    n = (std::min<size_t>(n, need) + 255) & ~(size_t)255;
    return (int)std::min<size_t>(n, (size_t)L.pool_cap);
}

// One batch = n_iters iterations of every local pixel through the wavefronts of all lanes, then the ordered resolve.
// Lane streams are forked from / joined into lane 0's stream with events, so the caller sees one ordered stream.
// `capture`: the call is being recorded into a CUDA graph (exact-step mode only: no allocation, no host polling, and
// the per-call inputs are read from c->d_frame).
static int run_batch(b200pt_ctx* c, const Camera& cam, uint32_t first_iter, uint32_t n_iters, int reset, float* out_dev, int write_out,
                     double* launches, double* steps, bool capture = false) {
    CK(cudaEventRecord(c->ev_fork, c->stream));
    for (size_t k = 0; k < c->lanes.size(); ++k) {
        Lane& L = c->lanes[k];
        L.done = L.map.n_local_pixels == 0; L.chunk = 0; L.exact_pending = false; L.nearly_done = false;
        if (L.done) continue;
        if (k) CK(cudaStreamWaitEvent(L.stream, c->ev_fork, 0));
        const size_t need = (size_t)n_iters * L.map.n_local_pixels;
        if (need > L.samples_cap) {
            if (capture) return fail(B200PT_ECUDA, "internal: sample planes must be allocated before capture");
            int rc = ensure_samples(c, L, need);
            if (rc) return rc;
        }
        BatchParams bp; bp.first_iter = first_iter; bp.n_iters = n_iters; bp.total = need;
        if (c->fused) {
            // CTA-local wavefront: the whole batch of this lane is ONE persistent launch (slots, queue and scene in shared
            // memory); ~3/4 of the samples are handed out statically per slot, the rest from the global counter
            const unsigned long long P = (unsigned long long)c->wave_blocks * wave_threads(c);
            bp.k_static = (uint32_t)((need - need / 4) / P);
            L.h_init[0] = (unsigned long long)bp.k_static * P; L.h_init[1] = 0ull;
            CK(cudaMemcpyAsync(L.counters, L.h_init, 2 * sizeof(unsigned long long), cudaMemcpyHostToDevice, L.stream));
            fill_args(c, L, cam, bp);
            L.sa.frame = capture ? c->d_frame : nullptr;
            WaveArgs wa; wa.sa = L.sa; wa.ta = L.ta;
            launch_wave(c, L, wa);
            *launches += 1; *steps += 1;
            if (capture) CK(cudaMemcpyAsync(&L.h_counters[0], L.counters, sizeof(Counters), cudaMemcpyDeviceToHost, L.stream));
            L.done = true;
            continue;
        }
        L.pool.n = batch_pool_slots(c, L, need);
        bp.k_static = (uint32_t)((need - need / 4) / (unsigned long long)L.pool.n);      // ~3/4 of the batch by static assignment
        // next_sample starts behind the statically assigned range; done_samples at 0 (rays keeps counting)
        L.h_init[0] = (unsigned long long)bp.k_static * (unsigned long long)L.pool.n; L.h_init[1] = 0ull;
        CK(cudaMemcpyAsync(L.counters, L.h_init, 2 * sizeof(unsigned long long), cudaMemcpyHostToDevice, L.stream));
        CK(cudaMemsetAsync(L.pool.li_t, 0, sizeof(float4) * (size_t)L.pool.n, L.stream));   // no static samples consumed yet
        CK(cudaMemsetAsync(L.q.ctl, 0, sizeof(QueueCtl), L.stream));
        // all slots start dead: the first shade pass only regenerates
        CK(cudaMemsetAsync(L.pool.d_flags, 0, sizeof(float4) * (size_t)L.pool.n, L.stream));
        fill_args(c, L, cam, bp);
        L.sa.frame = capture ? c->d_frame : nullptr;
        L.sa.parity = 0; launch_shade(c, L, L.sa); *launches += 1;
        // single-pass batch: that first pass handed out every sample, so from now on dead tiles may leave early
        if (bp.total <= (unsigned long long)L.pool.n) L.sa.drain_hint = 1;
    }
    // Launch in chunks, lanes interleaved, and poll each lane's retired-sample counter one chunk behind, so the GPU
    // never waits on the host.  Step i of a lane: trace consumes queue set (i & 1), the shade after it fills the other.
    for (;;) {
        bool all_done = true;
        for (Lane& L : c->lanes) {
            if (L.done) continue;
            all_done = false;
            // `pt` with every sample of the batch in flight at once (the reference's one-Render-per-frame usage): a path
            // needs at most max_depth bounces, so exactly max_depth + 1 steps retire everything — launch those and
            // check once, instead of polling in chunks of steps_per_poll
            // (`vpt` has no such bound — medium boundaries do not count as bounces — so a single-pass `vpt` batch is
            // checked synchronously after 8 steps and then every 4)
            const bool single_pass = L.sa.batch.total <= (unsigned long long)L.pool.n;
            const bool exact = single_pass && (!c->vol ? L.chunk == 0 : true);
            // (a lane that was almost finished at the last poll is launched in short chunks: fewer empty tail steps)
            const int n_steps = !exact ? (L.nearly_done ? 2 : c->steps_per_poll) : (!c->vol ? c->sc.max_depth + 1 : (L.chunk == 0 ? 8 : 4));
            for (int s = 0; s < n_steps; ++s) {
                launch_trace(c, L, L.ta); L.ta.parity ^= 1u;
                L.sa.parity ^= 1u; launch_shade(c, L, L.sa);
            }
            *launches += (c->sort_rays && L.sort_out && !c->small_scene ? 5.0 : 2.0) * n_steps; *steps += n_steps;
            const int slot = L.chunk & 1;
            CK(cudaMemcpyAsync(&L.h_counters[slot], L.counters, sizeof(Counters), cudaMemcpyDeviceToHost, L.stream));
            if (capture) { L.done = true; ++L.chunk; continue; }      // verdict is read by the caller after the graph has run
            CK(cudaEventRecord(L.ev_poll[slot], L.stream));
            if (exact) L.exact_pending = true;
            else if (L.chunk >= 1) {
                const int prev = (L.chunk - 1) & 1;
                CK(cudaEventSynchronize(L.ev_poll[prev]));
                if (L.h_counters[prev].done_samples >= L.sa.batch.total) L.done = true;
                else if ((double)L.h_counters[prev].done_samples >= 0.97 * (double)L.sa.batch.total) L.nearly_done = true;
                if (L.h_counters[prev].next_sample >= L.sa.batch.total) L.sa.drain_hint = 1;
            }
            if (++L.chunk > (1 << 22)) return fail(B200PT_ECUDA, "wavefront did not converge (internal error)");
        }
        if (all_done) break;
        // exact-step lanes: all lanes have been launched; now wait for their verdicts (normally "done")
        for (Lane& L : c->lanes) {
            if (!L.exact_pending) continue;
            L.exact_pending = false;
            CK(cudaEventSynchronize(L.ev_poll[(L.chunk - 1) & 1]));
            if (L.h_counters[(L.chunk - 1) & 1].done_samples >= L.sa.batch.total) L.done = true;
        }
    }
    for (size_t k = 0; k < c->lanes.size(); ++k) {
        Lane& L = c->lanes[k];
        if (L.map.n_local_pixels == 0) continue;
        ResolveArgs ra; ra.samples = L.samples; ra.acc = c->acc; ra.color = c->color; ra.out = out_dev ? out_dev : c->out;
        ra.map = L.map; ra.batch = L.sa.batch; ra.reset = reset; ra.filmic = cam.filmic; ra.write_out = write_out;
        ra.frame = capture ? c->d_frame : nullptr;
        PT_LAUNCH(k_resolve, (L.map.n_local_pixels + 255) / 256, 256, 0, L.stream, ra);
        *launches += 1;
        if (k) { CK(cudaEventRecord(L.ev_done, L.stream)); CK(cudaStreamWaitEvent(c->stream, L.ev_done, 0)); }
    }
    CK(cudaGetLastError());
    return 0;
}

// Rays traced so far (statistics): one asynchronous read-back per lane on the context stream, one wait.
static int read_rays(b200pt_ctx* c, unsigned long long* total) {
    for (Lane& L : c->lanes) CK(cudaMemcpyAsync(&L.h_counters[1].rays, &L.counters->rays, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    *total = 0;
    for (Lane& L : c->lanes) *total += L.h_counters[1].rays;
    return 0;
}

extern "C" int b200pt_render(b200pt_ctx* c, const void* camera, uint32_t first_iter, uint32_t spp, int reset,
                             float* output, int output_is_device) {
    if (!c || !camera) return fail(B200PT_EINVAL, "null argument");
    if (spp == 0) return fail(B200PT_EINVAL, "spp must be >= 1");
    if (first_iter == 0) return fail(B200PT_EINVAL, "iter is 1-based (src/main.cpp:178 increments before the first Render)");
    CK(cudaSetDevice(c->device));
    Camera cam; std::memcpy(&cam, camera, sizeof(Camera));      // re-read every call, like src/pathtracer.cu:2706
    if (c->vol && (cam.medium >= c->sc.n_mediums || cam.medium < -1)) return fail(B200PT_EINVAL, "camera medium index out of range");
    c->last_filmic = cam.filmic;
    float* out_dev = (output && output_is_device) ? output : c->out;
    const size_t npix = (size_t)c->width * c->height;
    const unsigned long long rays0 = c->rays_seen;
    unsigned long long rays1 = 0;
    int rc = 0;
    CK(cudaEventRecord(c->ev0, c->stream));
    if (reset && c->map.n_shards > 1) CK(cudaMemsetAsync(c->acc, 0, 3 * npix * sizeof(float), c->stream));
    if (c->map.n_shards > 1 && out_dev) CK(cudaMemsetAsync(out_dev, 0, 3 * npix * sizeof(float), c->stream));
    // batch size: as many iterations as fit the sample-plane budget
    size_t per_iter = (size_t)std::max(1, c->map.n_local_pixels) * sizeof(float4);
    uint32_t max_iters = (uint32_t)std::max<size_t>(1, c->max_batch_bytes / per_iter);
    double launches = 0, steps = 0;
    uint32_t done = 0;
#ifndef B200PT_EMULATE
    // one Render per frame (`pt`, 1 spp, every lane single-pass): the whole frame — uploads, shade / trace steps of all
    // lanes, resolves — is ONE captured CUDA graph, replayed with the per-call inputs refreshed through c->d_frame
    bool graph_frame = c->use_graph && !c->vol && spp == 1;
    if (!c->fused) for (Lane& L : c->lanes) if (L.map.n_local_pixels > L.pool_cap) graph_frame = false;
    if (graph_frame) {
        for (Lane& L : c->lanes) {
            int rc2 = ensure_samples(c, L, (size_t)L.map.n_local_pixels);
            if (rc2) return rc2;
        }
        if (!c->graph_exec) {
            cudaGraph_t graph = nullptr;
            CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
            cudaMemcpyAsync(c->d_frame, c->h_frame, sizeof(FrameParams), cudaMemcpyHostToDevice, c->stream);
            c->graph_launches = 0; c->graph_steps = 0;
            int rc2 = run_batch(c, cam, first_iter, 1, reset, out_dev, 1, &c->graph_launches, &c->graph_steps, true);
            cudaError_t e2 = cudaStreamEndCapture(c->stream, &graph);
            if (rc2 || e2 != cudaSuccess || !graph) { if (graph) cudaGraphDestroy(graph); return rc2 ? rc2 : fail(B200PT_ECUDA, std::string("graph capture failed: ") + cudaGetErrorString(e2)); }
            e2 = cudaGraphInstantiate(&c->graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e2 != cudaSuccess) { c->graph_exec = nullptr; return fail(B200PT_ECUDA, std::string("graph instantiate failed: ") + cudaGetErrorString(e2)); }
        }
        c->h_frame->cam = cam; c->h_frame->first_iter = first_iter; c->h_frame->reset = reset ? 1 : 0;
        c->h_frame->filmic = cam.filmic; c->h_frame->out = out_dev;
        CK(cudaGraphLaunch(c->graph_exec, c->stream));
        launches = c->graph_launches; steps = c->graph_steps;
        done = spp;
    }
#endif
    while (done < spp) {
        uint32_t n = std::min(max_iters, spp - done);
        bool last = done + n == spp;
        rc = run_batch(c, cam, first_iter + done, n, reset && done == 0, out_dev, last ? 1 : 0, &launches, &steps);
        if (rc) return rc;
        done += n;
    }
    CK(cudaEventRecord(c->ev1, c->stream));
    if (output && !output_is_device) CK(cudaMemcpyAsync(output, out_dev, 3 * npix * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
#ifndef B200PT_EMULATE
    if (graph_frame)
        for (Lane& L : c->lanes)
            if (L.map.n_local_pixels && L.h_counters[0].done_samples < (unsigned long long)L.map.n_local_pixels)
                return fail(B200PT_ECUDA, "captured frame did not retire every sample (internal error)");
#endif
#ifndef B200PT_EMULATE
    if (graph_frame) { for (Lane& L : c->lanes) rays1 += L.map.n_local_pixels ? L.h_counters[0].rays : L.h_counters[1].rays; }   // read back by the graph itself
    else
#endif
    if ((rc = read_rays(c, &rays1))) return rc;
    c->rays_seen = rays1;
    CK(cudaGetLastError());
    float ms = 0; cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->stats[0] = (double)spp * c->map.n_local_pixels; c->stats[1] = launches; c->stats[2] = (double)(rays1 - rays0);
    c->stats[3] = steps / (double)c->lanes.size(); c->stats[4] = ms;
    c->total_ms += ms;
    return B200PT_OK;
}

extern "C" int b200pt_get_accum(b200pt_ctx* c, float* dst, int dst_is_device) {
    if (!c || !dst) return fail(B200PT_EINVAL, "null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(dst, c->acc, 3 * (size_t)c->width * c->height * sizeof(float),
                       dst_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}
extern "C" int b200pt_accum_device_ptr(b200pt_ctx* c, float** out_ptr) {
    if (!c || !out_ptr) return fail(B200PT_EINVAL, "null argument");
    *out_ptr = c->acc;
    return 0;
}
extern "C" int b200pt_get_color(b200pt_ctx* c, float* dst_host) {
    if (!c || !dst_host) return fail(B200PT_EINVAL, "null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(dst_host, c->color, 3 * (size_t)c->width * c->height * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}
extern "C" int b200pt_tonemap(b200pt_ctx* c, const float* acc_device, uint32_t iter, float* out_device) {
    if (!c || !acc_device || !out_device || iter == 0) return fail(B200PT_EINVAL, "bad argument");
    CK(cudaSetDevice(c->device));
    uint32_t npix = c->width * c->height;
    PT_LAUNCH(k_tonemap, (npix + 255) / 256, 256, 0, c->stream, acc_device, out_device, npix, iter, c->last_filmic);
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// Primary-visibility query through the production traversal kernel: regenerate one iteration, trace once.
// Uses lane 0's pool with a whole-image map.
extern "C" int b200pt_trace_primary(b200pt_ctx* c, const void* camera, uint32_t iter, float* hits_host) {
    if (!c || !camera || !hits_host || iter == 0) return fail(B200PT_EINVAL, "bad argument");
    CK(cudaSetDevice(c->device));
    if (c->map.n_shards != 1) return fail(B200PT_EINVAL, "trace_primary needs an unsharded context");
    Camera cam; std::memcpy(&cam, camera, sizeof(Camera));
    Lane& L = c->lanes[0];
    const ShardMap saved_map = L.map;
    L.map = c->map;
    const uint32_t npix = c->width * c->height;
    std::vector<float4> hits(npix);
    L.pool.n = L.pool_cap;
    const uint32_t P = (uint32_t)L.pool.n;
    int rc = 0;
    if ((rc = ensure_samples(c, L, npix))) { L.map = saved_map; return rc; }
    std::vector<float4> tmp(P), bs(P), fl(P);
    for (uint32_t base = 0; base < npix && !rc; base += P) {
        // hand out exactly the samples [base, base + P) of this iteration
        uint32_t cnt = std::min(P, npix - base);
        BatchParams bp; bp.first_iter = iter; bp.n_iters = 1; bp.total = base + cnt; bp.k_static = 0;
        Counters z; std::memset(&z, 0, sizeof(z)); z.next_sample = base;
        cudaMemcpyAsync(L.counters, &z, sizeof(unsigned long long) * 2, cudaMemcpyHostToDevice, L.stream);
        cudaMemsetAsync(L.pool.d_flags, 0, sizeof(float4) * (size_t)P, L.stream);
        cudaMemsetAsync(L.q.ctl, 0, sizeof(QueueCtl), L.stream);
        fill_args(c, L, cam, bp);
        launch_shade(c, L, L.sa);
        launch_trace(c, L, L.ta);
        cudaMemcpyAsync(tmp.data(), L.pool.hit0, sizeof(float4) * P, cudaMemcpyDeviceToHost, L.stream);
        cudaMemcpyAsync(bs.data(), L.pool.beta_s, sizeof(float4) * P, cudaMemcpyDeviceToHost, L.stream);
        cudaMemcpyAsync(fl.data(), L.pool.d_flags, sizeof(float4) * P, cudaMemcpyDeviceToHost, L.stream);
        if (cudaStreamSynchronize(L.stream) != cudaSuccess) rc = fail(B200PT_ECUDA, std::string("trace_primary: ") + cudaGetErrorString(cudaGetLastError()));
        for (uint32_t s = 0; s < P && !rc; ++s) {
            uint32_t sample, flags; std::memcpy(&sample, &bs[s].w, 4); std::memcpy(&flags, &fl[s].w, 4);
            if ((flags & F_ALIVE) && sample >= base && sample < base + cnt) hits[sample] = tmp[s];
        }
    }
    L.map = saved_map;
    if (rc) return rc;
    CK(cudaGetLastError());
    std::memcpy(hits_host, hits.data(), sizeof(float4) * npix);
    return 0;
}

extern "C" int b200pt_stats(b200pt_ctx* c, double* out5) {
    if (!c || !out5) return fail(B200PT_EINVAL, "null argument");
    std::memcpy(out5, c->stats, sizeof(c->stats));
    return 0;
}

#ifndef B200PT_EMULATE
static void comm_destroy(void* comm);
#endif
extern "C" int b200pt_destroy(b200pt_ctx* c) {
    if (!c) return fail(B200PT_EINVAL, "null context");
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
#ifndef B200PT_EMULATE
    if (c->comm) { comm_destroy(c->comm); c->comm = nullptr; }
#endif
    for (void* p : c->allocs) cudaFree(p);
    for (Lane& L : c->lanes) {
        free_pool(L);
        if (L.samples) cudaFree(L.samples);
        if (L.h_counters) cudaFreeHost(L.h_counters);
        if (L.h_init) cudaFreeHost(L.h_init);
        for (auto& e : L.ev_poll) if (e) cudaEventDestroy(e);
        if (L.ev_done) cudaEventDestroy(L.ev_done);
        if (L.stream) cudaStreamDestroy(L.stream);
    }
#ifndef B200PT_EMULATE
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
#endif
    if (c->h_frame) cudaFreeHost(c->h_frame);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    delete c;
    return 0;
}

// ---- multi-GPU: NCCL reduce of the accumulation framebuffers inside the library -------------------------------
#ifndef B200PT_EMULATE
namespace {
// NCCL is bound at run time: the library has no link-time dependency on it, and a process that already carries an NCCL
// (torch's bundled one) shares it.
struct Id128 { char b[128]; };                         // ncclUniqueId (passed BY VALUE to ncclCommInitRank)
struct Nccl {
    void* h = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, struct Id128, int) = nullptr;
    int (*CommInitAll)(void**, int, const int*) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
Nccl g_nccl;
int nccl_load() {
    if (g_nccl.h) return 0;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(B200PT_EUNSUPPORTED, std::string("NCCL not available: ") + dlerror());
    auto sym = [&](const char* n) { return dlsym(h, n); };
    g_nccl.GetUniqueId = (int (*)(void*))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, Id128, int))sym("ncclCommInitRank");
    g_nccl.CommInitAll = (int (*)(void**, int, const int*))sym("ncclCommInitAll");
    g_nccl.Reduce = (int (*)(const void*, void*, size_t, int, int, int, void*, cudaStream_t))sym("ncclReduce");
    g_nccl.GroupStart = (int (*)())sym("ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())sym("ncclGroupEnd");
    g_nccl.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommInitAll || !g_nccl.Reduce || !g_nccl.GroupStart || !g_nccl.GroupEnd || !g_nccl.CommDestroy)
        return fail(B200PT_EUNSUPPORTED, "libnccl.so.2 lacks a required symbol");
    g_nccl.h = h;
    return 0;
}
int nccl_fail(int rc, const char* what) {
    return fail(B200PT_ECUDA, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error") + " (" + std::to_string(rc) + ")");
}
constexpr int kNcclFloat32 = 7, kNcclSum = 0;          // ncclDataType_t / ncclRedOp_t values (nccl.h)
}  // namespace
static void comm_destroy(void* comm) { if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm); }
#endif

static int ensure_reduced(b200pt_ctx* c) {
    if (c->reduced) return 0;
    return dev_alloc(c, &c->reduced, 3 * (size_t)c->width * c->height, true);
}
// root side of a reduce: tonemap the full-image sum into `output` (what Output does after the last iteration)
static int finish_reduced(b200pt_ctx* c, uint32_t iter, float* output, int output_is_device) {
    const size_t npix = (size_t)c->width * c->height;
    c->reduced_iter = iter;
    if (!output) { CK(cudaStreamSynchronize(c->stream)); return 0; }
    float* out_dev = output_is_device ? output : c->out;
    PT_LAUNCH(k_tonemap, ((uint32_t)npix + 255) / 256, 256, 0, c->stream, c->reduced, out_dev, (uint32_t)npix, iter, c->last_filmic);
    if (!output_is_device) CK(cudaMemcpyAsync(output, out_dev, 3 * npix * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int b200pt_comm_unique_id(void* id128) {
#ifdef B200PT_EMULATE
    (void)id128; return fail(B200PT_EUNSUPPORTED, "NCCL communicators need a GPU build");
#else
    if (!id128) return fail(B200PT_EINVAL, "null argument");
    int rc = nccl_load();
    if (rc) return rc;
    int nr = g_nccl.GetUniqueId(id128);
    return nr ? nccl_fail(nr, "ncclGetUniqueId") : 0;
#endif
}
extern "C" int b200pt_comm_init(b200pt_ctx* c, int n_ranks, int rank, const void* id128) {
#ifdef B200PT_EMULATE
    (void)c; (void)n_ranks; (void)rank; (void)id128; return fail(B200PT_EUNSUPPORTED, "NCCL communicators need a GPU build");
#else
    if (!c || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(B200PT_EINVAL, "bad argument");
    if (c->comm) return fail(B200PT_EINVAL, "context already has a communicator");
    if (c->map.n_shards != n_ranks || c->map.shard != rank) return fail(B200PT_EINVAL, "the context's shard must be {rank, n_ranks}");
    int rc = nccl_load();
    if (rc) return rc;
    CK(cudaSetDevice(c->device));
    Id128 id; std::memcpy(&id, id128, sizeof(id));
    int nr = g_nccl.CommInitRank(&c->comm, n_ranks, id, rank);
    if (nr) { c->comm = nullptr; return nccl_fail(nr, "ncclCommInitRank"); }
    c->comm_rank = rank; c->comm_size = n_ranks;
    return 0;
#endif
}
extern "C" int b200pt_render_reduce(b200pt_ctx* c, const void* camera, uint32_t first_iter, uint32_t spp, int reset, int root,
                                    float* output, int output_is_device) {
#ifdef B200PT_EMULATE
    (void)c; (void)camera; (void)first_iter; (void)spp; (void)reset; (void)root; (void)output; (void)output_is_device;
    return fail(B200PT_EUNSUPPORTED, "NCCL communicators need a GPU build");
#else
    if (!c || !c->comm) return fail(B200PT_EINVAL, "b200pt_comm_init first");
    if (root < 0 || root >= c->comm_size) return fail(B200PT_EINVAL, "bad root");
    int rc = b200pt_render(c, camera, first_iter, spp, reset, nullptr, 0);
    if (rc) return rc;
    const bool is_root = c->comm_rank == root;
    if (is_root && (rc = ensure_reduced(c))) return rc;
    // the ONE collective of the data path: float3 framebuffer, sum, onto root, on the context's stream (NVLink / NVSwitch)
    int nr = g_nccl.Reduce(c->acc, is_root ? c->reduced : nullptr, 3 * (size_t)c->width * c->height, kNcclFloat32, kNcclSum, root, c->comm, c->stream);
    if (nr) return nccl_fail(nr, "ncclReduce");
    if (is_root) return finish_reduced(c, first_iter + spp - 1, output, output_is_device);
    CK(cudaStreamSynchronize(c->stream));
    return 0;
#endif
}
extern "C" int b200pt_reduced_accum(b200pt_ctx* c, float* dst, int dst_is_device) {
    if (!c || !dst) return fail(B200PT_EINVAL, "null argument");
    if (!c->reduced) return fail(B200PT_EINVAL, "no reduced image on this context (not the root of a reduce)");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(dst, c->reduced, 3 * (size_t)c->width * c->height * sizeof(float),
                       dst_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// single process, several GPUs
struct b200pt_multi {
    std::vector<b200pt_ctx*> ctx;
    std::vector<void*> comms;
    double stats[5] = {0, 0, 0, 0, 0};
};
extern "C" int b200pt_multi_destroy(b200pt_multi* m) {
    if (!m) return fail(B200PT_EINVAL, "null argument");
#ifndef B200PT_EMULATE
    for (void* cm : m->comms) if (cm && g_nccl.CommDestroy) g_nccl.CommDestroy(cm);
#endif
    for (b200pt_ctx* c : m->ctx) if (c) { c->comm = nullptr; b200pt_destroy(c); }
    delete m;
    return 0;
}
extern "C" int b200pt_create_multi(const b200pt_scene_view* scene, uint32_t width, uint32_t height, float epsilon,
                                   int n_gpus, const int* devices, b200pt_multi** out_multi) {
    if (!scene || !out_multi || n_gpus < 1 || n_gpus > 64) return fail(B200PT_EINVAL, "bad argument");
    b200pt_multi* m = new b200pt_multi();
    m->ctx.assign(n_gpus, nullptr);
    std::vector<int> devs(n_gpus);
    for (int i = 0; i < n_gpus; ++i) devs[i] = devices ? devices[i] : i;
#ifdef B200PT_EMULATE
    for (int i = 0; i < n_gpus; ++i) devs[i] = 0;              // the emulation has one "device"
#endif
    for (int i = 0; i < n_gpus; ++i) {
        b200pt_shard sh{i, n_gpus, 32, 32};
        int rc = b200pt_create(scene, width, height, epsilon, devs[i], n_gpus > 1 ? &sh : nullptr, &m->ctx[i]);
        if (rc) { std::string keep = g_err; b200pt_multi_destroy(m); g_err = keep; return rc; }
    }
#ifndef B200PT_EMULATE
    if (n_gpus > 1) {
        int rc = nccl_load();
        if (rc) { std::string keep = g_err; b200pt_multi_destroy(m); g_err = keep; return rc; }
        m->comms.assign(n_gpus, nullptr);
        int nr = g_nccl.CommInitAll(m->comms.data(), n_gpus, devs.data());
        if (nr) { m->comms.clear(); nccl_fail(nr, "ncclCommInitAll"); std::string keep = g_err; b200pt_multi_destroy(m); g_err = keep; return B200PT_ECUDA; }
        for (int i = 0; i < n_gpus; ++i) { m->ctx[i]->comm = m->comms[i]; m->ctx[i]->comm_rank = i; m->ctx[i]->comm_size = n_gpus; }
    }
#endif
    *out_multi = m;
    return 0;
}
extern "C" int b200pt_multi_render(b200pt_multi* m, const void* camera, uint32_t first_iter, uint32_t spp, int reset,
                                   float* output, int output_is_device) {
    if (!m || !camera) return fail(B200PT_EINVAL, "null argument");
    const int n = (int)m->ctx.size();
    if (n == 1) {
        int rc = b200pt_render(m->ctx[0], camera, first_iter, spp, reset, output, output_is_device);
        if (!rc) b200pt_stats(m->ctx[0], m->stats);
        return rc;
    }
    // all shards render concurrently: one host thread per GPU drives that context's (blocking) render call
    std::vector<int> rcs(n, 0);
    std::vector<std::string> errs(n);
#ifndef B200PT_EMULATE
    {
        std::vector<std::thread> th;
        for (int i = 0; i < n; ++i)
            th.emplace_back([&, i]() { rcs[i] = b200pt_render(m->ctx[i], camera, first_iter, spp, reset, nullptr, 0); if (rcs[i]) errs[i] = b200pt_last_error(); });
        for (auto& t : th) t.join();
    }
#else
    for (int i = 0; i < n; ++i) { rcs[i] = b200pt_render(m->ctx[i], camera, first_iter, spp, reset, nullptr, 0); if (rcs[i]) errs[i] = b200pt_last_error(); }
#endif
    for (int i = 0; i < n; ++i) if (rcs[i]) return fail(rcs[i], "GPU " + std::to_string(i) + ": " + errs[i]);
    b200pt_ctx* root = m->ctx[0];
    int rc = ensure_reduced(root);
    if (rc) return rc;
    const size_t count = 3 * (size_t)root->width * root->height;
#ifdef B200PT_EMULATE
    std::memset(root->reduced, 0, count * sizeof(float));
    for (int i = 0; i < n; ++i) for (size_t k = 0; k < count; ++k) root->reduced[k] += m->ctx[i]->acc[k];
#else
    int nr = g_nccl.GroupStart();
    for (int i = 0; i < n && !nr; ++i) {
        cudaSetDevice(m->ctx[i]->device);
        nr = g_nccl.Reduce(m->ctx[i]->acc, i == 0 ? root->reduced : nullptr, count, kNcclFloat32, kNcclSum, 0, m->ctx[i]->comm, m->ctx[i]->stream);
    }
    int ne = g_nccl.GroupEnd();
    if (nr || ne) return nccl_fail(nr ? nr : ne, "ncclReduce (group)");
    for (int i = 1; i < n; ++i) { cudaSetDevice(m->ctx[i]->device); CK(cudaStreamSynchronize(m->ctx[i]->stream)); }
    CK(cudaSetDevice(root->device));
#endif
    rc = finish_reduced(root, first_iter + spp - 1, output, output_is_device);
    if (rc) return rc;
    double agg[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
        double s5[5]; b200pt_stats(m->ctx[i], s5);
        agg[0] += s5[0]; agg[1] += s5[1]; agg[2] += s5[2]; agg[3] = std::max(agg[3], s5[3]); agg[4] = std::max(agg[4], s5[4]);
    }
    std::memcpy(m->stats, agg, sizeof(agg));
    return 0;
}
extern "C" int b200pt_multi_get_accum(b200pt_multi* m, float* dst_host) {
    if (!m || !dst_host) return fail(B200PT_EINVAL, "null argument");
    if (m->ctx.size() == 1) return b200pt_get_accum(m->ctx[0], dst_host, 0);
    return b200pt_reduced_accum(m->ctx[0], dst_host, 0);
}
extern "C" int b200pt_multi_stats(b200pt_multi* m, double* out5) {
    if (!m || !out5) return fail(B200PT_EINVAL, "null argument");
    std::memcpy(out5, m->stats, sizeof(m->stats));
    return 0;
}

extern "C" const char* b200pt_last_error(void) { return g_err.c_str(); }
extern "C" int b200pt_version(void) { return 101; }
