// k_trace.cuh — persistent-threads BVH traversal kernel (replaces Intersect / IntersectP / the ray walks of
// Tr, src/pathtracer.cu:214-322).
//
// * One thread = one ray; three ray classes per path slot (continuation closest-hit, shadow any-hit, MIS
//   closest-hit) are laid out class-major so a warp holds one class only.
// * Grid = a multiple of the SM count; every CTA first stages the acceleration structure (two-child 64-B
//   nodes + 48-B primitive records) into shared memory with one TMA bulk copy (cp.async.bulk + mbarrier) when
//   it fits, else traverses from L2/HBM with 16-B vector loads; then it grid-strides over the ray list.
// * Ordered traversal (near child first) with the reference's exact slab and Moeller-Trumbore arithmetic, so
//   hit/miss decisions are bit-identical to the reference's unordered DFS; exact-t ties resolve to the higher
//   primitive index, which is what the reference's visiting order produces (src/mesh.h:64 accepts tt == tmax).
#pragma once
#include "wavefront.cuh"

namespace pt {

#ifndef B200PT_EMULATE
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- TMA 1-D bulk copy global -> shared, completion on an mbarrier --------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
#endif  // !B200PT_EMULATE

// ---- the reference's slab test, BBox::Intersect (src/bbox.h:77-96), on one child box ---------------------
// inv = 1/d is hoisted out of the node loop (the reference recomputes the same value per node).
__device__ __forceinline__ bool slab(float bminx, float bminy, float bminz, float bmaxx, float bmaxy, float bmaxz,
                                     f3 o, f3 inv, float ray_tmax, float& tnear) {
    float t1 = (bminx - o.x) * inv.x;
    float t2 = (bmaxx - o.x) * inv.x;
    float t3 = (bminy - o.y) * inv.y;
    float t4 = (bmaxy - o.y) * inv.y;
    float t5 = (bminz - o.z) * inv.z;
    float t6 = (bmaxz - o.z) * inv.z;
    float tmin = fmaxf(fmaxf(fminf(t1, t2), fminf(t3, t4)), fminf(t5, t6));
    float tmax = fminf(fminf(fmaxf(t1, t2), fmaxf(t3, t4)), fmaxf(t5, t6));
    tnear = tmin;
    if (tmax <= 0.00001f) return false;
    if (tmin > tmax) return false;
    if (tmin > ray_tmax) return false;
    return true;
}

struct Hit { float t; int prim; float b1, b2; };

// Triangle::Intersect (src/mesh.h:45-66) / Sphere::Intersect (src/sphere.h:26-72) on a WPrim record.
// Returns nonzero and sets (t, b1, b2) when the primitive is accepted for the interval [tmin, tmax].
// Return value 2 = the sphere's far root was taken: the reference then overwrites ray.tmax and the hit record
// WITHOUT comparing against the current tmax (src/sphere.h:66-69) — kept, callers accept it unconditionally.
__device__ __forceinline__ int prim_test(const float4 q0, const float4 q1, const float4 q2, f3 o, f3 d,
                                          float tmin, float tmax, float& t_out, float& b1_out, float& b2_out) {
    if (__float_as_int(q2.y) == 0) {             // triangle
        f3 v0 = mk3(q0.x, q0.y, q0.z);
        f3 e1 = mk3(q0.w, q1.x, q1.y);
        f3 e2 = mk3(q1.z, q1.w, q2.x);
        f3 s1 = cross(d, e2);
        float divisor = dot(s1, e1);
        if (fabsf(divisor) < 1e-8f) return 0;
        float invDivisor = 1.0f / divisor;       // == (float)(1.0 / (double)divisor): IEEE division, 53 >= 2*24+2
        f3 s = o - v0;
        float b1 = dot(s, s1) * invDivisor;
        if (b1 < 0.0f || b1 > 1.0f) return 0;
        f3 s2 = cross(s, e1);
        float b2 = dot(d, s2) * invDivisor;
        if (b2 < 0.0f || b1 + b2 > 1.0f) return 0;
        float tt = dot(e2, s2) * invDivisor;
        if (tt < tmin || tt > tmax) return 0;
        t_out = tt; b1_out = b1; b2_out = b2;
        return 1;
    } else {                                     // sphere: q0 = centre.xyz, radius
        f3 op = o - mk3(q0.x, q0.y, q0.z);
        float radius = q0.w;
        float B = dot(op, d);
        float C = dot(op, op) - radius * radius;
        float delta = B * B - C;
        if (delta < 0.f) return 0;
        float sqrDelta = sqrtf(delta);
        float t1 = -B - sqrDelta;
        float t2 = -B + sqrDelta;
        if (t1 < 0.f && t2 < 0.f) return 0;
        if (t1 < 0.f || t2 < 0.f) {
            float tt1 = t1, tt2 = t2;
            t1 = tt1 < 0.f ? tt2 : tt1;
            t2 = tt1 < 0.f ? tt1 : tt2;
        } else if (t1 > t2) { float tmp = t2; t2 = t1; t1 = tmp; }
        if (t1 > tmax) return 0;
        int r = 1;
        float tt;
        if (t1 > tmin) tt = t1;
        else if (t2 > 0.f) { tt = t2; r = 2; }
        else return 0;
        t_out = tt; b1_out = 0.f; b2_out = 0.f;
        return r;
    }
}

// Closest hit (ANY == false) or any hit (ANY == true) of one ray against the staged structure.
template <bool ANY>
__device__ __forceinline__ bool traverse(const SceneDev& sc, const WNode* __restrict__ nodes, const WPrim* __restrict__ prims,
                                         f3 o, f3 d, float tmin, float tmax, Hit& hit) {
    hit.t = -1.f; hit.prim = -1; hit.b1 = 0.f; hit.b2 = 0.f;
    const f3 inv = mk3(1.f / d.x, 1.f / d.y, 1.f / d.z);
    float tn;
    // the reference tests the root's own box first (node 0, src/pathtracer.cu:222-223)
    if (!slab(sc.root_min[0], sc.root_min[1], sc.root_min[2], sc.root_max[0], sc.root_max[1], sc.root_max[2], o, inv, tmax, tn))
        return false;
    bool found = false;
    int stack[64];
    int sp = 0;
    // `cur` >= 0: inner node to visit; `leaf` >= 0: first primitive of a leaf to test (run ends at the record
    // flagged "last in leaf")
    int cur = sc.root_leaf_count > 0 ? -1 : 0;
    int leaf = sc.root_leaf_count > 0 ? 0 : -1;
    for (;;) {
        // ---- leaf: test its primitives (reference leaf loop, src/pathtracer.cu:230-245)
        if (leaf >= 0) {
            for (int pi = leaf;; ++pi) {
                const float4* pp = reinterpret_cast<const float4*>(prims + pi);
                const float4 q0 = pp[0], q1 = pp[1], q2 = pp[2];
                float t, b1, b2;
                const int acc = prim_test(q0, q1, q2, o, d, tmin, tmax, t, b1, b2);
                if (acc) {
                    if (ANY) return true;
                    // tt == tmax is accepted by the reference; the later (higher index) primitive then wins
                    if (acc == 2 || t < tmax || pi > hit.prim) { hit.t = t; hit.prim = pi; hit.b1 = b1; hit.b2 = b2; }
                    tmax = t; found = true;
                }
                if (__float_as_int(q2.z) != 0) break;    // last primitive of this leaf
            }
            leaf = -1;
        }
        if (cur < 0) {
            if (sp == 0) break;
            const int e = stack[--sp];
            if (e < 0) { leaf = ~e; continue; }
            cur = e;
        }
        // ---- inner node: two slab tests from one 64-B record
        const float4* np = reinterpret_cast<const float4*>(nodes + cur);
        const float4 q0 = np[0], q1 = np[1], q2 = np[2];
        const int4 link = reinterpret_cast<const int4*>(np)[3];
        float tl = 0.f, tr = 0.f;
        bool hl = slab(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, o, inv, tmax, tl);
        bool hr = link.y != kEmptyChild && slab(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, o, inv, tmax, tr);
        int c0 = link.x, c1 = link.y;
        if (hl && hr) {
            if (tr < tl) { int t_ = c0; c0 = c1; c1 = t_; }   // near child first, far child on the stack
            stack[sp++] = c1;
        } else if (hr) {
            c0 = c1;
        } else if (!hl) {
            cur = -1;
            continue;
        }
        if (c0 >= 0) cur = c0;
        else { cur = -1; leaf = ~c0; }
    }
    return found;
}

struct TraceArgs {
    SceneDev sc;
    Pool pool;
    Counters* counters;
    uint32_t stage_bytes_nodes, stage_bytes_prims;   // > 0: stage into shared memory with TMA
};

// Transmittance walk of Tr() (src/pathtracer.cu:298-322) for `vpt` shadow rays: closest hits until an opaque
// surface (matIdx != -1) blocks the ray, multiplying exp(-sigmaT * segment) of the current homogeneous medium
// and switching medium at every boundary crossed.
__device__ __forceinline__ f3 transmittance_walk(const SceneDev& sc, const WNode* nodes, const WPrim* prims,
                                                 f3 o, f3 d, float tmax_total, int medium, uint32_t& nrays) {
    f3 tr = mk3(1, 1, 1);
    float tmax = tmax_total;
    float seg_max = tmax_total;
    for (;;) {
        Hit h;
        ++nrays;
        bool invisible = traverse<false>(sc, nodes, prims, o, d, sc.eps, seg_max, h);
        float seg = invisible ? h.t : seg_max;
        if (invisible && sc.shade[h.prim].matIdx != -1) return mk3(0, 0, 0);
        if (medium >= 0) {
            f3 sigmaT = ld3(sc.mediums[medium].sigmaT);
            f3 c = sigmaT * (-seg);                                   // Homogeneous::Tr, src/medium.h:14
            tr *= mk3(expf(c.x), expf(c.y), expf(c.z));
        }
        if (!invisible) break;
        const WShade& s = sc.shade[h.prim];
        f3 nor;
        if (s.type == 0) nor = normalize(ld3(s.n1) * (1.f - h.b1 - h.b2) + ld3(s.n2) * h.b1 + ld3(s.n3) * h.b2);
        else nor = normalize((o + seg * d) - ld3(s.n1));
        medium = dot(d, nor) > 0 ? s.mediumOutside : s.mediumInside;
        tmax -= seg;
        o = o + seg * d;                                              // Ray(ray(ray.tmax), ray.d, m, eps, tmax)
        seg_max = tmax;
    }
    return tr;
}

template <bool VOL>
__global__ void __launch_bounds__(256) k_trace(const TraceArgs a) {
    const WNode* nodes = a.sc.nodes;
    const WPrim* prims = a.sc.prims;
#ifndef B200PT_EMULATE
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    if (a.stage_bytes_nodes + a.stage_bytes_prims > 0) {
        // stage the whole acceleration structure with one TMA bulk transaction per array
        WNode* s_nodes = reinterpret_cast<WNode*>(smem_raw);
        WPrim* s_prims = reinterpret_cast<WPrim*>(smem_raw + a.stage_bytes_nodes);
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, a.stage_bytes_nodes + a.stage_bytes_prims);
            if (a.stage_bytes_nodes) tma_bulk_g2s(s_nodes, a.sc.nodes, a.stage_bytes_nodes, &bar);
            tma_bulk_g2s(s_prims, a.sc.prims, a.stage_bytes_prims, &bar);
        }
        mbar_wait(&bar, 0);
        nodes = s_nodes; prims = s_prims;
    }
#endif
    const uint32_t P = (uint32_t)a.pool.n;
    const uint32_t total = 3u * P;
    uint32_t nrays = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t kind = i / P;              // 0 continuation, 1 shadow, 2 MIS  (class-major: warps are homogeneous)
        const uint32_t slot = i - kind * P;
        const float4 df = a.pool.d_flags[slot];
        const uint32_t flags = __float_as_uint(df.w);
        const uint32_t need = kind == 0 ? F_CONT : (kind == 1 ? F_SHADOW : F_MIS);
        if (!(flags & need)) continue;
        const float4 orng = a.pool.o_rng[slot];
        const f3 o = mk3(orng.x, orng.y, orng.z);
        if (kind == 0) {
            Hit h;
            traverse<false>(a.sc, nodes, prims, o, mk3(df.x, df.y, df.z), a.sc.eps, INFINITY, h);
            a.pool.hit0[slot] = make_float4(h.t, __int_as_float(h.prim), h.b1, h.b2);
            ++nrays;
        } else if (kind == 1) {
            const float4 sd = a.pool.shd[slot];
            f3 tr;
            if (!VOL) {
                Hit h;
                bool occluded = traverse<true>(a.sc, nodes, prims, o, mk3(sd.x, sd.y, sd.z), a.sc.eps, sd.w, h);
                tr = occluded ? mk3(0, 0, 0) : mk3(1, 1, 1);
                ++nrays;
            } else {
                int medium = (int)((flags >> kMedium2Shift) & 0xffu) - 1;
                tr = transmittance_walk(a.sc, nodes, prims, o, mk3(sd.x, sd.y, sd.z), sd.w, medium, nrays);
            }
            a.pool.vis[slot] = make_float4(tr.x, tr.y, tr.z, 0.f);
        } else {
            const float4 md = a.pool.misd[slot];
            Hit h;
            traverse<false>(a.sc, nodes, prims, o, mk3(md.x, md.y, md.z), a.sc.eps, INFINITY, h);
            a.pool.hit1[slot] = make_float4(h.t, __int_as_float(h.prim), h.b1, h.b2);
            ++nrays;
        }
    }
    // ray statistics: one atomic per warp
    for (int off = 16; off > 0; off >>= 1) nrays += __shfl_down_sync(0xffffffffu, nrays, off);
    if ((threadIdx.x & 31) == 0 && nrays) atomicAdd(&a.counters->rays, (unsigned long long)nrays);
}

}  // namespace pt
