// wavefront.cuh — device-side data layout of the wavefront integrator.
//
// The reference keeps one path per thread in registers for its whole life (src/pathtracer.cu:880-1021).
// Here a path lives in a slot of a structure-of-arrays POOL in HBM/L2 and moves through two kernels per
// bounce:  k_trace (all rays of all slots: continuation closest-hit, shadow any-hit, MIS closest-hit) and
// k_shade (finish the previous bounce's direct light, shade the new hit, emit the next three rays, or
// retire the sample and regenerate the slot from the global sample counter).
#pragma once
#include <stdint.h>
#include "cuda_compat.h"
#include "pt_math.cuh"

namespace pt {

// ---- re-laid-out scene (built once per context by k_prepare_* from the reference's arrays) ---------------
// Two-child BVH node, 64 B: both child boxes in one record so one visit = two slab tests + one 64-B fetch
// (the reference fetches a 40-B node per box test, src/pathtracer.cu:222).
//   q0 = lmin.xyz, lmax.x   q1 = lmax.yz, rmin.xy   q2 = rmin.z, rmax.xyz   q3 = {left, right, -, -}
// child >= 0: inner node index; child < 0: leaf whose first primitive is ~child (the run ends at the WPrim
// flagged last-in-leaf).
struct WNode { float4 q0, q1, q2; int4 link; };
static_assert(sizeof(WNode) == 64, "WNode");
constexpr int kEmptyChild = 0x7fffffff;

// Four-child node, 128 B: the boxes of up to four (grand)children of a reference node, structure-of-arrays so that each
// plane is one 16-byte load, plus their links (same encoding as WNode).  Built by collapsing every other level of the
// reference's binary tree (b200pt_api.cu: build_scene): a child box is inside its parent's box and BBox::Intersect is
// monotonic in the box, so a ray that passes a grandchild's test passes the skipped child's test as well — the SAME
// primitives reach the exact primitive test, with half the dependent node fetches per ray.
struct WNode4 { float4 minx, miny, minz, maxx, maxy, maxz; int4 link; int4 _pad; };
static_assert(sizeof(WNode4) == 128, "WNode4");

// Intersection record per primitive (leaf order), 48 B: v0, e1 = v1-v0, e2 = v2-v0 for Moeller-Trumbore
// (src/mesh.h:45-66 recomputes the edges per test from a 176-B Primitive), or centre+radius for a sphere, or the
// end points and radii of a hair segment (src/line.h:8).
//   triangle: q0 = v0.xyz, e1.x   q1 = e1.yz, e2.xy   q2 = e2.z, type(bits) = 0, last-in-leaf(bits), -
//   sphere:   q0 = centre.xyz, radius                 q2 = -, type = 1, last, -
//   line:     q0 = p0.xyz, p1.x   q1 = p1.yz, width0, width1   q2 = -, type = 2, last, -
struct WPrim { float4 q0, q1, q2; };
static_assert(sizeof(WPrim) == 48, "WPrim");

// Shading record per primitive, 96 B: fetched once per FINAL hit (the reference interpolates normal/uv/dpdu
// for every accepted candidate, src/mesh.h:68-95).
struct WShade {
    float n1[3], n2[3], n3[3];      // vertex normals            (sphere: n1 = centre, n2.x = radius)
    float uv1[2], uv2[2], uv3[2];
    float ndpdv[3];                 // normalize(dpdv) of the triangle (constant per triangle)
    int32_t matIdx, lightIdx, mediumInside, mediumOutside;
    int32_t type, _pad;
};
static_assert(sizeof(WShade) == 96, "WShade");

// Emitter record per Area light (src/area.h:7), 96 B.
struct WLight {
    float v1[3], v2[3], v3[3];
    float n1[3], n2[3], n3[3];
    float radiance[3];
    float area;                     // Triangle::GetSurfaceArea (src/mesh.h:39), computed once on the device
    float _pad[2];
};
static_assert(sizeof(WLight) == 96, "WLight");

struct WMedium { float sigmaA[3], sigmaS[3], sigmaT[3]; float g; int32_t type; int32_t _pad; };   // type 0 homogeneous (src/medium.h:9), 1 heterogeneous (+ WHetero)

// Heterogeneous medium (src/medium.h:52): density grid in the box p0..p1 (device pointer), indexed like WMedium.
struct WHetero {
    const float* density; int32_t nx, ny, nz;
    float invMaxDensity, p0[3], p1[3];
    int32_t iterMax, evalTransmittanceType;
};

struct WInfinite {                  // src/infinite.h:6 with a device texel pointer
    const float* data; int32_t width, height;
    float center[3], radius, u[3], v[3], w[3];
    int32_t isvalid;
};

struct SceneDev {
    const WNode* nodes; const WNode4* nodes4; const WPrim* prims; const WShade* shade; const WLight* lights;
    const Material* mats; const WMedium* mediums; const float* cdf;
    const unsigned char* texels;    // all uchar4 textures back to back (src/texture.h:9)
    const int4* tex_info;           // per texture: {first texel, width, height, -}
    WInfinite inf;
    float root_min[3], root_max[3];
    int32_t n_nodes, n_prims, n_lights, n_cdf, n_mats, n_mediums;
    int32_t root_leaf_count;        // > 0 when the whole scene is a single leaf (no inner node)
    int32_t integrator, max_depth;
    float eps;
    const float4* emit_boxes;       // boxes (group record layout: min.xyz, max.x | max.yz, -, -) a ray must pass to reach an emitter
    int32_t n_emit_boxes;           // 0: no culling of MIS rays (environment light, too many boxes, or switched off)
    const unsigned char* prim_key;  // per primitive: shade-stage sort key (material type, +8 for an emitter; 7 = no material)
    const WHetero* het;             // non-null when some medium is heterogeneous (then k_volpath_seq renders the scene)
};

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

// A BSDF-sampled light ray (the MIS ray of a bounce) adds radiance only if its CLOSEST hit is an emitter (or, with an
// environment light, if it escapes).  The traversal tests a primitive only after the ray passed the slab test of the box
// it sits in — its primitive group (small scenes) or its BVH leaf (tree kernel) — so a ray that fails that test for every
// box holding an emitter can never report an emitter: it is not traced at all.  Same test, same operands, same
// decision as the traversal would make; no random number is involved.  The host lists the boxes (SceneDev::emit_boxes,
// at most kMaxEmitBoxes; none when the scene has an environment light).  Cornell box: one box, 9 of 10 MIS rays dropped.
constexpr int kMaxEmitBoxes = 16;
__device__ __forceinline__ bool mis_ray_may_reach_emitter(const SceneDev& sc, const f3 o, const f3 d) {
    if (sc.n_emit_boxes <= 0) return true;
    const f3 inv = mk3(1.f / d.x, 1.f / d.y, 1.f / d.z);
    for (int i = 0; i < sc.n_emit_boxes; ++i) {
        const float4 q0 = sc.emit_boxes[2 * i], q1 = sc.emit_boxes[2 * i + 1];
        float tn;
        if (slab(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, o, inv, INFINITY, tn)) return true;
    }
    return false;
}

// ---- path pool -------------------------------------------------------------------------------------------
// flags word
constexpr uint32_t F_ALIVE = 1u << 0;        // slot holds a live sample
constexpr uint32_t F_CONT = 1u << 1;         // continuation ray to trace this step (closest hit -> hit0)
constexpr uint32_t F_SHADOW = 1u << 2;       // shadow ray to trace (any hit, or Tr() walk for vpt)
constexpr uint32_t F_MIS = 1u << 3;          // BSDF-sampled MIS ray to trace (closest hit -> hit1)
constexpr uint32_t F_PENDING = 1u << 4;      // previous bounce's direct light still to be added
constexpr uint32_t F_TERMINATE = 1u << 5;    // retire after the pending direct light is added
constexpr uint32_t F_SPECULAR = 1u << 6;     // last bounce was a delta BSDF
constexpr uint32_t F_MEDSCATTER = 1u << 7;   // pending contribution is a medium in-scatter (vpt)
constexpr uint32_t F_CARRY = 1u << 15;       // the pending direct light belongs to the slot's PREVIOUS sample (see Pool::carry)
constexpr int kBounceShift = 8;              // bits 8..14: bounce counter
constexpr int kMediumShift = 16;             // bits 16..23: continuation ray's medium index + 1 (0 = none)
constexpr int kMedium2Shift = 24;            // bits 24..31: medium of the shadow / MIS rays + 1

struct Pool {
    float4* o_rng;        // ray origin xyz (shared by all three rays of the slot), rng state bits
    float4* d_flags;      // continuation direction xyz, flags bits
    float4* beta_s;       // throughput xyz, sample index bits
    float4* li_t;         // radiance accumulated so far xyz, static samples this slot has consumed (bits)
    float4* shd;          // shadow ray direction xyz, tmax
    float4* misd;         // MIS ray direction xyz, pdf
    float4* ldl;          // direct-light term if unoccluded xyz, |cos| of the MIS direction
    float4* misf;         // BSDF value of the MIS sample xyz, -
    float4* beta_old;     // throughput in front of the pending direct light xyz, -
    float4* hit0;         // continuation hit: t (<0 miss), prim bits, b1, b2
    float4* hit1;         // MIS hit
    float4* vis;          // shadow result: transmittance xyz (0 = occluded; 1 for `pt`), -
    float4* aux;          // vpt only: emitter radiance xyz, MIS weight of the light sample
    // A path that ends with direct light still pending (Russian roulette, depth limit, black BSDF) does not idle for a
    // step: its slot is regenerated at once and CARRIES the old sample along — the shadow / MIS rays keep their own
    // origin, and the next shade pass adds their result to the carried radiance and retires the old sample.
    float4* pend_o;       // origin of the shadow / MIS rays xyz, carried sample index (bits)
    float4* carry;        // radiance of the carried sample so far xyz, -
    int32_t n;
};

// ---- TMA bulk copy + mbarrier primitives (sm_90+/sm_100a PTX) ---------------------------------------------------
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

// ---- ray queue ---------------------------------------------------------------------------------------------
// k_shade appends one 4-byte entry per ray it emits (slot | kind << 30; kind 0 = continuation closest-hit,
// 1 = shadow any-hit / transmittance walk, 2 = MIS closest-hit) with one warp-aggregated atomic per warp; the
// persistent k_trace warps pull entries with one atomic per refill, so every lane of a traversal warp carries a
// live ray.  Two control sets alternate by step parity: the trace kernel of step i consumes set i & 1 and zeroes
// the other set for the shade kernel that follows it.
constexpr int kKindShift = 30;
constexpr uint32_t kSlotMask = (1u << kKindShift) - 1u;
struct QueueCtl { uint32_t tail[2]; uint32_t head[2]; };
struct RayQueue { uint32_t* entries; QueueCtl* ctl; };

// ---- warp helpers (a "warp" is one lane wide in the CPU emulation build of the test suite) ------------------
#ifndef B200PT_EMULATE
__device__ __forceinline__ uint32_t pt_lane() { return threadIdx.x & 31u; }
#else
static inline uint32_t pt_lane() { return 0u; }
#endif
constexpr uint32_t kFullMask = 0xffffffffu;

struct Counters {
    unsigned long long next_sample;    // next (iteration, pixel) pair to hand out
    unsigned long long done_samples;   // retired samples
    unsigned long long rays;           // rays traced (statistics)
    unsigned long long _pad;
};

// Which pixels this context renders: interleaved screen tiles (multi-GPU sharding).
// Which tiles a shard owns is decided once on the host (shard_owner) and handed to the kernels as a table of tile
// origins, one 32-bit entry (tile x | tile y << 16) per local tile.
struct ShardMap {
    int32_t width, height;
    int32_t tile_w, tile_h, tiles_x, tiles_y;
    int32_t shard, n_shards;
    int32_t n_local_tiles, n_local_pixels;
    const uint32_t* tiles;          // [n_local_tiles] (device memory); unused when n_shards == 1
};
// Owner of tile k (row-major tile index).  Every group of n consecutive tiles holds one tile of each shard, and the
// assignment inside a group is rotated by a HASH of the group index: a shard's tiles are scattered over the image
// without any period.  (Plain k % n gave each of 8 ranks four fixed COLUMNS of a 32-tile-wide image — shard 3 of the
// Cornell box needed 9 % longer than shard 0, most of the 1 -> 8 GPU scaling loss of round 1; a fixed rotation of 3 per
// group still left a two-row period and 3 %: profiles/r02b_shard_tax.txt, r02l_shard_balance.txt.)
PT_HD int shard_owner(int k, int n) { return (int)(((uint32_t)(k % n) + wang_hash((uint32_t)(k / n)) % (uint32_t)n) % (uint32_t)n); }
// local pixel index -> global pixel (x, y). Local order: tile-major, row-major inside the tile.
__device__ __forceinline__ void local_to_xy(const ShardMap& m, uint32_t local, uint32_t& x, uint32_t& y) {
    if (m.n_shards == 1) { x = local % (uint32_t)m.width; y = local / (uint32_t)m.width; return; }
    uint32_t per_tile = (uint32_t)(m.tile_w * m.tile_h);
    uint32_t lt = local / per_tile, in = local - lt * per_tile;
    const uint32_t e = m.tiles[lt];
    uint32_t iy = in / (uint32_t)m.tile_w, ix = in - iy * (uint32_t)m.tile_w;
    x = (e & 0xffffu) * (uint32_t)m.tile_w + ix; y = (e >> 16) * (uint32_t)m.tile_h + iy;
}

// Per-call inputs of a captured (CUDA graph) frame: read through a pointer so that the graph's kernel arguments stay
// constant from frame to frame; a memcpy node refreshes it from pinned host memory at the head of the graph.
struct FrameParams { Camera cam; uint32_t first_iter; int32_t reset, filmic, _pad; float* out; };

struct BatchParams {
    uint32_t first_iter;     // iteration number of sample plane 0 (1-based, part of the RNG seed)
    uint32_t n_iters;        // iterations in this batch
    unsigned long long total;  // n_iters * n_local_pixels
    // Sample hand-out: slot s takes samples s, s + P, ... for its first k_static regenerations (no atomic, ~3/4 of
    // the batch); the rest comes from the global counter, which starts at k_static * P and evens out the tail.
    uint32_t k_static;
};

}  // namespace pt
