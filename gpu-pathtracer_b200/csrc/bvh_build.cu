// bvh_build.cu — binned-SAH BVH2 construction on the GPU: replaces the reference's host
// builder BVH::build / split / flatten (src/bvh.cpp:16-173), which recurses on the host and copies 176-B
// primitives at every level (seconds for 10^6 triangles).
//
// Same decisions as the reference, different schedule.  The reference splits one node at a time, depth first;
// here ALL nodes of one tree level are split together (level-synchronous, breadth first):
//   k_hist     one thread per primitive: for each axis, the primitive's bucket (12 per axis, src/bvh.cpp:66) in
//              its node's box; bucket boxes and counts accumulate with integer atomics on an order-preserving
//              encoding of the floats (min/max are exact, so the order of accumulation does not matter); a CTA
//              whose primitives all belong to one node accumulates in shared memory first.
//   k_split    one thread per node: the 3 x 11 candidate planes in the reference's order (axis major, strict <),
//              leaf criteria of src/bvh.cpp:43/:113, children allocated pairwise.
//   k_flags + exclusive scan + k_scatter: STABLE partition of every node's primitive range at once (the
//              reference's left/right push_back order, src/bvh.cpp:131-148), so the final permutation is the
//              reference's leaf order.
// Afterwards subtree sizes (bottom-up over the levels) give every node its pre-order index, which is the layout
// BVH::flatten emits (left child = index + 1, second_child_offset = right child); the permutation (4 B per
// primitive) goes back and the host reorders its 176-B primitives with it.  Output: field-for-field byte-identical LinearBVHNode[] / Primitive[] to b200pt_bvh_build
// (and therefore to the reference).  That includes the SIGN OF ZERO of a box coordinate, which in the host code
// depends on the visiting order (its min/max keep the first of two equal operands when boxes are merged): every
// stable partition keeps a node's primitives in input order, so "first" is "smallest input index", and
// k_zero_ids / k_zero_sign restore exactly that after the atomics (which fold -0 onto +0).  Every float
// expression that decides a bucket or a cost is written with explicitly rounded operations: the reference is
// host code without FMA contraction.
#include "cuda_compat.h"
#include "b200pt.h"
#include "ref_layouts.h"

#include <chrono>
#include <cstring>
#include <string>
#include <thread>
#include <vector>


int b200pt_internal_fail(int code, const char* msg);   // b200pt_api.cu: sets b200pt_last_error()
extern "C" int b200pt_internal_prims_finite(const void* prims, int n, int* bad);   // host_prep.cpp

namespace bvhb {

constexpr int kBuckets = 12;                           // src/bvh.cpp:66
constexpr int kBucketWords = 7;                        // ~enc(lo.xyz), enc(hi.xyz), count
constexpr int kNodeWords = 3 * kBuckets * kBucketWords;
constexpr int kLeaf = -1, kCandidate = -2;             // values of Tree::left besides a child index
constexpr int kThreads = 256;

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ float rmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float radd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float rsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float rdiv(float a, float b) { return __fdiv_rn(a, b); }
#else
static inline float rmul(float a, float b) { return a * b; }     // emulation build: -ffp-contract=off
static inline float radd(float a, float b) { return a + b; }
static inline float rsub(float a, float b) { return a - b; }
static inline float rdiv(float a, float b) { return a / b; }
#endif

// order-preserving float -> unsigned (so that integer atomicMax is a float max); -0 is folded onto +0
__device__ __forceinline__ unsigned enc(float f) {
    f = (f == 0.f) ? 0.f : f;
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec(unsigned e) { return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e); }

struct Tree {
    float4* lo;        // xyz = box min, w = bits of first primitive position
    float4* hi;        // xyz = box max, w = bits of primitive count
    int* left;         // left child (right = left + 1), kLeaf, or kCandidate (waiting for this level's split)
    int* slot;         // candidate: index of its bucket block in this level; split node: axis | bucket << 2
    int* size;         // nodes in the subtree
    int* pre;          // pre-order index
    int* parent;
    int* zero_id;      // 6 per node: smallest input index among the primitives whose bound is the node's zero bound
};
struct Counters { int n_nodes, n_candidates; unsigned root[6]; };

struct Box3 { float lo[3], hi[3]; };
__device__ __forceinline__ void box_reset(Box3& b) {
    b.lo[0] = b.lo[1] = b.lo[2] = INFINITY; b.hi[0] = b.hi[1] = b.hi[2] = -INFINITY;
}
// BBox::Expand(point) with the host's min/max (src/cutil_math.h:36-44): on a tie the NEW point wins.  The select
// is done on the bit patterns: a float select would become FMNMX, which orders -0 below +0 instead.
__device__ __forceinline__ void box_grow_point(Box3& b, const float* p) {
    for (int a = 0; a < 3; ++a) {
        b.lo[a] = __uint_as_float(b.lo[a] < p[a] ? __float_as_uint(b.lo[a]) : __float_as_uint(p[a]));
        b.hi[a] = __uint_as_float(b.hi[a] > p[a] ? __float_as_uint(b.hi[a]) : __float_as_uint(p[a]));
    }
}
__device__ __forceinline__ void box_grow(Box3& b, const Box3& o) {
    for (int a = 0; a < 3; ++a) { b.lo[a] = fminf(b.lo[a], o.lo[a]); b.hi[a] = fmaxf(b.hi[a], o.hi[a]); }
}
// BBox::SurfaceArea (src/bbox.h:63-66)
__device__ __forceinline__ float box_area(const Box3& b) {
    float dx = rsub(b.hi[0], b.lo[0]), dy = rsub(b.hi[1], b.lo[1]), dz = rsub(b.hi[2], b.lo[2]);
    return rmul(2.f, radd(radd(rmul(dx, dy), rmul(dy, dz)), rmul(dz, dx)));
}
// bucket of a primitive on one axis (src/bvh.cpp:74-76, :133-135): centre = (lo + hi) * 0.5f (BBox::Centric)
__device__ __forceinline__ int bucket_of(float plo, float phi, float v0, float v1) {
    float c = rmul(radd(plo, phi), 0.5f);
    int no = (int)rmul(rdiv(rsub(c, v0), rsub(v1, v0)), (float)kBuckets);
    return no == kBuckets ? no - 1 : no;
}
// leaf criterion of src/bvh.cpp:43
__device__ __forceinline__ bool small_or_thin(int count, const float* lo, const float* hi) {
    return count <= 4 || rsub(hi[0], lo[0]) < 0.0001f || rsub(hi[1], lo[1]) < 0.0001f || rsub(hi[2], lo[2]) < 0.0001f;
}

// GetBBox(Primitive&) (src/bvh.cpp:3-10): triangle mesh.h:28, sphere sphere.h:17, line line.h:16
__global__ void k_prim_boxes(const RefPrimitive* prims, int n, float4* blo, float4* bhi, Counters* ctr) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    Box3 b; box_reset(b);
    if (i < n) {
        const RefPrimitive& p = prims[i];
        if (p.type == REF_GT_TRIANGLE) {
            box_grow_point(b, p.u.triangle.v1.v); box_grow_point(b, p.u.triangle.v2.v); box_grow_point(b, p.u.triangle.v3.v);
        } else if (p.type == REF_GT_SPHERE) {
            for (int a = 0; a < 3; ++a) {
                b.lo[a] = rsub(p.u.sphere.origin[a], p.u.sphere.radius);
                b.hi[a] = radd(p.u.sphere.origin[a], p.u.sphere.radius);
            }
        } else {
            const RefLine& l = p.u.line;
            float w = l.width0 > l.width1 ? l.width0 : l.width1;
            float q[3];
            for (int a = 0; a < 3; ++a) q[a] = rsub(l.p0[a], w); box_grow_point(b, q);
            for (int a = 0; a < 3; ++a) q[a] = radd(l.p0[a], w); box_grow_point(b, q);
            for (int a = 0; a < 3; ++a) q[a] = rsub(l.p1[a], w); box_grow_point(b, q);
            for (int a = 0; a < 3; ++a) q[a] = radd(l.p1[a], w); box_grow_point(b, q);
        }
        blo[i] = make_float4(b.lo[0], b.lo[1], b.lo[2], 0.f);
        bhi[i] = make_float4(b.hi[0], b.hi[1], b.hi[2], 0.f);
    }
    // scene box: warp reduction of the encoded bounds, one atomic per warp and word
    for (int a = 0; a < 3; ++a) {
        unsigned l = i < n ? ~enc(b.lo[a]) : 0u, h = i < n ? enc(b.hi[a]) : 0u;
#if defined(__CUDA_ARCH__)
        l = __reduce_max_sync(0xffffffffu, l); h = __reduce_max_sync(0xffffffffu, h);
        if ((threadIdx.x & 31) != 0) continue;
#endif
        if (l) atomicMax(&ctr->root[a], l);
        if (h) atomicMax(&ctr->root[3 + a], h);
    }
}

__global__ void k_init(int n, int* ids, int* seg, Tree t, Counters* ctr) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { ids[i] = i; seg[i] = 0; }
    if (i == 0) {
        float lo[3], hi[3];
        for (int a = 0; a < 3; ++a) { lo[a] = dec(~ctr->root[a]); hi[a] = dec(ctr->root[3 + a]); }
        t.lo[0] = make_float4(lo[0], lo[1], lo[2], __int_as_float(0));
        t.hi[0] = make_float4(hi[0], hi[1], hi[2], __int_as_float(n));
        bool leaf = small_or_thin(n, lo, hi);
        t.left[0] = leaf ? kLeaf : kCandidate;
        t.slot[0] = 0;
        t.parent[0] = -1;
        ctr->n_nodes = 1;
        ctr->n_candidates = leaf ? 0 : 1;
    }
}

__global__ void __launch_bounds__(kThreads) k_hist(int n, const int* __restrict__ ids, const int* __restrict__ seg, Tree t,
                                                   const float4* __restrict__ blo, const float4* __restrict__ bhi,
                                                   unsigned* buckets) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool priv = false;
#ifndef B200PT_EMULATE
    __shared__ unsigned sh[kNodeWords];
    const int first = blockIdx.x * blockDim.x, last = min(n - 1, first + (int)blockDim.x - 1);
    const int s_first = seg[first];
    priv = s_first == seg[last] && t.left[s_first] == kCandidate;    // ranges are contiguous: the whole CTA is in one node
    if (priv) {
        for (int w = threadIdx.x; w < kNodeWords; w += blockDim.x) sh[w] = 0u;
        __syncthreads();
    }
#endif
    if (i < n) {
        const int s = seg[i];
        if (t.left[s] == kCandidate) {
            const int id = ids[i];
            const float4 l4 = blo[id], h4 = bhi[id];
            const float4 nlo = t.lo[s], nhi = t.hi[s];
            const float pl[3] = {l4.x, l4.y, l4.z}, ph[3] = {h4.x, h4.y, h4.z};
            const float v0[3] = {nlo.x, nlo.y, nlo.z}, v1[3] = {nhi.x, nhi.y, nhi.z};
            unsigned e[6];
            for (int a = 0; a < 3; ++a) { e[a] = ~enc(pl[a]); e[3 + a] = enc(ph[a]); }
#ifndef B200PT_EMULATE
            unsigned* base = priv ? sh : buckets + (size_t)t.slot[s] * kNodeWords;
#else
            unsigned* base = buckets + (size_t)t.slot[s] * kNodeWords;
#endif
            for (int axis = 0; axis < 3; ++axis) {
                unsigned* b = base + (axis * kBuckets + bucket_of(pl[axis], ph[axis], v0[axis], v1[axis])) * kBucketWords;
                for (int k = 0; k < 6; ++k) atomicMax(&b[k], e[k]);
                atomicAdd(&b[6], 1u);
            }
        }
    }
#ifndef B200PT_EMULATE
    if (priv) {
        __syncthreads();
        unsigned* g = buckets + (size_t)t.slot[s_first] * kNodeWords;
        for (int w = threadIdx.x; w < kNodeWords; w += blockDim.x) {
            const unsigned v = sh[w];
            if (v) { if (w % kBucketWords == 6) atomicAdd(&g[w], v); else atomicMax(&g[w], v); }
        }
    }
#endif
}

// One thread per node of the level: choose the split (src/bvh.cpp:60-113) and create the children.
__global__ void k_split(int level_start, int level_end, Tree t, const unsigned* __restrict__ buckets, Counters* ctr,
                        int node_capacity) {
    const int s = level_start + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= level_end || t.left[s] != kCandidate) return;
    const float4 nlo = t.lo[s], nhi = t.hi[s];
    const int begin = __float_as_int(nlo.w), count = __float_as_int(nhi.w);
    Box3 bbox; bbox.lo[0] = nlo.x; bbox.lo[1] = nlo.y; bbox.lo[2] = nlo.z; bbox.hi[0] = nhi.x; bbox.hi[1] = nhi.y; bbox.hi[2] = nhi.z;
    const unsigned* B = buckets + (size_t)t.slot[s] * kNodeWords;

    float best_cost = rmul((float)count, box_area(bbox));
    int best_axis = -1, best_bucket = 0, best_left = 0;
    Box3 best_l, best_r; box_reset(best_l); box_reset(best_r);
    for (int axis = 0; axis < 3; ++axis) {
        Box3 bb[kBuckets]; int cnt[kBuckets];
        for (int k = 0; k < kBuckets; ++k) {
            const unsigned* w = B + (axis * kBuckets + k) * kBucketWords;
            cnt[k] = (int)w[6];
            box_reset(bb[k]);
            if (cnt[k]) for (int a = 0; a < 3; ++a) { bb[k].lo[a] = dec(~w[a]); bb[k].hi[a] = dec(w[3 + a]); }
        }
        Box3 suffix[kBuckets]; int csuf[kBuckets];        // union / count of buckets k >= j
        suffix[kBuckets - 1] = bb[kBuckets - 1]; csuf[kBuckets - 1] = cnt[kBuckets - 1];
        for (int k = kBuckets - 2; k >= 0; --k) { suffix[k] = suffix[k + 1]; box_grow(suffix[k], bb[k]); csuf[k] = csuf[k + 1] + cnt[k]; }
        Box3 b0; box_reset(b0); int c0 = 0;
        for (int j = 1; j < kBuckets; ++j) {
            box_grow(b0, bb[j - 1]); c0 += cnt[j - 1];
            const int c1 = csuf[j];
            const float sa = c0 == 0 ? 0.f : rmul(box_area(b0), (float)c0);
            const float sb = c1 == 0 ? 0.f : rmul(box_area(suffix[j]), (float)c1);
            const float cost = radd(sa, sb);
            if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bucket = j; best_left = c0; best_l = b0; best_r = suffix[j]; }
        }
    }
    if (best_axis == -1) { t.left[s] = kLeaf; return; }                               // src/bvh.cpp:113
    const int c = atomicAdd(&ctr->n_nodes, 2);
    if (c + 2 > node_capacity) { t.left[s] = kLeaf; return; }                         // host reports ENOMEM from n_nodes
    t.left[s] = c;
    t.slot[s] = best_axis | (best_bucket << 2);
    const int cnts[2] = {best_left, count - best_left};
    const int begins[2] = {begin, begin + best_left};
    const Box3* boxes[2] = {&best_l, &best_r};
    for (int k = 0; k < 2; ++k) {
        const Box3& b = *boxes[k];
        t.lo[c + k] = make_float4(b.lo[0], b.lo[1], b.lo[2], __int_as_float(begins[k]));
        t.hi[c + k] = make_float4(b.hi[0], b.hi[1], b.hi[2], __int_as_float(cnts[k]));
        const bool leaf = small_or_thin(cnts[k], b.lo, b.hi);
        t.left[c + k] = leaf ? kLeaf : kCandidate;
        t.slot[c + k] = leaf ? 0 : atomicAdd(&ctr->n_candidates, 1);
        t.parent[c + k] = s;
    }
}

// 1 for primitives that go to the left child of a node split on this level (src/bvh.cpp:131-148)
__global__ void k_flags(int n, int level_start, const int* __restrict__ ids, const int* __restrict__ seg, Tree t,
                        const float4* __restrict__ blo, const float4* __restrict__ bhi, int* flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = seg[i];
    int f = 0;
    if (s >= level_start && t.left[s] >= 0) {
        const int axis = t.slot[s] & 3, bucket = t.slot[s] >> 2;
        const int id = ids[i];
        const float pl = (&blo[id].x)[axis], ph = (&bhi[id].x)[axis];
        f = bucket_of(pl, ph, (&t.lo[s].x)[axis], (&t.hi[s].x)[axis]) < bucket;
    }
    flags[i] = f;
}

__global__ void k_scatter(int n, int level_start, const int* __restrict__ ids, const int* __restrict__ seg, Tree t,
                          const int* __restrict__ flags, const int* __restrict__ scan, int* ids_out, int* seg_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = seg[i];
    int pos = i, ns = s;
    if (s >= level_start && t.left[s] >= 0) {
        const int l = t.left[s];
        const int begin = __float_as_int(t.lo[s].w), n_left = __float_as_int(t.hi[l].w);
        const int left_rank = scan[i] - scan[begin];
        if (flags[i]) { pos = begin + left_rank; ns = l; }
        else { pos = begin + n_left + (i - begin - left_rank); ns = l + 1; }
    }
    ids_out[pos] = ids[i];
    seg_out[pos] = ns;
}

// Sign of zero, part 1: a primitive with a zero bound walks from its leaf to the root and offers its input index
// to every node whose same bound is zero too (a node's lo is <= its descendants' lo: once it is negative it stays so).
__global__ void k_zero_ids(int n, const int* __restrict__ ids, const int* __restrict__ seg, Tree t,
                           const float4* __restrict__ blo, const float4* __restrict__ bhi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int id = ids[i];
    const float4 l4 = blo[id], h4 = bhi[id];
    const float pb[6] = {l4.x, l4.y, l4.z, h4.x, h4.y, h4.z};
    unsigned live = 0;
    for (int a = 0; a < 6; ++a) live |= (pb[a] == 0.f) << a;
    for (int s = seg[i]; s >= 0 && live; s = t.parent[s]) {
        const float4 nl = t.lo[s], nh = t.hi[s];
        const float nb[6] = {nl.x, nl.y, nl.z, nh.x, nh.y, nh.z};
        for (int a = 0; a < 6; ++a) {
            if (!(live >> a & 1)) continue;
            if (nb[a] == 0.f) atomicMin(&t.zero_id[6 * (size_t)s + a], id);
            else live &= ~(1u << a);
        }
    }
}
// part 2: the node's zero takes the sign of that primitive's zero (first operand kept on ties, src/cutil_math.h:36-44)
__global__ void k_zero_sign(int total, Tree t, const float4* __restrict__ blo, const float4* __restrict__ bhi) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= total) return;
    float4 nl = t.lo[s], nh = t.hi[s];
    float* nb[6] = {&nl.x, &nl.y, &nl.z, &nh.x, &nh.y, &nh.z};
    bool any = false;
    for (int a = 0; a < 6; ++a) {
        if (*nb[a] != 0.f) continue;
        const int id = t.zero_id[6 * (size_t)s + a];
        if (id == 0x7f7f7f7f) continue;
        *nb[a] = a < 3 ? (&blo[id].x)[a] : (&bhi[id].x)[a - 3];
        any = true;
    }
    if (any) { t.lo[s] = nl; t.hi[s] = nh; }
}

__global__ void k_sizes(int level_start, int level_end, Tree t) {
    const int s = level_start + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= level_end) return;
    const int l = t.left[s];
    t.size[s] = l < 0 ? 1 : 1 + t.size[l] + t.size[l + 1];
}

// BVH::flatten (src/bvh.cpp:150-173): pre-order records, left child adjacent.
__global__ void k_emit(int level_start, int level_end, Tree t, RefLinearBVHNode* out) {
    const int s = level_start + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= level_end) return;
    const int me = s == 0 ? 0 : t.pre[s];
    const int l = t.left[s];
    const float4 lo = t.lo[s], hi = t.hi[s];
    RefLinearBVHNode nd;
    nd.fmin[0] = lo.x; nd.fmin[1] = lo.y; nd.fmin[2] = lo.z; nd.fmax[0] = hi.x; nd.fmax[1] = hi.y; nd.fmax[2] = hi.z;
    nd._pad[0] = nd._pad[1] = nd._pad[2] = 0;
    if (l < 0) {
        const int begin = __float_as_int(lo.w), count = __float_as_int(hi.w);
        nd.second_child_offset = -1; nd.is_leaf = 1;
        nd.start = count ? begin : -1; nd.end = count ? begin + count - 1 : -1;
    } else {
        t.pre[l] = me + 1;
        t.pre[l + 1] = me + 1 + t.size[l];
        nd.second_child_offset = me + 1 + t.size[l]; nd.is_leaf = 0; nd.start = nd.end = -1;
    }
    out[me] = nd;
}

#ifdef B200PT_EMULATE
static void exclusive_scan(const int* in, int* out, int n) { int s = 0; for (int i = 0; i < n; ++i) { out[i] = s; s += in[i]; } }
#endif

// one device allocation, carved up: pass 0 (base == nullptr) only sizes it
struct Arena {
    char* base = nullptr; size_t off = 0;
    template <class T> T* take(size_t count) {
        T* p = base ? (T*)(base + off) : nullptr;
        off += (sizeof(T) * count + 255) & ~(size_t)255;
        return p;
    }
};

inline int grid_for(long long n) { return (int)((n + kThreads - 1) / kThreads); }

}  // namespace bvhb

using namespace bvhb;

// ---- exclusive prefix sum of the partition flags (one per level): three small kernels — per-tile scan + tile totals,
// scan of the totals, add.  The arrays are <= 16 MB and L2-resident; the levels are launch- and atomic-latency bound,
// so nothing fancier (decoupled look-back) pays here.
constexpr int kScanThreads = 256, kScanItems = 8, kScanTile = kScanThreads * kScanItems;
#ifndef B200PT_EMULATE
__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int& total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int x = v;
    for (int off = 1; off < 32; off <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, off); if (lane >= off) x += y; }
    if (lane == 31) s_warp[w] = x;
    __syncthreads();
    if (w == 0) {
        int t = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0;
        for (int off = 1; off < 32; off <<= 1) { int y = __shfl_up_sync(0xffffffffu, t, off); if (lane >= off) t += y; }
        s_warp[lane] = t;                       // inclusive scan of the warp totals
    }
    __syncthreads();
    total = s_warp[(blockDim.x >> 5) - 1];
    const int base = w ? s_warp[w - 1] : 0;
    __syncthreads();
    return base + x - v;
}
__global__ void __launch_bounds__(kScanThreads) k_scan_tiles(const int* __restrict__ in, int* __restrict__ out, int* __restrict__ tile_sums, int n) {
    __shared__ int s_warp[32];
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int v[kScanItems], sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) { v[k] = base + k < n ? in[base + k] : 0; sum += v[k]; }
    int total;
    int run = block_exclusive_scan(sum, s_warp, total);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) { if (base + k < n) out[base + k] = run; run += v[k]; }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) k_scan_sums(int* tile_sums, int n_tiles) {         // one CTA
    __shared__ int s_warp[32];
    int carry = 0;
    for (int base = 0; base < n_tiles; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n_tiles ? tile_sums[i] : 0;
        int total;
        const int ex = block_exclusive_scan(v, s_warp, total);
        if (i < n_tiles) tile_sums[i] = carry + ex;
        carry += total;
    }
}
__global__ void __launch_bounds__(kScanThreads) k_scan_add(int* __restrict__ out, const int* __restrict__ tile_sums, int n) {
    const int off = tile_sums[blockIdx.x];
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) if (base + k < n) out[base + k] += off;
}
#endif

#define BCK(call)                                                                                             \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) {                                                                              \
            std::string m_ = std::string(#call) + ": " + cudaGetErrorString(e_) + " (bvh_build.cu:" + std::to_string(__LINE__) + ")"; \
            for (void* p_ : allocs) cudaFree(p_);                                                             \
            return b200pt_internal_fail(B200PT_ECUDA, m_.c_str());                                            \
        }                                                                                                     \
    } while (0)

extern "C" int b200pt_bvh_build_gpu(const void* prims_in, int32_t n_prims, void* prims_out, void* nodes_out,
                                    int32_t nodes_capacity, int32_t* n_nodes, float* root_box6, int32_t device,
                                    double* timing_ms4) {
    if (!prims_in || !prims_out || !nodes_out || !n_nodes || n_prims <= 0 || nodes_capacity <= 0)
        return b200pt_internal_fail(B200PT_EINVAL, "b200pt_bvh_build_gpu: null argument or empty input");
    int bad = -1;
    if (!b200pt_internal_prims_finite(prims_in, n_prims, &bad))
        return b200pt_internal_fail(B200PT_EINVAL, ("b200pt_bvh_build_gpu: primitive " + std::to_string(bad) + " has a non-finite bounding box").c_str());
    std::vector<void*> allocs;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return b200pt_internal_fail(B200PT_ECUDA, "b200pt_bvh_build_gpu: no CUDA device");
    if (device < 0 || device >= ndev) return b200pt_internal_fail(B200PT_EINVAL, "b200pt_bvh_build_gpu: bad device index");
    BCK(cudaSetDevice(device));
    const auto wall0 = std::chrono::steady_clock::now();
    const int n = n_prims;
    const int cap = nodes_capacity < 2 * n ? nodes_capacity : 2 * n;      // a BVH2 over n primitives has < 2n nodes
    const size_t max_candidates = (size_t)n / 5 + 1;                      // a candidate holds >= 5 primitives

    // one device allocation, carved up (allocation calls dominate the wall clock of small builds otherwise)
    const int n_tiles = (n + kScanTile - 1) / kScanTile;
    RefPrimitive* d_prims; float4 *d_blo, *d_bhi; int *d_ids0, *d_ids1, *d_seg0, *d_seg1, *d_flags, *d_scan;
    unsigned* d_buckets; Counters* d_ctr; RefLinearBVHNode* d_out; int* d_tile_sums; Tree t;
    Arena arena;
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
            void* base = nullptr;
            if (cudaMalloc(&base, arena.off) != cudaSuccess)
                return b200pt_internal_fail(B200PT_ENOMEM, "b200pt_bvh_build_gpu: device allocation failed");
            allocs.push_back(base);
            arena.base = (char*)base; arena.off = 0;
        }
        d_prims = arena.take<RefPrimitive>(n);
        d_blo = arena.take<float4>(n); d_bhi = arena.take<float4>(n);
        d_ids0 = arena.take<int>(n); d_ids1 = arena.take<int>(n);
        d_seg0 = arena.take<int>(n); d_seg1 = arena.take<int>(n);
        d_flags = arena.take<int>(n); d_scan = arena.take<int>(n);
        d_buckets = arena.take<unsigned>(max_candidates * kNodeWords);
        d_ctr = arena.take<Counters>(1);
        d_out = arena.take<RefLinearBVHNode>(cap);
        d_tile_sums = arena.take<int>((size_t)n_tiles + 1);
        t.lo = arena.take<float4>(cap); t.hi = arena.take<float4>(cap);
        t.left = arena.take<int>(cap); t.slot = arena.take<int>(cap); t.size = arena.take<int>(cap);
        t.pre = arena.take<int>(cap); t.parent = arena.take<int>(cap); t.zero_id = arena.take<int>(6 * (size_t)cap);
    }
    cudaStream_t st = 0;
    cudaEvent_t ev[4];
    for (int k = 0; k < 4; ++k) BCK(cudaEventCreate(&ev[k]));

    BCK(cudaEventRecord(ev[0], st));
    BCK(cudaMemcpyAsync(d_prims, prims_in, sizeof(RefPrimitive) * (size_t)n, cudaMemcpyHostToDevice, st));
    BCK(cudaEventRecord(ev[1], st));
    BCK(cudaMemsetAsync(d_ctr, 0, sizeof(Counters), st));
    PT_LAUNCH(k_prim_boxes, grid_for(n), kThreads, 0, st, d_prims, n, d_blo, d_bhi, d_ctr);
    PT_LAUNCH(k_init, grid_for(n), kThreads, 0, st, n, d_ids0, d_seg0, t, d_ctr);

    int* ids = d_ids0; int* ids_next = d_ids1; int* seg = d_seg0; int* seg_next = d_seg1;
    std::vector<int> level_start{0};
    int level_end = 1;
    Counters hc;
    BCK(cudaMemcpyAsync(&hc, d_ctr, sizeof(hc), cudaMemcpyDeviceToHost, st));
    BCK(cudaStreamSynchronize(st));
    while (hc.n_candidates > 0) {
        const int ls = level_start.back();
        if ((size_t)hc.n_candidates > max_candidates) { for (void* p_ : allocs) cudaFree(p_); return b200pt_internal_fail(B200PT_ECUDA, "b200pt_bvh_build_gpu: candidate count out of range"); }
        BCK(cudaMemsetAsync(d_buckets, 0, sizeof(unsigned) * kNodeWords * (size_t)hc.n_candidates, st));
        PT_LAUNCH(k_hist, grid_for(n), kThreads, 0, st, n, ids, seg, t, d_blo, d_bhi, d_buckets);
        BCK(cudaMemsetAsync(&d_ctr->n_candidates, 0, sizeof(int), st));
        PT_LAUNCH(k_split, grid_for(level_end - ls), kThreads, 0, st, ls, level_end, t, d_buckets, d_ctr, cap);
        PT_LAUNCH(k_flags, grid_for(n), kThreads, 0, st, n, ls, ids, seg, t, d_blo, d_bhi, d_flags);
#ifndef B200PT_EMULATE
        k_scan_tiles<<<n_tiles, kScanThreads, 0, st>>>(d_flags, d_scan, d_tile_sums, n);
        k_scan_sums<<<1, 1024, 0, st>>>(d_tile_sums, n_tiles);
        k_scan_add<<<n_tiles, kScanThreads, 0, st>>>(d_scan, d_tile_sums, n);
#else
        exclusive_scan(d_flags, d_scan, n);
#endif
        PT_LAUNCH(k_scatter, grid_for(n), kThreads, 0, st, n, ls, ids, seg, t, d_flags, d_scan, ids_next, seg_next);
        std::swap(ids, ids_next); std::swap(seg, seg_next);
        BCK(cudaMemcpyAsync(&hc, d_ctr, sizeof(hc), cudaMemcpyDeviceToHost, st));
        BCK(cudaStreamSynchronize(st));
        if (hc.n_nodes > cap) { for (void* p_ : allocs) cudaFree(p_); return b200pt_internal_fail(B200PT_ENOMEM, "b200pt_bvh_build_gpu: nodes_capacity too small"); }
        if (hc.n_nodes == level_end) break;                  // nothing split on this level
        level_start.push_back(level_end);
        level_end = hc.n_nodes;
    }
    const int total = hc.n_nodes;
    const int levels = (int)level_start.size();
    BCK(cudaMemsetAsync(t.zero_id, 0x7f, sizeof(int) * 6 * (size_t)total, st));
    PT_LAUNCH(k_zero_ids, grid_for(n), kThreads, 0, st, n, ids, seg, t, d_blo, d_bhi);
    PT_LAUNCH(k_zero_sign, grid_for(total), kThreads, 0, st, total, t, d_blo, d_bhi);
    for (int l = levels - 1; l >= 0; --l) {
        const int ls = level_start[l], le = l + 1 < levels ? level_start[l + 1] : total;
        PT_LAUNCH(k_sizes, grid_for(le - ls), kThreads, 0, st, ls, le, t);
    }
    for (int l = 0; l < levels; ++l) {
        const int ls = level_start[l], le = l + 1 < levels ? level_start[l + 1] : total;
        PT_LAUNCH(k_emit, grid_for(le - ls), kThreads, 0, st, ls, le, t, d_out);
    }
    BCK(cudaEventRecord(ev[2], st));
    BCK(cudaMemcpyAsync(nodes_out, d_out, sizeof(RefLinearBVHNode) * (size_t)total, cudaMemcpyDeviceToHost, st));
    std::vector<int> order(n);
    BCK(cudaMemcpyAsync(order.data(), ids, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
    BCK(cudaEventRecord(ev[3], st));
    BCK(cudaStreamSynchronize(st));
    BCK(cudaGetLastError());
    {   // leaf order of the 176-B primitives: the permutation came back (4 B each), the records never left the host
        const RefPrimitive* src = (const RefPrimitive*)prims_in;
        RefPrimitive* dst = (RefPrimitive*)prims_out;
        unsigned nt = std::thread::hardware_concurrency();
        nt = nt < 1 ? 1 : (nt > 16 ? 16 : nt);
        if (n < 65536) nt = 1;
        std::vector<std::thread> pool;
        for (unsigned k = 0; k < nt; ++k) {
            const size_t i0 = (size_t)n * k / nt, i1 = (size_t)n * (k + 1) / nt;
            auto work = [=, &order]() { for (size_t i = i0; i < i1; ++i) std::memcpy(&dst[i], &src[order[i]], sizeof(RefPrimitive)); };
            if (nt == 1) work(); else pool.emplace_back(work);
        }
        for (auto& th : pool) th.join();
    }
    *n_nodes = total;
    if (root_box6) for (int a = 0; a < 3; ++a) { root_box6[a] = 0; root_box6[3 + a] = 0; }
    if (root_box6) {
        const RefLinearBVHNode* root = (const RefLinearBVHNode*)nodes_out;
        std::memcpy(root_box6, root->fmin, 12); std::memcpy(root_box6 + 3, root->fmax, 12);
    }
    if (timing_ms4) {
        float a = 0, b = 0, c = 0;
        cudaEventElapsedTime(&a, ev[0], ev[1]); cudaEventElapsedTime(&b, ev[1], ev[2]); cudaEventElapsedTime(&c, ev[2], ev[3]);
        timing_ms4[0] = a; timing_ms4[1] = b; timing_ms4[2] = c;
        timing_ms4[3] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    }
    for (int k = 0; k < 4; ++k) cudaEventDestroy(ev[k]);
    for (void* p : allocs) cudaFree(p);
    return B200PT_OK;
}
