// k_trace.cuh — persistent-threads BVH traversal kernel (replaces Intersect / IntersectP / the ray walks of
// Tr, src/pathtracer.cu:214-322).
//
// * One lane = one ray.  Rays come from the compact ray queue that k_shade fills (continuation closest-hit,
//   shadow any-hit / transmittance walk, MIS closest-hit); a persistent grid of SM-count x resident-CTA warps
//   pulls them with warp-ballot refills, so idle lanes are re-armed instead of waiting for the slowest ray.
// * Every CTA first stages the top of the breadth-first numbered acceleration structure (two-child 64-B nodes)
//   and, for small scenes, the 48-B primitive records into shared memory with TMA bulk copies
//   (cp.async.bulk + mbarrier); deeper nodes / primitives are read from L2/HBM with 16-B vector loads.
// * Ordered traversal (near child first) with the reference's exact slab and Moeller-Trumbore arithmetic, so
//   hit/miss decisions are bit-identical to the reference's unordered DFS; exact-t ties resolve to the higher
//   primitive index, which is what the reference's visiting order produces (src/mesh.h:64 accepts tt == tmax).
#pragma once
#include "wavefront.cuh"

namespace pt {


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
        f3 s1 = cross_pinned(d, e2);
        float divisor = dot_pinned(s1, e1);
        if (fabsf(divisor) < 1e-8f) return 0;
        float invDivisor = 1.0f / divisor;       // == (float)(1.0 / (double)divisor): IEEE division, 53 >= 2*24+2
        f3 s = o - v0;
        float b1 = dot_pinned(s, s1) * invDivisor;
        if (b1 < 0.0f || b1 > 1.0f) return 0;
        f3 s2 = cross_pinned(s, e1);
        float b2 = dot_pinned(d, s2) * invDivisor;
        if (b2 < 0.0f || b1 + b2 > 1.0f) return 0;
        float tt = dot_pinned(e2, s2) * invDivisor;
        if (tt < tmin || tt > tmax) return 0;
        t_out = tt; b1_out = b1; b2_out = b2;
        return 1;
    } else if (__float_as_int(q2.y) == 2) {      // hair segment, Line::Intersect (src/line.h:33-86)
        f3 p0 = mk3(q0.x, q0.y, q0.z), p1 = mk3(q0.w, q1.x, q1.y);
        float width0 = q1.z, width1 = q1.w;
        f3 u = d;
        f3 v = p1 - p0;
        f3 w = o - p0;
        float a = dot(u, u);
        float b = dot(u, v);
        float c = dot(v, v);
        float dd = dot(u, w);
        float e = dot(v, w);
        float det = a * c - b * b;
        if (det == 0) return 0;
        float t = (b * e - c * dd) / det;
        float s = (a * e - b * dd) / det;
        if (t < tmin || t > tmax) return 0;
        s = clampf(s, 0.f, 1.f);
        f3 pr = o + d * t;
        f3 pl = p0 + (p1 - p0) * s;
        f3 prl = pr - pl;
        float d2 = dot(prl, prl);
        float r = width0 * (1 - s) + width1 * s;
        if (d2 > r * r) return 0;
        t_out = t; b1_out = s; b2_out = sqrtf(d2) / r;          // (b1, b2) carry the segment uv (src/line.h:77)
        return 1;
    } else {                                     // sphere: q0 = centre.xyz, radius
        f3 op = o - mk3(q0.x, q0.y, q0.z);
        float radius = q0.w;
        float B = dot_pinned(op, d);
#if defined(__CUDA_ARCH__)
        // B*B - (dot(op,op) - r*r) as the reference build contracts it at all seven Sphere::Intersect sites of Path and
        // Volpath (SASS: FFMA c, r, r, -dot ; FFMA delta, B, B, c): fma(B, B, fma(r, r, -dot(op,op))) — BOTH products fused
        float delta = __fmaf_rn(B, B, __fmaf_rn(radius, radius, -dot_pinned(op, op)));
#else
        float C = dot(op, op) - radius * radius;
        float delta = B * B - C;
#endif
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

struct TraceArgs {
    SceneDev sc;
    Pool pool;
    RayQueue q;
    Counters* counters;
    uint32_t parity;                                 // which QueueCtl set this step consumes
    int32_t refill_below;                            // refill a warp when fewer lanes than this still carry a ray
    uint32_t stage_bytes_nodes, stage_bytes_prims;   // > 0: bytes staged into shared memory with TMA
    uint32_t small_prim_bytes;                       // k_trace_small: bytes of ALL primitive records (always staged)
    const float4* leaves;                            // small scenes: primitive-group records (box + <= 4 primitives)
    int32_t n_leaves;
    int32_t sec_tmax;                                // kind-2 rays carry their own tmax in misd.w (heterogeneous-media wavefront: Tr() segments)
    // `pt`, scene without area emitters (environment light only): the MIS ray adds radiance iff it ESCAPES, so it is an
    // any-hit query — "some primitive is hit" and "a closest hit exists" are the same event (a larger tmax only lets more
    // boxes and primitives pass), and the primitive reported carries no light index either way
    int32_t mis_anyhit;
};

constexpr int kDone = (int)0x80000000;               // traversal cursor: nothing left to visit (also "no postponed leaf")
constexpr int kPop = (int)0x80000001;                // traversal cursor: take the next entry from the stack
constexpr int kTraceThreads = 256;
constexpr int kSmemStack = 12;                       // stack levels kept in shared memory (8 B x 256 threads each)
constexpr size_t kTraceStackBytes = (size_t)kSmemStack * kTraceThreads * 8;

// One persistent warp = 32 independent rays in flight.  Lanes whose ray has finished are re-armed from the ray
// queue as soon as fewer than `refill_below` lanes are busy (warp ballot + one aggregated atomic), so the
// SIMT width stays full even though incoherent rays need very different numbers of node visits.  The traversal
// is "while-while": descend inner nodes until a leaf is reached, then test that leaf's primitives; all lanes
// reconverge after each (descend, leaf) round.
//
// Nodes are numbered breadth-first; the first `stage_bytes_nodes / 64` of them (the top of the tree, or the whole
// tree for small scenes) and, when they fit, all primitive records are staged into shared memory by one TMA bulk
// copy per CTA.  Everything else is read from L2/HBM with 16-byte vector loads.
// WIDE: four-child nodes (WNode4) instead of two-child ones.
template <bool VOL, bool WIDE>
__global__ void __launch_bounds__(kTraceThreads) k_trace(const TraceArgs a) {
    const WNode* __restrict__ gnodes = a.sc.nodes;
    const WNode4* __restrict__ gnodes4 = a.sc.nodes4;
    const WPrim* __restrict__ gprims = a.sc.prims;
    int n_staged = 0;
    bool prims_staged = false;
#ifndef B200PT_EMULATE
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    const WNode* s_nodes = reinterpret_cast<const WNode*>(smem_raw);
    const WNode4* s_nodes4 = reinterpret_cast<const WNode4*>(smem_raw);
    const WPrim* s_prims = reinterpret_cast<const WPrim*>(smem_raw + a.stage_bytes_nodes);
    if (a.stage_bytes_nodes + a.stage_bytes_prims > 0) {
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, a.stage_bytes_nodes + a.stage_bytes_prims);
            if (a.stage_bytes_nodes) tma_bulk_g2s(smem_raw, WIDE ? (const void*)a.sc.nodes4 : (const void*)a.sc.nodes, a.stage_bytes_nodes, &bar);
            if (a.stage_bytes_prims) tma_bulk_g2s(smem_raw + a.stage_bytes_nodes, a.sc.prims, a.stage_bytes_prims, &bar);
        }
        mbar_wait(&bar, 0);
        n_staged = (int)(a.stage_bytes_nodes / (WIDE ? sizeof(WNode4) : sizeof(WNode)));
        prims_staged = a.stage_bytes_prims > 0;
    }
#endif
    const uint32_t lane = pt_lane();
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t par = a.parity & 1u;
    const uint32_t tail = a.q.ctl->tail[par];
    uint32_t* head = &a.q.ctl->head[par];
    if (blockIdx.x == 0 && threadIdx.x == 0) { a.q.ctl->tail[par ^ 1u] = 0u; a.q.ctl->head[par ^ 1u] = 0u; }

    // per-lane ray state
    bool active = false, exhausted = false, anyhit = false;
    uint32_t entry = 0u;
    f3 o = mk3(0, 0, 0), d = mk3(0, 0, 1), inv = mk3(0, 0, 0);
    float tmax = 0.f, hb1 = 0.f, hb2 = 0.f;
    int hprim = -1, cur = kDone, leaf = kDone, sp = 0;
    // Traversal stack: (subtree, entry distance) pairs.  The first kSmemStack levels live in shared memory, laid out
    // [level][thread] so any mix of depths across a warp is bank-conflict free; deeper levels (rare with near-first
    // ordering) spill to thread-local memory.  Keeping the hot levels out of local memory also keeps them out of L1,
    // which the node / primitive fetches of a large scene need for themselves.
    int stack[64];
    float stack_t[64];
#ifndef B200PT_EMULATE
    int* const s_stk = reinterpret_cast<int*>(smem_raw + a.stage_bytes_nodes + a.stage_bytes_prims) + threadIdx.x;
    float* const s_stk_t = reinterpret_cast<float*>(s_stk + kSmemStack * kTraceThreads);
#define STK_PUSH(node_, t_) do { if (sp < kSmemStack) { s_stk[sp * kTraceThreads] = (node_); s_stk_t[sp * kTraceThreads] = (t_); } \
                                 else { stack[sp - kSmemStack] = (node_); stack_t[sp - kSmemStack] = (t_); } ++sp; } while (0)
#define STK_POP(node_, t_) do { --sp; if (sp < kSmemStack) { (node_) = s_stk[sp * kTraceThreads]; (t_) = s_stk_t[sp * kTraceThreads]; } \
                                else { (node_) = stack[sp - kSmemStack]; (t_) = stack_t[sp - kSmemStack]; } } while (0)
#else
#define STK_PUSH(node_, t_) do { stack[sp] = (node_); stack_t[sp] = (t_); ++sp; } while (0)
#define STK_POP(node_, t_) do { --sp; (node_) = stack[sp]; (t_) = stack_t[sp]; } while (0)
#endif
    f3 tr = mk3(1, 1, 1);          // vpt shadow rays: transmittance so far, remaining length, current medium
    float remain = 0.f;
    int medium = -1;
    uint32_t nrays = 0;
    const float eps = a.sc.eps;

    for (;;) {
        // ---- refill idle lanes from the queue
        const uint32_t idle = __ballot_sync(kFullMask, !active);
        if (idle != 0u && !exhausted) {
            const uint32_t n = (uint32_t)__popc(idle);
            uint32_t base = 0u;
            if (lane == 0u) base = atomicAdd(head, n);
            base = __shfl_sync(kFullMask, base, 0);
            const uint32_t idx = base + (uint32_t)__popc(idle & lt);
            if (!active && idx < tail) {
                entry = a.q.entries[idx];
                const uint32_t slot = entry & kSlotMask, kind = entry >> kKindShift;
                const float4 orng = kind == 0u ? a.pool.o_rng[slot] : a.pool.pend_o[slot];     // shadow / MIS rays keep their own origin
                o = mk3(orng.x, orng.y, orng.z);
                float4 dv;
                if (kind == 0u) { dv = a.pool.d_flags[slot]; dv.w = INFINITY; }
                else if (kind == 1u) dv = a.pool.shd[slot];
                else { dv = a.pool.misd[slot]; if (!a.sec_tmax) dv.w = INFINITY; }
                d = mk3(dv.x, dv.y, dv.z);
                tmax = dv.w;
                anyhit = !VOL && (kind == 1u || (kind == 2u && a.mis_anyhit != 0));
                if (VOL && kind == 1u) {
                    const uint32_t flags = __float_as_uint(a.pool.d_flags[slot].w);
                    medium = (int)((flags >> kMedium2Shift) & 0xffu) - 1;
                    tr = mk3(1, 1, 1);
                    remain = tmax;
                }
                inv = mk3(1.f / d.x, 1.f / d.y, 1.f / d.z);
                hprim = -1; sp = 0; leaf = kDone;
                float tn;
                // the reference tests the root's own box first (node 0, src/pathtracer.cu:222-223)
                const bool in = slab(a.sc.root_min[0], a.sc.root_min[1], a.sc.root_min[2], a.sc.root_max[0], a.sc.root_max[1], a.sc.root_max[2], o, inv, tmax, tn);
                cur = !in ? kDone : (a.sc.root_leaf_count > 0 ? ~0 : 0);
                active = true;
                ++nrays;
            }
            exhausted = base + n >= tail;
        }
        uint32_t busy = __ballot_sync(kFullMask, active);
        if (busy == 0u) break;

        // ---- traversal rounds until too few lanes are busy (then refill) or, with the queue drained, all are done
        do {
            if (active) {
                // descend: inner nodes, two slab tests per 64-B record, near child first.  The first leaf a lane
                // reaches is POSTPONED and the lane keeps descending speculatively while any other lane of the warp
                // is still looking for its leaf, so the (expensive) primitive tests start with most lanes on board.
                for (;;) {
                    if (WIDE && cur >= 0) {
                        // four child boxes per 128-B record: the reference's slab test on each, hits ordered near to far,
                        // the far ones pushed (farthest first), the nearest visited next
                        const float4* np = reinterpret_cast<const float4*>(gnodes4 + cur);
#ifndef B200PT_EMULATE
                        if (cur < n_staged) np = reinterpret_cast<const float4*>(s_nodes4 + cur);
#endif
                        const float4 mnx = np[0], mny = np[1], mnz = np[2], mxx = np[3], mxy = np[4], mxz = np[5];
                        const int4 lk = *reinterpret_cast<const int4*>(np + 6);
                        float t0 = INFINITY, t1 = INFINITY, t2 = INFINITY, t3 = INFINITY, tn;
                        int c0 = lk.x, c1 = lk.y, c2 = lk.z, c3 = lk.w, n = 0;
                        if (lk.x != kEmptyChild && slab(mnx.x, mny.x, mnz.x, mxx.x, mxy.x, mxz.x, o, inv, tmax, tn)) { t0 = fminf(tn, 3.0e38f); ++n; }
                        if (lk.y != kEmptyChild && slab(mnx.y, mny.y, mnz.y, mxx.y, mxy.y, mxz.y, o, inv, tmax, tn)) { t1 = fminf(tn, 3.0e38f); ++n; }
                        if (lk.z != kEmptyChild && slab(mnx.z, mny.z, mnz.z, mxx.z, mxy.z, mxz.z, o, inv, tmax, tn)) { t2 = fminf(tn, 3.0e38f); ++n; }
                        if (lk.w != kEmptyChild && slab(mnx.w, mny.w, mnz.w, mxx.w, mxy.w, mxz.w, o, inv, tmax, tn)) { t3 = fminf(tn, 3.0e38f); ++n; }
#define PT_CAS(ta, ca, tb, cb) { if (tb < ta) { const float f_ = ta; ta = tb; tb = f_; const int i_ = ca; ca = cb; cb = i_; } }
                        PT_CAS(t0, c0, t1, c1) PT_CAS(t2, c2, t3, c3) PT_CAS(t0, c0, t2, c2) PT_CAS(t1, c1, t3, c3) PT_CAS(t1, c1, t2, c2)
#undef PT_CAS
                        if (n > 3) STK_PUSH(c3, t3);
                        if (n > 2) STK_PUSH(c2, t2);
                        if (n > 1) STK_PUSH(c1, t1);
                        cur = n ? c0 : kPop;
                    }
                    if (!WIDE && cur >= 0) {
                        float4 q0, q1, q2; int2 link;
#ifndef B200PT_EMULATE
                        if (cur < n_staged) {
                            const float4* np = reinterpret_cast<const float4*>(s_nodes + cur);
                            q0 = np[0]; q1 = np[1]; q2 = np[2]; link = *reinterpret_cast<const int2*>(np + 3);
                        } else
#endif
                        {
                            const float4* np = reinterpret_cast<const float4*>(gnodes + cur);
                            q0 = np[0]; q1 = np[1]; q2 = np[2]; link = *reinterpret_cast<const int2*>(np + 3);
                        }
                        float tl = 0.f, tr_ = 0.f;
                        const bool hl = slab(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, o, inv, tmax, tl);
                        const bool hr = link.y != kEmptyChild && slab(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, o, inv, tmax, tr_);
                        int c0 = link.x, c1 = link.y;
                        if (hl && hr) {
                            if (tr_ < tl) { const int t_ = c0; c0 = c1; c1 = t_; const float f_ = tl; tl = tr_; tr_ = f_; }
                            STK_PUSH(c1, tr_);
                            cur = c0;
                        } else if (hl) cur = c0;
                        else if (hr) cur = c1;
                        else cur = kPop;
                    }
                    if (cur == kPop) {
                        // one pop per iteration (right after the visit that ran dry); subtrees whose entry distance is
                        // already behind the closest hit are dropped (the same `tmin > ray.tmax` rejection
                        // BBox::Intersect would make on visiting them, src/bbox.h:93)
                        if (sp > 0) { int n_; float t_; STK_POP(n_, t_); cur = (t_ > tmax) ? kPop : n_; }
                        else cur = kDone;
                    }
                    if (cur < 0 && cur > kPop && leaf == kDone) { leaf = cur; cur = kPop; }   // postpone the first leaf
                    if (cur < 0 && cur != kPop) break;                                // second leaf, or nothing left
#ifndef B200PT_EMULATE
                    if (!__any_sync(__activemask(), leaf == kDone)) break;            // every lane has its leaf
#else
                    if (leaf != kDone) break;
#endif
                }
                // leaves: Moeller-Trumbore / sphere test, one primitive per iteration (reference leaf loop, :230-245);
                // a lane that found a second leaf while speculating continues straight into it
                int pi = ~leaf;
                while (leaf != kDone) {
                    float4 p0, p1, p2;
#ifndef B200PT_EMULATE
                    if (prims_staged) {
                        const float4* pp = reinterpret_cast<const float4*>(s_prims + pi);
                        p0 = pp[0]; p1 = pp[1]; p2 = pp[2];
                    } else
#endif
                    {
                        const float4* pp = reinterpret_cast<const float4*>(gprims + pi);
                        p0 = pp[0]; p1 = pp[1]; p2 = pp[2];
                    }
                    float t, b1, b2;
                    const int acc = prim_test(p0, p1, p2, o, d, eps, tmax, t, b1, b2);
                    if (acc) {
                        if (anyhit) { hprim = pi; cur = kDone; sp = 0; leaf = kDone; break; }
                        // tt == tmax is accepted by the reference; the later (higher index) primitive then wins
                        if (acc == 2 || t < tmax || pi > hprim) { hprim = pi; hb1 = b1; hb2 = b2; }
                        tmax = t;
                    }
                    if (__float_as_int(p2.z) == 0) { ++pi; continue; }                 // more primitives in this leaf
                    if (cur < 0 && cur > kPop) { leaf = cur; pi = ~cur; cur = kPop; }  // the leaf found while speculating
                    else leaf = kDone;
                }
                if (cur == kPop) {                                                    // resolve a pending pop before the round ends
                    cur = kDone;
                    while (sp > 0) { int n_; float t_; STK_POP(n_, t_); if (!(t_ > tmax)) { cur = n_; break; } }
                    if (cur < 0 && cur != kDone) { leaf = cur; cur = kPop; }          // a leaf: handled first thing next round
                }
                // ---- finished: write the result (or start the next segment of a transmittance walk)
                if (cur == kDone) {
                    const uint32_t slot = entry & kSlotMask, kind = entry >> kKindShift;
                    active = false;
                    if (kind != 1u) {
                        const float4 h = make_float4(hprim >= 0 ? tmax : -1.f, __int_as_float(hprim), hb1, hb2);
                        if (kind == 0u) a.pool.hit0[slot] = h; else a.pool.hit1[slot] = h;
                    } else if (!VOL) {
                        const float v = hprim >= 0 ? 0.f : 1.f;
                        a.pool.vis[slot] = make_float4(v, v, v, 0.f);
                    } else {
                        // Tr() (src/pathtracer.cu:298-322): closest hits until an opaque surface blocks the ray,
                        // exp(-sigmaT * segment) of the current homogeneous medium, medium switch at boundaries
                        const bool invisible = hprim >= 0;
                        const float seg = invisible ? tmax : remain;
                        bool again = false;
                        if (invisible && a.sc.shade[hprim].matIdx != -1) tr = mk3(0, 0, 0);
                        else {
                            if (medium >= 0) {
                                const f3 c = ld3(a.sc.mediums[medium].sigmaT) * (-seg);        // Homogeneous::Tr, src/medium.h:14
                                tr *= mk3(expf(c.x), expf(c.y), expf(c.z));
                            }
                            if (invisible) {
                                const WShade& s = a.sc.shade[hprim];
                                f3 nor;
                                if (s.type == 0) nor = normalize(lin3_seq(1.f - hb1 - hb2, ld3(s.n1), hb1, ld3(s.n2), hb2, ld3(s.n3)));
                                else nor = normalize((o + seg * d) - ld3(s.n1));
                                medium = dot(d, nor) > 0 ? s.mediumOutside : s.mediumInside;
                                remain -= seg;
                                o = o + seg * d;                                               // Ray(ray(ray.tmax), ray.d, m, eps, tmax)
                                tmax = remain;
                                hprim = -1; sp = 0; leaf = kDone;
                                float tn;
                                const bool in = slab(a.sc.root_min[0], a.sc.root_min[1], a.sc.root_min[2], a.sc.root_max[0], a.sc.root_max[1], a.sc.root_max[2], o, inv, tmax, tn);
                                cur = !in ? kDone : (a.sc.root_leaf_count > 0 ? ~0 : 0);
                                again = true; active = true;
                                ++nrays;
                            }
                        }
                        if (!again) a.pool.vis[slot] = make_float4(tr.x, tr.y, tr.z, 0.f);
                    }
                }
            }
            busy = __ballot_sync(kFullMask, active);
        } while (busy != 0u && (exhausted || __popc(busy) >= a.refill_below));
    }
    // ray statistics: one atomic per warp
#ifndef B200PT_EMULATE
    for (int off = 16; off > 0; off >>= 1) nrays += __shfl_down_sync(kFullMask, nrays, off);
#endif
    if (lane == 0u && nrays) atomicAdd(&a.counters->rays, (unsigned long long)nrays);
}

#undef STK_PUSH
#undef STK_POP

// ---- ray sorting for the tree kernel ------------------------------------------------------------------------------------------
// After the first bounce the 32 rays a warp pulls from the queue have nothing in common: they start anywhere in the scene
// and point anywhere, every lane walks its own part of the tree, and the node loop runs at a third of the SIMT width (C4:
// 11.7 of 32 lanes).  Before k_trace, the step's queue is therefore counting-sorted by
//     key = any-hit query | Morton code of the origin's cell (16^3 cells over the root box) | octant of the direction,
// so that the rays of a warp enter the tree through the same top nodes and tend to stay together.  Three small kernels on
// the lane's stream (histogram, one-CTA scan, scatter); the queue itself is not touched — k_trace reads the sorted copy.
// Pure scheduling: which lane traces a ray changes, nothing about the ray does.
// MEASURED AND LEFT OFF (B200PT_SORT_RAYS=1 switches it on; profiles/r03c_sort_rays.txt): the node loop of C4 goes from 11.7
// to 12.3 of 32 lanes only — rays that start in the same cell and octant still part ways within a few levels of a
// 20-level tree over a triangle soup, the divergence is in trip counts, not in entry points — and the three extra
// kernels cost more than that returns: C4 126 -> 106 Msamples/s, C3 460 -> 220 (its steps are 0.3 ms long), hair 464 -> 240.
constexpr int kRaySortBins = 1 << 16;
struct RaySortArgs {
    SceneDev sc; Pool pool; RayQueue q; uint32_t parity; int32_t sec_tmax;
    uint16_t* keys; uint32_t* hist; uint32_t* offs; uint32_t* sorted;
};
__device__ __forceinline__ uint32_t morton3_4bit(uint32_t x, uint32_t y, uint32_t z) {
    uint32_t m = 0u;
    for (int b = 0; b < 4; ++b) m |= (((x >> b) & 1u) << (3 * b)) | (((y >> b) & 1u) << (3 * b + 1)) | (((z >> b) & 1u) << (3 * b + 2));
    return m;
}
__global__ void __launch_bounds__(256) k_ray_hist(const RaySortArgs a) {
    const uint32_t tail = a.q.ctl->tail[a.parity & 1u];
    const float sx = 16.f / fmaxf(a.sc.root_max[0] - a.sc.root_min[0], 1e-30f), sy = 16.f / fmaxf(a.sc.root_max[1] - a.sc.root_min[1], 1e-30f),
                sz = 16.f / fmaxf(a.sc.root_max[2] - a.sc.root_min[2], 1e-30f);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < tail; i += gridDim.x * blockDim.x) {
        const uint32_t entry = a.q.entries[i];
        const uint32_t slot = entry & kSlotMask, kind = entry >> kKindShift;
        const float4 o = kind == 0u ? a.pool.o_rng[slot] : a.pool.pend_o[slot];
        const float4 d = kind == 0u ? a.pool.d_flags[slot] : (kind == 1u ? a.pool.shd[slot] : a.pool.misd[slot]);
        const float fx = fminf(fmaxf((o.x - a.sc.root_min[0]) * sx, 0.f), 15.f), fy = fminf(fmaxf((o.y - a.sc.root_min[1]) * sy, 0.f), 15.f),
                    fz = fminf(fmaxf((o.z - a.sc.root_min[2]) * sz, 0.f), 15.f);
        const int cx = (int)fx, cy = (int)fy, cz = (int)fz;           // (a NaN origin lands in cell 0: fmaxf drops it)
        const uint32_t oct = (d.x < 0.f ? 1u : 0u) | (d.y < 0.f ? 2u : 0u) | (d.z < 0.f ? 4u : 0u);
        const uint32_t key = (kind == 1u ? 0x8000u : 0u) | (morton3_4bit((uint32_t)cx, (uint32_t)cy, (uint32_t)cz) << 3) | oct;
        a.keys[i] = (uint16_t)key;
        atomicAdd(&a.hist[key], 1u);
    }
}
// one CTA: counts -> exclusive start offsets (into `offs`), counts cleared for the next step
__global__ void __launch_bounds__(1024) k_ray_scan(const RaySortArgs a) {
    constexpr int kPer = kRaySortBins / 1024;
    __shared__ uint32_t s_part[1024];
    const uint32_t t = threadIdx.x;
    uint32_t sum = 0u;
    for (int k = 0; k < kPer; ++k) sum += a.hist[t * kPer + k];
    s_part[t] = sum;
    __syncthreads();
#ifndef B200PT_EMULATE
    for (uint32_t off = 1; off < 1024u; off <<= 1) {            // Hillis-Steele inclusive scan of the 1024 partial sums
        const uint32_t v = t >= off ? s_part[t - off] : 0u;
        __syncthreads();
        s_part[t] += v;
        __syncthreads();
    }
    uint32_t run = s_part[t] - sum;
#else
    uint32_t run = 0u;
    for (uint32_t k = 0; k < t; ++k) run += s_part[k];
#endif
    for (int k = 0; k < kPer; ++k) { const uint32_t c = a.hist[t * kPer + k]; a.offs[t * kPer + k] = run; run += c; a.hist[t * kPer + k] = 0u; }
}
__global__ void __launch_bounds__(256) k_ray_scatter(const RaySortArgs a) {
    const uint32_t tail = a.q.ctl->tail[a.parity & 1u];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < tail; i += gridDim.x * blockDim.x)
        a.sorted[atomicAdd(&a.offs[a.keys[i]], 1u)] = a.q.entries[i];
}

// All group boxes of a small scene against one ray, warp-uniform: the hit groups as a 64-bit mask and the group the ray
// enters first.
__device__ __forceinline__ void small_ray_boxes(const TraceArgs& a, const float4* __restrict__ leaves, const f3 o, const f3 inv, const float tmax,
                                                unsigned long long& mask, int& best) {
    const int n_leaves = a.n_leaves;
    mask = 0ull;
    float tn, best_t = INFINITY;
    best = -1;
    if (slab(a.sc.root_min[0], a.sc.root_min[1], a.sc.root_min[2], a.sc.root_max[0], a.sc.root_max[1], a.sc.root_max[2], o, inv, tmax, tn)) {
        // BBox::Intersect (src/bbox.h:77-96) on every group box, with the min / max of each axis' two plane distances
        // replaced by a SELECTION: (plane - o) * (1/d) is monotonic in `plane`, so for a finite 1/d the smaller of
        // the two distances belongs to the plane the sign of d picks — bit-identical tmin / tmax, but the near and far
        // planes come from two 4-byte shared-memory loads at a per-lane offset and the 6 two-input min / max per box
        // are gone (the loop was 35 % of the kernel's instructions, profiles/r02d_wave_c2_lines.txt).  A direction
        // with a zero or denormal component (1/d infinite: the reference's test then leans on NaN-dropping fminf /
        // fmaxf) takes the literal form.
        // (A cheaper conservative test with one fused multiply-add per plane was tried and dropped: it needs grown
        // boxes, and a box that is hit where the reference's is not lets the exact primitive test accept a grazing hit
        // the reference culls — 1 sample in 12 288 of the 64 x 64 vol_caustic image, a direct view of the emitter.)
        const bool fast = fabsf(inv.x) < INFINITY && fabsf(inv.y) < INFINITY && fabsf(inv.z) < INFINITY;
        uint32_t m0 = 0u, m1 = 0u;
        if (fast) {
            const int nx = inv.x >= 0.f ? 0 : 3, ny = inv.y >= 0.f ? 1 : 4, nz = inv.z >= 0.f ? 2 : 5;      // word of the near plane
            const float* lw = reinterpret_cast<const float*>(leaves);
#define PT_BOX_TEST(l_, bit_, m_)                                                                                     \
            {                                                                                                     \
                const int l = (l_);                                                                               \
                const float* g_ = lw + 8 * l;                                                                     \
                const float tnear = fmaxf(fmaxf((g_[nx] - o.x) * inv.x, (g_[ny] - o.y) * inv.y), (g_[nz] - o.z) * inv.z);          \
                const float tfar = fminf(fminf((g_[3 - nx] - o.x) * inv.x, (g_[5 - ny] - o.y) * inv.y), (g_[7 - nz] - o.z) * inv.z); \
                if (!(tfar <= 0.00001f) && !(tnear > tfar) && !(tnear > tmax)) {                                  \
                    m_ |= 1u << (bit_);                                                                           \
                    if (tnear < best_t) { best_t = tnear; best = l; }                                             \
                }                                                                                                 \
            }
            const int n0 = n_leaves < 32 ? n_leaves : 32;
            for (int i = 0; i < n0; ++i) PT_BOX_TEST(i, i, m0)
            for (int i = 32; i < n_leaves; ++i) PT_BOX_TEST(i, i - 32, m1)
#undef PT_BOX_TEST
        } else {
            for (int l = 0; l < n_leaves; ++l) {
                const float4 q0 = leaves[2 * l], q1 = leaves[2 * l + 1];
                if (slab(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, o, inv, tmax, tn)) {
                    if (l < 32) m0 |= 1u << l; else m1 |= 1u << (l - 32);
                    if (tn < best_t) { best_t = tn; best = l; }
                }
            }
        }
        mask = (unsigned long long)m0 | ((unsigned long long)m1 << 32);
    }
}

// The ray a queue entry stands for: origin, direction, tmax (continuation / MIS rays are unbounded).
__device__ __forceinline__ void small_ray_fetch(const TraceArgs& a, const Pool& pool, const uint32_t entry, f3& o, f3& d, float& tmax) {
    const uint32_t slot = entry & kSlotMask, kind = entry >> kKindShift;
    const float4 orng = kind == 0u ? pool.o_rng[slot] : pool.pend_o[slot];             // shadow / MIS rays keep their own origin
    o = mk3(orng.x, orng.y, orng.z);
    float4 dv;
    if (kind == 0u) { dv = pool.d_flags[slot]; dv.w = INFINITY; }
    else if (kind == 1u) dv = pool.shd[slot];
    else { dv = pool.misd[slot]; if (!a.sec_tmax) dv.w = INFINITY; }
    d = mk3(dv.x, dv.y, dv.z);
    tmax = dv.w;
}

// One ray of the flat small-scene traversal (see k_trace_small below): fetch the ray of queue entry `entry` from the
// pool planes, test all group boxes, run the flat primitive loop, write the hit / visibility back.  `prims` / `leaves`
// point at the staged (shared-memory) copies; the pool planes are global (k_trace_small) or shared (k_wave.cuh).
template <bool VOL>
__device__ __forceinline__ void trace_small_ray(const TraceArgs& a, const Pool& pool, const WPrim* __restrict__ prims,
                                                const float4* __restrict__ leaves, const uint32_t entry, uint32_t& nrays,
                                                bool have_first = false, const unsigned long long first_mask = 0ull, const int first_best = -1) {
    const float eps = a.sc.eps;
    const uint32_t slot = entry & kSlotMask, kind = entry >> kKindShift;
    f3 o, d;
    float tmax;
    small_ray_fetch(a, pool, entry, o, d, tmax);
    const bool anyhit = !VOL && kind == 1u;
    f3 tr = mk3(1, 1, 1);
    float remain = tmax;
    int medium = -1;
    if (VOL && kind == 1u) medium = (int)((__float_as_uint(pool.d_flags[slot].w) >> kMedium2Shift) & 0xffu) - 1;
    const f3 inv = mk3(1.f / d.x, 1.f / d.y, 1.f / d.z);
    int hprim; float hb1 = 0.f, hb2 = 0.f;
    for (;;) {                                   // one pass per ray; vpt shadow rays repeat it per segment (Tr())
        ++nrays;
        hprim = -1;
        // ---- all group boxes, warp-uniform; remember the group the ray enters first (the CTA-local kernel has done this
        // pass for the ray's first leg already and sorted the rays by its result: k_wave.cuh)
        unsigned long long mask;
        int best;
        float tn;
        if (have_first) { mask = first_mask; best = first_best; have_first = false; }
        else small_ray_boxes(a, leaves, o, inv, tmax, mask, best);
        // ---- flat primitive loop over the hit groups: nearest group first, then the others in index order; once
        // a closest hit is known, a group is re-tested against the shrunken interval before its primitives are
        // (the `tmin > ray.tmax` rejection of BBox::Intersect, src/bbox.h:93)
        uint32_t w0 = 0u, w1 = 0u;        // remaining primitive indices of the current group, 16 bits each
        int left = 0;
        for (;;) {
            if (left == 0) {
                int l;
                if (best >= 0) { l = best; mask &= ~(1ull << best); best = -1; }
                else {
                    if (mask == 0ull) break;
                    l = __ffsll((long long)mask) - 1;
                    mask &= mask - 1ull;
                    if (hprim >= 0) {
                        const float4 q0 = leaves[2 * l], q1 = leaves[2 * l + 1];
                        if (!slab(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, o, inv, tmax, tn)) continue;
                    }
                }
                const float4 q1 = leaves[2 * l + 1];
                w0 = __float_as_uint(q1.z); w1 = __float_as_uint(q1.w);
                left = (w0 >> 16) == 0xffffu ? 1 : ((w1 & 0xffffu) == 0xffffu ? 2 : ((w1 >> 16) == 0xffffu ? 3 : 4));
            }
            const int pi = (int)(w0 & 0xffffu);
            w0 = (w0 >> 16) | (w1 << 16); w1 >>= 16;
            --left;
            const float4* pp = reinterpret_cast<const float4*>(prims + pi);
            const float4 p0 = pp[0], p1 = pp[1], p2 = pp[2];
            float t, b1, b2;
            const int acc = prim_test(p0, p1, p2, o, d, eps, tmax, t, b1, b2);
            if (acc) {
                if (anyhit) { hprim = pi; break; }
                if (acc == 2 || t < tmax || pi > hprim) { hprim = pi; hb1 = b1; hb2 = b2; }
                tmax = t;
            }
        }
        if (!(VOL && kind == 1u)) break;
        // Tr() (src/pathtracer.cu:298-322), same walk as in k_trace
        const bool invisible = hprim >= 0;
        const float seg = invisible ? tmax : remain;
        if (invisible && a.sc.shade[hprim].matIdx != -1) { tr = mk3(0, 0, 0); break; }
        if (medium >= 0) {
            const f3 c = ld3(a.sc.mediums[medium].sigmaT) * (-seg);
            tr *= mk3(expf(c.x), expf(c.y), expf(c.z));
        }
        if (!invisible) break;
        const WShade& sh = a.sc.shade[hprim];
        f3 nor;
        if (sh.type == 0) nor = normalize(lin3_seq(1.f - hb1 - hb2, ld3(sh.n1), hb1, ld3(sh.n2), hb2, ld3(sh.n3)));
        else nor = normalize((o + seg * d) - ld3(sh.n1));
        medium = dot(d, nor) > 0 ? sh.mediumOutside : sh.mediumInside;
        remain -= seg;
        o = o + seg * d;
        tmax = remain;
    }
    if (kind != 1u) {
        const float4 h = make_float4(hprim >= 0 ? tmax : -1.f, __int_as_float(hprim), hb1, hb2);
        if (kind == 0u) pool.hit0[slot] = h; else pool.hit1[slot] = h;
    } else if (!VOL) {
        const float v = hprim >= 0 ? 0.f : 1.f;
        pool.vis[slot] = make_float4(v, v, v, 0.f);
    } else {
        pool.vis[slot] = make_float4(tr.x, tr.y, tr.z, 0.f);
    }
}

// ---- small scenes (<= 256 primitives): flat list of tight primitive groups instead of a tree walk ----------
// A Cornell-box-sized scene has a few dozen primitives.  Walking its tree costs more in divergence (every lane is
// at a different depth) than the tree saves, so this kernel tests ALL group boxes in one warp-uniform loop (box
// records come out of shared memory as broadcasts), keeps the hit groups of a ray as a 64-bit mask, and then runs
// one flat primitive loop in which every iteration is one Moeller-Trumbore / sphere test for every lane that still
// has a primitive.  Groups (<= 64 per scene, <= 4 primitives each) are built at upload by greedy agglomeration
// under a box-test/primitive-test cost model, so a quad's two triangles share one tight box while unrelated
// triangles are not lumped together as the SAH leaves do.  Same slab and primitive arithmetic as the tree kernel;
// a primitive is tested iff its group's box is hit, which (up to rounding at box faces) is a superset of the
// primitives that can be hit, so the closest / any hit is the same.
//   group record (32 B): q0 = bmin.xyz, bmax.x   q1 = bmax.yz, prims 0|1 (2 x u16 bits), prims 2|3 (0xffff = none)
template <bool VOL>
__global__ void __launch_bounds__(kTraceThreads) k_trace_small(const TraceArgs a) {
    const WPrim* __restrict__ prims = a.sc.prims;
    const float4* __restrict__ leaves = a.leaves;
#ifndef B200PT_EMULATE
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    {
        const uint32_t lb = (uint32_t)a.n_leaves * 32u;
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, a.small_prim_bytes + lb);
            tma_bulk_g2s(smem_raw, a.sc.prims, a.small_prim_bytes, &bar);
            tma_bulk_g2s(smem_raw + a.small_prim_bytes, a.leaves, lb, &bar);
        }
        mbar_wait(&bar, 0);
        prims = reinterpret_cast<const WPrim*>(smem_raw);
        leaves = reinterpret_cast<const float4*>(smem_raw + a.small_prim_bytes);
    }
#endif
    const uint32_t lane = pt_lane();
    const uint32_t par = a.parity & 1u;
    const uint32_t tail = a.q.ctl->tail[par];
    if (blockIdx.x == 0 && threadIdx.x == 0) { a.q.ctl->tail[par ^ 1u] = 0u; a.q.ctl->head[par ^ 1u] = 0u; }
    uint32_t nrays = 0;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < tail; idx += stride) {
        trace_small_ray<VOL>(a, a.pool, prims, leaves, a.q.entries[idx], nrays);
    }
#ifndef B200PT_EMULATE
    for (int off = 16; off > 0; off >>= 1) nrays += __shfl_down_sync(kFullMask, nrays, off);
#endif
    if (lane == 0u && nrays) atomicAdd(&a.counters->rays, (unsigned long long)nrays);
}

}  // namespace pt
