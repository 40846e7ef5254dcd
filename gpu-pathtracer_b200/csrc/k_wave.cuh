// k_wave.cuh — CTA-LOCAL wavefront for scenes whose acceleration structure fits in shared memory (<= 256 primitives:
// the Cornell-box family, BASELINE configs C1 / C2 / C5).
//
// The global wavefront (k_shade + k_trace over a path pool in HBM) streams ~1.8 KB of path state per sample through
// HBM for a scene of 3 KB, and pays two kernel launches per bounce.  Here a persistent CTA owns kWaveThreads path
// SLOTS whose whole state (the same 14 / 15 float4 planes) lives in SHARED memory next to the TMA-staged scene, and
// runs the SAME two stages as phases of one kernel:
//
//   shade phase   thread t runs shade_slot() on slot t (finish the previous bounce's direct light, shade the hit,
//                 retire + regenerate from the batch's sample counter) and appends the slot's <= 3 rays to the
//                 CTA's ray queue in shared memory (one warp-aggregated shared-memory atomic per warp);
//   trace phase   the CTA's threads take the queued rays in order — a COMPACT list, so every lane of every warp but
//                 the last carries a ray whatever mix of live / dead / shadow-less slots produced it — and run
//                 trace_small_ray() (flat box list + Moeller-Trumbore, k_trace.cuh) with ray and hit records in
//                 shared memory.
//
// Nothing but the finished samples (16 B each) and three counters ever goes to global memory; one launch renders a
// whole spp batch.  Paths still move wavefront-fashion (all slots shade, then all rays trace), which keeps the SIMT
// width the megakernel of the reference loses to path-length divergence.
#pragma once
#include "k_shade.cuh"
#include "k_trace.cuh"
#include "k_het.cuh"
#ifdef B200PT_EMULATE
#include <vector>
#endif

namespace pt {

#ifndef PT_WAVE_THREADS
#define PT_WAVE_THREADS 256                       // 128: C2 -6 %, six-BSDF scene -43 % (smaller sort domain); 64: -20 % (profiles/r02p_cta_size.txt)
#endif
constexpr int kWaveThreads = PT_WAVE_THREADS;     // slots (= threads) per CTA of the surface / homogeneous-medium wavefront
#ifndef PT_WAVE_HET_THREADS
#define PT_WAVE_HET_THREADS 320                   // heterogeneous media: 2 CTAs x 10 warps at 96 registers (~200 B of spills) against 2 x 8 at 123:
#endif                                            // more slots per sort, more warps per SM — smoke 175 -> 203 Msamples/s, shipped scene 131 -> 137 at
                                                  // 1024^2 (288: 205 / 130, 352: 197 / 129, 384: 177 / 128, 128: 113; profiles/r06_het_cta.txt)
#ifndef PT_WAVE_ALLMATS_THREADS
#define PT_WAVE_ALLMATS_THREADS 288               // all-BSDF instantiations: the larger sort domain of the material binning is worth more than
#endif                                            // the registers (27 warps per SM, 72 registers): six-BSDF scene +10 % `pt`, +12 % `vpt`;
                                                  // C2 +0.5 %, C5 -5 % stay at 256 (profiles/r03b_t288.txt)
#ifndef PT_WAVE_LAMBERT_THREADS
#define PT_WAVE_LAMBERT_THREADS 448               // lambertian `pt` (C1 / C2): 2 CTAs x 14 warps instead of 3 x 8 — what shared memory allows once the
#endif                                            // sort arrays of the two-pass trace phase are gone: C2 1513 -> 1552 Msamples/s, C1 1773 -> 1846
                                                  // (2 x 384: +1 %, 2 x 416: -2.5 %, 3 x 288: +0.5 %; profiles/r07_lambert_cta.txt)
template <bool HET, uint32_t MATS = 0u, bool VOL = true> struct WaveThreads {
    static constexpr int value = HET ? PT_WAVE_HET_THREADS
                                     : (MATS == 0x3fu ? PT_WAVE_ALLMATS_THREADS : ((MATS == 1u && !VOL) ? PT_WAVE_LAMBERT_THREADS : kWaveThreads));
};
#ifndef PT_WAVE_LAMBERT_CTAS
#define PT_WAVE_LAMBERT_CTAS 2      // resident CTAs per SM the lambertian `pt` instantiation is compiled for
#endif
#ifndef PT_WAVE_MATS_CTAS
#define PT_WAVE_MATS_CTAS 3         // resident CTAs per SM the `vpt` / several-BSDF instantiations are compiled for: 80 registers and
                                    // ~100 B of spills against 96-99 at 2 — C5 662 -> 706 Msamples/s, six-BSDF scenes unchanged (profiles/r02s_ctas3.txt)
#endif
#ifndef PT_WAVE_HET_CTAS
#define PT_WAVE_HET_CTAS 2          // resident CTAs per SM the heterogeneous-media instantiation is compiled for (3: 80 registers,
                                    // 530 B of spills — smoke 141 vs 156, shipped scene 94 vs 125: profiles/r02s_ctas3.txt)
#endif

// The shade / trace bodies read scene, camera, shard map and batch from the same argument structs as the global
// wavefront (kernel parameters: constant bank); their pool / queue members are unused here — the CTA's own planes and
// queue in shared memory are handed to the bodies explicitly.
struct WaveArgs {
    ShadeArgs sa;
    TraceArgs ta;
};

// Shared-memory footprint of one CTA: scene (primitive records + group boxes) + path planes + ray queue.
template <bool VOL> struct WavePlanes { static constexpr int value = VOL ? 15 : 14; };
// ... + the per-ray box results of the sorted trace phase: hit-group mask (8 B), order (2 B) and nearest group (1 B) per queue entry
// PT_WAVE_SORT = 1: the trace phase of the CTA-local wavefront runs in two passes — the box pass of every queued ray (the
// same work for every ray: full SIMT width), a counting sort of the rays by (kind of query, number of hit groups) in shared
// memory, and the primitive pass in that order: every ray needs a different number of Moeller-Trumbore tests, and in queue
// order the primitive loop ran at 5-10 of 32 lanes (profiles/r02d_wave_c2_lines.txt).  It was worth +3.4 % on C2 while a
// sample cost 10.2 rays; with the pruned ray set (6.4 rays per sample, wavefront.cuh) the three extra barriers per step cost
// more than the fuller warps return — single pass: C2 1422 -> 1508 Msamples/s, C1 1651 -> 1768 (profiles/r02z_sort_again.txt).
#ifndef PT_WAVE_SORT
#define PT_WAVE_SORT 0
#endif
template <bool VOL> inline size_t wave_smem_bytes(uint32_t prim_bytes, int n_leaves, int threads) {
    return (size_t)((prim_bytes + (uint32_t)n_leaves * 32u + 127u) & ~127u) + (size_t)WavePlanes<VOL>::value * threads * sizeof(float4) +
           (size_t)3 * threads * sizeof(uint32_t) +
           ((PT_WAVE_SORT != 0 && !VOL) ? (size_t)3 * threads * (sizeof(unsigned long long) + sizeof(uint16_t) + 2) : (size_t)0);   // sort arrays of the two-pass trace phase
}

#ifndef B200PT_EMULATE
#define PT_WAVE_FOR_THREADS(t) for (uint32_t t = threadIdx.x, once_ = 1u; once_; once_ = 0u)
#define PT_WAVE_SYNC() __syncthreads()
#else
// CPU emulation of the test suite: one host thread plays a whole CTA, phase by phase
#define PT_WAVE_FOR_THREADS(t) for (uint32_t t = 0; t < (uint32_t)kT; ++t)
#define PT_WAVE_SYNC() do { } while (0)
#endif

template <bool VOL> __device__ __forceinline__ uint32_t wave_sort_key(uint32_t entry, unsigned long long mask) {
    const uint32_t kind = entry >> kKindShift;
    const uint32_t n = (uint32_t)__popc((uint32_t)mask) + (uint32_t)__popc((uint32_t)(mask >> 32));
    const uint32_t any = (!VOL && kind == 1u) ? 16u : 0u;             // any-hit queries after the closest-hit ones
    return any + 15u - (n < 15u ? n : 15u);                           // most hit groups first
}

// Sort key of a heterogeneous-media slot: slots whose resumed stage starts with a walk through a heterogeneous medium
// (het_slot's tracking prelude — one loop for all of them) come first, ordered by wait state; then the others by wait
// state; dead slots last.
__device__ __forceinline__ uint32_t het_sort_key(const Pool& P, const SceneDev& sc, uint32_t slot) {
    const uint32_t f = __float_as_uint(P.d_flags[slot].w);
    if (!(f & H_ALIVE)) return 15u;
    const uint32_t st = (f >> kHStateShift) & 3u;
    if (st == HS_MIS) return 6u;
    return st + (het_slot_tracks(sc, P, slot) ? 0u : 3u);
}

// Sort key of a slot for the shade phase of a scene with several BSDFs (material binning, see k_shade): dead slots, misses,
// then the material class of the hit primitive; `vpt` paths inside a medium (free-flight sampling, phase function) form
// a second set of classes.
template <bool VOL>
__device__ __forceinline__ uint32_t wave_shade_key(const Pool& P, const SceneDev& sc, uint32_t slot) {
    const uint32_t f = __float_as_uint(P.d_flags[slot].w);
    if (!(f & F_ALIVE)) return 0u;
    const float4 h0 = P.hit0[slot];
    const uint32_t prim = __float_as_uint(h0.y);
    uint32_t key = (h0.x < 0.f || prim >= (uint32_t)sc.n_prims) ? 1u : 2u + (sc.prim_key[prim] & 15u);
    if (VOL && ((f >> kMediumShift) & 0xffu) != 0u) key += 18u;
    return key;
}
constexpr uint32_t kWaveShadeKeys = 36u;

// HET: the slots run the heterogeneous-media coroutine (k_het.cuh) instead of the surface / homogeneous shade stage.
template <bool VOL, uint32_t MATS, bool HET>
__global__ void __launch_bounds__(WaveThreads<HET, MATS, VOL>::value, HET ? PT_WAVE_HET_CTAS : ((VOL || MATS != kMatsLambertOnly) ? PT_WAVE_MATS_CTAS : PT_WAVE_LAMBERT_CTAS) * (kWaveThreads <= 128 ? 256 / kWaveThreads : 1)) k_wave_small(const WaveArgs a) {
    constexpr int kT = WaveThreads<HET, MATS, VOL>::value;
    const ShadeArgs& sa = a.sa;
    const TraceArgs& ta = a.ta;
#ifndef B200PT_EMULATE
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
#else
    static thread_local std::vector<unsigned char> smem_vec;
    smem_vec.assign(wave_smem_bytes<VOL>(ta.small_prim_bytes, ta.n_leaves, kT) + 128, 0);
    unsigned char* smem_raw = smem_vec.data();
#endif
    __shared__ uint32_t s_busy[2], s_retired, s_rays, s_hist[32];
    __shared__ QueueCtl s_ctl;
    __shared__ unsigned long long s_next;

    const uint32_t scene_bytes = (ta.small_prim_bytes + (uint32_t)ta.n_leaves * 32u + 127u) & ~127u;
    float4* const planes = reinterpret_cast<float4*>(smem_raw + scene_bytes);
    uint32_t* const s_queue = reinterpret_cast<uint32_t*>(planes + (size_t)WavePlanes<VOL>::value * kT);
    unsigned long long* const s_mask = reinterpret_cast<unsigned long long*>(s_queue + 3 * kT);
    uint16_t* const s_sorted = reinterpret_cast<uint16_t*>(s_mask + 3 * kT);
    signed char* const s_best = reinterpret_cast<signed char*>(s_sorted + 3 * kT);
    const WPrim* prims = ta.sc.prims;
    const float4* leaves = ta.leaves;

    // ---- stage the scene with TMA, clear the slots
#ifndef B200PT_EMULATE
    {
        const uint32_t lb = (uint32_t)ta.n_leaves * 32u;
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            s_ctl.tail[0] = s_ctl.tail[1] = 0u; s_ctl.head[0] = s_ctl.head[1] = 0u;
            s_busy[0] = s_busy[1] = 0u; s_retired = 0u; s_rays = 0u;
            s_next = sa.counters->next_sample;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, ta.small_prim_bytes + lb);
            tma_bulk_g2s(smem_raw, ta.sc.prims, ta.small_prim_bytes, &bar);
            tma_bulk_g2s(smem_raw + ta.small_prim_bytes, ta.leaves, lb, &bar);
        }
        // all slots start dead with no static sample consumed (the flags word and li_t.w are what shade_slot looks at)
        for (int k = 0; k < WavePlanes<VOL>::value; ++k) planes[(size_t)k * kT + threadIdx.x] = make_float4(0.f, 0.f, 0.f, 0.f);
        mbar_wait(&bar, 0);
        prims = reinterpret_cast<const WPrim*>(smem_raw);
        leaves = reinterpret_cast<const float4*>(smem_raw + ta.small_prim_bytes);
        __syncthreads();
    }
#else
    s_ctl.tail[0] = s_ctl.tail[1] = 0u; s_ctl.head[0] = s_ctl.head[1] = 0u;
    s_busy[0] = s_busy[1] = 0u; s_retired = 0u; s_rays = 0u;
    s_next = sa.counters->next_sample;
#endif

    Pool P;
    P.o_rng = planes; P.d_flags = planes + 1 * kT; P.beta_s = planes + 2 * kT; P.li_t = planes + 3 * kT;
    P.shd = planes + 4 * kT; P.misd = planes + 5 * kT; P.ldl = planes + 6 * kT; P.misf = planes + 7 * kT;
    P.beta_old = planes + 8 * kT; P.hit0 = planes + 9 * kT; P.hit1 = planes + 10 * kT; P.vis = planes + 11 * kT;
    P.pend_o = planes + 12 * kT; P.carry = planes + 13 * kT; P.aux = planes + (VOL ? 14 : 0) * kT;
    P.n = kT;
    RayQueue Q; Q.entries = s_queue; Q.ctl = &s_ctl;
    const uint32_t pool_n = gridDim.x * (uint32_t)kT;

    if (HET) {
        // Heterogeneous media: the slots are coroutines in one of four wait states (k_het.cuh), and each state resumes into
        // different code (free-flight sampling, a leg of a transmittance walk, the MIS hit).  Executed slot-per-thread,
        // a warp ran every state's code with a fifth of its lanes (6.2 of 32 active lanes per instruction,
        // profiles/r02_counters.json, "smoke").  So every step starts with a COUNTING SORT of the CTA's slots by state
        // (shared memory; key = state x "the resumed stage walks a medium") and thread j resumes slot order[j]: the lanes of a warp resume the same stage and walk
        // their tracking loops together.  Then the usual two phases: glue posts <= 1 query per slot into the compact
        // queue, the trace phase runs the queue.
        __shared__ uint32_t s_cnt[16], s_pos[16];
        uint16_t* const s_order = reinterpret_cast<uint16_t*>(s_queue + 2 * kT);       // the queue holds <= 1 entry per slot here
        PT_WAVE_FOR_THREADS(t) { if (t < 16u) { s_cnt[t] = 0u; s_pos[t] = 0u; } }
        PT_WAVE_SYNC();
        for (uint32_t step = 0;; ++step) {
            const uint32_t par = step & 1u;
            // ---- sort the slots by wait state (dead slots last): count, then scatter (every thread sums the counts below
            // its key itself — 16 broadcast reads instead of a prefix phase and its barrier)
            PT_WAVE_FOR_THREADS(t) { atomicAdd(&s_cnt[het_sort_key(P, sa.sc, t)], 1u); }
            PT_WAVE_SYNC();
            PT_WAVE_FOR_THREADS(t) {
                const uint32_t key = het_sort_key(P, sa.sc, t);
                uint32_t base = 0u;
                for (uint32_t k = 0; k < key; ++k) base += s_cnt[k];
                s_order[base + atomicAdd(&s_pos[key], 1u)] = (uint16_t)t;
            }
            PT_WAVE_SYNC();
            const uint32_t n_track = PT_HET_OVERLAP ? s_cnt[0] + s_cnt[1] + s_cnt[2] : 0u;       // slots waiting for a walk: first in s_order
            // ---- glue phase
            PT_WAVE_FOR_THREADS(t) {
#ifdef B200PT_EMULATE
                threadIdx.x = t;
#endif
                const uint32_t slot = s_order[t];
                uint32_t posted = 0u;
                het_slot<MATS, true, true>(sa, P, Q, par, &s_retired, &s_busy[par], slot, blockIdx.x * (uint32_t)kT + slot, pool_n, s_next, posted);
            }
            PT_WAVE_SYNC();
            const uint32_t tail = s_ctl.tail[par];
            if (tail == 0u && s_busy[par] == 0u) break;
            // ---- trace phase
            PT_WAVE_FOR_THREADS(t) {
#ifdef B200PT_EMULATE
                threadIdx.x = t;
#endif
                if (t == 0u) { s_ctl.tail[par ^ 1u] = 0u; s_busy[par ^ 1u] = 0u; s_next = *(volatile unsigned long long*)&sa.counters->next_sample; }
                if (t < 16u) { s_cnt[t] = 0u; s_pos[t] = 0u; }                      // re-arm the sort counters for the next step
                uint32_t nrays = 0;
                // the first n_track threads walk one chunk of a waiting slot's medium walk (full warps: those slots are
                // sorted in front), the others trace the queue — tracking slots post no ray, so tail <= kT - n_track
                if (t < n_track) het_track_step(sa.sc, P, s_order[t]);
                else for (uint32_t idx = t - n_track; idx < tail; idx += (uint32_t)kT - n_track) trace_small_ray<VOL>(ta, P, prims, leaves, s_queue[idx], nrays);
#ifndef B200PT_EMULATE
                for (int off = 16; off > 0; off >>= 1) nrays += __shfl_down_sync(kFullMask, nrays, off);
#endif
                if (pt_lane() == 0u && nrays) atomicAdd(&s_rays, nrays);
            }
            PT_WAVE_SYNC();
        }
    } else {
    // material binning (scenes with several BSDFs): thread j shades the j-th slot in material order — same counting sort
    __shared__ uint32_t s_mcnt[kWaveShadeKeys], s_mpos[kWaveShadeKeys];
    __shared__ uint16_t s_morder[kT];
    const bool bin = sa.bin_materials != 0;       // (forced on for a lambertian-only scene it costs 5 %: profiles/r02o_bin_lambert.txt)
    if (bin) {
        PT_WAVE_FOR_THREADS(t) { if (t < kWaveShadeKeys) { s_mcnt[t] = 0u; s_mpos[t] = 0u; } }
        PT_WAVE_SYNC();
    }
    for (uint32_t step = 0;; ++step) {
        const uint32_t par = step & 1u;
        if (bin) {
            PT_WAVE_FOR_THREADS(t) { atomicAdd(&s_mcnt[wave_shade_key<VOL>(P, sa.sc, t)], 1u); }
            PT_WAVE_SYNC();
            PT_WAVE_FOR_THREADS(t) {
                const uint32_t key = wave_shade_key<VOL>(P, sa.sc, t);
                uint32_t base = 0u;
                for (uint32_t k = 0; k < key; ++k) base += s_mcnt[k];
                s_morder[base + atomicAdd(&s_mpos[key], 1u)] = (uint16_t)t;
            }
            PT_WAVE_SYNC();
        }
        // ---- shade phase
        PT_WAVE_FOR_THREADS(t) {
#ifdef B200PT_EMULATE
            threadIdx.x = t;
#endif
            const uint32_t slot = bin ? (uint32_t)s_morder[t] : t;
            SlotRec r;
            load_slot<VOL>(P, slot, r);
            shade_slot<VOL, MATS, true>(sa, P, Q, par, &s_retired, &s_busy[par], slot, blockIdx.x * (uint32_t)kT + slot, pool_n, r, s_next);
        }
        PT_WAVE_SYNC();
        const uint32_t tail = s_ctl.tail[par];
        if (tail == 0u && s_busy[par] == 0u) break;          // no ray in flight, no slot alive, no sample left
        // ---- trace phase (thread 0 also refreshes the sample-counter snapshot and re-arms the other control set)
        PT_WAVE_FOR_THREADS(t) {
            if (t == 0u) { s_ctl.tail[par ^ 1u] = 0u; s_busy[par ^ 1u] = 0u; s_next = *(volatile unsigned long long*)&sa.counters->next_sample; }
            if (t < 32u) s_hist[t] = 0u;
            if (bin && t < kWaveShadeKeys) { s_mcnt[t] = 0u; s_mpos[t] = 0u; }
        }
        // (two-pass trace phase, compiled out by default — see PT_WAVE_SORT above; never used for `vpt`, whose shadow queries are
        // multi-leg transmittance walks: -5 % on C5, profiles/r02i_sort_ab.txt)
        constexpr bool kSort = (PT_WAVE_SORT != 0) && !VOL;
        if (kSort) {
            PT_WAVE_SYNC();
            // box pass: every ray against all group boxes (uniform work), result kept per queue entry; histogram of the sort key
            PT_WAVE_FOR_THREADS(t) {
                for (uint32_t idx = t; idx < tail; idx += (uint32_t)kT) {
                    const uint32_t entry = s_queue[idx];
                    f3 o, d; float tmax;
                    small_ray_fetch(ta, P, entry, o, d, tmax);
                    unsigned long long mask; int best;
                    small_ray_boxes(ta, leaves, o, mk3(1.f / d.x, 1.f / d.y, 1.f / d.z), tmax, mask, best);
                    s_mask[idx] = mask; s_best[idx] = (signed char)best;
                    atomicAdd(&s_hist[wave_sort_key<VOL>(entry, mask)], 1u);
                }
            }
            PT_WAVE_SYNC();
            PT_WAVE_FOR_THREADS(t) {
                if (t == 0u) { uint32_t run = 0u; for (int k = 0; k < 32; ++k) { const uint32_t c = s_hist[k]; s_hist[k] = run; run += c; } }
            }
            PT_WAVE_SYNC();
            PT_WAVE_FOR_THREADS(t) {
                for (uint32_t idx = t; idx < tail; idx += (uint32_t)kT)
                    s_sorted[atomicAdd(&s_hist[wave_sort_key<VOL>(s_queue[idx], s_mask[idx])], 1u)] = (uint16_t)idx;
            }
            PT_WAVE_SYNC();
        }
        // primitive pass — in sorted order the lanes of a warp carry rays of one kind with (nearly) the same number of hit groups
        PT_WAVE_FOR_THREADS(t) {
#ifdef B200PT_EMULATE
            threadIdx.x = t;
#endif
            uint32_t nrays = 0;
            if (kSort) {
                for (uint32_t j = t; j < tail; j += (uint32_t)kT) {
                    const uint32_t idx = s_sorted[j];
                    trace_small_ray<VOL>(ta, P, prims, leaves, s_queue[idx], nrays, true, s_mask[idx], (int)s_best[idx]);
                }
            } else {
                for (uint32_t idx = t; idx < tail; idx += (uint32_t)kT) trace_small_ray<VOL>(ta, P, prims, leaves, s_queue[idx], nrays);
            }
#ifndef B200PT_EMULATE
            for (int off = 16; off > 0; off >>= 1) nrays += __shfl_down_sync(kFullMask, nrays, off);
#endif
            if (pt_lane() == 0u && nrays) atomicAdd(&s_rays, nrays);
        }
        PT_WAVE_SYNC();
    }
    }
    // ---- per-CTA statistics: one atomic each
    PT_WAVE_FOR_THREADS(t) {
        if (t == 0u) {
            if (s_retired) atomicAdd(&sa.counters->done_samples, (unsigned long long)s_retired);
            if (s_rays) atomicAdd(&sa.counters->rays, (unsigned long long)s_rays);
        }
    }
}

}  // namespace pt
