// k_het.cuh — the shade stage of the wavefront for `vpt` scenes with HETEROGENEOUS media (what the
// reference's shipped scenes/cornell_box/scene.json renders; Volpath, src/pathtracer.cu:1025-1242 with
// Heterogeneous::Sample / Tr, src/medium.h:52-179).
//
// Delta / ratio tracking draws a DATA-DEPENDENT number of random numbers from the path's one RNG stream in the middle of
// a bounce — the transmittance walk Tr() of the shadow ray sits between the light sample and the BSDF samples, the
// transmittance towards a BSDF-sampled emitter between those and the continuation sample — and every leg of such a walk
// needs a closest-hit query first.  k_shade's fixed "draw the bounce, then trace its three rays" schedule cannot replay
// that, so here a path is a COROUTINE that is cut at every closest-hit query:
//
//      MAIN --hit--> [free-flight sampling]  --medium scatter--> light sample --> WALK(med) ... --> phase sample --> MAIN
//                                            --surface-->  emitter | boundary (--> MAIN) |
//                                                          light sample --> WALK(surf) ... --> MIS sample --> MIS --> continuation --> MAIN
//
// One call of het_slot() resumes the slot's coroutine with the hit its last query returned, runs the glue (tracking
// loops included — they only touch the density grid) up to the NEXT query, and posts that query as ONE ray in the
// ordinary ray queue: kind 0 = the path's own ray (origin / direction planes, result -> hit0), kind 2 = a secondary
// closest-hit query with its own tmax (a Tr() leg or the MIS ray: pend_o / misd planes, result -> hit1).  The
// traversal kernels are the wavefront's own (k_trace / k_trace_small bodies): every lane of a traversal warp carries a
// ray, instead of one thread walking its private stack at 4 of 32 active lanes (the sequential kernel this replaces,
// and the reference's megakernel).  A path that ends is regenerated in the same call.
#pragma once
#include "k_shade.cuh"

namespace pt {

// ---- heterogeneous-medium arithmetic (leaf math: same float expressions as src/medium.h) -------------------------
// Heterogeneous::d / getDensity (src/medium.h:159-178): trilinear lookup, 0 outside the grid.  The reference calls
// d(p + corner) eight times, each converting its own float coordinates to int and checking all six bounds; psi is
// integral, so (int)(psi + 0) == (int)psi and the eight lookups share two conversions and two bound checks per axis —
// every conversion the reference makes is still made on the same float, a third of the instructions.
__device__ __forceinline__ float lerpf(float a, float b, float t) { return a + t * (b - a); }       // src/cutil_math.h:1008
__device__ __forceinline__ float het_density(const WHetero& H, f3 p) {
    f3 ps = mk3(p.x * H.nx, p.y * H.ny, p.z * H.nz);
    f3 psi = mk3(floorf(ps.x), floorf(ps.y), floorf(ps.z));
    f3 delta = ps - psi;
    const int x0 = (int)psi.x, x1 = (int)(psi.x + 1.f);
    const int y0 = (int)psi.y, y1 = (int)(psi.y + 1.f);
    const int z0 = (int)psi.z, z1 = (int)(psi.z + 1.f);
    const bool bx0 = !(x0 < 0 || x0 > H.nx - 1), bx1 = !(x1 < 0 || x1 > H.nx - 1);
    const bool by0 = !(y0 < 0 || y0 > H.ny - 1), by1 = !(y1 < 0 || y1 > H.ny - 1);
    const bool bz0 = !(z0 < 0 || z0 > H.nz - 1), bz1 = !(z1 < 0 || z1 > H.nz - 1);
    const float* D = H.density;
    const int r00 = z0 * H.ny * H.nx + y0 * H.nx, r10 = z0 * H.ny * H.nx + y1 * H.nx;
    const int r01 = z1 * H.ny * H.nx + y0 * H.nx, r11 = z1 * H.ny * H.nx + y1 * H.nx;
    const float d000 = (bx0 && by0 && bz0) ? D[r00 + x0] : 0.f, d100 = (bx1 && by0 && bz0) ? D[r00 + x1] : 0.f;
    const float d010 = (bx0 && by1 && bz0) ? D[r10 + x0] : 0.f, d110 = (bx1 && by1 && bz0) ? D[r10 + x1] : 0.f;
    const float d001 = (bx0 && by0 && bz1) ? D[r01 + x0] : 0.f, d101 = (bx1 && by0 && bz1) ? D[r01 + x1] : 0.f;
    const float d011 = (bx0 && by1 && bz1) ? D[r11 + x0] : 0.f, d111 = (bx1 && by1 && bz1) ? D[r11 + x1] : 0.f;
    float d00 = lerpf(d000, d100, delta.x);
    float d10 = lerpf(d010, d110, delta.x);
    float d01 = lerpf(d001, d101, delta.x);
    float d11 = lerpf(d011, d111, delta.x);
    float d0 = lerpf(d00, d10, delta.y);
    float d1 = lerpf(d01, d11, delta.y);
    return lerpf(d0, d1, delta.z);
}
// Heterogeneous::Tr (src/medium.h:64-135): delta (0) / ratio (1) / residual-ratio (2) tracking over [0, tmax]
__device__ __noinline__ f3 het_tr(const WMedium& M, const WHetero& H, f3 o, f3 dir, float tmax, uint32_t& rng) {
    float sigma = dot(ld3(M.sigmaT), mk3(0.212671f, 0.715160f, 0.072169f));
    const f3 p0 = ld3(H.p0);
    f3 d = ld3(H.p1) - p0;
    float tr = 1.f;
    float dist = 0.f;
    int iter = H.iterMax;
    if (H.evalTransmittanceType == 0) {
        while (true) {
            dist += -logf(rng_next(rng)) * H.invMaxDensity / sigma;
            if (dist >= tmax) break;
            f3 p = o + dir * dist;
            p = (p - p0) / d;
            if (het_density(H, p) * H.invMaxDensity > rng_next(rng)) { tr = 0; break; }
            if (--iter == 0) { tr = 0; break; }
        }
    } else if (H.evalTransmittanceType == 1) {
        while (true) {
            dist += -logf(rng_next(rng)) * H.invMaxDensity / sigma;
            if (dist >= tmax) break;
            f3 p = o + dir * dist;
            p = (p - p0) / d;
            tr *= 1.f - het_density(H, p) * H.invMaxDensity;
            if (tr < 0.1f) {
                float q = 1.f - tr;
                if (rng_next(rng) < q) return mk3(0.f, 0.f, 0.f);
                tr = 1;
            }
            if (--iter == 0) break;
        }
    } else {
        float maxDensity = 1 / H.invMaxDensity;
        float ce = 0.5f * maxDensity;                    // (float)(0.5 * (double)maxDensity): exact either way
        float tc = expf(-tmax * ce * sigma);
        while (true) {
            dist += -logf(rng_next(rng)) * (1 / (maxDensity - ce) / sigma);
            if (dist >= tmax) break;
            f3 p = o + dir * dist;
            p = (p - p0) / d;
            tr *= 1.f - (het_density(H, p) - ce) / (maxDensity - ce);
            if (tr < 0.1f) {
                float q = 1.f - tr;
                if (rng_next(rng) < q) return mk3(0.f, 0.f, 0.f);
                tr /= (1.f - q);
            }
            if (--iter == 0) break;
        }
        tr *= tc;
    }
    return mk3(tr, tr, tr);
}
// the type dispatch of every Tr call site (src/pathtracer.cu:308-311, :1107-1110, :1180-1183, :1200-1203)
__device__ __forceinline__ f3 medium_tr(const SceneDev& sc, int medium, f3 o, f3 d, float tmax, uint32_t& rng) {
    const WMedium& M = sc.mediums[medium];
    if (M.type == 0) return exp3(ld3(M.sigmaT) * (-tmax));                                      // Homogeneous::Tr, src/medium.h:14
    return het_tr(M, sc.het[medium], o, d, tmax, rng);
}
// Homogeneous::Sample (src/medium.h:19-49) / Heterogeneous::Sample (:137-157): free-flight distance along (o, dir) up to
// the surface at tmax; returns the throughput factor
__device__ __noinline__ f3 medium_sample(const SceneDev& sc, int medium, f3 o, f3 dir, float tmax, uint32_t& rng, float& t, bool& sampled) {
    const WMedium& M = sc.mediums[medium];
    f3 sigmaT = ld3(M.sigmaT), sigmaS = ld3(M.sigmaS);
    float sigma = dot(sigmaT, mk3(0.212671f, 0.715160f, 0.072169f));
    if (M.type == 0) {
        float dist = -logf(rng_next(rng)) / sigma;
        f3 Tr = exp3(sigmaT * -dist);
        float pdf = sigma * expf(sigma * -dist);
        sampled = dist < tmax;
        t = dist;
        return sampled ? (Tr * sigmaS / pdf) : sigmaT * Tr / pdf;
    }
    const WHetero& H = sc.het[medium];
    const f3 p0 = ld3(H.p0);
    f3 d = ld3(H.p1) - p0;
    float dist = 0.f;
    int iter = H.iterMax;
    while (true) {
        dist += -logf(rng_next(rng)) * H.invMaxDensity / sigma;
        if (dist >= tmax) break;
        f3 p = o + dir * dist;
        p = (p - p0) / d;
        if (het_density(H, p) * H.invMaxDensity > rng_next(rng)) {
            t = dist;
            sampled = true;
            return sigmaS / sigmaT;
        }
        if (--iter == 0) break;
    }
    t = dist;
    sampled = false;
    return mk3(1.f, 1.f, 1.f);
}

// ---- the same two loops, resumable: at most `max_steps` tracking steps per call ------------------------------------------------
// The trip count of a tracking loop is data dependent (the free path through the majorant), and inside the CTA-local
// wavefront a lane that tracks for 200 steps keeps its whole CTA at the barrier.  So the coroutine runs a walk through a
// heterogeneous medium in CHUNKS: the loop state (distance so far, running transmittance, remaining iteration budget) is
// saved in the slot, the slot stays in its wait state without posting a query, and the next step — after the slots have
// been re-sorted — continues where it stopped.  Same operations on the same values in the same order as het_tr /
// medium_sample above (the random-number stream is the path's own and is saved with the slot).
struct TrackState { float dist, tr; int iter; };
// Both loops as ONE: mode 0 = free-flight sampling (medium_sample's loop), 1 + evalTransmittanceType = transmittance
// (het_tr's three estimators).  The distance update, the grid lookup and the iteration budget — 4/5 of a tracking
// step's instructions — are the same code for every mode, so slots that track for different reasons (the path's free
// flight, a shadow leg, a medium-scatter shadow leg) share the lanes of a warp; what a mode does with the density is a
// few instructions behind a branch.  Same operations on the same values in the same order as het_tr / medium_sample above.
// `reason` on return: 0 = budget of this call used up (resume later), 1 = reached tmax, 2 = collision / roulette kill,
// 3 = iteration limit of the medium.
enum { TRK_SAMPLE = 0, TRK_TR0 = 1, TRK_TR1 = 2, TRK_TR2 = 3 };
__device__ __forceinline__ int het_track_chunk(const WMedium& M, const WHetero& H, const int mode, f3 o, f3 dir, float tmax, uint32_t& rng,
                                               TrackState& k, int max_steps) {
    float sigma = dot(ld3(M.sigmaT), mk3(0.212671f, 0.715160f, 0.072169f));
    const f3 p0 = ld3(H.p0);
    f3 d = ld3(H.p1) - p0;
    float maxDensity = 1 / H.invMaxDensity;
    float ce = 0.5f * maxDensity;
    for (int s_ = 0; s_ < max_steps; ++s_) {
        if (mode == TRK_TR2) k.dist += -logf(rng_next(rng)) * (1 / (maxDensity - ce) / sigma);
        else k.dist += -logf(rng_next(rng)) * H.invMaxDensity / sigma;
        if (k.dist >= tmax) return 1;
        f3 p = o + dir * k.dist;
        p = (p - p0) / d;
        const float dens = het_density(H, p);
        if (mode <= TRK_TR0) {
            if (dens * H.invMaxDensity > rng_next(rng)) return 2;
        } else {
            if (mode == TRK_TR1) k.tr *= 1.f - dens * H.invMaxDensity;
            else k.tr *= 1.f - (dens - ce) / (maxDensity - ce);
            if (k.tr < 0.1f) {
                float q = 1.f - k.tr;
                if (rng_next(rng) < q) return 2;
                if (mode == TRK_TR1) k.tr = 1;
                else k.tr /= (1.f - q);
            }
        }
        if (--k.iter == 0) return 3;
    }
    return 0;
}
// what a finished transmittance run is worth (the tails of het_tr's three loops)
__device__ __forceinline__ f3 het_track_tr_result(const WMedium& M, const WHetero& H, const int mode, const int reason, const TrackState& k, float tmax) {
    if (reason == 2) return mk3(0.f, 0.f, 0.f);
    if (mode == TRK_TR0) { const float r = reason == 1 ? k.tr : 0.f; return mk3(r, r, r); }
    if (mode == TRK_TR1) return mk3(k.tr, k.tr, k.tr);
    float sigma = dot(ld3(M.sigmaT), mk3(0.212671f, 0.715160f, 0.072169f));
    float maxDensity = 1 / H.invMaxDensity;
    float ce = 0.5f * maxDensity;
    float tc = expf(-tmax * ce * sigma);
    const float r = k.tr * tc;
    return mk3(r, r, r);
}
#ifndef PT_TRACK_CHUNK
#define PT_TRACK_CHUNK 16
#endif
constexpr int kTrackChunk = PT_TRACK_CHUNK;                  // tracking steps per wavefront step and slot (CTA-local kernel)

// ---- the coroutine ------------------------------------------------------------------------------------------------------
// Slot record (the wavefront's own planes, re-used):
//   o_rng     path ray origin, rng state             d_flags   path ray direction, state word (below)
//   beta_s    throughput, sample index               li_t      radiance so far, static samples consumed
//   hit0      hit of the path ray                    hit1      hit of the last secondary query
//   pend_o    secondary origin, -                    misd      secondary direction, tmax of the query
//   vis       transmittance of the walk so far, medium of the walk (bits, +1)
//   ldl       emitter radiance of the light sample, lightPdf * choicePdf
//   misf      BSDF value (light-sample evaluation / MIS sample), its pdf
//   beta_old  direct light Ld of this bounce so far, free-flight distance of a medium scatter
//   aux       |cos| of the MIS direction, -, -, -
//   carry     tracking loop in progress (CHUNKED): distance so far, running transmittance, iterations left (bits), 1 / 0
constexpr uint32_t H_ALIVE = 1u << 0, H_SPECULAR = 1u << 1;
constexpr int kHStateShift = 24;                 // bits 24..26: what the slot is waiting for
enum { HS_MAIN = 0, HS_WALK_MED = 1, HS_WALK_SURF = 2, HS_MIS = 3 };
// (bits 8..14 bounce counter and 16..23 medium + 1 as in the surface wavefront: kBounceShift, kMediumShift)

// Medium::Phase / SamplePhase (src/medium.h:197-246)
__device__ __forceinline__ float hg_phase(float g, f3 wo, f3 wi) {
    float phase = kInvFourPi;
    if (g != 0) {
        float costheta = dot(wo, wi);
        float cubicTerm = (1.f + g * g - 2.f * g * costheta);
        phase = kInvFourPi * (1.f - g * g) / sqrtf(cubicTerm * cubicTerm * cubicTerm);
    }
    return phase;
}
__device__ __forceinline__ f3 hg_sample(float g, float pa, float pb) {
    if (g == 0) { float pdf_; return uniform_sphere(pa, pb, pdf_); }
    float costheta;
    if (fabsf(g) < 1e-3f) costheta = 1.f - 2.f * pa;
    else {
        float sqrtTerm = (1.f - g * g) / (1.f - g + 2.f * g * pa);
        costheta = (1.f + g * g - sqrtTerm * sqrtTerm) / (2.f * g);
    }
    float sintheta = sqrtf(1.f - costheta * costheta);
    float phi = kTwoPi * pb;
    float sinphi = sinf(phi), cosphi = cosf(phi);
    return mk3(sintheta * cosphi, costheta, sintheta * sinphi);
}

// The walk through a heterogeneous medium a slot's resumed stage starts with, if any: the path's free flight up to the hit
// (HS_MAIN) or the transmittance of the current shadow leg (HS_WALK_*).  ONE definition for the three places that must
// agree on it: the sort key, the glue, and the tracking jobs of the trace phase.
struct TrackJob { int mode, med; float tmax; f3 o, d; };
__device__ __forceinline__ bool het_track_job(const SceneDev& sc, const Pool& pool, const uint32_t slot, const int state, const int medium,
                                              const float h0x, const f3 o, const f3 d, TrackJob& j) {
    if (state == HS_MAIN) {
        if (h0x < 0.f || medium < 0 || sc.mediums[medium].type == 0) return false;
        j.mode = TRK_SAMPLE; j.med = medium; j.tmax = h0x; j.o = o; j.d = d;
        return true;
    }
    if (state == HS_MIS) return false;
    const float4 h1 = pool.hit1[slot];
    const int mw = (int)__float_as_uint(pool.vis[slot].w) - 1;
    const bool invisible = h1.x >= 0.f;
    if ((invisible && sc.shade[__float_as_int(h1.y)].matIdx != -1) || mw < 0 || sc.mediums[mw].type == 0) return false;
    const float4 po = pool.pend_o[slot], md = pool.misd[slot];
    j.mode = TRK_TR0 + sc.het[mw].evalTransmittanceType; j.med = mw; j.tmax = invisible ? h1.x : md.w;
    j.o = mk3(po.x, po.y, po.z); j.d = mk3(md.x, md.y, md.z);
    return true;
}
#ifndef PT_HET_OVERLAP
#define PT_HET_OVERLAP 1
#endif
// carry.w of a slot: 0 = no walk in progress, 1 = walk in progress (distance, transmittance, iterations left saved),
// 2 + k = walk finished with reason 1 + k (PT_HET_OVERLAP: the trace phase walks, the next glue consumes)
// true when the slot waits for (more of) a walk: it is sorted in front, the glue leaves it alone, the trace phase walks it
__device__ __forceinline__ bool het_slot_tracks(const SceneDev& sc, const Pool& pool, const uint32_t slot) {
    const float4 df = pool.d_flags[slot];
    const uint32_t f = __float_as_uint(df.w);
    if (!(f & H_ALIVE)) return false;
    const float4 orng = pool.o_rng[slot];
    TrackJob j;
    if (!het_track_job(sc, pool, slot, (int)((f >> kHStateShift) & 7u), (int)((f >> kMediumShift) & 0xffu) - 1, pool.hit0[slot].x,
                       mk3(orng.x, orng.y, orng.z), mk3(df.x, df.y, df.z), j)) return false;
    return !PT_HET_OVERLAP || pool.carry[slot].w < 2.f;
}
// One chunk of a slot's walk, as a work item of the TRACE phase (PT_HET_OVERLAP): the slots that track post no ray, so
// their walks run next to the other slots' traversals instead of in front of them.
__device__ __forceinline__ void het_track_step(const SceneDev& sc, const Pool& pool, const uint32_t slot) {
    const float4 df = pool.d_flags[slot], orng = pool.o_rng[slot];
    const uint32_t f = __float_as_uint(df.w);
    uint32_t rng = __float_as_uint(orng.w);
    TrackJob j;
    if (!het_track_job(sc, pool, slot, (int)((f >> kHStateShift) & 7u), (int)((f >> kMediumShift) & 0xffu) - 1, pool.hit0[slot].x,
                       mk3(orng.x, orng.y, orng.z), mk3(df.x, df.y, df.z), j)) return;
    const float4 cy = pool.carry[slot];
    TrackState tk;
    if (cy.w != 0.f) { tk.dist = cy.x; tk.tr = cy.y; tk.iter = __float_as_int(cy.z); }
    else { tk.dist = 0.f; tk.tr = 1.f; tk.iter = sc.het[j.med].iterMax; }
    const int reason = het_track_chunk(sc.mediums[j.med], sc.het[j.med], j.mode, j.o, j.d, j.tmax, rng, tk, kTrackChunk);
    pool.carry[slot] = make_float4(tk.dist, tk.tr, __int_as_float(tk.iter), reason == 0 ? 1.f : 1.f + (float)reason);
    pool.o_rng[slot].w = __uint_as_float(rng);
}

// QUEUED: the query goes into the ray queue (global wavefront: k_trace / k_trace_small pull it).  Otherwise the caller
// traces it itself right away (warp-local stepping of the CTA-local kernel, k_wave.cuh) and gets its kind in `posted`
// (0 none, 1 path ray = queue kind 0, 2 secondary query = queue kind 2).
template <uint32_t MATS, bool FUSED, bool QUEUED>
__device__ __forceinline__ void het_slot(const ShadeArgs& a, const Pool& pool, const RayQueue& q, const uint32_t parity,
                                         uint32_t* cta_retired, uint32_t* cta_busy, const uint32_t slot, const uint32_t gslot,
                                         const uint32_t pool_n, const unsigned long long next_snapshot, uint32_t& posted) {
    const SceneDev& sc = a.sc;
    const uint32_t lane = pt_lane(), lt = (1u << lane) - 1u;
    const float4 df = pool.d_flags[slot], orng = pool.o_rng[slot], lt4 = pool.li_t[slot];
    uint32_t flags = __float_as_uint(df.w);
    uint32_t rng = __float_as_uint(orng.w);
    f3 o = mk3(orng.x, orng.y, orng.z), d = mk3(df.x, df.y, df.z);
    const bool alive = (flags & H_ALIVE) != 0u;
    uint32_t kdone = __float_as_uint(lt4.w);
    f3 beta = mk3(1, 1, 1), Li = mk3(0, 0, 0);
    uint32_t sample = 0u;
    int bounces = (int)((flags >> kBounceShift) & 0x7fu);
    int medium = (int)((flags >> kMediumShift) & 0xffu) - 1;
    bool specular = (flags & H_SPECULAR) != 0u;
    int state = (int)((flags >> kHStateShift) & 7u);
    bool finished = false;
    uint32_t post = 0u;                  // 0 nothing, 1 path ray (kind 0), 2 secondary query (kind 2)

    if (alive) {
        const float4 bs = pool.beta_s[slot];
        beta = mk3(bs.x, bs.y, bs.z); sample = __float_as_uint(bs.w);
        Li = mk3(lt4.x, lt4.y, lt4.z);
        const float4 h0 = pool.hit0[slot];
        // stages of one bounce; a stage either falls through to a later one, posts a query, or ends the path
        enum { G_MAIN, G_WALK, G_AFTER_NEE, G_MIS, G_CONT, G_BOUNCE_END, G_OUT };
        int g = state == HS_MAIN ? G_MAIN : (state == HS_MIS ? G_MIS : G_WALK);
        // ---- tracking prelude (CTA-local kernel): the walk through a heterogeneous medium the resumed stage starts with —
        // the path's free flight up to the hit (G_MAIN) or the transmittance of a shadow leg (G_WALK) — runs HERE, in one
        // loop for every mode, before the stages diverge; the stage then only consumes the result.  (Both stages draw their
        // first random number inside that walk, so the stream is consumed in the reference's order.)  At most kTrackChunk
        // steps per call: an unfinished walk is saved in `carry`, the slot keeps its state and posts nothing.
        int trk_mode = -1, trk_reason = 0;
        TrackState tk; tk.dist = 0.f; tk.tr = 1.f; tk.iter = 0;
        if (FUSED) {
            TrackJob j;
            if (het_track_job(sc, pool, slot, state, medium, h0.x, o, d, j)) {
                trk_mode = j.mode;
                const float4 cy = pool.carry[slot];
#if PT_HET_OVERLAP
                // the walk itself is a work item of the TRACE phase (het_track_step): consume a finished one, else wait
                if (cy.w >= 2.f) {
                    tk.dist = cy.x; tk.tr = cy.y; tk.iter = __float_as_int(cy.z); trk_reason = (int)cy.w - 1;
                    st_rec<FUSED>(pool.carry + slot, make_float4(0.f, 0.f, 0.f, 0.f));
                } else { post = 0u; g = G_OUT; }
#else
                if (cy.w != 0.f) { tk.dist = cy.x; tk.tr = cy.y; tk.iter = __float_as_int(cy.z); }
                else tk.iter = sc.het[j.med].iterMax;
                trk_reason = het_track_chunk(sc.mediums[j.med], sc.het[j.med], j.mode, j.o, j.d, j.tmax, rng, tk, kTrackChunk);
                if (trk_reason == 0) {
                    st_rec<FUSED>(pool.carry + slot, make_float4(tk.dist, tk.tr, __int_as_float(tk.iter), 1.f));
                    post = 0u; g = G_OUT;                                                       // same state, no query: resume next step
                } else if (cy.w != 0.f) st_rec<FUSED>(pool.carry + slot, make_float4(0.f, 0.f, 0.f, 0.f));
#endif
            }
        }
        SurfaceHit h;
        bool have_h = false;
        f3 Ld = mk3(0, 0, 0);
        if (state != HS_MAIN) {
            if (state != HS_WALK_MED) { reconstruct_hit(sc, o, d, h0.x, __float_as_int(h0.y), h0.z, h0.w, h); have_h = true; }
            const float4 bo = pool.beta_old[slot]; Ld = mk3(bo.x, bo.y, bo.z);
        }
        while (g != G_OUT) {
            if (g == G_MAIN) {                                                                  // :1051-1124
                if (h0.x < 0.f) {
                    if ((bounces == 0 || specular) && sc.inf.isvalid) Li += beta * inf_le(sc.inf, d);
                    finished = true; g = G_OUT; continue;
                }
                reconstruct_hit(sc, o, d, h0.x, __float_as_int(h0.y), h0.z, h0.w, h); have_h = true;
                float sampledDist = 0.f; bool sampledMedium = false;
                if (medium >= 0) {
                    const WMedium& M_ = sc.mediums[medium];
                    if (FUSED && M_.type != 0) {                 // heterogeneous, CTA-local kernel: the prelude walked the free flight
                        sampledDist = tk.dist; sampledMedium = trk_reason == 2;
                        beta *= sampledMedium ? ld3(M_.sigmaS) / ld3(M_.sigmaT) : mk3(1.f, 1.f, 1.f);
                    } else beta *= medium_sample(sc, medium, o, d, h0.x, rng, sampledDist, sampledMedium);
                }
                if (is_black(beta)) { finished = true; g = G_OUT; continue; }                    // :1070
                if (sampledMedium) {                                                            // :1071-1088: light sample, then Tr()
                    float u = rng_next(rng);
                    float choicePdf;
                    int idx = lookup_light(sc, u, choicePdf);
                    if (idx < 0) idx = 0;
                    f3 samplePos = o + sampledDist * d;
                    float ua = rng_next(rng), ub = rng_next(rng);
                    LightSample ls;
                    if (idx != sc.n_lights) area_sample(sc.lights[idx], samplePos, ua, ub, sc.eps, ls);
                    else inf_sample(sc.inf, ua, ub, sc.eps, ls);
                    st_rec<FUSED>(pool.ldl + slot, make_float4(ls.radiance.x, ls.radiance.y, ls.radiance.z, ls.pdf * choicePdf));
                    st_rec<FUSED>(pool.beta_old + slot, make_float4(0.f, 0.f, 0.f, sampledDist));
                    st_rec<FUSED>(pool.pend_o + slot, make_float4(samplePos.x, samplePos.y, samplePos.z, 0.f));
                    st_rec<FUSED>(pool.misd + slot, make_float4(ls.dir.x, ls.dir.y, ls.dir.z, ls.tmax));
                    st_rec<FUSED>(pool.vis + slot, make_float4(1.f, 1.f, 1.f, __uint_as_float((uint32_t)(medium + 1))));
                    state = HS_WALK_MED; post = 2u; g = G_OUT; continue;
                }
                if ((bounces == 0 || specular) && h.lightIdx != -1) {                           // emitter, :1103-1115
                    const WLight& L = sc.lights[h.lightIdx];
                    f3 le = dot(h.nor, -d) > 0.f ? ld3(L.radiance) : mk3(0.f, 0.f, 0.f);
                    f3 tr = mk3(1.f, 1.f, 1.f);
                    if (medium >= 0) tr = medium_tr(sc, medium, o, d, h0.x, rng);
                    Li += tr * beta * le;
                    finished = true; g = G_OUT; continue;
                }
                if (h.matIdx == -1) {                                                           // medium boundary, :1117-1124
                    medium = dot(d, h.nor) > 0 ? h.mediumOutside : h.mediumInside;
                    o = h.pos;
                    state = HS_MAIN; post = 1u; g = G_OUT; continue;                           // `bounces--; continue;`
                }
                const Material mat = sc.mats[h.matIdx];
                if (is_delta(mat.type)) { g = G_CONT; continue; }
                float u = rng_next(rng);                                                        // :1128-1160
                float choicePdf;
                int idx = lookup_light(sc, u, choicePdf);
                if (idx < 0) idx = 0;
                float ua = rng_next(rng), ub = rng_next(rng);
                LightSample ls;
                if (idx != sc.n_lights) area_sample(sc.lights[idx], h.pos, ua, ub, sc.eps, ls);
                else inf_sample(sc.inf, ua, ub, sc.eps, ls);
                Ld = mk3(0.f, 0.f, 0.f);
                if (is_black(ls.radiance)) { g = G_AFTER_NEE; continue; }
                f3 fr; float samplePdf;
                eval_bsdf_m<MATS>(mat, material_albedo(sc, mat, h.uv), -d, ls.dir, h.nor, h.dpdu, fr, samplePdf);
                st_rec<FUSED>(pool.ldl + slot, make_float4(ls.radiance.x, ls.radiance.y, ls.radiance.z, ls.pdf * choicePdf));
                st_rec<FUSED>(pool.misf + slot, make_float4(fr.x, fr.y, fr.z, samplePdf));
                st_rec<FUSED>(pool.beta_old + slot, make_float4(0.f, 0.f, 0.f, 0.f));
                st_rec<FUSED>(pool.pend_o + slot, make_float4(h.pos.x, h.pos.y, h.pos.z, 0.f));
                st_rec<FUSED>(pool.misd + slot, make_float4(ls.dir.x, ls.dir.y, ls.dir.z, ls.tmax));
                st_rec<FUSED>(pool.vis + slot, make_float4(1.f, 1.f, 1.f, __uint_as_float((uint32_t)(medium + 1))));
                state = HS_WALK_SURF; post = 2u; g = G_OUT; continue;
            }
            if (g == G_WALK) {                                                                  // one leg of Tr(), :298-322
                const float4 h1 = pool.hit1[slot], po = pool.pend_o[slot], md = pool.misd[slot], pv = pool.vis[slot];
                f3 ow = mk3(po.x, po.y, po.z), dw = mk3(md.x, md.y, md.z);
                float remain = md.w;
                f3 tr = mk3(pv.x, pv.y, pv.z);
                int mw = (int)__float_as_uint(pv.w) - 1;
                const bool invisible = h1.x >= 0.f;
                bool again = false;
                if (invisible && sc.shade[__float_as_int(h1.y)].matIdx != -1) tr = mk3(0, 0, 0);
                else {
                    const float seg = invisible ? h1.x : remain;
                    if (mw >= 0) {
                        const WMedium& M_ = sc.mediums[mw];
                        if (FUSED && M_.type != 0) tr *= het_track_tr_result(M_, sc.het[mw], trk_mode, trk_reason, tk, seg);   // the prelude walked this leg
                        else tr *= medium_tr(sc, mw, ow, dw, seg, rng);
                    }
                    if (invisible) {
                        const int prim = __float_as_int(h1.y);
                        const WShade& s = sc.shade[prim];
                        f3 nor;
                        if (s.type == 0) nor = normalize(lin3_seq(1.f - h1.z - h1.w, ld3(s.n1), h1.z, ld3(s.n2), h1.w, ld3(s.n3)));
                        else nor = normalize((ow + seg * dw) - ld3(s.n1));
                        mw = dot(dw, nor) > 0 ? s.mediumOutside : s.mediumInside;
                        remain -= seg;
                        ow = ow + seg * dw;                                                     // Ray(ray(ray.tmax), ray.d, m, eps, tmax)
                        st_rec<FUSED>(pool.pend_o + slot, make_float4(ow.x, ow.y, ow.z, 0.f));
                        st_rec<FUSED>(pool.misd + slot, make_float4(dw.x, dw.y, dw.z, remain));
                        st_rec<FUSED>(pool.vis + slot, make_float4(tr.x, tr.y, tr.z, __uint_as_float((uint32_t)(mw + 1))));
                        again = true;
                    }
                }
                if (again) { post = 2u; g = G_OUT; continue; }                                  // same state: next leg
                const float4 l = pool.ldl[slot];
                const f3 radiance = mk3(l.x, l.y, l.z);
                if (state == HS_WALK_MED) {                                                     // :1089-1101
                    const float g_ = sc.mediums[medium].g;
                    float phase = hg_phase(g_, -d, dw);
                    if (!is_black(radiance)) Li += tr * beta * phase * radiance / l.w;
                    float pa = rng_next(rng), pb = rng_next(rng);
                    f3 dir = hg_sample(g_, pa, pb);
                    o = o + pool.beta_old[slot].w * d; d = dir;
                    specular = false;
                    g = G_BOUNCE_END; continue;
                }
                const float4 mf = pool.misf[slot];                                               // :1148-1155
                float weight = power_heuristic(1, l.w, 1, mf.w);
                Ld += weight * tr * mk3(mf.x, mf.y, mf.z) * radiance * fabsf(dot(h.nor, dw)) / l.w;
                g = G_AFTER_NEE; continue;
            }
            if (g == G_AFTER_NEE) {                                                             // :1157-1163: BSDF-sampled light ray
                const Material mat = sc.mats[h.matIdx];
                float s0 = rng_next(rng), s1 = rng_next(rng), s2 = rng_next(rng);
                f3 out, fr; float pdf;
                sample_bsdf_m<MATS>(mat, material_albedo(sc, mat, h.uv), -d, h.nor, h.dpdu, mk3(s0, s1, s2), out, fr, pdf);
                if (is_black(fr) || pdf == 0 || !mis_ray_may_reach_emitter(sc, h.pos, out)) { g = G_MIS + 100; }   // no light ray: straight to the accumulation
                else {
                    st_rec<FUSED>(pool.misf + slot, make_float4(fr.x, fr.y, fr.z, pdf));
                    st_rec<FUSED>(pool.beta_old + slot, make_float4(Ld.x, Ld.y, Ld.z, 0.f));
                    st_rec<FUSED>(pool.aux + slot, make_float4(fabsf(dot(out, h.nor)), 0.f, 0.f, 0.f));
                    st_rec<FUSED>(pool.pend_o + slot, make_float4(h.pos.x, h.pos.y, h.pos.z, 0.f));
                    st_rec<FUSED>(pool.misd + slot, make_float4(out.x, out.y, out.z, INFINITY));
                    state = HS_MIS; post = 2u; g = G_OUT; continue;
                }
            }
            if (g == G_MIS) {                                                                   // :1164-1209
                const float4 h1 = pool.hit1[slot], md = pool.misd[slot], mf = pool.misf[slot];
                const f3 out = mk3(md.x, md.y, md.z), fr = mk3(mf.x, mf.y, mf.z);
                const float pdf = mf.w, absdot = pool.aux[slot].x;
                if (h1.x >= 0.f) {
                    const int prim = __float_as_int(h1.y);
                    const int lightIdx = sc.shade[prim].lightIdx;
                    f3 p = h.pos + h1.x * out;
                    f3 n = hit_normal(sc, p, prim, h1.z, h1.w);
                    f3 radiance = mk3(0.f, 0.f, 0.f);
                    if (lightIdx != -1) {
                        const WLight& L = sc.lights[lightIdx];
                        if (dot(n, -out) > 0.f) radiance = ld3(L.radiance);
                        if (!is_black(radiance)) {
                            float pdfA = 1.f / L.area;
                            float cp = sc.cdf[lightIdx + 1] - sc.cdf[lightIdx];
                            float lenSquare = dot(p - h.pos, p - h.pos);
                            float costheta = fabsf(dot(n, out));
                            float lPdf = pdfA * lenSquare / (costheta);
                            float weight = power_heuristic(1, pdf, 1, lPdf * cp);
                            f3 tr = mk3(1.f, 1.f, 1.f);
                            if (medium >= 0) tr = medium_tr(sc, medium, h.pos, out, h1.x, rng);
                            Ld += weight * tr * fr * radiance * absdot / pdf;                    // :1185
                        }
                    }
                } else if (sc.inf.isvalid) {
                    f3 radiance = inf_le(sc.inf, out);
                    float cp = sc.cdf[sc.n_lights + 1] - sc.cdf[sc.n_lights];
                    float weight = power_heuristic(1, pdf, 1, kInvFourPi * cp);
                    f3 tr = mk3(1.f, 1.f, 1.f);
                    if (medium >= 0) tr = medium_tr(sc, medium, h.pos, out, INFINITY, rng);
                    Ld += weight * tr * fr * radiance * absdot / pdf;                            // :1205
                }
            }
            if (g == G_MIS || g == G_MIS + 100) {                                               // Li += beta * Ld, :1211
#if defined(__CUDA_ARCH__)
                Li = mk3(__fmaf_rn(beta.x, Ld.x, Li.x), __fmaf_rn(beta.y, Ld.y, Li.y), __fmaf_rn(beta.z, Ld.z, Li.z));
#else
                Li += beta * Ld;
#endif
                g = G_CONT;
            }
            if (g == G_CONT) {                                                                  // :1213-1228
                const Material mat = sc.mats[h.matIdx];
                float c0 = rng_next(rng), c1 = rng_next(rng), c2 = rng_next(rng);
                f3 out, fr; float pdf;
                sample_bsdf_m<MATS>(mat, material_albedo(sc, mat, h.uv), -d, h.nor, h.dpdu, mk3(c0, c1, c2), out, fr, pdf);
                if (is_black(fr)) { finished = true; g = G_OUT; continue; }
                beta *= fr * fabsf(dot(h.nor, out)) / pdf;
                specular = is_delta(mat.type);
                int m = dot(out, h.nor) > 0 ? h.mediumOutside : h.mediumInside;                  // :1224-1226
                m = dot(-d, h.nor) * dot(out, h.nor) > 0 ? medium : m;
                medium = m;
                o = h.pos; d = out;
                g = G_BOUNCE_END;
            }
            if (g == G_BOUNCE_END) {                                                            // :1230-1236 and the loop bound
                if (bounces > 3) {
                    float illumate = clampf(1.f - luminance_rr<true>(beta), 0.f, 1.f);
                    if (rng_next(rng) < illumate) { finished = true; g = G_OUT; continue; }
                    beta /= (1 - illumate);
                }
                ++bounces;
                if (bounces >= sc.max_depth) { finished = true; g = G_OUT; continue; }
                state = HS_MAIN; post = 1u; g = G_OUT;
            }
        }
        (void)have_h;
    }

    // ---- retire + regenerate (same hand-out as the surface wavefront: static share per slot, then the global counter)
    const bool have_more = kdone < a.batch.k_static || next_snapshot < a.batch.total;
    const bool want_new = (finished || !alive) && have_more;
    if (finished) st_rec<FUSED>(a.samples + sample, make_float4(Li.x, Li.y, Li.z, 1.f));
    const bool take_static = want_new && kdone < a.batch.k_static;
    const uint32_t m_fin = __ballot_sync(kFullMask, want_new && !take_static);
    const uint32_t m_ret = __ballot_sync(kFullMask, finished);
    unsigned long long sbase = 0ull;
    if (lane == 0u) {
        if (m_fin) sbase = atomicAdd(&a.counters->next_sample, (unsigned long long)__popc(m_fin));
        if (m_ret) {
            if (FUSED) atomicAdd(cta_retired, (uint32_t)__popc(m_ret));
            else atomicAdd(&a.counters->done_samples, (unsigned long long)__popc(m_ret));
        }
    }
    sbase = __shfl_sync(kFullMask, sbase, 0);
    bool now_alive = alive && !finished;
    if (want_new) {
        unsigned long long s;
        if (take_static) { s = (unsigned long long)gslot + (unsigned long long)kdone * (unsigned long long)pool_n; ++kdone; }
        else s = sbase + (unsigned long long)__popc(m_fin & lt);
        if (s < a.batch.total) {
            sample = (uint32_t)s;
            const uint32_t npix = (uint32_t)a.map.n_local_pixels;
            const uint32_t it_local = sample / npix, local = sample - it_local * npix;
            uint32_t x, y;
            local_to_xy(a.map, local, x, y);
            const uint32_t pixel = x + y * (uint32_t)a.map.width;                               // :1028
            rng = rng_seed(pixel, a.batch.first_iter + it_local);                               // :1033
            float offsetx = rng_next(rng) - 0.5f;                                               // :1037-1043
            float offsety = rng_next(rng) - 0.5f;
            float a0 = rng_next(rng), a1 = rng_next(rng);
            f2 aperture = mk2(0.f, 0.f);
            if (a.cam.apertureRadius > 0.00001f) aperture = uniform_disk(a0, a1);
            camera_ray(a.cam, x + offsetx, y + offsety, aperture, o, d);
            beta = mk3(1.f, 1.f, 1.f); Li = mk3(0.f, 0.f, 0.f);
            bounces = 0; specular = false; medium = a.cam.medium; state = HS_MAIN;
            post = 1u; now_alive = true;
        }
    }
    // ---- post the query: one aggregated queue reservation per warp
    posted = post;
    if (QUEUED) {
        const uint32_t mp = __ballot_sync(kFullMask, post != 0u);
        uint32_t qbase = 0u;
        if (lane == 0u && mp) qbase = atomicAdd(&q.ctl->tail[parity & 1u], (uint32_t)__popc(mp));
        qbase = __shfl_sync(kFullMask, qbase, 0);
        if (post) q.entries[qbase + (uint32_t)__popc(mp & lt)] = slot | ((post == 2u ? 2u : 0u) << kKindShift);
    }
    if (!now_alive && !alive) {
        // stayed dead; keep kdone (the dynamic hand-out may have consumed nothing)
        if (want_new) st_rec<FUSED>(pool.li_t + slot, make_float4(0.f, 0.f, 0.f, __uint_as_float(kdone)));
        return;
    }
    if (FUSED && cta_busy) *cta_busy = 1u;
    uint32_t nf = now_alive ? H_ALIVE : 0u;
    if (specular) nf |= H_SPECULAR;
    nf |= ((uint32_t)bounces & 0x7fu) << kBounceShift;
    nf |= ((uint32_t)(medium + 1) & 0xffu) << kMediumShift;
    nf |= ((uint32_t)state & 7u) << kHStateShift;
    st_rec<FUSED>(pool.o_rng + slot, make_float4(o.x, o.y, o.z, __uint_as_float(rng)));
    st_rec<FUSED>(pool.d_flags + slot, make_float4(d.x, d.y, d.z, __uint_as_float(nf)));
    st_rec<FUSED>(pool.beta_s + slot, make_float4(beta.x, beta.y, beta.z, __uint_as_float(sample)));
    st_rec<FUSED>(pool.li_t + slot, make_float4(Li.x, Li.y, Li.z, __uint_as_float(kdone)));
}

// The same stage as a kernel of the GLOBAL wavefront (path pool in HBM, k_trace / k_trace_small behind it): scenes with a
// heterogeneous medium whose primitives do not fit in shared memory.
template <uint32_t MATS>
__global__ void __launch_bounds__(128, 6) k_het_shade(const ShadeArgs a) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;      // the pool size is a multiple of the block size
    uint32_t posted;
    het_slot<MATS, false, true>(a, a.pool, a.q, a.parity, nullptr, nullptr, slot, slot, (uint32_t)a.pool.n, a.counters->next_sample, posted);
}

}  // namespace pt
