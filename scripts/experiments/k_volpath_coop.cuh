// NOT COMPILED, NOT SHIPPED — kept as the starting point of next round's three-stage wavefront for heterogeneous media.
// Status when it left the tree (round 1): bit-exact against the oracle in the 1-lane CPU emulation
// (tests/test_wavefront_emu.py -k "smoke or heterogeneous" with this kernel wired in), 4.25 ms per launch on B200 where
// k_volpath_seq takes 2.42 ms (profiles/r01z_het_seq.txt).  Its value is the state machine: the bounce
// of Volpath cut at every traversal / tracking call site, with the state that has to survive each cut made explicit —
// the per-slot record a k_het_shade kernel would load and store around k_trace_small.
// (Call sites below predate the seq_light_sample / seq_sample_bsdf / seq_eval_bsdf helpers of k_volpath_seq.cuh.)
//
// k_volpath_coop.cuh — `vpt` with HETEROGENEOUS media as a warp-cooperative state machine (Volpath,
// src/pathtracer.cu:1025-1242, Heterogeneous::Sample / Tr, src/medium.h:52-179).
//
// k_volpath_seq.cuh runs one path per lane in the reference's own control flow; ncu shows what that costs: 3.8 of 32
// lanes active per issued instruction (profiles/r01z_het_seq.txt; the reference's Volpath: 4.2), because the lanes of
// a warp sit in different call sites of the same two expensive loops — closest-hit traversal (7 call sites per
// bounce once Tr()'s boundary walk is unrolled) and delta / ratio tracking through the density grid (5 call sites).
//
// Here every lane is a coroutine.  The bounce is cut at each of those call sites into short glue states; a lane that
// needs a traversal or a tracking run posts it as an OPERATION (ray + interval + medium in fixed registers) and all
// lanes of the warp that have the same operation pending execute it TOGETHER, in one copy of the loop:
//
//     for (;;) {  glue: switch (state) ... until the lane has posted an operation (or has run out of samples)
//                 vote:  lanes waiting for a traversal / for tracking steps
//                 TRAV:  seq_closest_hit for all waiting lanes at once
//                 TRACK: up to kTrackChunk tracking steps for all waiting lanes at once; unfinished runs stay posted,
//                        their (dist, tr, iterations left) live in registers }
//
// Draw order, arithmetic and results per path are those of k_volpath_seq (bit-exact in the 1-lane CPU emulation
// against the oracle, tests/test_wavefront_emu.py); only the interleaving of different paths changes.
// A finished path takes the next (iteration, pixel) from the global counter in its next glue phase.
#pragma once
#include "k_volpath_seq.cuh"

namespace pt {

enum : int { OP_NONE = 0, OP_TRAV = 1, OP_TRACK = 2, OP_IDLE = 3 };
enum : int { TK_SAMPLE = 0, TK_DELTA = 1, TK_RATIO = 2, TK_RESIDUAL = 3 };
enum : int {
    ST_START = 0,      // (re)generate, post the bounce's closest hit
    ST_HIT0,           // hit record of the bounce ray is in; post the free-flight sampling
    ST_SAMPLED,        // Medium::Sample done: scatter / emitter / boundary / surface
    ST_WALK_START,     // Tr(): post the closest hit of the current segment
    ST_WALK_HIT,       //       segment end known: post the segment's transmittance
    ST_WALK_TR,        //       accumulate, cross the boundary or return
    ST_SCATTER_LIT,    // medium scattering: direct light arrived, sample the phase function
    ST_EMIT_TR,        // emitter seen directly: transmittance of the bounce segment arrived
    ST_NEE_DONE,       // surface: light-sampled estimate arrived; sample the BSDF for MIS and post that ray
    ST_MIS_HIT,        // surface: hit record of the MIS ray is in; post its transmittance
    ST_MIS_TR,         // surface: add the BSDF-sampled estimate
    ST_CONT            // surface: continuation sample, throughput, Russian roulette
};

#ifdef PT_COOP_STATS
static unsigned long long g_coop_stats[8];   // 0 trav ops, 1 track ops, 2 track steps, 3 track chunks, 4 samples, 5 glue states
#define PT_STAT(i, n) (g_coop_stats[i] += (n))
#else
#define PT_STAT(i, n) ((void)0)
#endif
constexpr int kTrackChunk = 8;     // tracking steps per cooperative phase before the warp votes again

template <uint32_t MATS>
__global__ void __launch_bounds__(128) k_volpath_coop(const __grid_constant__ SeqArgs a) {
    const SceneDev& sc = a.sc;
    const uint32_t npix = (uint32_t)a.map.n_local_pixels;
    const f3 kLum = mk3(0.212671f, 0.715160f, 0.072169f);
    // ---- path
    bool alive = false, specular = false;
    uint32_t sample = 0u, rng = 0u, nrays = 0u;
    f3 o = mk3(0, 0, 0), d = mk3(0, 0, 1), beta = mk3(1, 1, 1), Li = mk3(0, 0, 0);
    int medium = -1, bounces = 0;
    // ---- coroutine
    int st = ST_START, op = OP_NONE;
    // operation registers: the ray / interval / medium of the posted operation, and its results
    f3 q_o = o, q_d = d; float q_tmax = 0.f; int q_medium = -1;
    bool r_hit = false; Hit r_h; r_h.t = 0.f; r_h.prim = -1; r_h.b1 = r_h.b2 = 0.f;
    int tk_kind = TK_DELTA, tk_iter = 0; float tk_dist = 0.f, tk_tr = 1.f;
    f3 r_trv = mk3(1, 1, 1);                       // transmittance result (any medium type)
    f3 w3 = mk3(1, 1, 1); float sampledDist = 0.f; bool sampledMedium = false;      // Medium::Sample result
    // ---- state that lives across the operations of one bounce
    Hit h0; h0.t = 0.f; h0.prim = -1; h0.b1 = h0.b2 = 0.f;
    SurfaceHit h; h.pos = h.nor = h.dpdu = mk3(0, 0, 0); h.uv = mk2(0, 0); h.matIdx = h.lightIdx = h.mediumInside = h.mediumOutside = -1;
    LightSample ls; ls.radiance = ls.dir = mk3(0, 0, 0); ls.tmax = ls.pdf = 0.f;
    float choicePdf = 0.f, samplePdf = 0.f;
    f3 fr_l = mk3(0, 0, 0), Ld = mk3(0, 0, 0);
    bool nee_valid = false;
    f3 out_m = mk3(0, 0, 0), fr_m = mk3(0, 0, 0), mis_rad = mk3(0, 0, 0); float pdf_m = 0.f, absdot = 0.f, mis_w = 0.f;
    // Tr() walk
    f3 tw_o = o, tw_d = d, tw_tr = mk3(1, 1, 1); float tw_remain = 0.f, tw_seg = 0.f; int tw_medium = -1, tw_ret = ST_START; bool tw_hit = false;

    // posts the transmittance of [0, tmax] along (ro, rd) in medium m, or resolves it on the spot (no / homogeneous medium)
#define PT_POST_TR(m_, ro_, rd_, tmax_)                                                                              \
    do {                                                                                                             \
        const int m__ = (m_);                                                                                        \
        if (m__ < 0) r_trv = mk3(1.f, 1.f, 1.f);                                                                     \
        else if (sc.mediums[m__].type == 0) r_trv = exp3(ld3(sc.mediums[m__].sigmaT) * (-(tmax_)));                  \
        else {                                                                                                       \
            q_medium = m__; q_o = (ro_); q_d = (rd_); q_tmax = (tmax_);                                              \
            tk_kind = TK_DELTA + sc.het[m__].evalTransmittanceType; tk_dist = 0.f; tk_tr = 1.f; tk_iter = sc.het[m__].iterMax; \
            op = OP_TRACK;                                                                                           \
        }                                                                                                            \
    } while (0)
#define PT_FINISH()                                                                                                  \
    do { st_pool(a.samples + sample, make_float4(Li.x, Li.y, Li.z, 1.f)); alive = false; st = ST_START; } while (0)
    // :1230-1236 + the loop condition; `bounces--; continue;` of the boundary case never comes here
#define PT_END_BOUNCE()                                                                                              \
    do {                                                                                                             \
        bool fin__ = false;                                                                                          \
        if (bounces > 3) {                                                                                           \
            float illumate = clampf(1.f - luminance_rr<true>(beta), 0.f, 1.f);                                       \
            if (rng_next(rng) < illumate) fin__ = true;                                                              \
            else beta /= (1 - illumate);                                                                             \
        }                                                                                                            \
        if (!fin__) { ++bounces; if (bounces >= sc.max_depth) fin__ = true; }                                        \
        if (fin__) PT_FINISH(); else st = ST_START;                                                                  \
    } while (0)

    for (;;) {
        // ================= glue: run this lane's state machine until it has posted an operation =================
        while (op == OP_NONE) {
            PT_STAT(5, 1);
            switch (st) {
            case ST_START: {
                if (!alive) {                                                                   // ray generation, :1026-1048
                    const unsigned long long s = atomicAdd(&a.counters->next_sample, 1ull);
                    if (s >= a.batch.total) { op = OP_IDLE; break; }
                    sample = (uint32_t)s;
                    const uint32_t it_local = sample / npix, local = sample - it_local * npix;
                    uint32_t x, y;
                    local_to_xy(a.map, local, x, y);
                    const uint32_t pixel = x + y * (uint32_t)a.map.width;
                    rng = rng_seed(pixel, a.batch.first_iter + it_local);
                    float offsetx = rng_next(rng) - 0.5f;
                    float offsety = rng_next(rng) - 0.5f;
                    float a0 = rng_next(rng), a1 = rng_next(rng);
                    f2 aperture = mk2(0.f, 0.f);
                    if (a.cam.apertureRadius > 0.00001f) aperture = uniform_disk(a0, a1);
                    camera_ray(a.cam, x + offsetx, y + offsety, aperture, o, d);
                    beta = mk3(1.f, 1.f, 1.f); Li = mk3(0.f, 0.f, 0.f);
                    bounces = 0; specular = false; medium = a.cam.medium;
                    alive = true;
                    if (sc.max_depth <= 0) {                                                    // the bounce loop never runs
                        st_pool(a.samples + sample, make_float4(0.f, 0.f, 0.f, 1.f));
                        alive = false;
                        break;
                    }
                }
                q_o = o; q_d = d; q_tmax = INFINITY;
                op = OP_TRAV; st = ST_HIT0;
                break;
            }
            case ST_HIT0: {
                if (!r_hit) {
                    if ((bounces == 0 || specular) && sc.inf.isvalid) Li += beta * inf_le(sc.inf, d);
                    PT_FINISH();
                    break;
                }
                h0 = r_h;
                reconstruct_hit(sc, o, d, h0.t, h0.prim, h0.b1, h0.b2, h);
                w3 = mk3(1.f, 1.f, 1.f); sampledDist = 0.f; sampledMedium = false;
                st = ST_SAMPLED;
                if (medium >= 0) {
                    const WMedium& M = sc.mediums[medium];
                    if (M.type == 0) {                                                          // Homogeneous::Sample, src/medium.h:19-49
                        f3 sigmaT = ld3(M.sigmaT), sigmaS = ld3(M.sigmaS);
                        float sigma = dot(sigmaT, kLum);
                        float dist = -logf(rng_next(rng)) / sigma;
                        f3 Tr = exp3(sigmaT * -dist);
                        float pdf = sigma * expf(sigma * -dist);
                        sampledMedium = dist < h0.t;
                        sampledDist = dist;
                        w3 = sampledMedium ? (Tr * sigmaS / pdf) : sigmaT * Tr / pdf;
                    } else {                                                                    // Heterogeneous::Sample, :137-157
                        q_medium = medium; q_o = o; q_d = d; q_tmax = h0.t;
                        tk_kind = TK_SAMPLE; tk_dist = 0.f; tk_tr = 1.f; tk_iter = sc.het[medium].iterMax;
                        op = OP_TRACK;
                    }
                }
                break;
            }
            case ST_SAMPLED: {
                if (medium >= 0) beta *= w3;
                if (is_black(beta)) { PT_FINISH(); break; }                                     // :1070
                if (sampledMedium) {                                                            // :1071-1088
                    float u = rng_next(rng);
                    f3 samplePos = o + sampledDist * d;
                    float ua = rng_next(rng), ub = rng_next(rng);
                    seq_light_sample(sc, samplePos, u, ua, ub, ls, choicePdf);
                    tw_o = samplePos; tw_d = ls.dir; tw_remain = ls.tmax; tw_medium = medium; tw_tr = mk3(1, 1, 1);
                    tw_ret = ST_SCATTER_LIT; st = ST_WALK_START;                                // unconditional, :1088
                    break;
                }
                const bool emitter_hit = (bounces == 0 || specular) && h.lightIdx != -1;
                if (emitter_hit) {                                                              // :1103-1115
                    st = ST_EMIT_TR;
                    PT_POST_TR(medium, o, d, h0.t);
                    break;
                }
                if (h.matIdx == -1) {                                                           // medium boundary, :1117-1124
                    medium = dot(d, h.nor) > 0 ? h.mediumOutside : h.mediumInside;
                    o = h.pos;
                    st = ST_START;                                                              // `bounces--; continue;`
                    break;
                }
                Ld = mk3(0.f, 0.f, 0.f);
                if (is_delta(sc.mats[h.matIdx].type)) { st = ST_CONT; break; }
                {                                                                               // :1128-1150
                    float u = rng_next(rng);
                    float ua = rng_next(rng), ub = rng_next(rng);
                    seq_light_sample(sc, h.pos, u, ua, ub, ls, choicePdf);
                    nee_valid = !is_black(ls.radiance);
                    st = ST_NEE_DONE;
                    if (nee_valid) {
                        seq_eval_bsdf<MATS>(sc, h.matIdx, h.uv, -d, ls.dir, h.nor, h.dpdu, fr_l, samplePdf);
                        tw_o = h.pos; tw_d = ls.dir; tw_remain = ls.tmax; tw_medium = medium; tw_tr = mk3(1, 1, 1);
                        tw_ret = ST_NEE_DONE; st = ST_WALK_START;
                    }
                }
                break;
            }
            // ---- Tr() (src/pathtracer.cu:298-322), one segment per round
            case ST_WALK_START: {
                q_o = tw_o; q_d = tw_d; q_tmax = tw_remain;
                op = OP_TRAV; st = ST_WALK_HIT;
                break;
            }
            case ST_WALK_HIT: {
                tw_hit = r_hit;
                if (tw_hit && sc.shade[r_h.prim].matIdx != -1) { tw_tr = mk3(0, 0, 0); st = tw_ret; break; }
                tw_seg = tw_hit ? r_h.t : tw_remain;
                st = ST_WALK_TR;
                PT_POST_TR(tw_medium, tw_o, tw_d, tw_seg);
                break;
            }
            case ST_WALK_TR: {
                if (tw_medium >= 0) tw_tr *= r_trv;
                if (!tw_hit) { st = tw_ret; break; }
                const WShade& s = sc.shade[r_h.prim];
                f3 nor;
                if (s.type == 0) nor = normalize(lin3(1.f - r_h.b1 - r_h.b2, ld3(s.n1), r_h.b1, ld3(s.n2), r_h.b2, ld3(s.n3)));
                else nor = normalize((tw_o + tw_seg * tw_d) - ld3(s.n1));
                tw_medium = dot(tw_d, nor) > 0 ? s.mediumOutside : s.mediumInside;
                tw_remain -= tw_seg;
                tw_o = tw_o + tw_seg * tw_d;                                                    // Ray(ray(ray.tmax), ray.d, m, eps, tmax)
                st = ST_WALK_START;
                break;
            }
            case ST_SCATTER_LIT: {                                                              // :1089-1101
                const WMedium& M = sc.mediums[medium];
                const f3 tr = tw_tr;
                float phase = kInvFourPi;                                                       // Medium::Phase, src/medium.h:222
                if (M.g != 0) {
                    float costheta = dot(-d, ls.dir);
                    float cubicTerm = (1.f + M.g * M.g - 2.f * M.g * costheta);
                    phase = kInvFourPi * (1.f - M.g * M.g) / sqrtf(cubicTerm * cubicTerm * cubicTerm);
                }
                if (!is_black(ls.radiance)) Li += tr * beta * phase * ls.radiance / (ls.pdf * choicePdf);
                float pa = rng_next(rng), pb = rng_next(rng);                                   // Medium::SamplePhase, src/medium.h:197
                f3 dir;
                if (M.g == 0) { float pdf_; dir = uniform_sphere(pa, pb, pdf_); }
                else {
                    float costheta;
                    if (fabsf(M.g) < 1e-3f) costheta = 1.f - 2.f * pa;
                    else {
                        float sqrtTerm = (1.f - M.g * M.g) / (1.f - M.g + 2.f * M.g * pa);
                        costheta = (1.f + M.g * M.g - sqrtTerm * sqrtTerm) / (2.f * M.g);
                    }
                    float sintheta = sqrtf(1.f - costheta * costheta);
                    float phi = kTwoPi * pb;
                    float sinphi = sinf(phi), cosphi = cosf(phi);
                    dir = mk3(sintheta * cosphi, costheta, sintheta * sinphi);
                }
                o = o + sampledDist * d; d = dir;
                specular = false;
                PT_END_BOUNCE();
                break;
            }
            case ST_EMIT_TR: {
                const WLight& L = sc.lights[h.lightIdx];
                f3 le = dot(h.nor, -d) > 0.f ? ld3(L.radiance) : mk3(0.f, 0.f, 0.f);
                f3 tr = mk3(1.f, 1.f, 1.f);
                if (medium >= 0) tr = r_trv;
                Li += tr * beta * le;
                PT_FINISH();
                break;
            }
            case ST_NEE_DONE: {                                                                 // :1151-1160
                if (nee_valid) {
                    const f3 tr = tw_tr;
                    float weight = power_heuristic(1, ls.pdf * choicePdf, 1, samplePdf);
                    Ld += weight * tr * fr_l * ls.radiance * fabsf(dot(h.nor, ls.dir)) / (ls.pdf * choicePdf);
                }
                float s0 = rng_next(rng), s1 = rng_next(rng), s2 = rng_next(rng);
                seq_sample_bsdf<MATS>(sc, h.matIdx, h.uv, -d, h.nor, h.dpdu, mk3(s0, s1, s2), out_m, fr_m, pdf_m);
                st = ST_CONT;
                if (!(is_black(fr_m) || pdf_m == 0)) {
                    absdot = fabsf(dot(out_m, h.nor));
                    q_o = h.pos; q_d = out_m; q_tmax = INFINITY;
                    op = OP_TRAV; st = ST_MIS_HIT;
                }
                break;
            }
            case ST_MIS_HIT: {                                                                  // :1163-1207
                st = ST_CONT;
                if (r_hit) {
                    const int lightIdx = sc.shade[r_h.prim].lightIdx;
                    f3 p = h.pos + r_h.t * out_m;
                    f3 n = hit_normal(sc, p, r_h.prim, r_h.b1, r_h.b2);
                    if (lightIdx != -1) {
                        const WLight& L = sc.lights[lightIdx];
                        mis_rad = mk3(0.f, 0.f, 0.f);
                        if (dot(n, -out_m) > 0.f) mis_rad = ld3(L.radiance);
                        if (!is_black(mis_rad)) {
                            float pdfA = 1.f / L.area;
                            float cp = sc.cdf[lightIdx + 1] - sc.cdf[lightIdx];
                            float lenSquare = dot(p - h.pos, p - h.pos);
                            float costheta = fabsf(dot(n, out_m));
                            float lPdf = pdfA * lenSquare / (costheta);
                            mis_w = power_heuristic(1, pdf_m, 1, lPdf * cp);
                            st = ST_MIS_TR;
                            PT_POST_TR(medium, h.pos, out_m, r_h.t);
                        }
                    }
                } else if (sc.inf.isvalid) {
                    mis_rad = inf_le(sc.inf, out_m);
                    float cp = sc.cdf[sc.n_lights + 1] - sc.cdf[sc.n_lights];
                    mis_w = power_heuristic(1, pdf_m, 1, kInvFourPi * cp);
                    st = ST_MIS_TR;
                    PT_POST_TR(medium, h.pos, out_m, INFINITY);
                }
                break;
            }
            case ST_MIS_TR: {
                f3 tr = mk3(1.f, 1.f, 1.f);
                if (medium >= 0) tr = r_trv;
                Ld += mis_w * tr * fr_m * mis_rad * absdot / pdf_m;                               // :1185 / :1205
                st = ST_CONT;
                break;
            }
            default: {                                                                          // ST_CONT, :1209-1236
                const int mat_type = sc.mats[h.matIdx].type;
                if (!is_delta(mat_type)) {
#if defined(__CUDA_ARCH__)
                    Li = mk3(__fmaf_rn(beta.x, Ld.x, Li.x), __fmaf_rn(beta.y, Ld.y, Li.y), __fmaf_rn(beta.z, Ld.z, Li.z));
#else
                    Li += beta * Ld;
#endif
                }
                float c0 = rng_next(rng), c1 = rng_next(rng), c2 = rng_next(rng);                // :1213-1219
                f3 out, fr; float pdf;
                seq_sample_bsdf<MATS>(sc, h.matIdx, h.uv, -d, h.nor, h.dpdu, mk3(c0, c1, c2), out, fr, pdf);
                if (is_black(fr)) { PT_FINISH(); break; }
                beta *= fr * fabsf(dot(h.nor, out)) / pdf;
                specular = is_delta(mat_type);
                int m = dot(out, h.nor) > 0 ? h.mediumOutside : h.mediumInside;                  // :1224-1226
                m = dot(-d, h.nor) * dot(out, h.nor) > 0 ? medium : m;
                medium = m;
                o = h.pos; d = out;
                PT_END_BOUNCE();
                break;
            }
            }
        }
        // ================= vote =================
        const unsigned want_trav = __ballot_sync(0xffffffffu, op == OP_TRAV);
        const unsigned want_track = __ballot_sync(0xffffffffu, op == OP_TRACK);
        if (!(want_trav | want_track)) break;                                                   // every lane has run out of samples
        // ================= TRAV: closest hit for all lanes that wait for one (Intersect, :214-262) =================
        if (op == OP_TRAV) {
            r_hit = seq_closest_hit(sc, q_o, q_d, q_tmax, r_h);
            PT_STAT(0, 1);
            ++nrays;
            op = OP_NONE;
        }
        // ================= TRACK: a chunk of tracking steps for all lanes that wait for them =================
        if (op == OP_TRACK) {
            const WMedium& M = sc.mediums[q_medium];
            const WHetero& H = sc.het[q_medium];
            const float sigma = dot(ld3(M.sigmaT), kLum);
            const f3 p0 = ld3(H.p0);
            const f3 ext = ld3(H.p1) - p0;
            const float maxDensity = 1 / H.invMaxDensity;
            const float ce = 0.5f * maxDensity;
            int status = 0;                         // 1: left through `break` (interval / iteration budget), 2: collision
            PT_STAT(3, 1);
            for (int k = 0; k < kTrackChunk; ++k) {
                PT_STAT(2, 1);
                const float u = rng_next(rng);
                if (tk_kind == TK_RESIDUAL) tk_dist += -logf(u) * (1 / (maxDensity - ce) / sigma);
                else tk_dist += -logf(u) * H.invMaxDensity / sigma;
                if (tk_dist >= q_tmax) { status = 1; break; }
                f3 p = q_o + q_d * tk_dist;
                p = (p - p0) / ext;
                const float dens = het_density(H, p);
                if (tk_kind <= TK_DELTA) {                                                      // Sample (:137-157), Tr delta (:72-83)
                    if (dens * H.invMaxDensity > rng_next(rng)) { status = 2; break; }
                    if (--tk_iter == 0) { status = tk_kind == TK_DELTA ? 2 : 1; break; }
                } else {
                    if (tk_kind == TK_RATIO) tk_tr *= 1.f - dens * H.invMaxDensity;             // :85-101
                    else tk_tr *= 1.f - (dens - ce) / (maxDensity - ce);                        // :103-131
                    if (tk_tr < 0.1f) {
                        float q = 1.f - tk_tr;
                        if (rng_next(rng) < q) { status = 2; break; }
                        if (tk_kind == TK_RATIO) tk_tr = 1;
                        else tk_tr /= (1.f - q);
                    }
                    if (--tk_iter == 0) { status = 1; break; }
                }
            }
            if (status) {
                PT_STAT(1, 1);
                if (tk_kind == TK_SAMPLE) {
                    sampledDist = tk_dist;
                    sampledMedium = status == 2;
                    w3 = sampledMedium ? ld3(M.sigmaS) / ld3(M.sigmaT) : mk3(1.f, 1.f, 1.f);
                } else {
                    float tr = status == 2 ? 0.f : tk_tr;
                    if (tk_kind == TK_RESIDUAL && status == 1) tr *= expf(-q_tmax * ce * sigma);
                    r_trv = mk3(tr, tr, tr);
                }
                op = OP_NONE;
            }
        }
    }
#undef PT_POST_TR
#undef PT_FINISH
#undef PT_END_BOUNCE
    if (nrays) atomicAdd(&a.counters->rays, (unsigned long long)nrays);
}

}  // namespace pt
