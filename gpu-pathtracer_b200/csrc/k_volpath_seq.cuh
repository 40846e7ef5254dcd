// k_volpath_seq.cuh — `vpt` for scenes with HETEROGENEOUS media (Volpath, src/pathtracer.cu:1025-1242,
// with Heterogeneous::Sample / Tr, src/medium.h:52-179).
//
// Why not the wavefront: delta / ratio tracking draws a DATA-DEPENDENT number of random numbers from the path's one
// RNG stream in the middle of a bounce — Tr() of the shadow ray between the light sample and the BSDF samples, Tr() of
// the BSDF-sampled light ray between those and the continuation sample — and each of them needs a traversal first.
// k_shade draws a bounce's whole fixed sequence before any of its three rays is traced; with these media the order
// would have to be shade / trace / shade / trace / shade / trace per bounce.  Until that three-phase wavefront
// exists, such scenes run here: one thread = one path at a time, bounce by bounce in the reference's own order, with
//   * the same re-laid-out scene (64-B two-child nodes, 48-B intersection records, WShade / WLight),
//   * ordered stack traversal with the wavefront's slab / Moeller-Trumbore arithmetic and tie rule,
//   * per-THREAD regeneration: a thread whose path ends takes the next (iteration, pixel) from the global counter at
//     the top of the bounce loop, so a warp never idles behind its longest path (the reference's megakernel does),
//   * the same sample planes + ordered k_resolve behind it (NaN handling, accumulation, tonemap unchanged),
//   * the big loops (closest hit, tracking, Tr walk) and the shading calls as real functions: fully inlined the kernel
//     was instruction-fetch bound (profiles/r01z_het_seq.txt).
// A context with such a medium runs ONE lane (b200pt_api.cu).  A warp-cooperative variant (lanes post traversal /
// tracking operations and execute them together after a vote) was measured slower and is not in the tree.
// No warp collectives: the kernel also runs under the 1-lane CPU emulation of the test suite.
#pragma once
#include "k_shade.cuh"
#include "k_trace.cuh"

// The kernel is instruction-fetch bound when everything is inlined (closest hit at 3 call sites, tracking at 5: ncu
// shows `no_instruction` as the top stall, 7.7 warps per issue slot), so the three big loops are real functions.
#if defined(PT_SEQ_FORCEINLINE)
#define PT_SEQ_FN __device__ __forceinline__
#else
#define PT_SEQ_FN __device__ __noinline__ inline
#endif

namespace pt {

struct SeqArgs {
    SceneDev sc;
    float4* samples;          // [n_iters][n_local_pixels]
    Counters* counters;
    Camera cam;
    ShardMap map;
    BatchParams batch;
};

// Closest hit in [eps, tmax] (Intersect, src/pathtracer.cu:214-262): near child first, subtrees behind the current
// hit are dropped on pop; exact-t ties go to the higher primitive index like k_trace (the reference accepts tt == tmax).
PT_SEQ_FN bool seq_closest_hit(const SceneDev& sc, f3 o, f3 d, float tmax, Hit& out) {
    const f3 inv = mk3(1.f / d.x, 1.f / d.y, 1.f / d.z);
    float tn;
    if (!slab(sc.root_min[0], sc.root_min[1], sc.root_min[2], sc.root_max[0], sc.root_max[1], sc.root_max[2], o, inv, tmax, tn)) return false;
    int stack[64]; float stack_t[64];
    int sp = 0;
    int cur = sc.root_leaf_count > 0 ? ~0 : 0;
    int hprim = -1; float hb1 = 0.f, hb2 = 0.f;
    for (;;) {
        if (cur >= 0) {
            const float4* np = reinterpret_cast<const float4*>(sc.nodes + cur);
            const float4 q0 = np[0], q1 = np[1], q2 = np[2]; const int2 link = *reinterpret_cast<const int2*>(np + 3);
            float tl = 0.f, tr_ = 0.f;
            const bool hl = slab(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, o, inv, tmax, tl);
            const bool hr = link.y != kEmptyChild && slab(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, o, inv, tmax, tr_);
            int c0 = link.x, c1 = link.y;
            if (hl && hr) {
                if (tr_ < tl) { const int t_ = c0; c0 = c1; c1 = t_; const float f_ = tl; tl = tr_; tr_ = f_; }
                if (sp < 64) { stack[sp] = c1; stack_t[sp] = tr_; ++sp; }
                cur = c0;
                continue;
            }
            if (hl) { cur = c0; continue; }
            if (hr) { cur = c1; continue; }
        } else {
            int pi = ~cur;
            for (;;) {
                const float4* pp = reinterpret_cast<const float4*>(sc.prims + pi);
                const float4 p0 = pp[0], p1 = pp[1], p2 = pp[2];
                float t, b1, b2;
                const int acc = prim_test(p0, p1, p2, o, d, sc.eps, tmax, t, b1, b2);
                if (acc) {
                    if (acc == 2 || t < tmax || pi > hprim) { hprim = pi; hb1 = b1; hb2 = b2; }
                    tmax = t;
                }
                if (__float_as_int(p2.z) != 0) break;
                ++pi;
            }
        }
        cur = 0x7fffffff;
        while (sp > 0) { --sp; if (!(stack_t[sp] > tmax)) { cur = stack[sp]; break; } }
        if (cur == 0x7fffffff) break;
    }
    out.t = tmax; out.prim = hprim; out.b1 = hb1; out.b2 = hb2;
    return hprim >= 0;
}

// Heterogeneous::d / getDensity (src/medium.h:159-178): trilinear lookup, 0 outside the grid.
// The reference calls d(p + corner) eight times, each converting its own float coordinates to int and checking all six
// bounds; psi is integral, so (int)(psi + 0) == (int)psi and the eight lookups share two conversions and two bound
// checks per axis — same values (also for NaN / out-of-range inputs: every conversion the reference makes is still
// made on the same float), a third of the instructions (the lookup was ~25 % of the kernel's).
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
PT_SEQ_FN f3 het_tr(const WMedium& M, const WHetero& H, f3 o, f3 dir, float tmax, uint32_t& rng) {
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
__device__ __forceinline__ f3 seq_medium_tr(const SceneDev& sc, int medium, f3 o, f3 d, float tmax, uint32_t& rng) {
    const WMedium& M = sc.mediums[medium];
    if (M.type == 0) return exp3(ld3(M.sigmaT) * (-tmax));                                      // Homogeneous::Tr, src/medium.h:14
    return het_tr(M, sc.het[medium], o, d, tmax, rng);
}
// Homogeneous::Sample (src/medium.h:19-49) / Heterogeneous::Sample (:137-157)
PT_SEQ_FN f3 seq_medium_sample(const SceneDev& sc, int medium, f3 o, f3 dir, float tmax, uint32_t& rng, float& t, bool& sampled) {
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

// Tr() (src/pathtracer.cu:298-322): closest hits until an opaque surface blocks the ray; the transmittance of every
// segment comes from the medium the segment runs in, medium switch at invisible boundaries.
PT_SEQ_FN f3 seq_transmittance(const SceneDev& sc, f3 o, f3 d, float tmax, int medium, uint32_t& rng, uint32_t& nrays) {
    f3 tr = mk3(1, 1, 1);
    float remain = tmax;
    for (;;) {
        Hit h;
        const bool invisible = seq_closest_hit(sc, o, d, remain, h);
        ++nrays;
        if (invisible && sc.shade[h.prim].matIdx != -1) return mk3(0, 0, 0);
        const float seg = invisible ? h.t : remain;
        if (medium >= 0) tr *= seq_medium_tr(sc, medium, o, d, seg, rng);
        if (!invisible) break;
        const WShade& s = sc.shade[h.prim];
        f3 nor;
        if (s.type == 0) nor = normalize(lin3_seq(1.f - h.b1 - h.b2, ld3(s.n1), h.b1, ld3(s.n2), h.b2, ld3(s.n3)));
        else nor = normalize((o + seg * d) - ld3(s.n1));
        medium = dot(d, nor) > 0 ? s.mediumOutside : s.mediumInside;
        remain -= seg;
        o = o + seg * d;                                                                        // Ray(ray(ray.tmax), ray.d, m, eps, tmax)
    }
    return tr;
}

// Shading pieces used at several places of a bounce, as real functions for the same reason (instruction footprint).
PT_SEQ_FN void seq_light_sample(const SceneDev& sc, f3 pos, float u, float ua, float ub, LightSample& ls, float& choicePdf) {
    int idx = lookup_light(sc, u, choicePdf);
    if (idx < 0) idx = 0;
    if (idx != sc.n_lights) area_sample(sc.lights[idx], pos, ua, ub, sc.eps, ls);
    else inf_sample(sc.inf, ua, ub, sc.eps, ls);
}
template <uint32_t MATS>
PT_SEQ_FN void seq_sample_bsdf(const SceneDev& sc, int matIdx, f2 uv, f3 wo, f3 nor, f3 dpdu, f3 u, f3& out, f3& fr, float& pdf) {
    const Material mat = sc.mats[matIdx];
    const f3 albedo = material_albedo(sc, mat, uv);
    sample_bsdf_m<MATS>(mat, albedo, wo, nor, dpdu, u, out, fr, pdf);
}
template <uint32_t MATS>
PT_SEQ_FN void seq_eval_bsdf(const SceneDev& sc, int matIdx, f2 uv, f3 wo, f3 wi, f3 nor, f3 dpdu, f3& fr, float& pdf) {
    const Material mat = sc.mats[matIdx];
    const f3 albedo = material_albedo(sc, mat, uv);
    eval_bsdf_m<MATS>(mat, albedo, wo, wi, nor, dpdu, fr, pdf);
}

// 12 CTAs per SM (40 registers, ~870 B of spills into L1): the kernel waits on fixed-latency dependencies at 4 of 32
// active lanes, so resident warps pay more than spills cost.  Measured (smoke 1024^2 / shipped scene 512^2, Msamples/s):
// 4 CTAs (126 registers) 83 / 51 before the lookup change; then 5: 90 / 58, 6: 92 / 59, 8: 99 / 62, 10: 103 / 63,
// 12: 105 / 62, 16: 106 / 61.
template <uint32_t MATS>
__global__ void __launch_bounds__(128, 12) k_volpath_seq(const __grid_constant__ SeqArgs a) {
    const SceneDev& sc = a.sc;
    const uint32_t npix = (uint32_t)a.map.n_local_pixels;
    bool alive = false;
    uint32_t sample = 0u, rng = 0u, nrays = 0u;
    f3 o = mk3(0, 0, 0), d = mk3(0, 0, 1), beta = mk3(1, 1, 1), Li = mk3(0, 0, 0);
    int medium = -1, bounces = 0;
    bool specular = false;
    for (;;) {
        if (!alive) {                                                                           // ray generation, :1026-1048
            const unsigned long long s = atomicAdd(&a.counters->next_sample, 1ull);
            if (s >= a.batch.total) break;
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
            if (sc.max_depth <= 0) {                                                            // the bounce loop never runs
                st_pool(a.samples + sample, make_float4(0.f, 0.f, 0.f, 1.f));
                alive = false;
                continue;
            }
        }
        // ---- one iteration of the reference's bounce loop (:1050-1238)
        bool finished = false;
        bool count_bounce = true;
        Hit h0;
        const bool hit = seq_closest_hit(sc, o, d, INFINITY, h0);
        ++nrays;
        if (!hit) {
            if ((bounces == 0 || specular) && sc.inf.isvalid) Li += beta * inf_le(sc.inf, d);
            finished = true;
        } else {
            SurfaceHit h;
            reconstruct_hit(sc, o, d, h0.t, h0.prim, h0.b1, h0.b2, h);
            float sampledDist = 0.f; bool sampledMedium = false;
            if (medium >= 0) beta *= seq_medium_sample(sc, medium, o, d, h0.t, rng, sampledDist, sampledMedium);
            if (is_black(beta)) finished = true;                                                // :1070
            else if (sampledMedium) {                                                           // :1071-1101
                const WMedium& M = sc.mediums[medium];
                float u = rng_next(rng);
                float choicePdf;
                f3 samplePos = o + sampledDist * d;
                float ua = rng_next(rng), ub = rng_next(rng);
                LightSample ls;
                seq_light_sample(sc, samplePos, u, ua, ub, ls, choicePdf);
                f3 tr = seq_transmittance(sc, samplePos, ls.dir, ls.tmax, medium, rng, nrays);  // unconditional, :1088
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
                o = samplePos; d = dir;
                specular = false;
            } else {
                const bool emitter_hit = (bounces == 0 || specular) && h.lightIdx != -1;
                if (emitter_hit) {                                                              // :1103-1115
                    const WLight& L = sc.lights[h.lightIdx];
                    f3 le = dot(h.nor, -d) > 0.f ? ld3(L.radiance) : mk3(0.f, 0.f, 0.f);
                    f3 tr = mk3(1.f, 1.f, 1.f);
                    if (medium >= 0) tr = seq_medium_tr(sc, medium, o, d, h0.t, rng);
                    Li += tr * beta * le;
                    finished = true;
                } else if (h.matIdx == -1) {                                                    // medium boundary, :1117-1124
                    medium = dot(d, h.nor) > 0 ? h.mediumOutside : h.mediumInside;
                    o = h.pos;
                    count_bounce = false;                                                       // `bounces--; continue;`
                } else {
                    const int mat_type = sc.mats[h.matIdx].type;
                    const f3 wo = -d;
                    if (!is_delta(mat_type)) {                                                  // :1128-1211
                        float u = rng_next(rng);
                        float choicePdf;
                        float ua = rng_next(rng), ub = rng_next(rng);
                        LightSample ls;
                        seq_light_sample(sc, h.pos, u, ua, ub, ls, choicePdf);
                        f3 Ld = mk3(0.f, 0.f, 0.f);
                        if (!is_black(ls.radiance)) {
                            f3 fr; float samplePdf;
                            seq_eval_bsdf<MATS>(sc, h.matIdx, h.uv, wo, ls.dir, h.nor, h.dpdu, fr, samplePdf);
                            f3 tr = seq_transmittance(sc, h.pos, ls.dir, ls.tmax, medium, rng, nrays);
                            float weight = power_heuristic(1, ls.pdf * choicePdf, 1, samplePdf);
                            Ld += weight * tr * fr * ls.radiance * fabsf(dot(h.nor, ls.dir)) / (ls.pdf * choicePdf);
                        }
                        float s0 = rng_next(rng), s1 = rng_next(rng), s2 = rng_next(rng);
                        f3 out, fr; float pdf;
                        seq_sample_bsdf<MATS>(sc, h.matIdx, h.uv, wo, h.nor, h.dpdu, mk3(s0, s1, s2), out, fr, pdf);
                        if (!(is_black(fr) || pdf == 0)) {
                            const float absdot = fabsf(dot(out, h.nor));
                            Hit h1;
                            const bool lhit = seq_closest_hit(sc, h.pos, out, INFINITY, h1);
                            ++nrays;
                            if (lhit) {
                                const int lightIdx = sc.shade[h1.prim].lightIdx;
                                f3 p = h.pos + h1.t * out;
                                f3 n = hit_normal(sc, p, h1.prim, h1.b1, h1.b2);
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
                                        if (medium >= 0) tr = seq_medium_tr(sc, medium, h.pos, out, h1.t, rng);
                                        Ld += weight * tr * fr * radiance * absdot / pdf;        // :1185
                                    }
                                }
                            } else if (sc.inf.isvalid) {
                                f3 radiance = inf_le(sc.inf, out);
                                float cp = sc.cdf[sc.n_lights + 1] - sc.cdf[sc.n_lights];
                                float weight = power_heuristic(1, pdf, 1, kInvFourPi * cp);
                                f3 tr = mk3(1.f, 1.f, 1.f);
                                if (medium >= 0) tr = seq_medium_tr(sc, medium, h.pos, out, INFINITY, rng);
                                Ld += weight * tr * fr * radiance * absdot / pdf;                // :1205
                            }
                        }
#if defined(__CUDA_ARCH__)
                        Li = mk3(__fmaf_rn(beta.x, Ld.x, Li.x), __fmaf_rn(beta.y, Ld.y, Li.y), __fmaf_rn(beta.z, Ld.z, Li.z));
#else
                        Li += beta * Ld;
#endif
                    }
                    float c0 = rng_next(rng), c1 = rng_next(rng), c2 = rng_next(rng);            // :1213-1219
                    f3 out, fr; float pdf;
                    seq_sample_bsdf<MATS>(sc, h.matIdx, h.uv, wo, h.nor, h.dpdu, mk3(c0, c1, c2), out, fr, pdf);
                    if (is_black(fr)) finished = true;
                    else {
                        beta *= fr * fabsf(dot(h.nor, out)) / pdf;
                        specular = is_delta(mat_type);
                        int m = dot(out, h.nor) > 0 ? h.mediumOutside : h.mediumInside;          // :1224-1226
                        m = dot(-d, h.nor) * dot(out, h.nor) > 0 ? medium : m;
                        medium = m;
                        o = h.pos; d = out;
                    }
                }
            }
            if (!finished && count_bounce) {
                if (bounces > 3) {                                                              // :1230-1236
                    float illumate = clampf(1.f - luminance_rr<true>(beta), 0.f, 1.f);
                    if (rng_next(rng) < illumate) finished = true;
                    else beta /= (1 - illumate);
                }
                ++bounces;
                if (bounces >= sc.max_depth) finished = true;
            }
        }
        if (finished) {
            st_pool(a.samples + sample, make_float4(Li.x, Li.y, Li.z, 1.f));
            alive = false;
        }
    }
    if (nrays) atomicAdd(&a.counters->rays, (unsigned long long)nrays);
}

}  // namespace pt
