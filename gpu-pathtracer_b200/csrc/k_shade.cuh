// k_shade.cuh — the shading / next-event / regeneration stage of the wavefront (replaces the body of the
// `Path` and `Volpath` bounce loops, src/pathtracer.cu:904-1017 and :1050-1238, minus the ray traversals).
//
// Per slot and step:
//   A. finish the PREVIOUS bounce: add beta_old * Ld where Ld = light-sampled term (if the shadow ray was
//      unoccluded) + BSDF-sampled MIS term (needs the MIS ray's hit)                      (:942-995 / :1145-1211)
//   B. shade the continuation ray's hit: escape / emitter / delta or non-delta BSDF, draw the random numbers in
//      the reference's order (1 light pick, 2 light uv, 3 MIS-BSDF, 3 continuation-BSDF, +1 Russian roulette
//      when bounces > 3), emit the shadow, MIS and continuation rays                        (:905-1016)
//   C. when the sample ends, write its radiance to the sample plane and REGENERATE the slot with the next
//      (iteration, pixel) pair from the global counter (ray generation, :881-903)
#pragma once
#include "wavefront.cuh"
#ifdef B200PT_PROBE
#include <cstdio>
#endif

namespace pt {

struct ShadeArgs {
    SceneDev sc;
    Pool pool;
    Counters* counters;
    float4* samples;          // [n_iters][n_local_pixels] radiance of every finished sample (w = 1)
    RayQueue q;               // rays emitted by this step, consumed by the k_trace launch that follows
    uint32_t parity;          // QueueCtl set of this step
    Camera cam;
    ShardMap map;
    BatchParams batch;
    const FrameParams* frame;  // non-null inside a captured frame: camera and first_iter come from here
    int32_t drain_hint;        // the host saw the sample counter exhausted: dead tiles may leave early
    int32_t bin_materials;     // k_shade: thread j takes the j-th record of the CTA's tile in material order (see k_shade)
};

// One slot's record as the shade stage reads it (13 planes, 14 for vpt).
struct SlotRec { float4 df, orng, bs, lt4, h0, bo, pv, pl, pmd, pmf, h1, po, cy, pax; };

struct SurfaceHit { f3 pos, nor, dpdu; f2 uv; int matIdx, lightIdx, mediumInside, mediumOutside; };

// getTexel / GetTexel (src/pathtracer.cu:324-359): wrap + clamp bilinear lookup of a uchar4 texture.
__device__ __forceinline__ f3 tex_texel(const SceneDev& sc, const int4 info, int x, int y) {
    const float inv = 1.f / 255.f;
    const int w = info.y, h = info.z;
    float rx = x - (x / w) * w;
    float ry = y - (y / h) * h;
    x = (rx < 0) ? rx + w : rx;
    y = (ry < 0) ? ry + h : ry;
    if (x < 0) x = 0;
    if (x > w - 1) x = w - 1;
    if (y < 0) y = 0;
    if (y > h - 1) y = h - 1;
    const unsigned char* c = sc.texels + 4 * ((size_t)info.x + (size_t)y * w + x);
    return mk3(c[0] * inv, c[1] * inv, c[2] * inv);
}
__device__ __forceinline__ f3 material_albedo(const SceneDev& sc, const Material& m, f2 uv) {
    if (m.textureIdx == -1) return ld3(m.diffuse);
    const int4 info = sc.tex_info[m.textureIdx];
    float xx = info.y * uv.x;
    float yy = info.z * uv.y;
    int x = floorf(xx);
    int y = floorf(yy);
    float dx = fabsf(xx - x);
    float dy = fabsf(yy - y);
    f3 c00 = tex_texel(sc, info, x, y), c10 = tex_texel(sc, info, x + 1, y), c01 = tex_texel(sc, info, x, y + 1), c11 = tex_texel(sc, info, x + 1, y + 1);
    return lin2(1 - dy, lin2(1 - dx, c00, dx, c10), dy, lin2(1 - dx, c01, dx, c11));
}

// Epilogue of Triangle::Intersect (src/mesh.h:68-95) / Sphere::Intersect (src/sphere.h:74-91), evaluated once
// for the final hit from (t, prim, b1, b2).
__device__ __forceinline__ void reconstruct_hit(const SceneDev& sc, f3 o, f3 d, float t, int prim, float b1, float b2, SurfaceHit& h) {
    const WShade& s = sc.shade[prim];
#if defined(__CUDA_ARCH__)
    h.pos = mk3(__fmaf_rn(t, d.x, o.x), __fmaf_rn(t, d.y, o.y), __fmaf_rn(t, d.z, o.z));     // Ray::operator(), fused (as the reference build does)
#else
    h.pos = o + t * d;
#endif
    if (s.type == 0) {
        h.nor = normalize(lin3_seq(1.f - b1 - b2, ld3(s.n1), b1, ld3(s.n2), b2, ld3(s.n3)));
        h.uv = lin3_seq2(1.f - b1 - b2, mk2(s.uv1[0], s.uv1[1]), b1, mk2(s.uv2[0], s.uv2[1]), b2, mk2(s.uv3[0], s.uv3[1]));
        h.dpdu = normalize(cross(h.nor, ld3(s.ndpdv)));
    } else if (s.type == 2) {                                            // hair segment, src/line.h:74-83
        h.nor = -d;
        h.uv = mk2(b1, b2);
        f3 dpdv;
        make_coordinate(h.nor, h.dpdu, dpdv);
    } else {                                                             // sphere, src/sphere.h:74-91
        h.nor = normalize(h.pos - ld3(s.n1));
        const f3 normal = h.nor;
        float costheta = dot(normal, mk3(0.f, 1.f, 0.f));
        float v = acosf(costheta) * kInvPi;
        float cosphi = dot(mk3(1.f, 0.f, 0.f), mk3(normal.x, 0.f, normal.z));
        float phi = acosf(cosphi);
        phi = normal.z > 0.f ? kTwoPi - phi : phi;
        h.uv = mk2(phi * kInvTwoPi, v);
        h.dpdu = normalize(mk3(-kTwoPi * h.pos.y, kTwoPi * h.pos.x, 0));
    }
    h.matIdx = s.matIdx; h.lightIdx = s.lightIdx; h.mediumInside = s.mediumInside; h.mediumOutside = s.mediumOutside;
}
// shading normal only (MIS hit, :962)
__device__ __forceinline__ f3 hit_normal(const SceneDev& sc, f3 pos, int prim, float b1, float b2) {
    const WShade& s = sc.shade[prim];
    if (s.type == 0) return normalize(lin3_seq(1.f - b1 - b2, ld3(s.n1), b1, ld3(s.n2), b2, ld3(s.n3)));
    return normalize(pos - ld3(s.n1));     // sphere (hair segments never carry a light, so their normal is not needed here)
}

// Infinite::getTexel / getTexelBilinear (src/infinite.h:66-94)
__device__ __forceinline__ f3 inf_texel(const WInfinite& I, int x, int y) {
    int width = I.width, height = I.height;
    float rx = x - (x / width) * width;
    float ry = y - (y / height) * height;
    x = (rx < 0) ? rx + width : rx;
    y = (ry < 0) ? ry + height : ry;
    if (x < 0) x = 0;
    if (x > width - 1) x = width - 1;
    if (y < 0) y = 0;
    if (y > height - 1) y = height - 1;
    const float* c = I.data + 3 * ((size_t)y * width + x);
    return mk3(c[0], c[1], c[2]);
}
__device__ __forceinline__ f3 inf_bilinear(const WInfinite& I, f2 uv) {
    float xx = I.width * uv.x;
    float yy = I.height * uv.y;
    int x = floorf(xx);
    int y = floorf(yy);
    float dx = fabsf(xx - x);
    float dy = fabsf(yy - y);
    f3 c00 = inf_texel(I, x, y), c10 = inf_texel(I, x + 1, y), c01 = inf_texel(I, x, y + 1), c11 = inf_texel(I, x + 1, y + 1);
    return lin2(1 - dy, lin2(1 - dx, c00, dx, c10), dy, lin2(1 - dx, c01, dx, c11));
}
// direction -> lat-long uv, shared by Infinite::Le (:47) and Infinite::SampleLight (:17)
__device__ __forceinline__ f2 inf_dir_to_uv(const WInfinite& I, f3 dir) {
    f3 u = ld3(I.u), v = ld3(I.v), w = ld3(I.w);
    float costheta = dot(dir, v);
    float theta = acosf(costheta);
    f3 d = normalize(dir - costheta * v);
    float cosphi = dot(d, u);
    float phi = acosf(cosphi);
    float c = dot(d, w);
    phi = c > 0 ? kTwoPi - phi : phi;
    float uu = phi / kTwoPi;
    float vv = theta / kPi;
    return mk2(1.f - uu, vv);
}
__device__ __forceinline__ f3 inf_le(const WInfinite& I, f3 dir) { return inf_bilinear(I, inf_dir_to_uv(I, dir)); }

// LookUpLightDistribution (src/pathtracer.cu:172): first interval [cdf[i], cdf[i+1]] containing u.
__device__ __forceinline__ int lookup_light(const SceneDev& sc, float u, float& pdf) {
    const int n = sc.n_cdf - 1;
    if (n <= 8) {
        for (int i = 0; i < n; ++i) {
            float s = sc.cdf[i], e = sc.cdf[i + 1];
            if (u >= s && u <= e) { pdf = e - s; return i; }
        }
        pdf = 0.f; return -1;
    }
    // many emissive triangles: binary search for the first i with cdf[i+1] >= u — the same index the linear
    // scan returns, because the CDF is non-decreasing (cdf[i] < u for that i, or i == 0)
    int lo = 0, hi = n - 1;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (sc.cdf[mid + 1] >= u) hi = mid; else lo = mid + 1; }
    float s = sc.cdf[lo], e = sc.cdf[lo + 1];
    if (u >= s && u <= e) { pdf = e - s; return lo; }
    pdf = 0.f; return -1;
}

struct LightSample { f3 radiance, dir; float tmax, pdf; };
// Area::SampleLight (src/area.h:14) -> Triangle::SampleShape (src/mesh.h:100)
__device__ __forceinline__ void area_sample(const WLight& L, f3 pos, float ux, float uy, float eps, LightSample& ls) {
    f2 uv = uniform_triangle(ux, uy);
    f3 p = lin3(uv.x, ld3(L.v1), uv.y, ld3(L.v2), 1 - uv.x - uv.y, ld3(L.v3));
    f3 normal = normalize(lin3(uv.x, ld3(L.n1), uv.y, ld3(L.n2), 1 - uv.x - uv.y, ld3(L.n3)));
    f3 dir = p - pos;
    float pdf = 1.f / (L.area * fabsf(dot(normal, normalize(dir)))) * dot(dir, dir);
    if (dot(normal, dir) >= 0.f) pdf = 0.f;
    ls.pdf = pdf;
    ls.radiance = pdf != 0.f ? ld3(L.radiance) : mk3(0.f, 0.f, 0.f);
    ls.dir = normalize(dir);
    ls.tmax = sqrtf(dot(dir, dir) - eps);
}
// Infinite::SampleLight (src/infinite.h:17)
__device__ __forceinline__ void inf_sample(const WInfinite& I, float ux, float uy, float eps, LightSample& ls) {
    float pdfW;
    f3 dir = uniform_sphere(ux, uy, pdfW);
    f2 uv = inf_dir_to_uv(I, dir);
    ls.dir = dir;
    ls.tmax = 2.f * I.radius - eps;
    ls.pdf = pdfW;
    ls.radiance = inf_bilinear(I, uv);
}

__device__ __forceinline__ f3 exp3(f3 c) { return mk3(expf(c.x), expf(c.y), expf(c.z)); }

// Pool accesses are streaming (each record is read and written once per step).  Reads: the whole CTA tile of every
// pool array comes in through TMA bulk copies into shared memory, so ONE asynchronous round trip covers the record
// (per-thread loads were sunk by the compiler behind each dependent branch: ~6 serial DRAM round trips) and L1 is
// left to the scene records (WShade / Material / WLight / CDF).  Writes bypass L1 allocation.
#ifndef B200PT_EMULATE
__device__ __forceinline__ void st_pool(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
#else
static inline void st_pool(float4* p, float4 v) { *p = v; }
#endif

// ---- parity probes (diagnostic build only, -DB200PT_PROBE -> libb200pt_probe.so; scripts/parity_diag.py) ---------------
// printf of the per-bounce state of ONE (pixel, iteration) sample, in the format of the probe build of the reference
// (oracle/build_ref_debug.sh), so the first diverging intermediate of a flipped sample can be read off a diff.
#ifdef B200PT_PROBE
__device__ int g_probe[2] = {-1, -1};            // global pixel index, iteration
#define PT_P3(v) (double)(v).x, (double)(v).y, (double)(v).z
#define PT_PROBE(...) do { if (probe_on) printf(__VA_ARGS__); } while (0)
#else
#define PT_PROBE(...) do { } while (0)
#endif

// FUSED (k_wave.cuh): the record lives in shared memory, reached through a generic pointer.
template <bool FUSED> __device__ __forceinline__ void st_rec(float4* p, float4 v) {
    if (FUSED) *p = v; else st_pool(p, v);
}

// Material specialisation: a scene whose materials are all lambertian gets a kernel without the GGX / dielectric
// code (a quarter of the instructions and registers of the general one); MATS is the set of MaterialTypes present.
#ifndef PT_CULL_ZERO_SHADOW
#define PT_CULL_ZERO_SHADOW 1
#endif
constexpr uint32_t kMatsLambertOnly = 1u << MT_LAMBERTIAN;
constexpr uint32_t kMatsLDC = (1u << MT_LAMBERTIAN) | (1u << MT_DIELECTRIC) | (1u << MT_ROUGHCONDUCTOR);   // e.g. veach_bidir
constexpr uint32_t kMatsAll = 0x3fu;
constexpr uint32_t kShadeKeys = 18u;       // k_shade's sort keys: dead, miss, 2 + SceneDev::prim_key (0..15)

// SampleBSDF / Fr restricted to the material types in MATS: each enabled type is dispatched with a compile-time
// constant, so the switch inside sample_bsdf / eval_bsdf folds to that one case and the others are never emitted.
template <uint32_t MATS>
__device__ __forceinline__ void sample_bsdf_m(const Material& m, f3 albedo, f3 in, f3 nor, f3 dpdu, f3 u, f3& out, f3& fr, float& pdf) {
#define PT_CASE(T) if (((MATS >> (T)) & 1u) && m.type == (T)) { Material mm = m; mm.type = (T); sample_bsdf(mm, albedo, in, nor, dpdu, u, out, fr, pdf); return; }
    PT_CASE(MT_LAMBERTIAN) PT_CASE(MT_MIRROR) PT_CASE(MT_DIELECTRIC) PT_CASE(MT_ROUGHDIELECTRIC) PT_CASE(MT_ROUGHCONDUCTOR) PT_CASE(MT_SUBSTRATE)
#undef PT_CASE
    fr = mk3(0, 0, 0); pdf = 0.f; out = mk3(0, 0, 0);
}
template <uint32_t MATS>
__device__ __forceinline__ void eval_bsdf_m(const Material& m, f3 albedo, f3 in, f3 out, f3 nor, f3 dpdu, f3& fr, float& pdf) {
#define PT_CASE(T) if (((MATS >> (T)) & 1u) && m.type == (T)) { Material mm = m; mm.type = (T); eval_bsdf(mm, albedo, in, out, nor, dpdu, fr, pdf); return; }
    PT_CASE(MT_LAMBERTIAN) PT_CASE(MT_MIRROR) PT_CASE(MT_DIELECTRIC) PT_CASE(MT_ROUGHDIELECTRIC) PT_CASE(MT_ROUGHCONDUCTOR) PT_CASE(MT_SUBSTRATE)
#undef PT_CASE
    fr = mk3(0, 0, 0); pdf = 0.f;
}

// plain loads of one slot's record (emulation build; CTA-local wavefront, where the planes are in shared memory)
template <bool VOL> __device__ __forceinline__ void load_slot(const Pool& p, uint32_t slot, SlotRec& r) {
    r.df = p.d_flags[slot]; r.orng = p.o_rng[slot]; r.bs = p.beta_s[slot]; r.lt4 = p.li_t[slot];
    r.h0 = p.hit0[slot]; r.bo = p.beta_old[slot]; r.pv = p.vis[slot]; r.pl = p.ldl[slot];
    r.pmd = p.misd[slot]; r.pmf = p.misf[slot]; r.h1 = p.hit1[slot];
    r.po = p.pend_o[slot]; r.cy = p.carry[slot];
    r.pax = VOL ? p.aux[slot] : make_float4(0.f, 0.f, 0.f, 0.f);
}

// The shade stage of ONE slot (sections A-C above).  `slot` indexes the pool planes / queue entries, `gslot` is the
// slot's number among all `pool_n` slots of the wavefront (static sample hand-out); FUSED: called from the CTA-local
// wavefront kernel (k_wave.cuh) with the planes in shared memory.
template <bool VOL, uint32_t MATS, bool FUSED>
__device__ __forceinline__ void shade_slot(const ShadeArgs& a, const Pool& pool, const RayQueue& q, const uint32_t parity,
                                           uint32_t* cta_retired, uint32_t* cta_busy, const uint32_t slot, const uint32_t gslot,
                                           const uint32_t pool_n, const SlotRec& r, const unsigned long long next_snapshot) {
    const SceneDev& sc = a.sc;
    const uint32_t lane = pt_lane(), lt = (1u << lane) - 1u;
    const float4 df = r.df, orng = r.orng, bs = r.bs, lt4 = r.lt4, h0 = r.h0, bo = r.bo, pv = r.pv, pl = r.pl;
    const float4 pmd = r.pmd, pmf = r.pmf, h1 = r.h1, po = r.po, cy = r.cy, pax = r.pax;

    uint32_t flags = __float_as_uint(df.w);
    uint32_t rng = __float_as_uint(orng.w);
    f3 o = mk3(orng.x, orng.y, orng.z);
    f3 d = mk3(df.x, df.y, df.z);
    const bool alive = (flags & F_ALIVE) != 0;
    f3 beta = alive ? mk3(bs.x, bs.y, bs.z) : mk3(1, 1, 1);
    f3 Li = alive ? mk3(lt4.x, lt4.y, lt4.z) : mk3(0, 0, 0);
    uint32_t sample = alive ? __float_as_uint(bs.w) : 0u;
    uint32_t kdone = __float_as_uint(lt4.w);                        // static samples this slot has consumed
#ifdef B200PT_PROBE
    bool probe_on = false;
    if (alive) {
        const uint32_t npix_ = (uint32_t)a.map.n_local_pixels;
        const uint32_t itl_ = sample / npix_; uint32_t px_, py_;
        local_to_xy(a.map, sample - itl_ * npix_, px_, py_);
        probe_on = (int)(px_ + py_ * (uint32_t)a.map.width) == g_probe[0] && (int)(a.batch.first_iter + itl_) == g_probe[1];
    }
#endif
    // a dead slot regenerates while samples are left for it (its static share, then the global counter — a snapshot
    // of it is exact enough, see below)
    bool finished = !alive && (kdone < a.batch.k_static || next_snapshot < a.batch.total);
    const bool idle_dead = !alive && !finished;
    int bounces = (int)((flags >> kBounceShift) & 0x7fu);
    int medium = (int)((flags >> kMediumShift) & 0xffu) - 1;      // medium of the continuation ray (vpt)
    bool retire_carry = false;                                     // the slot's previous sample is completed by section A

    // ---------------------------------------------------------------- A. pending direct light of the previous bounce
    if (alive && (flags & F_PENDING)) {
        const f3 beta_old = mk3(bo.x, bo.y, bo.z);
        const f3 po3 = mk3(po.x, po.y, po.z);                          // where the shadow / MIS rays started
        const bool carried = (flags & F_CARRY) != 0;
        f3 Lacc = carried ? mk3(cy.x, cy.y, cy.z) : Li;                // the radiance this direct light belongs to
        const int medium2 = (int)((flags >> kMedium2Shift) & 0xffu) - 1;
        if (VOL && (flags & F_MEDSCATTER)) {
            // Li += tr*beta*phase*radiance / (lightPdf*choicePdf)   (src/pathtracer.cu:1092-1093)
            if (flags & F_SHADOW) {
                const float4 v = pv; const float4 l = pl; const float4 mf = pmf;
                f3 tr = mk3(v.x, v.y, v.z), radiance = mk3(l.x, l.y, l.z);
                Lacc += tr * beta_old * mf.x * radiance / mf.y;
            }
        } else {
            f3 Ld = mk3(0.f, 0.f, 0.f);
            if (flags & F_SHADOW) {
                const float4 v = pv; const float4 l = pl;
                if (!VOL) {
                    if (v.x != 0.f) Ld += mk3(l.x, l.y, l.z);                                   // :942-951
                } else {
                    // Ld += weight*tr*fr*radiance*|cos| / (lightPdf*choicePdf)                  // :1153-1154
                    const float4 ax = pax;                 // radiance xyz, weight
                    const float4 mf = pmf;
                    f3 tr = mk3(v.x, v.y, v.z), fr = mk3(l.x, l.y, l.z), radiance = mk3(ax.x, ax.y, ax.z);
                    Ld += ax.w * tr * fr * radiance * bo.w / mf.w;
                }
            }
            if (flags & F_MIS) {
                const float4 md = pmd; const float4 mf = pmf;
                const float absdot = pl.w;
                const f3 out = mk3(md.x, md.y, md.z), fr = mk3(mf.x, mf.y, mf.z);
                const float pdf = md.w;
                if (h1.x >= 0.f) {
                    const int prim = __float_as_int(h1.y);
                    const int lightIdx = sc.shade[prim].lightIdx;
                    f3 p = po3 + h1.x * out;
                    f3 n = hit_normal(sc, p, prim, h1.z, h1.w);
                    f3 radiance = mk3(0.f, 0.f, 0.f);
                    if (lightIdx != -1) {
                        const WLight& L = sc.lights[lightIdx];
                        if (dot(n, -out) > 0.f) radiance = ld3(L.radiance);                     // Area::Le, src/area.h:38
                        if (!is_black(radiance)) {
                            float pdfA = 1.f / L.area;                                          // Area::Pdf, src/area.h:28
                            float choicePdf = sc.cdf[lightIdx + 1] - sc.cdf[lightIdx];
                            float lenSquare = dot(p - po3, p - po3);
                            float costheta = fabsf(dot(n, out));
                            float lPdf = pdfA * lenSquare / (costheta);
                            float weight = power_heuristic(1, pdf, 1, lPdf * choicePdf);
                            if (!VOL) Ld += weight * fr * radiance * absdot / pdf;               // :975
                            else {
                                f3 tr = mk3(1.f, 1.f, 1.f);
                                if (medium2 >= 0) tr = exp3(ld3(sc.mediums[medium2].sigmaT) * (-h1.x));
                                Ld += weight * tr * fr * radiance * absdot / pdf;                // :1185
                            }
                        }
                    }
                } else if (sc.inf.isvalid) {
                    f3 radiance = inf_le(sc.inf, out);
                    float choicePdf = sc.cdf[sc.n_lights + 1] - sc.cdf[sc.n_lights];
                    float weight = power_heuristic(1, pdf, 1, kInvFourPi * choicePdf);
                    if (!VOL) Ld += weight * fr * radiance * absdot / pdf;                       // :989
                    else {
                        f3 tr = mk3(1.f, 1.f, 1.f);
                        if (medium2 >= 0) tr = exp3(ld3(sc.mediums[medium2].sigmaT) * (-INFINITY));
                        Ld += weight * tr * fr * radiance * absdot / pdf;                        // :1205
                    }
                }
            }
#if defined(__CUDA_ARCH__)
            Lacc = mk3(__fmaf_rn(beta_old.x, Ld.x, Lacc.x), __fmaf_rn(beta_old.y, Ld.y, Lacc.y), __fmaf_rn(beta_old.z, Ld.z, Lacc.z));   // :994, fused
#else
            Lacc += beta_old * Ld;                                                                // :994
#endif
            if (!carried) PT_PROBE("D Li %a %a %a Ld %a %a %a\n", PT_P3(Lacc), PT_P3(Ld));
        }
        if (carried) {
            st_rec<FUSED>(a.samples + __float_as_uint(po.w), make_float4(Lacc.x, Lacc.y, Lacc.z, 1.f));
            retire_carry = true;
        } else {
            Li = Lacc;
            if (flags & F_TERMINATE) finished = true;
        }
    }

    // ---------------------------------------------------------------- B. shade the continuation hit
    uint32_t nf = F_ALIVE;          // flags of the next step
    f3 new_o = o, new_d = d;
    f3 pend_origin = o;             // origin of the shadow / MIS rays emitted by this step
    bool specular = (flags & F_SPECULAR) != 0;
    if (alive && !finished) {
        if (h0.x < 0.f) {                                                                       // miss, :905-909
            if ((bounces == 0 || specular) && sc.inf.isvalid) Li += beta * inf_le(sc.inf, d);
            finished = true;
        } else {
            SurfaceHit h;
            reconstruct_hit(sc, o, d, h0.x, __float_as_int(h0.y), h0.z, h0.w, h);
            PT_PROBE("H %d o %a %a %a d %a %a %a t %a pos %a %a %a nor %a %a %a uv %a %a dpdu %a %a %a beta %a %a %a Li %a %a %a prim %d %d\n", bounces,
                     PT_P3(o), PT_P3(d), (double)h0.x, PT_P3(h.pos), PT_P3(h.nor), (double)h.uv.x, (double)h.uv.y, PT_P3(h.dpdu), PT_P3(beta), PT_P3(Li), h.matIdx, h.lightIdx);
            bool shade_surface = true;
            if (VOL) {
                float sampledDist = 0.f; bool sampledMedium = false;
                if (medium >= 0) {                                                              // Homogeneous::Sample, src/medium.h:19
                    const WMedium& M = sc.mediums[medium];
                    f3 sigmaT = ld3(M.sigmaT), sigmaS = ld3(M.sigmaS);
                    float sigma = dot(sigmaT, mk3(0.212671f, 0.715160f, 0.072169f));
                    float dist = -logf(rng_next(rng)) / sigma;
                    f3 Tr = exp3(sigmaT * -dist);
                    float pdf = sigma * expf(sigma * -dist);
                    sampledMedium = dist < h0.x;
                    sampledDist = dist;
                    beta *= sampledMedium ? (Tr * sigmaS / pdf) : sigmaT * Tr / pdf;
                }
                PT_PROBE("V beta %a %a %a dist %a med %d\n", PT_P3(beta), (double)sampledDist, (int)sampledMedium);
                if (is_black(beta)) { finished = true; shade_surface = false; }                 // :1070
                else if (sampledMedium) {                                                       // :1071-1101
                    shade_surface = false;
                    const WMedium& M = sc.mediums[medium];
                    float u = rng_next(rng);
                    float choicePdf;
                    int idx = lookup_light(sc, u, choicePdf);
                    if (idx < 0) idx = 0;
                    f3 samplePos = o + sampledDist * d;
                    float ua = rng_next(rng), ub = rng_next(rng);
                    LightSample ls;
                    if (idx != sc.n_lights) area_sample(sc.lights[idx], samplePos, ua, ub, sc.eps, ls);
                    else inf_sample(sc.inf, ua, ub, sc.eps, ls);
                    PT_PROBE("L lpdf %a cpdf %a sd %a %a %a tmax %a rad %a %a %a idx %d\n", (double)ls.pdf, (double)choicePdf, PT_P3(ls.dir), (double)ls.tmax, PT_P3(ls.radiance), idx);
                    float phase = kInvFourPi;                                                   // Medium::Phase, src/medium.h:222
                    if (M.g != 0) {
                        float costheta = dot(-d, ls.dir);
                        float cubicTerm = (1.f + M.g * M.g - 2.f * M.g * costheta);
                        phase = kInvFourPi * (1.f - M.g * M.g) / sqrtf(cubicTerm * cubicTerm * cubicTerm);
                    }
                    nf |= F_PENDING | F_MEDSCATTER;
                    pend_origin = samplePos;
                    st_rec<FUSED>(pool.beta_old + slot, make_float4(beta.x, beta.y, beta.z, 0.f));
                    if (!is_black(ls.radiance)) {                                               // Tr() is side-effect free otherwise
                        nf |= F_SHADOW;
                        st_rec<FUSED>(pool.shd + slot, make_float4(ls.dir.x, ls.dir.y, ls.dir.z, ls.tmax));
                        st_rec<FUSED>(pool.ldl + slot, make_float4(ls.radiance.x, ls.radiance.y, ls.radiance.z, 0.f));
                        st_rec<FUSED>(pool.misf + slot, make_float4(phase, ls.pdf * choicePdf, 0.f, 0.f));
                    }
                    float pa = rng_next(rng), pb = rng_next(rng);                               // Medium::SamplePhase, src/medium.h:197
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
                    new_o = samplePos; new_d = dir;
                    nf |= F_CONT | ((uint32_t)(medium + 1) << kMediumShift) | ((uint32_t)(medium + 1) << kMedium2Shift);
                    specular = false;
                    // Russian roulette + depth limit shared with the surface branch below
                    int b_old = bounces;
                    bounces = bounces + 1;
                    if (b_old > 3) {
                        float illumate = clampf(1.f - luminance_rr<VOL>(beta), 0.f, 1.f);
                        if (rng_next(rng) < illumate) nf = (nf | F_TERMINATE) & ~F_CONT;
                        else beta /= (1 - illumate);
                    }
                    if (bounces >= sc.max_depth) nf = (nf | F_TERMINATE) & ~F_CONT;
                }
            }
            if (shade_surface) {
                bool emitter_hit = (bounces == 0 || specular) && h.lightIdx != -1;
                if (emitter_hit) {                                                              // :917-922 / :1103-1115
                    const WLight& L = sc.lights[h.lightIdx];
                    f3 le = dot(h.nor, -d) > 0.f ? ld3(L.radiance) : mk3(0.f, 0.f, 0.f);
                    if (!VOL) Li += beta * le;
                    else {
                        f3 tr = mk3(1.f, 1.f, 1.f);
                        if (medium >= 0) tr = exp3(ld3(sc.mediums[medium].sigmaT) * (-h0.x));
                        Li += tr * beta * le;
                    }
                    finished = true;
                } else if (VOL && h.matIdx == -1) {                                             // medium boundary, :1117-1124
                    int m = dot(d, h.nor) > 0 ? h.mediumOutside : h.mediumInside;
                    new_o = h.pos; new_d = d;
                    nf |= F_CONT | ((uint32_t)(m + 1) << kMediumShift);
                    // bounces unchanged, no Russian roulette (the reference `continue`s)
                } else {
                    const Material mat = sc.mats[h.matIdx];
                    const f3 albedo = material_albedo(sc, mat, h.uv);                          // GetTexel, :341-359
                    const f3 wo = -d;
                    if (!is_delta(mat.type)) {                                                  // :925-995
                        float u = rng_next(rng);
                        float choicePdf;
                        int idx = lookup_light(sc, u, choicePdf);
                        if (idx < 0) idx = 0;      // u outside every interval (NaN): the reference falls off its loop (UB)
                        float ua = rng_next(rng), ub = rng_next(rng);
                        LightSample ls;
                        if (idx != sc.n_lights) area_sample(sc.lights[idx], h.pos, ua, ub, sc.eps, ls);
                        else inf_sample(sc.inf, ua, ub, sc.eps, ls);
                        PT_PROBE("L lpdf %a cpdf %a sd %a %a %a tmax %a rad %a %a %a idx %d\n", (double)ls.pdf, (double)choicePdf, PT_P3(ls.dir), (double)ls.tmax, PT_P3(ls.radiance), idx);
                        nf |= F_PENDING;
                        pend_origin = h.pos;
                        st_rec<FUSED>(pool.beta_old + slot, make_float4(beta.x, beta.y, beta.z, fabsf(dot(h.nor, ls.dir))));
                        float mis_absdot = 0.f;
                        f3 ldl = mk3(0, 0, 0);
                        if (!is_black(ls.radiance)) {
                            f3 fr; float samplePdf;
                            eval_bsdf_m<MATS>(mat, albedo, wo, ls.dir, h.nor, h.dpdu, fr, samplePdf);
                            float weight = power_heuristic(1, ls.pdf * choicePdf, 1, samplePdf);
                            nf |= F_SHADOW;
                            st_rec<FUSED>(pool.shd + slot, make_float4(ls.dir.x, ls.dir.y, ls.dir.z, ls.tmax));
                            if (!VOL) {
                                ldl = weight * fr * ls.radiance * fabsf(dot(h.nor, ls.dir)) / (ls.pdf * choicePdf);
                                // a light sample worth exactly zero (BSDF black: the light is below the surface's horizon) adds
                                // +0 whether it is visible or not — its shadow ray is not traced (a NaN term is not zero: traced).
                                // `pt` only: under `vpt` the term is formed after the walk (weight * Tr * fr * ...), the rule saved
                                // 0.1 % of C5's rays, and with it 5 samples of 2 M of the six-BSDF `vpt` scene took another path on
                                // the GPU (profiles/r02w_diag.txt; exact in emulation — most likely the added code moved an FMA
                                // contraction in that instantiation): dropped rather than chased
                                if (PT_CULL_ZERO_SHADOW && ldl.x == 0.f && ldl.y == 0.f && ldl.z == 0.f) nf &= ~F_SHADOW;
                            } else {
                                ldl = fr;
                                st_rec<FUSED>(pool.aux + slot, make_float4(ls.radiance.x, ls.radiance.y, ls.radiance.z, weight));
                            }
                        }
                        float s0 = rng_next(rng), s1 = rng_next(rng), s2 = rng_next(rng);
                        f3 out, fr; float pdf;
                        sample_bsdf_m<MATS>(mat, albedo, wo, h.nor, h.dpdu, mk3(s0, s1, s2), out, fr, pdf);
                        PT_PROBE("M out %a %a %a fr %a %a %a pdf %a\n", PT_P3(out), PT_P3(fr), (double)pdf);
                        float denom = ls.pdf * choicePdf;
                        if (!(is_black(fr) || pdf == 0) && mis_ray_may_reach_emitter(sc, h.pos, out)) {
                            nf |= F_MIS;
                            mis_absdot = fabsf(dot(out, h.nor));
                            st_rec<FUSED>(pool.misd + slot, make_float4(out.x, out.y, out.z, pdf));
                            st_rec<FUSED>(pool.misf + slot, make_float4(fr.x, fr.y, fr.z, denom));
                        } else if (VOL) {
                            st_rec<FUSED>(pool.misf + slot, make_float4(0.f, 0.f, 0.f, denom));
                        }
                        st_rec<FUSED>(pool.ldl + slot, make_float4(ldl.x, ldl.y, ldl.z, mis_absdot));
                        nf |= ((uint32_t)(medium + 1) << kMedium2Shift);
                    }
                    float c0 = rng_next(rng), c1 = rng_next(rng), c2 = rng_next(rng);            // :997-1003
                    f3 out, fr; float pdf;
                    sample_bsdf_m<MATS>(mat, albedo, wo, h.nor, h.dpdu, mk3(c0, c1, c2), out, fr, pdf);
                    PT_PROBE("C out %a %a %a fr %a %a %a pdf %a\n", PT_P3(out), PT_P3(fr), (double)pdf);
                    if (is_black(fr)) {
                        nf |= F_TERMINATE;
                    } else {
                        beta *= fr * fabsf(dot(h.nor, out)) / pdf;                              // :1005
                        specular = is_delta(mat.type);
                        int m = -1;
                        if (VOL) {                                                              // :1224-1226
                            m = dot(out, h.nor) > 0 ? h.mediumOutside : h.mediumInside;
                            m = dot(-d, h.nor) * dot(out, h.nor) > 0 ? medium : m;
                        }
                        new_o = h.pos; new_d = out;
                        nf |= F_CONT | ((uint32_t)(m + 1) << kMediumShift);
                        int b_old = bounces;
                        bounces = bounces + 1;
                        if (b_old > 3) {                                                        // :1010-1016
                            float illumate = clampf(1.f - luminance_rr<VOL>(beta), 0.f, 1.f);
                            if (rng_next(rng) < illumate) nf = (nf | F_TERMINATE) & ~F_CONT;
                            else beta /= (1 - illumate);
                        }
                        if (bounces >= sc.max_depth) nf = (nf | F_TERMINATE) & ~F_CONT;        // loop bound, :904
                    }
                    // the shadow/MIS rays of this bounce start at the hit point as well
                    if (!(nf & F_CONT)) new_o = h.pos;
                }
            }
            // a path that ends with nothing pending retires right away
            if (!finished && (nf & F_TERMINATE) && !(nf & F_PENDING)) finished = true;
        }
    }


    // ---------------------------------------------------------------- C. retire + regenerate
    // One aggregated atomic per warp hands out the next samples; the ray-queue reservation is issued right behind
    // it (a regenerated slot always emits exactly one continuation ray), so both round trips overlap.
    // A path that ends with direct light still pending does not idle for a step: if samples are left, the slot starts
    // its next sample now and carries the old one (F_CARRY) until the next pass has added the pending light.
    const bool have_more = kdone < a.batch.k_static || next_snapshot < a.batch.total;
    const bool emitted_pending = alive && !finished && (nf & F_PENDING) != 0u;
    const bool carry_now = emitted_pending && (nf & F_TERMINATE) != 0u && have_more;
    const bool want_new = finished || carry_now;
    if (finished && alive) {
        PT_PROBE("E Li %a %a %a\n", PT_P3(Li));
        st_rec<FUSED>(a.samples + sample, make_float4(Li.x, Li.y, Li.z, 1.f));
    }
    const bool take_static = want_new && kdone < a.batch.k_static;
    const uint32_t m_fin = __ballot_sync(kFullMask, want_new && !take_static);      // lanes that need the global counter
    const uint32_t m_ret = __ballot_sync(kFullMask, finished && alive);
    const uint32_t m_ret2 = __ballot_sync(kFullMask, retire_carry);
    uint32_t rays = want_new ? (F_CONT | (carry_now ? (nf & (F_SHADOW | F_MIS)) : 0u))
                             : (idle_dead ? 0u : (nf & (F_CONT | F_SHADOW | F_MIS)));
    const uint32_t mc = __ballot_sync(kFullMask, (rays & F_CONT) != 0u);
    const uint32_t ms = __ballot_sync(kFullMask, (rays & F_SHADOW) != 0u);
    const uint32_t mm = __ballot_sync(kFullMask, (rays & F_MIS) != 0u);
    const uint32_t nc = (uint32_t)__popc(mc), ns = (uint32_t)__popc(ms), nm = (uint32_t)__popc(mm);
    unsigned long long sbase = 0ull;
    uint32_t qbase = 0u;
    if (lane == 0u) {
        if (m_fin) sbase = atomicAdd(&a.counters->next_sample, (unsigned long long)__popc(m_fin));
        if (nc + ns + nm) qbase = atomicAdd(&q.ctl->tail[parity & 1u], nc + ns + nm);
        if (m_ret | m_ret2) {
            if (FUSED) atomicAdd(cta_retired, (uint32_t)(__popc(m_ret) + __popc(m_ret2)));
            else atomicAdd(&a.counters->done_samples, (unsigned long long)(__popc(m_ret) + __popc(m_ret2)));
        }
    }
    sbase = __shfl_sync(kFullMask, sbase, 0);
    qbase = __shfl_sync(kFullMask, qbase, 0);
    if (rays & F_CONT) q.entries[qbase + (uint32_t)__popc(mc & lt)] = slot;
    if (rays & F_SHADOW) q.entries[qbase + nc + (uint32_t)__popc(ms & lt)] = slot | (1u << kKindShift);
    if (rays & F_MIS) q.entries[qbase + nc + ns + (uint32_t)__popc(mm & lt)] = slot | (2u << kKindShift);
    if (idle_dead) return;
    if (FUSED) *cta_busy = 1u;                 // (benign race: every writer stores the same value)

    uint32_t carried_sample = 0u;
    if (want_new) {
        unsigned long long s;
        if (take_static) { s = (unsigned long long)gslot + (unsigned long long)kdone * (unsigned long long)pool_n; ++kdone; }
        else s = sbase + (unsigned long long)__popc(m_fin & lt);
        if (s >= a.batch.total) {
            // the batch ran out between the snapshot and the atomic (at most one step per batch); the reserved queue
            // entry stays and traces one harmless ray
            if (finished) {                                   // the slot dies
                st_rec<FUSED>(pool.o_rng + slot, make_float4(0.f, 0.f, 0.f, __uint_as_float(rng)));
                st_rec<FUSED>(pool.d_flags + slot, make_float4(0.f, 0.f, 1.f, __uint_as_float(0u)));
                return;
            }
            // carry_now: fall back to the idle step — the state computed by section B (TERMINATE | PENDING) is kept
        } else {
            uint32_t pending_bits = 0u;
            if (carry_now) {
                st_rec<FUSED>(pool.carry + slot, make_float4(Li.x, Li.y, Li.z, 0.f));
                carried_sample = sample;
                pending_bits = F_CARRY | (nf & (F_PENDING | F_SHADOW | F_MIS | F_MEDSCATTER)) | (nf & (0xffu << kMedium2Shift));
            }
            sample = (uint32_t)s;
            const uint32_t npix = (uint32_t)a.map.n_local_pixels;
            const uint32_t it_local = sample / npix, local = sample - it_local * npix;
            uint32_t x, y;
            local_to_xy(a.map, local, x, y);
            const uint32_t pixel = x + y * (uint32_t)a.map.width;                               // :883
            const Camera& cam = a.frame ? a.frame->cam : a.cam;
            rng = rng_seed(pixel, (a.frame ? a.frame->first_iter : a.batch.first_iter) + it_local);   // :888
            float offsetx = rng_next(rng) - 0.5f;                                               // :892-897
            float offsety = rng_next(rng) - 0.5f;
            float a0 = rng_next(rng), a1 = rng_next(rng);
            f2 aperture = mk2(0.f, 0.f);
            if (cam.apertureRadius > 0.00001f) aperture = uniform_disk(a0, a1);                  // unused otherwise (camera.h:63)
            camera_ray(cam, x + offsetx, y + offsety, aperture, new_o, new_d);
            beta = mk3(1.f, 1.f, 1.f); Li = mk3(0.f, 0.f, 0.f);
            bounces = 0; specular = false;
            int m = VOL ? cam.medium : -1;                                                      // :1043
            nf = F_ALIVE | F_CONT | ((uint32_t)(m + 1) << kMediumShift) | pending_bits;
        }
    }
    if (emitted_pending)
        st_rec<FUSED>(pool.pend_o + slot, make_float4(pend_origin.x, pend_origin.y, pend_origin.z, __uint_as_float(carried_sample)));
    if (specular) nf |= F_SPECULAR;
    nf |= ((uint32_t)bounces & 0x7fu) << kBounceShift;
    st_rec<FUSED>(pool.o_rng + slot, make_float4(new_o.x, new_o.y, new_o.z, __uint_as_float(rng)));
    st_rec<FUSED>(pool.d_flags + slot, make_float4(new_d.x, new_d.y, new_d.z, __uint_as_float(nf)));
    st_rec<FUSED>(pool.beta_s + slot, make_float4(beta.x, beta.y, beta.z, __uint_as_float(sample)));
    st_rec<FUSED>(pool.li_t + slot, make_float4(Li.x, Li.y, Li.z, __uint_as_float(kdone)));
}

// Resident CTAs per SM: the kernel is latency bound (long scoreboard), so occupancy beats registers — 8 CTAs (64
// registers; the lambertian kernel does not even spill) is +3 % on C2 and +6.5 % on C3 over the unconstrained 70 / 96
// registers; the volumetric instantiations carry one more staged plane and more live state: 6 CTAs (80 registers) is
// +3.4 % on C5 where 8 is -9 %; the all-materials kernel spills 144 B at 64 registers (-2 % on the textured-hair
// scene), so it stays at 6 as well (same-box A/B, profiles/r01z_perf_all_configs.txt).
template <bool VOL, uint32_t MATS>
__global__ void __launch_bounds__(128, (VOL || MATS == kMatsAll) ? 6 : 8) k_shade(const ShadeArgs a) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;      // the pool size is a multiple of the block size

    // ---------------------------------------------------------------- loads: the CTA's 128 pool records, all arrays,
    // staged into shared memory by TMA bulk copies (one elected thread issues them, everyone waits on the mbarrier)
    constexpr int kArrays = VOL ? 14 : 13;
#ifndef B200PT_EMULATE
    // drain phase (no sample left to hand out): a tile whose slots are all dead has nothing to do — find that out
    // with one 16-byte load per thread instead of staging 13 planes
    // (drain_hint comes from the host's last poll of the sample counter, so steps before the drain phase pay nothing)
    if (a.drain_hint) {
        const uint32_t f = __float_as_uint(a.pool.d_flags[slot].w);
        const uint32_t k = __float_as_uint(a.pool.li_t[slot].w);
        if (!__syncthreads_or((f & F_ALIVE) != 0u || k < a.batch.k_static)) return;
    }
    __shared__ __align__(128) float4 s_rec[kArrays][128];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const float4* src[14] = {a.pool.d_flags, a.pool.o_rng, a.pool.beta_s, a.pool.li_t, a.pool.hit0, a.pool.beta_old,
                                 a.pool.vis, a.pool.ldl, a.pool.misd, a.pool.misf, a.pool.hit1, a.pool.pend_o, a.pool.carry, a.pool.aux};
        mbar_expect_tx(&bar, (uint32_t)(kArrays * 128 * sizeof(float4)));
#pragma unroll
        for (int k = 0; k < kArrays; ++k) tma_bulk_g2s(&s_rec[k][0], src[k] + (size_t)blockIdx.x * 128, 128 * sizeof(float4), &bar);
    }
    __syncthreads();                       // the barrier object is initialised before anyone polls it
    mbar_wait(&bar, 0);
    // ---------------------------------------------------------------- material binning (north_star's "sorted by BSDF"):
    // a scene with several BSDFs runs every material's code in every warp, each with a fraction of the lanes, and the
    // instruction stream of the all-materials kernel does not fit the instruction cache (`no_instruction` stalls).  The
    // tile's 128 records are already in shared memory, so a counting sort costs three barriers: key = dead / miss /
    // material type of the hit primitive (+ emitter), one byte per primitive prepared by the host (SceneDev::prim_key —
    // a single gather, not the primitive -> material -> type chain), and thread j shades the j-th record in key order.
    // The slot a record belongs to — its pool address, sample and random-number stream — does not change; only which lane runs it.
    uint32_t t = threadIdx.x;
    if (MATS != kMatsLambertOnly && a.bin_materials) {
        __shared__ uint32_t s_cnt[kShadeKeys];
        __shared__ unsigned char s_order[128];
        if (threadIdx.x < kShadeKeys) s_cnt[threadIdx.x] = 0u;
        __syncthreads();
        const uint32_t f = __float_as_uint(s_rec[0][t].w);
        const float4 h0 = s_rec[4][t];
        const uint32_t prim = __float_as_uint(h0.y);
        uint32_t key = 0u;                                                               // dead: regenerate or idle
        if (f & F_ALIVE) key = (h0.x < 0.f || prim >= (uint32_t)a.sc.n_prims) ? 1u : 2u + a.sc.prim_key[prim];
        const uint32_t rank = atomicAdd(&s_cnt[key], 1u);
        __syncthreads();
        uint32_t base = 0u;
        for (uint32_t k = 0; k < key; ++k) base += s_cnt[k];
        s_order[base + rank] = (unsigned char)t;
        __syncthreads();
        t = s_order[threadIdx.x];
    }
    const uint32_t my = blockIdx.x * blockDim.x + t;
    SlotRec r;
    r.df = s_rec[0][t]; r.orng = s_rec[1][t]; r.bs = s_rec[2][t]; r.lt4 = s_rec[3][t];
    r.h0 = s_rec[4][t]; r.bo = s_rec[5][t]; r.pv = s_rec[6][t]; r.pl = s_rec[7][t];
    r.pmd = s_rec[8][t]; r.pmf = s_rec[9][t]; r.h1 = s_rec[10][t];
    r.po = s_rec[11][t]; r.cy = s_rec[12][t];
    r.pax = VOL ? s_rec[kArrays - 1][t] : make_float4(0.f, 0.f, 0.f, 0.f);
#else
    const uint32_t my = slot;
    SlotRec r;
    load_slot<VOL>(a.pool, my, r);
#endif
    shade_slot<VOL, MATS, false>(a, a.pool, a.q, a.parity, nullptr, nullptr, my, my, (uint32_t)a.pool.n, r, a.counters->next_sample);
}

// ---- Output (src/pathtracer.cu:2516-2531) over a whole batch of iterations -------------------------------
// Per pixel, in iteration order: NaN/Inf samples keep the previous iteration's colour (:1019-1020),
// acc += colour, and the last iteration's tonemapped acc/iter goes to `out`.
struct ResolveArgs {
    const float4* samples; float* acc; float* color; float* out;
    ShardMap map; BatchParams batch;
    int reset; int filmic; int write_out;
    const FrameParams* frame;  // non-null inside a captured frame: reset / filmic / out / first_iter come from here
};
__global__ void __launch_bounds__(256) k_resolve(const ResolveArgs a) {
    const uint32_t local = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t npix = (uint32_t)a.map.n_local_pixels;
    if (local >= npix) return;
    uint32_t x, y;
    local_to_xy(a.map, local, x, y);
    const size_t pixel = (size_t)x + (size_t)y * (size_t)a.map.width;
    f3 color = mk3(a.color[3 * pixel], a.color[3 * pixel + 1], a.color[3 * pixel + 2]);
    const int reset = a.frame ? a.frame->reset : a.reset, filmic = a.frame ? a.frame->filmic : a.filmic;
    const uint32_t first_iter = a.frame ? a.frame->first_iter : a.batch.first_iter;
    float* const out = a.frame ? a.frame->out : a.out;
    f3 acc = reset ? mk3(0.f, 0.f, 0.f) : mk3(a.acc[3 * pixel], a.acc[3 * pixel + 1], a.acc[3 * pixel + 2]);
    for (uint32_t k = 0; k < a.batch.n_iters; ++k) {
        const float4 s = a.samples[(size_t)k * npix + local];
        const f3 L = mk3(s.x, s.y, s.z);
        if (!is_inf3(L) && !is_nan3(L)) color = L;
        acc += color;
    }
    a.color[3 * pixel] = color.x; a.color[3 * pixel + 1] = color.y; a.color[3 * pixel + 2] = color.z;
    a.acc[3 * pixel] = acc.x; a.acc[3 * pixel + 1] = acc.y; a.acc[3 * pixel + 2] = acc.z;
    if (a.write_out) {
        const uint32_t iter = first_iter + a.batch.n_iters - 1;
        f3 c = acc / (float)(int)iter;
        c = filmic ? filmic_tonemap(c) : gamma_correct(c);
        out[3 * pixel] = c.x; out[3 * pixel + 1] = c.y; out[3 * pixel + 2] = c.z;
    }
}

// out = tonemap(acc / iter) for an externally reduced accumulation image (multi-GPU, after the NCCL reduce)
__global__ void k_tonemap(const float* acc, float* out, uint32_t npix, uint32_t iter, int filmic) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    f3 c = mk3(acc[3 * p], acc[3 * p + 1], acc[3 * p + 2]) / (float)(int)iter;
    c = filmic ? filmic_tonemap(c) : gamma_correct(c);
    out[3 * p] = c.x; out[3 * p + 1] = c.y; out[3 * p + 2] = c.z;
}

}  // namespace pt
