// pt_math.cuh — device math of the B200 path tracer: vectors, RNG stream, camera, BSDF sampling/evaluation,
// emitter sampling, tone mapping.  Every function states which reference function it must agree with
// (brickray/gpu-pathtracer, paths relative to the reference root); parity is "same random numbers, same
// discrete decisions, same IEEE-float expression order" (SURVEY.md §7 hard parts), so expression shapes here
// deliberately keep the reference's association order.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PT_HD __host__ __device__ __forceinline__
#else
#define PT_HD inline
#endif

namespace pt {

constexpr float kPi = 3.14159265358f;            // src/common.h:22-27
constexpr float kTwoPi = 6.28318530716f;
constexpr float kFourPi = 12.56637061432f;
constexpr float kInvPi = 0.3183098861847f;
constexpr float kInvTwoPi = 0.1591549430923f;
constexpr float kInvFourPi = 0.0795774715461f;

struct f2 { float x, y; };
struct f3 { float x, y, z; };

PT_HD f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
PT_HD f3 mk3(float s) { return mk3(s, s, s); }
PT_HD f2 mk2(float x, float y) { f2 r; r.x = x; r.y = y; return r; }
PT_HD f3 ld3(const float* p) { return mk3(p[0], p[1], p[2]); }

// component-wise operators, same definitions as src/cutil_math.h (true division, no reciprocal tricks)
PT_HD f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
PT_HD f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
PT_HD f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
PT_HD f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
PT_HD f3 operator*(f3 a, float b) { return mk3(a.x * b, a.y * b, a.z * b); }
PT_HD f3 operator*(float b, f3 a) { return mk3(b * a.x, b * a.y, b * a.z); }
PT_HD f3 operator/(f3 a, f3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
PT_HD f3 operator/(f3 a, float b) { return mk3(a.x / b, a.y / b, a.z / b); }
PT_HD f3 operator+(f3 a, float b) { return mk3(a.x + b, a.y + b, a.z + b); }
PT_HD f3 operator-(f3 a, float b) { return mk3(a.x - b, a.y - b, a.z - b); }
PT_HD void operator+=(f3& a, f3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
PT_HD void operator*=(f3& a, f3 b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; }
PT_HD void operator*=(f3& a, float b) { a.x *= b; a.y *= b; a.z *= b; }
PT_HD void operator/=(f3& a, float b) { a.x /= b; a.y /= b; a.z /= b; }
PT_HD f2 operator-(f2 a, f2 b) { return mk2(a.x - b.x, a.y - b.y); }
PT_HD f2 operator+(f2 a, f2 b) { return mk2(a.x + b.x, a.y + b.y); }
PT_HD f2 operator*(f2 a, float b) { return mk2(a.x * b, a.y * b); }

// dot / cross: on the device the FMA contraction is spelled out in the form nvcc gives these expressions at (almost)
// every site of the reference's sm_100a build — dot = fma(z, fma(x, mul(y))), cross component = fma(first product,
// -mul(second)) — instead of leaving it to the inlining context (see dot_pinned below and DESIGN.md section 1).
PT_HD float dot(f3 a, f3 b) {                                                                              // cutil_math.h:1126
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a.z, b.z, __fmaf_rn(a.x, b.x, __fmul_rn(a.y, b.y)));
#else
    return a.x * b.x + a.y * b.y + a.z * b.z;
#endif
}
PT_HD f3 cross(f3 a, f3 b) {                                                                               // :1298
#if defined(__CUDA_ARCH__)
    return mk3(__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)), __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)),
               __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x)));
#else
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
#endif
}
// Linear combinations a*U + b*V (+ c*W) of vectors, contraction spelled out the way nvcc contracts a sum of products
// written in this order: the FIRST product is fused onto the plainly rounded SECOND one, a third is fused on top.
PT_HD f3 lin2(float a, f3 U, float b, f3 V) {
#if defined(__CUDA_ARCH__)
    return mk3(__fmaf_rn(a, U.x, __fmul_rn(b, V.x)), __fmaf_rn(a, U.y, __fmul_rn(b, V.y)), __fmaf_rn(a, U.z, __fmul_rn(b, V.z)));
#else
    return a * U + b * V;
#endif
}
PT_HD f3 lin3(float a, f3 U, float b, f3 V, float c, f3 W) {
#if defined(__CUDA_ARCH__)
    return mk3(__fmaf_rn(c, W.x, __fmaf_rn(a, U.x, __fmul_rn(b, V.x))), __fmaf_rn(c, W.y, __fmaf_rn(a, U.y, __fmul_rn(b, V.y))),
               __fmaf_rn(c, W.z, __fmaf_rn(a, U.z, __fmul_rn(b, V.z))));
#else
    return a * U + b * V + c * W;
#endif
}
// a*U + b*V + c*W contracted in SOURCE order — mul(a,U), then b*V fused on, then c*W fused on.  This is what the
// reference's sm_100a build does for the barycentric interpolation of the shading normal and uv in
// Triangle::Intersect's epilogue (src/mesh.h:87-88) at every site (Path closest-hit and MIS hit, Volpath closest-hit;
// read off the SASS with scripts/sass_expr.py: `fma(n3, b2, fma(n2, b1, mul(n1, 1-b1-b2)))`), unlike
// Triangle::SampleShape (src/mesh.h:102-103), where the SECOND product is the plain one (lin3 above).
PT_HD f3 lin3_seq(float a, f3 U, float b, f3 V, float c, f3 W) {
#if defined(__CUDA_ARCH__)
    return mk3(__fmaf_rn(c, W.x, __fmaf_rn(b, V.x, __fmul_rn(a, U.x))), __fmaf_rn(c, W.y, __fmaf_rn(b, V.y, __fmul_rn(a, U.y))),
               __fmaf_rn(c, W.z, __fmaf_rn(b, V.z, __fmul_rn(a, U.z))));
#else
    return a * U + b * V + c * W;
#endif
}
PT_HD f2 lin3_seq2(float a, f2 U, float b, f2 V, float c, f2 W) {
#if defined(__CUDA_ARCH__)
    return mk2(__fmaf_rn(c, W.x, __fmaf_rn(b, V.x, __fmul_rn(a, U.x))), __fmaf_rn(c, W.y, __fmaf_rn(b, V.y, __fmul_rn(a, U.y))));
#else
    return U * a + V * b + W * c;
#endif
}
// Hit/miss decisions must not depend on how the compiler happens to contract a*b+c in a given inlining context
// (nvcc's choice of WHICH product of a dot/cross gets fused changes with the surrounding code).  These two spell
// out the contraction the reference's own sm_100a build uses at every Triangle::Intersect site
// (verified in its PTX/SASS: dot = fma(z, fma(x, mul(y))), cross component = fma(first product, -mul(second))),
// with intrinsics that ptxas never re-fuses.  Host builds (oracle parity, no FMA) keep the plain expressions.
PT_HD float dot_pinned(f3 a, f3 b) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a.z, b.z, __fmaf_rn(a.x, b.x, __fmul_rn(a.y, b.y)));
#else
    return a.x * b.x + a.y * b.y + a.z * b.z;
#endif
}
PT_HD f3 cross_pinned(f3 a, f3 b) {
#if defined(__CUDA_ARCH__)
    return mk3(__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)), __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)),
               __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x)));
#else
    return cross(a, b);
#endif
}
PT_HD float rsqrt_ref(float x) {
#if defined(__CUDA_ARCH__)
    return rsqrtf(x);          // device normalize uses the hardware rsqrt approximation (cutil_math.h:1189)
#else
    return 1.0f / sqrtf(x);    // host definition (cutil_math.h:46-49)
#endif
}
PT_HD f3 normalize(f3 v) { float inv = rsqrt_ref(dot(v, v)); return v * inv; }                               // :1187
PT_HD float length(f3 v) { return sqrtf(dot(v, v)); }                                                        // :1169
PT_HD float clampf(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }                               // :1030
PT_HD bool is_black(f3 c) { return c.x == 0 && c.y == 0 && c.z == 0; }                                       // common.h:69
PT_HD bool is_nan3(f3 c) { return isnan(c.x) || isnan(c.y) || isnan(c.z); }
PT_HD bool is_inf3(f3 c) { return isinf(c.x) || isinf(c.y) || isinf(c.z); }
PT_HD float luminance(f3 c) { return dot(c, mk3(0.212671f, 0.715160f, 0.072169f)); }                         // pathtracer.cu:206
// Luminance of the throughput in the Russian-roulette test (`u < 1 - Y(beta)` decides whether the path lives):
// contraction pinned to what the reference's sm_100a build does at that site — Path: fma(z, fma(x, mul(y)));
// Volpath: fma(z, fma(y, mul(x))) (read off its PTX; the two kernels differ).
template <bool VOL> PT_HD float luminance_rr(f3 c) {
#if defined(__CUDA_ARCH__)
    if (VOL) return __fmaf_rn(c.z, 0.072169f, __fmaf_rn(c.y, 0.715160f, __fmul_rn(c.x, 0.212671f)));
    return __fmaf_rn(c.z, 0.072169f, __fmaf_rn(c.x, 0.212671f, __fmul_rn(c.y, 0.715160f)));
#else
    return luminance(c);
#endif
}

// ------------------------------------------------------------------------------------------------ RNG (a1)
// WangHash, src/pathtracer.cu:40
PT_HD uint32_t wang_hash(uint32_t seed) {
    seed = (seed ^ 61u) ^ (seed >> 16);
    seed = seed + (seed << 3);
    seed = seed ^ (seed >> 4);
    seed = seed * 0x27d4eb2du;
    seed = seed ^ (seed >> 15);
    return seed;
}
// thrust::default_random_engine = minstd_rand: x <- 48271 x mod (2^31 - 1); seeded as src/pathtracer.cu:888.
PT_HD uint32_t rng_seed(uint32_t pixel, uint32_t iter) {
    uint32_t s = wang_hash(pixel) + wang_hash(iter);
    s %= 2147483647u;
    return s == 0u ? 1u : s;
}
// thrust::uniform_real_distribution<float>(0,1): float(x - 1) / 2^31 — can return exactly 1.0f.
PT_HD float rng_next(uint32_t& x) {
    // x * 48271 mod (2^31 - 1) by Mersenne folding: p = hi * 2^31 + lo  ==>  p mod M = (hi + lo) mod M, and
    // hi + lo < 2M because p < 2^47.  Bit-identical to the 64-bit `%` (thrust's Schrage form gives the same value).
    const uint64_t p = (uint64_t)x * 48271ull;
    uint32_t r = (uint32_t)(p & 0x7fffffffull) + (uint32_t)(p >> 31);
    x = r >= 2147483647u ? r - 2147483647u : r;
    return (float)(x - 1u) * 4.656612873077393e-10f;   // exact power of two: identical to the division
}

// ------------------------------------------------------------------------------------------------ sampling (wrap.h)
PT_HD void make_coordinate(f3 n, f3& u, f3& w) {                                                            // wrap.h:6
    if (fabsf(n.x) > fabsf(n.y)) {
        float invLen = 1.0f / sqrtf(n.x * n.x + n.z * n.z);
        w = mk3(n.z * invLen, 0.0f, -n.x * invLen);
    } else {
        float invLen = 1.0f / sqrtf(n.y * n.y + n.z * n.z);
        w = mk3(0.0f, n.z * invLen, -n.y * invLen);
    }
    u = cross(w, n);
}
// ToWorld (wrap.h:18).  In the reference's Path kernel both inlined lambertian copies contract
// dir.x*u + dir.y*v + dir.z*w as fma(dir.z, w, fma(dir.y, v, mul(dir.x, u))) — first product plain (dir.x is itself
// a product, sin(theta)*cos(phi)) — unlike its dot products; measured: 82 % vs 71 % bit-identical pixels on Cornell.
PT_HD f3 to_world(f3 dir, f3 u, f3 v, f3 w) {
#if defined(__CUDA_ARCH__)
    return mk3(__fmaf_rn(dir.z, w.x, __fmaf_rn(dir.y, v.x, __fmul_rn(dir.x, u.x))),
               __fmaf_rn(dir.z, w.y, __fmaf_rn(dir.y, v.y, __fmul_rn(dir.x, u.y))),
               __fmaf_rn(dir.z, w.z, __fmaf_rn(dir.y, v.z, __fmul_rn(dir.x, u.z))));
#else
    return dir.x * u + dir.y * v + dir.z * w;
#endif
}
PT_HD f3 uniform_sphere(float u1, float u2, float& pdf) {                                                   // wrap.h:26
    float costheta = 1.f - 2.f * u1;
    float sintheta = sqrtf(1.f - costheta * costheta);
    float phi = kTwoPi * u2;
    float cosphi = cosf(phi);
    float sinphi = sinf(phi);
    pdf = kInvFourPi;
    return mk3(sintheta * cosphi, costheta, sintheta * sinphi);
}
PT_HD f3 cosine_hemisphere(float u1, float u2, float& pdf) {                                                // wrap.h:51
    float sintheta = sqrtf(u1);
    float costheta = sqrtf(1.f - u1);
    float phi = kTwoPi * u2;
    float cosphi = cosf(phi);
    float sinphi = sinf(phi);
    pdf = costheta * kInvPi;
    return mk3(sintheta * cosphi, costheta, sintheta * sinphi);
}
PT_HD f2 uniform_disk(float u1, float u2) {                                                                 // wrap.h:78
    float r = sqrtf(u1);
    float phi = kTwoPi * u2;
    return mk2(r * cosf(phi), r * sinf(phi));
}
PT_HD f2 uniform_triangle(float u1, float u2) {                                                             // wrap.h:110
    float su1 = sqrtf(u1);
    return mk2(1.f - su1, u2 * su1);
}

// ------------------------------------------------------------------------------------------------ camera (a2)
struct Camera {            // field-for-field the 104-B reference Camera (src/camera.h:8)
    float position[3], u[3], v[3], w[3];
    float resolution[2];
    float distance, fov, apertureRadius, focalDistance;
    uint8_t filmic, environment, _pad[2];
    int32_t medium;
    float width, height;
    float pixel2screen[2];
    float ratio, area;
};
// Camera::GeneratePrimaryRay, src/camera.h:48-84
PT_HD void camera_ray(const Camera& c, float x, float y, f2 xy, f3& orig, f3& dir) {
    f3 cu = ld3(c.u), cv = ld3(c.v), cw = ld3(c.w);
    orig = ld3(c.position);
    if (c.environment) {
        float theta = kPi * (1.f - y / c.resolution[1]);
        float phi = kTwoPi * (1.f - x / c.resolution[0]);
        f3 d = mk3(sinf(theta) * cosf(phi), cosf(theta), sinf(theta) * sinf(phi));
        dir = d.x * cu + d.y * cv - d.z * cw;
        return;
    }
    float xx = x * c.pixel2screen[0] - c.width;
    float yy = y * c.pixel2screen[1] - c.height;
    if (c.apertureRadius > 0.00001f) {
        f2 aperture_xy = xy * c.apertureRadius;
        float focal_x = c.ratio * xx;
        float focal_y = c.ratio * yy;
        f3 aperture = mk3(aperture_xy.x, aperture_xy.y, 0);
        f3 focal = mk3(focal_x, focal_y, -c.focalDistance);
        dir = focal - aperture;
        dir = lin3(dir.x, cu, dir.y, cv, dir.z, cw);
        orig += lin2(aperture.x, cu, aperture.y, cv);
    } else {
        dir = lin3(xx, cu, yy, cv, -c.distance, cw);
    }
    dir = normalize(dir);
}

// ------------------------------------------------------------------------------------------------ materials (a8-a10)
enum { MT_LAMBERTIAN = 0, MT_MIRROR, MT_DIELECTRIC, MT_ROUGHDIELECTRIC, MT_ROUGHCONDUCTOR, MT_SUBSTRATE };
struct Material {          // 72-B reference Material (src/material.h:19)
    int32_t type;
    float alphaU, alphaV, insideIOR, outsideIOR;
    float k[3], eta[3], diffuse[3], specular[3];
    int32_t textureIdx;
};
PT_HD bool is_delta(int type) { return type == MT_MIRROR || type == MT_DIELECTRIC; }                          // material.h:37

PT_HD float dielectric_fresnel(float cosi, float cost, float etai, float etat) {                            // pathtracer.cu:51
#if defined(__CUDA_ARCH__)
    // feeds the reflect/refract decision `u > fresnel`: contraction pinned to the reference build's SASS (the same at
    // every inlined copy in Path and Volpath, read with scripts/sass_expr.py): in numerator and denominator the FIRST
    // product is fused onto the plainly rounded second one — ptxas does that on top of the PTX, which still shows
    // separate mul / sub — and Rparl^2 is fused onto mul(Rperp^2)
    const float b = __fmul_rn(etai, cost), d = __fmul_rn(etat, cost);
    const float Rparl = __fdiv_rn(__fmaf_rn(etat, cosi, -b), __fmaf_rn(etat, cosi, b));
    const float Rperp = __fdiv_rn(__fmaf_rn(etai, cosi, -d), __fmaf_rn(etai, cosi, d));
    return __fmul_rn(__fmaf_rn(Rparl, Rparl, __fmul_rn(Rperp, Rperp)), 0.5f);
#else
    float Rparl = (etat * cosi - etai * cost) / (etat * cosi + etai * cost);
    float Rperp = (etai * cosi - etat * cost) / (etai * cosi + etat * cost);
    return (Rparl * Rparl + Rperp * Rperp) * 0.5f;
#endif
}
PT_HD f3 conduct_fresnel(float cosi, f3 eta, f3 k) {                                                        // pathtracer.cu:58
    f3 tmp = (eta * eta + k * k) * cosi * cosi;
    f3 Rparl2 = (tmp - eta * cosi * 2.f + 1.f) / (tmp + eta * cosi * 2.f + 1.f);
    f3 tmp_f = (eta * eta + k * k);
    f3 Rperp2 = (tmp_f - eta * cosi * 2.f + cosi * cosi) / (tmp_f + eta * cosi * 2.f + cosi * cosi);
    return (Rparl2 + Rperp2) * 0.5f;
}
PT_HD float ggx_d(f3 wh, f3 normal, f3 dpdu, float alphaU, float alphaV) {                                   // pathtracer.cu:68
    float costheta = dot(wh, normal);
    if (costheta <= 0.f) return 0.f;
    costheta = clampf(costheta, 0.f, 1.f);
    float costheta2 = costheta * costheta;
    float sintheta2 = 1.f - costheta2;
    float costheta4 = costheta2 * costheta2;
    float tantheta2 = sintheta2 / costheta2;
    f3 dir = normalize(wh - costheta * normal);
    float cosphi = dot(dir, dpdu);
    float cosphi2 = cosphi * cosphi;
    float sinphi2 = 1.f - cosphi2;
    float sqrD = 1.f + tantheta2 * (cosphi2 / (alphaU * alphaU) + sinphi2 / (alphaV * alphaV));
    return 1.f / (kPi * alphaU * alphaV * costheta4 * sqrD * sqrD);
}
PT_HD float smith_g(f3 w, f3 normal, f3 wh, f3 dpdu, float alphaU, float alphaV) {                           // pathtracer.cu:86
    float wdn = dot(w, normal);
    if (wdn * dot(w, wh) < 0.f) return 0.f;
    float sintheta = sqrtf(clampf(1.f - wdn * wdn, 0.f, 1.f));
    float tantheta = sintheta / wdn;
    if (isinf(tantheta)) return 0.f;
    f3 dir = normalize(w - wdn * normal);
    float cosphi = dot(dir, dpdu);
    float cosphi2 = cosphi * cosphi;
    float sinphi2 = 1.f - cosphi2;
    float alpha2 = cosphi2 * (alphaU * alphaU) + sinphi2 * (alphaV * alphaV);
    float sqrD = alpha2 * tantheta * tantheta;
    return 2.f / (1.f + sqrtf(1 + sqrD));
}
PT_HD float ggx_g(f3 wo, f3 wi, f3 normal, f3 wh, f3 dpdu, float aU, float aV) {                              // pathtracer.cu:103
    return smith_g(wo, normal, wh, dpdu, aU, aV) * smith_g(wi, normal, wh, dpdu, aU, aV);
}
PT_HD f3 sample_ggx(float alphaU, float alphaV, float u1, float u2) {                                        // pathtracer.cu:107
    if (alphaU == alphaV) {
        float costheta = sqrtf((1.f - u1) / (u1 * (alphaU * alphaV - 1.f) + 1.f));
        float sintheta = sqrtf(1.f - costheta * costheta);
        float phi = 2 * kPi * u2;
        float cosphi = cosf(phi);
        float sinphi = sinf(phi);
        return mk3(sintheta * cosphi, costheta, sintheta * sinphi);
    } else {
        float phi;
        if (u2 <= 0.25) phi = atanf(alphaV / alphaU * tanf(kTwoPi * u2));
        else if (u2 >= 0.75f) phi = atanf(alphaV / alphaU * tanf(kTwoPi * u2)) + kTwoPi;
        else phi = atanf(alphaV / alphaU * tanf(kTwoPi * u2)) + kPi;
        float sinphi = sinf(phi), cosphi = cosf(phi);
        float sinphi2 = sinphi * sinphi;
        float cosphi2 = 1.0f - sinphi2;
        float inverseA = 1.0f / (cosphi2 / (alphaU * alphaU) + sinphi2 / (alphaV * alphaV));
        float theta = atanf(sqrtf(inverseA * u1 / (1.0f - u1)));
        float sintheta = sinf(theta), costheta = cosf(theta);
        return mk3(sintheta * cosphi, costheta, sintheta * sinphi);
    }
}
PT_HD f3 reflect(f3 in, f3 nor) { return 2.f * dot(in, nor) * nor - in; }                                   // pathtracer.cu:140
PT_HD f3 refract(f3 in, f3 nor, float etai, float etat) {                                                   // pathtracer.cu:144
    float cosi = dot(in, nor);
    bool enter = cosi > 0;
    if (!enter) { float t = etai; etai = etat; etat = t; }
    float eta = etai / etat;
    float sini2 = 1.f - cosi * cosi;
    float sint2 = sini2 * eta * eta;
    float cost = sqrtf(1.f - sint2);
    return normalize((nor * cosi - in) * eta + (enter ? -cost : cost) * nor);
}
PT_HD f3 schlick_fresnel(f3 rs, float costheta) {                                                           // pathtracer.cu:160
    float c = 1.f - costheta;
    return rs + c * c * c * c * c * (mk3(1.f, 1.f, 1.f) - rs);
}
PT_HD float power_heuristic(int nf, float fPdf, int ng, float gPdf) {                                       // pathtracer.cu:166
    float f = nf * fPdf, g = ng * gPdf;
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(__fmul_rn(f, f), __fmaf_rn(g, g, __fmul_rn(f, f)));      // f*f is shared, g*g is fused (reference build)
#else
    return (f * f) / (f * f + g * g);
#endif
}
PT_HD bool same_hemisphere(f3 in, f3 out, f3 nor) { return dot(in, nor) * dot(out, nor) > 0 ? true : false; }  // :210

// SampleBSDF, src/pathtracer.cu:491-695 (TransportMode::Radiance). `albedo` = GetTexel(material, uv).xyz
// (constant material.diffuse when textureIdx == -1, src/pathtracer.cu:341-343).
PT_HD void sample_bsdf(const Material& m, f3 albedo, f3 in, f3 nor, f3 dpdu, f3 u, f3& out, f3& fr, float& pdf) {
    const f3 specular = ld3(m.specular);
    switch (m.type) {
    case MT_LAMBERTIAN: {
        f3 n = nor;
        if (dot(nor, in) < 0) n = -n;
        out = cosine_hemisphere(u.x, u.y, pdf);
        f3 uu = dpdu, ww;
        ww = cross(uu, n);
        out = to_world(out, uu, n, ww);
        fr = albedo * kInvPi;
        break;
    }
    case MT_MIRROR:
        out = reflect(in, nor);
        fr = specular / fabsf(dot(out, nor));
        pdf = 1.f;
        break;
    case MT_DIELECTRIC: {
        f3 wi = -in;
        f3 normal = nor;
        float ei = m.outsideIOR, et = m.insideIOR;
#if defined(__CUDA_ARCH__)
        // Contraction of this branch as the reference's sm_100a build has it in the copy of SampleBSDF that feeds the path
        // continuation (the only one a delta material reaches; identical in Path and Volpath — SASS read with
        // scripts/sass_expr.py): dot(in, nor) = (in.x*n.x (+) in.y*n.y) + in.z*n.z with the z product rounded separately;
        // Reflect = fma(n, 2*dot, -in); Refract = fma(fma(n, dot, -in), eta, n * (+-cost)) with 1 - sint2 contracted as
        // fma(-(sini2*eta), eta, 1), normalised with x*x as the plain product; |dot(out, n)| of fr = fma(z, fma(y, mul(x))).
        const float E = __fadd_rn(__fmaf_rn(in.x, nor.x, __fmul_rn(in.y, nor.y)), __fmul_rn(in.z, nor.z));
        const float cosi = -E;
        bool enter = cosi < 0;
        if (!enter) { float t = ei; ei = et; et = t; }
        float eta = __fdiv_rn(ei, et), cost;
        float sint2 = __fmul_rn(__fmul_rn(eta, eta), __fmaf_rn(-cosi, cosi, 1.f));
        cost = sqrtf(1.f - sint2 < 0.f ? 0.f : 1.f - sint2);
        const float E2 = __fadd_rn(E, E);
        f3 rdir = mk3(__fmaf_rn(nor.x, E2, wi.x), __fmaf_rn(nor.y, E2, wi.y), __fmaf_rn(nor.z, E2, wi.z));
        f3 tdir;
        {
            const bool enter_r = E > 0;
            const float etai = enter_r ? m.outsideIOR : m.insideIOR, etat = enter_r ? m.insideIOR : m.outsideIOR;
            const float eta_r = __fdiv_rn(etai, etat);
            const float sini2 = __fmaf_rn(-E, E, 1.f);
            const float cost_r = sqrtf(__fmaf_rn(__fmul_rn(sini2, eta_r), -eta_r, 1.f));
            const float sc = enter_r ? -cost_r : cost_r;
            const f3 tv = mk3(__fmaf_rn(__fmaf_rn(nor.x, E, wi.x), eta_r, __fmul_rn(nor.x, sc)),
                              __fmaf_rn(__fmaf_rn(nor.y, E, wi.y), eta_r, __fmul_rn(nor.y, sc)),
                              __fmaf_rn(__fmaf_rn(nor.z, E, wi.z), eta_r, __fmul_rn(nor.z, sc)));
            const float inv = rsqrtf(__fmaf_rn(tv.z, tv.z, __fmaf_rn(tv.y, tv.y, __fmul_rn(tv.x, tv.x))));
            tdir = mk3(__fmul_rn(tv.x, inv), __fmul_rn(tv.y, inv), __fmul_rn(tv.z, inv));
        }
#define PT_DOT_XYZ(a, b) __fmaf_rn((a).z, (b).z, __fmaf_rn((a).y, (b).y, __fmul_rn((a).x, (b).x)))
#else
        float cosi = dot(wi, normal);
        bool enter = cosi < 0;
        if (!enter) { float t = ei; ei = et; et = t; }
        float eta = ei / et, cost;
        float sint2 = eta * eta * (1.f - cosi * cosi);
        cost = sqrtf(1.f - sint2 < 0.f ? 0.f : 1.f - sint2);
        f3 rdir = reflect(-wi, normal);
        f3 tdir = refract(in, nor, m.outsideIOR, m.insideIOR);
#define PT_DOT_XYZ(a, b) dot(a, b)
#endif
        if (sint2 > 1.f) {  // total reflection
            out = rdir;
            fr = specular / fabsf(PT_DOT_XYZ(out, normal));
            pdf = 1.f;
            return;
        }
        float fresnel = dielectric_fresnel(fabsf(cost), fabsf(cosi), et, ei);
        if (u.x > fresnel) {  // refract
            out = tdir;
            fr = specular / fabsf(PT_DOT_XYZ(out, normal)) * (1.f - fresnel);
            fr *= eta * eta;
            pdf = 1.f - fresnel;
        } else {              // reflect
            out = rdir;
            fr = specular / fabsf(PT_DOT_XYZ(out, normal)) * fresnel;
            pdf = fresnel;
        }
#undef PT_DOT_XYZ
        break;
    }
    case MT_ROUGHCONDUCTOR: {
        f3 n = nor;
        if (dot(nor, in) < 0) n = -n;
        f3 wh = sample_ggx(m.alphaU, m.alphaV, u.x, u.y);
        f3 uu = dpdu, ww;
        ww = cross(uu, n);
        wh = to_world(wh, uu, n, ww);
        out = reflect(in, wh);
        if (!same_hemisphere(in, out, nor)) { fr = mk3(0, 0, 0); pdf = 0.f; return; }
        float cosi = dot(out, wh);
        f3 F = conduct_fresnel(fabsf(cosi), ld3(m.eta), ld3(m.k));
        float D = ggx_d(wh, n, dpdu, m.alphaU, m.alphaV);
        float G = ggx_g(in, out, n, wh, dpdu, m.alphaU, m.alphaV);
        fr = specular * F * D * G / (4.f * fabsf(dot(in, n)) * fabsf(dot(out, n)));
        pdf = D * fabsf(dot(wh, n)) / (4.f * fabsf(dot(in, wh)));
        break;
    }
    case MT_SUBSTRATE: {
        f3 n = nor;
        if (dot(nor, in) < 0) n = -n;
        if (u.x < 0.5) {
            float ux = u.x * 2.f;
            out = cosine_hemisphere(ux, u.y, pdf);
            f3 uu = dpdu, ww;
            ww = cross(uu, n);
            out = to_world(out, uu, n, ww);
        } else {
            float ux = (u.x - 0.5f) * 2.f;
            f3 wh = sample_ggx(m.alphaU, m.alphaV, ux, u.y);
            f3 uu = dpdu, ww;
            ww = cross(uu, n);
            wh = to_world(wh, uu, n, ww);
            out = reflect(in, wh);
        }
        if (!same_hemisphere(in, out, n)) { fr = mk3(0.f, 0.f, 0.f); pdf = 0.f; return; }
        float c0 = fabsf(dot(in, n));
        float c1 = fabsf(dot(out, n));
        f3 Rd = albedo;
        f3 Rs = specular;
        float cons0 = 1 - 0.5f * c0;
        float cons1 = 1 - 0.5f * c1;
        f3 diffuse = (28.f / (23.f * kPi)) * Rd * (mk3(1.f, 1.f, 1.f) - Rs) *
                     (1 - cons0 * cons0 * cons0 * cons0 * cons0) * (1 - cons1 * cons1 * cons1 * cons1 * cons1);
        f3 wh = normalize(in + out);
        float D = ggx_d(wh, n, dpdu, m.alphaU, m.alphaV);
        f3 spec = D / (4.f * fabsf(dot(out, wh)) * (c0 > c1 ? c0 : c1)) * schlick_fresnel(Rs, dot(out, wh));
        fr = diffuse + spec;
        pdf = 0.5f * (fabsf(dot(out, n)) * kInvPi + D * fabsf(dot(wh, n)) / (4.f * dot(in, wh)));
        break;
    }
    case MT_ROUGHDIELECTRIC: {
        f3 wi = -in;
        f3 n = nor;
        f3 wh = sample_ggx(m.alphaU, m.alphaV, u.x, u.y);
        f3 uu = dpdu, ww;
        ww = cross(uu, n);
        wh = to_world(wh, uu, n, ww);
        float ei = m.outsideIOR, et = m.insideIOR;
        float cosi = dot(wi, n);
        bool enter = cosi < 0;
        if (!enter) { float t = ei; ei = et; et = t; }
        float D = ggx_d(wh, n, dpdu, m.alphaU, m.alphaV);
        float eta = ei / et, cost;
        cosi = dot(wi, wh);
        float sint2 = eta * eta * (1.f - cosi * cosi);
        cost = sqrtf(1.f - sint2 < 0.f ? 0.f : 1.f - sint2);
        f3 rdir = reflect(-wi, wh);
        f3 tdir = normalize((wi - wh * cosi) * eta + (enter ? -cost : cost) * wh);
        if (sint2 > 1.f) {  // total reflection
            out = rdir;
            float G = ggx_g(in, out, n, wh, dpdu, m.alphaU, m.alphaV);
            fr = specular * D * G / (4.f * fabsf(dot(in, n)) * fabsf(dot(out, n)));
            pdf = D * fabsf(dot(wh, n)) / (4.f * fabsf(dot(wh, in)));
            return;
        }
        float fresnel = dielectric_fresnel(fabsf(cost), fabsf(cosi), et, ei);
        if (u.z > fresnel) {  // refract
            out = tdir;
            float G = ggx_g(in, out, n, wh, dpdu, m.alphaU, m.alphaV);
            float c = et * dot(out, wh) + ei * dot(in, wh);
            fr = specular * ei * ei * D * G * (1.f - fresnel) * fabsf(dot(in, wh)) * fabsf(dot(out, wh)) /
                 (fabsf(dot(out, n)) * fabsf(dot(in, n)) * c * c);
            fr *= (1.f / (eta * eta));
            pdf = (1.f - fresnel) * D * fabsf(dot(wh, n)) * et * et * fabsf(dot(out, wh)) / (c * c);
        } else {              // reflect
            out = rdir;
            float G = ggx_g(in, out, n, wh, dpdu, m.alphaU, m.alphaV);
            fr = specular * fresnel * D * G / (4.f * fabsf(dot(in, n)) * fabsf(dot(out, n)));
            pdf = D * fabsf(dot(wh, n)) / (4.f * fabsf(dot(wh, in))) * fresnel;
        }
        break;
    }
    default:
        fr = mk3(0, 0, 0); pdf = 0.f; out = mk3(0, 0, 0);
    }
}

// Fr, src/pathtracer.cu:698-826 (TransportMode::Radiance)
PT_HD void eval_bsdf(const Material& m, f3 albedo, f3 in, f3 out, f3 nor, f3 dpdu, f3& fr, float& pdf) {
    const f3 specular = ld3(m.specular);
    switch (m.type) {
    case MT_LAMBERTIAN:
        if (!same_hemisphere(in, out, nor)) { fr = mk3(0.f, 0.f, 0.f); pdf = 0.f; return; }
        fr = albedo * kInvPi;
        pdf = fabsf(dot(out, nor)) * kInvPi;
        break;
    case MT_MIRROR:
    case MT_DIELECTRIC:
        fr = mk3(0.f, 0.f, 0.f);
        pdf = 0.f;
        break;
    case MT_ROUGHCONDUCTOR: {
        if (!same_hemisphere(in, out, nor)) { fr = mk3(0, 0, 0); pdf = 0; return; }
        f3 n = nor;
        if (dot(nor, in) < 0) n = -n;
        f3 wh = normalize(in + out);
        float cosi = dot(out, wh);
        float D = ggx_d(wh, n, dpdu, m.alphaU, m.alphaV);
        float G = ggx_g(in, out, n, wh, dpdu, m.alphaU, m.alphaV);
        f3 F = conduct_fresnel(fabsf(cosi), ld3(m.eta), ld3(m.k));
        fr = specular * F * D * G / (4.f * fabsf(dot(in, n)) * fabsf(dot(out, n)));
        pdf = D * fabsf(dot(wh, n)) / (4.f * fabsf(dot(in, wh)));
        break;
    }
    case MT_SUBSTRATE: {
        if (!same_hemisphere(in, out, nor)) { fr = mk3(0, 0, 0); pdf = 0; return; }
        f3 n = nor;
        if (dot(nor, in) < 0) n = -n;
        float c0 = fabsf(dot(in, n));
        float c1 = fabsf(dot(out, n));
        f3 Rd = albedo;
        f3 Rs = specular;
        float cons0 = 1 - 0.5f * c0;
        float cons1 = 1 - 0.5f * c1;
        f3 wh = normalize(in + out);
        float D = ggx_d(wh, n, dpdu, m.alphaU, m.alphaV);
        f3 diffuse = (28.f / (23.f * kPi)) * Rd * (mk3(1.f, 1.f, 1.f) - Rs) *
                     (1 - cons0 * cons0 * cons0 * cons0 * cons0) * (1 - cons1 * cons1 * cons1 * cons1 * cons1);
        f3 spec = D / (4.f * fabsf(dot(out, wh)) * (c0 > c1 ? c0 : c1)) * schlick_fresnel(Rs, dot(out, wh));
        fr = diffuse + spec;
        pdf = 0.5f * (fabsf(dot(out, n)) * kInvPi + D * fabsf(dot(wh, n)) / (4.f * dot(in, wh)));
        break;
    }
    case MT_ROUGHDIELECTRIC: {
        f3 wi = -in;
        f3 n = nor;
        bool is_reflect = dot(in, n) * dot(out, n) > 0;
        float ei = m.outsideIOR, et = m.insideIOR;
        float cosi = dot(wi, n);
        bool enter = cosi < 0;
        if (!enter) { float t = ei; ei = et; et = t; }
        f3 wh = normalize(-(ei * in + et * out));
        float eta = ei / et, cost;
        cosi = dot(wi, wh);
        float sint2 = eta * eta * (1.f - cosi * cosi);
        cost = sqrtf(1.f - sint2 < 0.f ? 0.f : 1.f - sint2);
        float fresnel = dielectric_fresnel(fabsf(cost), fabsf(cosi), et, ei);
        float D = ggx_d(wh, n, dpdu, m.alphaU, m.alphaV);
        if (!is_reflect) {
            float G = ggx_g(in, out, n, wh, dpdu, m.alphaU, m.alphaV);
            float c = et * dot(out, wh) + ei * dot(in, wh);
            fr = specular * ei * ei * D * G * (1.f - fresnel) * fabsf(dot(in, wh)) * fabsf(dot(out, wh)) /
                 (fabsf(dot(out, n)) * fabsf(dot(in, n)) * c * c);
            fr *= (1.f / (eta * eta));
            pdf = (1.f - fresnel) * D * fabsf(dot(wh, n)) * et * et * fabsf(dot(out, wh)) / (c * c);
        } else {
            float G = ggx_g(in, out, n, wh, dpdu, m.alphaU, m.alphaV);
            fr = specular * fresnel * D * G / (4.f * fabsf(dot(in, n)) * fabsf(dot(out, n)));
            pdf = fresnel * D * fabsf(dot(wh, n)) / (4.f * fabsf(dot(wh, in)));
        }
        break;
    }
    default:
        fr = mk3(0, 0, 0); pdf = 0.f;
    }
}

// ------------------------------------------------------------------------------------------------ tone mapping (a17)
PT_HD f3 filmic_tonemap(f3 in) {                                                                            // pathtracer.cu:199
    f3 c = in - mk3(0.004f, 0.004f, 0.004f);
    c = mk3(fmaxf(0.f, c.x), fmaxf(0.f, c.y), fmaxf(0.f, c.z));
    c = (c * (6.2f * c + 0.5f)) / (c * (6.2f * c + 1.7f) + 0.06f);
    return c;
}
PT_HD float fast_pow(float x, float y) {
#if defined(__CUDA_ARCH__)
    return __powf(x, y);
#else
    return powf(x, y);     // host shim of the reference build maps __powf to powf
#endif
}
PT_HD f3 gamma_correct(f3 in) {                                                                             // pathtracer.cu:187
    float one_over_gamma = 1.f / 2.2f;
    float exposure = 1.41421356f;
    in = mk3(fmaxf(in.x, 1e-5f), fmaxf(in.y, 1e-5f), fmaxf(in.z, 1e-5f));
    in.x = fast_pow(in.x * exposure, one_over_gamma);
    in.y = fast_pow(in.y * exposure, one_over_gamma);
    in.z = fast_pow(in.z * exposure, one_over_gamma);
    return in;
}

}  // namespace pt
