// TEST INFRASTRUCTURE — CPU oracle of the hot path: a restatement of the reference's unidirectional path
// tracer (brickray/gpu-pathtracer `Path` src/pathtracer.cu:880-1021, `Volpath` :1025-1242, `Output`
// :2516-2531 and everything they reach) as a one-sample-at-a-time loop over the reference's own scene arrays.
// It exists only to check the CUDA product (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg);
// the product never links, imports or calls it.
//
// Parity status: PINNED — tests/test_oracle_pinning.py checks this file bit-for-bit against the reference's
// own kernel bodies compiled for the host (oracle/_ref/libref_host.so, built from /root/reference by
// oracle/build_ref.sh) and against the committed golden vectors those produced (tests/golden/*.npz).
// Build: g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC (no FMA contraction: the pinning is bit-exact).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <omp.h>

#include "oracle_layouts.h"
#include "oracle_math.h"
#include "../include/b200pt.h"

using namespace orc;

namespace {

struct Ray { f3 o, d; float tmin, tmax; int medium; };          // src/ray.h:7 (medium as an index, -1 = none)
struct Isect {                                                   // src/intersection.h:6
    f3 pos, nor; f2 uv; f3 dpdu;
    int matIdx, bssrdf, lightIdx, mediumInside, mediumOutside;
};

struct SceneO {
    RefCamera cam;
    std::vector<RefPrimitive> prims;
    std::vector<RefLinearBVHNode> nodes;
    std::vector<RefMaterial> mats;
    std::vector<RefMedium> mediums;
    std::vector<std::vector<float>> densities;   // heterogeneous media: private copies of the nx*ny*nz grids
    std::vector<RefArea> lights;
    RefInfinite inf; bool has_inf = false;
    std::vector<float> inf_texels;
    std::vector<float> cdf;
    struct Tex { std::vector<unsigned char> rgba; int w, h; };
    std::vector<Tex> textures;
    int integrator = 1, max_depth = 5;
    float eps = 0.001f;
    unsigned w = 0, h = 0;
    std::vector<f3> acc, color;
};
SceneO* g = nullptr;

Ray mk_ray(f3 o, f3 d, int medium, float tmin, float tmax = INFINITY) { Ray r; r.o = o; r.d = d; r.medium = medium; r.tmin = tmin; r.tmax = tmax; return r; }
f3 at(const Ray& r, float t) { return r.o + t * r.d; }          // Ray::operator(), src/ray.h:29

// BBox::Intersect, src/bbox.h:77-96
bool bbox_intersect(const float* fmin, const float* fmax, const Ray& r) {
    f3 inv_dir = mk3(1.f / r.d.x, 1.f / r.d.y, 1.f / r.d.z);
    float t1 = (fmin[0] - r.o.x) * inv_dir.x;
    float t2 = (fmax[0] - r.o.x) * inv_dir.x;
    float t3 = (fmin[1] - r.o.y) * inv_dir.y;
    float t4 = (fmax[1] - r.o.y) * inv_dir.y;
    float t5 = (fmin[2] - r.o.z) * inv_dir.z;
    float t6 = (fmax[2] - r.o.z) * inv_dir.z;
    float tmin = fmaxf(fmaxf(fminf(t1, t2), fminf(t3, t4)), fminf(t5, t6));
    float tmax = fminf(fminf(fmaxf(t1, t2), fmaxf(t3, t4)), fmaxf(t5, t6));
    if (tmax <= 0.00001f) return false;
    if (tmin > tmax) return false;
    if (tmin > r.tmax) return false;
    return true;
}

// Triangle::Intersect, src/mesh.h:45-98
bool triangle_intersect(const RefTriangle& T, Ray& ray, Isect* isect) {
    f3 v1 = ld3(T.v1.v), v2 = ld3(T.v2.v), v3 = ld3(T.v3.v);
    f3 e1 = v2 - v1;
    f3 e2 = v3 - v1;
    f3 s1 = cross(ray.d, e2);
    float divisor = dot(s1, e1);
    if (fabsf(divisor) < 1e-8f) return false;
    float invDivisor = (float)(1.0 / (double)divisor);
    f3 s = ray.o - v1;
    float b1 = dot(s, s1) * invDivisor;
    if (b1 < 0.0 || b1 > 1.0) return false;
    f3 s2 = cross(s, e1);
    float b2 = dot(ray.d, s2) * invDivisor;
    if (b2 < 0.0 || b1 + b2 > 1.0) return false;
    float tt = dot(e2, s2) * invDivisor;
    if (tt < ray.tmin || tt > ray.tmax) return false;
    ray.tmax = tt;
    if (isect) {
        f3 dpdu, dpdv;
        f2 uv1 = mk2(T.v1.uv[0], T.v1.uv[1]), uv2 = mk2(T.v2.uv[0], T.v2.uv[1]), uv3 = mk2(T.v3.uv[0], T.v3.uv[1]);
        f2 duv1 = uv2 - uv1;
        f2 duv2 = uv3 - uv1;
        float det = duv1.x * duv2.y - duv1.y * duv2.x;
        if (fabs((double)det) < 1e-8) {
            f3 nn = normalize(cross(e1, e2));
            make_coordinate(nn, dpdu, dpdv);
        } else {
            float invDet = 1 / det;
            dpdu = (duv2.y * e1 - duv1.y * e2) * invDet;
            dpdv = (-duv2.x * e1 + duv1.x * e2) * invDet;
        }
        isect->pos = at(ray, tt);
        isect->nor = normalize(ld3(T.v1.n) * (1.f - b1 - b2) + ld3(T.v2.n) * b1 + ld3(T.v3.n) * b2);
        isect->uv = uv1 * (1.f - b1 - b2) + uv2 * b1 + uv3 * b2;
        isect->matIdx = T.matIdx;
        isect->lightIdx = T.lightIdx;
        isect->dpdu = normalize(cross(isect->nor, normalize(dpdv)));
        isect->bssrdf = T.bssrdfIdx;
        isect->mediumInside = T.mediumInside;
        isect->mediumOutside = T.mediumOutside;
    }
    return true;
}

// Sphere::Intersect, src/sphere.h:26-94
bool sphere_intersect(const RefSphere& S, Ray& ray, Isect* isect) {
    f3 origin = ld3(S.origin);
    f3 op = ray.o - origin;
    float B = dot(op, ray.d);
    float C = dot(op, op) - S.radius * S.radius;
    float delta = B * B - C;
    if (delta < 0.f) return false;
    float sqrDelta = sqrtf(delta);
    float t1 = -B - sqrDelta;
    float t2 = -B + sqrDelta;
    if (t1 < 0.f && t2 < 0.f) return false;
    if (t1 < 0.f || t2 < 0.f) {
        float tt1 = t1, tt2 = t2;
        t1 = tt1 < 0.f ? tt2 : tt1;
        t2 = tt1 < 0.f ? tt1 : tt2;
    } else {
        if (t1 > t2) { float temp = t2; t2 = t1; t1 = temp; }
    }
    if (t1 > ray.tmax) return false;
    if (t1 > ray.tmin) ray.tmax = t1;
    else if (t2 > 0.f) ray.tmax = t2;
    else return false;
    if (isect) {
        isect->pos = at(ray, ray.tmax);
        isect->nor = normalize(isect->pos - origin);
        f3 normal = isect->nor;
        float costheta = dot(normal, mk3(0.f, 1.f, 0.f));
        float v = acosf(costheta) * kInvPi;
        float cosphi = dot(mk3(1.f, 0.f, 0.f), mk3(normal.x, 0.f, normal.z));
        float phi = acosf(cosphi);
        phi = normal.z > 0.f ? kTwoPi - phi : phi;
        float u = phi * kInvTwoPi;
        isect->dpdu = normalize(mk3(-kTwoPi * isect->pos.y, kTwoPi * isect->pos.x, 0));
        isect->uv = mk2(u, v);
        isect->matIdx = S.matIdx;
        isect->lightIdx = -1;
        isect->bssrdf = S.bssrdfIdx;
        isect->mediumInside = S.mediumInside;
        isect->mediumOutside = S.mediumOutside;
    }
    return true;
}

// Line::Intersect, src/line.h:33-86 (hair segment: closest approach of ray and segment against the lerped radius)
bool line_intersect(const RefLine& L, Ray& ray, Isect* isect) {
    f3 u = ray.d;
    f3 v = ld3(L.p1) - ld3(L.p0);
    f3 w = ray.o - ld3(L.p0);
    float a = dot(u, u);
    float b = dot(u, v);
    float c = dot(v, v);
    float d = dot(u, w);
    float e = dot(v, w);
    float det = a * c - b * b;
    if (det == 0) return false;
    float t = (b * e - c * d) / det;
    float s = (a * e - b * d) / det;
    if (t < ray.tmin || t > ray.tmax) return false;
    s = clampf(s, 0.f, 1.f);
    f3 pr = ray.o + ray.d * t;
    f3 pl = ld3(L.p0) + (ld3(L.p1) - ld3(L.p0)) * s;
    f3 prl = pr - pl;
    float d2 = dot(prl, prl);
    float r = L.width0 * (1 - s) + L.width1 * s;
    if (d2 > r * r) return false;
    ray.tmax = t;
    if (isect) {
        isect->pos = at(ray, t);
        isect->nor = -ray.d;
        isect->uv = mk2(s, sqrtf(d2) / r);
        f3 dpdu, dpdv;
        make_coordinate(isect->nor, dpdu, dpdv);
        isect->dpdu = dpdu;
        isect->matIdx = L.matIdx;
        isect->lightIdx = -1;
        isect->bssrdf = -1;
        // mediumInside / mediumOutside are left untouched by the reference (only `pt` scenes use lines)
    }
    return true;
}

bool prim_intersect(const RefPrimitive& p, Ray& ray, Isect* isect) {
    if (p.type == REF_GT_TRIANGLE) return triangle_intersect(p.u.triangle, ray, isect);
    if (p.type == REF_GT_SPHERE) return sphere_intersect(p.u.sphere, ray, isect);
    return line_intersect(p.u.line, ray, isect);
}

// Traversal statistics of SURVEY 8(d) — rays, box tests and primitive tests AS EXECUTED BY THE REFERENCE TRAVERSAL (the
// R, N, P, H of the algorithmic-bytes formula bench.py uses).  Off by default; oracle_count(1) switches them on.
// [0] closest-hit queries, [1] any-hit queries, [2] box tests, [3] primitive tests.
static bool g_counting = false;
static unsigned long long g_counts[4];
#define ORACLE_COUNT(i, n) do { if (g_counting) { _Pragma("omp atomic") g_counts[i] += (n); } } while (0)

// Intersect (closest hit), src/pathtracer.cu:214-255: fixed left-then-right DFS over LinearBVHNode[]
bool intersect(Ray& ray, Isect* isect) {
    int stack[64]; int top = 0;
    bool ret = false;
    int node_idx = 0;
    unsigned nb = 0, np = 0;
    struct Flush { unsigned& b; unsigned& p; ~Flush() { ORACLE_COUNT(0, 1); ORACLE_COUNT(2, b); ORACLE_COUNT(3, p); } } flush{nb, np};
    for (;;) {
        const RefLinearBVHNode& node = g->nodes[node_idx];
        ++nb;
        if (bbox_intersect(node.fmin, node.fmax, ray)) {
            if (node.is_leaf) np += (unsigned)(node.end - node.start + 1);
            if (!node.is_leaf) {
                stack[top++] = node.second_child_offset;
                stack[top++] = node_idx + 1;
            } else {
                for (int i = node.start; i <= node.end; ++i)
                    if (prim_intersect(g->prims[i], ray, isect)) ret = true;
            }
        }
        if (top == 0) break;
        node_idx = stack[--top];
    }
    return ret;
}
// IntersectP (any hit), src/pathtracer.cu:257-296
bool intersect_p(Ray& ray) {
    int stack[64]; int top = 0;
    int node_idx = 0;
    unsigned nb = 0, np = 0;
    struct Flush { unsigned& b; unsigned& p; ~Flush() { ORACLE_COUNT(1, 1); ORACLE_COUNT(2, b); ORACLE_COUNT(3, p); } } flush{nb, np};
    for (;;) {
        const RefLinearBVHNode& node = g->nodes[node_idx];
        ++nb;
        if (bbox_intersect(node.fmin, node.fmax, ray)) {
            if (!node.is_leaf) {
                stack[top++] = node.second_child_offset;
                stack[top++] = node_idx + 1;
            } else {
                for (int i = node.start; i <= node.end; ++i) {
                    ++np;
                    if (prim_intersect(g->prims[i], ray, nullptr)) return true;
                }
            }
        }
        if (top == 0) break;
        node_idx = stack[--top];
    }
    return false;
}

// ---- emitters ------------------------------------------------------------------------------------------
float tri_area(const RefTriangle& T) {                           // Triangle::GetSurfaceArea, src/mesh.h:39
    f3 e1 = ld3(T.v2.v) - ld3(T.v1.v);
    f3 e2 = ld3(T.v3.v) - ld3(T.v1.v);
    return length(cross(e1, e2)) * 0.5f;
}
// Area::SampleLight, src/area.h:14-19 -> Triangle::SampleShape, src/mesh.h:100-109
void area_sample_light(const RefArea& A, f3 pos, f2 u, f3& rad, Ray& ray, f3& nor, float& pdf, float epsilon) {
    const RefTriangle& T = A.triangle;
    f2 uv = uniform_triangle(u.x, u.y);
    f3 p = uv.x * ld3(T.v1.v) + uv.y * ld3(T.v2.v) + (1 - uv.x - uv.y) * ld3(T.v3.v);
    f3 normal = normalize(uv.x * ld3(T.v1.n) + uv.y * ld3(T.v2.n) + (1 - uv.x - uv.y) * ld3(T.v3.n));
    f3 dir = p - pos;
    nor = normal;
    pdf = 1.f / (tri_area(T) * fabsf(dot(normal, normalize(dir)))) * dot(dir, dir);
    if (dot(normal, dir) >= 0.f) pdf = 0.f;
    rad = pdf != 0.f ? ld3(A.radiance) : mk3(0.f, 0.f, 0.f);
    ray = mk_ray(pos, normalize(dir), -1, epsilon, sqrtf(dot(dir, dir) - epsilon));
}
f3 area_le(const RefArea& A, f3 nor, f3 dir) {                   // Area::Le, src/area.h:38
    if (dot(nor, dir) > 0.f) return ld3(A.radiance);
    return mk3(0.f, 0.f, 0.f);
}
// Infinite::getTexel / getTexelBilinear, src/infinite.h:66-94
f3 inf_texel(int x, int y) {
    const RefInfinite& I = g->inf;
    int width = I.width, height = I.height;
    float rx = x - (x / width) * width;
    float ry = y - (y / height) * height;
    x = (rx < 0) ? rx + width : rx;
    y = (ry < 0) ? ry + height : ry;
    if (x < 0) x = 0;
    if (x > width - 1) x = width - 1;
    if (y < 0) y = 0;
    if (y > height - 1) y = height - 1;
    return ld3(&g->inf_texels[3 * ((size_t)y * width + x)]);
}
f3 inf_bilinear(f2 uv) {
    const RefInfinite& I = g->inf;
    float xx = I.width * uv.x;
    float yy = I.height * uv.y;
    int x = floor(xx);
    int y = floor(yy);
    float dx = fabs(xx - x);
    float dy = fabs(yy - y);
    f3 c00 = inf_texel(x, y), c10 = inf_texel(x + 1, y), c01 = inf_texel(x, y + 1), c11 = inf_texel(x + 1, y + 1);
    return (1 - dy) * ((1 - dx) * c00 + dx * c10) + dy * ((1 - dx) * c01 + dx * c11);
}
f2 inf_dir_to_uv(f3 dir) {                                       // shared body of Infinite::Le / SampleLight, src/infinite.h:21-30,48-57
    const RefInfinite& I = g->inf;
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
f3 inf_le(f3 dir) { return inf_bilinear(inf_dir_to_uv(dir)); }  // Infinite::Le, src/infinite.h:47
void inf_sample_light(f3 pos, f2 uniform, f3& rad, Ray& ray, f3& nor, float& pdf, float epsilon) {   // src/infinite.h:17-36
    float pdfW;
    f3 dir = uniform_sphere(uniform.x, uniform.y, pdfW);
    f2 uv = inf_dir_to_uv(dir);
    nor = -dir;
    ray = mk_ray(pos, dir, -1, epsilon, 2.f * g->inf.radius - epsilon);
    pdf = pdfW;
    rad = inf_bilinear(uv);
}
// LookUpLightDistribution, src/pathtracer.cu:172-181 (first interval containing u, inclusive both ends)
int lookup_light(float u, float& pdf) {
    int n = (int)g->cdf.size();
    for (int i = 0; i + 1 < n; ++i) {     // the reference also reads cdf[n] in its last iteration; never selected for u <= 1
        float s = g->cdf[i], e = g->cdf[i + 1];
        if (u >= s && u <= e) { pdf = e - s; return i; }
    }
    pdf = 0.f;
    return -1;
}
float light_choice_pdf(int idx) { return g->cdf[idx + 1] - g->cdf[idx]; }   // PdfFromLightDistribution, :183

// ---- media (Volpath) -------------------------------------------------------------------------------------
f3 exp3(f3 c) { return mk3(expf(c.x), expf(c.y), expf(c.z)); }   // Exp, src/common.h:81
f3 homogeneous_tr(const RefMedium& m, const Ray& ray) {          // Homogeneous::Tr, src/medium.h:14
    f3 c = ld3(m.sigmaT) * (-ray.tmax);
    return exp3(c);
}
// Homogeneous::Sample, src/medium.h:19-49 (the surface branch's weight is evaluated at the sampled distance)
f3 homogeneous_sample(const RefMedium& m, const Ray& ray, uint32_t& rng, float& t, bool& sampled) {
    f3 sigmaT = ld3(m.sigmaT), sigmaS = ld3(m.sigmaS);
    float sigma = dot(sigmaT, mk3(0.212671f, 0.715160f, 0.072169f));
    float dist = -logf(rng_next(rng)) / sigma;                     // Exponential, src/wrap.h:158
    f3 Tr = exp3(sigmaT * -dist);
    float pdf = sigma * expf(sigma * -dist);
    bool sampledMedium = dist < ray.tmax;
    sampled = sampledMedium;
    t = dist;
    return sampledMedium ? (Tr * sigmaS / pdf) : sigmaT * Tr / pdf;
}
// ---- Heterogeneous (src/medium.h:52-179): a density grid in the box p0..p1, sigmaT uniform across channels ----
float het_d(const RefMedium& m, f3 p) {                           // Heterogeneous::d, src/medium.h:173-178
    int x = p.x, y = p.y, z = p.z;
    if (x < 0 || x > m.nx - 1 || y < 0 || y > m.ny - 1 || z < 0 || z > m.nz - 1) return 0.f;
    int idx = z * m.ny * m.nx + y * m.nx + x;
    return m.density[idx];
}
float lerpf(float a, float b, float t) { return a + t * (b - a); }   // src/cutil_math.h:1008
float het_density(const RefMedium& m, f3 p) {                     // Heterogeneous::getDensity, src/medium.h:159-171
    f3 ps = mk3(p.x * m.nx, p.y * m.ny, p.z * m.nz);
    f3 psi = mk3(floorf(ps.x), floorf(ps.y), floorf(ps.z));
    f3 delta = ps - psi;
    float d00 = lerpf(het_d(m, psi), het_d(m, psi + mk3(1, 0, 0)), delta.x);
    float d10 = lerpf(het_d(m, psi + mk3(0, 1, 0)), het_d(m, psi + mk3(1, 1, 0)), delta.x);
    float d01 = lerpf(het_d(m, psi + mk3(0, 0, 1)), het_d(m, psi + mk3(1, 0, 1)), delta.x);
    float d11 = lerpf(het_d(m, psi + mk3(0, 1, 1)), het_d(m, psi + mk3(1, 1, 1)), delta.x);
    float d0 = lerpf(d00, d10, delta.y);
    float d1 = lerpf(d01, d11, delta.y);
    return lerpf(d0, d1, delta.z);
}
// Heterogeneous::Tr, src/medium.h:64-135: delta (0), ratio (1) or residual-ratio (2) tracking; draws from the path's RNG
f3 heterogeneous_tr(const RefMedium& m, const Ray& ray, uint32_t& rng) {
    float sigma = dot(ld3(m.sigmaT), mk3(0.212671f, 0.715160f, 0.072169f));
    f3 d = ld3(m.p1) - ld3(m.p0);
    float tr = 1.f;
    float dist = 0.f;
    int iter = m.iterMax;
    if (m.evalTransmittanceType == 0) {
        while (true) {
            dist += -logf(rng_next(rng)) * m.invMaxDensity / sigma;
            if (dist >= ray.tmax) break;
            f3 p = at(ray, dist);
            p = (p - ld3(m.p0)) / d;
            if (het_density(m, p) * m.invMaxDensity > rng_next(rng)) { tr = 0; break; }
            if (--iter == 0) { tr = 0; break; }
        }
    } else if (m.evalTransmittanceType == 1) {
        while (true) {
            dist += -logf(rng_next(rng)) * m.invMaxDensity / sigma;
            if (dist >= ray.tmax) break;
            f3 p = at(ray, dist);
            p = (p - ld3(m.p0)) / d;
            tr *= 1.f - het_density(m, p) * m.invMaxDensity;
            if (tr < 0.1f) {
                float q = 1.f - tr;
                if (rng_next(rng) < q) return mk3(0.f, 0.f, 0.f);
                tr = 1;
            }
            if (--iter == 0) break;
        }
    } else {
        float maxDensity = 1 / m.invMaxDensity;
        float ce = 0.5 * maxDensity;
        float tc = expf(-ray.tmax * ce * sigma);
        while (true) {
            dist += -logf(rng_next(rng)) * (1 / (maxDensity - ce) / sigma);
            if (dist >= ray.tmax) break;
            f3 p = at(ray, dist);
            p = (p - ld3(m.p0)) / d;
            tr *= 1.f - (het_density(m, p) - ce) / (maxDensity - ce);
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
// Heterogeneous::Sample, src/medium.h:137-157: delta tracking; a real collision scatters with weight sigmaS / sigmaT
f3 heterogeneous_sample(const RefMedium& m, const Ray& ray, uint32_t& rng, float& t, bool& sampled) {
    float sigma = dot(ld3(m.sigmaT), mk3(0.212671f, 0.715160f, 0.072169f));
    f3 d = ld3(m.p1) - ld3(m.p0);
    float dist = 0.f;
    int iter = m.iterMax;
    while (true) {
        dist += -logf(rng_next(rng)) * m.invMaxDensity / sigma;
        if (dist >= ray.tmax) break;
        f3 p = at(ray, dist);
        p = (p - ld3(m.p0)) / d;
        if (het_density(m, p) * m.invMaxDensity > rng_next(rng)) {
            t = dist;
            sampled = true;
            return ld3(m.sigmaS) / ld3(m.sigmaT);
        }
        if (--iter == 0) break;
    }
    t = dist;
    sampled = false;
    return mk3(1.f, 1.f, 1.f);
}
// the type dispatch at every call site (src/pathtracer.cu:308-311, :1065-1068, :1107-1110, :1180-1183, :1200-1203)
f3 medium_tr(const RefMedium& m, const Ray& ray, uint32_t& rng) {
    return m.type == REF_MEDIUM_HOMOGENEOUS ? homogeneous_tr(m, ray) : heterogeneous_tr(m, ray, rng);
}
f3 medium_sample(const RefMedium& m, const Ray& ray, uint32_t& rng, float& t, bool& sampled) {
    return m.type == REF_MEDIUM_HOMOGENEOUS ? homogeneous_sample(m, ray, rng, t, sampled) : heterogeneous_sample(m, ray, rng, t, sampled);
}
void medium_phase(const RefMedium& m, f3 in, f3 out, float& phase, float& pdf) {   // Medium::Phase, src/medium.h:222
    float gg = m.g;
    if (gg == 0) { phase = kInvFourPi; pdf = phase; return; }
    float costheta = dot(in, out);
    float cubicTerm = (1.f + gg * gg - 2.f * gg * costheta);
    phase = kInvFourPi * (1.f - gg * gg) / sqrtf(cubicTerm * cubicTerm * cubicTerm);
    pdf = phase;
}
void medium_sample_phase(const RefMedium& m, f2 u, f3& dir, float& phase, float& pdf) {   // Medium::SamplePhase, :197
    float gg = m.g;
    if (gg == 0) { phase = kInvFourPi; dir = uniform_sphere(u.x, u.y, pdf); return; }
    float costheta;
    if (fabs((double)gg) < 1e-3) costheta = 1.f - 2.f * u.x;
    else {
        float sqrtTerm = (1.f - gg * gg) / (1.f - gg + 2.f * gg * u.x);
        costheta = (1.f + gg * gg - sqrtTerm * sqrtTerm) / (2.f * gg);
    }
    float sintheta = sqrtf(1.f - costheta * costheta);
    float phi = kTwoPi * u.y;
    float sinphi = sinf(phi), cosphi = cosf(phi);
    dir = mk3(sintheta * cosphi, costheta, sintheta * sinphi);
    float cubicTerm = (1.f + gg * gg - 2.f * gg * costheta);
    phase = kInvFourPi * (1.f - gg * gg) / sqrtf(cubicTerm * cubicTerm * cubicTerm);
    pdf = phase;
}
// Tr, src/pathtracer.cu:298-322
f3 transmittance(Ray ray, uint32_t& rng) {
    f3 tr = mk3(1, 1, 1);
    float tmax = ray.tmax;
    while (true) {
        Isect isect; isect.lightIdx = -1; isect.matIdx = 0;
        bool invisible = intersect(ray, &isect);
        if (invisible && isect.matIdx != -1) return mk3(0, 0, 0);
        if (ray.medium >= 0) tr *= medium_tr(g->mediums[ray.medium], ray, rng);
        if (!invisible) break;
        int m = dot(ray.d, isect.nor) > 0 ? isect.mediumOutside : isect.mediumInside;
        tmax -= ray.tmax;
        ray = mk_ray(at(ray, ray.tmax), ray.d, m, g->eps, tmax);
    }
    return tr;
}

const RefMaterial& material(int idx) { return g->mats[idx]; }
// getTexel / GetTexel, src/pathtracer.cu:324-359: constant colour, or wrap+clamp bilinear lookup of a uchar4 texture
f3 texel_at(const SceneO::Tex& T, int x, int y) {
    const float inv = 1.f / 255.f;
    const int w = T.w, h = T.h;
    float rx = x - (x / w) * w;
    float ry = y - (y / h) * h;
    x = (rx < 0) ? rx + w : rx;
    y = (ry < 0) ? ry + h : ry;
    if (x < 0) x = 0;
    if (x > w - 1) x = w - 1;
    if (y < 0) y = 0;
    if (y > h - 1) y = h - 1;
    const unsigned char* c = &T.rgba[4 * ((size_t)y * w + x)];
    return mk3(c[0] * inv, c[1] * inv, c[2] * inv);
}
f3 albedo_of(const RefMaterial& m, f2 uv) {
    if (m.textureIdx == -1) return ld3(m.diffuse);
    const SceneO::Tex& T = g->textures[m.textureIdx];
    float xx = T.w * uv.x;
    float yy = T.h * uv.y;
    int x = floorf(xx);
    int y = floorf(yy);
    float dx = fabsf(xx - x);
    float dy = fabsf(yy - y);
    f3 c00 = texel_at(T, x, y), c10 = texel_at(T, x + 1, y), c01 = texel_at(T, x, y + 1), c11 = texel_at(T, x + 1, y + 1);
    return (1 - dy) * ((1 - dx) * c00 + dx * c10) + dy * ((1 - dx) * c01 + dx * c11);
}
const Material& M(const RefMaterial& m) { return *reinterpret_cast<const Material*>(&m); }
const Camera& C(const RefCamera& c) { return *reinterpret_cast<const Camera*>(&c); }

// The NEE + BSDF-sampled MIS block shared by Path (:924-995) and Volpath's surface branch (:1128-1211).
// `vol` adds the Tr() factors of Volpath.
f3 direct_light(const Ray& r, const Isect& isect, const RefMaterial& mat, uint32_t& rng, bool vol) {
    f3 pos = isect.pos, nor = isect.nor, dpdu = isect.dpdu;
    f3 Ld = mk3(0.f, 0.f, 0.f);
    bool inf = false;
    float u = rng_next(rng);
    float choicePdf;
    int idx = lookup_light(u, choicePdf);
    const int light_size = (int)g->lights.size();
    if (idx == light_size) inf = true;
    float ua = rng_next(rng), ub = rng_next(rng);
    f2 u1 = mk2(ua, ub);
    f3 radiance, lightNor;
    Ray shadowRay;
    float lightPdf;
    if (!inf) area_sample_light(g->lights[idx], pos, u1, radiance, shadowRay, lightNor, lightPdf, g->eps);
    else inf_sample_light(pos, u1, radiance, shadowRay, lightNor, lightPdf, g->eps);
    shadowRay.medium = r.medium;
    f3 wo = -r.d;
    if (!vol) {
        if (!is_black(radiance)) {
            Ray sr = shadowRay;
            if (!intersect_p(sr)) {
                f3 fr; float samplePdf;
                eval_bsdf(M(mat), albedo_of(mat, isect.uv), wo, shadowRay.d, nor, dpdu, fr, samplePdf);
                float weight = power_heuristic(1, lightPdf * choicePdf, 1, samplePdf);
                Ld += weight * fr * radiance * fabsf(dot(nor, shadowRay.d)) / (lightPdf * choicePdf);
            }
        }
    } else {
        if (!is_black(radiance)) {
            f3 fr; float samplePdf;
            eval_bsdf(M(mat), albedo_of(mat, isect.uv), wo, shadowRay.d, nor, dpdu, fr, samplePdf);
            f3 tr = transmittance(shadowRay, rng);
            float weight = power_heuristic(1, lightPdf * choicePdf, 1, samplePdf);
            Ld += weight * tr * fr * radiance * fabsf(dot(nor, shadowRay.d)) / (lightPdf * choicePdf);
        }
    }
    float s0 = rng_next(rng), s1 = rng_next(rng), s2 = rng_next(rng);   // device evaluates the argument list left to right
    f3 us = mk3(s0, s1, s2);
    f3 out, fr; float pdf;
    sample_bsdf(M(mat), albedo_of(mat, isect.uv), wo, nor, dpdu, us, out, fr, pdf);
    if (!(is_black(fr) || pdf == 0)) {
        Isect lightIsect; lightIsect.lightIdx = -1;
        Ray lightRay = mk_ray(pos, out, r.medium, g->eps);
        if (intersect(lightRay, &lightIsect)) {
            f3 p = lightIsect.pos;
            f3 n = lightIsect.nor;
            f3 rad = mk3(0.f, 0.f, 0.f);
            if (lightIsect.lightIdx != -1) rad = area_le(g->lights[lightIsect.lightIdx], n, -lightRay.d);
            if (!is_black(rad)) {
                float pdfA = 1.f / tri_area(g->lights[lightIsect.lightIdx].triangle);   // Area::Pdf, src/area.h:28
                float cp = light_choice_pdf(lightIsect.lightIdx);
                float lenSquare = dot(p - pos, p - pos);
                float costheta = fabsf(dot(n, lightRay.d));
                float lPdf = pdfA * lenSquare / (costheta);
                float weight = power_heuristic(1, pdf, 1, lPdf * cp);
                if (!vol) Ld += weight * fr * rad * fabsf(dot(out, nor)) / pdf;
                else {
                    f3 tr = mk3(1.f, 1.f, 1.f);
                    if (lightRay.medium >= 0) tr = medium_tr(g->mediums[lightRay.medium], lightRay, rng);
                    Ld += weight * tr * fr * rad * fabsf(dot(out, nor)) / pdf;
                }
            }
        } else if (g->has_inf && g->inf.isvalid) {
            f3 rad = inf_le(lightRay.d);
            float cp = light_choice_pdf(light_size);
            float lightPdf2 = kInvFourPi;                            // Infinite::Pdf, src/infinite.h:38
            float weight = power_heuristic(1, pdf, 1, lightPdf2 * cp);
            if (!vol) Ld += weight * fr * rad * fabsf(dot(out, nor)) / pdf;
            else {
                f3 tr = mk3(1.f, 1.f, 1.f);
                if (lightRay.medium >= 0) tr = medium_tr(g->mediums[lightRay.medium], lightRay, rng);
                Ld += weight * tr * fr * rad * fabsf(dot(out, nor)) / pdf;
            }
        }
    }
    return Ld;
}

Ray primary_ray(unsigned x, unsigned y, uint32_t& rng) {        // src/pathtracer.cu:892-897
    float offsetx = rng_next(rng) - 0.5f;
    float offsety = rng_next(rng) - 0.5f;
    float a0 = rng_next(rng), a1 = rng_next(rng);
    f2 aperture = uniform_disk(a0, a1);
    Ray ray; ray.medium = -1; ray.tmax = INFINITY;
    camera_ray(C(g->cam), x + offsetx, y + offsety, aperture, ray.o, ray.d);
    ray.tmin = g->eps;
    return ray;
}

// Path, src/pathtracer.cu:880-1021.  Returns false when Li is NaN/Inf (the caller then keeps the stale colour).
bool path_sample(unsigned x, unsigned y, unsigned pixel, unsigned iter, f3& Li_out) {
    uint32_t rng = rng_seed(pixel, iter);
    Ray r = primary_ray(x, y, rng);
    f3 Li = mk3(0.f, 0.f, 0.f), beta = mk3(1.f, 1.f, 1.f);
    Isect isect; isect.lightIdx = -1;
    bool specular = false;
    const int maxDepth = g->max_depth;
    for (int bounces = 0; bounces < maxDepth; ++bounces) {
        if (!intersect(r, &isect)) {
            if ((bounces == 0 || specular) && g->has_inf && g->inf.isvalid) Li += beta * inf_le(r.d);
            break;
        }
        f3 pos = isect.pos, nor = isect.nor, dpdu = isect.dpdu;
        const RefMaterial& mat = material(isect.matIdx);
        if (bounces == 0 || specular) {
            if (isect.lightIdx != -1) { Li += beta * area_le(g->lights[isect.lightIdx], nor, -r.d); break; }
        }
        if (!is_delta(mat.type)) {
            f3 Ld = direct_light(r, isect, mat, rng, false);
            Li += beta * Ld;
        }
        float c0 = rng_next(rng), c1 = rng_next(rng), c2 = rng_next(rng);
        f3 out, fr; float pdf;
        sample_bsdf(M(mat), albedo_of(mat, isect.uv), -r.d, nor, dpdu, mk3(c0, c1, c2), out, fr, pdf);
        if (is_black(fr)) break;
        beta *= fr * fabsf(dot(nor, out)) / pdf;
        specular = is_delta(mat.type);
        r = mk_ray(pos, out, -1, g->eps);
        if (bounces > 3) {
            float illumate = clampf(1.f - luminance(beta), 0.f, 1.f);
            if (rng_next(rng) < illumate) break;
            beta /= (1 - illumate);
        }
    }
    Li_out = Li;
    return !is_inf3(Li) && !is_nan3(Li);
}

// Volpath, src/pathtracer.cu:1025-1242
bool volpath_sample(unsigned x, unsigned y, unsigned pixel, unsigned iter, f3& Li_out) {
    uint32_t rng = rng_seed(pixel, iter);
    Ray r = primary_ray(x, y, rng);
    r.medium = g->cam.medium;
    f3 Li = mk3(0.f, 0.f, 0.f), beta = mk3(1.f, 1.f, 1.f);
    Isect isect; isect.lightIdx = -1;
    bool specular = false;
    const int maxDepth = g->max_depth;
    const int light_size = (int)g->lights.size();
    for (int bounces = 0; bounces < maxDepth; ++bounces) {
        if (!intersect(r, &isect)) {
            if ((bounces == 0 || specular) && g->has_inf && g->inf.isvalid) Li += beta * inf_le(r.d);
            break;
        }
        f3 pos = isect.pos, nor = isect.nor, dpdu = isect.dpdu;
        float sampledDist = 0.f;
        bool sampledMedium = false;
        if (r.medium >= 0) beta *= medium_sample(g->mediums[r.medium], r, rng, sampledDist, sampledMedium);
        if (is_black(beta)) break;
        if (sampledMedium) {
            const RefMedium& med = g->mediums[r.medium];
            bool inf = false;
            float u = rng_next(rng);
            float choicePdf;
            int idx = lookup_light(u, choicePdf);
            if (idx == light_size) inf = true;
            f3 samplePos = at(r, sampledDist);
            float ua = rng_next(rng), ub = rng_next(rng);
            f3 radiance, lightNor; Ray shadowRay; float lightPdf;
            if (!inf) area_sample_light(g->lights[idx], samplePos, mk2(ua, ub), radiance, shadowRay, lightNor, lightPdf, g->eps);
            else inf_sample_light(samplePos, mk2(ua, ub), radiance, shadowRay, lightNor, lightPdf, g->eps);
            shadowRay.medium = r.medium;
            f3 tr = transmittance(shadowRay, rng);
            float phase, unuse;
            medium_phase(med, -r.d, shadowRay.d, phase, unuse);
            if (!is_black(radiance)) Li += tr * beta * phase * radiance / (lightPdf * choicePdf);
            float pdf;
            float pa = rng_next(rng), pb = rng_next(rng);
            f3 dir;
            medium_sample_phase(med, mk2(pa, pb), dir, phase, pdf);
            r = mk_ray(samplePos, dir, r.medium, g->eps);
            specular = false;
        } else {
            if (bounces == 0 || specular) {
                if (isect.lightIdx != -1) {
                    f3 tr = mk3(1.f, 1.f, 1.f);
                    if (r.medium >= 0) tr = medium_tr(g->mediums[r.medium], r, rng);
                    Li += tr * beta * area_le(g->lights[isect.lightIdx], nor, -r.d);
                    break;
                }
            }
            if (isect.matIdx == -1) {
                bounces--;
                int m = dot(r.d, isect.nor) > 0 ? isect.mediumOutside : isect.mediumInside;
                r = mk_ray(pos, r.d, m, g->eps);
                continue;
            }
            const RefMaterial& mat = material(isect.matIdx);
            if (!is_delta(mat.type)) {
                f3 Ld = direct_light(r, isect, mat, rng, true);
                Li += beta * Ld;
            }
            float c0 = rng_next(rng), c1 = rng_next(rng), c2 = rng_next(rng);
            f3 out, fr; float pdf;
            sample_bsdf(M(mat), albedo_of(mat, isect.uv), -r.d, nor, dpdu, mk3(c0, c1, c2), out, fr, pdf);
            if (is_black(fr)) break;
            beta *= fr * fabsf(dot(nor, out)) / pdf;
            specular = is_delta(mat.type);
            int m = dot(out, nor) > 0 ? isect.mediumOutside : isect.mediumInside;
            m = dot(-r.d, nor) * dot(out, nor) > 0 ? r.medium : m;
            r = mk_ray(pos, out, m, g->eps);
        }
        if (bounces > 3) {
            float illumate = clampf(1.f - luminance(beta), 0.f, 1.f);
            if (rng_next(rng) < illumate) break;
            beta /= (1 - illumate);
        }
    }
    Li_out = Li;
    return !is_inf3(Li) && !is_nan3(Li);
}

}  // namespace

extern "C" int oracle_begin(const b200pt_scene_view* v, unsigned w, unsigned h, float eps) {
    if (g) return -1;
    if (v->integrator_type != B200PT_IT_PT && v->integrator_type != B200PT_IT_VPT) return -4;
    g = new SceneO();
    std::memcpy(&g->cam, v->camera, sizeof(RefCamera));
    g->prims.assign((const RefPrimitive*)v->prims, (const RefPrimitive*)v->prims + v->n_prims);
    g->nodes.assign((const RefLinearBVHNode*)v->nodes, (const RefLinearBVHNode*)v->nodes + v->n_nodes);
    g->mats.assign((const RefMaterial*)v->materials, (const RefMaterial*)v->materials + v->n_materials);
    if (v->n_mediums) g->mediums.assign((const RefMedium*)v->mediums, (const RefMedium*)v->mediums + v->n_mediums);
    g->densities.resize(g->mediums.size());
    for (size_t i = 0; i < g->mediums.size(); ++i) {
        RefMedium& m = g->mediums[i];
        if (m.type != REF_MEDIUM_HETEROGENEOUS) continue;
        g->densities[i].assign(m.density, m.density + (size_t)m.nx * m.ny * m.nz);
        m.density = g->densities[i].data();
    }
    if (v->n_lights) g->lights.assign((const RefArea*)v->lights, (const RefArea*)v->lights + v->n_lights);
    if (v->infinite) {
        std::memcpy(&g->inf, v->infinite, sizeof(RefInfinite));
        g->has_inf = true;
        if (g->inf.isvalid) g->inf_texels.assign(g->inf.data, g->inf.data + 3 * (size_t)g->inf.width * g->inf.height);
    }
    g->cdf.assign(v->light_distribution, v->light_distribution + v->n_light_distribution);
    for (int i = 0; i < v->n_textures; ++i) {
        SceneO::Tex T; T.w = v->textures[i].width; T.h = v->textures[i].height;
        const unsigned char* src = (const unsigned char*)v->textures[i].texels;
        T.rgba.assign(src, src + 4 * (size_t)T.w * T.h);
        g->textures.push_back(T);
    }
    g->integrator = v->integrator_type; g->max_depth = v->max_depth; g->eps = eps; g->w = w; g->h = h;
    g->acc.assign((size_t)w * h, mk3(0, 0, 0));
    g->color.assign((size_t)w * h, mk3(0, 0, 0));
    return 0;
}
extern "C" int oracle_set_camera(const void* cam104) { if (!g) return -1; std::memcpy(&g->cam, cam104, sizeof(RefCamera)); return 0; }

// Render(iter) for iter = first..first+n-1 followed by Output each time (src/pathtracer.cu:2705-2750, :2516-2531).
extern "C" int oracle_render(unsigned first_iter, unsigned n, int reset_first, float* out_host, int nthreads) {
    if (!g) return -1;
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    const int W = (int)g->w, H = (int)g->h;
    for (unsigned it = first_iter; it < first_iter + n; ++it) {
        bool reset = reset_first && it == first_iter;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                unsigned pixel = (unsigned)x + (unsigned)y * (unsigned)W;
                f3 Li;
                bool ok = g->integrator == B200PT_IT_PT ? path_sample(x, y, pixel, it, Li) : volpath_sample(x, y, pixel, it, Li);
                if (ok) g->color[pixel] = Li;                       // else: stale colour kept (src/pathtracer.cu:1019)
                if (reset) g->acc[pixel] = mk3(0, 0, 0);
                g->acc[pixel] += g->color[pixel];
                if (out_host && it == first_iter + n - 1) {
                    f3 c = g->acc[pixel] / (float)(int)it;
                    c = g->cam.filmic ? filmic_tonemap(c) : gamma_correct(c);
                    out_host[3 * (size_t)pixel + 0] = c.x; out_host[3 * (size_t)pixel + 1] = c.y; out_host[3 * (size_t)pixel + 2] = c.z;
                }
            }
    }
    return 0;
}
extern "C" int oracle_get_accum(float* host) { if (!g) return -1; std::memcpy(host, g->acc.data(), sizeof(f3) * g->acc.size()); return 0; }
extern "C" int oracle_get_color(float* host) { if (!g) return -1; std::memcpy(host, g->color.data(), sizeof(f3) * g->color.size()); return 0; }
extern "C" void oracle_count(int on) { g_counting = on != 0; for (auto& c : g_counts) c = 0; }
extern "C" void oracle_get_counts(unsigned long long* out4) { for (int i = 0; i < 4; ++i) out4[i] = g_counts[i]; }
extern "C" int oracle_end() { if (!g) return -1; delete g; g = nullptr; return 0; }

// ---- function-level entry points (same signatures as the refhost_* ones in oracle/refbuild/ref_host_harness.cpp)
static Ray ray_from8(const float* r8) { Ray r; r.o = mk3(r8[0], r8[1], r8[2]); r.d = mk3(r8[3], r8[4], r8[5]); r.tmin = r8[6]; r.tmax = r8[7]; r.medium = -1; return r; }
static void put_isect(const Isect& is, float* o) {
    float tmp[16] = {is.pos.x, is.pos.y, is.pos.z, is.nor.x, is.nor.y, is.nor.z, is.uv.x, is.uv.y, is.dpdu.x, is.dpdu.y, is.dpdu.z};
    std::memcpy(o, tmp, 11 * 4);
    int ints[5] = {is.matIdx, is.bssrdf, is.lightIdx, is.mediumInside, is.mediumOutside};
    std::memcpy(o + 11, ints, 5 * 4);
}
extern "C" int oracle_intersect(const float* ray8, float* t_out, float* isect16) {
    Ray r = ray_from8(ray8);
    Isect is; std::memset(&is, 0, sizeof(is)); is.lightIdx = -1;
    bool hit = intersect(r, &is);
    *t_out = r.tmax; put_isect(is, isect16);
    return hit ? 1 : 0;
}
extern "C" int oracle_intersect_p(const float* ray8) { Ray r = ray_from8(ray8); return intersect_p(r) ? 1 : 0; }
extern "C" void oracle_sample_bsdf(const void* mat72, const float* in3, const float* nor3, const float* uv2, const float* dpdu3,
                                   const float* u3, float* out3, float* fr3, float* pdf) {
    RefMaterial m; std::memcpy(&m, mat72, sizeof(m));
    f3 out = mk3(0, 0, 0), fr = mk3(0, 0, 0); float p = 0;
    sample_bsdf(M(m), albedo_of(m, mk2(uv2[0], uv2[1])), ld3(in3), ld3(nor3), ld3(dpdu3), ld3(u3), out, fr, p);
    out3[0] = out.x; out3[1] = out.y; out3[2] = out.z; fr3[0] = fr.x; fr3[1] = fr.y; fr3[2] = fr.z; *pdf = p;
}
extern "C" void oracle_fr(const void* mat72, const float* in3, const float* out3, const float* nor3, const float* uv2,
                          const float* dpdu3, float* fr3, float* pdf) {
    RefMaterial m; std::memcpy(&m, mat72, sizeof(m));
    f3 fr = mk3(0, 0, 0); float p = 0;
    eval_bsdf(M(m), albedo_of(m, mk2(uv2[0], uv2[1])), ld3(in3), ld3(out3), ld3(nor3), ld3(dpdu3), fr, p);
    fr3[0] = fr.x; fr3[1] = fr.y; fr3[2] = fr.z; *pdf = p;
}
extern "C" void oracle_camera_ray(const void* cam104, float x, float y, float ax, float ay, float* o3, float* d3) {
    RefCamera c; std::memcpy(&c, cam104, sizeof(c));
    f3 o, d;
    camera_ray(C(c), x, y, mk2(ax, ay), o, d);
    o3[0] = o.x; o3[1] = o.y; o3[2] = o.z; d3[0] = d.x; d3[1] = d.y; d3[2] = d.z;
}
extern "C" void oracle_rng(unsigned pixel, unsigned iter, int n, float* out) {
    uint32_t s = rng_seed(pixel, iter);
    for (int i = 0; i < n; ++i) out[i] = rng_next(s);
}
extern "C" void oracle_area_sample(const void* area192, const float* pos3, const float* u2, float eps, float* rad3, float* ray8,
                                   float* nor3, float* pdf) {
    RefArea a; std::memcpy(&a, area192, sizeof(a));
    f3 rad, nor; Ray r; float p;
    area_sample_light(a, ld3(pos3), mk2(u2[0], u2[1]), rad, r, nor, p, eps);
    rad3[0] = rad.x; rad3[1] = rad.y; rad3[2] = rad.z; nor3[0] = nor.x; nor3[1] = nor.y; nor3[2] = nor.z; *pdf = p;
    ray8[0] = r.o.x; ray8[1] = r.o.y; ray8[2] = r.o.z; ray8[3] = r.d.x; ray8[4] = r.d.y; ray8[5] = r.d.z; ray8[6] = r.tmin; ray8[7] = r.tmax;
}
extern "C" void oracle_tonemap(const float* in3, int filmic, float* out3) {
    f3 c = filmic ? filmic_tonemap(ld3(in3)) : gamma_correct(ld3(in3));
    out3[0] = c.x; out3[1] = c.y; out3[2] = c.z;
}
