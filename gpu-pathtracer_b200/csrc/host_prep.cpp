// Host-side scene preparation behind the C ABI (include/b200pt.h): binned-SAH BVH build + flatten into the
// reference's LinearBVHNode layout, camera constructor arithmetic, and the light-selection CDF.
//
// These are the callers on the host side of the hot path.  They are
// written index-based (no 176-B primitive copies per level as in src/bvh.cpp:131-148) but follow the same
// decisions so that the output is bit-identical to the reference builder (tests/test_host_prep.py).
// Compile with -ffp-contract=off: every float expression must round like the reference host build.
#include "b200pt.h"
#include "ref_layouts.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

int b200pt_internal_fail(int code, const char* msg);   // b200pt_api.cu: sets b200pt_last_error()
extern "C" int b200pt_internal_prims_finite(const void* prims, int n, int* bad);
namespace {

// Host-side min/max as the reference's host build defines them (src/cutil_math.h:36-44): on ties (and NaN)
// the SECOND operand wins — this decides the sign of zero in node boxes, so it is part of bit-parity.
inline float minf2(float a, float b) { return a < b ? a : b; }
inline float maxf2(float a, float b) { return a > b ? a : b; }

struct Box {
    float lo[3], hi[3];
    void reset() { lo[0] = lo[1] = lo[2] = INFINITY; hi[0] = hi[1] = hi[2] = -INFINITY; }
    void grow(const float* p) {
        for (int a = 0; a < 3; ++a) { lo[a] = minf2(lo[a], p[a]); hi[a] = maxf2(hi[a], p[a]); }
    }
    void grow(const Box& b) {
        for (int a = 0; a < 3; ++a) { lo[a] = minf2(b.lo[a], lo[a]); hi[a] = maxf2(b.hi[a], hi[a]); }
    }
    // BBox::SurfaceArea (src/bbox.h:63-66)
    float area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return 2.f * (dx * dy + dy * dz + dz * dx);
    }
};

// GetBBox(Primitive&) (src/bvh.cpp:3-10): triangle mesh.h:28, sphere sphere.h:17, line line.h:16
Box prim_box(const RefPrimitive& p) {
    Box b; b.reset();
    if (p.type == REF_GT_TRIANGLE) {
        b.grow(p.u.triangle.v1.v); b.grow(p.u.triangle.v2.v); b.grow(p.u.triangle.v3.v);
    } else if (p.type == REF_GT_SPHERE) {
        const RefSphere& s = p.u.sphere;
        for (int a = 0; a < 3; ++a) { b.lo[a] = s.origin[a] - s.radius; b.hi[a] = s.origin[a] + s.radius; }
    } else {
        const RefLine& l = p.u.line;
        float w = l.width0 > l.width1 ? l.width0 : l.width1;
        float q[3];
        for (int a = 0; a < 3; ++a) q[a] = l.p0[a] - w; b.grow(q);
        for (int a = 0; a < 3; ++a) q[a] = l.p0[a] + w; b.grow(q);
        for (int a = 0; a < 3; ++a) q[a] = l.p1[a] - w; b.grow(q);
        for (int a = 0; a < 3; ++a) q[a] = l.p1[a] + w; b.grow(q);
    }
    return b;
}

// Bucket indices (src/bvh.cpp:74-76) are computed from primitive boxes without range checks — in the reference a NaN /
// Inf vertex indexes outside its 12-entry arrays.  Behind a public C ABI the input is checked instead: every box, its
// centre (lo + hi) and its extent (hi - lo) must be finite, which keeps every bucket index in [0, 12].
bool prims_finite(const RefPrimitive* prims, int n, int* bad) {
    for (int i = 0; i < n; ++i) {
        const Box b = prim_box(prims[i]);
        for (int a = 0; a < 3; ++a)
            if (!std::isfinite(b.lo[a]) || !std::isfinite(b.hi[a]) || !std::isfinite(b.lo[a] + b.hi[a]) || !std::isfinite(b.hi[a] - b.lo[a])) { *bad = i; return false; }
    }
    return true;
}

struct Builder {
    const RefPrimitive* prims;
    std::vector<Box> boxes;          // per input primitive
    std::vector<float> centre;       // 3 per primitive: (lo+hi)*0.5f  (BBox::Centric, src/bbox.h:49)
    RefLinearBVHNode* nodes;
    int capacity;
    int n_nodes = 0;
    std::vector<int> leaf_order;     // input primitive index, in emitted (leaf) order
    bool overflow = false;

    static const int kBuckets = 12;  // src/bvh.cpp:66

    // Emits nodes in pre-order, which is exactly the layout BVH::flatten produces (left child = idx+1).
    void build(std::vector<int>& ids, const Box& bbox) {
        if (n_nodes >= capacity) { overflow = true; return; }
        const int me = n_nodes++;
        RefLinearBVHNode& nd = nodes[me];
        std::memset(&nd, 0, sizeof(nd));
        std::memcpy(nd.fmin, bbox.lo, 12); std::memcpy(nd.fmax, bbox.hi, 12);
        nd.start = nd.end = -1; nd.second_child_offset = -1;

        const size_t n = ids.size();
        float diag[3] = {bbox.hi[0] - bbox.lo[0], bbox.hi[1] - bbox.lo[1], bbox.hi[2] - bbox.lo[2]};
        bool leaf = n <= 4 || diag[0] < 0.0001f || diag[1] < 0.0001f || diag[2] < 0.0001f;   // src/bvh.cpp:43

        int best_axis = -1, best_bucket = 0;
        if (!leaf) {
            float best_cost = n * bbox.area();
            for (int axis = 0; axis < 3; ++axis) {
                Box bb[kBuckets]; int cnt[kBuckets];
                for (int k = 0; k < kBuckets; ++k) { bb[k].reset(); cnt[k] = 0; }
                const float v0 = bbox.lo[axis], v1 = bbox.hi[axis];
                for (size_t j = 0; j < n; ++j) {
                    const int id = ids[j];
                    int no = (int)((centre[3 * id + axis] - v0) / (v1 - v0) * kBuckets);
                    no = (no == 12) ? no - 1 : no;
                    no = no < 0 ? 0 : (no > kBuckets - 1 ? kBuckets - 1 : no);     // unreachable for finite boxes (checked on entry)
                    cnt[no]++; bb[no].grow(boxes[id]);
                }
                for (int j = 1; j < kBuckets; ++j) {
                    Box b0, b1; b0.reset(); b1.reset();
                    int c0 = 0, c1 = 0;
                    for (int k = 0; k < j; ++k) { b0.grow(bb[k]); c0 += cnt[k]; }
                    for (int k = j; k < kBuckets; ++k) { b1.grow(bb[k]); c1 += cnt[k]; }
                    float sa = (c0 == 0) ? 0 : b0.area() * c0;
                    float sb = (c1 == 0) ? 0 : b1.area() * c1;
                    float cost = sa + sb;
                    if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bucket = j; }
                }
            }
            if (best_axis == -1) leaf = true;                                                 // src/bvh.cpp:113
        }
        if (leaf) {
            nd.is_leaf = 1;
            if (n) {
                nd.start = (int)leaf_order.size();
                for (size_t i = 0; i < n; ++i) leaf_order.push_back(ids[i]);
                nd.end = (int)leaf_order.size() - 1;
            }
            return;
        }
        std::vector<int> left, right;
        Box bl, br; bl.reset(); br.reset();
        const float v0 = bbox.lo[best_axis], v1 = bbox.hi[best_axis];
        for (size_t i = 0; i < n; ++i) {
            const int id = ids[i];
            int no = (int)((centre[3 * id + best_axis] - v0) / (v1 - v0) * kBuckets);
            no = (no == kBuckets) ? no - 1 : no;
            if (no < best_bucket) { left.push_back(id); bl.grow(boxes[id]); }
            else { right.push_back(id); br.grow(boxes[id]); }
        }
        std::vector<int>().swap(ids);       // release before recursing
        build(left, bl);
        nodes[me].second_child_offset = n_nodes;
        build(right, br);
    }
};

inline void v3sub(const float* a, const float* b, float* o) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
inline void v3cross(const float* a, const float* b, float* o) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
inline float v3dot(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
// host normalize of cutil_math.h:1187 is v * (1.0f / sqrtf(dot))
inline void v3normalize(float* v) { float inv = 1.0f / sqrtf(v3dot(v, v)); v[0] *= inv; v[1] *= inv; v[2] *= inv; }

}  // namespace

extern "C" int b200pt_internal_prims_finite(const void* prims, int n, int* bad) { return prims_finite((const RefPrimitive*)prims, n, bad) ? 1 : 0; }

extern "C" int b200pt_bvh_build(const void* prims_in, int32_t n_prims, void* prims_out, void* nodes_out,
                                int32_t nodes_capacity, int32_t* n_nodes, float* root_box6) {
    if (!prims_in || !prims_out || !nodes_out || !n_nodes || n_prims <= 0)
        return b200pt_internal_fail(B200PT_EINVAL, "b200pt_bvh_build: null argument or empty input");
    int bad = -1;
    if (!prims_finite((const RefPrimitive*)prims_in, n_prims, &bad))
        return b200pt_internal_fail(B200PT_EINVAL, ("b200pt_bvh_build: primitive " + std::to_string(bad) + " has a non-finite bounding box").c_str());
    Builder b;
    b.prims = (const RefPrimitive*)prims_in;
    b.nodes = (RefLinearBVHNode*)nodes_out;
    b.capacity = nodes_capacity;
    b.boxes.resize(n_prims); b.centre.resize(3 * (size_t)n_prims);
    Box root; root.reset();
    std::vector<int> ids(n_prims);
    for (int i = 0; i < n_prims; ++i) {
        b.boxes[i] = prim_box(b.prims[i]);
        for (int a = 0; a < 3; ++a) b.centre[3 * (size_t)i + a] = (b.boxes[i].lo[a] + b.boxes[i].hi[a]) * 0.5f;
        root.grow(b.boxes[i]);
        ids[i] = i;
    }
    b.leaf_order.reserve(n_prims);
    b.build(ids, root);
    if (b.overflow) return b200pt_internal_fail(B200PT_ENOMEM, "b200pt_bvh_build: nodes_capacity too small");
    RefPrimitive* out = (RefPrimitive*)prims_out;
    for (int i = 0; i < n_prims; ++i) std::memcpy(&out[i], &b.prims[b.leaf_order[i]], sizeof(RefPrimitive));
    *n_nodes = b.n_nodes;
    if (root_box6) { std::memcpy(root_box6, root.lo, 12); std::memcpy(root_box6 + 3, root.hi, 12); }
    return B200PT_OK;
}

// Camera::Lookat (src/camera.h:124-129) followed by the constructor (src/camera.h:31-47), as the reference
// app chains them (src/parsescene.cpp:162-176, src/main.cpp:268-270).
extern "C" int b200pt_camera_init(void* camera104, const float* position3, const float* lookat3, const float* up3,
                                  float res_x, float res_y, float distance, float fov, float aperture_radius,
                                  float focal_distance, int filmic, int environment, int medium) {
    if (!camera104 || !position3 || !lookat3 || !up3) return B200PT_EINVAL;
    RefCamera c;
    std::memset(&c, 0, sizeof(c));
    std::memcpy(c.position, position3, 12);
    v3sub(position3, lookat3, c.w); v3normalize(c.w);
    v3cross(up3, c.w, c.u); v3normalize(c.u);
    v3cross(c.w, c.u, c.v); v3normalize(c.v);
    c.resolution[0] = res_x; c.resolution[1] = res_y;
    c.distance = distance; c.fov = fov; c.apertureRadius = aperture_radius; c.focalDistance = focal_distance;
    c.filmic = filmic ? 1 : 0; c.environment = environment ? 1 : 0; c.medium = medium;
    float half_fov = fov * .5f;
    float radians = half_fov / 180.0 * 3.14159265358f;       // DegreesToRadians, double intermediate (src/common.h:48-51)
    c.height = tanf(radians) * distance;
    c.width = c.height * res_x / res_y;
    c.area = 4.f * c.width * c.height;
    c.pixel2screen[0] = 2.f * c.width / res_x;
    c.pixel2screen[1] = 2.f * c.height / res_y;
    c.ratio = focal_distance / distance;
    std::memcpy(camera104, &c, sizeof(c));
    return B200PT_OK;
}

// Scene::Init light CDF (src/scene.h:65-82): running sum of luminance(power), normalised by the total.
extern "C" int b200pt_light_distribution(const void* lights, int32_t n_lights, const void* infinite_or_null,
                                         float* out, int32_t* n_out) {
    if (!out || !n_out || n_lights < 0) return B200PT_EINVAL;
    const RefArea* L = (const RefArea*)lights;
    const float luma[3] = {0.212671f, 0.715160f, 0.072169f};
    float sum = 0.f;
    int n = 0;
    out[n++] = 0.f;
    for (int i = 0; i < n_lights; ++i) {
        float e1[3], e2[3], cr[3];
        v3sub(L[i].triangle.v2.v, L[i].triangle.v1.v, e1);
        v3sub(L[i].triangle.v3.v, L[i].triangle.v1.v, e2);
        v3cross(e1, e2, cr);
        float area = sqrtf(v3dot(cr, cr)) * 0.5f;                 // Triangle::GetSurfaceArea, src/mesh.h:39
        float power[3];
        for (int a = 0; a < 3; ++a) power[a] = L[i].radiance[a] * area * 3.14159265358f;   // Area::GetPower, src/area.h:34
        sum += v3dot(luma, power);
        out[n++] = sum;
    }
    if (infinite_or_null) {
        const RefInfinite* inf = (const RefInfinite*)infinite_or_null;
        if (inf->isvalid) {
            const float* texel0 = (const float*)inf->data;      // only the first texel (src/infinite.h:43)
            float s = 12.56637061432f * inf->radius * inf->radius;
            float power[3] = {s * texel0[0], s * texel0[1], s * texel0[2]};
            sum += v3dot(luma, power);
            out[n++] = sum;
        }
    }
    for (int i = 0; i < n; ++i) out[i] /= sum;
    *n_out = n;
    return B200PT_OK;
}

// Infinite::Init(root_box) (src/infinite.h:61-63): bounding sphere of the scene box (src/bbox.h:98-101).
extern "C" int b200pt_infinite_init(void* infinite72, const float* root_box6) {
    if (!infinite72 || !root_box6) return B200PT_EINVAL;
    RefInfinite inf; std::memcpy(&inf, infinite72, sizeof(inf));
    float d[3];
    for (int a = 0; a < 3; ++a) {
        inf.center[a] = (root_box6[a] + root_box6[3 + a]) * 0.5f;
        d[a] = root_box6[3 + a] - inf.center[a];
    }
    inf.radius = sqrtf(v3dot(d, d));
    std::memcpy(infinite72, &inf, sizeof(inf));
    return B200PT_OK;
}

// ---- bvh.cache (BVH::LoadOrBuildBVH, src/bvh.cpp:189-217) -------------------------------------------------------------
// Byte stream: int total_nodes | int n_prims | float[3] root min | float[3] root max | Primitive[n_prims] |
// LinearBVHNode[total_nodes].  The reference reads it without any validation; here the header is checked against
// the file size before anything is copied.
namespace {
int cache_fail(int code, const std::string& msg) { return b200pt_internal_fail(code, msg.c_str()); }
struct CacheHeader { int32_t n_nodes; int32_t n_prims; float box[6]; };
static_assert(sizeof(CacheHeader) == 32, "bvh.cache header is 32 bytes");

struct File {
    FILE* fp = nullptr;
    explicit File(const char* path, const char* mode) { fp = path ? std::fopen(path, mode) : nullptr; }
    ~File() { if (fp) std::fclose(fp); }
};

int cache_read_header(FILE* fp, CacheHeader& h) {
    if (std::fread(&h, sizeof(h), 1, fp) != 1) return cache_fail(B200PT_EINVAL, "bvh.cache: shorter than its 32-byte header");
    if (h.n_nodes <= 0 || h.n_prims <= 0 || (int64_t)h.n_nodes > 2 * (int64_t)h.n_prims + 1)
        return cache_fail(B200PT_EINVAL, "bvh.cache: implausible header (" + std::to_string(h.n_nodes) + " nodes, " + std::to_string(h.n_prims) + " primitives)");
    if (std::fseek(fp, 0, SEEK_END) != 0) return cache_fail(B200PT_EINVAL, "bvh.cache: seek failed");
    const long long size = std::ftell(fp);
    const long long want = (long long)sizeof(h) + (long long)h.n_prims * (long long)sizeof(RefPrimitive) +
                           (long long)h.n_nodes * (long long)sizeof(RefLinearBVHNode);
    if (size != want) return cache_fail(B200PT_EINVAL, "bvh.cache: file size " + std::to_string(size) + " B does not match its header (" + std::to_string(want) + " B)");
    if (std::fseek(fp, (long)sizeof(h), SEEK_SET) != 0) return cache_fail(B200PT_EINVAL, "bvh.cache: seek failed");
    return B200PT_OK;
}
}  // namespace

extern "C" int b200pt_bvh_cache_save(const char* path, const void* prims, int32_t n_prims, const void* nodes,
                                     int32_t n_nodes, const float* root_box6) {
    if (!path || !prims || !nodes || !root_box6 || n_prims <= 0 || n_nodes <= 0) return cache_fail(B200PT_EINVAL, "bvh_cache_save: null or empty argument");
    const std::string tmp = std::string(path) + ".tmp";
    {
        File f(tmp.c_str(), "wb");
        if (!f.fp) return cache_fail(B200PT_EINVAL, "bvh_cache_save: cannot create " + tmp);
        CacheHeader h; h.n_nodes = n_nodes; h.n_prims = n_prims; std::memcpy(h.box, root_box6, sizeof(h.box));
        bool ok = std::fwrite(&h, sizeof(h), 1, f.fp) == 1;
        ok = ok && std::fwrite(prims, sizeof(RefPrimitive), (size_t)n_prims, f.fp) == (size_t)n_prims;
        ok = ok && std::fwrite(nodes, sizeof(RefLinearBVHNode), (size_t)n_nodes, f.fp) == (size_t)n_nodes;
        ok = ok && std::fflush(f.fp) == 0;
        if (!ok) { std::remove(tmp.c_str()); return cache_fail(B200PT_EINVAL, "bvh_cache_save: short write to " + tmp); }
    }
    if (std::rename(tmp.c_str(), path) != 0) { std::remove(tmp.c_str()); return cache_fail(B200PT_EINVAL, std::string("bvh_cache_save: cannot rename onto ") + path); }
    return B200PT_OK;
}

extern "C" int b200pt_bvh_cache_info(const char* path, int32_t* n_prims, int32_t* n_nodes, float* root_box6) {
    File f(path, "rb");
    if (!f.fp) return cache_fail(B200PT_EINVAL, std::string("bvh.cache: cannot open ") + (path ? path : "(null)"));
    CacheHeader h;
    const int rc = cache_read_header(f.fp, h);
    if (rc != B200PT_OK) return rc;
    if (n_prims) *n_prims = h.n_prims;
    if (n_nodes) *n_nodes = h.n_nodes;
    if (root_box6) std::memcpy(root_box6, h.box, sizeof(h.box));
    return B200PT_OK;
}

extern "C" int b200pt_bvh_cache_load(const char* path, void* prims_out, int32_t prims_capacity, void* nodes_out,
                                     int32_t nodes_capacity, int32_t* n_prims, int32_t* n_nodes, float* root_box6) {
    if (!prims_out || !nodes_out || !n_prims || !n_nodes) return cache_fail(B200PT_EINVAL, "bvh_cache_load: null argument");
    File f(path, "rb");
    if (!f.fp) return cache_fail(B200PT_EINVAL, std::string("bvh.cache: cannot open ") + (path ? path : "(null)"));
    CacheHeader h;
    const int rc = cache_read_header(f.fp, h);
    if (rc != B200PT_OK) return rc;
    if (h.n_prims > prims_capacity || h.n_nodes > nodes_capacity) return cache_fail(B200PT_ENOMEM, "bvh_cache_load: destination arrays are too small");
    if (std::fread(prims_out, sizeof(RefPrimitive), (size_t)h.n_prims, f.fp) != (size_t)h.n_prims) return cache_fail(B200PT_EINVAL, "bvh.cache: short read (primitives)");
    if (std::fread(nodes_out, sizeof(RefLinearBVHNode), (size_t)h.n_nodes, f.fp) != (size_t)h.n_nodes) return cache_fail(B200PT_EINVAL, "bvh.cache: short read (nodes)");
    // Structural check of what the traversal will index with (the reference trusts the file blindly).
    const RefLinearBVHNode* nd = (const RefLinearBVHNode*)nodes_out;
    for (int i = 0; i < h.n_nodes; ++i) {
        if (nd[i].is_leaf) {
            if (nd[i].start < 0 || nd[i].end < nd[i].start || nd[i].end >= h.n_prims)
                return cache_fail(B200PT_EINVAL, "bvh.cache: leaf " + std::to_string(i) + " ranges outside the primitive array");
        } else if (nd[i].second_child_offset <= i + 1 || nd[i].second_child_offset >= h.n_nodes) {
            return cache_fail(B200PT_EINVAL, "bvh.cache: node " + std::to_string(i) + " links outside the node array");
        }
    }
    *n_prims = h.n_prims; *n_nodes = h.n_nodes;
    if (root_box6) std::memcpy(root_box6, h.box, sizeof(h.box));
    return B200PT_OK;
}

extern "C" int b200pt_bvh_load_or_build(const char* path, const void* prims_in, int32_t n_prims, void* prims_out,
                                        void* nodes_out, int32_t nodes_capacity, int32_t* n_nodes, float* root_box6,
                                        int32_t device, int32_t* was_loaded) {
    if (!path || !prims_in || !prims_out || !nodes_out || !n_nodes || n_prims <= 0) return cache_fail(B200PT_EINVAL, "bvh_load_or_build: null or empty argument");
    float box[6];
    int32_t np = 0, nn = 0;
    if (b200pt_bvh_cache_info(path, &np, &nn, box) == B200PT_OK && np == n_prims && nn <= nodes_capacity &&
        b200pt_bvh_cache_load(path, prims_out, n_prims, nodes_out, nodes_capacity, &np, n_nodes, box) == B200PT_OK) {
        if (root_box6) std::memcpy(root_box6, box, sizeof(box));
        if (was_loaded) *was_loaded = 1;
        return B200PT_OK;
    }
    if (was_loaded) *was_loaded = 0;
    const int rc = device >= 0
        ? b200pt_bvh_build_gpu(prims_in, n_prims, prims_out, nodes_out, nodes_capacity, n_nodes, box, device, nullptr)
        : b200pt_bvh_build(prims_in, n_prims, prims_out, nodes_out, nodes_capacity, n_nodes, box);
    if (rc != B200PT_OK) return rc;
    if (root_box6) std::memcpy(root_box6, box, sizeof(box));
    return b200pt_bvh_cache_save(path, prims_out, n_prims, nodes_out, *n_nodes, box);
}
