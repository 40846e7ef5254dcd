// Plain-C mirrors of the reference's data records, byte-for-byte, as they cross the drop-in boundary
// (sizes in include/b200pt.h).  Only layout — no behaviour.  float[3] stands for float3
// (4-byte aligned), float[2] members that were float2 in the reference are 8-byte aligned there, which the
// explicit padding below reproduces.
#pragma once
#include <stdint.h>

#define REF_GT_TRIANGLE 0
#define REF_GT_LINES    1
#define REF_GT_SPHERE   2

#define REF_MT_LAMBERTIAN      0
#define REF_MT_MIRROR          1
#define REF_MT_DIELECTRIC      2
#define REF_MT_ROUGHDIELECTRIC 3
#define REF_MT_ROUGHCONDUCTOR  4
#define REF_MT_SUBSTRATE       5

#define REF_MEDIUM_HOMOGENEOUS   0
#define REF_MEDIUM_HETEROGENEOUS 1

struct RefVertex {            // src/mesh.h:13, 48 B
    float v[3];
    float n[3];
    float uv[2];
    float t[3];
    float _pad;
};
struct RefTriangle {          // src/mesh.h:20, 168 B
    RefVertex v1, v2, v3;
    int32_t matIdx, bssrdfIdx, lightIdx, mediumInside, mediumOutside;
    int32_t _pad;
};
struct RefLine {              // src/line.h:8
    float p0[3], p1[3];
    float width0, width1;
    int32_t matIdx;
};
struct RefSphere {            // src/sphere.h:8, 32 B
    float origin[3];
    float radius;
    int32_t matIdx, bssrdfIdx, mediumInside, mediumOutside;
};
struct RefPrimitive {         // src/primitive.h:15, 176 B
    int32_t type;
    int32_t _pad;
    union { RefTriangle triangle; RefLine line; RefSphere sphere; } u;
};
struct RefLinearBVHNode {     // src/bvh.h:19, 40 B
    float fmin[3], fmax[3];
    int32_t second_child_offset;
    uint8_t is_leaf; uint8_t _pad[3];
    int32_t start, end;       // inclusive primitive range
};
struct RefMaterial {          // src/material.h:19, 72 B
    int32_t type;
    float alphaU, alphaV;
    float insideIOR, outsideIOR;
    float k[3], eta[3];
    float diffuse[3], specular[3];
    int32_t textureIdx;
};
struct RefMedium {            // src/medium.h:186, 104 B (Homogeneous :9 / Heterogeneous :52 union at 8)
    int32_t type;
    float g;
    float sigmaA[3], sigmaS[3], sigmaT[3];
    int32_t nx, ny, nz;       // heterogeneous only from here on
    const float* density;
    float invMaxDensity;
    float p0[3], p1[3];
    int32_t iterMax;
    int32_t evalTransmittanceType;
    int32_t _pad;
};
struct RefArea {              // src/area.h:7, 192 B
    float radiance[3];
    float _pad0;
    RefTriangle triangle;
    int32_t medium;
    int32_t _pad1;
};
struct RefInfinite {          // src/infinite.h:6, 72 B
    const float* data;        // width*height float3
    int32_t width, height;
    float center[3];
    float radius;
    float u[3], v[3], w[3];
    uint8_t isvalid; uint8_t _pad[3];
};
struct RefCamera {            // src/camera.h:8, 104 B
    float position[3];
    float u[3], v[3], w[3];
    float resolution[2];
    float distance, fov, apertureRadius, focalDistance;
    uint8_t filmic, environment; uint8_t _pad[2];
    int32_t medium;
    float width, height;      // private in the reference: tan(fov/2)*distance and its aspect multiple
    float pixel2screen[2];
    float ratio, area;
};

static_assert(sizeof(RefVertex) == 48, "Vertex");
static_assert(sizeof(RefTriangle) == 168, "Triangle");
static_assert(sizeof(RefPrimitive) == 176, "Primitive");
static_assert(sizeof(RefLinearBVHNode) == 40, "LinearBVHNode");
static_assert(sizeof(RefMaterial) == 72, "Material");
static_assert(sizeof(RefMedium) == 104, "Medium");
static_assert(sizeof(RefArea) == 192, "Area");
static_assert(sizeof(RefInfinite) == 72, "Infinite");
static_assert(sizeof(RefCamera) == 104, "Camera");
