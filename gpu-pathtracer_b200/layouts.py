"""numpy views of the reference's plain-data structs, byte-for-byte.

These are the records that cross the drop-in boundary; sizes and offsets were checked
against the reference headers with sizeof/offsetof (tests/test_layouts.py re-checks them against the compiled
reference when oracle/_ref is present).

    Vertex         src/mesh.h:13        48 B
    Triangle       src/mesh.h:20       168 B
    Primitive      src/primitive.h:15  176 B   (tag@0, union@8)
    LinearBVHNode  src/bvh.h:19         40 B
    Material       src/material.h:19    72 B
    Medium         src/medium.h:186    104 B
    Area           src/area.h:7        192 B
    Infinite       src/infinite.h:6     72 B
    Camera         src/camera.h:8      104 B
"""
import numpy as np

f3 = (np.float32, 3)
f2 = (np.float32, 2)

Vertex = np.dtype({"names": ["v", "n", "uv", "t"], "formats": [f3, f3, f2, f3],
                   "offsets": [0, 12, 24, 32], "itemsize": 48})

Triangle = np.dtype({"names": ["v1", "v2", "v3", "matIdx", "bssrdfIdx", "lightIdx", "mediumInside", "mediumOutside"],
                     "formats": [Vertex, Vertex, Vertex, np.int32, np.int32, np.int32, np.int32, np.int32],
                     "offsets": [0, 48, 96, 144, 148, 152, 156, 160], "itemsize": 168})

Sphere = np.dtype({"names": ["origin", "radius", "matIdx", "bssrdfIdx", "mediumInside", "mediumOutside"],
                   "formats": [f3, np.float32, np.int32, np.int32, np.int32, np.int32],
                   "offsets": [0, 12, 16, 20, 24, 28], "itemsize": 32})

Line = np.dtype({"names": ["p0", "p1", "width0", "width1", "matIdx"],
                 "formats": [f3, f3, np.float32, np.float32, np.int32],
                 "offsets": [0, 12, 24, 28, 32], "itemsize": 36})

GT_TRIANGLE, GT_LINES, GT_SPHERE = 0, 1, 2

# union { Triangle; Line; Sphere } at offset 8: one dtype per union member, same 176-B record
# (numpy repacks dtypes with overlapping fields on concatenate, so the members are separate views)
Primitive = np.dtype({"names": ["type", "triangle"], "formats": [np.int32, Triangle], "offsets": [0, 8], "itemsize": 176})
PrimitiveSphere = np.dtype({"names": ["type", "sphere"], "formats": [np.int32, Sphere], "offsets": [0, 8], "itemsize": 176})
PrimitiveLine = np.dtype({"names": ["type", "line"], "formats": [np.int32, Line], "offsets": [0, 8], "itemsize": 176})


def cat(arrays, dtype):
    """Concatenate record arrays byte-wise (keeps explicit offsets/padding intact)."""
    arrays = [np.ascontiguousarray(a) for a in arrays if a is not None and len(a)]
    if not arrays:
        return np.zeros(0, dtype)
    raw = np.concatenate([a.view(np.uint8).reshape(len(a), dtype.itemsize) for a in arrays])
    return np.ascontiguousarray(raw).view(dtype).reshape(-1)

LinearBVHNode = np.dtype({"names": ["fmin", "fmax", "second_child_offset", "is_leaf", "start", "end"],
                          "formats": [f3, f3, np.int32, np.uint8, np.int32, np.int32],
                          "offsets": [0, 12, 24, 28, 32, 36], "itemsize": 40})

MT_LAMBERTIAN, MT_MIRROR, MT_DIELECTRIC, MT_ROUGHDIELECTRIC, MT_ROUGHCONDUCTOR, MT_SUBSTRATE = range(6)
MATERIAL_TYPES = {"lambertian": 0, "mirror": 1, "dielectric": 2, "roughdielectric": 3, "roughconduct": 4, "substrate": 5}

Material = np.dtype({"names": ["type", "alphaU", "alphaV", "insideIOR", "outsideIOR", "k", "eta", "diffuse", "specular", "textureIdx"],
                     "formats": [np.int32, np.float32, np.float32, np.float32, np.float32, f3, f3, f3, f3, np.int32],
                     "offsets": [0, 4, 8, 12, 16, 20, 32, 44, 56, 68], "itemsize": 72})

MT_HOMOGENEOUS, MT_HETEROGENEOUS = 0, 1
Medium = np.dtype({"names": ["type", "g", "sigmaA", "sigmaS", "sigmaT", "nx", "ny", "nz", "density", "invMaxDensity",
                             "p0", "p1", "iterMax", "evalTransmittanceType"],
                   "formats": [np.int32, np.float32, f3, f3, f3, np.int32, np.int32, np.int32, np.uint64, np.float32,
                               f3, f3, np.int32, np.int32],
                   "offsets": [0, 4, 8, 20, 32, 44, 48, 52, 56, 64, 68, 80, 92, 96], "itemsize": 104})

Area = np.dtype({"names": ["radiance", "triangle", "medium"], "formats": [f3, Triangle, np.int32],
                 "offsets": [0, 16, 184], "itemsize": 192})

Infinite = np.dtype({"names": ["data", "width", "height", "center", "radius", "u", "v", "w", "isvalid"],
                     "formats": [np.uint64, np.int32, np.int32, f3, np.float32, f3, f3, f3, np.uint8],
                     "offsets": [0, 8, 12, 16, 28, 32, 44, 56, 68], "itemsize": 72})

Camera = np.dtype({"names": ["position", "u", "v", "w", "resolution", "distance", "fov", "apertureRadius", "focalDistance",
                             "filmic", "environment", "medium", "width", "height", "pixel2screen", "ratio", "area"],
                   "formats": [f3, f3, f3, f3, f2, np.float32, np.float32, np.float32, np.float32,
                               np.uint8, np.uint8, np.int32, np.float32, np.float32, f2, np.float32, np.float32],
                   "offsets": [0, 12, 24, 36, 48, 56, 60, 64, 68, 72, 73, 76, 80, 84, 88, 96, 100], "itemsize": 104})

# Intersection (src/intersection.h:6), 64 B — returned by the known-answer entry points of the test oracle
Intersection = np.dtype({"names": ["pos", "nor", "uv", "dpdu", "matIdx", "bssrdf", "lightIdx", "mediumInside", "mediumOutside"],
                         "formats": [f3, f3, f2, f3, np.int32, np.int32, np.int32, np.int32, np.int32],
                         "offsets": [0, 12, 24, 32, 44, 48, 52, 56, 60], "itemsize": 64})

IT_AO, IT_PT, IT_VPT = 0, 1, 2

assert Primitive.itemsize == 176 and Area.itemsize == 192 and Camera.itemsize == 104
