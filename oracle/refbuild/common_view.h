// TEST INFRASTRUCTURE — builds the reference's `Scene` (src/scene.h:26) from the flat C-ABI scene view
// (include/b200pt.h) so that the reference's own BeginRender/Render (CUDA build) or kernel bodies (host
// build) can be driven with exactly the same arrays as the product.  Our code, not reference code.
#pragma once
#include "b200pt.h"
#include <new>
static_assert(sizeof(Camera) == B200PT_SIZEOF_CAMERA, "Camera");
static_assert(sizeof(Primitive) == B200PT_SIZEOF_PRIMITIVE, "Primitive");
static_assert(sizeof(LinearBVHNode) == B200PT_SIZEOF_BVHNODE, "LinearBVHNode");
static_assert(sizeof(Material) == B200PT_SIZEOF_MATERIAL, "Material");
static_assert(sizeof(Medium) == B200PT_SIZEOF_MEDIUM, "Medium");
static_assert(sizeof(Area) == B200PT_SIZEOF_AREA, "Area");
static_assert(sizeof(Infinite) == B200PT_SIZEOF_INFINITE, "Infinite");

template <class T> static void fill_vec(std::vector<T>& v, const void* src, int n) {
    v.clear();
    if (n <= 0) return;
    T* tmp = (T*)malloc(sizeof(T) * (size_t)n);
    memcpy((void*)tmp, src, sizeof(T) * (size_t)n);
    for (int i = 0; i < n; ++i) v.push_back(tmp[i]);
    free(tmp);
}

// Fills everything BeginRender reads (src/pathtracer.cu:2578-2671).  `cam` must outlive the scene.
static void scene_from_view(Scene& scene, Camera* cam, const b200pt_scene_view* v) {
    memcpy((void*)cam, v->camera, sizeof(Camera));
    scene.camera = cam;
    fill_vec(scene.bvh.prims, v->prims, v->n_prims);
    scene.bvh.total_nodes = v->n_nodes;
    scene.bvh.linear_root = (LinearBVHNode*)malloc(sizeof(LinearBVHNode) * (size_t)(v->n_nodes > 0 ? v->n_nodes : 1));
    if (v->n_nodes > 0) memcpy((void*)scene.bvh.linear_root, v->nodes, sizeof(LinearBVHNode) * (size_t)v->n_nodes);
    fill_vec(scene.materials, v->materials, v->n_materials);
    fill_vec(scene.mediums, v->mediums, v->n_mediums);
    for (int i = 0; i < v->n_mediums; ++i) {
        if (scene.mediums[i].type == MT_HETEROGENEOUS) {   // BeginRender delete[]s the host grid (:2621)
            Heterogeneous& m = scene.mediums[i].heterogeneous;
            size_t n = (size_t)m.nx * m.ny * m.nz;
            float* copy = new float[n];
            memcpy(copy, m.density, n * sizeof(float));
            m.density = copy;
        }
    }
    fill_vec(scene.lights, v->lights, v->n_lights);
    if (v->infinite) memcpy((void*)&scene.infinite, v->infinite, sizeof(Infinite));
    else { memset((void*)&scene.infinite, 0, sizeof(Infinite)); scene.infinite.isvalid = false; }
    scene.lightDistribution.assign(v->light_distribution, v->light_distribution + v->n_light_distribution);
    scene.textures.clear();
    for (int i = 0; i < v->n_textures; ++i) {      // Texture only has a from-file ctor (src/texture.h:15)
        Texture* t = (Texture*)operator new(sizeof(Texture));
        new (&t->data) std::vector<uchar4>((const uchar4*)v->textures[i].texels,
                                           (const uchar4*)v->textures[i].texels + (size_t)v->textures[i].width * v->textures[i].height);
        t->width = v->textures[i].width; t->height = v->textures[i].height;
        scene.textures.push_back(*t);
    }
    scene.integrator.type = (IntegratorType)v->integrator_type;
    scene.integrator.maxDepth = v->max_depth;
}
