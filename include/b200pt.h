/* b200pt.h — C ABI of the B200-native unidirectional path-tracing integrator.
 *
 * Drop-in boundary: these entry points are what an FFI / adapter for the reference's
 *     void BeginRender(Scene&, unsigned w, unsigned h, float ep);                      (src/pathtracer.h:11)
 *     void Render(Scene&, unsigned w, unsigned h, Camera*, unsigned iter, bool reset,
 *                 float3* output);                                                     (src/pathtracer.h:10)
 *     void EndRender();                                                                (src/pathtracer.h:12)
 * binds.  `gpu-pathtracer_b200/csrc/pathtracer_adapter.cpp` implements exactly those three C++ symbols on
 * top of this header.
 *
 * All scene arrays are passed in the REFERENCE's own struct layouts (byte-for-byte what the reference's
 * BeginRender cudaMemcpy's, src/pathtracer.cu:2578-2671); the implementation re-lays them out on upload.
 * Plain pointers and sizes only; no C++ / torch types.  Every function returns 0 on success or a negative
 * B200PT_E* code; b200pt_last_error() returns a human-readable message for the calling thread's last failure.
 */
#ifndef B200PT_H
#define B200PT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- reference struct sizes (verified with sizeof/offsetof on the reference headers) ------------------ */
#define B200PT_SIZEOF_CAMERA      104  /* src/camera.h:8      */
#define B200PT_SIZEOF_PRIMITIVE   176  /* src/primitive.h:15  (tag@0, union@8; Triangle 168 B, src/mesh.h:20) */
#define B200PT_SIZEOF_BVHNODE      40  /* src/bvh.h:19        (bbox@0, second_child_offset@24, is_leaf@28, start@32, end@36) */
#define B200PT_SIZEOF_MATERIAL     72  /* src/material.h:19   */
#define B200PT_SIZEOF_MEDIUM      104  /* src/medium.h:186    */
#define B200PT_SIZEOF_AREA        192  /* src/area.h:7        */
#define B200PT_SIZEOF_INFINITE     72  /* src/infinite.h:6    */

/* IntegratorType values the hot path implements (src/scene.h:15-24). */
#define B200PT_IT_PT   1   /* IT_PT  -> Path    (src/pathtracer.cu:880)  */
#define B200PT_IT_VPT  2   /* IT_VPT -> Volpath (src/pathtracer.cu:1025) */

/* error codes */
#define B200PT_OK            0
#define B200PT_EINVAL       -1   /* bad argument / unsupported scene feature */
#define B200PT_ECUDA        -2   /* CUDA runtime error (message in b200pt_last_error) */
#define B200PT_ENOMEM       -3
#define B200PT_EUNSUPPORTED -4   /* integrator / primitive combination outside the hot path (ao/lt/bdpt/..., lines under vpt) */

/* One uchar4 texture (src/texture.h:9, uploaded at src/pathtracer.cu:2646-2661). */
typedef struct b200pt_texture {
    const void* texels;   /* width*height uchar4, row-major */
    int32_t width, height;
} b200pt_texture;

/* Everything BeginRender reads from `Scene&` (src/pathtracer.cu:2578-2671, :2711-2717). Host pointers. */
typedef struct b200pt_scene_view {
    const void*  camera;             /* 1 x Camera (104 B) — scene.camera                                   */
    const void*  prims;              /* n_prims x Primitive (176 B), in BVH leaf order — scene.bvh.prims     */
    const void*  nodes;              /* n_nodes x LinearBVHNode (40 B) — scene.bvh.linear_root               */
    const void*  materials;          /* n_materials x Material (72 B)                                        */
    const void*  mediums;            /* n_mediums x Medium (104 B); heterogeneous density = host pointer     */
    const void*  lights;             /* n_lights x Area (192 B)                                              */
    const void*  infinite;           /* 1 x Infinite (72 B; data = host float3 texels) or NULL               */
    const float* light_distribution; /* n_light_distribution floats — scene.lightDistribution (CDF)          */
    const b200pt_texture* textures;  /* n_textures entries or NULL                                           */
    int32_t n_prims, n_nodes, n_materials, n_mediums, n_lights, n_light_distribution, n_textures;
    int32_t integrator_type;         /* scene.integrator.type  (B200PT_IT_*)                                 */
    int32_t max_depth;               /* scene.integrator.maxDepth                                            */
} b200pt_scene_view;

/* Pixel-tile shard of the image this context renders (multi-GPU: the tile_w x tile_h screen tiles are dealt out to the
 * n_shards ranks so that every rank's tiles are spread over the whole image — every group of n consecutive tiles in
 * row-major order holds one tile per rank, rotating from group to group; the shards of ranks 0..n-1 are disjoint and
 * cover the image).  n_shards == 1 renders everything. */
typedef struct b200pt_shard {
    int32_t shard, n_shards;
    int32_t tile_w, tile_h;
} b200pt_shard;

typedef struct b200pt_ctx b200pt_ctx;

/* Replaces BeginRender (src/pathtracer.cu:2568): deep-copies + re-lays-out the scene onto `device`.
 * `shard` may be NULL (whole image).  width % 32 == 0 and height % 4 == 0 as in the reference's launch
 * (src/pathtracer.cu:2707-2709). */
int b200pt_create(const b200pt_scene_view* scene, uint32_t width, uint32_t height, float epsilon,
                  int device, const b200pt_shard* shard, b200pt_ctx** out_ctx);

/* Replaces `spp` consecutive calls Render(scene,w,h,camera,iter,reset&&iter==first_iter,output) for
 * iter = first_iter .. first_iter+spp-1 (src/pathtracer.cu:2705-2750).  spp == 1 is exactly one Render call.
 * `camera` (104-B reference Camera, host) is re-read every call like src/pathtracer.cu:2706.
 * `output` receives the tonemapped w*h float3 image of the LAST iteration (what Output writes,
 * src/pathtracer.cu:2516-2531); it may be NULL, a device pointer (output_is_device=1; the reference's
 * contract) or a host pointer (0).  Pixels outside this context's shard are written as 0.
 * Returns after the results are visible. */
int b200pt_render(b200pt_ctx* ctx, const void* camera, uint32_t first_iter, uint32_t spp, int reset,
                  float* output, int output_is_device);

/* Linear accumulation buffer kernel_acc_image (src/pathtracer.cu:10,2525): w*h float3, sum over iterations.
 * dst may be host (dst_is_device=0) or device.  Pixels outside the shard are 0 (so shards sum exactly). */
int b200pt_get_accum(b200pt_ctx* ctx, float* dst, int dst_is_device);
/* Device pointer of the same buffer (for an in-place NCCL reduce over NVLink by the caller). */
int b200pt_accum_device_ptr(b200pt_ctx* ctx, float** out_ptr);
/* Last per-iteration radiance kernel_color (src/pathtracer.cu:1019-1020), w*h float3, host. */
int b200pt_get_color(b200pt_ctx* ctx, float* dst_host);
/* Tonemap an externally reduced accumulation image: out = tonemap(acc / iter) (src/pathtracer.cu:2526-2530). */
int b200pt_tonemap(b200pt_ctx* ctx, const float* acc_device, uint32_t iter, float* out_device);

/* Primary-visibility debug/parity query: for every pixel of iteration `iter` the closest hit of the camera
 * ray — t (or -1), primitive index, barycentrics b1,b2 — as computed by the traversal kernel
 * (parity target: Intersect, src/pathtracer.cu:214).  hits_host: w*h * 4 floats (t, prim as int bits, b1, b2). */
int b200pt_trace_primary(b200pt_ctx* ctx, const void* camera, uint32_t iter, float* hits_host);

/* Counters of the last render call: [0]=samples, [1]=kernel launches, [2]=rays traced, [3]=wavefront steps,
 * [4]=device ms of the call (CUDA events on the context's stream around all kernels of the call). */
int b200pt_stats(b200pt_ctx* ctx, double* out5);

/* Read-only facts about a context: "lanes" (independent wavefronts), "pool_per_lane" (path slots of one lane),
 * "groups" (primitive groups of the small-scene kernel, 0 if the tree kernel is used), "small_kernel", "lambert_only",
 * "fused" (CTA-local wavefront in use), "wave_blocks", "bin_materials" (shade stage sorts its slots by BSDF class),
 * "emit_boxes" (boxes a MIS ray must pass to reach an emitter; 0 = such rays are not culled), "wide" / "nodes4". */
int b200pt_get_info(b200pt_ctx* ctx, const char* name, int64_t* out_value);

/* Tunables (pool = number of path slots in flight; 0 keeps default).  "reserve_iters" = n pre-sizes the sample planes
 * for batches of up to n iterations, so that no later b200pt_render allocates inside the call.  Scheduling switches —
 * none of them changes a bit of the image: "fused", "bin_materials", "wave_ctas_per_sm", "trace_ctas_per_sm",
 * "refill_below", "graph", "steps_per_poll", "max_batch_bytes", "stage_smem".  Read at b200pt_create from the environment
 * (A/B measurements): B200PT_FUSED, B200PT_BIN_MATERIALS, B200PT_CULL_MIS, B200PT_WIDE, B200PT_LANES, B200PT_WAVE_CTAS,
 * B200PT_STAGE_BYTES. */
int b200pt_set_option(b200pt_ctx* ctx, const char* name, int64_t value);

/* Replaces EndRender (src/pathtracer.cu:2697); frees ALL device memory of the context. */
int b200pt_destroy(b200pt_ctx* ctx);

/* ---- multi-GPU: the image's screen tiles are sharded over the GPUs of one box; after every spp batch
 * the float3 accumulation framebuffers are summed onto one GPU by ONE NCCL reduce over NVLink, inside the library.
 * The reference has no multi-GPU path (one process, one GPU: src/pathtracer.cu:2568); these entry points are what a
 * caller of BeginRender / Render adds to use the whole box.  NCCL is loaded at run time (libnccl.so.2).
 *
 * (1) one process per GPU (torchrun / MPI style): every rank creates its context with shard = {rank, n_ranks, 32, 32},
 *     rank 0 calls b200pt_comm_unique_id and ships the 128 bytes to the others by any means, every rank calls
 *     b200pt_comm_init (collective), then b200pt_render_reduce per batch (collective): this rank's shard is rendered,
 *     the accumulation buffers are reduced onto `root`, and on `root` `output` receives the tonemapped FULL image
 *     (NULL / ignored elsewhere).  Disjoint tiles => one non-zero contributor per pixel => bit-identical to one GPU. */
#define B200PT_COMM_ID_BYTES 128
int b200pt_comm_unique_id(void* id128);
int b200pt_comm_init(b200pt_ctx* ctx, int n_ranks, int rank, const void* id128);
int b200pt_render_reduce(b200pt_ctx* ctx, const void* camera, uint32_t first_iter, uint32_t spp, int reset, int root,
                         float* output, int output_is_device);
/* root only: the reduced linear accumulation image of the last b200pt_render_reduce (w*h float3, host or device). */
int b200pt_reduced_accum(b200pt_ctx* ctx, float* dst, int dst_is_device);

/* (2) one process, n_gpus GPUs (what the BeginRender / Render / EndRender adapter uses when B200PT_GPUS > 1):
 *     b200pt_create_multi makes one sharded context per device (devices == NULL: 0 .. n_gpus-1) and the NCCL
 *     communicators (ncclCommInitAll); b200pt_multi_render renders all shards concurrently, reduces onto the first
 *     device and writes the tonemapped full image to `output` (host, or device memory of the first device). */
typedef struct b200pt_multi b200pt_multi;
int b200pt_create_multi(const b200pt_scene_view* scene, uint32_t width, uint32_t height, float epsilon,
                        int n_gpus, const int* devices, b200pt_multi** out_multi);
int b200pt_multi_render(b200pt_multi* m, const void* camera, uint32_t first_iter, uint32_t spp, int reset,
                        float* output, int output_is_device);
int b200pt_multi_get_accum(b200pt_multi* m, float* dst_host);
int b200pt_multi_stats(b200pt_multi* m, double* out5);      /* as b200pt_stats; samples / launches / rays summed, ms = max */
int b200pt_multi_destroy(b200pt_multi* m);

const char* b200pt_last_error(void);
int b200pt_version(void);

/* ---- scene preparation (SURVEY §8(f) "next" rows; host side, C++) ------------------------------------- */

/* Binned-SAH BVH2 build + flatten, same algorithm and output layout as BVH::build/split/flatten
 * (src/bvh.cpp:16-173): reorders prims into leaf order and emits LinearBVHNode[].
 * prims_in/out: n x 176-B Primitive; nodes_out capacity must be >= 2*n; returns node count in *n_nodes. */
int b200pt_bvh_build(const void* prims_in, int32_t n_prims, void* prims_out, void* nodes_out,
                     int32_t nodes_capacity, int32_t* n_nodes, float* root_box6);

/* The same tree built on the GPU: level-synchronous binned SAH — every node of a level is
 * split at once (bucket histograms with integer atomics, stable segmented partition by prefix sum), then
 * flattened into the reference's pre-order LinearBVHNode layout.  Same arguments and same output as
 * b200pt_bvh_build (byte-identical, except that a zero box coordinate is always +0), host pointers in and
 * out.  timing_ms4 (optional): [0] upload, [1] device build, [2] download (CUDA events), [3] wall clock of the
 * call including allocation. */
int b200pt_bvh_build_gpu(const void* prims_in, int32_t n_prims, void* prims_out, void* nodes_out,
                         int32_t nodes_capacity, int32_t* n_nodes, float* root_box6, int32_t device,
                         double* timing_ms4);

/* The reference's `bvh.cache` file (BVH::LoadOrBuildBVH, src/bvh.cpp:189-217): int total_nodes, int n_prims,
 * float3 root min, float3 root max, Primitive[n_prims] (176 B each, leaf order), LinearBVHNode[total_nodes] (40 B
 * each); native endianness, no magic, no checksum.  Files written here load in the reference and vice versa.
 *   _save: writes exactly that byte stream (temporary file + rename, so a reader never sees a torn file).
 *   _info: reads the 32-byte header only and checks it against the file's size (the reference does not: a
 *          truncated or foreign file there is undefined behaviour) -> B200PT_EINVAL on mismatch.
 *   _load: fills caller arrays (capacities in records); B200PT_ENOMEM if they are too small.
 *   _load_or_build: the reference's entry point — load `path` if it exists and holds exactly n_prims
 *          primitives, else build (GPU builder when device >= 0, host builder when device < 0) and write it.
 *          *was_loaded tells which.  Like the reference, a cache with the right primitive COUNT is trusted. */
int b200pt_bvh_cache_save(const char* path, const void* prims, int32_t n_prims, const void* nodes, int32_t n_nodes,
                          const float* root_box6);
int b200pt_bvh_cache_info(const char* path, int32_t* n_prims, int32_t* n_nodes, float* root_box6);
int b200pt_bvh_cache_load(const char* path, void* prims_out, int32_t prims_capacity, void* nodes_out,
                          int32_t nodes_capacity, int32_t* n_prims, int32_t* n_nodes, float* root_box6);
int b200pt_bvh_load_or_build(const char* path, const void* prims_in, int32_t n_prims, void* prims_out, void* nodes_out,
                             int32_t nodes_capacity, int32_t* n_nodes, float* root_box6, int32_t device,
                             int32_t* was_loaded);

/* Camera constructor arithmetic (src/camera.h:31-47 + Lookat :124-129): fills a 104-B reference Camera. */
int b200pt_camera_init(void* camera104, const float* position3, const float* lookat3, const float* up3,
                       float res_x, float res_y, float distance, float fov, float aperture_radius,
                       float focal_distance, int filmic, int environment, int medium);

/* Light-selection CDF of Scene::Init (src/scene.h:65-82). out has n_lights+1 (+1 if infinite valid) floats. */
int b200pt_light_distribution(const void* lights, int32_t n_lights, const void* infinite_or_null,
                              float* out, int32_t* n_out);

/* Infinite::Init (src/infinite.h:61 -> BBox::boundingSphere, src/bbox.h:98): sets center/radius in place. */
int b200pt_infinite_init(void* infinite72, const float* root_box6);

#ifdef __cplusplus
}
#endif
#endif /* B200PT_H */
