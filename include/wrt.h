/*
 * wrt.h — C ABI of the B200 render back end ("weekend ray tracer", wrt).
 *
 * This is the drop-in boundary for the reference's render hot path.  The reference
 * (j-helland/zig-weekend-raytracer, Zig) has no plugin/FFI seam on that path; the seam is
 *
 *     pub fn render(self: *const Renderer, camera: *const Camera, entity: *const IEntity,
 *                   framebuffer: *Framebuffer) !void                      (src/render.zig:29)
 *
 * reached from Scene.draw (src/scene.zig:57-61) and main (src/main.zig:96).  A Zig maintainer
 * replaces the body of Renderer.render with: walk the IEntity / IMaterial / ITexture tree into
 * the POD arrays below (wrt_scene), then wrt_upload_scene + wrt_render.  The binding follows the
 * only FFI idiom the reference has (extern fn + opaque pointer + int status, as in
 * libs/zstbi/src/zstbi.zig:454-560); see INTEGRATION.md for the Zig stub.
 *
 * Conventions
 *   - every function returns 0 on success, a negative WRT_E_* code on failure; the text of the
 *     last failure is available from wrt_last_error().  No exceptions or callbacks cross the ABI.
 *   - all arrays are caller-owned and copied during the call (ownership is one-directional).
 *   - all reals are IEEE-754 binary64 (math.zig:40 `pub const Real = f64`), vectors are xyz triples.
 *   - indices are uint32_t; WRT_NONE means "null pointer / optional absent".
 *   - a wrt_ctx is bound to ONE CUDA device and is thread-compatible (one caller at a time); any number of contexts may
 *     share a device (every launch carries its own constants as kernel arguments).  Several devices: wrt_group (one
 *     process) or wrt_comm_init + wrt_render_sharded (one process per device).
 *   - there is no CPU fallback: every entry point that renders fails with WRT_E_CUDA if the device
 *     or the kernels are unavailable.
 */
#ifndef WRT_H
#define WRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WRT_ABI_VERSION 3u  /* 3: wrt_stats and wrt_scene_info grew at the tail, wrt_build_trees added (round 2) */
#define WRT_NONE 0xFFFFFFFFu

/* status codes */
enum {
    WRT_OK = 0,
    WRT_E_INVALID = -1,  /* bad argument / malformed scene */
    WRT_E_CUDA = -2,     /* CUDA runtime error (text in wrt_last_error) */
    WRT_E_NOMEM = -3,    /* host or device allocation failed */
    WRT_E_STATE = -4,    /* call order violated (e.g. render before upload) */
    WRT_E_LIMIT = -5     /* scene exceeds a compiled-in limit (transform nesting, ...) */
};

/* IEntity variants (src/entity.zig:17-24) */
enum {
    WRT_ENT_SPHERE = 0,
    WRT_ENT_QUAD = 1,
    WRT_ENT_COLLECTION = 2,
    WRT_ENT_BVH_NODE = 3,
    WRT_ENT_TRANSLATE = 4,
    WRT_ENT_ROTATE_Y = 5
};

/* IMaterial variants (src/material.zig:25-32) */
enum {
    WRT_MAT_LAMBERTIAN = 0,
    WRT_MAT_ISOTROPIC = 1,
    WRT_MAT_METAL = 2,
    WRT_MAT_DIELECTRIC = 3,
    WRT_MAT_DIFFUSE_EMISSIVE = 4
};

/* ITexture variants (src/texture.zig:11-16) */
enum { WRT_TEX_SOLID = 0, WRT_TEX_CHECKER = 1, WRT_TEX_IMAGE = 2 };

/* BVH culling rule used by the device traversal.  TIGHT and REFERENCE return the same closest hit
 * whenever the reference's own boxes are conservative; REFERENCE also reproduces the reference when they are not
 * (AABB.offset shrinks an instanced box, src/math/aabb.zig:52-60, SURVEY.md A.9-4: scene `rtw_final`).
 * wrt_upload_scene compares every reference box with the box recomputed from its subtree (x and y, the axes the reference
 * tests) and counts the ones that do not contain it (wrt_scene_info / wrt_stats .ref_boxes_loose).  The default, AUTO,
 * is the REFERENCE's RESULT at the best speed: TIGHT when that count is 0, REFERENCE otherwise.  wrt_stats.cull_mode_used
 * says which one ran. */
enum {
    WRT_CULL_AUTO = 0,      /* default: TIGHT if every reference box contains its subtree, else REFERENCE */
    WRT_CULL_REFERENCE = 1, /* the reference's test: its cached boxes, x and y only, each axis on its own
                               (src/math/aabb.zig:80-101 + math.zig:186-190) */
    WRT_CULL_TIGHT = 2      /* 3-axis slab test on boxes recomputed from the primitives (fast path); opt-in where
                               ref_boxes_loose > 0: returns hits the reference's loose boxes drop */
};

/* One node of the reference's entity tree (src/entity.zig).  `bbox_*` is the cached AABB.min/max the
 * reference's AABB.hit reads (aabb.zig:23-24,84-85), NOT a recomputed box. */
typedef struct wrt_entity {
    uint32_t kind; /* WRT_ENT_* */
    uint32_t a;    /* sphere/quad: index into spheres[]/quads[]; collection: first slot in children[];
                      bvh_node: left entity; translate/rotate_y: wrapped entity */
    uint32_t b;    /* collection: number of children; bvh_node: right entity */
    uint32_t c;    /* collection: bvh_root entity or WRT_NONE (entity.zig:311) */
    double p[3];   /* translate: offset (entity.zig:71); rotate_y: {sin_theta, cos_theta, 0} (entity.zig:115-116) */
    double bbox_min[3];
    double bbox_max[3];
} wrt_entity;

/* SphereEntity (src/entity.zig:533-543) */
typedef struct wrt_sphere {
    double center[3];
    double radius;
    double movement[3]; /* movement_direction; used iff is_moving (entity.zig:590-594,653-656) */
    uint32_t material;
    uint32_t is_moving;
} wrt_sphere;

/* QuadEntity (src/entity.zig:428-442): the derived fields are passed as the reference computed them
 * in initEntity (entity.zig:444-475) so that device arithmetic starts from identical bits. */
typedef struct wrt_quad {
    double start[3];  /* start_point */
    double u[3];      /* basis.u = axis1 */
    double v[3];      /* basis.v = axis2 */
    double w[3];      /* basis.w = n / dot(n,n) */
    double normal[3]; /* unit normal */
    double offset;    /* dot(normal, start) */
    double area;      /* |axis1 x axis2| */
    uint32_t material;
    uint32_t _pad;
} wrt_quad;

/* IMaterial payloads (src/material.zig:79-226) */
typedef struct wrt_material {
    uint32_t kind;    /* WRT_MAT_* */
    uint32_t texture; /* lambertian / isotropic / diffuse_emissive: texture index */
    double albedo[3]; /* metal */
    double param;     /* metal: fuzz; dielectric: refraction_index */
} wrt_material;

/* ITexture payloads (src/texture.zig:33-119) */
typedef struct wrt_texture {
    uint32_t kind;  /* WRT_TEX_* */
    uint32_t even;  /* checker: tex_even index */
    uint32_t odd;   /* checker: tex_odd index */
    uint32_t image; /* image: index into images[] */
    double color[3];  /* solid */
    double inv_scale; /* checker */
} wrt_texture;

/* zstbi.Image as consumed by Image.getPixel (src/image.zig:23-36): 8-bit interleaved rows. */
typedef struct wrt_image {
    uint32_t width;
    uint32_t height; /* 0 => the reference's magenta debug colour (texture.zig:53-55) */
    uint32_t num_components;
    uint32_t bytes_per_row;
    uint64_t texel_offset; /* byte offset of row 0 inside wrt_scene.texels */
} wrt_image;

typedef struct wrt_scene {
    uint32_t abi_version; /* WRT_ABI_VERSION */
    uint32_t root;        /* entity index of Scene.scene (scene.zig:41) */
    uint32_t lights;      /* entity index of Scene.lights (a collection) or WRT_NONE (scene.zig:42) */
    uint32_t n_entities, n_children, n_spheres, n_quads, n_materials, n_textures, n_images;
    const wrt_entity* entities;
    const uint32_t* children; /* concatenated EntityCollection.entities lists (entity indices) */
    const wrt_sphere* spheres;
    const wrt_quad* quads;
    const wrt_material* materials;
    const wrt_texture* textures;
    const wrt_image* images;
    const uint8_t* texels;
    uint64_t texel_bytes;
} wrt_scene;

/* The camera view the render jobs read: exactly the RenderThreadContext fields (src/render.zig:94-102)
 * with the Viewport members they dereference (src/camera.zig:112-114).  Computed on the host in f64 by
 * Camera.init / Viewport.init (camera.zig:61-90,117-157) so device primary rays are bit-identical. */
typedef struct wrt_camera {
    double position[3];
    double pixel00_loc[3];
    double pixel_delta_u[3];
    double pixel_delta_v[3];
    double defocus_disk_u[3];
    double defocus_disk_v[3];
    uint32_t is_depth_of_field;
    uint32_t _pad;
} wrt_camera;

typedef struct wrt_params {
    uint32_t width;   /* framebuffer.num_cols */
    uint32_t height;  /* framebuffer.num_rows */
    uint32_t samples_per_pixel;    /* Renderer.samples_per_pixel   (--samples_per_pixel) */
    uint32_t max_ray_bounce_depth; /* Renderer.max_ray_bounce_depth (--ray_bounce_max_depth) */
    double background_color[3];    /* Renderer.background_color */
    double clear_color[3];         /* Renderer.clear_color */
    uint64_t seed;                 /* keys the counter-based RNG; the reference seeds from getrandom (rng.zig:16-26) */
    /* sharding (multi-GPU): this context renders rows  r = row_shard_index + k*row_shard_count  only and
     * writes them densely (row k of the output = image row r).  {0,1} = whole image. */
    uint32_t row_shard_index;
    uint32_t row_shard_count;
    /* sample range [sample_begin, sample_end) of each pixel, still scaled by 1/samples_per_pixel;
     * {0,0} means the full range.  Used for sample-range sharding / progressive accumulation. */
    uint32_t sample_begin;
    uint32_t sample_end;
    uint32_t cull_mode; /* WRT_CULL_* */
    uint32_t flags;     /* WRT_FLAG_* */
} wrt_params;

#define WRT_FLAG_NO_CLEAR 1u       /* add onto the existing framebuffer contents instead of clear_color */
#define WRT_FLAG_DISABLE_DOF 2u    /* treat the camera as a pinhole (gate-1 dumps) */
#define WRT_FLAG_FORCE_LANE 4u     /* closest-hit scan per lane even for small programs (default: chosen by program size) */
#define WRT_FLAG_FORCE_PACKET 8u   /* warp-uniform packet scan even for large programs */
#define WRT_FLAG_ENGINE_MEGAKERNEL 16u /* force the persistent megakernel (default: chosen by the amount of work) */
#define WRT_FLAG_ENGINE_WAVEFRONT 32u  /* force the wavefront engine (path pool + per-material queues in HBM) */
#define WRT_FLAG_ENGINE_SYNC 64u       /* phase-synchronous megakernel: one 16-warp block per SM, block barriers between phases */
#define WRT_FLAG_ENGINE_REGROUP 128u   /* phase-synchronous megakernel + per-material regrouping of the block's paths in shared memory */
#define WRT_FLAG_SHARD_SAMPLES 256u    /* wrt_group_render / wrt_render_sharded: split the sample range instead of the rows */
/* Sampler upgrade the reference sketches (src/math/sampler.zig:203-247): every random decision of a path (lens, time,
 * mixture choice, light pick, direction) takes the next Owen-scrambled Sobol dimension (get1D / get2D, dimensions 2, 3, ...
 * of the pixel sample's Sobol index, wrapping at 1024) instead of the pseudo-random stream.  Changes the estimator's noise,
 * not its mean; default off (parity mode). */
#define WRT_FLAG_SAMPLER_SOBOL 512u
/* Sample chunks: a pixel's samples are summed in order inside a chunk and the chunk sums are added in order, so the
 * chunk count fixes the last bits of the frame.  By default it is chosen from the full frame size, the sample count and
 * the engine (never from the shard or the GPU), so a frame is bit-identical on 1..8 GPUs; WRT_FLAG_CHUNKS(n), n in
 * 1..255, pins it — two engines given the same n produce the same bits. */
#define WRT_FLAG_CHUNKS(n) (((uint32_t)(n) & 0xFFu) << 24)
/* the same two switches for wrt_trace_rays, OR-ed into its cull_mode argument */
#define WRT_TRAV_FORCE_LANE 0x100u
#define WRT_TRAV_FORCE_PACKET 0x200u

typedef struct wrt_stats {
    uint64_t paths;          /* camera samples started */
    uint64_t rays;           /* closest-hit queries issued by the integrator (render.zig:215) */
    double render_ms;        /* device time of the last wrt_render* call (CUDA events) */
    double kernel_ms;        /* device time of the path-tracing kernel alone */
    double upload_ms;        /* host wall time of the last wrt_upload_scene */
    uint32_t kernel_launches;/* kernels launched by the last wrt_render* call */
    uint32_t program_ops;    /* size of the compiled traversal program */
    uint32_t n_prims;        /* leaf primitives in DFS order */
    uint32_t cull_mode_used; /* WRT_CULL_REFERENCE or WRT_CULL_TIGHT: what the last render / gate call resolved AUTO to */
    uint64_t traversal_steps;/* node records + ops visited by the per-lane ordered traversal (0 for the packet scan) */
    uint32_t ref_boxes_loose;/* bvh_node boxes of the uploaded scene that do not contain their subtree in x / y */
    uint32_t n_devices;      /* devices that took part in the last render (1 for a plain wrt_ctx) */
    double gather_ms;        /* device time of the shard gather of the last wrt_group_render (0 otherwise) */
    double kernel_ms_min;    /* min / max of kernel_ms over the devices of a group (== kernel_ms for one device) */
    double kernel_ms_max;
    double tree_build_ms;    /* time of the ordered traversal's tree build inside the last upload (device events or host wall) */
    uint32_t tree_build_device; /* 1: the trees were built on the device (wrt_build.cu), 0: on the host threads */
    uint32_t n_tree_records; /* records of the trees the ordered traversal walks (four-wide or child-pair) */
} wrt_stats;

typedef struct wrt_ctx wrt_ctx;

/* exported symbols (the library is built with -fvisibility=hidden) */
#if defined(__GNUC__)
#define WRT_API __attribute__((visibility("default")))
#else
#define WRT_API
#endif

/* Lifetime ------------------------------------------------------------------------------------ */
WRT_API int wrt_create(int cuda_device, wrt_ctx** out);
WRT_API void wrt_destroy(wrt_ctx* ctx);
/* Text of the most recent failure on `ctx` (or of the last failed wrt_create when ctx == NULL). */
WRT_API const char* wrt_last_error(const wrt_ctx* ctx);
WRT_API uint32_t wrt_abi_version(void);

/* Scene --------------------------------------------------------------------------------------- */
/* Validates the tree, assigns primitive ids (DFS order), compiles the stack-less
 * traversal program and copies everything to the device. */
WRT_API int wrt_upload_scene(wrt_ctx* ctx, const wrt_scene* scene);

/* Host-only: compiles `scene` exactly as wrt_upload_scene does (no device, no context) and checks the result's structure —
 * skip links, transform nesting, the packet program against the full one, every ordered-traversal tree reaching each
 * primitive of its BVH exactly once.  Returns WRT_OK, or an error code with the reason in err[0..err_cap). */
typedef struct wrt_scene_info {
    uint32_t n_ops;          /* ops of the traversal program (OP_END included) */
    uint32_t n_ops_packet;   /* ops of the pruned program the packet scan reads */
    uint32_t n_prims;
    uint32_t n_boxes;        /* bvh_node + instance bounds */
    uint32_t n_tree_records; /* child-pair records of the ordered-traversal trees */
    uint32_t tree_depth;     /* deepest tree, in records */
    uint32_t max_nesting;    /* stack bound of the ordered traversal */
    uint32_t n_lights;
    uint32_t ref_boxes_loose;/* bvh_node boxes that do not contain their subtree in x / y (0 => AUTO culling = TIGHT) */
    uint32_t stack_depth;    /* exact worst-case stack use of the ordered traversal on the rebuilt trees */
    uint32_t compact_stack;  /* 1: one tree of single-primitive leaves without transforms — the wavefront's extend kernel uses 8-byte stack entries */
    uint32_t quantised_records; /* number of 64-byte quantised four-wide records (0: the scene walks the binary32 records) */
} wrt_scene_info;
WRT_API int wrt_check_scene(const wrt_scene* scene, wrt_scene_info* info, char* err, size_t err_cap);

/* The tree build on its own (replaces BVHNodeEntity.init, src/entity.zig:226-259, for the ordered traversal): compiles
 * `scene` and builds the trees over the leaves of every reference BVH — binned surface-area heuristic, then the four-wide
 * collapse — on CUDA device `cuda_device`, or on the host threads when cuda_device < 0.  Both builders write the same
 * bytes.  records2 / records4 (may be NULL) receive up to cap2 / cap4 BYTES of the child-pair (64 B) and four-wide (128 B)
 * records; the counts come back in `info` either way.  wrt_upload_scene runs the same code (device build for scenes of
 * >= 32768 leaves, WRT_DEVICE_BUILD=0/1 overrides). */
typedef struct wrt_tree_info {
    uint32_t n_records2;     /* child-pair records (reference topology records included) */
    uint32_t n_records4;     /* four-wide records */
    uint32_t stack_depth;    /* exact worst-case stack use of the ordered traversal */
    uint32_t use_wide;       /* the ordered traversal walks the four-wide records */
    uint32_t on_device;      /* where the build ran */
    uint32_t max_nesting;
    double build_ms;         /* device events (device build) or host wall time (host build), tree build only */
    double total_ms;         /* host wall time of the whole call */
} wrt_tree_info;
WRT_API int wrt_build_trees(const wrt_scene* scene, int cuda_device, wrt_tree_info* info, void* records2, size_t cap2,
                    void* records4, size_t cap4, char* err, size_t err_cap);

/* Render (replaces Renderer.render, src/render.zig:29-74) ------------------------------------- */
/* Host framebuffer: `framebuffer` points at Framebuffer.buffer ([]Color, camera.zig:14); one pixel every
 * `pixel_stride_bytes` bytes (= @sizeOf(Vec3): 32 with 4 lanes, 64 with 8; >= 24), lanes 0..2 = R,G,B
 * linear radiance, remaining lanes written as 0.  Result per pixel: clear_color + mean of rayColor. */
WRT_API int wrt_render(wrt_ctx* ctx, const wrt_camera* cam, const wrt_params* params, void* framebuffer,
               size_t pixel_stride_bytes);
/* Same, but `d_framebuffer` is a device pointer on the context's device (dense rows of this shard);
 * no device->host copy is made.  Used under NCCL gathers. */
WRT_API int wrt_render_device(wrt_ctx* ctx, const wrt_camera* cam, const wrt_params* params, void* d_framebuffer,
                      size_t pixel_stride_bytes);
/* Quantise the frame produced by the last wrt_render* call exactly like encodeColor
 * (src/writer/writer.zig:68-94: NaN->0, sqrt, clamp [0,0.999], *256, truncate) into 3 bytes/pixel (host). */
WRT_API int wrt_encode_rgb8(wrt_ctx* ctx, uint8_t* rgb_out);

/* Gates / diagnostics ------------------------------------------------------------------------- */
/* The PPM writer's body on the device (writer/writer.zig:16-123): formats an RGB8 frame as the reference's file image —
 * "P3\n{W} {H}\n255\n" followed by one "{r} {g} {b}\n" line per pixel, packed — into `out`.  rgb8 == NULL formats the
 * frame of the last wrt_render* call (quantised on the device by the resolve pass; width x height must be that frame);
 * otherwise width*height*3 host bytes are uploaded first (e.g. a frame gathered from several GPUs).  `capacity` must be
 * at least header + 12 bytes per pixel, the size the reference gives its file (writer.zig:20); the bytes between
 * *content_bytes and that size are zeroed, as in the reference's mmap'ed file.  `out` may be the mapping itself. */
WRT_API int wrt_format_ppm(wrt_ctx* ctx, const uint8_t* rgb8, uint32_t width, uint32_t height, uint8_t* out, uint64_t capacity,
                   uint64_t* content_bytes);

/* Gate 1: closest hit of the PRIMARY ray of samples [0, n_samples) of every pixel (row-major, sample
 * innermost): primitive id (WRT_NONE = miss) and t.  Depth of field is disabled. */
WRT_API int wrt_primary_hits(wrt_ctx* ctx, const wrt_camera* cam, const wrt_params* params, uint32_t n_samples,
                     uint32_t* prim_ids, double* t);
/* Closest hit of arbitrary rays against the uploaded scene with range (tmin, +inf):
 * origins/directions are n xyz triples; outputs may be NULL.  point/normal are n xyz triples. */
WRT_API int wrt_trace_rays(wrt_ctx* ctx, const double* origins, const double* directions, uint64_t n, double tmin,
                   uint32_t cull_mode, uint32_t* prim_ids, double* t, double* point, double* normal,
                   double* uv, uint32_t* front_face);
/* Sobol pixel sampling (src/math/sampler.zig:197-201,222-234,267-298): for each listed pixel and sample
 * index returns the global Sobol index and the [0,1) pixel offsets. */
WRT_API int wrt_sobol_pixel_samples(wrt_ctx* ctx, uint32_t width, uint32_t height, const uint32_t* cols,
                            const uint32_t* rows, const uint32_t* sample_idx, uint64_t n, uint64_t* sobol_index,
                            double* offsets_xy);
/* Sobol higher dimensions (sampler.zig:203-247): sampleDimension(dim) for a given global index with the
 * noop or owen_fast randomiser (Murmur2 per-dimension seed + Laine-Karras hash, sampler.zig:39-53). */
WRT_API int wrt_sobol_dimension_samples(wrt_ctx* ctx, const uint64_t* sobol_index, const uint32_t* dimension,
                                uint64_t n, uint32_t owen_fast, uint32_t seed, float* out);
WRT_API int wrt_get_stats(const wrt_ctx* ctx, wrt_stats* out);
/* Measures the device's binary64 FMA issue rate (thread-level DFMA per second) with a micro-kernel: the denominator of
 * the issue-bound roofline for the cache-resident configs. */
WRT_API int wrt_fp64_issue_peak(wrt_ctx* ctx, double* fma_per_second);
/* The same probe on the binary32 pipe (the conservative culler runs there). */
WRT_API int wrt_fp32_issue_peak(wrt_ctx* ctx, double* fma_per_second);

/* Multi-GPU ----------------------------------------------------------------------------------- */
/* The frame shards like the reference's own job fan-out (src/render.zig:55-73: disjoint row segments, no
 * synchronisation): device r of n renders image rows r, r+n, r+2n, ... (interleaved for load balance), the shards are
 * gathered into one device's frame with NCCL point-to-point transfers over NVLink (3 binary64 lanes per pixel on the
 * wire), and one fused pass writes the caller's framebuffer layout and the RGB8 frame.  Random numbers and Sobol indices
 * are keyed by the global pixel, and the sample-chunk summation tree by the full frame, so the assembled frame is
 * bit-identical for every n.  With WRT_FLAG_SHARD_SAMPLES every device renders ALL rows for 1/n of the sample range and the
 * partial means are added with ncclReduce(ncclSum, ncclDouble) — for small frames at very high sample counts; the sum
 * order then depends on n (last-bit differences).
 *
 * (1) one process, n devices: a wrt_group owns one wrt_ctx per device and an ncclCommInitAll communicator clique. */
typedef struct wrt_group wrt_group;
WRT_API int wrt_group_create(const int* device_ids, int n_devices, wrt_group** out);
WRT_API void wrt_group_destroy(wrt_group* g);
WRT_API const char* wrt_group_last_error(const wrt_group* g); /* g == NULL: last failed wrt_group_create */
WRT_API int wrt_group_size(const wrt_group* g);
/* The context of member i (0 = the root that holds the assembled frame): for wrt_format_ppm, wrt_get_stats, gates. */
WRT_API wrt_ctx* wrt_group_ctx(wrt_group* g, int i);
WRT_API int wrt_group_upload_scene(wrt_group* g, const wrt_scene* scene); /* compiled once, copied to every device */
/* Replaces Renderer.render on n devices; params->row_shard_* are ignored (the group sets them).  `framebuffer` as in
 * wrt_render; NULL leaves the assembled frame on the root device (wrt_group_encode_rgb8 / wrt_format_ppm read it). */
WRT_API int wrt_group_render(wrt_group* g, const wrt_camera* cam, const wrt_params* params, void* framebuffer,
                             size_t pixel_stride_bytes);
WRT_API int wrt_group_encode_rgb8(wrt_group* g, uint8_t* rgb_out);
/* Whole-job numbers of the last wrt_group_render: rays/paths summed over the devices, render_ms = max over the devices
 * + gather, kernel_ms_min/max over the devices, gather_ms, n_devices. */
WRT_API int wrt_group_get_stats(const wrt_group* g, wrt_stats* out);

/* (2) one process per device (torchrun / MPI): rank 0 makes an id, the caller broadcasts it by its own means, every rank
 * attaches its context; wrt_render_sharded then renders this rank's shard and gathers to rank 0 inside the library. */
#define WRT_COMM_ID_BYTES 128
WRT_API int wrt_comm_unique_id(uint8_t id[WRT_COMM_ID_BYTES]);
WRT_API int wrt_comm_init(wrt_ctx* ctx, const uint8_t id[WRT_COMM_ID_BYTES], int rank, int n_ranks);
/* Collective over the ranks of wrt_comm_init.  Rank 0: `framebuffer` = host buffer of the FULL frame, or NULL to leave it
 * on the device; other ranks: ignored.  params->row_shard_* are ignored (rank / n_ranks are used). */
WRT_API int wrt_render_sharded(wrt_ctx* ctx, const wrt_camera* cam, const wrt_params* params, void* framebuffer,
                               size_t pixel_stride_bytes);

#ifdef __cplusplus
}
#endif
#endif /* WRT_H */
