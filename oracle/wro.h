/*
 * wro.h — public C API of the ORACLE (test infrastructure, not product code).
 *
 * "wro" = weekend-raytracer oracle: a CPU restatement of the reference's render hot path
 * (j-helland/zig-weekend-raytracer, src/render.zig et al.), used ONLY as the checker for the CUDA back end
 * (tests/, __graft_entry__.smoke()) and as the timed CPU baseline (bench.py cpu_baseline / --impl reference).
 * The product (libwrt.so, include/wrt.h) never links or calls it.
 *
 * PARITY UNPINNED: the reference is Zig and cannot be built here (no toolchain), and its own tests hold no
 * golden vector for this path.  The oracle is pinned instead by (i) the reference's math /
 * writer known answers restated in tests/, (ii) Sobol van-der-Corput known answers and the pixel-containment
 * invariant, (iii) the BVH topologies of SURVEY.md A.10, (iv) BVH-vs-brute-force and furnace self-checks.
 */
#ifndef WRO_H
#define WRO_H

#include <stddef.h>
#include <stdint.h>

#include "../include/wrt.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wro_scene wro_scene;

/* Image bytes a scene builder may need (decoded by the caller: reference zstbi / PIL / procedural). */
typedef struct wro_image_in {
    const char* name; /* "wap.jpg", "me.jpg", "earth.png" */
    uint32_t width, height, num_components;
    const uint8_t* data;
} wro_image_in;

/* Scene catalogue: "balls", "shrek_quads", "emissive", "cornell_box", "rtw_final" (src/scene.zig:18-24) plus the
 * harness scenes "earth" and "synthetic" (SURVEY.md §8d).  `seed` replaces the reference's getrandom-seeded
 * scene RNG; `n_prims` is only read by "synthetic".  Returns NULL on unknown names. */
wro_scene* wro_scene_build(const char* name, uint64_t seed, uint32_t n_prims, const wro_image_in* images,
                           uint32_t n_images);
/* Rebuild a pointer tree from flat arrays (any producer of include/wrt.h scenes). */
wro_scene* wro_scene_from_flat(const wrt_scene* flat);
void wro_scene_destroy(wro_scene* s);
void wro_scene_set_no_cull(wro_scene* s, int no_cull);

/* Scene.camera viewed through a W x H framebuffer (Camera.init + Viewport.init). */
void wro_scene_camera(const wro_scene* s, uint32_t width, uint32_t height, wrt_camera* out);
void wro_scene_camera_desc(const wro_scene* s, double out[12]); /* from, at, up, vfov, focus, defocus */
void wro_scene_background(const wro_scene* s, double out[3]);
uint32_t wro_scene_n_prims(const wro_scene* s);

/* Flat export: what the Zig shim's tree walk produces.  The view stays valid until wro_flat_free. */
typedef struct wro_flat {
    wrt_scene scene;
    void* owner;
} wro_flat;
int wro_scene_flatten(const wro_scene* s, wro_flat* out);
void wro_flat_free(wro_flat* f);

/* DFS leaf order description for topology known-answer tests: for each prim id, kind (0 sphere, 1 quad),
 * material index, and the centre of its reference AABB. */
int wro_scene_prim_table(const wro_scene* s, uint32_t* kinds, uint32_t* materials, double* centers_xyz);

typedef struct wro_render_stats {
    uint64_t paths, rays;
    double seconds;
    uint32_t threads;
} wro_render_stats;

enum { WRO_RNG_MODE_REFERENCE = 0, WRO_RNG_MODE_COUNTER = 1 };

/* Renderer.render (src/render.zig:29-74): same job decomposition (row x 32-column blocks) on n_threads. */
int wro_render(const wro_scene* s, const wrt_camera* cam, const wrt_params* params, int rng_mode,
               uint32_t n_threads, void* framebuffer, size_t pixel_stride_bytes, wro_render_stats* stats);
int wro_primary_hits(const wro_scene* s, const wrt_camera* cam, const wrt_params* params, uint32_t n_samples,
                     uint32_t n_threads, uint32_t* prim_ids, double* t);
int wro_trace_rays(const wro_scene* s, const double* origins, const double* directions, uint64_t n, double tmin,
                   uint32_t* prim_ids, double* t, double* point, double* normal, double* uv,
                   uint32_t* front_face);
/* lights.pdfValue / sampleDirectionToSurface hooks on given origins (EntityPdf, src/pdf.zig:68-90) */
int wro_light_pdf_values(const wro_scene* s, const double* origins, const double* directions, uint64_t n,
                         double* out);

int wro_sobol_pixel_samples(uint32_t width, uint32_t height, const uint32_t* cols, const uint32_t* rows,
                            const uint32_t* sample_idx, uint64_t n, uint64_t* sobol_index, double* offsets_xy);
int wro_sobol_dimension_samples(const uint64_t* sobol_index, const uint32_t* dimension, uint64_t n,
                                uint32_t owen_fast, uint32_t seed, float* out);
/* get1D/get2D sequence of one pixel sample (sampler.zig:203-220): n values from get1D after startPixelSample */
int wro_sobol_get1d_sequence(uint32_t width, uint32_t height, uint32_t col, uint32_t row, uint32_t sample_idx,
                             uint32_t owen_fast, uint32_t seed, uint32_t n, double* out);

/* writer.zig:68-94 encodeColor, :96-100 sizeOfLine, :107-114 sizeOfDigit */
void wro_encode_color(const double rgb[3], uint8_t out[3]);
void wro_encode_image(const void* framebuffer, size_t pixel_stride_bytes, uint64_t n_pixels, uint8_t* rgb_out);
uint32_t wro_size_of_line(const uint8_t pixel[3]);
uint32_t wro_size_of_digit(uint8_t digit);

/* counter RNG stream (shared definition with the device): 64 bits of (seed, pixel, sample, draw) */
uint64_t wro_counter_rng_bits(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t draw);

/* math known answers (src/math/math.zig tests) */
void wro_math_cross(const double u[3], const double v[3], double out[3]);
double wro_math_dot(const double u[3], const double v[3]);
double wro_math_length(const double u[3]);
void wro_math_normalize(const double u[3], double out[3]);

#ifdef __cplusplus
}
#endif
#endif /* WRO_H */
