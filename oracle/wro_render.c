/*
 * wro_render.c — ORACLE (test infrastructure, not product code).
 * CPU restatement of the render hot path: src/render.zig (job fan-out, sample loop, recursive rayColor),
 * src/camera.zig (Camera.init / Viewport.init), src/material.zig, src/pdf.zig, src/writer/writer.zig:68-123
 * (encodeColor, sizeOfLine, sizeOfDigit).
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "wro.h"
#include "wro_scene.h"
#include "wro_sobol.h"

/* ---- camera (camera.zig:61-90, 117-157) ------------------------------------------------------------- */
void wro_camera_view(const wro_camera_desc* c, uint32_t image_width, uint32_t image_height, wrt_camera* out) {
    /* Camera.init */
    v3 w = v3_normalize(v3_sub(c->look_from, c->look_at));
    v3 u = v3_normalize(v3_cross(c->view_up, w));
    v3 v = v3_cross(w, u);
    double defocus_radius = c->lens_focus_dist * tan((c->defocus_angle_degrees / 2.0) * (WRO_PI / 180.0));
    v3 disk_u = v3_scale(u, defocus_radius);
    v3 disk_v = v3_scale(v, defocus_radius);
    /* Framebuffer.getAspectRatio, camera.zig:35-39 */
    double aspect_ratio = (double)image_width / (double)image_height;
    /* Viewport.init */
    double theta = c->fov_vertical * (WRO_PI / 180.0);
    double h = tan(theta / 2.0);
    double viewport_height = 2.0 * h * c->lens_focus_dist;
    double viewport_width = viewport_height * aspect_ratio;
    v3 viewport_u = v3_scale(u, viewport_width);
    v3 viewport_v = v3_scale(v, -viewport_height);
    v3 upper_left = v3_sub(v3_sub(v3_sub(c->look_from, v3_scale(w, c->lens_focus_dist)), v3_div(viewport_u, v3_splat(2))),
                           v3_div(viewport_v, v3_splat(2)));
    v3 du = v3_div(viewport_u, v3_splat((double)image_width));
    v3 dv = v3_div(viewport_v, v3_splat((double)image_height));
    v3 p00 = v3_add(upper_left, v3_scale(v3_add(du, dv), 0.5));

    memset(out, 0, sizeof *out);
    out->position[0] = c->look_from.x; out->position[1] = c->look_from.y; out->position[2] = c->look_from.z;
    out->pixel00_loc[0] = p00.x; out->pixel00_loc[1] = p00.y; out->pixel00_loc[2] = p00.z;
    out->pixel_delta_u[0] = du.x; out->pixel_delta_u[1] = du.y; out->pixel_delta_u[2] = du.z;
    out->pixel_delta_v[0] = dv.x; out->pixel_delta_v[1] = dv.y; out->pixel_delta_v[2] = dv.z;
    out->defocus_disk_u[0] = disk_u.x; out->defocus_disk_u[1] = disk_u.y; out->defocus_disk_u[2] = disk_u.z;
    out->defocus_disk_v[0] = disk_v.x; out->defocus_disk_v[1] = disk_v.y; out->defocus_disk_v[2] = disk_v.z;
    out->is_depth_of_field = (c->defocus_angle_degrees > 0.0);
}

static v3 arr3(const double a[3]) { return v3_make(a[0], a[1], a[2]); }

/* ---- materials (material.zig) and pdfs (pdf.zig) ---------------------------------------------------- */
enum { PDF_NONE = 0, PDF_COSINE, PDF_SPHERE };

typedef struct scatter_record { /* material.zig:11-22 */
    v3 attenuation;
    int pdf_kind;
    onb pdf_basis; /* CosinePdf.basis */
    int has_specular;
    ray ray_specular;
} scatter_record;

/* material.zig:34-45 + :88-96 */
static v3 material_emitted(const wro_material* m, const wro_hit* rec) {
    if (m->kind != WRO_MAT_DIFFUSE_EMISSIVE) return v3_splat(0);
    if (!rec->front_face) return v3_splat(0);
    return wro_texture_value(m->texture, rec->uv, rec->point);
}

/* material.zig:221-225 Schlick reflectance (std.math.pow(f64, x, 5)) */
static double reflectance(double refraction_index, double cosine) {
    double r0 = (1 - refraction_index) / (1 + refraction_index);
    r0 *= r0;
    return r0 + (1 - r0) * pow(1 - cosine, 5);
}

/* material.zig:47-58 dispatch; :108-116 lambertian, :136-143 isotropic, :163-178 metal, :190-218 dielectric */
static int material_scatter(const wro_material* m, const ray* ray_in, const wro_hit* rec, wro_rng* rng, scatter_record* sr) {
    switch (m->kind) {
        case WRO_MAT_LAMBERTIAN:
            sr->attenuation = wro_texture_value(m->texture, rec->uv, rec->point);
            sr->pdf_kind = PDF_COSINE;
            sr->pdf_basis = onb_init(rec->normal); /* CosinePdf.initPdf, pdf.zig:51-56 */
            return 1;
        case WRO_MAT_ISOTROPIC:
            sr->attenuation = wro_texture_value(m->texture, rec->uv, rec->point);
            sr->pdf_kind = PDF_SPHERE;
            return 1;
        case WRO_MAT_METAL: {
            double blur = wro_clamp(m->param, 0.0, 1.0);
            v3 refl = v3_reflect(ray_in->direction, rec->normal); /* unnormalised incoming direction (A.9-8) */
            wro_rng_slot(rng, 2);
            v3 dir = v3_add(refl, v3_scale(wro_sample_unit_sphere(rng), blur));
            sr->attenuation = m->albedo;
            sr->pdf_kind = PDF_NONE;
            sr->has_specular = 1;
            sr->ray_specular.origin = rec->point;
            sr->ray_specular.direction = dir;
            sr->ray_specular.time = ray_in->time;
            return v3_dot(dir, rec->normal) > 0.0;
        }
        case WRO_MAT_DIELECTRIC: {
            double index = rec->front_face ? 1.0 / m->param : m->param;
            v3 in_unit = v3_normalize(ray_in->direction);
            double cos_theta = fmin(v3_dot(v3_neg(in_unit), rec->normal), 1.0);
            double sin_theta = sqrt(1 - cos_theta * cos_theta);
            v3 dir;
            wro_rng_slot(rng, 0);
            /* `or` short-circuits: the uniform is drawn only when total internal reflection does not apply */
            if (index * sin_theta > 1.0 || reflectance(m->param, cos_theta) > wro_rng_float(rng))
                dir = v3_reflect(in_unit, rec->normal);
            else
                dir = v3_refract(in_unit, rec->normal, index);
            sr->attenuation = v3_make(1, 1, 1);
            sr->pdf_kind = PDF_NONE;
            sr->has_specular = 1;
            sr->ray_specular.origin = rec->point;
            sr->ray_specular.direction = dir;
            sr->ray_specular.time = ray_in->time;
            return 1;
        }
        default: return 0; /* diffuse_emissive has no scatter method (material.zig:51-57) */
    }
}

/* material.zig:60-71 dispatch; :118-125 lambertian, :145-150 isotropic */
static double material_scattering_pdf(const wro_material* m, const wro_hit* rec, const ray* scattered) {
    switch (m->kind) {
        case WRO_MAT_LAMBERTIAN: {
            v3 light_dir = v3_normalize(scattered->direction);
            double cos_theta = v3_dot(rec->normal, light_dir);
            return fmax(0.0, cos_theta / WRO_PI);
        }
        case WRO_MAT_ISOTROPIC: return 1.0 / (4.0 * WRO_PI);
        default: return 0.0;
    }
}

/* pdf.zig:58-61 CosinePdf.value, :36-38 SpherePdf.value */
static double surface_pdf_value(int kind, const onb* basis, v3 direction) {
    if (kind == PDF_COSINE) {
        double cos_theta = v3_dot(v3_normalize(direction), basis->w);
        return fmax(0, cos_theta / WRO_PI);
    }
    return 1.0 / (4.0 * WRO_PI);
}
/* pdf.zig:63-65 CosinePdf.generate, :40-42 SpherePdf.generate */
static v3 surface_pdf_generate(int kind, const onb* basis, wro_rng* rng) {
    if (kind == PDF_COSINE) return onb_transform(basis, wro_sample_cosine_direction_z(rng));
    return wro_sample_unit_sphere(rng);
}

/* ---- rayColor (render.zig:188-289) ------------------------------------------------------------------- */
typedef struct trace_ctx {
    const wro_scene* scene;
    v3 background;
    wro_rng* rng;
    uint64_t rays;
    uint32_t max_depth;
} trace_ctx;

static v3 ray_color(trace_ctx* tc, const ray* r, uint32_t depth) {
    if (depth == 0) return v3_splat(0);
    wro_rng_set_base(tc->rng, 4u + 4u * (tc->max_depth - depth)); /* counter mode: this bounce's draw slots */
    const double ray_correction_factor = 1e-4;
    wro_hit rec;
    memset(&rec, 0, sizeof rec);
    rec.t = INFINITY;
    ival trange = {ray_correction_factor, INFINITY};
    tc->rays++;
    if (!wro_entity_hit(tc->scene, tc->scene->root, r, trange, &rec)) return tc->background;

    scatter_record sr;
    memset(&sr, 0, sizeof sr);
    sr.attenuation = v3_make(1, 1, 1);
    const wro_material* material = rec.material;

    v3 emission = material_emitted(material, &rec);
    if (!material_scatter(material, r, &rec, tc->rng, &sr)) return emission;

    if (sr.has_specular) {
        v3 c = ray_color(tc, &sr.ray_specular, depth - 1);
        return v3_mul(sr.attenuation, c);
    }

    ray scattered;
    scattered.origin = rec.point;
    scattered.time = r->time;
    double pdf_value;
    const wro_entity* lights = tc->scene->lights;
    if (lights) {
        /* MixturePdf(EntityPdf(lights, point), material pdf): pdf.zig:99-118, render.zig:254-263 */
        wro_rng_slot(tc->rng, 0);
        double p = wro_rng_float(tc->rng);
        if (p < 0.5) {
            wro_rng_slot(tc->rng, 1);
            scattered.direction = wro_entity_sample_direction(lights, tc->rng, rec.point);
        } else {
            wro_rng_slot(tc->rng, 2);
            scattered.direction = surface_pdf_generate(sr.pdf_kind, &sr.pdf_basis, tc->rng);
        }
        double p1 = wro_entity_pdf_value(tc->scene, lights, rec.point, scattered.direction);
        double p2 = surface_pdf_value(sr.pdf_kind, &sr.pdf_basis, scattered.direction);
        pdf_value = 0.5 * p1 + 0.5 * p2;
    } else {
        /* render.zig:264-269: a cosine pdf about the normal whatever the material asked for */
        onb basis = onb_init(rec.normal);
        wro_rng_slot(tc->rng, 2);
        scattered.direction = surface_pdf_generate(PDF_COSINE, &basis, tc->rng);
        pdf_value = surface_pdf_value(PDF_COSINE, &basis, scattered.direction);
    }

    v3 scatter_color = ray_color(tc, &scattered, depth - 1);
    double sp = material_scattering_pdf(material, &rec, &scattered);
    scatter_color = v3_mul(scatter_color, v3_scale(sr.attenuation, sp));
    scatter_color = v3_div(scatter_color, v3_splat(pdf_value));
    return v3_add(emission, scatter_color);
}

/* ---- sampleRay (render.zig:144-174) ------------------------------------------------------------------ */
static ray sample_ray(const wrt_camera* cam, int dof, wro_rng* rng, wro_sobol_sampler* sampler, uint32_t col,
                      uint32_t row, uint32_t sample_idx) {
    double offset[2];
    wro_sobol_start_pixel_sample(sampler, col, row, sample_idx);
    wro_sobol_get_pixel_2d(sampler, offset);
    v3 p00 = arr3(cam->pixel00_loc), du = arr3(cam->pixel_delta_u), dv = arr3(cam->pixel_delta_v);
    v3 sample = v3_add(v3_add(p00, v3_scale(du, (double)col + offset[0])), v3_scale(dv, (double)row + offset[1]));
    v3 origin = arr3(cam->position);
    if (dof) { /* sampleDefocusDisk, render.zig:182-185 */
        v3 p = wro_sample_unit_disk_xy(rng, 1.0);
        origin = v3_add(v3_add(arr3(cam->position), v3_scale(arr3(cam->defocus_disk_u), p.x)),
                        v3_scale(arr3(cam->defocus_disk_v), p.y));
    }
    ray r;
    r.origin = origin;
    r.direction = v3_sub(sample, origin);
    if (rng) wro_rng_slot(rng, 2);
    r.time = rng ? wro_rng_float(rng) : 0.0;
    return r;
}

/* ---- Renderer.render (render.zig:29-74) + rayColorLine (:107-141) ------------------------------------ */
typedef struct render_job_ctx {
    const wro_scene* scene;
    const wrt_camera* cam;
    const wrt_params* params;
    int rng_mode;
    double* fb;
    size_t stride_doubles;
    uint32_t n_rows_local, n_col_blocks;
    uint32_t s_begin, s_end;
    volatile uint64_t next_job;
    uint64_t total_jobs;
    uint64_t rays, paths;
    pthread_mutex_t lock;
} render_job_ctx;

typedef struct worker_arg { render_job_ctx* ctx; int index; } worker_arg;

static int cam_dof(const wrt_camera* cam, const wrt_params* p) {
    return cam->is_depth_of_field && !(p->flags & WRT_FLAG_DISABLE_DOF);
}

static void* render_worker(void* argp) {
    worker_arg* wa = argp;
    render_job_ctx* c = wa->ctx;
    const wrt_params* p = c->params;
    wro_rng rng;
    memset(&rng, 0, sizeof rng);
    /* rng.zig:9-14: one generator per thread (getrandom there; seed-derived here so runs are repeatable) */
    wro_rng_seed_reference(&rng, p->seed * 0x9E3779B97F4A7C15ull + (uint64_t)wa->index + 1);
    const int dof = cam_dof(c->cam, p);
    uint64_t rays = 0, paths = 0;
    const double scale = 1.0 / (double)p->samples_per_pixel; /* render.zig:123 */
    for (;;) {
        uint64_t job = __atomic_fetch_add(&c->next_job, 1, __ATOMIC_RELAXED);
        if (job >= c->total_jobs) break;
        uint32_t local_row = (uint32_t)(job / c->n_col_blocks);
        uint32_t col0 = (uint32_t)(job % c->n_col_blocks) * 32u; /* pixel_block_size, render.zig:55 */
        uint32_t col1 = col0 + 32u < p->width ? col0 + 32u : p->width;
        uint32_t row = p->row_shard_index + local_row * p->row_shard_count;

        uint32_t sobol_seed = 0;
        if (c->rng_mode == WRO_RNG_REFERENCE) sobol_seed = (uint32_t)wro_xoshiro_next(&rng); /* render.zig:112 */
        wro_sobol_sampler sampler;
        wro_sobol_init(&sampler, p->samples_per_pixel, p->width, p->height, 1, sobol_seed); /* render.zig:115-121 */

        for (uint32_t col = col0; col < col1; ++col) {
            v3 color = v3_splat(0);
            for (uint32_t s = c->s_begin; s < c->s_end; ++s) {
                if (c->rng_mode == WRO_RNG_COUNTER) wro_rng_start_counter(&rng, p->seed, row * p->width + col, s);
                ray r = sample_ray(c->cam, dof, &rng, &sampler, col, row, s);
                trace_ctx tc = {c->scene, arr3(p->background_color), &rng, 0, p->max_ray_bounce_depth};
                v3 cs = ray_color(&tc, &r, p->max_ray_bounce_depth);
                color = v3_add(color, v3_scale(cs, scale));
                rays += tc.rays;
                paths++;
            }
            double* px = c->fb + ((size_t)local_row * p->width + col) * c->stride_doubles;
            px[0] += color.x; px[1] += color.y; px[2] += color.z; /* render.zig:139 */
        }
    }
    pthread_mutex_lock(&c->lock);
    c->rays += rays;
    c->paths += paths;
    pthread_mutex_unlock(&c->lock);
    return NULL;
}

static double now_seconds(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static uint32_t shard_rows(const wrt_params* p) {
    uint32_t cnt = p->row_shard_count ? p->row_shard_count : 1;
    if (p->row_shard_index >= p->height) return 0;
    return (p->height - p->row_shard_index + cnt - 1) / cnt;
}

int wro_render(const wro_scene* s, const wrt_camera* cam, const wrt_params* params, int rng_mode,
               uint32_t n_threads, void* framebuffer, size_t pixel_stride_bytes, wro_render_stats* stats) {
    if (!s || !cam || !params || !framebuffer || pixel_stride_bytes < 24 || pixel_stride_bytes % 8) return -1;
    wrt_params p = *params;
    if (p.row_shard_count == 0) p.row_shard_count = 1;
    if (p.sample_begin == 0 && p.sample_end == 0) p.sample_end = p.samples_per_pixel;
    if (n_threads == 0) n_threads = 1;
    wro_norm_table_init();

    render_job_ctx c;
    memset(&c, 0, sizeof c);
    c.scene = s; c.cam = cam; c.params = &p; c.rng_mode = rng_mode;
    c.fb = framebuffer;
    c.stride_doubles = pixel_stride_bytes / 8;
    c.n_rows_local = shard_rows(&p);
    c.n_col_blocks = (p.width + 31u) / 32u;
    c.s_begin = p.sample_begin; c.s_end = p.sample_end;
    c.total_jobs = (uint64_t)c.n_rows_local * c.n_col_blocks;
    pthread_mutex_init(&c.lock, NULL);

    /* framebuffer.clear(clear_color), render.zig:33 / camera.zig:31-33 */
    if (!(p.flags & WRT_FLAG_NO_CLEAR)) {
        for (size_t i = 0; i < (size_t)c.n_rows_local * p.width; ++i) {
            double* px = c.fb + i * c.stride_doubles;
            px[0] = p.clear_color[0]; px[1] = p.clear_color[1]; px[2] = p.clear_color[2];
            for (size_t k = 3; k < c.stride_doubles; ++k) px[k] = 0.0;
        }
    }

    double t0 = now_seconds();
    pthread_t* th = malloc(n_threads * sizeof *th);
    worker_arg* wa = malloc(n_threads * sizeof *wa);
    for (uint32_t i = 0; i < n_threads; ++i) {
        wa[i].ctx = &c; wa[i].index = (int)i;
        pthread_create(&th[i], NULL, render_worker, &wa[i]);
    }
    for (uint32_t i = 0; i < n_threads; ++i) pthread_join(th[i], NULL);
    double t1 = now_seconds();
    free(th); free(wa);
    pthread_mutex_destroy(&c.lock);
    if (stats) { stats->paths = c.paths; stats->rays = c.rays; stats->seconds = t1 - t0; stats->threads = n_threads; }
    return 0;
}

/* ---- gate 1: primary hits ---------------------------------------------------------------------------- */
typedef struct ph_ctx {
    const wro_scene* scene; const wrt_camera* cam; const wrt_params* p; uint32_t n_samples;
    uint32_t* ids; double* t; volatile uint32_t next_row;
} ph_ctx;

static void* ph_worker(void* argp) {
    ph_ctx* c = argp;
    wro_sobol_sampler sampler;
    wro_sobol_init(&sampler, c->p->samples_per_pixel, c->p->width, c->p->height, 1, 0);
    for (;;) {
        uint32_t row = __atomic_fetch_add(&c->next_row, 1, __ATOMIC_RELAXED);
        if (row >= c->p->height) break;
        for (uint32_t col = 0; col < c->p->width; ++col) {
            for (uint32_t s = 0; s < c->n_samples; ++s) {
                ray r = sample_ray(c->cam, 0, NULL, &sampler, col, row, s);
                wro_hit rec;
                memset(&rec, 0, sizeof rec);
                ival tr = {1e-4, INFINITY};
                size_t o = ((size_t)row * c->p->width + col) * c->n_samples + s;
                if (wro_entity_hit(c->scene, c->scene->root, &r, tr, &rec)) {
                    if (c->ids) c->ids[o] = rec.prim_id;
                    if (c->t) c->t[o] = rec.t;
                } else {
                    if (c->ids) c->ids[o] = WRT_NONE;
                    if (c->t) c->t[o] = INFINITY;
                }
            }
        }
    }
    return NULL;
}

int wro_primary_hits(const wro_scene* s, const wrt_camera* cam, const wrt_params* params, uint32_t n_samples,
                     uint32_t n_threads, uint32_t* prim_ids, double* t) {
    if (!s || !cam || !params) return -1;
    if (n_threads == 0) n_threads = 1;
    ph_ctx c = {s, cam, params, n_samples, prim_ids, t, 0};
    pthread_t* th = malloc(n_threads * sizeof *th);
    for (uint32_t i = 0; i < n_threads; ++i) pthread_create(&th[i], NULL, ph_worker, &c);
    for (uint32_t i = 0; i < n_threads; ++i) pthread_join(th[i], NULL);
    free(th);
    return 0;
}

int wro_trace_rays(const wro_scene* s, const double* origins, const double* directions, uint64_t n, double tmin,
                   uint32_t* prim_ids, double* t, double* point, double* normal, double* uv,
                   uint32_t* front_face) {
    if (!s || !origins || !directions) return -1;
    for (uint64_t i = 0; i < n; ++i) {
        ray r;
        r.origin = arr3(origins + 3 * i);
        r.direction = arr3(directions + 3 * i);
        r.time = 0.0;
        wro_hit rec;
        memset(&rec, 0, sizeof rec);
        ival tr = {tmin, INFINITY};
        int hit = wro_entity_hit(s, s->root, &r, tr, &rec);
        if (prim_ids) prim_ids[i] = hit ? rec.prim_id : WRT_NONE;
        if (t) t[i] = hit ? rec.t : INFINITY;
        if (point) { point[3 * i] = hit ? rec.point.x : 0; point[3 * i + 1] = hit ? rec.point.y : 0; point[3 * i + 2] = hit ? rec.point.z : 0; }
        if (normal) { normal[3 * i] = hit ? rec.normal.x : 0; normal[3 * i + 1] = hit ? rec.normal.y : 0; normal[3 * i + 2] = hit ? rec.normal.z : 0; }
        if (uv) { uv[2 * i] = hit ? rec.uv[0] : 0; uv[2 * i + 1] = hit ? rec.uv[1] : 0; }
        if (front_face) front_face[i] = hit ? (uint32_t)rec.front_face : 0;
    }
    return 0;
}

int wro_light_pdf_values(const wro_scene* s, const double* origins, const double* directions, uint64_t n,
                         double* out) {
    if (!s || !s->lights) return -1;
    for (uint64_t i = 0; i < n; ++i)
        out[i] = wro_entity_pdf_value(s, s->lights, arr3(origins + 3 * i), arr3(directions + 3 * i));
    return 0;
}

/* ---- Sobol entry points ------------------------------------------------------------------------------ */
int wro_sobol_pixel_samples(uint32_t width, uint32_t height, const uint32_t* cols, const uint32_t* rows,
                            const uint32_t* sample_idx, uint64_t n, uint64_t* sobol_index, double* offsets_xy) {
    wro_sobol_sampler sp;
    wro_sobol_init(&sp, 1, width, height, 1, 0);
    for (uint64_t i = 0; i < n; ++i) {
        wro_sobol_start_pixel_sample(&sp, cols[i], rows[i], sample_idx[i]);
        if (sobol_index) sobol_index[i] = sp.sobol_idx;
        if (offsets_xy) wro_sobol_get_pixel_2d(&sp, offsets_xy + 2 * i);
    }
    return 0;
}

int wro_sobol_dimension_samples(const uint64_t* sobol_index, const uint32_t* dimension, uint64_t n,
                                uint32_t owen_fast, uint32_t seed, float* out) {
    wro_sobol_sampler sp;
    wro_sobol_init(&sp, 1, 1, 1, (int)owen_fast, seed);
    for (uint64_t i = 0; i < n; ++i) {
        sp.sobol_idx = sobol_index[i];
        out[i] = wro_sobol_sample_dimension(&sp, dimension[i]);
    }
    return 0;
}

int wro_sobol_get1d_sequence(uint32_t width, uint32_t height, uint32_t col, uint32_t row, uint32_t sample_idx,
                             uint32_t owen_fast, uint32_t seed, uint32_t n, double* out) {
    wro_sobol_sampler sp;
    wro_sobol_init(&sp, 1, width, height, (int)owen_fast, seed);
    wro_sobol_start_pixel_sample(&sp, col, row, sample_idx);
    for (uint32_t i = 0; i < n; ++i) out[i] = wro_sobol_get_1d(&sp);
    return 0;
}

/* ---- writer.zig:68-123 ------------------------------------------------------------------------------- */
void wro_encode_color(const double rgb[3], uint8_t out[3]) {
    const double rgb_max = 256.0;
    for (int k = 0; k < 3; ++k) {
        double c = rgb[k];
        if (c != c) c = 0; /* clampNaN */
        c = sqrt(c);       /* gammaCorrection */
        out[k] = (uint8_t)(rgb_max * wro_clamp(c, 0.0, 0.999));
    }
}
void wro_encode_image(const void* framebuffer, size_t pixel_stride_bytes, uint64_t n_pixels, uint8_t* rgb_out) {
    const char* base = framebuffer;
    for (uint64_t i = 0; i < n_pixels; ++i) wro_encode_color((const double*)(base + i * pixel_stride_bytes), rgb_out + 3 * i);
}
uint32_t wro_size_of_digit(uint8_t digit) {
    uint32_t result = 0x1;
    result <<= (digit > 9);
    result |= (digit > 99);
    return result;
}
uint32_t wro_size_of_line(const uint8_t pixel[3]) {
    return 3u + wro_size_of_digit(pixel[0]) + wro_size_of_digit(pixel[1]) + wro_size_of_digit(pixel[2]);
}

uint64_t wro_counter_rng_bits(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t draw) {
    return wro_counter_bits(seed, pixel, sample, draw);
}

void wro_math_cross(const double u[3], const double v[3], double out[3]) {
    v3 c = v3_cross(arr3(u), arr3(v));
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
}
double wro_math_dot(const double u[3], const double v[3]) { return v3_dot(arr3(u), arr3(v)); }
double wro_math_length(const double u[3]) { return v3_length(arr3(u)); }
void wro_math_normalize(const double u[3], double out[3]) {
    v3 c = v3_normalize(arr3(u));
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
}

/* ---- known-answer hooks for the restated Zig std pieces (tests/test_oracle_kats.py) ---------------------------------
 * The generators below are what rng.zig:6,17 instantiate (std.Random.DefaultPrng = Xoshiro256++ seeded through SplitMix64);
 * the tests pin them against the PUBLIC vectors of the algorithms' authors (Vigna's splitmix64.c / xoshiro256plusplus.c). */
void wro_kat_splitmix64(uint64_t seed, uint32_t n, uint64_t* out) {
    uint64_t s = seed;
    for (uint32_t i = 0; i < n; ++i) out[i] = wro_splitmix64(&s);
}
void wro_kat_xoshiro256pp(const uint64_t state[4], uint32_t n, uint64_t* out) {
    wro_rng r;
    r.mode = WRO_RNG_REFERENCE;
    for (int i = 0; i < 4; ++i) r.s[i] = state[i];
    for (uint32_t i = 0; i < n; ++i) out[i] = wro_xoshiro_next(&r);
}
/* DefaultPrng.init(seed) followed by n draws of Random.float(f64) / next() */
void wro_kat_default_prng(uint64_t seed, uint32_t n, uint64_t* out_u64, double* out_float) {
    wro_rng r;
    wro_rng_seed_reference(&r, seed);
    for (uint32_t i = 0; i < n; ++i) {
        if (out_u64) out_u64[i] = wro_xoshiro_next(&r);
        else out_float[i] = wro_rng_float(&r);
    }
}
