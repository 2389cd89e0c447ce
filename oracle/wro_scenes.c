/*
 * wro_scenes.c — ORACLE (test infrastructure, not product code).
 * The reference's scene catalogue, src/scene.zig:68-517, restated constant for constant, plus the two harness
 * scenes the benchmark configs need ("earth" = C4, "synthetic" = C5).
 * The reference seeds scene randomness from getrandom (rng.zig:16-26); here an explicit seed feeds a restated
 * Xoshiro256++ so the same scene can be rebuilt by the product's host code.
 */
#include <stdlib.h>
#include <string.h>

#include "wro.h"
#include "wro_scene.h"

static const wro_image_in* find_image(const wro_image_in* images, uint32_t n, const char* name) {
    for (uint32_t i = 0; i < n; ++i)
        if (images[i].name && strcmp(images[i].name, name) == 0) return &images[i];
    return NULL;
}

/* ImageTexture.initTextureFromPath (texture.zig:38-42).  When the caller has no bytes for `name` the texture
 * gets the reference's "no image" state (magenta, texture.zig:53-55). */
static wro_texture* image_texture(wro_scene* s, const wro_image_in* images, uint32_t n, const char* name) {
    const wro_image_in* in = find_image(images, n, name);
    wro_image* im = in ? wro_add_image(s, in->width, in->height, in->num_components, in->data)
                       : wro_add_image(s, 0, 0, 3, NULL);
    return wro_tex_image(s, im);
}

static void set_camera(wro_scene* s, v3 from, v3 at, v3 up, double fov, double focus, double defocus) {
    s->camera.look_from = from; s->camera.look_at = at; s->camera.view_up = up;
    s->camera.fov_vertical = fov; s->camera.lens_focus_dist = focus; s->camera.defocus_angle_degrees = defocus;
}

/* scene.zig:68-174 */
static void load_balls(wro_scene* s, uint64_t seed) {
    wro_rng rng;
    wro_rng_seed_reference(&rng, seed);
    wro_texture* brown = wro_tex_solid(s, v3_make(0.4, 0.2, 0.1));
    wro_texture* even = wro_tex_solid(s, v3_make(0.2, 0.3, 0.1));
    wro_texture* odd = wro_tex_solid(s, v3_make(0.9, 0.9, 0.9));
    wro_texture* ground = wro_tex_checker(s, 0.32, even, odd);

    wro_entity* scene = wro_collection(s);
    wro_collection_add(scene, wro_sphere(s, v3_make(0, -1000, 0), 1000, wro_mat_lambertian(s, ground)));

    for (double a = -11.0; a < 11.0; a += 1.0) {
        for (double b = -11.0; b < 11.0; b += 1.0) {
            double choose_mat = wro_rng_float(&rng);
            double cx = a + 0.9 * wro_rng_float(&rng);
            double cz = b + 0.9 * wro_rng_float(&rng);
            v3 center = v3_make(cx, 0.2, cz);
            if (v3_length(v3_sub(center, v3_make(4, 0.2, 0))) > 0.9) {
                if (choose_mat < 0.8) {
                    v3 albedo = wro_sample_vec3(&rng);
                    wro_collection_add(scene, wro_sphere(s, center, 0.2, wro_mat_lambertian(s, wro_tex_solid(s, albedo))));
                } else if (choose_mat < 0.95) {
                    v3 albedo = wro_sample_vec3_interval(&rng, 0.5, 1.0);
                    double fuzz = wro_rng_float(&rng) * 0.8;
                    wro_collection_add(scene, wro_sphere(s, center, 0.2, wro_mat_metal(s, albedo, fuzz)));
                } else {
                    wro_collection_add(scene, wro_sphere(s, center, 0.2, wro_mat_dielectric(s, 1.5)));
                }
            }
        }
    }
    wro_collection_add(scene, wro_sphere(s, v3_make(0, 1, 0), 1.0, wro_mat_dielectric(s, 1.5)));
    wro_collection_add(scene, wro_sphere(s, v3_make(-4, 1, 0), 1, wro_mat_lambertian(s, brown)));
    wro_collection_add(scene, wro_sphere(s, v3_make(4, 1, 0), 1, wro_mat_metal(s, v3_make(0.7, 0.6, 0.5), 0.0)));
    wro_collection_create_bvh(s, scene);

    s->root = scene;
    s->lights = NULL;
    set_camera(s, v3_make(13, 2, 3), v3_make(0, 0, 0), v3_make(0, 1, 0), 20.0, 10.0, 0.6);
    s->background_color = v3_make(0.5, 0.7, 1.0);
}

/* scene.zig:176-230 */
static void load_shrek_quads(wro_scene* s, const wro_image_in* images, uint32_t n_images) {
    wro_texture* tex = image_texture(s, images, n_images, "wap.jpg");
    wro_material* left = wro_mat_lambertian(s, tex);
    wro_material* back = wro_mat_lambertian(s, tex);
    wro_material* right = wro_mat_lambertian(s, tex);
    wro_material* top = wro_mat_lambertian(s, tex);
    wro_material* bottom = wro_mat_lambertian(s, tex);
    wro_entity* scene = wro_collection(s);
    wro_collection_add(scene, wro_quad(s, v3_make(-3, -2, 5), v3_make(0, 0, -4), v3_make(0, 4, 0), left));
    wro_collection_add(scene, wro_quad(s, v3_make(-2, -2, 0), v3_make(4, 0, 0), v3_make(0, 4, 0), right));
    wro_collection_add(scene, wro_quad(s, v3_make(3, -2, 1), v3_make(0, 0, 4), v3_make(0, 4, 0), back));
    wro_collection_add(scene, wro_quad(s, v3_make(-2, 3, 1), v3_make(4, 0, 0), v3_make(0, 0, 4), top));
    wro_collection_add(scene, wro_quad(s, v3_make(-2, -3, 5), v3_make(4, 0, 0), v3_make(0, 0, -4), bottom));
    s->root = scene; /* no BVH: linear scan (entity.zig:351-367) */
    s->lights = NULL;
    set_camera(s, v3_make(0, 0, 9), v3_make(0, 0, 0), v3_make(0, 1, 0), 80.0, 10.0, 0.0);
    s->background_color = v3_make(0.5, 0.7, 1.0);
}

/* scene.zig:232-310 */
static void load_emissive(wro_scene* s) {
    wro_texture* even = wro_tex_solid(s, v3_make(0.2, 0.3, 0.1));
    wro_texture* odd = wro_tex_solid(s, v3_make(0.9, 0.9, 0.9));
    wro_texture* ground = wro_tex_checker(s, 0.32, even, odd);
    wro_texture* light_blue = wro_tex_solid(s, v3_make(1, 2, 4));
    wro_texture* light_green = wro_tex_solid(s, v3_make(2.3, 4, 2.3));
    wro_material* m_glass = wro_mat_dielectric(s, 1.5);
    wro_material* m_ground = wro_mat_lambertian(s, ground);
    wro_material* m_blue = wro_mat_diffuse_light(s, light_blue);
    wro_material* m_green = wro_mat_diffuse_light(s, light_green);

    wro_entity* scene = wro_collection(s);
    wro_entity* ground_sphere = wro_sphere(s, v3_make(0, -1000, 0), 1000, m_ground);
    wro_collection_add(scene, ground_sphere);
    wro_entity* glass_sphere = wro_sphere(s, v3_make(0, 2, 0), 1.5, m_glass);
    wro_collection_add(scene, glass_sphere);
    wro_entity* light_quad = wro_quad(s, v3_make(3, 1, -2), v3_make(2, 0, 0), v3_make(0, 2, 0), m_blue);
    wro_collection_add(scene, light_quad);
    wro_entity* light_sphere = wro_sphere(s, v3_make(0, 7, 0), 1, m_green);
    wro_collection_add(scene, light_sphere);
    wro_collection_create_bvh(s, scene);

    wro_entity* lights = wro_collection(s);
    wro_collection_add(lights, light_quad);
    wro_collection_add(lights, light_sphere);
    wro_collection_add(lights, glass_sphere);

    s->root = scene;
    s->lights = lights;
    set_camera(s, v3_make(26, 3, 6), v3_make(0, 2, 0), v3_make(0, 1, 0), 20.0, 10.0, 0.0);
    s->background_color = v3_make(0, 0, 0);
}

/* scene.zig:312-408 */
static void load_cornell_box(wro_scene* s) {
    wro_texture* red = wro_tex_solid(s, v3_make(0.65, 0.05, 0.05));
    wro_texture* white = wro_tex_solid(s, v3_make(0.73, 0.73, 0.73));
    wro_texture* green = wro_tex_solid(s, v3_make(0.12, 0.45, 0.15));
    wro_texture* light_t = wro_tex_solid(s, v3_make(15, 15, 15));
    wro_material* m_red = wro_mat_lambertian(s, red);
    wro_material* m_white = wro_mat_lambertian(s, white);
    wro_material* m_green = wro_mat_lambertian(s, green);
    wro_material* m_light = wro_mat_diffuse_light(s, light_t);
    wro_material* m_glass = wro_mat_dielectric(s, 1.5);
    wro_material* m_metal = wro_mat_metal(s, v3_make(0.8, 0.85, 0.88), 0);

    wro_entity* scene = wro_collection(s);
    wro_collection_add(scene, wro_quad(s, v3_make(555, 0, 0), v3_make(0, 555, 0), v3_make(0, 0, 555), m_green));
    wro_collection_add(scene, wro_quad(s, v3_make(0, 0, 0), v3_make(0, 555, 0), v3_make(0, 0, 555), m_red));
    wro_collection_add(scene, wro_quad(s, v3_make(0, 0, 0), v3_make(555, 0, 0), v3_make(0, 0, 555), m_white));
    wro_collection_add(scene, wro_quad(s, v3_make(555, 555, 555), v3_make(-555, 0, 0), v3_make(0, 0, -555), m_white));
    wro_collection_add(scene, wro_quad(s, v3_make(0, 0, 555), v3_make(555, 0, 0), v3_make(0, 555, 0), m_white));

    wro_entity* glass_sphere = wro_sphere(s, v3_make(190, 90, 190), 90, m_glass);
    wro_collection_add(scene, glass_sphere);

    wro_entity* box2 = wro_translate(s, v3_make(265, 0, 295),
                                     wro_rotate_y(s, 15.0, wro_box(s, v3_make(0, 0, 0), v3_make(165, 330, 165), m_metal)));
    wro_collection_add(scene, box2);

    wro_entity* light = wro_quad(s, v3_make(343, 554, 332), v3_make(-150, 0, 0), v3_make(0, 0, -125), m_light);
    wro_collection_add(scene, light);
    wro_collection_create_bvh(s, scene);

    wro_entity* lights = wro_collection(s);
    wro_collection_add(lights, glass_sphere);
    wro_collection_add(lights, light);

    s->root = scene;
    s->lights = lights;
    set_camera(s, v3_make(278, 278, -800), v3_make(278, 278, 0), v3_make(0, 1, 0), 40.0, 10.0, 0.0);
    s->background_color = v3_make(0, 0, 0);
}

/* scene.zig:410-517 */
static void load_rtw_final(wro_scene* s, uint64_t seed, const wro_image_in* images, uint32_t n_images) {
    wro_rng rng;
    wro_rng_seed_reference(&rng, seed);
    wro_entity* scene = wro_collection(s);
    wro_entity* lights = wro_collection(s);

    wro_material* m_ground = wro_mat_lambertian(s, wro_tex_solid(s, v3_make(0.4, 0.83, 0.53)));
    wro_entity* ground_boxes = wro_collection(s);
    wro_collection_add(scene, ground_boxes); /* added while still empty: scene.aabb ∪ default (scene.zig:428) */
    const int num_boxes_per_side = 20;
    for (int i = 0; i < num_boxes_per_side; ++i) {
        double fi = (double)i;
        for (int j = 0; j < num_boxes_per_side; ++j) {
            double fj = (double)j;
            const double w = 100.0;
            double x0 = -1000.0 + fi * w;
            double y0 = 0.0;
            double z0 = -1000.0 + fj * w;
            double x1 = x0 + w;
            double y1 = wro_rng_float(&rng) * 100.0 + 1.0;
            double z1 = z0 + w;
            wro_collection_add(ground_boxes, wro_box(s, v3_make(x0, y0, z0), v3_make(x1, y1, z1), m_ground));
        }
    }
    wro_collection_create_bvh(s, ground_boxes);

    wro_material* m_light = wro_mat_diffuse_light(s, wro_tex_solid(s, v3_make(7, 7, 7)));
    wro_entity* light = wro_quad(s, v3_make(123, 554, 147), v3_make(300, 0, 0), v3_make(0, 0, 265), m_light);
    wro_collection_add(scene, light);
    wro_collection_add(lights, light);

    wro_collection_add(scene, wro_sphere(s, v3_make(260, 150, 45), 50.0, wro_mat_dielectric(s, 1.5)));
    wro_collection_add(scene, wro_sphere(s, v3_make(0, 150, 145), 50, wro_mat_metal(s, v3_make(0.8, 0.8, 0.9), 1.0)));
    wro_collection_add(scene, wro_sphere(s, v3_make(360, 150, 145), 70, wro_mat_dielectric(s, 1.5)));

    wro_texture* t_shrek = image_texture(s, images, n_images, "wap.jpg");
    wro_collection_add(scene, wro_sphere(s, v3_make(400, 200, 400), 100, wro_mat_lambertian(s, t_shrek)));
    wro_texture* t_me = image_texture(s, images, n_images, "me.jpg");
    wro_collection_add(scene, wro_sphere(s, v3_make(220, 280, 300), 80, wro_mat_lambertian(s, t_me)));

    wro_entity* box_of_balls = wro_collection(s);
    wro_material* m_white = wro_mat_lambertian(s, wro_tex_solid(s, v3_make(0.73, 0.73, 0.73)));
    for (int i = 0; i < 1000; ++i) {
        v3 center = v3_mul(wro_sample_vec3(&rng), v3_splat(165.0));
        wro_collection_add(box_of_balls, wro_sphere(s, center, 10, m_white));
    }
    wro_collection_create_bvh(s, box_of_balls);
    wro_collection_add(scene, wro_translate(s, v3_make(-100, 270, 395), wro_rotate_y(s, 15.0, box_of_balls)));
    wro_collection_create_bvh(s, scene);

    s->root = scene;
    s->lights = lights;
    set_camera(s, v3_make(478, 278, -600), v3_make(278, 278, 0), v3_make(0, 1, 0), 40.0, 10.0, 0.0);
    s->background_color = v3_make(0, 0, 0);
}

/* Harness scene for config C4 (not in the reference; assets/earth.png exists there but no scene uses it):
 * the `emissive` layout with the glass sphere replaced by an earth-textured Lambertian sphere. */
static void load_earth(wro_scene* s, const wro_image_in* images, uint32_t n_images) {
    wro_texture* even = wro_tex_solid(s, v3_make(0.2, 0.3, 0.1));
    wro_texture* odd = wro_tex_solid(s, v3_make(0.9, 0.9, 0.9));
    wro_texture* ground = wro_tex_checker(s, 0.32, even, odd);
    wro_texture* t_earth = image_texture(s, images, n_images, "earth.png");
    wro_texture* light_a = wro_tex_solid(s, v3_make(4, 4, 4));
    wro_texture* light_b = wro_tex_solid(s, v3_make(3, 2.7, 2.3));
    wro_material* m_ground = wro_mat_lambertian(s, ground);
    wro_material* m_earth = wro_mat_lambertian(s, t_earth);
    wro_material* m_la = wro_mat_diffuse_light(s, light_a);
    wro_material* m_lb = wro_mat_diffuse_light(s, light_b);

    wro_entity* scene = wro_collection(s);
    wro_collection_add(scene, wro_sphere(s, v3_make(0, -1000, 0), 1000, m_ground));
    wro_collection_add(scene, wro_sphere(s, v3_make(0, 2, 0), 2.0, m_earth));
    wro_entity* light_quad = wro_quad(s, v3_make(3, 1, -2), v3_make(2, 0, 0), v3_make(0, 2, 0), m_la);
    wro_collection_add(scene, light_quad);
    wro_entity* light_sphere = wro_sphere(s, v3_make(0, 7, 0), 1, m_lb);
    wro_collection_add(scene, light_sphere);
    wro_collection_create_bvh(s, scene);

    wro_entity* lights = wro_collection(s);
    wro_collection_add(lights, light_quad);
    wro_collection_add(lights, light_sphere);

    s->root = scene;
    s->lights = lights;
    set_camera(s, v3_make(26, 3, 6), v3_make(0, 2, 0), v3_make(0, 1, 0), 20.0, 10.0, 0.0);
    s->background_color = v3_make(0, 0, 0);
}

/* Harness scene for config C5: n primitives alternating sphere / quad, centres uniform in
 * [-1000,1000]^3, sphere radius U[1,5], quad edges an axis-aligned pair with lengths U[2,10]; materials from a
 * 4096-entry palette drawn 70 % lambertian / 20 % metal (fuzz U[0,0.5]) / 10 % glass; 64 emissive quads
 * (radiance 15, edges U[20,60]) registered as lights; camera (0,0,-3000) -> origin, vfov 40. */
static void load_synthetic(wro_scene* s, uint64_t seed, uint32_t n_prims) {
    wro_rng rng;
    wro_rng_seed_reference(&rng, seed);
    enum { PALETTE = 4096, N_LIGHTS = 64 };
    wro_material** palette = malloc(PALETTE * sizeof *palette);
    for (int i = 0; i < PALETTE; ++i) {
        double choose = wro_rng_float(&rng);
        if (choose < 0.7) {
            v3 albedo = wro_sample_vec3(&rng);
            palette[i] = wro_mat_lambertian(s, wro_tex_solid(s, albedo));
        } else if (choose < 0.9) {
            v3 albedo = wro_sample_vec3_interval(&rng, 0.5, 1.0);
            double fuzz = wro_rng_float(&rng) * 0.5;
            palette[i] = wro_mat_metal(s, albedo, fuzz);
        } else {
            palette[i] = wro_mat_dielectric(s, 1.5);
        }
    }
    wro_material* m_light = wro_mat_diffuse_light(s, wro_tex_solid(s, v3_make(15, 15, 15)));

    wro_entity* scene = wro_collection(s);
    wro_entity* lights = wro_collection(s);
    if (n_prims < N_LIGHTS * 2) n_prims = N_LIGHTS * 2;
    uint32_t light_every = n_prims / N_LIGHTS;
    uint32_t n_lights = 0;
    for (uint32_t i = 0; i < n_prims; ++i) {
        v3 c = wro_sample_vec3_interval(&rng, -1000.0, 1000.0);
        int is_light = (i % light_every == 1) && n_lights < N_LIGHTS; /* odd index => a quad slot */
        if ((i & 1u) == 0) {
            double radius = wro_rng_float(&rng) * 4.0 + 1.0;
            uint32_t pm = wro_rng_pick(&rng, PALETTE);
            wro_collection_add(scene, wro_sphere(s, c, radius, palette[pm]));
        } else {
            double lo = is_light ? 20.0 : 2.0, span = is_light ? 40.0 : 8.0;
            double l1 = wro_rng_float(&rng) * span + lo;
            double l2 = wro_rng_float(&rng) * span + lo;
            uint32_t axis = wro_rng_pick(&rng, 3); /* the axis the quad is perpendicular to */
            uint32_t pm = wro_rng_pick(&rng, PALETTE);
            v3 a1 = v3_splat(0), a2 = v3_splat(0);
            if (axis == 0) { a1.y = l1; a2.z = l2; }
            else if (axis == 1) { a1.z = l1; a2.x = l2; }
            else { a1.x = l1; a2.y = l2; }
            wro_entity* q = wro_quad(s, c, a1, a2, is_light ? m_light : palette[pm]);
            wro_collection_add(scene, q);
            if (is_light) { wro_collection_add(lights, q); ++n_lights; }
        }
    }
    free(palette);
    wro_collection_create_bvh(s, scene);
    s->root = scene;
    s->lights = lights;
    set_camera(s, v3_make(0, 0, -3000), v3_make(0, 0, 0), v3_make(0, 1, 0), 40.0, 10.0, 0.0);
    s->background_color = v3_make(0, 0, 0);
}

wro_scene* wro_scene_build(const char* name, uint64_t seed, uint32_t n_prims, const wro_image_in* images,
                           uint32_t n_images) {
    wro_scene* s = wro_scene_new();
    if (strcmp(name, "balls") == 0) load_balls(s, seed);
    else if (strcmp(name, "shrek_quads") == 0) load_shrek_quads(s, images, n_images);
    else if (strcmp(name, "emissive") == 0) load_emissive(s);
    else if (strcmp(name, "cornell_box") == 0) load_cornell_box(s);
    else if (strcmp(name, "rtw_final") == 0) load_rtw_final(s, seed, images, n_images);
    else if (strcmp(name, "earth") == 0) load_earth(s, images, n_images);
    else if (strcmp(name, "synthetic") == 0) load_synthetic(s, seed, n_prims);
    else { wro_scene_free(s); return NULL; }
    wro_scene_finalize(s);
    return s;
}

void wro_scene_destroy(wro_scene* s) { wro_scene_free(s); }
void wro_scene_set_no_cull(wro_scene* s, int no_cull) { s->no_cull = no_cull; }
void wro_scene_background(const wro_scene* s, double out[3]) {
    out[0] = s->background_color.x; out[1] = s->background_color.y; out[2] = s->background_color.z;
}
uint32_t wro_scene_n_prims(const wro_scene* s) { return s->n_prims; }

void wro_camera_view(const wro_camera_desc* c, uint32_t image_width, uint32_t image_height, wrt_camera* out);
void wro_scene_camera(const wro_scene* s, uint32_t width, uint32_t height, wrt_camera* out) {
    wro_camera_view(&s->camera, width, height, out);
}
void wro_scene_camera_desc(const wro_scene* s, double out[12]) {
    const wro_camera_desc* c = &s->camera;
    out[0] = c->look_from.x; out[1] = c->look_from.y; out[2] = c->look_from.z;
    out[3] = c->look_at.x; out[4] = c->look_at.y; out[5] = c->look_at.z;
    out[6] = c->view_up.x; out[7] = c->view_up.y; out[8] = c->view_up.z;
    out[9] = c->fov_vertical; out[10] = c->lens_focus_dist; out[11] = c->defocus_angle_degrees;
}
