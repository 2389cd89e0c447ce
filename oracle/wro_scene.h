/*
 * wro_scene.h — ORACLE (test infrastructure, not product code).
 *
 * Pointer-tree scene model restating the reference's tagged unions:
 *   IEntity   src/entity.zig:17-24      IMaterial src/material.zig:25-32
 *   ITexture  src/texture.zig:11-16     Image     src/image.zig:7-48
 *   HitRecord src/hitrecord.zig:6-26    Camera / Viewport src/camera.zig:48-157
 */
#ifndef WRO_SCENE_H
#define WRO_SCENE_H

#include <stddef.h>
#include <stdint.h>

#include "wro_math.h"
#include "wro_rng.h"

enum { WRO_ENT_SPHERE = 0, WRO_ENT_QUAD, WRO_ENT_COLLECTION, WRO_ENT_BVH_NODE, WRO_ENT_TRANSLATE, WRO_ENT_ROTATE_Y };
enum { WRO_MAT_LAMBERTIAN = 0, WRO_MAT_ISOTROPIC, WRO_MAT_METAL, WRO_MAT_DIELECTRIC, WRO_MAT_DIFFUSE_EMISSIVE };
enum { WRO_TEX_SOLID = 0, WRO_TEX_CHECKER, WRO_TEX_IMAGE };

typedef struct wro_image {
    uint32_t width, height, num_components, bytes_per_row;
    uint8_t* data; /* owned; NULL + height 0 = the reference's "no image" state (image.zig:10) */
    uint32_t index;
} wro_image;

typedef struct wro_texture {
    int kind;
    v3 color;                       /* solid_color */
    double inv_scale;               /* checkerboard */
    const struct wro_texture* even; /* checkerboard */
    const struct wro_texture* odd;
    const wro_image* image;         /* image */
    uint32_t index;
} wro_texture;

typedef struct wro_material {
    int kind;
    const wro_texture* texture; /* lambertian, isotropic, diffuse_emissive */
    v3 albedo;                  /* metal */
    double param;               /* metal fuzz / dielectric refraction_index */
    uint32_t index;
} wro_material;

typedef struct wro_entity wro_entity;
struct wro_entity {
    int kind;
    aabb box;
    uint32_t prim_id; /* sphere/quad: DFS first-visit order from the scene root; else WRT_NONE */
    uint32_t flat_id; /* scratch for the flattener */
    union {
        struct { v3 center; double radius; const wro_material* material; int is_moving; v3 movement; } sphere;
        struct { v3 start; onb basis; v3 normal; double offset, area; const wro_material* material; } quad;
        struct { wro_entity** items; size_t len, cap; wro_entity* bvh_root; } collection;
        struct { wro_entity* left; wro_entity* right; } bvh;
        struct { v3 offset; wro_entity* child; } translate;
        struct { double sin_theta, cos_theta; wro_entity* child; } rotate_y;
    } u;
};

/* hitrecord.zig:6-14 (+ prim_id, which the reference does not have: SURVEY.md A.8) */
typedef struct wro_hit {
    v3 point, normal;
    const wro_material* material;
    double t;
    double uv[2];
    int front_face;
    uint32_t prim_id;
} wro_hit;

typedef struct wro_camera_desc { /* arguments of Camera.init, camera.zig:61-68 */
    v3 look_from, look_at, view_up;
    double fov_vertical, lens_focus_dist, defocus_angle_degrees;
} wro_camera_desc;

typedef struct wro_scene {
    /* ownership lists */
    wro_entity** pool; size_t n_pool, cap_pool;
    wro_material** materials; size_t n_materials, cap_materials;
    wro_texture** textures; size_t n_textures, cap_textures;
    wro_image** images; size_t n_images, cap_images;
    /* Scene fields, scene.zig:36-45 */
    wro_entity* root;
    wro_entity* lights; /* NULL when the scene has none */
    wro_camera_desc camera;
    v3 background_color;
    uint32_t n_prims;
    int no_cull; /* 0 = reference AABB.hit; 1 = skip every box test (brute force over all leaves; self-validation only) */
} wro_scene;

/* ---- construction (entity.zig initEntity functions) ------------------------------------------------ */
wro_scene* wro_scene_new(void);
void wro_scene_free(wro_scene* s);

wro_image* wro_add_image(wro_scene* s, uint32_t w, uint32_t h, uint32_t comps, const uint8_t* data);
wro_texture* wro_tex_solid(wro_scene* s, v3 color);
wro_texture* wro_tex_checker(wro_scene* s, double inv_scale, const wro_texture* even, const wro_texture* odd);
wro_texture* wro_tex_image(wro_scene* s, const wro_image* img);
wro_material* wro_mat_lambertian(wro_scene* s, const wro_texture* t);
wro_material* wro_mat_isotropic(wro_scene* s, const wro_texture* t);
wro_material* wro_mat_metal(wro_scene* s, v3 albedo, double fuzz);
wro_material* wro_mat_dielectric(wro_scene* s, double refraction_index);
wro_material* wro_mat_diffuse_light(wro_scene* s, const wro_texture* t);

wro_entity* wro_sphere(wro_scene* s, v3 center, double radius, const wro_material* m);
wro_entity* wro_sphere_animated(wro_scene* s, v3 c0, v3 c1, double radius, const wro_material* m);
wro_entity* wro_quad(wro_scene* s, v3 start, v3 axis1, v3 axis2, const wro_material* m);
wro_entity* wro_collection(wro_scene* s);
void wro_collection_add(wro_entity* coll, wro_entity* e);
void wro_collection_create_bvh(wro_scene* s, wro_entity* coll);
wro_entity* wro_box(wro_scene* s, v3 a, v3 b, const wro_material* m);
wro_entity* wro_translate(wro_scene* s, v3 offset, wro_entity* child);
wro_entity* wro_rotate_y(wro_scene* s, double angle_degrees, wro_entity* child);
/* raw constructors used when a tree is rebuilt from flat arrays (boxes and derived fields given) */
wro_entity* wro_entity_raw(wro_scene* s, int kind);
/* assigns prim ids by DFS from the root; call once after the tree is complete */
void wro_scene_finalize(wro_scene* s);

/* ---- hot path ----------------------------------------------------------------------------------- */
int wro_entity_hit(const wro_scene* s, const wro_entity* e, const ray* r, ival trange, wro_hit* rec);
double wro_entity_pdf_value(const wro_scene* s, const wro_entity* e, v3 origin, v3 direction);
v3 wro_entity_sample_direction(const wro_entity* e, wro_rng* rng, v3 origin);
v3 wro_texture_value(const wro_texture* t, const double uv[2], v3 point);

#endif /* WRO_SCENE_H */
