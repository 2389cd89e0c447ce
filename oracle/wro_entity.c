/*
 * wro_entity.c — ORACLE (test infrastructure, not product code).
 * CPU restatement of src/entity.zig (entities, BVH builder, closest hit, light-sampling hooks),
 * src/texture.zig and src/image.zig.  Quirks are kept on purpose.
 */
#include <stdlib.h>
#include <string.h>

#include "wro_scene.h"

#define WRO_NONE 0xFFFFFFFFu

/* ---- ownership ------------------------------------------------------------------------------------ */
#define PUSH(arr, n, cap, item)                                          \
    do {                                                                 \
        if ((n) == (cap)) {                                              \
            (cap) = (cap) ? (cap) * 2 : 64;                              \
            (arr) = realloc((arr), (cap) * sizeof *(arr));               \
        }                                                                \
        (arr)[(n)++] = (item);                                           \
    } while (0)

wro_scene* wro_scene_new(void) {
    wro_scene* s = calloc(1, sizeof *s);
    return s;
}

void wro_scene_free(wro_scene* s) {
    if (!s) return;
    for (size_t i = 0; i < s->n_pool; ++i) {
        if (s->pool[i]->kind == WRO_ENT_COLLECTION) free(s->pool[i]->u.collection.items);
        free(s->pool[i]);
    }
    for (size_t i = 0; i < s->n_materials; ++i) free(s->materials[i]);
    for (size_t i = 0; i < s->n_textures; ++i) free(s->textures[i]);
    for (size_t i = 0; i < s->n_images; ++i) { free(s->images[i]->data); free(s->images[i]); }
    free(s->pool); free(s->materials); free(s->textures); free(s->images);
    free(s);
}

wro_entity* wro_entity_raw(wro_scene* s, int kind) {
    wro_entity* e = calloc(1, sizeof *e);
    e->kind = kind;
    e->prim_id = WRO_NONE;
    e->flat_id = WRO_NONE;
    e->box = aabb_default();
    PUSH(s->pool, s->n_pool, s->cap_pool, e);
    return e;
}

wro_image* wro_add_image(wro_scene* s, uint32_t w, uint32_t h, uint32_t comps, const uint8_t* data) {
    wro_image* im = calloc(1, sizeof *im);
    im->width = w; im->height = h; im->num_components = comps; im->bytes_per_row = w * comps;
    if (data && h) {
        im->data = malloc((size_t)im->bytes_per_row * h);
        memcpy(im->data, data, (size_t)im->bytes_per_row * h);
    } else {
        im->height = 0; /* image.zig:38-47: no image => height 0 */
    }
    im->index = (uint32_t)s->n_images;
    PUSH(s->images, s->n_images, s->cap_images, im);
    return im;
}

static wro_texture* tex_new(wro_scene* s, int kind) {
    wro_texture* t = calloc(1, sizeof *t);
    t->kind = kind;
    t->index = (uint32_t)s->n_textures;
    PUSH(s->textures, s->n_textures, s->cap_textures, t);
    return t;
}
wro_texture* wro_tex_solid(wro_scene* s, v3 color) { wro_texture* t = tex_new(s, WRO_TEX_SOLID); t->color = color; return t; }
wro_texture* wro_tex_checker(wro_scene* s, double inv_scale, const wro_texture* even, const wro_texture* odd) {
    wro_texture* t = tex_new(s, WRO_TEX_CHECKER);
    t->inv_scale = inv_scale; t->even = even; t->odd = odd;
    return t;
}
wro_texture* wro_tex_image(wro_scene* s, const wro_image* img) { wro_texture* t = tex_new(s, WRO_TEX_IMAGE); t->image = img; return t; }

static wro_material* mat_new(wro_scene* s, int kind) {
    wro_material* m = calloc(1, sizeof *m);
    m->kind = kind;
    m->index = (uint32_t)s->n_materials;
    PUSH(s->materials, s->n_materials, s->cap_materials, m);
    return m;
}
wro_material* wro_mat_lambertian(wro_scene* s, const wro_texture* t) { wro_material* m = mat_new(s, WRO_MAT_LAMBERTIAN); m->texture = t; return m; }
wro_material* wro_mat_isotropic(wro_scene* s, const wro_texture* t) { wro_material* m = mat_new(s, WRO_MAT_ISOTROPIC); m->texture = t; return m; }
wro_material* wro_mat_metal(wro_scene* s, v3 albedo, double fuzz) { wro_material* m = mat_new(s, WRO_MAT_METAL); m->albedo = albedo; m->param = fuzz; return m; }
wro_material* wro_mat_dielectric(wro_scene* s, double ir) { wro_material* m = mat_new(s, WRO_MAT_DIELECTRIC); m->param = ir; return m; }
wro_material* wro_mat_diffuse_light(wro_scene* s, const wro_texture* t) { wro_material* m = mat_new(s, WRO_MAT_DIFFUSE_EMISSIVE); m->texture = t; return m; }

/* ---- entity constructors -------------------------------------------------------------------------- */
/* entity.zig:545-561 SphereEntity.initEntity */
wro_entity* wro_sphere(wro_scene* s, v3 center, double radius, const wro_material* m) {
    wro_entity* e = wro_entity_raw(s, WRO_ENT_SPHERE);
    v3 rvec = v3_splat(radius);
    e->u.sphere.center = center;
    e->u.sphere.radius = radius;
    e->u.sphere.material = m;
    e->box = aabb_init(v3_sub(center, rvec), v3_add(center, rvec));
    return e;
}
/* entity.zig:563-583 SphereEntity.initEntityAnimated */
wro_entity* wro_sphere_animated(wro_scene* s, v3 c0, v3 c1, double radius, const wro_material* m) {
    wro_entity* e = wro_entity_raw(s, WRO_ENT_SPHERE);
    v3 rvec = v3_splat(radius);
    e->u.sphere.center = c0;
    e->u.sphere.radius = radius;
    e->u.sphere.material = m;
    e->u.sphere.is_moving = 1;
    e->u.sphere.movement = v3_sub(c1, c0);
    aabb b0 = aabb_init(v3_sub(c0, rvec), v3_add(c0, rvec));
    aabb b1 = aabb_init(v3_sub(c1, rvec), v3_add(c1, rvec));
    e->box = aabb_union(&b0, &b1);
    return e;
}
/* entity.zig:444-475 QuadEntity.initEntity */
wro_entity* wro_quad(wro_scene* s, v3 start, v3 axis1, v3 axis2, const wro_material* m) {
    wro_entity* e = wro_entity_raw(s, WRO_ENT_QUAD);
    v3 normal = v3_cross(axis1, axis2);
    v3 axis3 = v3_div(normal, v3_splat(v3_dot(normal, normal)));
    v3 normal_unit = v3_normalize(normal);
    double offset = v3_dot(normal_unit, start);
    aabb d1 = aabb_init(start, v3_add(v3_add(start, axis1), axis2));
    aabb d2 = aabb_init(v3_add(start, axis1), v3_add(start, axis2));
    e->u.quad.start = start;
    e->u.quad.basis = onb_from_vectors(axis1, axis2, axis3);
    e->u.quad.normal = normal_unit;
    e->u.quad.offset = offset;
    e->u.quad.area = v3_length(normal);
    e->u.quad.material = m;
    e->box = aabb_union(&d1, &d2);
    return e;
}
/* entity.zig:306-320 */
wro_entity* wro_collection(wro_scene* s) { return wro_entity_raw(s, WRO_ENT_COLLECTION); }
/* entity.zig:327-335 add: append + aabb = aabb ∪ child (seed is the default AABB, A.9-3) */
void wro_collection_add(wro_entity* c, wro_entity* e) {
    if (c->u.collection.len == c->u.collection.cap) {
        c->u.collection.cap = c->u.collection.cap ? c->u.collection.cap * 2 : 8;
        c->u.collection.items = realloc(c->u.collection.items, c->u.collection.cap * sizeof(wro_entity*));
    }
    c->u.collection.items[c->u.collection.len++] = e;
    c->box = aabb_union(&c->box, &e->box);
}
/* entity.zig:390-426 createBoxEntity: front, right, back, left, top, bottom */
wro_entity* wro_box(wro_scene* s, v3 a, v3 b, const wro_material* m) {
    wro_entity* sides = wro_collection(s);
    v3 mn = v3_min(a, b), mx = v3_max(a, b);
    v3 diff = v3_sub(mx, mn);
    v3 dx = v3_make(diff.x, 0, 0), dy = v3_make(0, diff.y, 0), dz = v3_make(0, 0, diff.z);
    wro_collection_add(sides, wro_quad(s, v3_make(mn.x, mn.y, mx.z), dx, dy, m));
    wro_collection_add(sides, wro_quad(s, v3_make(mx.x, mn.y, mx.z), v3_neg(dz), dy, m));
    wro_collection_add(sides, wro_quad(s, v3_make(mx.x, mn.y, mn.z), v3_neg(dx), dy, m));
    wro_collection_add(sides, wro_quad(s, v3_make(mn.x, mn.y, mn.z), dz, dy, m));
    wro_collection_add(sides, wro_quad(s, v3_make(mn.x, mx.y, mx.z), dx, v3_neg(dz), m));
    wro_collection_add(sides, wro_quad(s, v3_make(mn.x, mn.y, mn.z), dx, dz, m));
    return sides;
}
/* entity.zig:75-87 Translate.initEntity */
wro_entity* wro_translate(wro_scene* s, v3 offset, wro_entity* child) {
    wro_entity* e = wro_entity_raw(s, WRO_ENT_TRANSLATE);
    e->u.translate.offset = offset;
    e->u.translate.child = child;
    e->box = aabb_offset(&child->box, offset);
    return e;
}
/* entity.zig:120-163 RotateY.initEntity: x.max is used as the upper value of y and z too (A.9-5) */
wro_entity* wro_rotate_y(wro_scene* s, double angle_degrees, wro_entity* child) {
    wro_entity* e = wro_entity_raw(s, WRO_ENT_ROTATE_Y);
    double theta = angle_degrees * (WRO_PI / 180.0); /* std.math.degreesToRadians */
    double sin_theta = sin(theta), cos_theta = cos(theta);
    const aabb* bbox = &child->box;
    v3 mn = v3_splat(INFINITY), mx = v3_splat(-INFINITY);
    for (int i = 0; i < 2; ++i) {
        double fi = (double)i;
        double x = fi * bbox->x.max + (1.0 - fi) * bbox->x.min;
        for (int j = 0; j < 2; ++j) {
            double fj = (double)j;
            double y = fj * bbox->x.max + (1.0 - fj) * bbox->y.min;
            for (int k = 0; k < 2; ++k) {
                double fk = (double)k;
                double z = fk * bbox->x.max + (1.0 - fk) * bbox->z.min;
                double newx = cos_theta * x + sin_theta * z;
                double newz = -sin_theta * x + cos_theta * z;
                v3 tester = v3_make(newx, y, newz);
                mn = v3_min(mn, tester);
                mx = v3_max(mx, tester);
            }
        }
    }
    e->u.rotate_y.sin_theta = sin_theta;
    e->u.rotate_y.cos_theta = cos_theta;
    e->u.rotate_y.child = child;
    e->box = aabb_init(mn, mx);
    return e;
}

/* ---- BVH builder (entity.zig:226-267) ------------------------------------------------------------- */
/* std.sort.pdq is an unstable sort from the Zig standard library (not under /root/reference); for slices of
 * <= 20 items it is an insertion sort, i.e. stable.  The oracle uses a stable merge sort for every size: tie
 * order only matters for exactly coincident hits and the device consumes the tree the host built. */
static void stable_sort_by_axis(wro_entity** items, size_t n, int axis, wro_entity** tmp) {
    if (n < 2) return;
    if (n <= 8) {
        for (size_t i = 1; i < n; ++i) {
            wro_entity* x = items[i];
            double key = aabb_axis(&x->box, axis).min;
            size_t j = i;
            while (j > 0 && key < aabb_axis(&items[j - 1]->box, axis).min) { items[j] = items[j - 1]; --j; }
            items[j] = x;
        }
        return;
    }
    size_t mid = n / 2;
    stable_sort_by_axis(items, mid, axis, tmp);
    stable_sort_by_axis(items + mid, n - mid, axis, tmp);
    size_t i = 0, j = mid, k = 0;
    while (i < mid && j < n) {
        /* boxCmp (entity.zig:212-216): a before b iff a.min < b.min; take right only when strictly smaller */
        if (aabb_axis(&items[j]->box, axis).min < aabb_axis(&items[i]->box, axis).min) tmp[k++] = items[j++];
        else tmp[k++] = items[i++];
    }
    while (i < mid) tmp[k++] = items[i++];
    while (j < n) tmp[k++] = items[j++];
    memcpy(items, tmp, n * sizeof *items);
}

static wro_entity* bvh_init(wro_scene* s, wro_entity** items, size_t start, size_t end, wro_entity** tmp) {
    wro_entity* node = wro_entity_raw(s, WRO_ENT_BVH_NODE);
    size_t span = end - start;
    if (span == 1) {
        node->u.bvh.left = items[start];
        node->u.bvh.right = items[start];
    } else if (span == 2) {
        node->u.bvh.left = items[start];
        node->u.bvh.right = items[start + 1];
    } else {
        aabb bbox = aabb_default(); /* entity.zig:240: seed includes the origin (A.9-3) */
        for (size_t i = start; i < end; ++i) bbox = aabb_union(&bbox, &items[i]->box);
        int axis = aabb_longest_axis(&bbox);
        stable_sort_by_axis(items + start, span, axis, tmp);
        size_t mid = start + span / 2;
        node->u.bvh.left = bvh_init(s, items, start, mid, tmp);
        node->u.bvh.right = bvh_init(s, items, mid, end, tmp);
    }
    node->box = aabb_union(&node->u.bvh.left->box, &node->u.bvh.right->box);
    return node;
}

/* entity.zig:338-340 createBvhTree: sorts the collection's own item list in place */
void wro_collection_create_bvh(wro_scene* s, wro_entity* c) {
    size_t n = c->u.collection.len;
    wro_entity** tmp = malloc((n ? n : 1) * sizeof *tmp);
    c->u.collection.bvh_root = bvh_init(s, c->u.collection.items, 0, n, tmp);
    free(tmp);
}

/* ---- prim ids --------------------------------------------------------------------- */
static void assign_ids(wro_entity* e, uint32_t* next) {
    switch (e->kind) {
        case WRO_ENT_SPHERE:
        case WRO_ENT_QUAD:
            if (e->prim_id == WRO_NONE) e->prim_id = (*next)++;
            break;
        case WRO_ENT_COLLECTION:
            if (e->u.collection.bvh_root) assign_ids(e->u.collection.bvh_root, next);
            else for (size_t i = 0; i < e->u.collection.len; ++i) assign_ids(e->u.collection.items[i], next);
            break;
        case WRO_ENT_BVH_NODE:
            assign_ids(e->u.bvh.left, next);
            assign_ids(e->u.bvh.right, next);
            break;
        case WRO_ENT_TRANSLATE: assign_ids(e->u.translate.child, next); break;
        case WRO_ENT_ROTATE_Y: assign_ids(e->u.rotate_y.child, next); break;
    }
}
void wro_scene_finalize(wro_scene* s) {
    uint32_t next = 0;
    for (size_t i = 0; i < s->n_pool; ++i) s->pool[i]->prim_id = WRO_NONE;
    if (s->root) assign_ids(s->root, &next);
    s->n_prims = next;
}

/* ---- hit ------------------------------------------------------------------------------------------ */
/* hitrecord.zig:16-21 */
static void set_front_face_normal(wro_hit* rec, const ray* r, v3 outward) {
    rec->front_face = (v3_dot(r->direction, outward) < 0.0);
    rec->normal = rec->front_face ? outward : v3_neg(outward);
}

/* entity.zig:659-666 getSphereUv */
static void sphere_uv(v3 v, double uv[2]) {
    double theta = acos(-v.y);
    double phi = atan2(-v.z, v.x) + WRO_PI;
    uv[0] = phi / (2 * WRO_PI);
    uv[1] = theta / WRO_PI;
}

/* entity.zig:585-623 SphereEntity.hit */
static int sphere_hit(const wro_entity* e, const ray* r, ival trange, wro_hit* rec) {
    v3 center = e->u.sphere.is_moving ? v3_add(e->u.sphere.center, v3_scale(e->u.sphere.movement, r->time))
                                      : e->u.sphere.center;
    v3 oc = v3_sub(center, r->origin);
    double a = v3_dot(r->direction, r->direction);
    double h = v3_dot(r->direction, oc);
    double c = v3_dot(oc, oc) - e->u.sphere.radius * e->u.sphere.radius;
    double discriminant = h * h - a * c;
    if (discriminant < 0.0) return 0;
    double disc_sqrt = sqrt(discriminant);
    double root = (h - disc_sqrt) / a;
    if (!ival_surrounds(trange, root)) {
        root = (h + disc_sqrt) / a;
        if (!ival_surrounds(trange, root)) return 0;
    }
    rec->t = root;
    rec->point = ray_at(r, rec->t);
    v3 outward = v3_div(v3_sub(rec->point, center), v3_splat(e->u.sphere.radius));
    set_front_face_normal(rec, r, outward);
    sphere_uv(outward, rec->uv);
    rec->material = e->u.sphere.material;
    rec->prim_id = e->prim_id;
    return 1;
}

/* entity.zig:477-501 QuadEntity.hit */
static int quad_hit(const wro_entity* e, const ray* r, ival trange, wro_hit* rec) {
    double denom = v3_dot(e->u.quad.normal, r->direction);
    if (fabs(denom) < 1e-8) return 0;
    double t = (e->u.quad.offset - v3_dot(e->u.quad.normal, r->origin)) / denom;
    if (!ival_contains(trange, t)) return 0;
    v3 hit_point = ray_at(r, t);
    v3 planar = v3_sub(hit_point, e->u.quad.start);
    double alpha = v3_dot(e->u.quad.basis.w, v3_cross(planar, e->u.quad.basis.v));
    double beta = v3_dot(e->u.quad.basis.w, v3_cross(e->u.quad.basis.u, planar));
    ival unit = {0, 1};
    if (!(ival_contains(unit, alpha) && ival_contains(unit, beta))) return 0;
    rec->t = t;
    rec->point = hit_point;
    rec->material = e->u.quad.material;
    set_front_face_normal(rec, r, e->u.quad.normal);
    rec->uv[0] = alpha;
    rec->uv[1] = beta;
    rec->prim_id = e->prim_id;
    return 1;
}

static int box_hit(const wro_scene* s, const aabb* b, const ray* r, ival ray_t) {
    if (!s->no_cull) return aabb_hit(b, r, ray_t);
    return 1; /* validation mode: no culling at all == brute force over every leaf */
}

int wro_entity_hit(const wro_scene* s, const wro_entity* e, const ray* r, ival trange, wro_hit* rec) {
    switch (e->kind) {
        case WRO_ENT_SPHERE: return sphere_hit(e, r, trange, rec);
        case WRO_ENT_QUAD: return quad_hit(e, r, trange, rec);
        case WRO_ENT_COLLECTION: { /* entity.zig:342-368 */
            if (e->u.collection.bvh_root) return wro_entity_hit(s, e->u.collection.bvh_root, r, trange, rec);
            wro_hit tmp;
            memset(&tmp, 0, sizeof tmp);
            int hit_anything = 0;
            for (size_t i = 0; i < e->u.collection.len; ++i) {
                if (wro_entity_hit(s, e->u.collection.items[i], r, trange, &tmp)) {
                    hit_anything = 1;
                    trange.max = tmp.t;
                    *rec = tmp;
                }
            }
            return hit_anything;
        }
        case WRO_ENT_BVH_NODE: { /* entity.zig:286-303 */
            if (!box_hit(s, &e->box, r, trange)) return 0;
            int hit_left = e->u.bvh.left ? wro_entity_hit(s, e->u.bvh.left, r, trange, rec) : 0;
            ival tr = trange;
            if (hit_left) tr.max = rec->t;
            int hit_right = e->u.bvh.right ? wro_entity_hit(s, e->u.bvh.right, r, tr, rec) : 0;
            return hit_left || hit_right;
        }
        case WRO_ENT_TRANSLATE: { /* entity.zig:93-109 */
            ray rt = *r;
            rt.origin = v3_sub(r->origin, e->u.translate.offset);
            if (!wro_entity_hit(s, e->u.translate.child, &rt, trange, rec)) return 0;
            rec->point = v3_add(rec->point, e->u.translate.offset);
            return 1;
        }
        case WRO_ENT_ROTATE_Y: { /* entity.zig:169-205 */
            double sn = e->u.rotate_y.sin_theta, cs = e->u.rotate_y.cos_theta;
            ray rr;
            rr.origin = v3_make(cs * r->origin.x - sn * r->origin.z, r->origin.y, sn * r->origin.x + cs * r->origin.z);
            rr.direction = v3_make(cs * r->direction.x - sn * r->direction.z, r->direction.y,
                                   sn * r->direction.x + cs * r->direction.z);
            rr.time = r->time;
            if (!wro_entity_hit(s, e->u.rotate_y.child, &rr, trange, rec)) return 0;
            v3 p = rec->point, n = rec->normal;
            rec->point = v3_make(cs * p.x + sn * p.z, p.y, -sn * p.x + cs * p.z);
            rec->normal = v3_make(cs * n.x + sn * n.z, n.y, -sn * n.x + cs * n.z);
            return 1;
        }
    }
    return 0;
}

/* ---- light-sampling hooks ------------------------------------------------------------------------- */
/* entity.zig:503-518 QuadEntity.pdfValue, :626-644 SphereEntity.pdfValue, :371-378 EntityCollection.pdfValue;
 * every other variant answers 0 (entity.zig:47-55). */
double wro_entity_pdf_value(const wro_scene* s, const wro_entity* e, v3 origin, v3 direction) {
    (void)s;
    switch (e->kind) {
        case WRO_ENT_QUAD: {
            ray r = {origin, direction, 0.0};
            ival tr = {1e-3, INFINITY};
            wro_hit rec;
            memset(&rec, 0, sizeof rec);
            if (!quad_hit(e, &r, tr, &rec)) return 0.0;
            double dir_length_sq = v3_dot(direction, direction);
            double dist_sq = rec.t * rec.t * dir_length_sq;
            double cosine = fabs(v3_dot(direction, rec.normal)) / sqrt(dir_length_sq);
            return dist_sq / (cosine * e->u.quad.area);
        }
        case WRO_ENT_SPHERE: {
            ray r = {origin, direction, 0.0};
            ival tr = {1e-3, INFINITY};
            wro_hit rec;
            memset(&rec, 0, sizeof rec);
            if (!sphere_hit(e, &r, tr, &rec)) return 0.0;
            v3 diff = v3_sub(e->u.sphere.center, origin);
            double dist_sq = v3_dot(diff, diff);
            double cos_theta_max = sqrt(1.0 - e->u.sphere.radius * e->u.sphere.radius / dist_sq);
            double solid_angle = 2.0 * WRO_PI * (1.0 - cos_theta_max);
            return 1.0 / solid_angle;
        }
        case WRO_ENT_COLLECTION: {
            double weight = 1.0 / (double)e->u.collection.len;
            double sum = 0.0;
            for (size_t i = 0; i < e->u.collection.len; ++i)
                sum += weight * wro_entity_pdf_value(s, e->u.collection.items[i], origin, direction);
            return sum;
        }
        default: return 0.0;
    }
}

/* entity.zig:668-679 randomToSphere */
static v3 random_to_sphere(wro_rng* rng, double radius, double dist_sq) {
    double r1 = wro_rng_float(rng);
    double r2 = wro_rng_float(rng);
    double z = 1.0 + r2 * (sqrt(1.0 - radius * radius / dist_sq) - 1.0);
    double phi = 2.0 * WRO_PI * r1;
    double sz2 = sqrt(1.0 - z * z);
    double x = cos(phi) * sz2;
    double y = sin(phi) * sz2;
    return v3_make(x, y, z);
}

/* entity.zig:520-525 quad, :646-651 sphere, :381-386 collection; default (1,0,0) (entity.zig:58-65) */
v3 wro_entity_sample_direction(const wro_entity* e, wro_rng* rng, v3 origin) {
    switch (e->kind) {
        case WRO_ENT_QUAD: {
            v3 u = v3_scale(e->u.quad.basis.u, wro_rng_float(rng));
            v3 v = v3_scale(e->u.quad.basis.v, wro_rng_float(rng));
            v3 p = v3_add(v3_add(e->u.quad.start, u), v);
            return v3_sub(p, origin);
        }
        case WRO_ENT_SPHERE: {
            v3 direction = v3_sub(e->u.sphere.center, origin);
            double dist_sq = v3_dot(direction, direction);
            onb basis = onb_init(direction);
            return onb_transform(&basis, random_to_sphere(rng, e->u.sphere.radius, dist_sq));
        }
        case WRO_ENT_COLLECTION: {
            uint32_t idx = wro_rng_pick(rng, (uint32_t)e->u.collection.len);
            return wro_entity_sample_direction(e->u.collection.items[idx], rng, origin);
        }
        default: return v3_make(1, 0, 0);
    }
}

/* ---- textures (texture.zig, image.zig) ------------------------------------------------------------- */
static const uint8_t ERR_COLOR[3] = {255, 0, 255}; /* image.zig:5 */

/* image.zig:23-36 getPixel */
static const uint8_t* image_get_pixel(const wro_image* im, uint64_t x, uint64_t y) {
    if (im && im->data && im->height) {
        uint64_t cidx = x > (uint64_t)im->width - 1 ? (uint64_t)im->width - 1 : x;
        uint64_t ridx = y > (uint64_t)im->height - 1 ? (uint64_t)im->height - 1 : y;
        return im->data + (size_t)im->bytes_per_row * ridx + (size_t)im->num_components * cidx;
    }
    return ERR_COLOR;
}

/* texture.zig:70-77 pixelToColor */
static v3 pixel_to_color(const uint8_t* px) {
    const double scale = 1.0 / 255.0;
    v3 c = v3_make(scale * (double)px[0], scale * (double)px[1], scale * (double)px[2]);
    return v3_mul(c, c); /* linearizeColorSpace, math.zig:172-174 */
}

v3 wro_texture_value(const wro_texture* t, const double uv[2], v3 point) {
    switch (t->kind) {
        case WRO_TEX_SOLID: return t->color; /* texture.zig:89-93 */
        case WRO_TEX_CHECKER: {              /* texture.zig:111-118 */
            int32_t xi = (int32_t)floor(t->inv_scale * point.x);
            int32_t yi = (int32_t)floor(t->inv_scale * point.y);
            int32_t zi = (int32_t)floor(t->inv_scale * point.z);
            int32_t sum = xi + yi + zi;
            int32_t m = sum % 2;
            if (m < 0) m += 2; /* @mod */
            return wro_texture_value(m == 0 ? t->even : t->odd, uv, point);
        }
        case WRO_TEX_IMAGE: { /* texture.zig:49-68 */
            const wro_image* im = t->image;
            if (!im || im->height == 0) return pixel_to_color(ERR_COLOR);
            double u = wro_clamp(uv[0], 0.0, 1.0);
            double v = 1.0 - wro_clamp(uv[1], 0.0, 1.0);
            double fwidth = (double)im->width, fheight = (double)im->height;
            return pixel_to_color(image_get_pixel(im, (uint64_t)(u * fwidth), (uint64_t)(v * fheight)));
        }
    }
    return v3_splat(0);
}
