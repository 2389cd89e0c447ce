/*
 * wro_flat.c — ORACLE (test infrastructure, not product code).
 * Tree <-> flat conversion at the C-ABI boundary (include/wrt.h).  wro_scene_flatten is the restatement of the
 * walk a Zig shim performs over IEntity / IMaterial / ITexture; wro_scene_from_flat lets the
 * oracle trace any scene that was handed to the product through the ABI.
 */
#include <stdlib.h>
#include <string.h>

#include "wro.h"
#include "wro_scene.h"

typedef struct flat_builder {
    wrt_entity* entities; size_t n_entities, cap_entities;
    uint32_t* children; size_t n_children, cap_children;
    wrt_sphere* spheres; size_t n_spheres, cap_spheres;
    wrt_quad* quads; size_t n_quads, cap_quads;
    wrt_material* materials;
    wrt_texture* textures;
    wrt_image* images;
    uint8_t* texels;
} flat_builder;

#define GROW(arr, n, cap)                                      \
    do {                                                       \
        if ((n) == (cap)) {                                    \
            (cap) = (cap) ? (cap) * 2 : 64;                    \
            (arr) = realloc((arr), (cap) * sizeof *(arr));     \
        }                                                      \
    } while (0)

static void put3(double dst[3], v3 v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; }
static v3 get3(const double a[3]) { return v3_make(a[0], a[1], a[2]); }

static uint32_t flatten_entity(flat_builder* fb, wro_entity* e) {
    if (e->flat_id != WRT_NONE) return e->flat_id;
    GROW(fb->entities, fb->n_entities, fb->cap_entities);
    uint32_t id = (uint32_t)fb->n_entities++;
    e->flat_id = id;
    wrt_entity rec;
    memset(&rec, 0, sizeof rec);
    rec.a = rec.b = rec.c = WRT_NONE;
    put3(rec.bbox_min, e->box.min);
    put3(rec.bbox_max, e->box.max);
    switch (e->kind) {
        case WRO_ENT_SPHERE: {
            rec.kind = WRT_ENT_SPHERE;
            GROW(fb->spheres, fb->n_spheres, fb->cap_spheres);
            wrt_sphere* sp = &fb->spheres[fb->n_spheres];
            memset(sp, 0, sizeof *sp);
            put3(sp->center, e->u.sphere.center);
            sp->radius = e->u.sphere.radius;
            put3(sp->movement, e->u.sphere.movement);
            sp->material = e->u.sphere.material->index;
            sp->is_moving = (uint32_t)e->u.sphere.is_moving;
            rec.a = (uint32_t)fb->n_spheres++;
            break;
        }
        case WRO_ENT_QUAD: {
            rec.kind = WRT_ENT_QUAD;
            GROW(fb->quads, fb->n_quads, fb->cap_quads);
            wrt_quad* q = &fb->quads[fb->n_quads];
            memset(q, 0, sizeof *q);
            put3(q->start, e->u.quad.start);
            put3(q->u, e->u.quad.basis.u);
            put3(q->v, e->u.quad.basis.v);
            put3(q->w, e->u.quad.basis.w);
            put3(q->normal, e->u.quad.normal);
            q->offset = e->u.quad.offset;
            q->area = e->u.quad.area;
            q->material = e->u.quad.material->index;
            rec.a = (uint32_t)fb->n_quads++;
            break;
        }
        case WRO_ENT_COLLECTION: {
            rec.kind = WRT_ENT_COLLECTION;
            size_t len = e->u.collection.len;
            uint32_t* ids = malloc((len ? len : 1) * sizeof *ids);
            /* children first (they may recurse and append their own lists) */
            if (e->u.collection.bvh_root) rec.c = flatten_entity(fb, e->u.collection.bvh_root);
            for (size_t i = 0; i < len; ++i) ids[i] = flatten_entity(fb, e->u.collection.items[i]);
            while (fb->n_children + len > fb->cap_children) {
                fb->cap_children = fb->cap_children ? fb->cap_children * 2 : 256;
                fb->children = realloc(fb->children, fb->cap_children * sizeof *fb->children);
            }
            rec.a = (uint32_t)fb->n_children;
            rec.b = (uint32_t)len;
            memcpy(fb->children + fb->n_children, ids, len * sizeof *ids);
            fb->n_children += len;
            free(ids);
            break;
        }
        case WRO_ENT_BVH_NODE:
            rec.kind = WRT_ENT_BVH_NODE;
            rec.a = flatten_entity(fb, e->u.bvh.left);
            rec.b = flatten_entity(fb, e->u.bvh.right);
            break;
        case WRO_ENT_TRANSLATE:
            rec.kind = WRT_ENT_TRANSLATE;
            put3(rec.p, e->u.translate.offset);
            rec.a = flatten_entity(fb, e->u.translate.child);
            break;
        case WRO_ENT_ROTATE_Y:
            rec.kind = WRT_ENT_ROTATE_Y;
            rec.p[0] = e->u.rotate_y.sin_theta;
            rec.p[1] = e->u.rotate_y.cos_theta;
            rec.a = flatten_entity(fb, e->u.rotate_y.child);
            break;
    }
    fb->entities[id] = rec; /* the array may have moved while recursing */
    return id;
}

int wro_scene_flatten(const wro_scene* s, wro_flat* out) {
    if (!s || !out || !s->root) return -1;
    flat_builder* fb = calloc(1, sizeof *fb);
    for (size_t i = 0; i < s->n_pool; ++i) s->pool[i]->flat_id = WRT_NONE;

    fb->materials = calloc(s->n_materials ? s->n_materials : 1, sizeof *fb->materials);
    for (size_t i = 0; i < s->n_materials; ++i) {
        const wro_material* m = s->materials[i];
        wrt_material* o = &fb->materials[i];
        o->kind = (uint32_t)m->kind; /* enum orders match include/wrt.h */
        o->texture = m->texture ? m->texture->index : WRT_NONE;
        put3(o->albedo, m->albedo);
        o->param = m->param;
    }
    fb->textures = calloc(s->n_textures ? s->n_textures : 1, sizeof *fb->textures);
    for (size_t i = 0; i < s->n_textures; ++i) {
        const wro_texture* t = s->textures[i];
        wrt_texture* o = &fb->textures[i];
        o->kind = (uint32_t)t->kind;
        o->even = t->even ? t->even->index : WRT_NONE;
        o->odd = t->odd ? t->odd->index : WRT_NONE;
        o->image = t->image ? t->image->index : WRT_NONE;
        put3(o->color, t->color);
        o->inv_scale = t->inv_scale;
    }
    fb->images = calloc(s->n_images ? s->n_images : 1, sizeof *fb->images);
    uint64_t texel_bytes = 0;
    for (size_t i = 0; i < s->n_images; ++i) texel_bytes += (uint64_t)s->images[i]->bytes_per_row * s->images[i]->height;
    fb->texels = malloc(texel_bytes ? texel_bytes : 1);
    uint64_t off = 0;
    for (size_t i = 0; i < s->n_images; ++i) {
        const wro_image* im = s->images[i];
        wrt_image* o = &fb->images[i];
        o->width = im->width; o->height = im->height;
        o->num_components = im->num_components; o->bytes_per_row = im->bytes_per_row;
        o->texel_offset = off;
        uint64_t nbytes = (uint64_t)im->bytes_per_row * im->height;
        if (nbytes) memcpy(fb->texels + off, im->data, nbytes);
        off += nbytes;
    }

    uint32_t root = flatten_entity(fb, s->root);
    uint32_t lights = s->lights ? flatten_entity(fb, s->lights) : WRT_NONE;

    memset(out, 0, sizeof *out);
    out->owner = fb;
    wrt_scene* f = &out->scene;
    f->abi_version = WRT_ABI_VERSION;
    f->root = root;
    f->lights = lights;
    f->n_entities = (uint32_t)fb->n_entities; f->entities = fb->entities;
    f->n_children = (uint32_t)fb->n_children; f->children = fb->children;
    f->n_spheres = (uint32_t)fb->n_spheres; f->spheres = fb->spheres;
    f->n_quads = (uint32_t)fb->n_quads; f->quads = fb->quads;
    f->n_materials = (uint32_t)s->n_materials; f->materials = fb->materials;
    f->n_textures = (uint32_t)s->n_textures; f->textures = fb->textures;
    f->n_images = (uint32_t)s->n_images; f->images = fb->images;
    f->texels = fb->texels; f->texel_bytes = texel_bytes;
    return 0;
}

void wro_flat_free(wro_flat* f) {
    if (!f || !f->owner) return;
    flat_builder* fb = f->owner;
    free(fb->entities); free(fb->children); free(fb->spheres); free(fb->quads);
    free(fb->materials); free(fb->textures); free(fb->images); free(fb->texels);
    free(fb);
    memset(f, 0, sizeof *f);
}

/* ---- flat -> tree ------------------------------------------------------------------------------------ */
static aabb box_from_flat(const wrt_entity* r) {
    aabb b;
    b.min = get3(r->bbox_min); b.max = get3(r->bbox_max);
    b.x.min = b.min.x; b.x.max = b.max.x;
    b.y.min = b.min.y; b.y.max = b.max.y;
    b.z.min = b.min.z; b.z.max = b.max.z;
    return b;
}

wro_scene* wro_scene_from_flat(const wrt_scene* f) {
    if (!f || f->abi_version != WRT_ABI_VERSION || f->root >= f->n_entities) return NULL;
    wro_scene* s = wro_scene_new();
    wro_image** images = calloc(f->n_images ? f->n_images : 1, sizeof *images);
    for (uint32_t i = 0; i < f->n_images; ++i) {
        const wrt_image* im = &f->images[i];
        images[i] = wro_add_image(s, im->width, im->height, im->num_components,
                                  im->height ? f->texels + im->texel_offset : NULL);
        if (im->height) images[i]->bytes_per_row = im->bytes_per_row;
    }
    wro_texture** textures = calloc(f->n_textures ? f->n_textures : 1, sizeof *textures);
    for (uint32_t i = 0; i < f->n_textures; ++i) { /* allocate first: checker children may follow their parent */
        textures[i] = wro_tex_solid(s, v3_splat(0));
    }
    for (uint32_t i = 0; i < f->n_textures; ++i) {
        const wrt_texture* t = &f->textures[i];
        wro_texture* o = textures[i];
        o->kind = (int)t->kind;
        o->color = get3(t->color);
        o->inv_scale = t->inv_scale;
        o->even = t->even != WRT_NONE ? textures[t->even] : NULL;
        o->odd = t->odd != WRT_NONE ? textures[t->odd] : NULL;
        o->image = t->image != WRT_NONE ? images[t->image] : NULL;
    }
    wro_material** materials = calloc(f->n_materials ? f->n_materials : 1, sizeof *materials);
    for (uint32_t i = 0; i < f->n_materials; ++i) {
        const wrt_material* m = &f->materials[i];
        wro_material* o = wro_mat_dielectric(s, m->param);
        o->kind = (int)m->kind;
        o->texture = m->texture != WRT_NONE ? textures[m->texture] : NULL;
        o->albedo = get3(m->albedo);
        materials[i] = o;
    }
    wro_entity** ents = calloc(f->n_entities, sizeof *ents);
    for (uint32_t i = 0; i < f->n_entities; ++i) ents[i] = wro_entity_raw(s, (int)f->entities[i].kind);
    for (uint32_t i = 0; i < f->n_entities; ++i) {
        const wrt_entity* r = &f->entities[i];
        wro_entity* e = ents[i];
        e->box = box_from_flat(r);
        switch (r->kind) {
            case WRT_ENT_SPHERE: {
                const wrt_sphere* sp = &f->spheres[r->a];
                e->u.sphere.center = get3(sp->center);
                e->u.sphere.radius = sp->radius;
                e->u.sphere.material = materials[sp->material];
                e->u.sphere.is_moving = (int)sp->is_moving;
                e->u.sphere.movement = get3(sp->movement);
                break;
            }
            case WRT_ENT_QUAD: {
                const wrt_quad* q = &f->quads[r->a];
                e->u.quad.start = get3(q->start);
                e->u.quad.basis = onb_from_vectors(get3(q->u), get3(q->v), get3(q->w));
                e->u.quad.normal = get3(q->normal);
                e->u.quad.offset = q->offset;
                e->u.quad.area = q->area;
                e->u.quad.material = materials[q->material];
                break;
            }
            case WRT_ENT_COLLECTION: {
                e->u.collection.len = e->u.collection.cap = r->b;
                e->u.collection.items = malloc((r->b ? r->b : 1) * sizeof(wro_entity*));
                for (uint32_t k = 0; k < r->b; ++k) e->u.collection.items[k] = ents[f->children[r->a + k]];
                e->u.collection.bvh_root = r->c != WRT_NONE ? ents[r->c] : NULL;
                break;
            }
            case WRT_ENT_BVH_NODE:
                e->u.bvh.left = ents[r->a];
                e->u.bvh.right = ents[r->b];
                break;
            case WRT_ENT_TRANSLATE:
                e->u.translate.offset = get3(r->p);
                e->u.translate.child = ents[r->a];
                break;
            case WRT_ENT_ROTATE_Y:
                e->u.rotate_y.sin_theta = r->p[0];
                e->u.rotate_y.cos_theta = r->p[1];
                e->u.rotate_y.child = ents[r->a];
                break;
        }
    }
    s->root = ents[f->root];
    s->lights = f->lights != WRT_NONE ? ents[f->lights] : NULL;
    free(ents); free(materials); free(textures); free(images);
    wro_scene_finalize(s);
    return s;
}

/* ---- topology known-answer helper --------------------------------------------------------------------- */
int wro_scene_prim_table(const wro_scene* s, uint32_t* kinds, uint32_t* materials, double* centers_xyz) {
    for (size_t i = 0; i < s->n_pool; ++i) {
        const wro_entity* e = s->pool[i];
        if (e->prim_id == WRT_NONE) continue;
        uint32_t id = e->prim_id;
        kinds[id] = (e->kind == WRO_ENT_QUAD);
        materials[id] = (e->kind == WRO_ENT_QUAD) ? e->u.quad.material->index : e->u.sphere.material->index;
        centers_xyz[3 * id + 0] = 0.5 * (e->box.min.x + e->box.max.x);
        centers_xyz[3 * id + 1] = 0.5 * (e->box.min.y + e->box.max.y);
        centers_xyz[3 * id + 2] = 0.5 * (e->box.min.z + e->box.max.z);
    }
    return 0;
}
