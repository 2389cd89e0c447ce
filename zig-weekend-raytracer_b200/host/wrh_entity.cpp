// wrh_entity.cpp — scene-construction half of the host mirror: entity / material / texture factories, the BVH
// builder, Camera / Viewport / Framebuffer.  Every function follows the reference routine it names, quirks included
// (SURVEY.md A.9), because the device consumes the tree exactly as the reference would have built it.
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <limits>

#include "wrh_scene.hpp"

namespace wrh {

// ---- images -------------------------------------------------------------------------------------------------------
Image Image::fromPixels(uint32_t w, uint32_t h, uint32_t comps, const uint8_t* px) {
    Image im;
    im.width = w; im.height = h; im.num_components = comps; im.bytes_per_row = w * comps;
    if (px && w && h) im.data.assign(px, px + static_cast<size_t>(im.bytes_per_row) * h);
    else im.height = 0;
    return im;
}

bool Image::loadPnm(const std::string& path, Image& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::string magic;
    f >> magic;
    if (magic != "P6" && magic != "P5") return false;
    auto next_int = [&](long& v) {
        for (;;) {
            int c = f.peek();
            if (c == '#') { std::string line; std::getline(f, line); }
            else if (c == ' ' || c == '\n' || c == '\r' || c == '\t') f.get();
            else break;
        }
        return static_cast<bool>(f >> v);
    };
    long w = 0, h = 0, maxv = 0;
    if (!next_int(w) || !next_int(h) || !next_int(maxv) || maxv != 255 || w <= 0 || h <= 0) return false;
    f.get();  // single whitespace before the raster
    const uint32_t comps = magic == "P6" ? 3 : 1;
    std::vector<uint8_t> px(static_cast<size_t>(w) * h * comps);
    if (!f.read(reinterpret_cast<char*>(px.data()), static_cast<std::streamsize>(px.size()))) return false;
    out = fromPixels(static_cast<uint32_t>(w), static_cast<uint32_t>(h), comps, px.data());
    return true;
}

// ---- textures / materials ----------------------------------------------------------------------------------------------
ITexture SolidColorTexture::initTexture(Color c) {  // texture.zig:85-87
    ITexture t;
    t.kind = TextureKind::solid_color;
    t.color = c;
    return t;
}
ITexture CheckerboardTexture::initTexture(Real inv_scale, const ITexture* even, const ITexture* odd) {  // texture.zig:103-109
    ITexture t;
    t.kind = TextureKind::checkerboard;
    t.inv_scale = inv_scale;
    t.tex_even = even;
    t.tex_odd = odd;
    return t;
}
ITexture ImageTexture::initTexture(const Image* image) {  // texture.zig:38-42
    ITexture t;
    t.kind = TextureKind::image;
    t.image = image;
    return t;
}
IMaterial LambertianMaterial::initMaterial(const ITexture* t) {  // material.zig:104-106
    IMaterial m;
    m.kind = MaterialKind::lambertian;
    m.texture = t;
    return m;
}
IMaterial IsotropicMaterial::initMaterial(const ITexture* t) {  // material.zig:132-134
    IMaterial m;
    m.kind = MaterialKind::isotropic;
    m.texture = t;
    return m;
}
IMaterial MetalMaterial::initMaterial(Color albedo, Real fuzz) {  // material.zig:159-161
    IMaterial m;
    m.kind = MaterialKind::metal;
    m.albedo = albedo;
    m.fuzz = fuzz;
    return m;
}
IMaterial DielectricMaterial::initMaterial(Real refraction_index) {  // material.zig:186-188
    IMaterial m;
    m.kind = MaterialKind::dielectric;
    m.refraction_index = refraction_index;
    return m;
}
IMaterial DiffuseLightEmissiveMaterial::initMaterial(const ITexture* t) {  // material.zig:84-86
    IMaterial m;
    m.kind = MaterialKind::diffuse_emissive;
    m.texture = t;
    return m;
}

// ---- entities ----------------------------------------------------------------------------------------------------------
IEntity* SphereEntity::initEntity(EntityPool& pool, Point3 center, Real radius, const IMaterial* material) {  // entity.zig:545-561
    const Vec3 rvec = Vec3::splat(radius);
    IEntity* e = pool.create();
    e->kind = EntityKind::sphere;
    e->center = center;
    e->radius = radius;
    e->material = material;
    e->aabb = AABB::init(center - rvec, center + rvec);
    return e;
}

IEntity* SphereEntity::initEntityAnimated(EntityPool& pool, Point3 c0, Point3 c1, Real radius, const IMaterial* material) {  // entity.zig:563-583
    const Vec3 rvec = Vec3::splat(radius);
    IEntity* e = pool.create();
    e->kind = EntityKind::sphere;
    e->center = c0;
    e->radius = radius;
    e->material = material;
    e->b_is_moving = true;
    e->movement_direction = c1 - c0;
    e->aabb = AABB::init(c0 - rvec, c0 + rvec).unionWith(AABB::init(c1 - rvec, c1 + rvec));
    return e;
}

IEntity* QuadEntity::initEntity(EntityPool& pool, Point3 start, Vec3 axis1, Vec3 axis2, const IMaterial* material) {  // entity.zig:444-475
    const Vec3 n = cross(axis1, axis2);
    const Vec3 axis3 = n / Vec3::splat(dot(n, n));
    const Vec3 normal_unit = normalize(n);
    const Real plane_offset = dot(normal_unit, start);
    const AABB diag1 = AABB::init(start, start + axis1 + axis2);
    const AABB diag2 = AABB::init(start + axis1, start + axis2);
    IEntity* e = pool.create();
    e->kind = EntityKind::quad;
    e->start_point = start;
    e->basis = OrthoBasis{axis1, axis2, axis3};
    e->normal = normal_unit;
    e->offset = plane_offset;
    e->area = length(n);
    e->material = material;
    e->aabb = diag1.unionWith(diag2);
    return e;
}

IEntity* EntityCollection::initEntity(EntityPool& pool) {  // entity.zig:317-321
    IEntity* e = pool.create();
    e->kind = EntityKind::collection;
    return e;
}
void EntityCollection::add(IEntity* self, IEntity* e) {  // entity.zig:328-336
    self->entities.push_back(e);
    self->aabb = self->aabb.unionWith(e->boundingBox());
}
void EntityCollection::createBvhTree(IEntity* self, EntityPool& pool) {  // entity.zig:338-340
    self->bvh_root = BVHNodeEntity::initEntity(pool, self->entities, 0, self->entities.size());
}

// entity.zig:226-267.  The split sorts with std.sort.pdq (unstable, part of the Zig standard library); a stable sort
// is used here — identical for slices <= 20 items, where pdq is an insertion sort, and ties beyond that only
// reorder exactly coincident hits.
IEntity* BVHNodeEntity::initEntity(EntityPool& pool, std::vector<IEntity*>& entities, size_t start, size_t end) {
    IEntity* node = pool.create();
    node->kind = EntityKind::bvh_node;
    const size_t span = end - start;
    if (span == 1) {
        node->left = entities[start];
        node->right = entities[start];
    } else if (span == 2) {
        node->left = entities[start];
        node->right = entities[start + 1];
    } else {
        AABB bbox;  // default box: the origin is part of every split extent (A.9-3)
        for (size_t i = start; i < end; ++i) bbox = bbox.unionWith(entities[i]->boundingBox());
        const Axis axis = bbox.longestAxis();
        std::stable_sort(entities.begin() + static_cast<long>(start), entities.begin() + static_cast<long>(end),
                         [axis](const IEntity* a, const IEntity* b) {  // boxCmp, entity.zig:212-216
                             return a->boundingBox().axisInterval(axis).min < b->boundingBox().axisInterval(axis).min;
                         });
        const size_t mid = start + span / 2;
        node->left = initEntity(pool, entities, start, mid);
        node->right = initEntity(pool, entities, mid, end);
    }
    node->aabb = node->left->boundingBox().unionWith(node->right->boundingBox());
    return node;
}

IEntity* Translate::initEntity(EntityPool& pool, Vec3 offset, IEntity* child) {  // entity.zig:75-87
    IEntity* e = pool.create();
    e->kind = EntityKind::translate;
    e->translate_offset = offset;
    e->entity = child;
    e->aabb = child->boundingBox().offset(offset);
    return e;
}

IEntity* RotateY::initEntity(EntityPool& pool, Real angle_degrees, IEntity* child) {  // entity.zig:120-163
    const Real theta = degreesToRadians(angle_degrees);
    const Real sin_theta = std::sin(theta), cos_theta = std::cos(theta);
    const AABB& bbox = child->boundingBox();
    Vec3 mn = Vec3::splat(std::numeric_limits<Real>::infinity());
    Vec3 mx = Vec3::splat(-std::numeric_limits<Real>::infinity());
    for (int i = 0; i < 2; ++i) {
        const Real fi = i;
        const Real x = fi * bbox.x.max + (1.0 - fi) * bbox.x.min;
        for (int j = 0; j < 2; ++j) {
            const Real fj = j;
            const Real y = fj * bbox.x.max + (1.0 - fj) * bbox.y.min;  // x.max on purpose: entity.zig:139 (A.9-5)
            for (int k = 0; k < 2; ++k) {
                const Real fk = k;
                const Real z = fk * bbox.x.max + (1.0 - fk) * bbox.z.min;  // entity.zig:143
                const Vec3 tester{cos_theta * x + sin_theta * z, y, -sin_theta * x + cos_theta * z};
                mn = vmin(mn, tester);
                mx = vmax(mx, tester);
            }
        }
    }
    IEntity* e = pool.create();
    e->kind = EntityKind::rotate_y;
    e->sin_theta = sin_theta;
    e->cos_theta = cos_theta;
    e->entity = child;
    e->aabb = AABB::init(mn, mx);
    return e;
}

IEntity* createBoxEntity(EntityPool& pool, Point3 a, Point3 b, const IMaterial* material) {  // entity.zig:390-426
    IEntity* sides = EntityCollection::initEntity(pool);
    const Vec3 mn = vmin(a, b), mx = vmax(a, b);
    const Vec3 diff = mx - mn;
    const Vec3 dx{diff.x, 0, 0}, dy{0, diff.y, 0}, dz{0, 0, diff.z};
    struct Side { Point3 p0; Vec3 u, v; };
    const Side sides_data[6] = {
        {{mn.x, mn.y, mx.z}, dx, dy},    // front
        {{mx.x, mn.y, mx.z}, -dz, dy},   // right
        {{mx.x, mn.y, mn.z}, -dx, dy},   // back
        {{mn.x, mn.y, mn.z}, dz, dy},    // left
        {{mn.x, mx.y, mx.z}, dx, -dz},   // top
        {{mn.x, mn.y, mn.z}, dx, dz},    // bottom
    };
    for (const Side& s : sides_data) EntityCollection::add(sides, QuadEntity::initEntity(pool, s.p0, s.u, s.v, material));
    return sides;
}

// ---- camera --------------------------------------------------------------------------------------------------------------
Framebuffer Framebuffer::init(size_t height, size_t width) {  // camera.zig:18-25
    Framebuffer fb;
    fb.num_rows = height;
    fb.num_cols = width;
    fb.buffer.assign(height * width * kLanes, 0.0);
    return fb;
}
void Framebuffer::clear(Color c) {  // camera.zig:31-33
    for (size_t i = 0; i < num_rows * num_cols; ++i) {
        Real* px = &buffer[i * kLanes];
        px[0] = c.x; px[1] = c.y; px[2] = c.z;
        for (size_t k = 3; k < kLanes; ++k) px[k] = 0.0;
    }
}

Camera Camera::init(Point3 look_from, Point3 look_at, Vec3 view_up, Real fov_vertical, Real lens_focus_dist,
                    Real defocus_angle_degrees) {  // camera.zig:61-90
    Camera c;
    const Vec3 w = normalize(look_from - look_at);
    const Vec3 u = normalize(cross(view_up, w));
    const Vec3 v = cross(w, u);
    const Vec3 defocus_radius = Vec3::splat(lens_focus_dist * std::tan(degreesToRadians(defocus_angle_degrees / 2.0)));
    c.coordinate_basis = {u, v, w};
    c.position = look_from;
    c.fov_vertical = fov_vertical;
    c.b_is_depth_of_field = defocus_angle_degrees > 0.0;
    c.lens_focus_dist = lens_focus_dist;
    c.defocus_radius = defocus_radius;
    c.defocus_disk_u = u * defocus_radius;
    c.defocus_disk_v = v * defocus_radius;
    return c;
}

Viewport Viewport::init(size_t image_width, size_t image_height, Real aspect_ratio, Real fov_vertical,
                        Real lens_focus_distance, Point3 look_from, const CoordinateBasis& basis) {  // camera.zig:117-157
    Viewport vp;
    const Real theta = degreesToRadians(fov_vertical);
    const Real h = std::tan(theta / 2.0);
    const Real viewport_height = 2.0 * h * lens_focus_distance;
    const Real viewport_width = viewport_height * aspect_ratio;
    const Vec3 viewport_u = Vec3::splat(viewport_width) * basis.u;
    const Vec3 viewport_v = Vec3::splat(-viewport_height) * basis.v;
    const Point3 upper_left = look_from - (Vec3::splat(lens_focus_distance) * basis.w) - viewport_u / Vec3::splat(2) -
                              viewport_v / Vec3::splat(2);
    vp.width = viewport_width;
    vp.height = viewport_height;
    vp.upper_left_corner = upper_left;
    vp.u = viewport_u;
    vp.v = viewport_v;
    vp.pixel_delta_u = viewport_u / Vec3::splat(static_cast<Real>(image_width));
    vp.pixel_delta_v = viewport_v / Vec3::splat(static_cast<Real>(image_height));
    vp.pixel00_loc = upper_left + Vec3::splat(0.5) * (vp.pixel_delta_u + vp.pixel_delta_v);
    return vp;
}

Viewport Camera::getViewport(const Framebuffer& fb) const {  // camera.zig:92-102
    return Viewport::init(fb.num_cols, fb.num_rows, fb.getAspectRatio(), fov_vertical, lens_focus_dist, position, coordinate_basis);
}

wrt_camera Camera::view(size_t image_width, size_t image_height) const {
    const Real aspect = static_cast<Real>(image_width) / static_cast<Real>(image_height);
    const Viewport vp = Viewport::init(image_width, image_height, aspect, fov_vertical, lens_focus_dist, position, coordinate_basis);
    wrt_camera c{};
    auto put = [](double dst[3], Vec3 v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; };
    put(c.position, position);
    put(c.pixel00_loc, vp.pixel00_loc);
    put(c.pixel_delta_u, vp.pixel_delta_u);
    put(c.pixel_delta_v, vp.pixel_delta_v);
    put(c.defocus_disk_u, defocus_disk_u);
    put(c.defocus_disk_v, defocus_disk_v);
    c.is_depth_of_field = b_is_depth_of_field ? 1u : 0u;
    return c;
}

}  // namespace wrh
