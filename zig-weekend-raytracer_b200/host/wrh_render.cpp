// wrh_render.cpp — Renderer.render (src/render.zig:29-74) on the B200 back end.
//
// The reference fans (row x 32-column) jobs out over a CPU thread pool and recurses through the IEntity tree per
// ray.  Here the same call flattens the tree into the POD arrays of include/wrt.h — the walk a Zig shim performs,
// INTEGRATION.md — and hands it to libwrt.so: wrt_upload_scene + wrt_render.  There is no CPU path.
#include <stdexcept>
#include <unordered_map>

#include "wrh_scene.hpp"

namespace wrh {

// ---- flattening ------------------------------------------------------------------------------------------------
namespace {

struct Flattener {
    FlatScene& out;
    std::unordered_map<const IEntity*, uint32_t> entity_ids;
    std::unordered_map<const IMaterial*, uint32_t> material_ids;
    std::unordered_map<const ITexture*, uint32_t> texture_ids;
    std::unordered_map<const Image*, uint32_t> image_ids;

    static void put(double dst[3], Vec3 v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; }

    uint32_t image(const Image* im) {
        auto it = image_ids.find(im);
        if (it != image_ids.end()) return it->second;
        wrt_image rec{};
        rec.width = im->width;
        rec.height = im->data.empty() ? 0 : im->height;
        rec.num_components = im->num_components;
        rec.bytes_per_row = im->bytes_per_row;
        rec.texel_offset = out.texels.size();
        out.texels.insert(out.texels.end(), im->data.begin(), im->data.end());
        out.images.push_back(rec);
        return image_ids[im] = static_cast<uint32_t>(out.images.size() - 1);
    }

    uint32_t texture(const ITexture* t) {
        auto it = texture_ids.find(t);
        if (it != texture_ids.end()) return it->second;
        const uint32_t id = static_cast<uint32_t>(out.textures.size());
        texture_ids[t] = id;
        out.textures.emplace_back();
        wrt_texture rec{};
        rec.kind = static_cast<uint32_t>(t->kind);
        rec.even = rec.odd = rec.image = WRT_NONE;
        put(rec.color, t->color);
        rec.inv_scale = t->inv_scale;
        if (t->kind == TextureKind::checkerboard) {
            rec.even = texture(t->tex_even);
            rec.odd = texture(t->tex_odd);
        } else if (t->kind == TextureKind::image) {
            rec.image = image(t->image);
        }
        out.textures[id] = rec;
        return id;
    }

    uint32_t material(const IMaterial* m) {
        auto it = material_ids.find(m);
        if (it != material_ids.end()) return it->second;
        wrt_material rec{};
        rec.kind = static_cast<uint32_t>(m->kind);
        rec.texture = m->texture ? texture(m->texture) : WRT_NONE;
        put(rec.albedo, m->albedo);
        rec.param = m->kind == MaterialKind::dielectric ? m->refraction_index : m->fuzz;
        out.materials.push_back(rec);
        return material_ids[m] = static_cast<uint32_t>(out.materials.size() - 1);
    }

    uint32_t entity(const IEntity* e) {
        auto it = entity_ids.find(e);
        if (it != entity_ids.end()) return it->second;
        const uint32_t id = static_cast<uint32_t>(out.entities.size());
        entity_ids[e] = id;
        out.entities.emplace_back();
        wrt_entity rec{};
        rec.kind = static_cast<uint32_t>(e->kind);
        rec.a = rec.b = rec.c = WRT_NONE;
        put(rec.bbox_min, e->aabb.min);
        put(rec.bbox_max, e->aabb.max);
        switch (e->kind) {
            case EntityKind::sphere: {
                wrt_sphere s{};
                put(s.center, e->center);
                s.radius = e->radius;
                put(s.movement, e->movement_direction);
                s.material = material(e->material);
                s.is_moving = e->b_is_moving ? 1u : 0u;
                out.spheres.push_back(s);
                rec.a = static_cast<uint32_t>(out.spheres.size() - 1);
                break;
            }
            case EntityKind::quad: {
                wrt_quad q{};
                put(q.start, e->start_point);
                put(q.u, e->basis.u);
                put(q.v, e->basis.v);
                put(q.w, e->basis.w);
                put(q.normal, e->normal);
                q.offset = e->offset;
                q.area = e->area;
                q.material = material(e->material);
                out.quads.push_back(q);
                rec.a = static_cast<uint32_t>(out.quads.size() - 1);
                break;
            }
            case EntityKind::collection: {
                if (e->bvh_root) rec.c = entity(e->bvh_root);
                std::vector<uint32_t> ids;
                ids.reserve(e->entities.size());
                for (const IEntity* child : e->entities) ids.push_back(entity(child));
                rec.a = static_cast<uint32_t>(out.children.size());
                rec.b = static_cast<uint32_t>(ids.size());
                out.children.insert(out.children.end(), ids.begin(), ids.end());
                break;
            }
            case EntityKind::bvh_node:
                rec.a = entity(e->left);
                rec.b = entity(e->right);
                break;
            case EntityKind::translate:
                put(rec.p, e->translate_offset);
                rec.a = entity(e->entity);
                break;
            case EntityKind::rotate_y:
                rec.p[0] = e->sin_theta;
                rec.p[1] = e->cos_theta;
                rec.a = entity(e->entity);
                break;
        }
        out.entities[id] = rec;
        return id;
    }
};

}  // namespace

void flattenScene(const IEntity& root, const IEntity* lights, FlatScene& out) {
    out = FlatScene();
    Flattener f{out, {}, {}, {}, {}};
    const uint32_t root_id = f.entity(&root);
    const uint32_t lights_id = lights ? f.entity(lights) : WRT_NONE;
    wrt_scene& v = out.view;
    v.abi_version = WRT_ABI_VERSION;
    v.root = root_id;
    v.lights = lights_id;
    v.n_entities = static_cast<uint32_t>(out.entities.size()); v.entities = out.entities.data();
    v.n_children = static_cast<uint32_t>(out.children.size()); v.children = out.children.data();
    v.n_spheres = static_cast<uint32_t>(out.spheres.size()); v.spheres = out.spheres.data();
    v.n_quads = static_cast<uint32_t>(out.quads.size()); v.quads = out.quads.data();
    v.n_materials = static_cast<uint32_t>(out.materials.size()); v.materials = out.materials.data();
    v.n_textures = static_cast<uint32_t>(out.textures.size()); v.textures = out.textures.data();
    v.n_images = static_cast<uint32_t>(out.images.size()); v.images = out.images.data();
    v.texels = out.texels.data();
    v.texel_bytes = out.texels.size();
}

uint64_t FlatScene::inputBytes() const {
    return entities.size() * sizeof(wrt_entity) + children.size() * sizeof(uint32_t) + spheres.size() * sizeof(wrt_sphere) +
           quads.size() * sizeof(wrt_quad) + materials.size() * sizeof(wrt_material) + textures.size() * sizeof(wrt_texture) +
           images.size() * sizeof(wrt_image) + texels.size() + sizeof(wrt_scene);
}

// ---- back end ----------------------------------------------------------------------------------------------------
Backend::Backend(int cuda_device) : Backend(std::vector<int>{cuda_device}) {}
Backend::Backend(const std::vector<int>& cuda_devices) {
    const int rc = wrt_group_create(cuda_devices.data(), static_cast<int>(cuda_devices.size()), &group_);
    if (rc != WRT_OK) throw std::runtime_error(std::string("wrt_group_create failed: ") + wrt_group_last_error(nullptr));
}
Backend::~Backend() { wrt_group_destroy(group_); }

void Renderer::render(const Camera& camera, const IEntity& entity, Framebuffer& framebuffer) {  // render.zig:29
    if (!backend) throw std::runtime_error("Renderer.render: no CUDA back end attached (there is no CPU fallback)");
    wrt_group* group = backend->group();
    FlatScene flat;
    flattenScene(entity, light_entities, flat);
    if (wrt_group_upload_scene(group, &flat.view) != WRT_OK)
        throw std::runtime_error(std::string("wrt_group_upload_scene: ") + wrt_group_last_error(group));

    const wrt_camera cam = camera.view(framebuffer.num_cols, framebuffer.num_rows);  // camera.getViewport, render.zig:47
    wrt_params p{};
    p.width = static_cast<uint32_t>(framebuffer.num_cols);
    p.height = static_cast<uint32_t>(framebuffer.num_rows);
    p.samples_per_pixel = static_cast<uint32_t>(samples_per_pixel);
    p.max_ray_bounce_depth = static_cast<uint32_t>(max_ray_bounce_depth);
    const Real bg[3] = {background_color.x, background_color.y, background_color.z};
    const Real cc[3] = {clear_color.x, clear_color.y, clear_color.z};
    for (int k = 0; k < 3; ++k) { p.background_color[k] = bg[k]; p.clear_color[k] = cc[k]; }
    p.seed = seed;
    p.row_shard_index = 0;
    p.row_shard_count = 1;
    p.cull_mode = cull_mode;
    p.flags = flags;
    // framebuffer.clear(clear_color) + the job fan-out over the devices + `buffer[..] += color` + the gather all happen on the
    // device side (render.zig:33, 55-73)
    if (wrt_group_render(group, &cam, &p, framebuffer.buffer.data(), framebuffer.pixelStrideBytes()) != WRT_OK)
        throw std::runtime_error(std::string("wrt_group_render: ") + wrt_group_last_error(group));
    wrt_stats st{};
    wrt_group_get_stats(group, &st);
    last_stats.paths = st.paths;
    last_stats.rays = st.rays;
    last_stats.render_ms = st.render_ms;
    last_stats.kernel_ms = st.kernel_ms;
    last_stats.upload_ms = st.upload_ms;
    last_stats.gather_ms = st.gather_ms;
    last_stats.kernel_ms_min = st.kernel_ms_min;
    last_stats.kernel_ms_max = st.kernel_ms_max;
    last_stats.n_devices = st.n_devices;
    last_stats.cull_mode_used = st.cull_mode_used;
    last_stats.ref_boxes_loose = st.ref_boxes_loose;
}

}  // namespace wrh
