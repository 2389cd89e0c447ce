// wrh_scene.hpp — host-side mirror of the reference's scene model and render entry point, so that code written against
// the reference reads the same here (no Zig toolchain exists in the build image; INTEGRATION.md has the Zig shim):
//
//   IEntity / SphereEntity / QuadEntity / EntityCollection / BVHNodeEntity / Translate / RotateY   src/entity.zig
//   IMaterial (5 variants)   src/material.zig        ITexture (3 variants)   src/texture.zig     Image  src/image.zig
//   Camera / Viewport / Framebuffer   src/camera.zig     Scene / SceneType / loadScene   src/scene.zig
//   Renderer.render   src/render.zig:29   -> flatten the tree, wrt_upload_scene, wrt_render (include/wrt.h)
//
// Nothing here traces rays: the hot path lives in libwrt.so (CUDA).  The host only builds the tree exactly as the
// reference does (same arithmetic, same quirks, same BVH split rule) and hands it across the C ABI.
#pragma once

#include <cstdint>
#include <deque>
#include <memory>
#include <string>
#include <vector>

#include "../../include/wrt.h"
#include "wrh_math.hpp"
#include "wrh_rng.hpp"

namespace wrh {

// ---- image / texture / material ------------------------------------------------------------------------------
struct Image {  // image.zig:7-48 over zstbi.Image
    uint32_t width = 0, height = 0, num_components = 0, bytes_per_row = 0;
    std::vector<uint8_t> data;  // empty => the reference's "no image" state (magenta)
    static Image fromPixels(uint32_t w, uint32_t h, uint32_t comps, const uint8_t* px);
    static bool loadPnm(const std::string& path, Image& out);                   // binary P6 / P5
    // Image.initFromFile (image.zig:12-17): JPEG / PNG through the reference's vendored stb_image (wrh_image.cpp)
    static bool loadFromFile(const std::string& path, Image& out, std::string& why);
    static bool decoderAvailable();
};

enum class TextureKind : uint32_t { solid_color = WRT_TEX_SOLID, checkerboard = WRT_TEX_CHECKER, image = WRT_TEX_IMAGE };
struct ITexture {  // texture.zig:11-16
    TextureKind kind = TextureKind::solid_color;
    Color color;                     // SolidColorTexture
    Real inv_scale = 0;              // CheckerboardTexture
    const ITexture* tex_even = nullptr;
    const ITexture* tex_odd = nullptr;
    const Image* image = nullptr;    // ImageTexture
};
struct SolidColorTexture { static ITexture initTexture(Color c); };
struct CheckerboardTexture { static ITexture initTexture(Real inv_scale, const ITexture* even, const ITexture* odd); };
struct ImageTexture { static ITexture initTexture(const Image* image); };

enum class MaterialKind : uint32_t {
    lambertian = WRT_MAT_LAMBERTIAN, isotropic = WRT_MAT_ISOTROPIC, metal = WRT_MAT_METAL,
    dielectric = WRT_MAT_DIELECTRIC, diffuse_emissive = WRT_MAT_DIFFUSE_EMISSIVE
};
struct IMaterial {  // material.zig:25-32
    MaterialKind kind = MaterialKind::lambertian;
    const ITexture* texture = nullptr;
    Color albedo;
    Real fuzz = 0;
    Real refraction_index = 0;
};
struct LambertianMaterial { static IMaterial initMaterial(const ITexture* t); };
struct IsotropicMaterial { static IMaterial initMaterial(const ITexture* t); };
struct MetalMaterial { static IMaterial initMaterial(Color albedo, Real fuzz); };
struct DielectricMaterial { static IMaterial initMaterial(Real refraction_index); };
struct DiffuseLightEmissiveMaterial { static IMaterial initMaterial(const ITexture* t); };

// ---- entities --------------------------------------------------------------------------------------------------
enum class EntityKind : uint32_t {
    sphere = WRT_ENT_SPHERE, quad = WRT_ENT_QUAD, collection = WRT_ENT_COLLECTION, bvh_node = WRT_ENT_BVH_NODE,
    translate = WRT_ENT_TRANSLATE, rotate_y = WRT_ENT_ROTATE_Y
};

struct IEntity {  // entity.zig:17-24; one record type with the union's payloads side by side
    EntityKind kind = EntityKind::sphere;
    AABB aabb;
    // sphere (entity.zig:533-543)
    Point3 center;
    Real radius = 0;
    bool b_is_moving = false;
    Vec3 movement_direction;
    // quad (entity.zig:428-442)
    Point3 start_point;
    OrthoBasis basis;
    Vec3 normal;
    Real offset = 0, area = 0;
    const IMaterial* material = nullptr;  // sphere and quad
    // collection (entity.zig:306-311)
    std::vector<IEntity*> entities;
    IEntity* bvh_root = nullptr;
    // bvh_node (entity.zig:222-224)
    IEntity* left = nullptr;
    IEntity* right = nullptr;
    // translate / rotate_y (entity.zig:68-73, 112-118)
    Vec3 translate_offset;
    Real sin_theta = 0, cos_theta = 0;
    IEntity* entity = nullptr;

    const AABB& boundingBox() const { return aabb; }
};

class EntityPool {  // std.heap.MemoryPool(IEntity), main.zig:57
   public:
    IEntity* create() {
        storage_.emplace_back();
        return &storage_.back();
    }
    size_t size() const { return storage_.size(); }

   private:
    std::deque<IEntity> storage_;
};

struct SphereEntity {
    static IEntity* initEntity(EntityPool& pool, Point3 center, Real radius, const IMaterial* material);
    static IEntity* initEntityAnimated(EntityPool& pool, Point3 c0, Point3 c1, Real radius, const IMaterial* material);
};
struct QuadEntity {
    static IEntity* initEntity(EntityPool& pool, Point3 start, Vec3 axis1, Vec3 axis2, const IMaterial* material);
};
struct EntityCollection {
    static IEntity* initEntity(EntityPool& pool);
    static void add(IEntity* self, IEntity* e);
    static void createBvhTree(IEntity* self, EntityPool& pool);
};
struct BVHNodeEntity {
    static IEntity* initEntity(EntityPool& pool, std::vector<IEntity*>& entities, size_t start, size_t end);
};
struct Translate { static IEntity* initEntity(EntityPool& pool, Vec3 offset, IEntity* e); };
struct RotateY { static IEntity* initEntity(EntityPool& pool, Real angle_degrees, IEntity* e); };
IEntity* createBoxEntity(EntityPool& pool, Point3 a, Point3 b, const IMaterial* material);  // entity.zig:390-426

// ---- camera ---------------------------------------------------------------------------------------------------------
struct Framebuffer {  // camera.zig:6-40; Color = 4 lanes here (pixel stride 32 bytes)
    static constexpr size_t kLanes = 4;
    std::vector<Real> buffer;  // num_rows * num_cols * kLanes
    size_t num_rows = 0, num_cols = 0;
    static Framebuffer init(size_t height, size_t width);
    void clear(Color c);
    Real getAspectRatio() const { return static_cast<Real>(num_cols) / static_cast<Real>(num_rows); }
    size_t pixelStrideBytes() const { return kLanes * sizeof(Real); }
};

struct CoordinateBasis { Vec3 u, v, w; };

struct Viewport {  // camera.zig:105-157
    Real width = 0, height = 0;
    Point3 upper_left_corner;
    Vec3 u, v, pixel_delta_u, pixel_delta_v;
    Point3 pixel00_loc;
    static Viewport init(size_t image_width, size_t image_height, Real aspect_ratio, Real fov_vertical,
                         Real lens_focus_distance, Point3 look_from, const CoordinateBasis& basis);
};

struct Camera {  // camera.zig:48-103
    CoordinateBasis coordinate_basis;
    Point3 position;
    Real fov_vertical = 0;
    bool b_is_depth_of_field = false;
    Real lens_focus_dist = 0;
    Vec3 defocus_radius, defocus_disk_u, defocus_disk_v;
    static Camera init(Point3 look_from, Point3 look_at, Vec3 view_up, Real fov_vertical, Real lens_focus_dist,
                       Real defocus_angle_degrees);
    Viewport getViewport(const Framebuffer& fb) const;
    wrt_camera view(size_t image_width, size_t image_height) const;  // the RenderThreadContext fields the device reads
};

// ---- renderer ---------------------------------------------------------------------------------------------------------
class ThreadPool;  // wrh_writer.hpp

struct RenderStats {
    uint64_t paths = 0, rays = 0;
    double render_ms = 0, kernel_ms = 0, upload_ms = 0;
    double gather_ms = 0, kernel_ms_min = 0, kernel_ms_max = 0;
    uint32_t n_devices = 1, cull_mode_used = 0, ref_boxes_loose = 0;
};

// Owns the device side: a wrt_group over one or more CUDA devices (include/wrt.h, Multi-GPU (1)).  No CPU fallback:
// construction throws when a device is unavailable.
class Backend {
   public:
    explicit Backend(int cuda_device = 0);
    explicit Backend(const std::vector<int>& cuda_devices);
    ~Backend();
    Backend(const Backend&) = delete;
    Backend& operator=(const Backend&) = delete;
    wrt_ctx* ctx() const { return wrt_group_ctx(group_, 0); }  // the root: holds the assembled frame
    wrt_group* group() const { return group_; }
    int size() const { return wrt_group_size(group_); }

   private:
    wrt_group* group_ = nullptr;
};

struct Renderer {  // render.zig:19-27
    ThreadPool* thread_pool = nullptr;  // only the PPM writer uses host threads now
    Color clear_color;
    Color background_color;
    size_t samples_per_pixel = 10;
    size_t max_ray_bounce_depth = 20;
    const IEntity* light_entities = nullptr;
    // back end (additions)
    Backend* backend = nullptr;
    uint64_t seed = 0;
    uint32_t cull_mode = WRT_CULL_AUTO;  // the reference's result at the best speed (include/wrt.h)
    uint32_t flags = 0;                  // WRT_FLAG_* (e.g. WRT_FLAG_SHARD_SAMPLES, WRT_FLAG_SAMPLER_SOBOL)
    RenderStats last_stats;

    // render.zig:29: throws std::runtime_error carrying wrt_last_error on failure (the reference's `!void`)
    void render(const Camera& camera, const IEntity& entity, Framebuffer& framebuffer);
};

// The tree walk a Zig shim performs: IEntity / IMaterial / ITexture pointers -> the POD arrays of include/wrt.h.
struct FlatScene {
    std::vector<wrt_entity> entities;
    std::vector<uint32_t> children;
    std::vector<wrt_sphere> spheres;
    std::vector<wrt_quad> quads;
    std::vector<wrt_material> materials;
    std::vector<wrt_texture> textures;
    std::vector<wrt_image> images;
    std::vector<uint8_t> texels;
    wrt_scene view{};
    uint64_t inputBytes() const;
};
void flattenScene(const IEntity& root, const IEntity* lights, FlatScene& out);

// ---- scenes -------------------------------------------------------------------------------------------------------------
enum class SceneType { balls, shrek_quads, emissive, cornell_box, rtw_final, earth, synthetic };  // scene.zig:18-24 + harness
bool parseSceneType(const std::string& name, SceneType& out);
const char* sceneTypeName(SceneType t);
std::vector<std::string> sceneTypeNames();

struct SceneLoadContext {  // scene.zig:12-16
    EntityPool* entity_pool = nullptr;
    uint64_t seed = 1;          // replaces the getrandom-seeded `rand`
    std::string asset_dir = "assets/";
    uint32_t synthetic_prims = 1u << 20;
    // decoded images handed in by the caller (name -> pixels); looked up before asset_dir
    std::vector<std::pair<std::string, Image>> images;
};

struct Scene {  // scene.zig:36-62
    std::deque<ITexture> textures;
    std::deque<IMaterial> materials;
    std::deque<Image> images;
    IEntity* scene = nullptr;
    IEntity* lights = nullptr;
    Camera camera;
    Vec3 background_color;
    void draw(Renderer& renderer, Framebuffer& framebuffer) const;  // scene.zig:57-61
};
void loadScene(SceneType type, const SceneLoadContext& ctx, Scene& out);  // scene.zig:26-34

}  // namespace wrh
