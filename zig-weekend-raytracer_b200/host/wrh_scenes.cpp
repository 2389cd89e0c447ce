// wrh_scenes.cpp — the reference's scene catalogue (src/scene.zig:68-517) for the host driver, plus the two harness
// scenes the benchmark configs name (BASELINE.json configs[3] "earth", configs[4] "synthetic").
// Scene definition is not part of the accelerated path (it runs once, on the host, and stays in Zig when the shim is
// used); it is restated here because no Zig toolchain exists in the build image.
#include <cstring>
#include <stdexcept>

#include "wrh_scene.hpp"

namespace wrh {

namespace {

const char* kSceneNames[] = {"balls", "shrek_quads", "emissive", "cornell_box", "rtw_final", "earth", "synthetic"};

struct Builder {
    Scene& s;
    const SceneLoadContext& ctx;
    EntityPool& pool;

    const ITexture* tex(ITexture t) { s.textures.push_back(t); return &s.textures.back(); }
    const IMaterial* mat(IMaterial m) { s.materials.push_back(m); return &s.materials.back(); }

    // ImageTexture.initTextureFromPath (texture.zig:38-42): caller-supplied pixels, else <asset_dir>/<file> decoded like
    // zstbi does (wrh_image.cpp), else a pre-converted <asset_dir>/<stem>.ppm.  A missing / undecodable image is an error, as
    // in the reference (`try img.Image.initFromFile`): there is no stand-in.
    const ITexture* imageTexture(const std::string& file) {
        for (const auto& kv : ctx.images)
            if (kv.first == file) { s.images.push_back(kv.second); return tex(ImageTexture::initTexture(&s.images.back())); }
        Image im;
        std::string why;
        const std::string stem = file.substr(0, file.find_last_of('.'));
        if (!Image::loadFromFile(ctx.asset_dir + file, im, why) && !Image::loadPnm(ctx.asset_dir + stem + ".ppm", im))
            throw std::runtime_error("ImageInitFailed: " + why);
        s.images.push_back(std::move(im));
        return tex(ImageTexture::initTexture(&s.images.back()));
    }
};

void loadSceneBalls(Builder& b) {  // scene.zig:68-174
    Scene& s = b.s;
    Random rand(b.ctx.seed);
    const ITexture* brown = b.tex(SolidColorTexture::initTexture({0.4, 0.2, 0.1}));
    const ITexture* even = b.tex(SolidColorTexture::initTexture({0.2, 0.3, 0.1}));
    const ITexture* odd = b.tex(SolidColorTexture::initTexture({0.9, 0.9, 0.9}));
    const ITexture* ground = b.tex(CheckerboardTexture::initTexture(0.32, even, odd));

    IEntity* scene = EntityCollection::initEntity(b.pool);
    EntityCollection::add(scene, SphereEntity::initEntity(b.pool, {0, -1000, 0}, 1000, b.mat(LambertianMaterial::initMaterial(ground))));

    for (Real a = -11.0; a < 11.0; a += 1.0) {
        for (Real c = -11.0; c < 11.0; c += 1.0) {
            const Real choose_mat = rand.floatReal();
            const Real cx = a + 0.9 * rand.floatReal();
            const Real cz = c + 0.9 * rand.floatReal();
            const Point3 center{cx, 0.2, cz};
            if (length(center - Vec3{4, 0.2, 0}) > 0.9) {
                if (choose_mat < 0.8) {
                    const Color albedo = rand.sampleVec3();
                    const IMaterial* m = b.mat(LambertianMaterial::initMaterial(b.tex(SolidColorTexture::initTexture(albedo))));
                    EntityCollection::add(scene, SphereEntity::initEntity(b.pool, center, 0.2, m));
                } else if (choose_mat < 0.95) {
                    const Color albedo = rand.sampleVec3Interval({0.5, 1.0});
                    const Real fuzz = rand.floatReal() * 0.8;
                    EntityCollection::add(scene, SphereEntity::initEntity(b.pool, center, 0.2, b.mat(MetalMaterial::initMaterial(albedo, fuzz))));
                } else {
                    EntityCollection::add(scene, SphereEntity::initEntity(b.pool, center, 0.2, b.mat(DielectricMaterial::initMaterial(1.5))));
                }
            }
        }
    }
    EntityCollection::add(scene, SphereEntity::initEntity(b.pool, {0, 1, 0}, 1.0, b.mat(DielectricMaterial::initMaterial(1.5))));
    EntityCollection::add(scene, SphereEntity::initEntity(b.pool, {-4, 1, 0}, 1, b.mat(LambertianMaterial::initMaterial(brown))));
    EntityCollection::add(scene, SphereEntity::initEntity(b.pool, {4, 1, 0}, 1, b.mat(MetalMaterial::initMaterial({0.7, 0.6, 0.5}, 0.0))));
    EntityCollection::createBvhTree(scene, b.pool);

    s.scene = scene;
    s.lights = nullptr;
    s.camera = Camera::init({13, 2, 3}, {0, 0, 0}, {0, 1, 0}, 20.0, 10.0, 0.6);
    s.background_color = {0.5, 0.7, 1.0};
}

void loadSceneShrekQuads(Builder& b) {  // scene.zig:176-230
    Scene& s = b.s;
    const ITexture* image = b.imageTexture("wap.jpg");
    const IMaterial* left = b.mat(LambertianMaterial::initMaterial(image));
    const IMaterial* back = b.mat(LambertianMaterial::initMaterial(image));
    const IMaterial* right = b.mat(LambertianMaterial::initMaterial(image));
    const IMaterial* top = b.mat(LambertianMaterial::initMaterial(image));
    const IMaterial* bottom = b.mat(LambertianMaterial::initMaterial(image));
    IEntity* scene = EntityCollection::initEntity(b.pool);
    EntityCollection::add(scene, QuadEntity::initEntity(b.pool, {-3, -2, 5}, {0, 0, -4}, {0, 4, 0}, left));
    EntityCollection::add(scene, QuadEntity::initEntity(b.pool, {-2, -2, 0}, {4, 0, 0}, {0, 4, 0}, right));
    EntityCollection::add(scene, QuadEntity::initEntity(b.pool, {3, -2, 1}, {0, 0, 4}, {0, 4, 0}, back));
    EntityCollection::add(scene, QuadEntity::initEntity(b.pool, {-2, 3, 1}, {4, 0, 0}, {0, 0, 4}, top));
    EntityCollection::add(scene, QuadEntity::initEntity(b.pool, {-2, -3, 5}, {4, 0, 0}, {0, 0, -4}, bottom));
    s.scene = scene;
    s.lights = nullptr;
    s.camera = Camera::init({0, 0, 9}, {0, 0, 0}, {0, 1, 0}, 80.0, 10.0, 0.0);
    s.background_color = {0.5, 0.7, 1.0};
}

void loadSceneEmissive(Builder& b) {  // scene.zig:232-310
    Scene& s = b.s;
    const ITexture* even = b.tex(SolidColorTexture::initTexture({0.2, 0.3, 0.1}));
    const ITexture* odd = b.tex(SolidColorTexture::initTexture({0.9, 0.9, 0.9}));
    const ITexture* ground = b.tex(CheckerboardTexture::initTexture(0.32, even, odd));
    const ITexture* light_blue = b.tex(SolidColorTexture::initTexture({1, 2, 4}));
    const ITexture* light_green = b.tex(SolidColorTexture::initTexture({2.3, 4, 2.3}));
    const IMaterial* m_glass = b.mat(DielectricMaterial::initMaterial(1.5));
    const IMaterial* m_ground = b.mat(LambertianMaterial::initMaterial(ground));
    const IMaterial* m_blue = b.mat(DiffuseLightEmissiveMaterial::initMaterial(light_blue));
    const IMaterial* m_green = b.mat(DiffuseLightEmissiveMaterial::initMaterial(light_green));

    IEntity* scene = EntityCollection::initEntity(b.pool);
    EntityCollection::add(scene, SphereEntity::initEntity(b.pool, {0, -1000, 0}, 1000, m_ground));
    IEntity* glass_sphere = SphereEntity::initEntity(b.pool, {0, 2, 0}, 1.5, m_glass);
    EntityCollection::add(scene, glass_sphere);
    IEntity* light_quad = QuadEntity::initEntity(b.pool, {3, 1, -2}, {2, 0, 0}, {0, 2, 0}, m_blue);
    EntityCollection::add(scene, light_quad);
    IEntity* light_sphere = SphereEntity::initEntity(b.pool, {0, 7, 0}, 1, m_green);
    EntityCollection::add(scene, light_sphere);
    EntityCollection::createBvhTree(scene, b.pool);

    IEntity* lights = EntityCollection::initEntity(b.pool);
    EntityCollection::add(lights, light_quad);
    EntityCollection::add(lights, light_sphere);
    EntityCollection::add(lights, glass_sphere);

    s.scene = scene;
    s.lights = lights;
    s.camera = Camera::init({26, 3, 6}, {0, 2, 0}, {0, 1, 0}, 20.0, 10.0, 0.0);
    s.background_color = {0, 0, 0};
}

void loadSceneCornellBox(Builder& b) {  // scene.zig:312-408
    Scene& s = b.s;
    const ITexture* red = b.tex(SolidColorTexture::initTexture({0.65, 0.05, 0.05}));
    const ITexture* white = b.tex(SolidColorTexture::initTexture({0.73, 0.73, 0.73}));
    const ITexture* green = b.tex(SolidColorTexture::initTexture({0.12, 0.45, 0.15}));
    const ITexture* light_tex = b.tex(SolidColorTexture::initTexture({15, 15, 15}));
    const IMaterial* m_red = b.mat(LambertianMaterial::initMaterial(red));
    const IMaterial* m_white = b.mat(LambertianMaterial::initMaterial(white));
    const IMaterial* m_green = b.mat(LambertianMaterial::initMaterial(green));
    const IMaterial* m_light = b.mat(DiffuseLightEmissiveMaterial::initMaterial(light_tex));
    const IMaterial* m_glass = b.mat(DielectricMaterial::initMaterial(1.5));
    const IMaterial* m_metal = b.mat(MetalMaterial::initMaterial({0.8, 0.85, 0.88}, 0));

    IEntity* scene = EntityCollection::initEntity(b.pool);
    EntityCollection::add(scene, QuadEntity::initEntity(b.pool, {555, 0, 0}, {0, 555, 0}, {0, 0, 555}, m_green));
    EntityCollection::add(scene, QuadEntity::initEntity(b.pool, {0, 0, 0}, {0, 555, 0}, {0, 0, 555}, m_red));
    EntityCollection::add(scene, QuadEntity::initEntity(b.pool, {0, 0, 0}, {555, 0, 0}, {0, 0, 555}, m_white));
    EntityCollection::add(scene, QuadEntity::initEntity(b.pool, {555, 555, 555}, {-555, 0, 0}, {0, 0, -555}, m_white));
    EntityCollection::add(scene, QuadEntity::initEntity(b.pool, {0, 0, 555}, {555, 0, 0}, {0, 555, 0}, m_white));

    IEntity* glass_sphere = SphereEntity::initEntity(b.pool, {190, 90, 190}, 90, m_glass);
    EntityCollection::add(scene, glass_sphere);
    IEntity* box2 = Translate::initEntity(b.pool, {265, 0, 295},
                                          RotateY::initEntity(b.pool, 15.0, createBoxEntity(b.pool, {0, 0, 0}, {165, 330, 165}, m_metal)));
    EntityCollection::add(scene, box2);
    IEntity* light = QuadEntity::initEntity(b.pool, {343, 554, 332}, {-150, 0, 0}, {0, 0, -125}, m_light);
    EntityCollection::add(scene, light);
    EntityCollection::createBvhTree(scene, b.pool);

    IEntity* lights = EntityCollection::initEntity(b.pool);
    EntityCollection::add(lights, glass_sphere);
    EntityCollection::add(lights, light);

    s.scene = scene;
    s.lights = lights;
    s.camera = Camera::init({278, 278, -800}, {278, 278, 0}, {0, 1, 0}, 40.0, 10.0, 0.0);
    s.background_color = {0, 0, 0};
}

void loadSceneRTWFinal(Builder& b) {  // scene.zig:410-517
    Scene& s = b.s;
    Random rand(b.ctx.seed);
    IEntity* scene = EntityCollection::initEntity(b.pool);
    IEntity* lights = EntityCollection::initEntity(b.pool);

    const IMaterial* m_ground = b.mat(LambertianMaterial::initMaterial(b.tex(SolidColorTexture::initTexture({0.4, 0.83, 0.53}))));
    IEntity* ground_boxes = EntityCollection::initEntity(b.pool);
    EntityCollection::add(scene, ground_boxes);
    const int num_boxes_per_side = 20;
    for (int i = 0; i < num_boxes_per_side; ++i) {
        const Real fi = i;
        for (int j = 0; j < num_boxes_per_side; ++j) {
            const Real fj = j;
            const Real w = 100.0;
            const Real x0 = -1000.0 + fi * w, y0 = 0.0, z0 = -1000.0 + fj * w;
            const Real x1 = x0 + w;
            const Real y1 = rand.floatReal() * 100.0 + 1.0;
            const Real z1 = z0 + w;
            EntityCollection::add(ground_boxes, createBoxEntity(b.pool, {x0, y0, z0}, {x1, y1, z1}, m_ground));
        }
    }
    EntityCollection::createBvhTree(ground_boxes, b.pool);

    const IMaterial* m_light = b.mat(DiffuseLightEmissiveMaterial::initMaterial(b.tex(SolidColorTexture::initTexture({7, 7, 7}))));
    IEntity* light = QuadEntity::initEntity(b.pool, {123, 554, 147}, {300, 0, 0}, {0, 0, 265}, m_light);
    EntityCollection::add(scene, light);
    EntityCollection::add(lights, light);

    EntityCollection::add(scene, SphereEntity::initEntity(b.pool, {260, 150, 45}, 50.0, b.mat(DielectricMaterial::initMaterial(1.5))));
    EntityCollection::add(scene, SphereEntity::initEntity(b.pool, {0, 150, 145}, 50, b.mat(MetalMaterial::initMaterial({0.8, 0.8, 0.9}, 1.0))));
    EntityCollection::add(scene, SphereEntity::initEntity(b.pool, {360, 150, 145}, 70, b.mat(DielectricMaterial::initMaterial(1.5))));

    const ITexture* shrek = b.imageTexture("wap.jpg");
    EntityCollection::add(scene, SphereEntity::initEntity(b.pool, {400, 200, 400}, 100, b.mat(LambertianMaterial::initMaterial(shrek))));
    const ITexture* me = b.imageTexture("me.jpg");
    EntityCollection::add(scene, SphereEntity::initEntity(b.pool, {220, 280, 300}, 80, b.mat(LambertianMaterial::initMaterial(me))));

    IEntity* box_of_balls = EntityCollection::initEntity(b.pool);
    const IMaterial* m_white = b.mat(LambertianMaterial::initMaterial(b.tex(SolidColorTexture::initTexture({0.73, 0.73, 0.73}))));
    for (int i = 0; i < 1000; ++i) {
        const Point3 center = rand.sampleVec3() * Vec3::splat(165.0);
        EntityCollection::add(box_of_balls, SphereEntity::initEntity(b.pool, center, 10, m_white));
    }
    EntityCollection::createBvhTree(box_of_balls, b.pool);
    EntityCollection::add(scene, Translate::initEntity(b.pool, {-100, 270, 395}, RotateY::initEntity(b.pool, 15.0, box_of_balls)));
    EntityCollection::createBvhTree(scene, b.pool);

    s.scene = scene;
    s.lights = lights;
    s.camera = Camera::init({478, 278, -600}, {278, 278, 0}, {0, 1, 0}, 40.0, 10.0, 0.0);
    s.background_color = {0, 0, 0};
}

// BASELINE.json configs[3]: image-textured sphere (assets/earth.png) plus diffuse lights — the `emissive` layout with the
// glass sphere replaced.
void loadSceneEarth(Builder& b) {
    Scene& s = b.s;
    const ITexture* even = b.tex(SolidColorTexture::initTexture({0.2, 0.3, 0.1}));
    const ITexture* odd = b.tex(SolidColorTexture::initTexture({0.9, 0.9, 0.9}));
    const ITexture* ground = b.tex(CheckerboardTexture::initTexture(0.32, even, odd));
    const ITexture* earth = b.imageTexture("earth.png");
    const ITexture* light_a = b.tex(SolidColorTexture::initTexture({4, 4, 4}));
    const ITexture* light_b = b.tex(SolidColorTexture::initTexture({3, 2.7, 2.3}));
    const IMaterial* m_ground = b.mat(LambertianMaterial::initMaterial(ground));
    const IMaterial* m_earth = b.mat(LambertianMaterial::initMaterial(earth));
    const IMaterial* m_la = b.mat(DiffuseLightEmissiveMaterial::initMaterial(light_a));
    const IMaterial* m_lb = b.mat(DiffuseLightEmissiveMaterial::initMaterial(light_b));

    IEntity* scene = EntityCollection::initEntity(b.pool);
    EntityCollection::add(scene, SphereEntity::initEntity(b.pool, {0, -1000, 0}, 1000, m_ground));
    EntityCollection::add(scene, SphereEntity::initEntity(b.pool, {0, 2, 0}, 2.0, m_earth));
    IEntity* light_quad = QuadEntity::initEntity(b.pool, {3, 1, -2}, {2, 0, 0}, {0, 2, 0}, m_la);
    EntityCollection::add(scene, light_quad);
    IEntity* light_sphere = SphereEntity::initEntity(b.pool, {0, 7, 0}, 1, m_lb);
    EntityCollection::add(scene, light_sphere);
    EntityCollection::createBvhTree(scene, b.pool);

    IEntity* lights = EntityCollection::initEntity(b.pool);
    EntityCollection::add(lights, light_quad);
    EntityCollection::add(lights, light_sphere);

    s.scene = scene;
    s.lights = lights;
    s.camera = Camera::init({26, 3, 6}, {0, 2, 0}, {0, 1, 0}, 20.0, 10.0, 0.0);
    s.background_color = {0, 0, 0};
}

// BASELINE.json configs[4]: synthetic random sphere/quad scene: n primitives alternating sphere / quad,
// centres uniform in [-1000,1000]^3, sphere radius U[1,5], axis-aligned quads with edges U[2,10]; materials from a
// 4096-entry palette (70 % lambertian, 20 % metal with fuzz U[0,0.5], 10 % glass); 64 emissive quads (radiance 15, edges
// U[20,60]) are the light list; camera (0,0,-3000) looking at the origin, vfov 40.
void loadSceneSynthetic(Builder& b) {
    Scene& s = b.s;
    Random rand(b.ctx.seed);
    constexpr uint32_t kPalette = 4096, kLights = 64;
    std::vector<const IMaterial*> palette(kPalette);
    for (uint32_t i = 0; i < kPalette; ++i) {
        const Real choose = rand.floatReal();
        if (choose < 0.7) {
            const Color albedo = rand.sampleVec3();
            palette[i] = b.mat(LambertianMaterial::initMaterial(b.tex(SolidColorTexture::initTexture(albedo))));
        } else if (choose < 0.9) {
            const Color albedo = rand.sampleVec3Interval({0.5, 1.0});
            const Real fuzz = rand.floatReal() * 0.5;
            palette[i] = b.mat(MetalMaterial::initMaterial(albedo, fuzz));
        } else {
            palette[i] = b.mat(DielectricMaterial::initMaterial(1.5));
        }
    }
    const IMaterial* m_light = b.mat(DiffuseLightEmissiveMaterial::initMaterial(b.tex(SolidColorTexture::initTexture({15, 15, 15}))));

    IEntity* scene = EntityCollection::initEntity(b.pool);
    IEntity* lights = EntityCollection::initEntity(b.pool);
    uint32_t n_prims = b.ctx.synthetic_prims;
    if (n_prims < kLights * 2) n_prims = kLights * 2;
    const uint32_t light_every = n_prims / kLights;
    uint32_t n_lights = 0;
    scene->entities.reserve(n_prims);
    for (uint32_t i = 0; i < n_prims; ++i) {
        const Point3 c = rand.sampleVec3Interval({-1000.0, 1000.0});
        const bool is_light = (i % light_every == 1) && n_lights < kLights;
        if ((i & 1u) == 0) {
            const Real radius = rand.floatReal() * 4.0 + 1.0;
            const uint32_t pm = rand.below(kPalette);
            EntityCollection::add(scene, SphereEntity::initEntity(b.pool, c, radius, palette[pm]));
        } else {
            const Real lo = is_light ? 20.0 : 2.0, span = is_light ? 40.0 : 8.0;
            const Real l1 = rand.floatReal() * span + lo;
            const Real l2 = rand.floatReal() * span + lo;
            const uint32_t axis = rand.below(3);
            const uint32_t pm = rand.below(kPalette);
            Vec3 a1, a2;
            if (axis == 0) { a1.y = l1; a2.z = l2; }
            else if (axis == 1) { a1.z = l1; a2.x = l2; }
            else { a1.x = l1; a2.y = l2; }
            IEntity* q = QuadEntity::initEntity(b.pool, c, a1, a2, is_light ? m_light : palette[pm]);
            EntityCollection::add(scene, q);
            if (is_light) { EntityCollection::add(lights, q); ++n_lights; }
        }
    }
    EntityCollection::createBvhTree(scene, b.pool);
    s.scene = scene;
    s.lights = lights;
    s.camera = Camera::init({0, 0, -3000}, {0, 0, 0}, {0, 1, 0}, 40.0, 10.0, 0.0);
    s.background_color = {0, 0, 0};
}

}  // namespace

bool parseSceneType(const std::string& name, SceneType& out) {
    for (size_t i = 0; i < sizeof kSceneNames / sizeof *kSceneNames; ++i)
        if (name == kSceneNames[i]) { out = static_cast<SceneType>(i); return true; }
    return false;
}
const char* sceneTypeName(SceneType t) { return kSceneNames[static_cast<size_t>(t)]; }
std::vector<std::string> sceneTypeNames() { return {std::begin(kSceneNames), std::end(kSceneNames)}; }

void loadScene(SceneType type, const SceneLoadContext& ctx, Scene& out) {  // scene.zig:26-34
    if (!ctx.entity_pool) throw std::invalid_argument("SceneLoadContext.entity_pool is null");
    Builder b{out, ctx, *ctx.entity_pool};
    switch (type) {
        case SceneType::balls: loadSceneBalls(b); break;
        case SceneType::shrek_quads: loadSceneShrekQuads(b); break;
        case SceneType::emissive: loadSceneEmissive(b); break;
        case SceneType::cornell_box: loadSceneCornellBox(b); break;
        case SceneType::rtw_final: loadSceneRTWFinal(b); break;
        case SceneType::earth: loadSceneEarth(b); break;
        case SceneType::synthetic: loadSceneSynthetic(b); break;
    }
}

void Scene::draw(Renderer& renderer, Framebuffer& framebuffer) const {  // scene.zig:57-61
    renderer.background_color = background_color;
    renderer.light_entities = lights;
    renderer.render(camera, *scene, framebuffer);
}

}  // namespace wrh
