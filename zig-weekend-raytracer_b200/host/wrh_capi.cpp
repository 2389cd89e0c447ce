// wrh_capi.cpp — plain-C handle API over the host mirror so that Python (bench.py, tests) can drive the product's own
// scene construction, flattening, Renderer.render and PPM writer without touching the checker.
#include <cstring>
#include <memory>
#include <string>

#include "wrh_scene.hpp"
#include "wrh_writer.hpp"

#define WRH_API extern "C" __attribute__((visibility("default")))

namespace {

struct SceneHandle {
    wrh::EntityPool pool;
    wrh::Scene scene;
    wrh::FlatScene flat;
    bool flattened = false;
    std::string error;
};

thread_local std::string g_error;

}  // namespace

WRH_API const char* wrh_last_error(void) { return g_error.c_str(); }

struct wrh_image_in {
    const char* name;
    uint32_t width, height, num_components;
    const uint8_t* data;
};

// loadScene (scene.zig:26-34) by name.  Returns NULL on failure (wrh_last_error).
WRH_API void* wrh_scene_load(const char* name, uint64_t seed, uint32_t synthetic_prims, const char* asset_dir,
                             const wrh_image_in* images, uint32_t n_images) {
    try {
        wrh::SceneType type;
        if (!name || !wrh::parseSceneType(name, type)) {
            g_error = std::string("unknown scene: ") + (name ? name : "(null)");
            return nullptr;
        }
        auto h = std::make_unique<SceneHandle>();
        wrh::SceneLoadContext ctx;
        ctx.entity_pool = &h->pool;
        ctx.seed = seed;
        if (synthetic_prims) ctx.synthetic_prims = synthetic_prims;
        if (asset_dir) ctx.asset_dir = asset_dir;
        for (uint32_t i = 0; i < n_images; ++i)
            ctx.images.emplace_back(images[i].name,
                                    wrh::Image::fromPixels(images[i].width, images[i].height, images[i].num_components, images[i].data));
        wrh::loadScene(type, ctx, h->scene);
        return h.release();
    } catch (const std::exception& e) {
        g_error = e.what();
        return nullptr;
    }
}

WRH_API void wrh_scene_free(void* handle) { delete static_cast<SceneHandle*>(handle); }

// The flattened view (valid until the handle is freed): what Renderer.render hands to wrt_upload_scene.
WRH_API const wrt_scene* wrh_scene_flat(void* handle) {
    auto* h = static_cast<SceneHandle*>(handle);
    if (!h) return nullptr;
    if (!h->flattened) {
        wrh::flattenScene(*h->scene.scene, h->scene.lights, h->flat);
        h->flattened = true;
    }
    return &h->flat.view;
}
WRH_API uint64_t wrh_scene_input_bytes(void* handle) {
    auto* h = static_cast<SceneHandle*>(handle);
    if (!h) return 0;
    wrh_scene_flat(handle);
    return h->flat.inputBytes();
}
WRH_API void wrh_scene_camera(void* handle, uint32_t width, uint32_t height, wrt_camera* out) {
    auto* h = static_cast<SceneHandle*>(handle);
    *out = h->scene.camera.view(width, height);
}
WRH_API void wrh_scene_background(void* handle, double out[3]) {
    auto* h = static_cast<SceneHandle*>(handle);
    out[0] = h->scene.background_color.x;
    out[1] = h->scene.background_color.y;
    out[2] = h->scene.background_color.z;
}

// Scene.draw -> Renderer.render (scene.zig:57-61, render.zig:29) on CUDA device `device` into a caller framebuffer of
// height*width pixels with 4 f64 lanes each.  Returns 0, or -1 with wrh_last_error set.
WRH_API int wrh_scene_draw(void* handle, int device, uint32_t width, uint32_t height, uint32_t samples_per_pixel,
                           uint32_t max_depth, uint64_t seed, uint32_t cull_mode, double* framebuffer, double* stats_out /*[5]*/) {
    try {
        auto* h = static_cast<SceneHandle*>(handle);
        wrh::Backend backend(device);
        wrh::Renderer renderer;
        renderer.samples_per_pixel = samples_per_pixel;
        renderer.max_ray_bounce_depth = max_depth;
        renderer.backend = &backend;
        renderer.seed = seed;
        renderer.cull_mode = cull_mode;
        wrh::Framebuffer fb = wrh::Framebuffer::init(height, width);
        h->scene.draw(renderer, fb);
        std::memcpy(framebuffer, fb.buffer.data(), fb.buffer.size() * sizeof(double));
        if (stats_out) {
            stats_out[0] = static_cast<double>(renderer.last_stats.paths);
            stats_out[1] = static_cast<double>(renderer.last_stats.rays);
            stats_out[2] = renderer.last_stats.render_ms;
            stats_out[3] = renderer.last_stats.kernel_ms;
            stats_out[4] = renderer.last_stats.upload_ms;
        }
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return -1;
    }
}

// WriterPPM.write (writer.zig:16).  Returns content bytes, or -1.
WRH_API long long wrh_write_ppm(const char* path, const double* data, uint32_t lanes, uint32_t width, uint32_t height,
                                uint32_t thread_pool_size, int truncate_to_content) {
    try {
        wrh::ThreadPool pool(thread_pool_size);
        wrh::WriterPPM writer;
        writer.thread_pool = &pool;
        writer.truncate_to_content = truncate_to_content != 0;
        return static_cast<long long>(writer.write(path, data, lanes, width, height));
    } catch (const std::exception& e) {
        g_error = e.what();
        return -1;
    }
}
WRH_API long long wrh_write_ppm_rgb8(const char* path, const uint8_t* rgb, uint32_t width, uint32_t height,
                                     uint32_t thread_pool_size, int truncate_to_content) {
    try {
        wrh::ThreadPool pool(thread_pool_size);
        wrh::WriterPPM writer;
        writer.thread_pool = &pool;
        writer.truncate_to_content = truncate_to_content != 0;
        return static_cast<long long>(writer.writeQuantised(path, rgb, width, height));
    } catch (const std::exception& e) {
        g_error = e.what();
        return -1;
    }
}

WRH_API void wrh_encode_color(const double rgb[3], uint8_t out[3]) {
    const auto px = wrh::encodeColor(rgb);
    out[0] = px[0]; out[1] = px[1]; out[2] = px[2];
}
WRH_API uint32_t wrh_size_of_line(const uint8_t px[3]) { return static_cast<uint32_t>(wrh::sizeOfLine({px[0], px[1], px[2]})); }
WRH_API uint32_t wrh_size_of_digit(uint8_t d) { return static_cast<uint32_t>(wrh::sizeOfDigit(d)); }
