// main.cpp — `weekend-raytracer` command line driver on the B200 back end (src/main.zig:49-106 with the argument
// conventions of src/argparser.zig): same flags, same defaults, same three timing lines, same PPM output.
//
//   --image_width=<usize> (required)   --image_height=<usize> (required)   --image_out_path=image.ppm
//   --thread_pool_size=8 (PPM writer threads only; rays are traced on the GPU)   --scene=emissive
//   --samples_per_pixel=10   --ray_bounce_max_depth=20
// Additions (non-breaking): --device=0  --devices=0,1,..  --shard=rows|samples  --seed=<u64>  --cull=auto|tight|reference  --asset_dir=assets/  --synthetic_prims=N
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>

#include "wrh_scene.hpp"
#include "wrh_writer.hpp"

namespace {

enum class ParseArgsError {  // argparser.zig:7-18
    HelpPassedInArgs, ParseIntFailed, ParseEnumFailed, InvalidArgument, ArgumentMissingValue, RequiredArgumentMissing,
    UnrecognizedArgument
};
const char* errorName(ParseArgsError e) {
    switch (e) {
        case ParseArgsError::HelpPassedInArgs: return "HelpPassedInArgs";
        case ParseArgsError::ParseIntFailed: return "ParseIntFailed";
        case ParseArgsError::ParseEnumFailed: return "ParseEnumFailed";
        case ParseArgsError::InvalidArgument: return "InvalidArgument";
        case ParseArgsError::ArgumentMissingValue: return "ArgumentMissingValue";
        case ParseArgsError::RequiredArgumentMissing: return "RequiredArgumentMissing";
        case ParseArgsError::UnrecognizedArgument: return "UnrecognizedArgument";
    }
    return "?";
}

struct UserArgs {  // main.zig:20-28 (+ additions)
    size_t image_width = 0, image_height = 0;
    std::string image_out_path = "image.ppm";
    size_t thread_pool_size = 8;
    wrh::SceneType scene = wrh::SceneType::emissive;
    size_t samples_per_pixel = 10;
    size_t ray_bounce_max_depth = 20;
    int device = 0;
    std::vector<int> devices;  // --devices=0,1,2,...: one process, several GPUs (wrt_group); empty = {device}
    bool shard_samples = false; // --shard=samples: split the sample range over the devices instead of the rows
    bool sampler_sobol = false; // --sampler=sobol: Owen-scrambled Sobol dimensions for every path decision (sampler.zig:203-247)
    uint64_t seed = 1;
    uint32_t cull = WRT_CULL_AUTO;
    std::string asset_dir = "assets/";
    uint32_t synthetic_prims = 1u << 20;
    bool writer_on_device = false;  // --writer=device: wrt_format_ppm instead of the host thread pool
};

void printUsage(FILE* out) {  // argparser.zig:94-113
    std::fprintf(out, "Usage:\n");
    std::fprintf(out, "\t--image_width=<usize>\n\t--image_height=<usize>\n\t--image_out_path=<[]const u8>\n");
    std::fprintf(out, "\t--thread_pool_size=<usize>\n\t--scene=<scene.SceneType>\n");
    for (const auto& n : wrh::sceneTypeNames()) std::fprintf(out, "\t\t%s\n", n.c_str());
    std::fprintf(out, "\t--samples_per_pixel=<usize>\n\t--ray_bounce_max_depth=<usize>\n");
    std::fprintf(out, "\t--device=<i32>\n\t--devices=<i32,i32,...>\n\t--shard=<rows|samples>\n\t--sampler=<random|sobol>\n\t--seed=<u64>\n\t--cull=<auto|tight|reference>\n\t--asset_dir=<[]const u8>\n\t--synthetic_prims=<u32>\n\t--writer=<host|device>\n");
}

bool parseUnsigned(const std::string& v, unsigned long long& out) {  // std.fmt.parseInt(.., 10)
    if (v.empty()) return false;
    size_t i = 0;
    if (v[0] == '+') i = 1;
    if (i >= v.size()) return false;
    unsigned long long r = 0;
    for (; i < v.size(); ++i) {
        if (v[i] == '_') continue;  // Zig's parseInt accepts digit separators
        if (v[i] < '0' || v[i] > '9') return false;
        r = r * 10 + static_cast<unsigned>(v[i] - '0');
    }
    out = r;
    return true;
}

// cacheArgVal + parse (argparser.zig:64-136): any number of leading '-', key=value, help / h, unknown keys rejected.
UserArgs parseUserArgs(int argc, char** argv) {
    static const char* known[] = {"image_width", "image_height", "image_out_path", "thread_pool_size", "scene", "samples_per_pixel",
                                  "ray_bounce_max_depth", "device", "devices", "shard", "sampler", "seed", "cull", "asset_dir", "synthetic_prims",
                                  "writer"};
    std::map<std::string, std::string> cache;
    for (int a = 1; a < argc; ++a) {
        std::string arg = argv[a];
        size_t start = 0;
        while (start < arg.size() && arg[start] == '-') ++start;
        arg = arg.substr(start);
        const size_t eq = arg.find('=');
        const std::string key = arg.substr(0, eq);
        if (key == "help" || key == "h") throw ParseArgsError::HelpPassedInArgs;
        bool ok = false;
        for (const char* k : known) ok = ok || key == k;
        if (!ok) throw ParseArgsError::UnrecognizedArgument;
        const std::string val = eq == std::string::npos ? "" : arg.substr(eq + 1);
        if (val.empty()) throw ParseArgsError::ArgumentMissingValue;
        cache[key] = val;
    }
    UserArgs args;
    auto get_size = [&](const char* key, size_t& dst, bool required) {
        auto it = cache.find(key);
        if (it == cache.end()) {
            if (required) throw ParseArgsError::RequiredArgumentMissing;
            return;
        }
        unsigned long long v = 0;
        if (!parseUnsigned(it->second, v)) throw ParseArgsError::ParseIntFailed;
        dst = static_cast<size_t>(v);
    };
    get_size("image_width", args.image_width, true);
    get_size("image_height", args.image_height, true);
    if (cache.count("image_out_path")) args.image_out_path = cache["image_out_path"];
    get_size("thread_pool_size", args.thread_pool_size, false);
    if (cache.count("scene") && !wrh::parseSceneType(cache["scene"], args.scene)) throw ParseArgsError::ParseEnumFailed;
    get_size("samples_per_pixel", args.samples_per_pixel, false);
    get_size("ray_bounce_max_depth", args.ray_bounce_max_depth, false);
    size_t tmp = 0;
    get_size("device", tmp, false); args.device = static_cast<int>(tmp);
    tmp = 1; get_size("seed", tmp, false); args.seed = tmp;
    tmp = args.synthetic_prims; get_size("synthetic_prims", tmp, false); args.synthetic_prims = static_cast<uint32_t>(tmp);
    if (cache.count("asset_dir")) args.asset_dir = cache["asset_dir"];
    if (cache.count("writer")) {
        if (cache["writer"] == "device") args.writer_on_device = true;
        else if (cache["writer"] == "host") args.writer_on_device = false;
        else throw ParseArgsError::ParseEnumFailed;
    }
    if (cache.count("devices")) {
        std::string list = cache["devices"];
        size_t pos = 0;
        while (pos <= list.size()) {
            const size_t comma = list.find(',', pos);
            const std::string item = list.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos);
            unsigned long long v = 0;
            if (!parseUnsigned(item, v)) throw ParseArgsError::ParseIntFailed;
            args.devices.push_back(static_cast<int>(v));
            if (comma == std::string::npos) break;
            pos = comma + 1;
        }
    }
    if (cache.count("shard")) {
        if (cache["shard"] == "samples") args.shard_samples = true;
        else if (cache["shard"] == "rows") args.shard_samples = false;
        else throw ParseArgsError::ParseEnumFailed;
    }
    if (cache.count("sampler")) {
        if (cache["sampler"] == "sobol") args.sampler_sobol = true;
        else if (cache["sampler"] == "random") args.sampler_sobol = false;
        else throw ParseArgsError::ParseEnumFailed;
    }
    if (cache.count("cull")) {
        if (cache["cull"] == "tight") args.cull = WRT_CULL_TIGHT;
        else if (cache["cull"] == "auto") args.cull = WRT_CULL_AUTO;
        else if (cache["cull"] == "reference") args.cull = WRT_CULL_REFERENCE;
        else throw ParseArgsError::ParseEnumFailed;
    }
    return args;
}

struct Timer {  // timer.zig:6-42
    using clock = std::chrono::steady_clock;
    clock::time_point last = clock::now();
    void logInfoElapsed(const char* msg) {
        const auto now = clock::now();
        const long long ms = std::chrono::duration_cast<std::chrono::milliseconds>(now - last).count();
        last = now;
        std::fprintf(stderr, "info: (%-5lld ms)%s\n", ms, msg);
    }
};

}  // namespace

int main(int argc, char** argv) {
    Timer timer;
    UserArgs args;
    try {
        args = parseUserArgs(argc, argv);
    } catch (ParseArgsError e) {
        printUsage(stderr);  // usage on any parse failure, main.zig:41-45
        if (e == ParseArgsError::HelpPassedInArgs) return 0;  // main.zig:60-64
        std::fprintf(stderr, "error: %s\n", errorName(e));
        return 1;
    }
    try {
        wrh::ThreadPool thread_pool(args.thread_pool_size);
        if (args.devices.empty()) args.devices.push_back(args.device);
        wrh::Backend backend(args.devices);  // fails loudly without a CUDA device

        wrh::Renderer renderer;  // main.zig:77-83
        renderer.thread_pool = &thread_pool;
        renderer.background_color = {0, 0, 0};
        renderer.clear_color = {0, 0, 0};
        renderer.samples_per_pixel = args.samples_per_pixel;
        renderer.max_ray_bounce_depth = args.ray_bounce_max_depth;
        renderer.backend = &backend;
        renderer.seed = args.seed;
        renderer.cull_mode = args.cull;
        renderer.flags = (args.shard_samples ? WRT_FLAG_SHARD_SAMPLES : 0u) | (args.sampler_sobol ? WRT_FLAG_SAMPLER_SOBOL : 0u);

        wrh::Framebuffer framebuffer = wrh::Framebuffer::init(args.image_height, args.image_width);

        wrh::EntityPool entity_pool;
        wrh::SceneLoadContext ctx;
        ctx.entity_pool = &entity_pool;
        ctx.seed = args.seed;
        ctx.asset_dir = args.asset_dir;
        ctx.synthetic_prims = args.synthetic_prims;
        wrh::Scene scene;
        wrh::loadScene(args.scene, ctx, scene);
        timer.logInfoElapsed("scene initialized");

        scene.draw(renderer, framebuffer);
        timer.logInfoElapsed("scene rendered");
        std::fprintf(stderr, "info: %llu paths, %llu rays, kernel %.3f ms, render call %.3f ms, %.1f Mrays/s\n",
                     static_cast<unsigned long long>(renderer.last_stats.paths), static_cast<unsigned long long>(renderer.last_stats.rays),
                     renderer.last_stats.kernel_ms, renderer.last_stats.render_ms,
                     renderer.last_stats.kernel_ms > 0 ? renderer.last_stats.rays / renderer.last_stats.kernel_ms / 1e3 : 0.0);
        std::fprintf(stderr, "info: %u device(s), kernel min/max %.3f / %.3f ms, gather %.3f ms, culling %s (%u loose reference boxes)\n",
                     renderer.last_stats.n_devices, renderer.last_stats.kernel_ms_min, renderer.last_stats.kernel_ms_max,
                     renderer.last_stats.gather_ms, renderer.last_stats.cull_mode_used == WRT_CULL_REFERENCE ? "reference" : "tight",
                     renderer.last_stats.ref_boxes_loose);

        wrh::WriterPPM writer;  // main.zig:100-104
        writer.thread_pool = &thread_pool;
        if (args.writer_on_device) writer.writeOnDevice(backend.ctx(), args.image_out_path, nullptr, framebuffer.num_cols, framebuffer.num_rows);
        else writer.write(args.image_out_path, framebuffer.buffer.data(), wrh::Framebuffer::kLanes, framebuffer.num_cols, framebuffer.num_rows);
        timer.logInfoElapsed("scene written to file");
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
