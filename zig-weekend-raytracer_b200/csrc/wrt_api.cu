// wrt_api.cu — the C ABI of include/wrt.h: context, scene upload, render, gates.
// No CPU fallback exists: every entry point that needs the device returns WRT_E_CUDA when it is unavailable.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "wrt_ctx.h"

extern "C" const unsigned char wrt_sobol_blob[];

// NVTX ranges around the phases the reference marks with Tracy zones (src/render.zig:30,108,151,195; scene.zig): visible in
// Nsight Systems / Compute timelines, free when no tool is attached.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

namespace {

using wrt::DevBuf;
using wrt::SobolBlob;

thread_local std::string g_create_error;

bool parse_sobol_blob(SobolBlob& b) {
    if (std::memcmp(wrt_sobol_blob, "WRTSOBL1", 8) != 0) return false;
    uint32_t hdr[4];
    std::memcpy(hdr, wrt_sobol_blob + 8, sizeof hdr);
    b.n_dims = hdr[0]; b.matrix_size = hdr[1]; b.n_vdc = hdr[2]; b.n_vdc_inv = hdr[3];
    if (b.n_dims != 1024 || b.matrix_size != 52 || b.n_vdc != 25 || b.n_vdc_inv != 26) return false;
    b.matrices32 = reinterpret_cast<const uint32_t*>(wrt_sobol_blob + 24);
    b.vdc = reinterpret_cast<const uint64_t*>(wrt_sobol_blob + 24 + 4ull * b.n_dims * b.matrix_size);
    b.vdc_inv = b.vdc + (size_t)b.n_vdc * b.matrix_size;
    return true;
}

uint32_t ceil_pow2(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}
uint32_t log2u(uint32_t v) {
    uint32_t l = 0;
    while (v >>= 1) ++l;
    return l;
}

}  // namespace

#ifndef WRT_WF_POOL_SLOTS
#define WRT_WF_POOL_SLOTS (64ull << 20)  // path slots of the wavefront pool (128 B of state + 32 B of queue / job space each)
#endif
#ifndef WRT_WF_PIPELINES
#define WRT_WF_PIPELINES 4u  // independent pipelines the pool is cut into (wrt_kernels.h: WavefrontArgs::shared)
#endif
#define CU(call)                                                   \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return ctx->cuda_fail(e__, #call); \
    } while (0)

int wrt::bind_device(wrt_ctx* ctx) {
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return ctx->cuda_fail(e, "cudaSetDevice");
    return WRT_OK;
}
using wrt::bind_device;

extern "C" uint32_t wrt_abi_version(void) { return WRT_ABI_VERSION; }

extern "C" int wrt_check_scene(const wrt_scene* scene, wrt_scene_info* info, char* err, size_t err_cap) {
    auto report = [&](const std::string& msg) {
        if (err && err_cap) { std::snprintf(err, err_cap, "%s", msg.c_str()); }
    };
    if (err && err_cap) err[0] = 0;
    if (!scene || !info) { report("scene / info is NULL"); return WRT_E_INVALID; }
    wrt::CompiledScene cs;
    std::string msg;
    const int rc = wrt::compile_scene(scene, cs, msg);
    if (rc != WRT_OK) { report(msg); return rc; }
    std::memset(info, 0, sizeof *info);
    info->n_ops = (uint32_t)cs.ops.size();
    info->n_ops_packet = (uint32_t)(cs.ops_pruned.empty() ? cs.ops.size() : cs.ops_pruned.size());
    info->n_prims = cs.n_prims;
    info->n_boxes = (uint32_t)cs.boxes_tight.size();
    info->n_tree_records = (uint32_t)(cs.use_wide ? cs.nodes4.size() : cs.nodes2.size());
    info->max_nesting = cs.max_nesting;
    info->n_lights = (uint32_t)cs.lights.size();
    info->ref_boxes_loose = cs.ref_boxes_loose;
    info->stack_depth = cs.stack_depth;
    info->compact_stack = cs.compact_ok ? 1u : 0u;
    info->quantised_records = (uint32_t)cs.nodes4q.size();
    if (!wrt::check_compiled_scene(cs, info->tree_depth, msg)) { report(msg); return WRT_E_STATE; }
    return WRT_OK;
}

extern "C" int wrt_build_trees(const wrt_scene* scene, int cuda_device, wrt_tree_info* info, void* records2, size_t cap2, void* records4,
                               size_t cap4, char* err, size_t err_cap) {
    auto report = [&](const std::string& msg) {
        if (err && err_cap) { std::snprintf(err, err_cap, "%s", msg.c_str()); }
    };
    if (err && err_cap) err[0] = 0;
    if (!scene || !info) { report("scene / info is NULL"); return WRT_E_INVALID; }
    try {
        const auto t0 = std::chrono::steady_clock::now();
        wrt::CompiledScene cs;
        std::string msg;
        int rc = wrt::compile_scene(scene, cs, msg, /*defer_trees=*/true);
        if (rc != WRT_OK) { report(msg); return rc; }
        std::memset(info, 0, sizeof *info);
        if (cuda_device >= 0) {
            int count = 0;
            if (cudaGetDeviceCount(&count) != cudaSuccess || cuda_device >= count) {
                report("wrt_build_trees: no such CUDA device (the device build has no CPU fallback; pass cuda_device < 0 for the host build)");
                return WRT_E_CUDA;
            }
            NvtxRange range("wrt_build_trees: device");
            rc = wrt::build_trees_device(cs, cuda_device, msg, &info->build_ms);
            if (rc != WRT_OK) { report(msg); return rc; }
            info->on_device = 1;
        } else {
            NvtxRange range("wrt_build_trees: host");
            const auto t1 = std::chrono::steady_clock::now();
            wrt::build_trees_host(cs);
            info->build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
        }
        info->n_records2 = (uint32_t)cs.nodes2.size();
        info->n_records4 = (uint32_t)cs.nodes4.size();
        info->stack_depth = cs.stack_depth;
        info->use_wide = cs.use_wide ? 1u : 0u;
        info->max_nesting = cs.max_nesting;
        if (records2 && cap2) std::memcpy(records2, cs.nodes2.data(), std::min(cap2, cs.nodes2.size() * sizeof(wrt::Node2)));
        if (records4 && cap4) std::memcpy(records4, cs.nodes4.data(), std::min(cap4, cs.nodes4.size() * sizeof(wrt::Node4)));
        info->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return WRT_OK;
    } catch (const std::bad_alloc&) {
        report("wrt_build_trees: out of host memory");
        return WRT_E_NOMEM;
    } catch (const std::exception& e) {
        report(std::string("wrt_build_trees: ") + e.what());
        return WRT_E_INVALID;
    }
}

extern "C" const char* wrt_last_error(const wrt_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int wrt_create(int cuda_device, wrt_ctx** out) {
    if (!out) { g_create_error = "wrt_create: out is NULL"; return WRT_E_INVALID; }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("wrt_create: no CUDA device (") + cudaGetErrorString(e) + "); this back end has no CPU fallback";
        return WRT_E_CUDA;
    }
    if (cuda_device < 0 || cuda_device >= count) { g_create_error = "wrt_create: device index out of range"; return WRT_E_INVALID; }
    wrt_ctx* ctx = new (std::nothrow) wrt_ctx();
    if (!ctx) { g_create_error = "wrt_create: out of host memory"; return WRT_E_NOMEM; }
    ctx->device = cuda_device;
    auto bail = [&](cudaError_t ce, const char* what) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(ce);
        wrt_destroy(ctx);
        return WRT_E_CUDA;
    };
    if ((e = cudaSetDevice(cuda_device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, cuda_device)) != cudaSuccess) return bail(e, "cudaGetDeviceProperties");
    ctx->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) {
        g_create_error = "wrt_create: kernels are built for sm_100a only; device is sm_" + std::to_string(prop.major * 10 + prop.minor);
        wrt_destroy(ctx);
        return WRT_E_CUDA;
    }
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    for (auto& ev : ctx->ev)
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if (!parse_sobol_blob(ctx->blob)) {
        g_create_error = "wrt_create: embedded Sobol table blob is corrupt";
        wrt_destroy(ctx);
        return WRT_E_INVALID;
    }
    if ((e = ctx->d_counters.ensure(4)) != cudaSuccess) return bail(e, "cudaMalloc(counters)");
    *out = ctx;
    return WRT_OK;
}

extern "C" void wrt_destroy(wrt_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    wrt::comm_release(ctx);
    ctx->d_shard.release(); ctx->d_staging.release();
    ctx->free_images();
    ctx->d_ops.release(); ctx->d_ops_pruned.release(); ctx->d_boxes_ref.release(); ctx->d_boxes_tight.release(); ctx->d_nodes2.release(); ctx->d_nodes4.release(); ctx->d_root4.release(); ctx->d_sphere_pc.release(); ctx->d_quad_pc.release(); ctx->d_nodes4q.release(); ctx->d_spheres.release();
    ctx->d_sphere_aux.release(); ctx->d_quads.release(); ctx->d_xforms.release(); ctx->d_xform_chains.release(); ctx->d_materials.release();
    ctx->d_textures.release(); ctx->d_images.release(); ctx->d_lights.release(); ctx->d_light_boxes.release(); ctx->d_sobol_matrices.release(); ctx->d_sobol_lut.release();
    ctx->d_accum.release(); ctx->d_fb.release(); ctx->d_rgb8.release(); ctx->d_counters.release();
    ctx->d_ppm_in.release(); ctx->d_ppm_body.release(); ctx->d_ppm_blocks.release(); ctx->d_ppm_offsets.release();
    ctx->d_wf_keys.release(); ctx->d_wf_sort_tmp.release(); ctx->d_wf_sort_hist.release();
    ctx->d_wf_paths.release(); ctx->d_wf_queues.release(); ctx->d_wf_slot_job.release(); ctx->d_wf_counters.release();
    if (ctx->h_wf_counters) cudaFreeHost(ctx->h_wf_counters);
    for (int k = 1; k < WRT_WF_MAX_PIPELINES; ++k) {
        if (ctx->wf_streams[k]) cudaStreamDestroy(ctx->wf_streams[k]);
        if (ctx->wf_events[k]) cudaEventDestroy(ctx->wf_events[k]);
    }
    for (auto& ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// Image textures as CUDA texture objects: uchar4 texels, point sampling, unnormalised coordinates, clamped —
// Image.getPixel's clamped fetch (image.zig:23-36).  pixelToColor reads bytes 0..2 of a num_components-wide
// pixel (texture.zig:70-77), which for 1- and 2-component images runs into the next pixel;
// the RGBA staging copy reproduces exactly those bytes.
static int upload_images(wrt_ctx* ctx, const wrt_scene* sc) {
    ctx->free_images();
    std::vector<wrt::ImageDesc> descs(sc->n_images);
    for (uint32_t i = 0; i < sc->n_images; ++i) {
        const wrt_image& im = sc->images[i];
        wrt::ImageDesc d;
        d.tex = 0; d.width = im.width; d.height = im.height;
        if (im.height == 0 || im.width == 0) { d.height = 0; descs[i] = d; continue; }
        const uint64_t need = im.texel_offset + (uint64_t)im.bytes_per_row * im.height;
        if (!sc->texels || need > sc->texel_bytes || im.num_components == 0 || (uint64_t)im.width * im.num_components > im.bytes_per_row)
            return ctx->fail(WRT_E_INVALID, "image texel range out of bounds");
        std::vector<uchar4> rgba;
        try {
            rgba.resize((size_t)im.width * im.height);
        } catch (const std::bad_alloc&) {
            return ctx->fail(WRT_E_NOMEM, "out of host memory staging an image texture");
        }
        const uint8_t* base = sc->texels + im.texel_offset;
        const uint64_t total = (uint64_t)im.bytes_per_row * im.height;
        for (uint32_t y = 0; y < im.height; ++y)
            for (uint32_t x = 0; x < im.width; ++x) {
                uint64_t o = (uint64_t)im.bytes_per_row * y + (uint64_t)im.num_components * x;
                uchar4 px;
                px.x = base[o];
                px.y = (o + 1 < total) ? base[o + 1] : 0;
                px.z = (o + 2 < total) ? base[o + 2] : 0;
                px.w = 255;
                rgba[(size_t)y * im.width + x] = px;
            }
        cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uchar4>();
        cudaArray_t arr = nullptr;
        CU(cudaMallocArray(&arr, &fmt, im.width, im.height));
        ctx->arrays.push_back(arr);
        CU(cudaMemcpy2DToArray(arr, 0, 0, rgba.data(), (size_t)im.width * sizeof(uchar4), (size_t)im.width * sizeof(uchar4), im.height,
                               cudaMemcpyHostToDevice));
        cudaResourceDesc res;
        std::memset(&res, 0, sizeof res);
        res.resType = cudaResourceTypeArray;
        res.res.array.array = arr;
        cudaTextureDesc td;
        std::memset(&td, 0, sizeof td);
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 0;
        cudaTextureObject_t tex = 0;
        CU(cudaCreateTextureObject(&tex, &res, &td, nullptr));
        ctx->texobjs.push_back(tex);
        d.tex = tex;
        descs[i] = d;
    }
    CU(ctx->d_images.upload(descs, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));  // descs is a local
    return WRT_OK;
}

int wrt::upload_compiled(wrt_ctx* ctx, const wrt::CompiledScene& cs, const wrt_scene* scene, double compile_ms) {
    NvtxRange range("wrt_upload_scene: H2D");
    auto t0 = std::chrono::steady_clock::now();
    int rc = bind_device(ctx);
    if (rc) return rc;
    ctx->have_scene = false;
    ctx->last_valid = false;
    const bool trace = std::getenv("WRT_TRACE_BUILD") != nullptr;
    auto t_last = t0;
    auto lap = [&](const char* what) {
        if (!trace) return;
        const auto t = std::chrono::steady_clock::now();
        std::fprintf(stderr, "wrt trace: upload: %s %.2f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
        t_last = t;
    };
    rc = upload_images(ctx, scene);
    if (rc != WRT_OK) return rc;
    lap("images");
    CU(ctx->d_ops.upload(cs.ops, ctx->stream));
    if (!cs.ops_pruned.empty()) CU(ctx->d_ops_pruned.upload(cs.ops_pruned, ctx->stream));
    CU(ctx->d_boxes_ref.upload(cs.boxes_ref, ctx->stream));
    CU(ctx->d_boxes_tight.upload(cs.boxes_tight, ctx->stream));
    CU(ctx->d_nodes2.upload(cs.nodes2, ctx->stream));
    CU(ctx->d_nodes4.upload(cs.nodes4, ctx->stream));
    CU(ctx->d_root4.upload(cs.root4, ctx->stream));
    CU(ctx->d_sphere_pc.upload(cs.sphere_pc, ctx->stream));
    CU(ctx->d_quad_pc.upload(cs.quad_pc, ctx->stream));
    CU(ctx->d_nodes4q.upload(cs.nodes4q, ctx->stream));
    CU(ctx->d_spheres.upload(cs.spheres, ctx->stream));
    CU(ctx->d_sphere_aux.upload(cs.sphere_aux, ctx->stream));
    CU(ctx->d_quads.upload(cs.quads, ctx->stream));
    CU(ctx->d_xforms.upload(cs.xforms, ctx->stream));
    CU(ctx->d_xform_chains.upload(cs.xform_chains, ctx->stream));
    CU(ctx->d_materials.upload(cs.materials, ctx->stream));
    CU(ctx->d_textures.upload(cs.textures, ctx->stream));
    CU(ctx->d_lights.upload(cs.lights, ctx->stream));
    CU(ctx->d_light_boxes.upload(cs.light_boxes, ctx->stream));
    lap("buffers sized, copies queued");
    CU(cudaStreamSynchronize(ctx->stream));
    lap("copies done");
    wrt::DeviceScene& ds = ctx->ds;
    ds.ops = ctx->d_ops.p; ds.boxes_ref = ctx->d_boxes_ref.p; ds.boxes_tight = ctx->d_boxes_tight.p; ds.nodes2 = ctx->d_nodes2.p; ds.nodes4 = ctx->d_nodes4.p; ds.root4 = ctx->d_root4.p;
    ds.spheres = ctx->d_spheres.p; ds.sphere_aux = ctx->d_sphere_aux.p; ds.quads = ctx->d_quads.p;
    ds.xforms = ctx->d_xforms.p; ds.xform_chains = ctx->d_xform_chains.p; ds.materials = ctx->d_materials.p; ds.textures = ctx->d_textures.p;
    ds.images = ctx->d_images.p; ds.lights = ctx->d_lights.p; ds.light_boxes = ctx->d_light_boxes.p;
    ds.n_ops = (uint32_t)cs.ops.size();
    ds.n_lights = (uint32_t)cs.lights.size();
    ds.has_lights = cs.has_lights ? 1u : 0u;
    ds.has_moving = cs.has_moving ? 1u : 0u;
    // ordered traversal only when its exact worst-case stack use fits (ordered_stack_depth walks the rebuilt trees)
    ds.use_ordered = (cs.stack_depth <= WRT_STACK_DEPTH) ? 1u : 0u;
    ds.use_wide = cs.use_wide ? 1u : 0u;
    ds.sphere_pc = ctx->d_sphere_pc.p; ds.quad_pc = ctx->d_quad_pc.p;
    {
        const char* env = std::getenv("WRT_COMPACT_STACK");  // A/B switch: 0 keeps the 16-byte entries
        ds.compact_ok = (cs.compact_ok && !(env && env[0] == '0')) ? 1u : 0u;
        const char* envq = std::getenv("WRT_QUANT_RECORDS");  // A/B switch: 0 keeps the binary32 records under the compact stack
        ds.nodes4q = (ds.compact_ok && !cs.nodes4q.empty() && !(envq && envq[0] == '0')) ? ctx->d_nodes4q.p : nullptr;
    }
    ctx->ds_pruned = ds;
    if (!cs.ops_pruned.empty()) { ctx->ds_pruned.ops = ctx->d_ops_pruned.p; ctx->ds_pruned.n_ops = (uint32_t)cs.ops_pruned.size(); }
    // Large trees: ask L2 to keep the four-wide records (the dependent fetches of every traversal step) in preference to the
    // streaming data (path pool, queues, primitives): an access-policy window over nodes4 on this context's stream.
    {
        cudaStreamAttrValue attr;
        std::memset(&attr, 0, sizeof attr);
        const char* env = std::getenv("WRT_L2_PERSIST");
        const bool want = cs.use_wide && env && env[0] == '1';  // opt-in: measured -5 % on the 2^20-primitive scene (the set-aside
                                                                 // starves the primitive records), kept as an experiment switch
        if (want) {
            int max_persist = 0, max_window = 0;
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
            cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
            const size_t bytes = cs.nodes4.size() * sizeof(wrt::Node4);
            const size_t window = std::min<size_t>(bytes, (size_t)std::max(max_window, 0));
            if (max_persist > 0 && window > 0) {
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist);
                attr.accessPolicyWindow.base_ptr = ctx->d_nodes4.p;
                attr.accessPolicyWindow.num_bytes = window;
                attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)max_persist / (double)window);
                attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            }
        }
        cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr);  // zero-sized window = none
        cudaGetLastError();  // best effort: a device without the feature just runs without the hint
    }
    ctx->n_ops = cs.ops.size();
    ctx->has_moving = cs.has_moving;
    ctx->ref_boxes_loose = cs.ref_boxes_loose;
    ctx->have_scene = true;
    ctx->stats.program_ops = ds.n_ops;
    ctx->stats.n_prims = cs.n_prims;
    ctx->stats.ref_boxes_loose = cs.ref_boxes_loose;
    ctx->stats.upload_ms = compile_ms + std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return WRT_OK;
}

extern "C" int wrt_upload_scene(wrt_ctx* ctx, const wrt_scene* scene) {
    if (!ctx) return WRT_E_INVALID;
    auto t0 = std::chrono::steady_clock::now();
    ctx->have_scene = false;
    ctx->last_valid = false;
    std::string err;
    wrt::CompiledScene cs;  // host arrays live only for the duration of the upload
    int rc;
    double tree_ms = 0.0;
    bool tree_on_device = false;
    {
        NvtxRange range("wrt_upload_scene: compile + tree build");
        rc = wrt::compile_scene_for_device(scene, cs, err, ctx->device, &tree_ms, &tree_on_device);
    }
    if (rc != WRT_OK) return ctx->fail(rc, "wrt_upload_scene: " + err);
    const double compile_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    rc = wrt::upload_compiled(ctx, cs, scene, compile_ms);
    ctx->stats.tree_build_ms = tree_ms;
    ctx->stats.tree_build_device = tree_on_device ? 1u : 0u;
    ctx->stats.n_tree_records = (uint32_t)(cs.use_wide ? cs.nodes4.size() : cs.nodes2.size());
    return rc;
}

// Build this context's Sobol rows (LaunchParams::sobol) for a W x H framebuffer (scale = ceilPowerOfTwo(max(W,H)), sampler.zig:188).
static int prepare_sobol(wrt_ctx* ctx, uint32_t width, uint32_t height) {
    if (ctx->sobol_w == width && ctx->sobol_h == height) return WRT_OK;
    if (width == 0 || height == 0) return ctx->fail(WRT_E_INVALID, "image dimensions must be non-zero");
    const uint32_t mx = width > height ? width : height;
    // VdCSobolMatrices has 25 rows, VdCSobolMatricesInv 26 (sobolmatrices.zig): resolutions up to 2^25 are addressable
    if (mx > (1u << 25)) return ctx->fail(WRT_E_LIMIT, "image side exceeds the Sobol table range (2^25)");
    ctx->sobol_w = ctx->sobol_h = 0;
    wrt::SobolTables& t = ctx->lp.sobol;
    std::memset(&t, 0, sizeof t);
    t.scale = ceil_pow2(mx);
    t.log2_scale = log2u(t.scale);
    if (t.log2_scale > 0) {
        std::memcpy(t.vdc, ctx->blob.vdc + (size_t)(t.log2_scale - 1) * 52, sizeof t.vdc);
        std::memcpy(t.vdc_inv, ctx->blob.vdc_inv + (size_t)(t.log2_scale - 1) * 52, sizeof t.vdc_inv);
    }
    std::memcpy(t.dim0, ctx->blob.matrices32, sizeof t.dim0);
    std::memcpy(t.dim1, ctx->blob.matrices32 + 52, sizeof t.dim1);
    for (int i = 0; i < 52; ++i)  // the device evaluates dimension 0 as a bit reversal
        if (t.dim0[i] != (i < 32 ? (0x80000000u >> i) : 0u)) return ctx->fail(WRT_E_INVALID, "Sobol dimension 0 is not van der Corput");
    // byte-indexed folds of the three matrices (wrt_device.cuh: SobolLut)
    std::vector<wrt::SobolLut> lut(1);
    std::memset(lut.data(), 0, sizeof(wrt::SobolLut));
    for (int k = 0; k < 7; ++k)
        for (int v = 0; v < 256; ++v) {
            uint64_t inv = 0, vd = 0;
            uint32_t d1 = 0;
            for (int j = 0; j < 8; ++j) {
                const int col = 8 * k + j;
                if (!((v >> j) & 1) || col >= 52) continue;
                inv ^= t.vdc_inv[col];
                vd ^= t.vdc[col];
                d1 ^= t.dim1[col];
            }
            lut[0].vdc_inv[k][v] = inv;
            lut[0].dim1[k][v] = d1;
            if (k < 2) lut[0].vdc[k][v] = vd;
        }
    // b = (px << m | py) ^ delta: delta is an XOR of VdC rows; find how many bytes can be non-zero
    uint64_t b_mask = ((uint64_t)1 << (2 * t.log2_scale)) - 1;
    for (int c = 0; c < 52; ++c) b_mask |= t.vdc[c];
    t.b_bytes = 0;
    while (t.b_bytes < 8 && (b_mask >> (8 * t.b_bytes)) != 0) ++t.b_bytes;
    if (t.b_bytes > 7) return ctx->fail(WRT_E_LIMIT, "Sobol index exceeds 56 bits");
    // increments of the dims 0/1 sample bits for s -> s + 1 (wrt_device.cuh: sobol_pixel_bits_next): the bits of the index
    // of "sample" 2^(k+1) - 1 in pixel (0, 0) — everything below is linear over GF(2), so the pixel term cancels
    for (int k = 0; k < 32; ++k) {
        const uint64_t q = (k == 31) ? 0xFFFFFFFFull : ((1ull << (k + 1)) - 1);
        uint64_t index = q;
        if (t.log2_scale > 0) {
            index = q << (2 * t.log2_scale);
            uint64_t delta = 0;
            for (int c = 0; c < 52; ++c) if ((q >> c) & 1) delta ^= t.vdc[c];
            for (int c = 0; c < 52; ++c) if ((delta >> c) & 1) index ^= t.vdc_inv[c];
        }
        uint32_t v0 = 0, v1 = 0;
        for (int c = 0; c < 52; ++c) if ((index >> c) & 1) { v0 ^= t.dim0[c]; v1 ^= t.dim1[c]; }
        t.inc0[k] = v0; t.inc1[k] = v1;
    }
    CU(ctx->d_sobol_lut.upload(lut, ctx->stream));  // per-context device memory; ordered on this context's stream
    t.lut = ctx->d_sobol_lut.p;
    CU(cudaStreamSynchronize(ctx->stream));  // `lut` is a local
    ctx->sobol_w = width; ctx->sobol_h = height;
    return WRT_OK;
}

// Small programs are scanned with the warp-uniform packet traversal, large ones per lane.
static bool use_packet(const wrt_ctx* ctx, uint32_t flags) {
    if (flags & WRT_FLAG_FORCE_LANE) return false;
    if (flags & WRT_FLAG_FORCE_PACKET) return true;
    return ctx->n_ops <= WRT_PACKET_MAX_OPS;
}

// The scene view a launch scans: the packet traversal under WRT_CULL_TIGHT reads the pruned program (wrt_program.cu,
// prune_program); everything else — reference culling, per-lane and ordered traversal (Node2 records index `ops`) — the full one.
static const wrt::DeviceScene& scene_view(const wrt_ctx* ctx, uint32_t cull_mode, bool packet) {
    return (packet && cull_mode == WRT_CULL_TIGHT) ? ctx->ds_pruned : ctx->ds;
}

static uint32_t shard_rows(const wrt_params& p) {
    if (p.row_shard_index >= p.height) return 0;
    return (p.height - p.row_shard_index + p.row_shard_count - 1) / p.row_shard_count;
}

static int normalise_params(wrt_ctx* ctx, const wrt_params* in, wrt_params& p) {
    if (!in) return ctx->fail(WRT_E_INVALID, "params is NULL");
    p = *in;
    if (p.row_shard_count == 0) p.row_shard_count = 1;
    if (p.sample_begin == 0 && p.sample_end == 0) p.sample_end = p.samples_per_pixel;
    if (p.width == 0 || p.height == 0) return ctx->fail(WRT_E_INVALID, "image dimensions must be non-zero");
    if ((uint64_t)p.width * p.height > 0xFFFFFFFFull) return ctx->fail(WRT_E_LIMIT, "more than 2^32 pixels");
    if (p.samples_per_pixel == 0) return ctx->fail(WRT_E_INVALID, "samples_per_pixel must be non-zero");
    if (p.sample_begin > p.sample_end) return ctx->fail(WRT_E_INVALID, "sample_begin > sample_end");
    if (p.sample_end > p.samples_per_pixel) return ctx->fail(WRT_E_INVALID, "sample_end > samples_per_pixel");
    if (p.row_shard_index >= p.row_shard_count) return ctx->fail(WRT_E_INVALID, "row_shard_index >= row_shard_count");
    if (p.cull_mode > WRT_CULL_TIGHT) return ctx->fail(WRT_E_INVALID, "unknown cull_mode");
    p.cull_mode = (uint32_t)ctx->resolve_cull(p.cull_mode);
    ctx->stats.cull_mode_used = p.cull_mode;
    return WRT_OK;
}

static void fill_constants(const wrt_camera& cam, const wrt_params& p, wrt::RenderConstants& rc) {
    std::memset(&rc, 0, sizeof rc);
    rc.cam = cam;
    for (int k = 0; k < 3; ++k) rc.background[k] = p.background_color[k];
    rc.seed = p.seed;
    rc.width = p.width; rc.height = p.height; rc.spp = p.samples_per_pixel; rc.max_depth = p.max_ray_bounce_depth;
    rc.dof = (cam.is_depth_of_field && !(p.flags & WRT_FLAG_DISABLE_DOF)) ? 1u : 0u;
    rc.row_shard_index = p.row_shard_index; rc.row_shard_count = p.row_shard_count;
    rc.n_rows_local = shard_rows(p);
    rc.n_col_blocks = (p.width + 31u) / 32u;
    rc.sample_begin = p.sample_begin; rc.sample_end = p.sample_end;
}

// Shared body of wrt_render / wrt_render_device.  `d_out` is a device pointer or NULL (internal buffer + D2H).
int wrt::render_impl(wrt_ctx* ctx, const wrt_camera* cam, const wrt_params* params, void* host_fb, void* d_out, size_t stride) {
    if (!ctx) return WRT_E_INVALID;
    NvtxRange range("wrt_render");
    int rc_ = bind_device(ctx);
    if (rc_) return rc_;
    if (!ctx->have_scene) return ctx->fail(WRT_E_STATE, "wrt_render: no scene uploaded");
    if (!cam) return ctx->fail(WRT_E_INVALID, "camera is NULL");
    if (!host_fb && !d_out) return ctx->fail(WRT_E_INVALID, "framebuffer is NULL");
    if (stride < 24 || stride % 8) return ctx->fail(WRT_E_INVALID, "pixel_stride_bytes must be a multiple of 8 and >= 24");
    wrt_params p;
    rc_ = normalise_params(ctx, params, p);
    if (rc_) return rc_;
    rc_ = prepare_sobol(ctx, p.width, p.height);
    if (rc_) return rc_;

    wrt::RenderConstants& rc = ctx->lp.rc;  // this launch's constants (kernel argument)
    fill_constants(*cam, p, rc);
    const wrt::LaunchParams& lp = ctx->lp;
    const uint32_t stride_d = (uint32_t)(stride / 8);
    const uint64_t n_pixels64 = (uint64_t)rc.n_rows_local * p.width;
    const uint32_t n_pixels = (uint32_t)n_pixels64;
    const uint32_t n_samples = p.sample_end - p.sample_begin;

    // Job decomposition: the reference's (row x column block) jobs (render.zig:55-73) times a sample split; render_kernel
    // hands the same work out per lane (pixel x sample chunk).
    int blocks_per_sm = 0;
    const bool packet = use_packet(ctx, p.flags);
    const wrt::DeviceScene& view = scene_view(ctx, p.cull_mode, packet);
    CU(wrt::render_occupancy(view, p.cull_mode, packet, &blocks_per_sm));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    const bool sync_engine = (p.flags & WRT_FLAG_ENGINE_SYNC) != 0;
    // regrouping kernel: packet programs without moving spheres (its staging area carries no ray time)
    const bool regroup_engine = !sync_engine && (p.flags & WRT_FLAG_ENGINE_REGROUP) && packet && !ctx->has_moving;
    const bool block_per_sm = sync_engine || regroup_engine;
    const uint32_t grid = block_per_sm ? (uint32_t)ctx->sm_count : (uint32_t)ctx->sm_count * (uint32_t)blocks_per_sm;
    const uint64_t base_jobs = (uint64_t)rc.n_rows_local * rc.n_col_blocks;
    // Engine: the persistent megakernel is the default — on the measured configs it matches the
    // wavefront (queues in HBM, one small kernel per stage) without its state traffic; the wavefront is selected by flag.
    // Large trees (the four-wide records, i.e. scenes whose records do not stay in L1) default to the wavefront: its persistent
    // extend kernel replaces every finished ray at once, which the heavy-tailed traversal lengths of such scenes need (the
    // megakernel's node loop runs with 5 of 32 lanes there).  Everything else defaults to the megakernel.
    bool wavefront = !packet && ctx->ds.use_wide && ctx->ds.use_ordered && p.cull_mode == WRT_CULL_TIGHT &&
                     p.max_ray_bounce_depth > 0 && n_pixels64 > 0 && n_samples > 0 &&
                     !(p.flags & (WRT_FLAG_ENGINE_SYNC | WRT_FLAG_ENGINE_REGROUP));
    if (p.flags & WRT_FLAG_ENGINE_MEGAKERNEL) wavefront = false;
    // the Sobol-dimension sampler lives in the wavefront kernels only (the megakernels sit at their register caps)
    const bool sobol_sampler = (p.flags & WRT_FLAG_SAMPLER_SOBOL) != 0;
    if ((p.flags & (WRT_FLAG_ENGINE_WAVEFRONT | WRT_FLAG_SAMPLER_SOBOL)) && p.max_ray_bounce_depth > 0 && n_pixels64 > 0 && n_samples > 0) wavefront = true;
    if (sobol_sampler && !wavefront) return ctx->fail(WRT_E_INVALID, "WRT_FLAG_SAMPLER_SOBOL needs depth > 0 and a non-empty frame / sample range");
    // Sample chunks (include/wrt.h, WRT_FLAG_CHUNKS): a function of the FULL frame, the sample count and the engine only —
    // not of the shard or the grid — so the per-pixel summation tree, and with it every bit of the frame, is the same on
    // 1 or 8 GPUs.  The accumulators (chunks x shard pixels x 24 B) stay under 0.8 GB for frames of up to 2^25 pixels
    // (6.4 GB for the wavefront, whose jobs are smaller).
    uint32_t n_chunks = 1;
    const uint64_t frame_pixels = (uint64_t)p.width * p.height;
    const uint32_t forced_chunks = p.flags >> 24;
    if (n_samples > 0) {
        uint64_t want, min_chunk;
        if (wavefront) {
            // wavefront: jobs are handed to a pool of WRT_WF_POOL_SLOTS path slots as slots fall free, so they are made small
            // (2^28 (pixel, chunk) jobs per frame, >= 8 samples each): the pool then stays full until the last few hundred
            // iterations of a frame, on one GPU or on eight.  Measured on the 4K frame of 2^20 primitives at 1 024 spp: 17 chunks
            // (2^27 jobs) against 32 (2^28): 859 -> 865 Mrays/s on one GPU, 6 082 -> 6 501 on eight — there a device's shard held
            // fewer jobs than the pool has slots and the pool thinned out towards the end of the frame.
            want = ((1ull << 28) + frame_pixels - 1) / frame_pixels;
            min_chunk = 8;
        } else {
            // 2^25 (pixel, chunk) jobs per frame: a few dozen per resident lane even when 8 GPUs share the frame.  The packet
            // kernel takes jobs per lane and is happy with chunks of 4 samples; the kernels that take jobs per warp wait for
            // the longest lane of each job, which only averages out over chunks of a few hundred samples.
            want = ((1ull << 25) + frame_pixels - 1) / frame_pixels;
            const bool lane_jobs = packet && !sync_engine && !regroup_engine;
            min_chunk = lane_jobs ? 4 : 256;
        }
        want = std::min<uint64_t>(want, 64);
        want = std::min<uint64_t>(want, (n_samples + min_chunk - 1) / min_chunk);
        if (forced_chunks) want = std::min<uint64_t>(forced_chunks, n_samples);
        n_chunks = (uint32_t)std::max<uint64_t>(want, 1);
    }
    if (n_pixels64 > 0 && (uint64_t)n_chunks * n_pixels64 > 0xFFFFFFFFull)  // lane jobs are indexed in 32 bits
        n_chunks = (uint32_t)std::max<uint64_t>(0xFFFFFFFFull / n_pixels64, 1);
    rc.chunk_size = n_samples ? (n_samples + n_chunks - 1) / n_chunks : 1;
    if (rc.chunk_size == 0) rc.chunk_size = 1;
    rc.n_chunks = n_samples ? (n_samples + rc.chunk_size - 1) / rc.chunk_size : 0;
    rc.total_jobs = (unsigned long long)rc.n_chunks * base_jobs;
    rc.n_pixels_local = n_pixels;
    rc.lane_jobs = (unsigned long long)rc.n_chunks * n_pixels64;

    CU(ctx->d_accum.ensure((size_t)std::max<uint32_t>(rc.n_chunks, 1) * n_pixels64 * 3));
    CU(ctx->d_rgb8.ensure((size_t)n_pixels64 * 3));
    double* d_fb = static_cast<double*>(d_out);
    if (!d_fb) {
        CU(ctx->d_fb.ensure((size_t)n_pixels64 * stride_d));
        d_fb = ctx->d_fb.p;
        if (p.flags & WRT_FLAG_NO_CLEAR)
            CU(cudaMemcpyAsync(d_fb, host_fb, (size_t)n_pixels64 * stride, cudaMemcpyHostToDevice, ctx->stream));
    }

    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    CU(cudaMemsetAsync(ctx->d_counters.p, 0, 4 * sizeof(unsigned long long), ctx->stream));
    uint32_t launches = 0;
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    unsigned long long wf_rays = 0, wf_paths = 0;
    if (wavefront && rc.total_jobs > 0) {
        // jobs = (chunk, pixel) accumulators; the pool holds at most WRT_WF_POOL_SLOTS of them at a time
        const uint64_t n_jobs64 = (uint64_t)rc.n_chunks * n_pixels64;
        if (n_jobs64 > 0xFFFFFF00ull) return ctx->fail(WRT_E_LIMIT, "wavefront: more than 2^32 (pixel, chunk) jobs");
        uint64_t pool_slots = WRT_WF_POOL_SLOTS;
        uint32_t n_pipes = WRT_WF_PIPELINES;
        if (const char* env = std::getenv("WRT_WF_POOL")) {  // development: pool size in units of 2^20 slots
            const long v = std::atol(env);
            if (v > 0 && v <= 1024) pool_slots = (uint64_t)v << 20;
        }
        if (const char* env = std::getenv("WRT_WF_PIPELINES")) {
            const long v = std::atol(env);
            if (v >= 1 && v <= WRT_WF_MAX_PIPELINES) n_pipes = (uint32_t)v;
        }
        const uint32_t n_slots = (uint32_t)std::min<uint64_t>(n_jobs64, pool_slots);
        if (n_slots < 4096u * n_pipes) n_pipes = 1;  // small frames: one pipeline
        CU(ctx->d_wf_paths.ensure(n_slots));
        CU(ctx->d_wf_queues.ensure((size_t)wrt::WQ_COUNT * n_slots));
        CU(ctx->d_wf_slot_job.ensure(n_slots));
        // ray reordering before the persistent extend kernel (wrt_kernels.h: WavefrontArgs::keys): on for the large trees it
        // serves; WRT_WF_SORT=0 switches it off, WRT_WF_SORT_SHIFT overrides the bucket width (ops per bucket = 2^shift)
        uint32_t sort_shift = 0, sort_buckets = 0;
        {
            const char* env = std::getenv("WRT_WF_SORT");
            const bool on = !packet && ctx->ds.use_wide && !(env && env[0] == '0');
            if (on) {
                sort_shift = 0;
                while ((((uint64_t)ctx->ds.n_ops >> sort_shift) + 1) * 8 > (1u << 18)) ++sort_shift;
                if (const char* s = std::getenv("WRT_WF_SORT_SHIFT")) sort_shift = std::max<uint32_t>(sort_shift, (uint32_t)std::atoi(s));
                sort_buckets = (uint32_t)((((uint64_t)ctx->ds.n_ops >> sort_shift) + 1) * 8);
                CU(ctx->d_wf_keys.ensure(2 * (size_t)n_slots));
                CU(ctx->d_wf_sort_tmp.ensure(n_slots));
                CU(ctx->d_wf_sort_hist.ensure((size_t)sort_buckets * WRT_WF_MAX_PIPELINES));
                CU(cudaMemsetAsync(ctx->d_wf_sort_hist.p, 0, (size_t)sort_buckets * WRT_WF_MAX_PIPELINES * sizeof(uint32_t), ctx->stream));
            }
        }
        CU(ctx->d_wf_counters.ensure(16 * WRT_WF_MAX_PIPELINES + wrt::WS_COUNT));
        if (!ctx->h_wf_counters) CU(cudaMallocHost(&ctx->h_wf_counters, 16 * sizeof(unsigned long long)));
        for (uint32_t k = 1; k < n_pipes; ++k)
            if (!ctx->wf_streams[k]) {
                CU(cudaStreamCreateWithFlags(&ctx->wf_streams[k], cudaStreamNonBlocking));
                CU(cudaEventCreateWithFlags(&ctx->wf_events[k], cudaEventDisableTiming));
            }
        ctx->wf_streams[0] = ctx->stream;
        unsigned long long* d_shared = ctx->d_wf_counters.p + 16 * WRT_WF_MAX_PIPELINES;
        const uint32_t* sobol_matrices = nullptr;
        if (sobol_sampler) {
            if (!ctx->d_sobol_matrices.p) {
                CU(ctx->d_sobol_matrices.ensure(1024 * 52));
                CU(cudaMemcpyAsync(ctx->d_sobol_matrices.p, ctx->blob.matrices32, 1024 * 52 * 4, cudaMemcpyHostToDevice, ctx->stream));
            }
            sobol_matrices = ctx->d_sobol_matrices.p;
        }
        // the pool is cut into n_pipes pipelines (own slots / queues / counters / stream; shared accumulators, job cursor, tallies)
        wrt::WavefrontArgs A[WRT_WF_MAX_PIPELINES];
        uint32_t wf_grid[WRT_WF_MAX_PIPELINES];
        uint32_t first_slot = 0;
        for (uint32_t k = 0; k < n_pipes; ++k) {
            const uint32_t cap = n_slots / n_pipes + (k < n_slots % n_pipes ? 1u : 0u);
            A[k].paths = ctx->d_wf_paths.p + first_slot;
            A[k].queues = ctx->d_wf_queues.p + (size_t)wrt::WQ_COUNT * first_slot;
            A[k].counters = ctx->d_wf_counters.p + 16 * k;
            A[k].accum = ctx->d_accum.p; A[k].capacity = cap; A[k].n_pixels = n_pixels;
            A[k].slot_job = ctx->d_wf_slot_job.p + first_slot; A[k].n_jobs = n_jobs64;
            A[k].sobol_matrices = sobol_matrices;
            A[k].shared = d_shared; A[k].job_base = first_slot;
            A[k].keys = sort_buckets ? ctx->d_wf_keys.p + 2 * (size_t)first_slot : nullptr;
            A[k].sort_tmp = sort_buckets ? ctx->d_wf_sort_tmp.p + first_slot : nullptr;
            A[k].sort_hist = sort_buckets ? ctx->d_wf_sort_hist.p + (size_t)k * sort_buckets : nullptr;
            A[k].sort_buckets = sort_buckets; A[k].sort_shift = sort_shift;
            auto env_u = [](const char* name) { const char* e = std::getenv(name); return e ? (uint32_t)std::atoi(e) : 0u; };
            A[k].node_burst = env_u("WRT_WF_NODE_BURST"); A[k].leaf_burst = env_u("WRT_WF_LEAF_BURST");
            A[k].node_shift = std::getenv("WRT_WF_NODE_SHIFT") ? env_u("WRT_WF_NODE_SHIFT") + 1u : 0u;
            wf_grid[k] = (uint32_t)std::min<uint64_t>((cap + 255) / 256, (uint64_t)ctx->sm_count * 8);
            first_slot += cap;
        }
        CU(cudaMemsetAsync(ctx->d_wf_counters.p, 0, (16 * WRT_WF_MAX_PIPELINES + wrt::WS_COUNT) * sizeof(unsigned long long), ctx->stream));
        const unsigned long long first_free_job = n_slots;  // jobs [0, n_slots) start in the slots
        CU(cudaMemcpyAsync(d_shared + wrt::WS_JOB_CURSOR, &first_free_job, sizeof first_free_job, cudaMemcpyHostToDevice, ctx->stream));
        int ext_blocks = 0;
        CU(wrt::wf_extend_occupancy(&ext_blocks));
        const uint32_t persist_grid = (uint32_t)ctx->sm_count * (uint32_t)std::max(ext_blocks, 1);
        CU(cudaEventRecord(ctx->ev[1], ctx->stream));  // (re-recorded: the pipelines' streams start behind the set-up)
        for (uint32_t k = 1; k < n_pipes; ++k) CU(cudaStreamWaitEvent(ctx->wf_streams[k], ctx->ev[1], 0));
        for (uint32_t k = 0; k < n_pipes; ++k) {
            CU(wrt::wf_launch_init(lp, A[k], wf_grid[k], ctx->wf_streams[k]));
            ++launches;
        }
        // a slot runs at most ceil(jobs / slots) + 1 jobs of chunk_size paths of at most max_depth segments, one segment per
        // iteration (+ one regeneration step per path)
        const uint64_t max_iters = ((n_jobs64 + n_slots - 1) / n_slots + 1) * (uint64_t)rc.chunk_size * (p.max_ray_bounce_depth + 1) + 64;
        const uint32_t check_every = 16;
        bool done = false;
        for (uint64_t it = 0; it < max_iters && !done; ++it) {
            for (uint32_t k = 0; k < n_pipes; ++k) {
                CU(wrt::wf_launch_iteration(lp, A[k], view, p.cull_mode, packet, (uint32_t)(it & 1), wf_grid[k], persist_grid, ctx->wf_streams[k]));
                launches += sort_buckets ? 10 : 6;
            }
            if ((it + 1) % check_every == 0 || it + 1 == max_iters) {
                for (uint32_t k = 1; k < n_pipes; ++k) {  // the tallies are read behind every pipeline
                    CU(cudaEventRecord(ctx->wf_events[k], ctx->wf_streams[k]));
                    CU(cudaStreamWaitEvent(ctx->stream, ctx->wf_events[k], 0));
                }
                CU(cudaMemcpyAsync(ctx->h_wf_counters, d_shared, wrt::WS_COUNT * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
                CU(cudaStreamSynchronize(ctx->stream));
                done = ctx->h_wf_counters[wrt::WS_JOBS_DONE] >= n_jobs64;
            }
        }
        if (!done) return ctx->fail(WRT_E_STATE, "wavefront did not drain within its iteration bound");
        wf_rays = ctx->h_wf_counters[wrt::WS_RAYS];
        wf_paths = ctx->h_wf_counters[wrt::WS_PATHS];
    } else if (rc.total_jobs > 0) {
        if (regroup_engine) CU(wrt::launch_render_regroup(lp, view, p.cull_mode, grid, ctx->d_accum.p, ctx->d_counters.p, ctx->stream));
        else if (sync_engine) CU(wrt::launch_render_sync(lp, view, p.cull_mode, packet, grid, ctx->d_accum.p, ctx->d_counters.p, ctx->stream));
        else CU(wrt::launch_render(lp, view, p.cull_mode, packet, grid, ctx->d_accum.p, ctx->d_counters.p, ctx->stream));
        ++launches;
    }
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    CU(wrt::launch_resolve(ctx->d_accum.p, rc.n_chunks, n_pixels, p.clear_color, (p.flags & WRT_FLAG_NO_CLEAR) ? 1 : 0, d_fb, stride_d,
                           ctx->d_rgb8.p, ctx->stream));
    if (n_pixels) ++launches;
    unsigned long long counters[4] = {0, 0, 0, 0};
    CU(cudaMemcpyAsync(counters, ctx->d_counters.p, sizeof counters, cudaMemcpyDeviceToHost, ctx->stream));
    if (!d_out) CU(cudaMemcpyAsync(host_fb, d_fb, (size_t)n_pixels64 * stride, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaEventRecord(ctx->ev[3], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));

    float ms_total = 0, ms_kernel = 0;
    CU(cudaEventElapsedTime(&ms_total, ctx->ev[0], ctx->ev[3]));
    CU(cudaEventElapsedTime(&ms_kernel, ctx->ev[1], ctx->ev[2]));
    ctx->stats.rays = wavefront ? wf_rays : counters[1];
    // the lane-job kernel does not count paths: every (pixel, chunk) job runs all its samples
    const bool lane_job_kernel = !wavefront && packet && !sync_engine && !regroup_engine;
    ctx->stats.paths = wavefront ? wf_paths : (lane_job_kernel ? n_pixels64 * n_samples : counters[2]);
    ctx->stats.traversal_steps = wavefront ? ctx->h_wf_counters[wrt::WS_STEPS] : counters[3];
    ctx->stats.render_ms = ms_total;
    ctx->stats.kernel_ms = ms_kernel;
    ctx->stats.kernel_ms_min = ctx->stats.kernel_ms_max = ms_kernel;
    ctx->stats.gather_ms = 0.0;
    ctx->stats.n_devices = 1;
    ctx->stats.kernel_launches = launches;
    ctx->last_pixels = n_pixels;
    ctx->last_valid = true;
    return WRT_OK;
}

extern "C" int wrt_render(wrt_ctx* ctx, const wrt_camera* cam, const wrt_params* params, void* framebuffer, size_t pixel_stride_bytes) {
    if (!ctx) return WRT_E_INVALID;
    if (!framebuffer) return ctx->fail(WRT_E_INVALID, "framebuffer is NULL");
    return wrt::render_impl(ctx, cam, params, framebuffer, nullptr, pixel_stride_bytes);
}

extern "C" int wrt_render_device(wrt_ctx* ctx, const wrt_camera* cam, const wrt_params* params, void* d_framebuffer,
                                 size_t pixel_stride_bytes) {
    if (!ctx) return WRT_E_INVALID;
    if (!d_framebuffer) return ctx->fail(WRT_E_INVALID, "d_framebuffer is NULL");
    return wrt::render_impl(ctx, cam, params, nullptr, d_framebuffer, pixel_stride_bytes);
}

extern "C" int wrt_encode_rgb8(wrt_ctx* ctx, uint8_t* rgb_out) {
    if (!ctx) return WRT_E_INVALID;
    int rc_ = bind_device(ctx);
    if (rc_) return rc_;
    if (!ctx->last_valid) return ctx->fail(WRT_E_STATE, "wrt_encode_rgb8: no frame rendered yet");
    if (!rgb_out) return ctx->fail(WRT_E_INVALID, "rgb_out is NULL");
    // the resolve pass already quantised the frame (fused final pass); just fetch it
    CU(cudaMemcpyAsync(rgb_out, ctx->d_rgb8.p, (size_t)ctx->last_pixels * 3, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return WRT_OK;
}

extern "C" int wrt_format_ppm(wrt_ctx* ctx, const uint8_t* rgb8, uint32_t width, uint32_t height, uint8_t* out, uint64_t capacity,
                              uint64_t* content_bytes) {
    if (!ctx) return WRT_E_INVALID;
    int rc_ = bind_device(ctx);
    if (rc_) return rc_;
    if (!out || !content_bytes) return ctx->fail(WRT_E_INVALID, "wrt_format_ppm: out / content_bytes is NULL");
    if (width == 0 || height == 0) return ctx->fail(WRT_E_INVALID, "image dimensions must be non-zero");
    const uint64_t n64 = (uint64_t)width * height;
    if (n64 > 0xFFFFFFFFull) return ctx->fail(WRT_E_LIMIT, "more than 2^32 pixels");
    const uint32_t n_pixels = (uint32_t)n64;
    if (!rgb8) {
        if (!ctx->last_valid) return ctx->fail(WRT_E_STATE, "wrt_format_ppm: no frame rendered yet");
        if (ctx->last_pixels != n_pixels) return ctx->fail(WRT_E_INVALID, "wrt_format_ppm: width x height is not the last rendered frame");
    }
    char header[64];
    const int header_len = std::snprintf(header, sizeof header, "P3\n%u %u\n255\n", width, height);  // writer.zig:9,18
    const uint64_t file_size = (uint64_t)header_len + 12ull * n64;                                     // writer.zig:20
    if (capacity < file_size) return ctx->fail(WRT_E_INVALID, "wrt_format_ppm: capacity < header + 12 bytes per pixel");
    const uint32_t n_blocks = wrt::ppm_block_count(n_pixels);
    const uint8_t* d_rgb = ctx->d_rgb8.p;
    if (rgb8) {
        CU(ctx->d_ppm_in.ensure((size_t)n64 * 3));
        CU(cudaMemcpyAsync(ctx->d_ppm_in.p, rgb8, (size_t)n64 * 3, cudaMemcpyHostToDevice, ctx->stream));
        d_rgb = ctx->d_ppm_in.p;
    }
    CU(ctx->d_ppm_body.ensure((size_t)n64 * 12));
    CU(ctx->d_ppm_blocks.ensure(n_blocks));
    CU(ctx->d_ppm_offsets.ensure((size_t)n_blocks + 1));
    CU(wrt::launch_format_ppm(d_rgb, n_pixels, ctx->d_ppm_blocks.p, ctx->d_ppm_offsets.p, ctx->d_ppm_body.p, ctx->stream));
    unsigned long long body = 0;
    CU(cudaMemcpyAsync(&body, ctx->d_ppm_offsets.p + n_blocks, sizeof body, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    std::memcpy(out, header, (size_t)header_len);
    CU(cudaMemcpyAsync(out + header_len, ctx->d_ppm_body.p, (size_t)body, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    std::memset(out + header_len + body, 0, (size_t)(file_size - header_len - body));  // the NUL tail of the reference's mmap'ed file
    *content_bytes = (uint64_t)header_len + body;
    return WRT_OK;
}

extern "C" int wrt_primary_hits(wrt_ctx* ctx, const wrt_camera* cam, const wrt_params* params, uint32_t n_samples, uint32_t* prim_ids,
                                double* t) {
    if (!ctx) return WRT_E_INVALID;
    int rc_ = bind_device(ctx);
    if (rc_) return rc_;
    if (!ctx->have_scene) return ctx->fail(WRT_E_STATE, "wrt_primary_hits: no scene uploaded");
    if (!cam) return ctx->fail(WRT_E_INVALID, "camera is NULL");
    wrt_params p;
    rc_ = normalise_params(ctx, params, p);
    if (rc_) return rc_;
    rc_ = prepare_sobol(ctx, p.width, p.height);
    if (rc_) return rc_;
    wrt::RenderConstants& rc = ctx->lp.rc;
    fill_constants(*cam, p, rc);
    rc.dof = 0;
    const uint64_t total = (uint64_t)p.width * p.height * n_samples;
    if (total == 0) return WRT_OK;
    DevBuf<uint32_t> d_ids;
    DevBuf<double> d_t;
    int ret = WRT_OK;
    do {
        cudaError_t e;
        if (prim_ids && (e = d_ids.ensure(total)) != cudaSuccess) { ret = ctx->cuda_fail(e, "cudaMalloc(ids)"); break; }
        if (t && (e = d_t.ensure(total)) != cudaSuccess) { ret = ctx->cuda_fail(e, "cudaMalloc(t)"); break; }
        uint32_t grid = (uint32_t)std::min<uint64_t>((total + 127) / 128, (uint64_t)ctx->sm_count * 32);
        if ((e = wrt::launch_primary_hits(ctx->lp, ctx->ds, p.cull_mode, n_samples, prim_ids ? d_ids.p : nullptr, t ? d_t.p : nullptr, grid,
                                          ctx->stream)) != cudaSuccess) { ret = ctx->cuda_fail(e, "primary_hits_kernel"); break; }
        if (prim_ids && (e = cudaMemcpyAsync(prim_ids, d_ids.p, total * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) { ret = ctx->cuda_fail(e, "D2H ids"); break; }
        if (t && (e = cudaMemcpyAsync(t, d_t.p, total * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) { ret = ctx->cuda_fail(e, "D2H t"); break; }
        if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) { ret = ctx->cuda_fail(e, "primary_hits sync"); break; }
    } while (0);
    d_ids.release();
    d_t.release();
    return ret;
}

extern "C" int wrt_trace_rays(wrt_ctx* ctx, const double* origins, const double* directions, uint64_t n, double tmin, uint32_t cull_mode,
                              uint32_t* prim_ids, double* t, double* point, double* normal, double* uv, uint32_t* front_face) {
    if (!ctx) return WRT_E_INVALID;
    int rc_ = bind_device(ctx);
    if (rc_) return rc_;
    if (!ctx->have_scene) return ctx->fail(WRT_E_STATE, "wrt_trace_rays: no scene uploaded");
    uint32_t trav_flags = 0;
    if (cull_mode & WRT_TRAV_FORCE_LANE) trav_flags |= WRT_FLAG_FORCE_LANE;
    if (cull_mode & WRT_TRAV_FORCE_PACKET) trav_flags |= WRT_FLAG_FORCE_PACKET;
    cull_mode &= 0xFFu;
    if (cull_mode > WRT_CULL_TIGHT) return ctx->fail(WRT_E_INVALID, "unknown cull_mode");
    cull_mode = (uint32_t)ctx->resolve_cull(cull_mode);
    ctx->stats.cull_mode_used = cull_mode;
    const bool packet = use_packet(ctx, trav_flags);
    if (n == 0) return WRT_OK;
    if (!origins || !directions) return ctx->fail(WRT_E_INVALID, "origins/directions is NULL");
    DevBuf<double> d_o, d_d, d_t, d_p, d_n, d_uv;
    DevBuf<uint32_t> d_ids, d_ff;
    int ret = WRT_OK;
    do {
        cudaError_t e;
#define TRY(x) if ((e = (x)) != cudaSuccess) { ret = ctx->cuda_fail(e, #x); break; }
        TRY(d_o.ensure(3 * n)); TRY(d_d.ensure(3 * n));
        if (prim_ids) TRY(d_ids.ensure(n));
        if (t) TRY(d_t.ensure(n));
        if (point) TRY(d_p.ensure(3 * n));
        if (normal) TRY(d_n.ensure(3 * n));
        if (uv) TRY(d_uv.ensure(2 * n));
        if (front_face) TRY(d_ff.ensure(n));
        TRY(cudaMemcpyAsync(d_o.p, origins, 3 * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        TRY(cudaMemcpyAsync(d_d.p, directions, 3 * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        uint32_t grid = (uint32_t)std::min<uint64_t>((n + 127) / 128, (uint64_t)ctx->sm_count * 32);
        TRY(wrt::launch_trace_rays(scene_view(ctx, cull_mode, packet), cull_mode, packet, d_o.p, d_d.p, n, tmin, prim_ids ? d_ids.p : nullptr, t ? d_t.p : nullptr,
                                   point ? d_p.p : nullptr, normal ? d_n.p : nullptr, uv ? d_uv.p : nullptr,
                                   front_face ? d_ff.p : nullptr, grid, ctx->stream));
        if (prim_ids) TRY(cudaMemcpyAsync(prim_ids, d_ids.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        if (t) TRY(cudaMemcpyAsync(t, d_t.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (point) TRY(cudaMemcpyAsync(point, d_p.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (normal) TRY(cudaMemcpyAsync(normal, d_n.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (uv) TRY(cudaMemcpyAsync(uv, d_uv.p, 2 * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (front_face) TRY(cudaMemcpyAsync(front_face, d_ff.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        TRY(cudaStreamSynchronize(ctx->stream));
#undef TRY
    } while (0);
    d_o.release(); d_d.release(); d_t.release(); d_p.release(); d_n.release(); d_uv.release(); d_ids.release(); d_ff.release();
    return ret;
}

extern "C" int wrt_sobol_pixel_samples(wrt_ctx* ctx, uint32_t width, uint32_t height, const uint32_t* cols, const uint32_t* rows,
                                       const uint32_t* sample_idx, uint64_t n, uint64_t* sobol_index, double* offsets_xy) {
    if (!ctx) return WRT_E_INVALID;
    int rc_ = bind_device(ctx);
    if (rc_) return rc_;
    if (n == 0) return WRT_OK;
    if (!cols || !rows || !sample_idx) return ctx->fail(WRT_E_INVALID, "cols/rows/sample_idx is NULL");
    rc_ = prepare_sobol(ctx, width, height);
    if (rc_) return rc_;
    DevBuf<uint32_t> d_c, d_r, d_s;
    DevBuf<uint64_t> d_i;
    DevBuf<double> d_o;
    int ret = WRT_OK;
    do {
        cudaError_t e;
#define TRY(x) if ((e = (x)) != cudaSuccess) { ret = ctx->cuda_fail(e, #x); break; }
        TRY(d_c.ensure(n)); TRY(d_r.ensure(n)); TRY(d_s.ensure(n));
        if (sobol_index) TRY(d_i.ensure(n));
        if (offsets_xy) TRY(d_o.ensure(2 * n));
        TRY(cudaMemcpyAsync(d_c.p, cols, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        TRY(cudaMemcpyAsync(d_r.p, rows, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        TRY(cudaMemcpyAsync(d_s.p, sample_idx, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        TRY(wrt::launch_sobol_pixel(ctx->lp, d_c.p, d_r.p, d_s.p, n, sobol_index ? d_i.p : nullptr, offsets_xy ? d_o.p : nullptr, ctx->stream));
        if (sobol_index) TRY(cudaMemcpyAsync(sobol_index, d_i.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (offsets_xy) TRY(cudaMemcpyAsync(offsets_xy, d_o.p, 2 * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        TRY(cudaStreamSynchronize(ctx->stream));
#undef TRY
    } while (0);
    d_c.release(); d_r.release(); d_s.release(); d_i.release(); d_o.release();
    return ret;
}

extern "C" int wrt_sobol_dimension_samples(wrt_ctx* ctx, const uint64_t* sobol_index, const uint32_t* dimension, uint64_t n,
                                           uint32_t owen_fast, uint32_t seed, float* out) {
    if (!ctx) return WRT_E_INVALID;
    int rc_ = bind_device(ctx);
    if (rc_) return rc_;
    if (n == 0) return WRT_OK;
    if (!sobol_index || !dimension || !out) return ctx->fail(WRT_E_INVALID, "argument is NULL");
    for (uint64_t i = 0; i < n; ++i)
        if (dimension[i] >= 1024) return ctx->fail(WRT_E_INVALID, "Sobol dimension >= NSobolDimensions (1024)");
    DevBuf<uint64_t> d_i;
    DevBuf<uint32_t> d_d;
    DevBuf<float> d_o;
    int ret = WRT_OK;
    do {
        cudaError_t e;
#define TRY(x) if ((e = (x)) != cudaSuccess) { ret = ctx->cuda_fail(e, #x); break; }
        if (!ctx->d_sobol_matrices.p) {
            TRY(ctx->d_sobol_matrices.ensure(1024 * 52));
            TRY(cudaMemcpyAsync(ctx->d_sobol_matrices.p, ctx->blob.matrices32, 1024 * 52 * 4, cudaMemcpyHostToDevice, ctx->stream));
        }
        TRY(d_i.ensure(n)); TRY(d_d.ensure(n)); TRY(d_o.ensure(n));
        TRY(cudaMemcpyAsync(d_i.p, sobol_index, n * 8, cudaMemcpyHostToDevice, ctx->stream));
        TRY(cudaMemcpyAsync(d_d.p, dimension, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        TRY(wrt::launch_sobol_dimension(ctx->d_sobol_matrices.p, d_i.p, d_d.p, n, owen_fast, seed, d_o.p, ctx->stream));
        TRY(cudaMemcpyAsync(out, d_o.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        TRY(cudaStreamSynchronize(ctx->stream));
#undef TRY
    } while (0);
    d_i.release(); d_d.release(); d_o.release();
    return ret;
}

static int issue_peak(wrt_ctx* ctx, double* fma_per_second, bool fp32) {
    if (!ctx) return WRT_E_INVALID;
    int rc_ = bind_device(ctx);
    if (rc_) return rc_;
    if (!fma_per_second) return ctx->fail(WRT_E_INVALID, "fma_per_second is NULL");
    const uint32_t block = 256, grid = (uint32_t)ctx->sm_count * 8, iters = 1u << 16;
    auto launch = fp32 ? wrt::launch_fp32_peak : wrt::launch_fp64_peak;
    DevBuf<double> d_out;
    int ret = WRT_OK;
    do {
        cudaError_t e;
#define TRY(x) if ((e = (x)) != cudaSuccess) { ret = ctx->cuda_fail(e, #x); break; }
        TRY(d_out.ensure((size_t)grid * block));
        TRY(launch(d_out.p, grid, block, 1u << 10, ctx->stream));  // warm-up
        double best = 0.0;
        for (int rep = 0; rep < 5; ++rep) {
            TRY(cudaEventRecord(ctx->ev[0], ctx->stream));
            TRY(launch(d_out.p, grid, block, iters, ctx->stream));
            TRY(cudaEventRecord(ctx->ev[1], ctx->stream));
            TRY(cudaStreamSynchronize(ctx->stream));
            float ms = 0;
            TRY(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
            double rate = (double)grid * block * iters * 8.0 / (ms * 1e-3);
            if (rate > best) best = rate;
        }
        if (ret == WRT_OK) *fma_per_second = best;
#undef TRY
    } while (0);
    d_out.release();
    return ret;
}
extern "C" int wrt_fp64_issue_peak(wrt_ctx* ctx, double* fma_per_second) { return issue_peak(ctx, fma_per_second, false); }
extern "C" int wrt_fp32_issue_peak(wrt_ctx* ctx, double* fma_per_second) { return issue_peak(ctx, fma_per_second, true); }

extern "C" int wrt_get_stats(const wrt_ctx* ctx, wrt_stats* out) {
    if (!ctx || !out) return WRT_E_INVALID;
    *out = ctx->stats;
    return WRT_OK;
}
