// wrt_multi.cu — the multi-GPU half of include/wrt.h: device groups (one process, n devices) and sharded rendering across
// processes (one device each), both ending in an NCCL gather of the shards over NVLink into one device's frame.
//
// The reference fans disjoint row segments out over a thread pool and joins them with a WaitGroup (src/render.zig:55-73); here
// device r of n renders rows r, r+n, ... (wrt_params.row_shard_*), ships them as 3 binary64 lanes per pixel with grouped
// ncclSend / ncclRecv, and the root interleaves them into the caller's framebuffer layout and quantises the RGB8 frame in one
// pass (assemble_kernel).  WRT_FLAG_SHARD_SAMPLES splits the sample range instead and adds the partial means with
// ncclReduce(ncclSum, ncclDouble).
//
// NCCL is bound at run time (dlopen): inside a PyTorch process that resolves to the libnccl.so.2 torch already loaded (one
// NCCL per process, its NCCL_DEBUG log shows these communicators too), elsewhere to the system library.
#include <dlfcn.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>

#include <chrono>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>

#include "wrt_ctx.h"

namespace {

struct NcclApi {
    void* handle = nullptr;
    std::string err;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    bool ok() const { return handle != nullptr; }
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        void* h = nullptr;
        for (const char* n : names)  // already in the process (PyTorch's bundled copy)?
            if ((h = dlopen(n, RTLD_NOW | RTLD_NOLOAD)) != nullptr) break;
        if (!h)
            for (const char* n : names)
                if ((h = dlopen(n, RTLD_NOW | RTLD_GLOBAL)) != nullptr) break;
        if (!h) { api.err = std::string("NCCL is not available: ") + (dlerror() ? dlerror() : "libnccl.so.2 not found"); return; }
        bool all = true;
        auto sym = [&](auto& fn, const char* name) {
            fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(h, name));
            if (!fn) { all = false; api.err = std::string("NCCL symbol missing: ") + name; }
        };
        sym(api.GetVersion, "ncclGetVersion"); sym(api.GetUniqueId, "ncclGetUniqueId"); sym(api.CommInitRank, "ncclCommInitRank");
        sym(api.CommInitAll, "ncclCommInitAll"); sym(api.CommDestroy, "ncclCommDestroy"); sym(api.GetErrorString, "ncclGetErrorString");
        sym(api.GroupStart, "ncclGroupStart"); sym(api.GroupEnd, "ncclGroupEnd"); sym(api.Send, "ncclSend"); sym(api.Recv, "ncclRecv");
        sym(api.Reduce, "ncclReduce");
        if (all) api.handle = h;
    });
    return api;
}

thread_local std::string g_group_error;

struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

uint32_t rows_of(uint32_t height, uint32_t index, uint32_t count) {
    return index >= height ? 0u : (height - index + count - 1) / count;
}

// Root: interleave the members' row shards (3 lanes per pixel, member r holds rows r, r+n, ...) into the caller's layout
// (one pixel every stride_d doubles, lanes >= 3 zero) and quantise like encodeColor (writer.zig:68-94).
__global__ void assemble_rows_kernel(const double* __restrict__ staging, uint32_t n_members, uint64_t member_stride, uint32_t width,
                                     uint32_t height, double* __restrict__ fb, uint32_t stride_d, uint8_t* __restrict__ rgb8) {
    const uint64_t n = (uint64_t)width * height;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t row = (uint32_t)(i / width), col = (uint32_t)(i - (uint64_t)row * width);
        const uint32_t member = row % n_members, local_row = row / n_members;
        const double* src = staging + member * member_stride + ((uint64_t)local_row * width + col) * 3;
        const double r = src[0], g = src[1], b = src[2];
        double* dst = fb + i * stride_d;
        dst[0] = r; dst[1] = g; dst[2] = b;
        for (uint32_t k = 3; k < stride_d; ++k) dst[k] = 0.0;
        rgb8[3 * i + 0] = wrt::encode_channel(r); rgb8[3 * i + 1] = wrt::encode_channel(g); rgb8[3 * i + 2] = wrt::encode_channel(b);
    }
}
// Root, sample split: frame = clear colour + sum over the members of their partial means (already reduced into `sum`).
__global__ void assemble_sum_kernel(const double* __restrict__ sum, uint64_t n_pixels, double cr, double cg, double cb,
                                    double* __restrict__ fb, uint32_t stride_d, uint8_t* __restrict__ rgb8) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_pixels; i += (uint64_t)gridDim.x * blockDim.x) {
        const double r = cr + sum[3 * i], g = cg + sum[3 * i + 1], b = cb + sum[3 * i + 2];
        double* dst = fb + i * stride_d;
        dst[0] = r; dst[1] = g; dst[2] = b;
        for (uint32_t k = 3; k < stride_d; ++k) dst[k] = 0.0;
        rgb8[3 * i + 0] = wrt::encode_channel(r); rgb8[3 * i + 1] = wrt::encode_channel(g); rgb8[3 * i + 2] = wrt::encode_channel(b);
    }
}

// The shard a member renders: its params, where it lands and how many doubles travel.
struct ShardPlan {
    wrt_params p;
    uint64_t doubles = 0;
};

// Splits `base` n ways.  Rows: member i takes rows i, i+n, ...; samples: member i takes a contiguous slice of the sample range
// and renders every row onto a zero clear colour (the root adds the clear colour once).
int plan_shards(const wrt_params& base, int n, std::vector<ShardPlan>& plans, bool& by_samples, std::string& err) {
    if (base.flags & WRT_FLAG_NO_CLEAR) { err = "WRT_FLAG_NO_CLEAR is not supported by group / sharded renders"; return WRT_E_INVALID; }
    if (base.width == 0 || base.height == 0) { err = "image dimensions must be non-zero"; return WRT_E_INVALID; }
    by_samples = (base.flags & WRT_FLAG_SHARD_SAMPLES) != 0;
    uint32_t s0 = base.sample_begin, s1 = base.sample_end;
    if (s0 == 0 && s1 == 0) s1 = base.samples_per_pixel;
    if (s0 > s1 || s1 > base.samples_per_pixel) { err = "bad sample range"; return WRT_E_INVALID; }
    plans.assign((size_t)n, ShardPlan{});
    for (int i = 0; i < n; ++i) {
        wrt_params p = base;
        p.flags &= ~(uint32_t)WRT_FLAG_SHARD_SAMPLES;
        if (by_samples) {
            const uint64_t span = s1 - s0;
            p.sample_begin = s0 + (uint32_t)(span * (uint64_t)i / (uint64_t)n);
            p.sample_end = s0 + (uint32_t)(span * (uint64_t)(i + 1) / (uint64_t)n);
            p.row_shard_index = 0; p.row_shard_count = 1;
            p.clear_color[0] = p.clear_color[1] = p.clear_color[2] = 0.0;
            plans[i].doubles = (uint64_t)base.width * base.height * 3;
            if (p.sample_begin == p.sample_end) { p.sample_begin = p.sample_end = 0; p.samples_per_pixel = base.samples_per_pixel; p.max_ray_bounce_depth = 0; }
        } else {
            p.sample_begin = s0; p.sample_end = s1;
            p.row_shard_index = (uint32_t)i; p.row_shard_count = (uint32_t)n;
            plans[i].doubles = (uint64_t)rows_of(base.height, (uint32_t)i, (uint32_t)n) * base.width * 3;
        }
        plans[i].p = p;
    }
    return WRT_OK;
}

#define CUG(ctx, call)                                              \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) return (ctx)->cuda_fail(e__, #call); \
    } while (0)

int nccl_fail(wrt_ctx* ctx, ncclResult_t r, const char* what) {
    ctx->err = std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error");
    return WRT_E_CUDA;
}
#define NC(ctx, call)                                              \
    do {                                                           \
        ncclResult_t r__ = (call);                                 \
        if (r__ != ncclSuccess) return nccl_fail((ctx), r__, #call); \
    } while (0)

// Root side after the transfers were enqueued on root->stream: assemble, optional D2H, bookkeeping.
int root_finish(wrt_ctx* root, const wrt_params& base, bool by_samples, int n, uint64_t member_stride, void* host_fb, size_t stride) {
    const uint32_t stride_d = (uint32_t)(stride / 8);
    const uint64_t n_pixels = (uint64_t)base.width * base.height;
    const uint32_t grid = (uint32_t)std::min<uint64_t>((n_pixels + 255) / 256, (uint64_t)root->sm_count * 16);
    if (by_samples)
        assemble_sum_kernel<<<grid, 256, 0, root->stream>>>(root->d_staging.p, n_pixels, base.clear_color[0], base.clear_color[1],
                                                            base.clear_color[2], root->d_fb.p, stride_d, root->d_rgb8.p);
    else
        assemble_rows_kernel<<<grid, 256, 0, root->stream>>>(root->d_staging.p, (uint32_t)n, member_stride, base.width, base.height,
                                                             root->d_fb.p, stride_d, root->d_rgb8.p);
    CUG(root, cudaGetLastError());
    CUG(root, cudaEventRecord(root->ev[1], root->stream));
    if (host_fb) {
        NvtxRange range("wrt: D2H frame");
        CUG(root, cudaMemcpyAsync(host_fb, root->d_fb.p, (size_t)n_pixels * stride, cudaMemcpyDeviceToHost, root->stream));
    }
    CUG(root, cudaStreamSynchronize(root->stream));
    float ms = 0;
    CUG(root, cudaEventElapsedTime(&ms, root->ev[0], root->ev[1]));
    root->stats.gather_ms = ms;
    root->stats.kernel_launches += 1;
    root->full_w = base.width; root->full_h = base.height;
    root->last_pixels = (uint32_t)n_pixels;
    root->last_valid = true;
    return WRT_OK;
}

int check_frame_args(wrt_ctx* ctx, const wrt_camera* cam, const wrt_params* params, size_t stride) {
    if (!ctx->have_scene) return ctx->fail(WRT_E_STATE, "no scene uploaded");
    if (!cam || !params) return ctx->fail(WRT_E_INVALID, "camera / params is NULL");
    if (stride < 24 || stride % 8) return ctx->fail(WRT_E_INVALID, "pixel_stride_bytes must be a multiple of 8 and >= 24");
    if ((uint64_t)params->width * params->height > 0xFFFFFFFFull) return ctx->fail(WRT_E_LIMIT, "more than 2^32 pixels");
    return WRT_OK;
}

}  // namespace

namespace wrt {
struct CommState {
    ncclComm_t comm = nullptr;
    int rank = 0, n_ranks = 1;
};
void comm_release(wrt_ctx* ctx) {
    if (!ctx->comm) return;
    if (ctx->comm->comm && nccl().ok()) nccl().CommDestroy(ctx->comm->comm);
    delete ctx->comm;
    ctx->comm = nullptr;
}
}  // namespace wrt

// ---------------------------------------------------------------------------------------------------------------------
// (1) device group: one process, n devices
// ---------------------------------------------------------------------------------------------------------------------
struct wrt_group {
    std::vector<wrt_ctx*> ctx;
    std::vector<ncclComm_t> comms;
    std::string err;
    wrt_stats stats{};
    int fail(int code, const std::string& msg) { err = msg; return code; }
};

extern "C" const char* wrt_group_last_error(const wrt_group* g) { return g ? g->err.c_str() : g_group_error.c_str(); }
extern "C" int wrt_group_size(const wrt_group* g) { return g ? (int)g->ctx.size() : 0; }
extern "C" wrt_ctx* wrt_group_ctx(wrt_group* g, int i) { return (g && i >= 0 && i < (int)g->ctx.size()) ? g->ctx[(size_t)i] : nullptr; }

extern "C" void wrt_group_destroy(wrt_group* g) {
    if (!g) return;
    for (size_t i = 0; i < g->comms.size(); ++i)
        if (g->comms[i] && nccl().ok()) {
            cudaSetDevice(g->ctx[i]->device);
            nccl().CommDestroy(g->comms[i]);
        }
    for (wrt_ctx* c : g->ctx) wrt_destroy(c);
    delete g;
}

extern "C" int wrt_group_create(const int* device_ids, int n_devices, wrt_group** out) {
    if (!out) { g_group_error = "wrt_group_create: out is NULL"; return WRT_E_INVALID; }
    *out = nullptr;
    if (!device_ids || n_devices < 1 || n_devices > 64) { g_group_error = "wrt_group_create: need 1..64 device ids"; return WRT_E_INVALID; }
    for (int i = 0; i < n_devices; ++i)
        for (int j = 0; j < i; ++j)
            if (device_ids[i] == device_ids[j]) { g_group_error = "wrt_group_create: duplicate device id"; return WRT_E_INVALID; }
    wrt_group* g = new (std::nothrow) wrt_group();
    if (!g) { g_group_error = "wrt_group_create: out of host memory"; return WRT_E_NOMEM; }
    try {
        for (int i = 0; i < n_devices; ++i) {
            wrt_ctx* c = nullptr;
            const int rc = wrt_create(device_ids[i], &c);
            if (rc != WRT_OK) {
                g_group_error = std::string("wrt_group_create: device ") + std::to_string(device_ids[i]) + ": " + wrt_last_error(nullptr);
                wrt_group_destroy(g);
                return rc;
            }
            g->ctx.push_back(c);
        }
        if (n_devices > 1) {
            if (!nccl().ok()) { g_group_error = "wrt_group_create: " + nccl().err; wrt_group_destroy(g); return WRT_E_CUDA; }
            g->comms.assign((size_t)n_devices, nullptr);
            const ncclResult_t r = nccl().CommInitAll(g->comms.data(), n_devices, device_ids);
            if (r != ncclSuccess) {
                g_group_error = std::string("ncclCommInitAll: ") + nccl().GetErrorString(r);
                g->comms.clear();
                wrt_group_destroy(g);
                return WRT_E_CUDA;
            }
        }
    } catch (const std::bad_alloc&) {
        g_group_error = "wrt_group_create: out of host memory";
        wrt_group_destroy(g);
        return WRT_E_NOMEM;
    }
    *out = g;
    return WRT_OK;
}

extern "C" int wrt_group_upload_scene(wrt_group* g, const wrt_scene* scene) {
    if (!g) return WRT_E_INVALID;
    try {
        auto t0 = std::chrono::steady_clock::now();
        wrt::CompiledScene cs;  // compiled once, copied to every device
        std::string err;
        int rc;
        double tree_ms = 0.0;
        bool tree_on_device = false;
        {
            NvtxRange range("wrt_group_upload_scene: compile + tree build");
            rc = wrt::compile_scene_for_device(scene, cs, err, g->ctx[0]->device, &tree_ms, &tree_on_device);
        }
        if (rc != WRT_OK) return g->fail(rc, "wrt_group_upload_scene: " + err);
        const double compile_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        const size_t n = g->ctx.size();
        std::vector<int> rcs(n, WRT_OK);
        std::vector<std::thread> threads;
        for (size_t i = 1; i < n; ++i) threads.emplace_back([&, i] { rcs[i] = wrt::upload_compiled(g->ctx[i], cs, scene, compile_ms); });
        rcs[0] = wrt::upload_compiled(g->ctx[0], cs, scene, compile_ms);
        for (auto& t : threads) t.join();
        for (size_t i = 0; i < n; ++i)
            if (rcs[i] != WRT_OK) return g->fail(rcs[i], "wrt_group_upload_scene: device " + std::to_string(g->ctx[i]->device) + ": " + g->ctx[i]->err);
        for (size_t i = 0; i < n; ++i) {
            g->ctx[i]->stats.tree_build_ms = tree_ms;
            g->ctx[i]->stats.tree_build_device = tree_on_device ? 1u : 0u;
            g->ctx[i]->stats.n_tree_records = (uint32_t)(cs.use_wide ? cs.nodes4.size() : cs.nodes2.size());
        }
        return WRT_OK;
    } catch (const std::bad_alloc&) {
        return g->fail(WRT_E_NOMEM, "wrt_group_upload_scene: out of host memory");
    } catch (const std::exception& e) {
        return g->fail(WRT_E_INVALID, std::string("wrt_group_upload_scene: ") + e.what());
    }
}

extern "C" int wrt_group_render(wrt_group* g, const wrt_camera* cam, const wrt_params* params, void* framebuffer, size_t stride) {
    if (!g) return WRT_E_INVALID;
    NvtxRange range("wrt_group_render");
    const int n = (int)g->ctx.size();
    wrt_ctx* root = g->ctx[0];
    int rc = check_frame_args(root, cam, params, stride);
    if (rc) return g->fail(rc, root->err);
    try {
        std::vector<ShardPlan> plans;
        bool by_samples = false;
        std::string err;
        rc = plan_shards(*params, n, plans, by_samples, err);
        if (rc) return g->fail(rc, "wrt_group_render: " + err);
        const uint64_t n_pixels = (uint64_t)params->width * params->height;
        uint64_t member_stride = 0;
        for (const ShardPlan& sp : plans) member_stride = std::max(member_stride, sp.doubles);
        if ((rc = wrt::bind_device(root))) return g->fail(rc, root->err);
        // root buffers: staging (member 0 renders straight into its slot), the assembled frame in the caller's stride, RGB8
        cudaError_t ce;
        if ((ce = root->d_staging.ensure((size_t)(by_samples ? member_stride : member_stride * (uint64_t)n))) != cudaSuccess ||
            (ce = root->d_fb.ensure((size_t)n_pixels * (stride / 8))) != cudaSuccess ||
            (ce = root->d_rgb8.ensure((size_t)n_pixels * 3)) != cudaSuccess)
            return g->fail(root->cuda_fail(ce, "cudaMalloc(group frame)"), root->err);
        for (int i = 1; i < n; ++i) {
            cudaSetDevice(g->ctx[(size_t)i]->device);
            if ((ce = g->ctx[(size_t)i]->d_shard.ensure((size_t)plans[(size_t)i].doubles)) != cudaSuccess)
                return g->fail(g->ctx[(size_t)i]->cuda_fail(ce, "cudaMalloc(shard)"), g->ctx[(size_t)i]->err);
        }
        // fan-out: one host thread per device, like the reference's thread pool over row jobs
        std::vector<int> rcs((size_t)n, WRT_OK);
        std::vector<std::thread> threads;
        auto run = [&](int i) {
            wrt_ctx* c = g->ctx[(size_t)i];
            double* dst = (i == 0) ? root->d_staging.p : c->d_shard.p;
            if (plans[(size_t)i].doubles == 0) { c->stats = wrt_stats{}; rcs[(size_t)i] = WRT_OK; return; }
            rcs[(size_t)i] = wrt::render_impl(c, cam, &plans[(size_t)i].p, nullptr, dst, 24);
        };
        for (int i = 1; i < n; ++i) threads.emplace_back(run, i);
        run(0);
        for (auto& t : threads) t.join();
        for (int i = 0; i < n; ++i)
            if (rcs[(size_t)i] != WRT_OK) return g->fail(rcs[(size_t)i], "wrt_group_render: device " + std::to_string(g->ctx[(size_t)i]->device) + ": " + g->ctx[(size_t)i]->err);
        // gather / reduce into the root
        if ((rc = wrt::bind_device(root))) return g->fail(rc, root->err);
        if ((ce = cudaEventRecord(root->ev[0], root->stream)) != cudaSuccess) return g->fail(root->cuda_fail(ce, "cudaEventRecord"), root->err);
        if (n > 1) {
            NvtxRange gather("wrt_group_render: NCCL gather");
            ncclResult_t r = nccl().GroupStart();
            for (int i = 0; i < n && r == ncclSuccess; ++i) {
                wrt_ctx* c = g->ctx[(size_t)i];
                if (by_samples) {
                    const void* src = (i == 0) ? (const void*)root->d_staging.p : (const void*)c->d_shard.p;
                    r = nccl().Reduce(src, root->d_staging.p, (size_t)member_stride, ncclDouble, ncclSum, 0, g->comms[(size_t)i], c->stream);
                } else if (i > 0 && plans[(size_t)i].doubles) {
                    r = nccl().Send(c->d_shard.p, (size_t)plans[(size_t)i].doubles, ncclDouble, 0, g->comms[(size_t)i], c->stream);
                    if (r == ncclSuccess)
                        r = nccl().Recv(root->d_staging.p + (uint64_t)i * member_stride, (size_t)plans[(size_t)i].doubles, ncclDouble, i,
                                        g->comms[0], root->stream);
                }
            }
            const ncclResult_t r2 = nccl().GroupEnd();
            if (r != ncclSuccess || r2 != ncclSuccess) return g->fail(nccl_fail(root, r != ncclSuccess ? r : r2, "NCCL gather"), root->err);
        }
        rc = root_finish(root, *params, by_samples, n, member_stride, framebuffer, stride);
        if (rc) return g->fail(rc, root->err);
        for (int i = 1; i < n; ++i) {  // the senders' streams
            cudaSetDevice(g->ctx[(size_t)i]->device);
            cudaStreamSynchronize(g->ctx[(size_t)i]->stream);
        }
        // whole-job numbers
        wrt_stats st = root->stats;
        st.paths = st.rays = 0; st.traversal_steps = 0;
        st.kernel_launches = 0;
        double k_min = 1e300, k_max = 0.0, r_max = 0.0;
        for (int i = 0; i < n; ++i) {
            const wrt_stats& s = g->ctx[(size_t)i]->stats;
            st.paths += s.paths; st.rays += s.rays; st.traversal_steps += s.traversal_steps; st.kernel_launches += s.kernel_launches;
            if (plans[(size_t)i].doubles) { k_min = std::min(k_min, s.kernel_ms); k_max = std::max(k_max, s.kernel_ms); r_max = std::max(r_max, s.render_ms); }
        }
        st.kernel_ms = k_max; st.kernel_ms_min = (k_min == 1e300) ? 0.0 : k_min; st.kernel_ms_max = k_max;
        st.gather_ms = root->stats.gather_ms;
        st.render_ms = r_max + st.gather_ms;
        st.n_devices = (uint32_t)n;
        g->stats = st;
        return WRT_OK;
    } catch (const std::bad_alloc&) {
        return g->fail(WRT_E_NOMEM, "wrt_group_render: out of host memory");
    } catch (const std::exception& e) {
        return g->fail(WRT_E_STATE, std::string("wrt_group_render: ") + e.what());
    }
}

extern "C" int wrt_group_encode_rgb8(wrt_group* g, uint8_t* rgb_out) {
    if (!g) return WRT_E_INVALID;
    const int rc = wrt_encode_rgb8(g->ctx[0], rgb_out);
    if (rc) g->err = g->ctx[0]->err;
    return rc;
}

extern "C" int wrt_group_get_stats(const wrt_group* g, wrt_stats* out) {
    if (!g || !out) return WRT_E_INVALID;
    *out = g->stats;
    return WRT_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// (2) one process per device
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int wrt_comm_unique_id(uint8_t id[WRT_COMM_ID_BYTES]) {
    static_assert(sizeof(ncclUniqueId) == WRT_COMM_ID_BYTES, "WRT_COMM_ID_BYTES must be sizeof(ncclUniqueId)");
    if (!id) return WRT_E_INVALID;
    if (!nccl().ok()) { g_group_error = nccl().err; return WRT_E_CUDA; }
    ncclUniqueId uid;
    const ncclResult_t r = nccl().GetUniqueId(&uid);
    if (r != ncclSuccess) { g_group_error = std::string("ncclGetUniqueId: ") + nccl().GetErrorString(r); return WRT_E_CUDA; }
    std::memcpy(id, &uid, sizeof uid);
    return WRT_OK;
}

extern "C" int wrt_comm_init(wrt_ctx* ctx, const uint8_t id[WRT_COMM_ID_BYTES], int rank, int n_ranks) {
    if (!ctx) return WRT_E_INVALID;
    if (!id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return ctx->fail(WRT_E_INVALID, "wrt_comm_init: bad id / rank / n_ranks");
    if (!nccl().ok()) return ctx->fail(WRT_E_CUDA, "wrt_comm_init: " + nccl().err);
    int rc = wrt::bind_device(ctx);
    if (rc) return rc;
    wrt::comm_release(ctx);
    wrt::CommState* cs = new (std::nothrow) wrt::CommState();
    if (!cs) return ctx->fail(WRT_E_NOMEM, "wrt_comm_init: out of host memory");
    cs->rank = rank; cs->n_ranks = n_ranks;
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof uid);
    const ncclResult_t r = nccl().CommInitRank(&cs->comm, n_ranks, uid, rank);
    if (r != ncclSuccess) { delete cs; return nccl_fail(ctx, r, "ncclCommInitRank"); }
    ctx->comm = cs;
    return WRT_OK;
}

extern "C" int wrt_render_sharded(wrt_ctx* ctx, const wrt_camera* cam, const wrt_params* params, void* framebuffer, size_t stride) {
    if (!ctx) return WRT_E_INVALID;
    NvtxRange range("wrt_render_sharded");
    if (!ctx->comm) return ctx->fail(WRT_E_STATE, "wrt_render_sharded: wrt_comm_init was not called on this context");
    int rc = check_frame_args(ctx, cam, params, stride);
    if (rc) return rc;
    try {
        const int n = ctx->comm->n_ranks, me = ctx->comm->rank;
        std::vector<ShardPlan> plans;
        bool by_samples = false;
        std::string err;
        rc = plan_shards(*params, n, plans, by_samples, err);
        if (rc) return ctx->fail(rc, "wrt_render_sharded: " + err);
        const uint64_t n_pixels = (uint64_t)params->width * params->height;
        uint64_t member_stride = 0;
        for (const ShardPlan& sp : plans) member_stride = std::max(member_stride, sp.doubles);
        if ((rc = wrt::bind_device(ctx))) return rc;
        double* dst = nullptr;
        if (me == 0) {
            CUG(ctx, ctx->d_staging.ensure((size_t)(by_samples ? member_stride : member_stride * (uint64_t)n)));
            CUG(ctx, ctx->d_fb.ensure((size_t)n_pixels * (stride / 8)));
            CUG(ctx, ctx->d_rgb8.ensure((size_t)n_pixels * 3));
            dst = ctx->d_staging.p;
        } else {
            CUG(ctx, ctx->d_shard.ensure((size_t)plans[(size_t)me].doubles));
            dst = ctx->d_shard.p;
        }
        if (plans[(size_t)me].doubles) {
            rc = wrt::render_impl(ctx, cam, &plans[(size_t)me].p, nullptr, dst, 24);
            if (rc) return rc;
        } else {
            ctx->stats.rays = ctx->stats.paths = 0; ctx->stats.kernel_ms = ctx->stats.render_ms = 0.0; ctx->stats.kernel_launches = 0;
        }
        ctx->stats.n_devices = (uint32_t)n;
        CUG(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
        if (n > 1) {
            NvtxRange gather("wrt_render_sharded: NCCL gather");
            ncclComm_t comm = ctx->comm->comm;
            if (by_samples) {
                NC(ctx, nccl().Reduce(dst, me == 0 ? (void*)ctx->d_staging.p : nullptr, (size_t)member_stride, ncclDouble, ncclSum, 0, comm, ctx->stream));
            } else if (me == 0) {
                NC(ctx, nccl().GroupStart());
                ncclResult_t r = ncclSuccess;
                for (int i = 1; i < n && r == ncclSuccess; ++i)
                    if (plans[(size_t)i].doubles)
                        r = nccl().Recv(ctx->d_staging.p + (uint64_t)i * member_stride, (size_t)plans[(size_t)i].doubles, ncclDouble, i, comm, ctx->stream);
                const ncclResult_t r2 = nccl().GroupEnd();
                if (r != ncclSuccess || r2 != ncclSuccess) return nccl_fail(ctx, r != ncclSuccess ? r : r2, "ncclRecv");
            } else if (plans[(size_t)me].doubles) {
                NC(ctx, nccl().Send(ctx->d_shard.p, (size_t)plans[(size_t)me].doubles, ncclDouble, 0, comm, ctx->stream));
            }
        }
        if (me == 0) return root_finish(ctx, *params, by_samples, n, member_stride, framebuffer, stride);
        CUG(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
        CUG(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0;
        CUG(ctx, cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        ctx->stats.gather_ms = ms;
        return WRT_OK;
    } catch (const std::bad_alloc&) {
        return ctx->fail(WRT_E_NOMEM, "wrt_render_sharded: out of host memory");
    }
}
