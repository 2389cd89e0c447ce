// wrt_ctx.h — the context object behind the opaque wrt_ctx of include/wrt.h, shared by wrt_api.cu (single device) and
// wrt_multi.cu (device groups / NCCL gather).  Internal: nothing here crosses the C ABI.
#pragma once

#include <string>
#include <vector>

#include "wrt_device.cuh"
#include "wrt_kernels.h"
#include "wrt_program.h"

#define WRT_WF_MAX_PIPELINES 4

namespace wrt {

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;  // capacity in elements
    void release() {
        if (p) cudaFree(p);
        p = nullptr; n = 0;
    }
    cudaError_t ensure(size_t count) {
        if (count <= n && p) return cudaSuccess;
        release();
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    cudaError_t upload(const std::vector<T>& v, cudaStream_t s) {
        cudaError_t e = ensure(v.size());
        if (e != cudaSuccess || v.empty()) return e;
        return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
    }
};

struct SobolBlob {
    const uint32_t* matrices32;  // [1024*52]
    const uint64_t* vdc;         // [25][52]
    const uint64_t* vdc_inv;     // [26][52]
    uint32_t n_dims, matrix_size, n_vdc, n_vdc_inv;
};

struct CommState;  // NCCL communicator of a multi-process context (wrt_multi.cu)

}  // namespace wrt

struct wrt_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::string err;
    wrt::SobolBlob blob{};

    // scene (the compiled host arrays are dropped after the upload; these facts stay)
    bool have_scene = false;
    size_t n_ops = 0;
    bool has_moving = false;
    uint32_t ref_boxes_loose = 0;
    wrt::DeviceScene ds{};
    wrt::DeviceScene ds_pruned{};  // ds with the pruned program (packet scan, WRT_CULL_TIGHT); == ds when nothing was dropped
    wrt::DevBuf<uint4> d_ops;
    wrt::DevBuf<uint4> d_ops_pruned;
    wrt::DevBuf<uint8_t> d_ppm_in, d_ppm_body;
    wrt::DevBuf<uint32_t> d_ppm_blocks;
    wrt::DevBuf<unsigned long long> d_ppm_offsets;
    wrt::DevBuf<wrt::BoxRef> d_boxes_ref;
    wrt::DevBuf<wrt::BoxTight> d_boxes_tight;
    wrt::DevBuf<wrt::Node2> d_nodes2;
    wrt::DevBuf<wrt::Node4> d_nodes4;
    wrt::DevBuf<uint32_t> d_root4;
    wrt::DevBuf<uint32_t> d_sphere_pc, d_quad_pc;
    wrt::DevBuf<wrt::Node4Q> d_nodes4q;
    wrt::DevBuf<wrt::SphereGeom> d_spheres;
    wrt::DevBuf<wrt::SphereAux> d_sphere_aux;
    wrt::DevBuf<wrt::QuadGeom> d_quads;
    wrt::DevBuf<wrt::Xform> d_xforms;
    wrt::DevBuf<uint32_t> d_xform_chains;
    wrt::DevBuf<wrt::Material> d_materials;
    wrt::DevBuf<wrt::Texture> d_textures;
    wrt::DevBuf<wrt::ImageDesc> d_images;
    wrt::DevBuf<wrt::Light> d_lights;
    wrt::DevBuf<wrt::BoxTight> d_light_boxes;
    wrt::DevBuf<uint32_t> d_sobol_matrices;
    wrt::DevBuf<wrt::SobolLut> d_sobol_lut;
    std::vector<cudaArray_t> arrays;
    std::vector<cudaTextureObject_t> texobjs;

    // render state
    wrt::DevBuf<double> d_accum;
    wrt::DevBuf<double> d_fb;
    wrt::DevBuf<uint8_t> d_rgb8;
    wrt::DevBuf<unsigned long long> d_counters;
    wrt::DevBuf<wrt::PathState> d_wf_paths;      // wavefront engine: path pool, queues, counters
    wrt::DevBuf<uint32_t> d_wf_queues;
    wrt::DevBuf<uint32_t> d_wf_slot_job;
    wrt::DevBuf<uint32_t> d_wf_keys, d_wf_sort_tmp, d_wf_sort_hist;  // ray reordering (wrt_kernels.h: WavefrontArgs::keys)
    wrt::DevBuf<unsigned long long> d_wf_counters;
    unsigned long long* h_wf_counters = nullptr;  // pinned
    cudaStream_t wf_streams[WRT_WF_MAX_PIPELINES] = {};  // [0] aliases `stream`
    cudaEvent_t wf_events[WRT_WF_MAX_PIPELINES] = {};
    uint32_t last_pixels = 0;        // pixels of the last render (this shard)
    bool last_valid = false;
    uint32_t sobol_w = 0, sobol_h = 0;  // resolution lp.sobol was built for
    wrt::LaunchParams lp{};             // the __grid_constant__ argument of this context's launches (rc refilled per call)
    wrt_stats stats{};

    // multi-GPU (wrt_multi.cu)
    wrt::CommState* comm = nullptr;     // wrt_comm_init: this context is rank `comm->rank` of a multi-process job
    wrt::DevBuf<double> d_shard;        // this device's rows, 3 lanes per pixel (what travels over NCCL)
    wrt::DevBuf<double> d_staging;      // root: every member's shard, [member][rows_pad * width * 3]
    uint32_t full_w = 0, full_h = 0;    // root: frame assembled by the last group / sharded render (d_fb, d_rgb8)

    int fail(int code, const std::string& msg) {
        err = msg;
        return code;
    }
    int cuda_fail(cudaError_t e, const char* what) {
        err = std::string(what) + ": " + cudaGetErrorString(e);
        return WRT_E_CUDA;
    }
    int resolve_cull(uint32_t mode) const {  // WRT_CULL_AUTO -> the reference's result at the best speed (include/wrt.h)
        if (mode == WRT_CULL_AUTO) return ref_boxes_loose == 0 ? WRT_CULL_TIGHT : WRT_CULL_REFERENCE;
        return (int)mode;
    }
    void free_images() {
        for (auto t : texobjs) cudaDestroyTextureObject(t);
        for (auto a : arrays) cudaFreeArray(a);
        texobjs.clear();
        arrays.clear();
    }
};


namespace wrt {
// shared by wrt_api.cu and wrt_multi.cu
int bind_device(wrt_ctx* ctx);
int upload_compiled(wrt_ctx* ctx, const CompiledScene& cs, const wrt_scene* scene, double compile_ms);
// `d_out` != NULL: device framebuffer of this shard (no D2H); else `host_fb`
int render_impl(wrt_ctx* ctx, const wrt_camera* cam, const wrt_params* params, void* host_fb, void* d_out, size_t stride);
void comm_release(wrt_ctx* ctx);
}  // namespace wrt
