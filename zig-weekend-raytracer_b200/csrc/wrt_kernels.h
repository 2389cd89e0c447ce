// wrt_kernels.h — host-visible interface of wrt_kernels.cu (launchers + constant blocks).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/wrt.h"

#define WRT_RENDER_BLOCK 128
#ifndef WRT_RENDER_MIN_BLOCKS
#define WRT_RENDER_MIN_BLOCKS 4  // per-lane kernels (traversal stack): <= 128 registers per thread, 16 warps per SM
#endif
#ifndef WRT_PACKET_MIN_BLOCKS
#define WRT_PACKET_MIN_BLOCKS 6  // packet kernel: <= 85 registers, 24 warps per SM (measured best of 3/4/5/6/8 on C2)
#endif
#define WRT_REGROUP_BLOCK 768  // regrouping variant: one block of 24 warps per SM (<= 85 registers)
#define WRT_SYNC_BLOCK 512  // phase-synchronous variant: one block of 16 warps per SM
// Programs up to this many ops are scanned with the warp-uniform packet traversal.
#define WRT_PACKET_MAX_OPS 96u

namespace wrt {

struct DeviceScene;
struct SobolTables;

// Per-render constants (one __constant__ block): the RenderThreadContext of the reference (render.zig:78-103)
// plus the job decomposition.
struct RenderConstants {
    wrt_camera cam;
    double background[3];
    unsigned long long seed;
    unsigned long long total_jobs;  // warp jobs: n_chunks * n_rows_local * n_col_blocks (lane / sync / regroup kernels)
    unsigned long long lane_jobs;   // lane jobs: n_chunks * n_pixels_local (render_kernel)
    uint32_t width, height, spp, max_depth, dof;
    uint32_t row_shard_index, row_shard_count, n_rows_local, n_col_blocks;
    uint32_t sample_begin, sample_end, chunk_size, n_chunks;
    uint32_t n_pixels_local, _pad0;
};

// ---- wavefront engine (wrt_kernels.cu) ----
struct __align__(16) PathState {  // 128 bytes, one L2 line per slot
    double ox, oy, oz, dx, dy, dz;  // ray
    double bx, by, bz;              // throughput
    double lx, ly, lz;              // radiance gathered by this path
    double t;                       // closest hit
    uint32_t hit_pc, hit_xf;
    uint32_t depth_left, sample;    // `sample` = index of the sample in flight
    double time;
};
static_assert(sizeof(PathState) == 128, "PathState must stay one cache line");

enum { WQ_EXTEND0 = 0, WQ_EXTEND1, WQ_REGEN0, WQ_REGEN1, WQ_SURFACE, WQ_METAL, WQ_OTHER, WQ_COUNT };
// per-pipeline counters: [0..WQ_COUNT) queue sizes, [11] read cursor of the persistent extend kernel
enum { WF_CURSOR = 11 };

struct WavefrontArgs {
    PathState* paths;
    uint32_t* queues;             // WQ_COUNT arrays of `capacity`
    unsigned long long* counters; // 16 entries
    double* accum;                // [slot][3]
    uint32_t capacity;            // number of slots of the pool
    uint32_t n_pixels;            // pixels of this shard
    const uint32_t* sobol_matrices;  // WRT_FLAG_SAMPLER_SOBOL: SobolMatrices32 (1024 x 52), else nullptr
    // A JOB is (sample chunk, pixel of the shard) = one accumulator triple, job = chunk * n_pixels + pixel; a SLOT works
    // through one job at a time and draws the next one from a counter when its job's samples are done, so the pool stays
    // full however unevenly the work is spread over the pixels (sky pixels finish their chunk in a few hundred iterations,
    // pixels deep in the scene need thousands).
    uint32_t* slot_job;           // [capacity]: the job a slot is working on
    unsigned long long n_jobs;    // n_chunks * n_pixels
    // The pool runs as independent PIPELINES (own slots, queues and counters, own stream) that share the accumulators, the job
    // cursor and the tallies below: while one pipeline's persistent extend kernel drains its last, longest rays — a few
    // milliseconds during which most of its lanes idle — the other pipeline's kernels fill the SMs.
    unsigned long long* shared;   // WS_* counters, common to all pipelines
    uint32_t job_base;            // the jobs this pipeline's slots start with: [job_base, job_base + capacity)
    // Ray reordering (large trees): secondary rays are bucketed by where they START — the primitive they leave (its position in
    // the program = a spatial order, the reference BVH's DFS order) and the octant of their direction — before the extend
    // kernel draws them, so the lanes of a warp walk neighbouring parts of the tree.  One counting-sort pass per iteration:
    // histogram, scan, scatter, copy back.  sort_buckets == 0: off.
    uint32_t* keys;               // [2][capacity], parallel to the two extend queues
    uint32_t* sort_tmp;           // [capacity]
    uint32_t* sort_hist;          // [sort_buckets]
    uint32_t sort_buckets;
    uint32_t sort_shift;          // key = ((hit op >> sort_shift) << 3) | direction octant
    // phase constants of the persistent extend kernel (0 = the compiled defaults; WRT_WF_NODE_BURST / _LEAF_BURST / _NODE_SHIFT in
    // the environment override them for sweeps)
    uint32_t node_burst, leaf_burst, node_shift;
};
enum { WS_JOB_CURSOR = 0, WS_JOBS_DONE = 1, WS_RAYS = 2, WS_PATHS = 3, WS_STEPS = 4, WS_COUNT = 8 };

struct LaunchParams;
cudaError_t wf_launch_init(const LaunchParams& lp, const WavefrontArgs& A, uint32_t grid, cudaStream_t stream);
cudaError_t wf_launch_iteration(const LaunchParams& lp, const WavefrontArgs& A, const DeviceScene& S, uint32_t cull_mode, bool packet, uint32_t parity,
                                uint32_t grid, uint32_t persist_grid, cudaStream_t stream);
cudaError_t wf_extend_occupancy(int* blocks_per_sm);

cudaError_t launch_render(const LaunchParams& lp, const DeviceScene& S, uint32_t cull_mode, bool packet, uint32_t grid, double* accum, unsigned long long* counters,
                          cudaStream_t stream);
cudaError_t launch_render_regroup(const LaunchParams& lp, const DeviceScene& S, uint32_t cull_mode, uint32_t grid, double* accum, unsigned long long* counters,
                                  cudaStream_t stream);
cudaError_t launch_render_sync(const LaunchParams& lp, const DeviceScene& S, uint32_t cull_mode, bool packet, uint32_t grid, double* accum,
                               unsigned long long* counters, cudaStream_t stream);
cudaError_t render_occupancy(const DeviceScene& S, uint32_t cull_mode, bool packet, int* blocks_per_sm);
cudaError_t launch_resolve(const double* accum, uint32_t n_chunks, uint32_t n_pixels, const double clear[3], int no_clear, double* fb,
                           uint32_t stride_doubles, uint8_t* rgb8, cudaStream_t stream);
cudaError_t launch_encode(const double* fb, uint32_t stride_doubles, uint32_t n_pixels, uint8_t* rgb8, cudaStream_t stream);
// PPM body of an RGB8 frame (writer.zig): per-block byte counts -> offsets[n_blocks + 1] (last = body size) -> text
cudaError_t launch_format_ppm(const uint8_t* rgb, uint32_t n_pixels, uint32_t* block_bytes, unsigned long long* offsets, uint8_t* body,
                              cudaStream_t stream);
uint32_t ppm_block_count(uint32_t n_pixels);
cudaError_t launch_primary_hits(const LaunchParams& lp, const DeviceScene& S, uint32_t cull_mode, uint32_t n_samples, uint32_t* ids, double* ts, uint32_t grid,
                                cudaStream_t stream);
cudaError_t launch_trace_rays(const DeviceScene& S, uint32_t cull_mode, bool packet, const double* origins, const double* dirs, uint64_t n,
                              double tmin, uint32_t* ids, double* ts, double* point, double* normal, double* uv, uint32_t* front_face,
                              uint32_t grid, cudaStream_t stream);
cudaError_t launch_sobol_pixel(const LaunchParams& lp, const uint32_t* cols, const uint32_t* rows, const uint32_t* sidx, uint64_t n, uint64_t* index_out,
                               double* offsets, cudaStream_t stream);
cudaError_t launch_sobol_dimension(const uint32_t* matrices, const uint64_t* index, const uint32_t* dimension, uint64_t n,
                                   uint32_t owen_fast, uint32_t seed, float* out, cudaStream_t stream);
cudaError_t launch_fp64_peak(double* out, uint32_t grid, uint32_t block, uint32_t iters, cudaStream_t stream);
cudaError_t launch_fp32_peak(double* out, uint32_t grid, uint32_t block, uint32_t iters, cudaStream_t stream);

}  // namespace wrt
