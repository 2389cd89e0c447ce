// wrt_build.cu — the trees of the ordered traversal built on the device (the reference's builder is the
// recursive, pointer-allocating median split of src/entity.zig:226-259, O(N log^2 N) on one host thread).
//
// Same algorithm and arithmetic as the host builder (wrt_treebuild.cuh spells out both; wrt_program.cu: HostTreeBuilder),
// re-scheduled for the GPU: the recursion becomes one pass per tree LEVEL over all segments of that level.
//
//   items stay where they are; `order[pos]` names the item at position pos, a segment is a range of positions.
//   per level:  bounds of the centroids -> 3 x 16 bins (population, box) -> cheapest plane           (per segment)
//               flag = item goes left; exclusive scan of the flags; stable scatter of `order`          (per position)
//               child boxes, leaf descriptors -> the child-pair record; child segments of the next level
//   segments of more than 1024 items accumulate bounds and bins with global 64-bit min / max atomics on order-preserving
//   keys of the doubles (at most N / 1024 of them exist); smaller ones are handled by one warp each with the bins in shared
//   memory.  Everything the records depend on is a min, a max, a count or a comparison of those — independent of the order
//   in which threads arrive — so the result equals the host's bit for bit.
//   Then the four-wide collapse, breadth first: per level, count the inner children of every new record, scan, emit.
//
// 2^20 primitives: ~25 levels, HBM traffic ~ 25 x 5 passes x 2^20 x O(100 B) = a few GB -> milliseconds, against ~1 s on
// the host threads.  Not a tensor-core shape; the rules that matter are coalesced per-position passes and keeping the
// per-segment state small (the bins of big segments: 2.7 KB each, <= N / 1024 segments).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "wrt_program.h"
#include "wrt_treebuild.cuh"

namespace wrt {
namespace {

constexpr uint32_t kSmallMax = 1024;  // segments up to this many items are handled by one warp
constexpr int kBinsPerSeg = 3 * kTreeBins;

// order-preserving 64-bit key of a double (min / max of keys = min / max of the doubles; -0 sorts below +0)
__host__ __device__ __forceinline__ unsigned long long key_of(double d) {
#if defined(__CUDA_ARCH__)
    const unsigned long long b = (unsigned long long)__double_as_longlong(d);
#else
    unsigned long long b;
    std::memcpy(&b, &d, 8);
#endif
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double double_of(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k ^ 0x8000000000000000ull) : ~k;
    return __longlong_as_double((long long)b);
}

struct Seg {
    uint32_t lo, hi;     // positions
    uint32_t rec, free;  // this segment's record; first of the hi - lo - 2 records below it
    uint32_t large;      // slot in the big-segment accumulators, WRT_NONE for a warp-handled segment
};
struct Split {
    int axis, bin;        // axis < 0: halve the current order
    uint32_t mid;
    uint32_t child[2];    // segment index at the next level, WRT_NONE = that side is a single item
    uint32_t rec[2];      // child records (WRT_NONE for a single item)
    double base, scale;   // of the chosen axis
};

struct LevelCounters { uint32_t n_segs, n_large; };

struct BuildArrays {  // device pointers of one build
    const TreeItem* items;
    uint32_t* order[2];
    uint32_t* seg_of[2];
    uint32_t* flag;
    uint32_t* pre;
    Seg* segs[2];
    Split* split;
    unsigned long long* sidebox;  // [seg][2][6] keys: min xyz, max xyz
    uint32_t* leafdesc;           // [seg][2][2]: start, end of a single-item side
    unsigned long long* lcb;      // [large][6]
    uint32_t* lbin_n;             // [large][48]
    unsigned long long* lbin_box; // [large][48][6]
    LevelCounters* counters;      // [2]: this level / next level
    Node2* nodes2;
};

__device__ __forceinline__ unsigned long long empty_min_key() { return key_of(INFINITY); }
__device__ __forceinline__ unsigned long long empty_max_key() { return key_of(-INFINITY); }

__global__ void init_large_kernel(BuildArrays A, uint32_t n_large) {
    const uint32_t total = n_large * (uint32_t)(kBinsPerSeg * 6);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
        A.lbin_box[i] = (i % 6 < 3) ? empty_min_key() : empty_max_key();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_large * (uint32_t)kBinsPerSeg; i += gridDim.x * blockDim.x) A.lbin_n[i] = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_large * 6u; i += gridDim.x * blockDim.x)
        A.lcb[i] = (i % 6 < 3) ? empty_min_key() : empty_max_key();
}

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
    return __shfl_xor_sync(0xFFFFFFFFu, v, m);
}

// centroid bounds of the big segments: one thread per position, one set of atomics per warp when the warp sits in one segment
__global__ void large_bounds_kernel(BuildArrays A, int cur, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t L = WRT_NONE;
    unsigned long long kmin[3], kmax[3];
    for (int k = 0; k < 3; ++k) { kmin[k] = empty_min_key(); kmax[k] = empty_max_key(); }
    if (i < n) {
        const uint32_t s = A.seg_of[cur][i];
        if (s != WRT_NONE) {
            L = A.segs[cur][s].large;
            if (L != WRT_NONE) {
                const TreeItem it = A.items[A.order[cur][i]];
                for (int k = 0; k < 3; ++k) kmin[k] = kmax[k] = key_of(tb_centroid(it, k));
            }
        }
    }
    const unsigned active = __ballot_sync(0xFFFFFFFFu, L != WRT_NONE);
    if (active == 0) return;
    const uint32_t L0 = __shfl_sync(0xFFFFFFFFu, L, __ffs(active) - 1);
    const bool uniform = __all_sync(0xFFFFFFFFu, L == WRT_NONE || L == L0);
    if (uniform) {
        for (int k = 0; k < 3; ++k)
            for (int m = 16; m > 0; m >>= 1) {
                const unsigned long long a = shfl_xor_u64(kmin[k], m), b = shfl_xor_u64(kmax[k], m);
                kmin[k] = a < kmin[k] ? a : kmin[k];
                kmax[k] = b > kmax[k] ? b : kmax[k];
            }
        if ((threadIdx.x & 31) == 0)
            for (int k = 0; k < 3; ++k) { atomicMin(&A.lcb[L0 * 6 + k], kmin[k]); atomicMax(&A.lcb[L0 * 6 + 3 + k], kmax[k]); }
    } else if (L != WRT_NONE) {
        for (int k = 0; k < 3; ++k) { atomicMin(&A.lcb[L * 6 + k], kmin[k]); atomicMax(&A.lcb[L * 6 + 3 + k], kmax[k]); }
    }
}

// bins of the big segments.  A block whose positions all lie in one segment (the rule on the first levels) accumulates in
// shared memory and adds its 48 bins to the segment's once.
__global__ void __launch_bounds__(256) large_bins_kernel(BuildArrays A, int cur, uint32_t n) {
    __shared__ uint32_t s_n[kBinsPerSeg];
    __shared__ unsigned long long s_box[kBinsPerSeg * 6];
    __shared__ uint32_t s_first;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t L = WRT_NONE;
    TreeItem it;
    if (i < n) {
        const uint32_t s = A.seg_of[cur][i];
        if (s != WRT_NONE) {
            L = A.segs[cur][s].large;
            if (L != WRT_NONE) it = A.items[A.order[cur][i]];
        }
    }
    if (threadIdx.x == 0) s_first = WRT_NONE;
    __syncthreads();
    if (L != WRT_NONE) s_first = L;  // any one of them
    __syncthreads();
    const uint32_t L0 = s_first;
    if (L0 == WRT_NONE) return;
    const bool uniform = __syncthreads_and(L == WRT_NONE || L == L0);
    if (uniform) {
        for (int b = threadIdx.x; b < kBinsPerSeg; b += blockDim.x) s_n[b] = 0;
        for (int b = threadIdx.x; b < kBinsPerSeg * 6; b += blockDim.x) s_box[b] = (b % 6 < 3) ? empty_min_key() : empty_max_key();
        __syncthreads();
    }
    if (L != WRT_NONE) {
        const bool valid = tb_valid(it.mn, it.mx);
        for (int axis = 0; axis < 3; ++axis) {
            const double mn = double_of(A.lcb[L * 6 + axis]), mx = double_of(A.lcb[L * 6 + 3 + axis]);
            const double ext = mx - mn;
            if (!(ext > 0.0)) continue;
            const double scale = (double)kTreeBins / ext;
            const int b = axis * kTreeBins + tb_bin(tb_centroid(it, axis), mn, scale);
            uint32_t* cnt = uniform ? &s_n[b] : &A.lbin_n[L * kBinsPerSeg + b];
            unsigned long long* box = uniform ? &s_box[b * 6] : &A.lbin_box[((size_t)L * kBinsPerSeg + b) * 6];
            atomicAdd(cnt, 1u);
            if (valid)
                for (int k = 0; k < 3; ++k) { atomicMin(&box[k], key_of(it.mn[k])); atomicMax(&box[3 + k], key_of(it.mx[k])); }
        }
    }
    if (uniform) {
        __syncthreads();
        for (int b = threadIdx.x; b < kBinsPerSeg; b += blockDim.x) {
            const uint32_t c = s_n[b];
            if (c == 0) continue;
            atomicAdd(&A.lbin_n[L0 * kBinsPerSeg + b], c);
            unsigned long long* box = &A.lbin_box[((size_t)L0 * kBinsPerSeg + b) * 6];
            for (int k = 0; k < 3; ++k) { atomicMin(&box[k], s_box[b * 6 + k]); atomicMax(&box[3 + k], s_box[b * 6 + 3 + k]); }
        }
    }
}

// Given a segment's centroid bounds and bins: choose the plane, lay out the children, start the next level's segments.
// Runs in one thread per segment (bins in global memory for a big segment, in the warp's shared memory otherwise).
__device__ void decide_split(const BuildArrays& A, int cur, uint32_t s, uint32_t level, const double* cmn, const double* cmx,
                             const uint32_t* bin_n, const unsigned long long* bin_keys) {
    const Seg seg = A.segs[cur][s];
    double best_cost = INFINITY, best_base = 0.0, best_scale = 0.0;
    int best_axis = -1, best_bin = 0;
    double box[kTreeBins * 6];
    if (level <= kTreeSahLevels) {
        for (int axis = 0; axis < 3; ++axis) {
            const double ext = cmx[axis] - cmn[axis];
            if (!(ext > 0.0)) continue;
            uint32_t cnt[kTreeBins];
            for (int b = 0; b < kTreeBins; ++b) {
                cnt[b] = bin_n[axis * kTreeBins + b];
                for (int k = 0; k < 6; ++k) box[b * 6 + k] = double_of(bin_keys[(axis * kTreeBins + b) * 6 + k]);
            }
            double cost;
            int bin;
            tb_sweep_axis(cnt, box, cost, bin);
            if (bin >= 0 && cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = bin; best_base = cmn[axis]; best_scale = (double)kTreeBins / ext; }
        }
    }
    Split sp;
    sp.axis = best_axis; sp.bin = best_bin; sp.base = best_base; sp.scale = best_scale;
    unsigned long long side[12];
    for (int k = 0; k < 12; ++k) side[k] = (k % 6 < 3) ? empty_min_key() : empty_max_key();
    if (best_axis >= 0) {
        uint32_t left = 0;
        for (int b = 0; b < kTreeBins; ++b) {
            const int sd = b <= best_bin ? 0 : 1;
            if (sd == 0) left += bin_n[best_axis * kTreeBins + b];
            const unsigned long long* kb = &bin_keys[(best_axis * kTreeBins + b) * 6];
            for (int k = 0; k < 3; ++k) {
                side[sd * 6 + k] = kb[k] < side[sd * 6 + k] ? kb[k] : side[sd * 6 + k];
                side[sd * 6 + 3 + k] = kb[3 + k] > side[sd * 6 + 3 + k] ? kb[3 + k] : side[sd * 6 + 3 + k];
            }
        }
        sp.mid = seg.lo + left;
    } else {
        sp.mid = seg.lo + (seg.hi - seg.lo) / 2;  // the scatter pass accumulates the side boxes of such a segment
    }
    for (int k = 0; k < 12; ++k) A.sidebox[(size_t)s * 12 + k] = side[k];
    const TreeChildren ch = tb_children(seg.lo, sp.mid, seg.hi, seg.free);
    for (int sd = 0; sd < 2; ++sd) {
        sp.rec[sd] = ch.rec[sd];
        sp.child[sd] = WRT_NONE;
        if (ch.rec[sd] == WRT_NONE) continue;
        Seg c;
        c.lo = sd == 0 ? seg.lo : sp.mid;
        c.hi = sd == 0 ? sp.mid : seg.hi;
        c.rec = ch.rec[sd];
        c.free = ch.free[sd];
        c.large = (c.hi - c.lo > kSmallMax) ? atomicAdd(&A.counters[1].n_large, 1u) : WRT_NONE;
        const uint32_t idx = atomicAdd(&A.counters[1].n_segs, 1u);
        A.segs[cur ^ 1][idx] = c;
        sp.child[sd] = idx;
    }
    A.split[s] = sp;
}

__global__ void large_split_kernel(BuildArrays A, int cur, uint32_t n_segs, uint32_t level) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_segs) return;
    const uint32_t L = A.segs[cur][s].large;
    if (L == WRT_NONE) return;
    double cmn[3], cmx[3];
    for (int k = 0; k < 3; ++k) { cmn[k] = double_of(A.lcb[L * 6 + k]); cmx[k] = double_of(A.lcb[L * 6 + 3 + k]); }
    decide_split(A, cur, s, level, cmn, cmx, &A.lbin_n[L * kBinsPerSeg], &A.lbin_box[(size_t)L * kBinsPerSeg * 6]);
}

// one warp per small segment: bounds, bins (shared memory), split
constexpr int kSmallWarps = 4;
__global__ void __launch_bounds__(kSmallWarps * 32) small_split_kernel(BuildArrays A, int cur, uint32_t n_segs, uint32_t level) {
    __shared__ uint32_t s_n[kSmallWarps][kBinsPerSeg];
    __shared__ unsigned long long s_box[kSmallWarps][kBinsPerSeg * 6];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t s = blockIdx.x * kSmallWarps + w;
    if (s >= n_segs) return;
    const Seg seg = A.segs[cur][s];
    if (seg.large != WRT_NONE) return;
    unsigned long long kmin[3], kmax[3];
    for (int k = 0; k < 3; ++k) { kmin[k] = empty_min_key(); kmax[k] = empty_max_key(); }
    for (uint32_t p = seg.lo + lane; p < seg.hi; p += 32) {
        const TreeItem it = A.items[A.order[cur][p]];
        for (int k = 0; k < 3; ++k) {
            const unsigned long long c = key_of(tb_centroid(it, k));
            kmin[k] = c < kmin[k] ? c : kmin[k];
            kmax[k] = c > kmax[k] ? c : kmax[k];
        }
    }
    double cmn[3], cmx[3];
    for (int k = 0; k < 3; ++k) {
        for (int m = 16; m > 0; m >>= 1) {
            const unsigned long long a = shfl_xor_u64(kmin[k], m), b = shfl_xor_u64(kmax[k], m);
            kmin[k] = a < kmin[k] ? a : kmin[k];
            kmax[k] = b > kmax[k] ? b : kmax[k];
        }
        cmn[k] = double_of(kmin[k]);
        cmx[k] = double_of(kmax[k]);
    }
    for (int b = lane; b < kBinsPerSeg; b += 32) s_n[w][b] = 0;
    for (int b = lane; b < kBinsPerSeg * 6; b += 32) s_box[w][b] = (b % 6 < 3) ? empty_min_key() : empty_max_key();
    __syncwarp();
    if (level <= kTreeSahLevels) {
        for (uint32_t p = seg.lo + lane; p < seg.hi; p += 32) {
            const TreeItem it = A.items[A.order[cur][p]];
            const bool valid = tb_valid(it.mn, it.mx);
            for (int axis = 0; axis < 3; ++axis) {
                const double ext = cmx[axis] - cmn[axis];
                if (!(ext > 0.0)) continue;
                const double scale = (double)kTreeBins / ext;
                const int b = axis * kTreeBins + tb_bin(tb_centroid(it, axis), cmn[axis], scale);
                atomicAdd(&s_n[w][b], 1u);
                if (valid)
                    for (int k = 0; k < 3; ++k) { atomicMin(&s_box[w][b * 6 + k], key_of(it.mn[k])); atomicMax(&s_box[w][b * 6 + 3 + k], key_of(it.mx[k])); }
            }
        }
    }
    __syncwarp();
    if (lane == 0) decide_split(A, cur, s, level, cmn, cmx, s_n[w], s_box[w]);
}

__global__ void flag_kernel(BuildArrays A, int cur, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = A.seg_of[cur][i];
    uint32_t f = 0;
    if (s != WRT_NONE) {
        const Split sp = A.split[s];
        if (sp.axis >= 0) {
            const TreeItem it = A.items[A.order[cur][i]];
            f = tb_bin(tb_centroid(it, sp.axis), sp.base, sp.scale) <= sp.bin ? 1u : 0u;
        } else {
            f = i < sp.mid ? 1u : 0u;
        }
    }
    A.flag[i] = f;
}

// stable scatter of `order` about each segment's plane; an item that ends up alone on its side hands over its op range
__global__ void scatter_kernel(BuildArrays A, int cur, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = A.seg_of[cur][i];
    const uint32_t item = A.order[cur][i];
    if (s == WRT_NONE) {  // already a leaf of its tree
        A.order[cur ^ 1][i] = item;
        A.seg_of[cur ^ 1][i] = WRT_NONE;
        return;
    }
    const Seg seg = A.segs[cur][s];
    const Split sp = A.split[s];
    const uint32_t left_before = A.pre[i] - A.pre[seg.lo];
    const uint32_t f = A.flag[i];
    const uint32_t pos = f ? seg.lo + left_before : sp.mid + (i - seg.lo - left_before);
    const int sd = f ? 0 : 1;
    A.order[cur ^ 1][pos] = item;
    A.seg_of[cur ^ 1][pos] = sp.child[sd];
    if (sp.child[sd] == WRT_NONE || sp.axis < 0) {
        const TreeItem it = A.items[item];
        if (sp.child[sd] == WRT_NONE) {
            A.leafdesc[(size_t)s * 4 + sd * 2] = it.start;
            A.leafdesc[(size_t)s * 4 + sd * 2 + 1] = it.end;
        }
        if (sp.axis < 0 && tb_valid(it.mn, it.mx)) {
            unsigned long long* box = &A.sidebox[(size_t)s * 12 + sd * 6];
            for (int k = 0; k < 3; ++k) { atomicMin(&box[k], key_of(it.mn[k])); atomicMax(&box[3 + k], key_of(it.mx[k])); }
        }
    }
}

__global__ void record_kernel(BuildArrays A, int cur, uint32_t n_segs) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_segs) return;
    const Seg seg = A.segs[cur][s];
    const Split sp = A.split[s];
    Node2 r;
    for (int sd = 0; sd < 2; ++sd) {
        double mn[3], mx[3];
        for (int k = 0; k < 3; ++k) { mn[k] = double_of(A.sidebox[(size_t)s * 12 + sd * 6 + k]); mx[k] = double_of(A.sidebox[(size_t)s * 12 + sd * 6 + 3 + k]); }
        float lo[3], hi[3];
        tb_padded(mn, mx, lo, hi);
        if (sp.rec[sd] == WRT_NONE) tb_set_child(r, sd, lo, hi, A.leafdesc[(size_t)s * 4 + sd * 2], A.leafdesc[(size_t)s * 4 + sd * 2 + 1]);
        else tb_set_child(r, sd, lo, hi, 0x80000000u | sp.rec[sd], 0u);
    }
    A.nodes2[seg.rec] = r;
}

__global__ void start_kernel(BuildArrays A, uint32_t n, Seg root) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { A.segs[0][0] = root; A.counters[0].n_segs = 1; A.counters[0].n_large = root.large != WRT_NONE ? 1u : 0u; A.counters[1].n_segs = 0; A.counters[1].n_large = 0; }
    if (i < n) { A.order[0][i] = i; A.seg_of[0][i] = 0; }
}
__global__ void next_level_kernel(BuildArrays A) {
    A.counters[0] = A.counters[1];
    A.counters[1].n_segs = 0;
    A.counters[1].n_large = 0;
}

// ---- exclusive scan of 32-bit counts (1024 per block, recursive over the block sums) -----------------------------------
constexpr int kScanThreads = 256, kScanPerThread = 4, kScanBlock = kScanThreads * kScanPerThread;
__global__ void __launch_bounds__(kScanThreads) scan_block_kernel(const uint32_t* in, uint32_t* out, uint32_t* block_sums, uint32_t n) {
    __shared__ uint32_t warp_sums[kScanThreads / 32];
    const uint32_t base = blockIdx.x * kScanBlock + threadIdx.x * kScanPerThread;
    uint32_t v[kScanPerThread], sum = 0;
    for (int k = 0; k < kScanPerThread; ++k) { v[k] = (base + k < n) ? in[base + k] : 0u; sum += v[k]; }
    uint32_t incl = sum;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int m = 1; m < 32; m <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, m); if (lane >= m) incl += t; }
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
        uint32_t ws = lane < kScanThreads / 32 ? warp_sums[lane] : 0u;
        for (int m = 1; m < 32; m <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, ws, m); if (lane >= m) ws += t; }
        if (lane < kScanThreads / 32) warp_sums[lane] = ws;  // inclusive
    }
    __syncthreads();
    uint32_t excl = incl - sum + (w ? warp_sums[w - 1] : 0u);
    for (int k = 0; k < kScanPerThread; ++k) { if (base + k < n) out[base + k] = excl; excl += v[k]; }
    if (threadIdx.x == kScanThreads - 1 && block_sums) block_sums[blockIdx.x] = warp_sums[kScanThreads / 32 - 1];
}
__global__ void scan_add_kernel(uint32_t* out, const uint32_t* block_offsets, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += block_offsets[i / kScanBlock];
}

struct Scanner {
    std::vector<uint32_t*> sums;  // block sums per recursion level (scanned in place)
    std::vector<uint32_t> cap;
    cudaError_t reserve(uint32_t n) {
        uint32_t m = n;
        while (m > (uint32_t)kScanBlock) {
            m = (m + kScanBlock - 1) / kScanBlock;
            uint32_t* p = nullptr;
            cudaError_t e = cudaMalloc(&p, (size_t)m * sizeof(uint32_t));
            if (e != cudaSuccess) return e;
            sums.push_back(p);
            cap.push_back(m);
        }
        return cudaSuccess;
    }
    void run(const uint32_t* in, uint32_t* out, uint32_t n, cudaStream_t st, size_t depth = 0) {
        if (n == 0) return;
        const uint32_t blocks = (n + kScanBlock - 1) / kScanBlock;
        uint32_t* bs = blocks > 1 ? sums[depth] : nullptr;
        scan_block_kernel<<<blocks, kScanThreads, 0, st>>>(in, out, bs, n);
        if (blocks > 1) {
            run(bs, bs, blocks, st, depth + 1);
            scan_add_kernel<<<(n + 255) / 256, 256, 0, st>>>(out, bs, n);
        }
    }
    ~Scanner() { for (uint32_t* p : sums) cudaFree(p); }
};

// ---- four-wide collapse -----------------------------------------------------------------------------------------------
struct Work4 { uint32_t rec2, rec4; };
__global__ void collapse_count_kernel(const Node2* nodes2, const Work4* work, uint32_t n_work, uint32_t* counts) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_work) return;
    TreeChild4 ch[4];
    const int n = tb_widen(nodes2, work[i].rec2, ch);
    uint32_t inner = 0;
    for (int k = 0; k < n; ++k) inner += (ch[k].desc & 0x80000000u) ? 1u : 0u;
    counts[i] = inner;
}
__global__ void collapse_emit_kernel(const Node2* nodes2, const uint4* ops, const Work4* work, uint32_t n_work, const uint32_t* offsets,
                                     uint32_t next_base, Node4* nodes4, Work4* next_work) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_work) return;
    TreeChild4 ch[4];
    const int n = tb_widen(nodes2, work[i].rec2, ch);
    const uint32_t first_child = next_base + offsets[i];
    nodes4[work[i].rec4] = tb_node4(ch, n, first_child, ops);
    uint32_t j = 0;
    for (int k = 0; k < n; ++k)
        if (ch[k].desc & 0x80000000u) { next_work[offsets[i] + j] = Work4{ch[k].desc & 0x7FFFFFFFu, first_child + j}; ++j; }
}

struct Freer {  // frees what was allocated when the build leaves, on every path
    std::vector<void*> ptrs;
    ~Freer() { for (void* p : ptrs) cudaFree(p); }
    template <class T> cudaError_t alloc(T** p, size_t count) {
        cudaError_t e = cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back((void*)*p);
        return e;
    }
};

}  // namespace

bool want_device_build(const CompiledScene& cs) {
    const char* env = std::getenv("WRT_DEVICE_BUILD");
    if (env && (env[0] == '0' || env[0] == '1')) return env[0] == '1';
    // The device build runs one tree after the other, ~10 launches and one host synchronisation per level: it pays for a big
    // tree (below 32 768 leaves the host build takes a few milliseconds and ~20 levels of launches cost as much), not for a
    // scene that is a crowd of small ones.
    size_t largest = 0, n_trees = 0;
    for (const TreeInput& r : cs.tree_inputs) {
        largest = std::max(largest, r.items.size());
        n_trees += r.items.size() >= 2 ? 1 : 0;
    }
    return largest >= 32768 && n_trees <= 64;
}

#define BCU(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) { err = std::string("device tree build: ") + cudaGetErrorString(e_); return e_ == cudaErrorMemoryAllocation ? WRT_E_NOMEM : WRT_E_CUDA; } \
    } while (0)

int build_trees_device(CompiledScene& cs, int device, std::string& err, double* build_ms) {
    const bool trace = std::getenv("WRT_TRACE_BUILD") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!trace) return;
        const auto t = std::chrono::steady_clock::now();
        std::fprintf(stderr, "wrt trace: device build: %s %.2f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
        t_last = t;
    };
    BCU(cudaSetDevice(device));
    cudaStream_t st = nullptr;
    BCU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } stream_guard{st};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    BCU(cudaEventCreate(&ev0));
    BCU(cudaEventCreate(&ev1));
    struct EventGuard { cudaEvent_t a, b; ~EventGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } event_guard{ev0, ev1};

    const bool rebuild = !keep_reference_trees();
    size_t max_items = 0, extra_records = 0;
    for (const TreeInput& r : cs.tree_inputs) {
        max_items = std::max(max_items, r.items.size());
        if (rebuild && r.items.size() >= 2) extra_records += r.items.size() - 2;
    }
    const size_t base_records = cs.nodes2.size();
    const size_t total_records = base_records + extra_records;
    // a four-wide record stands for one reachable child-pair record: n - 1 of a rebuilt tree, at most all base records otherwise
    size_t max_records4 = 0;
    for (const TreeInput& r : cs.tree_inputs) max_records4 += (rebuild && r.items.size() >= 2) ? r.items.size() - 1 : 1;
    if (!rebuild) max_records4 = base_records;
    max_records4 = std::max<size_t>(max_records4, 1);
    if (total_records >= 0x7FFFFFFFull || max_items >= 0x7FFFFFFFull) { err = "device tree build: too many records"; return WRT_E_LIMIT; }

    Freer mem;
    BuildArrays A;
    std::memset(&A, 0, sizeof A);
    const uint32_t n_max = (uint32_t)max_items;
    const size_t max_segs = std::max<size_t>(n_max / 2 + 1, 1), max_large = n_max / kSmallMax + 2;
    TreeItem* d_items = nullptr;
    uint4* d_ops = nullptr;
    Node4* d_nodes4 = nullptr;
    Work4* d_work[2] = {nullptr, nullptr};
    uint32_t* d_counts = nullptr;
    uint32_t* d_offsets = nullptr;
    BCU(mem.alloc(&d_items, n_max));
    for (int k = 0; k < 2; ++k) { BCU(mem.alloc(&A.order[k], n_max)); BCU(mem.alloc(&A.seg_of[k], n_max)); BCU(mem.alloc(&A.segs[k], max_segs)); }
    BCU(mem.alloc(&A.flag, n_max));
    BCU(mem.alloc(&A.pre, n_max));
    BCU(mem.alloc(&A.split, max_segs));
    BCU(mem.alloc(&A.sidebox, max_segs * 12));
    BCU(mem.alloc(&A.leafdesc, max_segs * 4));
    BCU(mem.alloc(&A.lcb, max_large * 6));
    BCU(mem.alloc(&A.lbin_n, max_large * kBinsPerSeg));
    BCU(mem.alloc(&A.lbin_box, max_large * kBinsPerSeg * 6));
    BCU(mem.alloc(&A.counters, 2));
    BCU(mem.alloc(&A.nodes2, total_records));
    BCU(mem.alloc(&d_ops, cs.ops.size()));
    BCU(mem.alloc(&d_nodes4, max_records4));
    for (int k = 0; k < 2; ++k) BCU(mem.alloc(&d_work[k], max_records4));
    BCU(mem.alloc(&d_counts, max_records4));
    BCU(mem.alloc(&d_offsets, max_records4));
    A.items = d_items;
    Scanner scan;
    BCU(scan.reserve((uint32_t)std::max<size_t>(max_records4, n_max)));
    LevelCounters* h_counters = nullptr;
    BCU(cudaMallocHost(&h_counters, 2 * sizeof(LevelCounters)));
    struct PinnedGuard { void* p; ~PinnedGuard() { cudaFreeHost(p); } } pinned_guard{h_counters};
    lap("allocations");

    // base records (reference topology) first; the rebuilt trees' other records behind them are all written by the build
    BCU(cudaMemcpyAsync(A.nodes2, cs.nodes2.data(), base_records * sizeof(Node2), cudaMemcpyHostToDevice, st));
    BCU(cudaMemcpyAsync(d_ops, cs.ops.data(), cs.ops.size() * sizeof(uint4), cudaMemcpyHostToDevice, st));
    double ms_total = 0.0;
    if (trace) cudaStreamSynchronize(st);
    lap("resize + H2D of the base records and the program");

    // ---- binned SAH, one tree after the other, one pass per level ------------------------------------------------------
    uint32_t free_rec = (uint32_t)base_records;
    if (rebuild) {
        for (TreeInput& r : cs.tree_inputs) {
            const uint32_t n = (uint32_t)r.items.size();
            if (n < 2) continue;  // a single leaf: the reference's record is already minimal
            BCU(cudaMemcpyAsync(d_items, r.items.data(), (size_t)n * sizeof(TreeItem), cudaMemcpyHostToDevice, st));
            BCU(cudaEventRecord(ev0, st));
            Seg root;
            root.lo = 0; root.hi = n; root.rec = r.record; root.free = free_rec;
            root.large = n > kSmallMax ? 0u : WRT_NONE;
            const uint32_t pos_blocks = (n + 255) / 256;
            start_kernel<<<pos_blocks, 256, 0, st>>>(A, n, root);
            uint32_t n_segs = 1, n_large = root.large != WRT_NONE ? 1u : 0u, level = 1;
            int cur = 0;
            while (n_segs > 0) {
                if (n_large > 0) {
                    init_large_kernel<<<std::min<uint32_t>((n_large * kBinsPerSeg * 6 + 255) / 256, 1184u), 256, 0, st>>>(A, n_large);
                    large_bounds_kernel<<<pos_blocks, 256, 0, st>>>(A, cur, n);
                    large_bins_kernel<<<pos_blocks, 256, 0, st>>>(A, cur, n);
                    large_split_kernel<<<(n_segs + 63) / 64, 64, 0, st>>>(A, cur, n_segs, level);
                }
                if (n_segs > n_large) small_split_kernel<<<(n_segs + kSmallWarps - 1) / kSmallWarps, kSmallWarps * 32, 0, st>>>(A, cur, n_segs, level);
                flag_kernel<<<pos_blocks, 256, 0, st>>>(A, cur, n);
                scan.run(A.flag, A.pre, n, st);
                scatter_kernel<<<pos_blocks, 256, 0, st>>>(A, cur, n);
                record_kernel<<<(n_segs + 255) / 256, 256, 0, st>>>(A, cur, n_segs);
                BCU(cudaMemcpyAsync(h_counters, A.counters, 2 * sizeof(LevelCounters), cudaMemcpyDeviceToHost, st));
                next_level_kernel<<<1, 1, 0, st>>>(A);
                BCU(cudaStreamSynchronize(st));
                n_segs = h_counters[1].n_segs;
                n_large = h_counters[1].n_large;
                if (n_segs > max_segs || n_large > max_large) { err = "device tree build: segment bookkeeping overflow"; return WRT_E_CUDA; }
                cur ^= 1;
                ++level;
            }
            BCU(cudaEventRecord(ev1, st));
            BCU(cudaEventSynchronize(ev1));
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, ev0, ev1);
            ms_total += ms;
            const uint32_t depth = level - 1;  // levels that held a segment = records on the longest root-to-leaf chain
            cs.max_nesting = std::max(cs.max_nesting, r.nest + depth);
            free_rec += n - 2;
        }
    }

    lap("SAH levels (incl. H2D of the items)");
    // ---- four-wide collapse, breadth first per tree ---------------------------------------------------------------------
    cs.root4.assign(total_records, WRT_NONE);
    uint32_t n4 = 0;
    BCU(cudaEventRecord(ev0, st));
    for (const TreeInput& r : cs.tree_inputs) {
        cs.root4[r.record] = n4;
        const Work4 first{r.record, n4};
        ++n4;
        BCU(cudaMemcpyAsync(d_work[0], &first, sizeof first, cudaMemcpyHostToDevice, st));
        BCU(cudaStreamSynchronize(st));  // `first` is a local
        uint32_t n_work = 1;
        int cur = 0;
        while (n_work > 0) {
            const uint32_t blocks = (n_work + 127) / 128;
            collapse_count_kernel<<<blocks, 128, 0, st>>>(A.nodes2, d_work[cur], n_work, d_counts);
            scan.run(d_counts, d_offsets, n_work, st);
            collapse_emit_kernel<<<blocks, 128, 0, st>>>(A.nodes2, d_ops, d_work[cur], n_work, d_offsets, n4, d_nodes4, d_work[cur ^ 1]);
            uint32_t last[2];
            BCU(cudaMemcpyAsync(&last[0], d_offsets + (n_work - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            BCU(cudaMemcpyAsync(&last[1], d_counts + (n_work - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            BCU(cudaStreamSynchronize(st));
            const uint32_t produced = last[0] + last[1];
            if ((size_t)n4 + produced > max_records4) { err = "device tree build: four-wide records overflow"; return WRT_E_CUDA; }
            n4 += produced;
            n_work = produced;
            cur ^= 1;
        }
    }
    BCU(cudaEventRecord(ev1, st));
    BCU(cudaEventSynchronize(ev1));
    {
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, ev0, ev1);
        ms_total += ms;
    }
    BCU(cudaGetLastError());
    lap("four-wide collapse");

    cs.nodes4.resize(n4);
    cs.nodes2.resize(total_records);
    if (extra_records) BCU(cudaMemcpyAsync(cs.nodes2.data() + base_records, A.nodes2 + base_records, extra_records * sizeof(Node2), cudaMemcpyDeviceToHost, st));
    for (const TreeInput& r : cs.tree_inputs)  // the roots of the rebuilt trees live among the base records
        BCU(cudaMemcpyAsync(&cs.nodes2[r.record], A.nodes2 + r.record, sizeof(Node2), cudaMemcpyDeviceToHost, st));
    if (n4) BCU(cudaMemcpyAsync(cs.nodes4.data(), d_nodes4, (size_t)n4 * sizeof(Node4), cudaMemcpyDeviceToHost, st));
    BCU(cudaStreamSynchronize(st));
    lap("D2H of the records");
    finish_trees_after_device_build(cs);
    lap("stack bound");
    if (build_ms) *build_ms = ms_total;
    return WRT_OK;
}

int compile_scene_for_device(const wrt_scene* scene, CompiledScene& out, std::string& err, int device, double* tree_ms, bool* on_device) {
    int rc = compile_scene(scene, out, err, /*defer_trees=*/true);
    if (rc != WRT_OK) return rc;
    const bool dev = want_device_build(out);
    if (on_device) *on_device = dev;
    try {
        if (dev) {
            double ms = 0.0;
            rc = build_trees_device(out, device, err, &ms);
            if (tree_ms) *tree_ms = ms;
        } else {
            const auto t0 = std::chrono::steady_clock::now();
            build_trees_host(out);
            if (tree_ms) *tree_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        }
    } catch (const std::bad_alloc&) {
        err = "out of host memory while building the trees";
        return WRT_E_NOMEM;
    } catch (const std::exception& e) {
        err = std::string("tree build failed: ") + e.what();
        return WRT_E_INVALID;
    }
    for (TreeInput& r : out.tree_inputs) std::vector<TreeItem>().swap(r.items);  // consumed
    return rc;
}

}  // namespace wrt
