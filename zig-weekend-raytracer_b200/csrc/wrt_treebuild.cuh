// wrt_treebuild.cuh — the arithmetic of the ordered traversal's tree build, shared word for word by the host builder
// (wrt_program.cu: build_trees_host) and the device builder (wrt_build.cu: build_trees_device), so that both produce the
// same records byte for byte (tests/test_gpu_build.py compares them).
//
// What is built (the reference's own builder is entity.zig:226-259, a median split of a random axis):
// over the leaf entities of every reference BVH, a binary tree by a binned surface-area heuristic — 3 axes x 16 bins over
// the centroid bounds, the cheapest of the 45 candidate planes, a STABLE partition of the items about it — then collapsed
// to four-wide records.  Closest hit and tie rule are properties of the primitives and their DFS positions in the
// reference topology, not of this tree (wrt_device.cuh, Trav), so any tree over the same leaves returns the same hits.
//
// Determinism rules that make the two builders agree:
//   * every quantity below is a pure function of the SET of items of a segment, except the fallback split (no plane
//     separates the centroids, or the depth cap is reached), which cuts the segment's current order in the middle; the
//     current order is the input order refined by stable partitions, the same on both sides;
//   * binary64 products and sums are never fused (nvcc -fmad=false, gcc -ffp-contract=off) and are written in one order;
//   * record indices follow from (lo, mid, hi, free) alone: a subtree over n items owns n - 1 consecutive records, its
//     root first, then the left subtree's, then the right subtree's;
//   * four-wide records are numbered breadth first per tree (parents in record order, children in slot order).
#pragma once

#include <math.h>
#include <stdint.h>

#include "wrt_device.cuh"

#if defined(__CUDACC__)
#define WRT_HD __host__ __device__ __forceinline__
#else
#define WRT_HD inline
#endif

namespace wrt {

constexpr int kTreeBins = 16;
constexpr uint32_t kTreeSahLevels = 30;  // deeper than this the build halves the current order (bounds the depth)

// a leaf entity of a reference BVH: its op range in the traversal program and its tight box in binary64
struct TreeItem {
    uint32_t start, end;
    double mn[3], mx[3];
};
static_assert(sizeof(TreeItem) == 56, "TreeItem is uploaded as is");

WRT_HD bool tb_valid(const double* mn, const double* mx) { return mn[0] <= mx[0]; }
WRT_HD double tb_centroid(const TreeItem& it, int k) { return tb_valid(it.mn, it.mx) ? 0.5 * (it.mn[k] + it.mx[k]) : 0.0; }
WRT_HD double tb_half_area(const double* mn, const double* mx) {
    if (!tb_valid(mn, mx)) return 0.0;
    const double dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
    const double a = dx * dy, b = dy * dz, c = dz * dx;
    return (a + b) + c;
}
WRT_HD int tb_bin(double c, double base, double scale) {
    const double f = (c - base) * scale;
    int b = (f >= (double)kTreeBins) ? kTreeBins - 1 : ((f > 0.0) ? (int)f : 0);  // also maps a NaN to bin 0
    return b > kTreeBins - 1 ? kTreeBins - 1 : b;
}

// One axis of the split search: bin populations and bin boxes (6 doubles per bin: min xyz, max xyz) -> the cheapest plane
// "after bin b" of this axis, or cost = +inf when no plane has items on both sides.
WRT_HD void tb_sweep_axis(const uint32_t* bin_n, const double* bin_box, double& best_cost, int& best_bin) {
    double right_area[kTreeBins];
    uint32_t right_n[kTreeBins];
    double amn[3], amx[3];
    for (int k = 0; k < 3; ++k) { amn[k] = INFINITY; amx[k] = -INFINITY; }
    uint32_t n = 0;
    for (int b = kTreeBins - 1; b > 0; --b) {
        for (int k = 0; k < 3; ++k) { amn[k] = fmin(amn[k], bin_box[b * 6 + k]); amx[k] = fmax(amx[k], bin_box[b * 6 + 3 + k]); }
        n += bin_n[b];
        right_area[b] = tb_half_area(amn, amx);
        right_n[b] = n;
    }
    for (int k = 0; k < 3; ++k) { amn[k] = INFINITY; amx[k] = -INFINITY; }
    n = 0;
    best_cost = INFINITY;
    best_bin = -1;
    for (int b = 0; b + 1 < kTreeBins; ++b) {
        for (int k = 0; k < 3; ++k) { amn[k] = fmin(amn[k], bin_box[b * 6 + k]); amx[k] = fmax(amx[k], bin_box[b * 6 + 3 + k]); }
        n += bin_n[b];
        if (n == 0 || right_n[b + 1] == 0) continue;
        const double cl = tb_half_area(amn, amx) * (double)n, cr = right_area[b + 1] * (double)right_n[b + 1];
        const double cost = cl + cr;
        if (cost < best_cost) { best_cost = cost; best_bin = b; }
    }
}

// binary32 box for the FP32 culler: padded by 4e-6 of the coordinate magnitude (covers |b*inv| * 2^-22 and the binary64
// rounding of the primitives' own arithmetic) and rounded outwards.
WRT_HD float tb_round_down(double v) {
#if defined(__CUDA_ARCH__)
    return __double2float_rd(v);
#else
    const float f = (float)v;
    return ((double)f > v) ? nextafterf(f, -INFINITY) : f;
#endif
}
WRT_HD float tb_round_up(double v) {
#if defined(__CUDA_ARCH__)
    return __double2float_ru(v);
#else
    const float f = (float)v;
    return ((double)f < v) ? nextafterf(f, INFINITY) : f;
#endif
}
WRT_HD void tb_padded(const double* mn, const double* mx, float* lo, float* hi) {
    if (!tb_valid(mn, mx)) {  // empty subtree: a box nothing can hit
        lo[0] = lo[1] = lo[2] = 1.0f;
        hi[0] = hi[1] = hi[2] = -1.0f;
        return;
    }
    for (int k = 0; k < 3; ++k) {
        const double mag = fmax(1.0, fmax(fabs(mn[k]), fabs(mx[k])));
        const double pad = mag * 4e-6;
        lo[k] = tb_round_down(mn[k] - pad);
        hi[k] = tb_round_up(mx[k] + pad);
    }
}

// Records of the two subtrees of a segment [lo, hi) cut at mid, whose root record is given and whose other hi - lo - 2
// records start at `free`: child record (WRT_NONE for a single item) and the first free record below each child.
struct TreeChildren { uint32_t rec[2], free[2]; };
WRT_HD TreeChildren tb_children(uint32_t lo, uint32_t mid, uint32_t hi, uint32_t free) {
    TreeChildren c;
    uint32_t next = free;
    const uint32_t n[2] = {mid - lo, hi - mid};
    for (int s = 0; s < 2; ++s) {
        if (n[s] == 1) { c.rec[s] = WRT_NONE; c.free[s] = next; }
        else { c.rec[s] = next; c.free[s] = next + 1; next += n[s] - 1; }
    }
    return c;
}
WRT_HD void tb_set_child(Node2& n, int side, const float* lo, const float* hi, uint32_t desc, uint32_t end) {
    if (side == 0) {
        n.lmin[0] = lo[0]; n.lmin[1] = lo[1]; n.lmin[2] = lo[2]; n.l_desc = desc;
        n.lmax[0] = hi[0]; n.lmax[1] = hi[1]; n.lmax[2] = hi[2]; n.l_end = end;
    } else {
        n.rmin[0] = lo[0]; n.rmin[1] = lo[1]; n.rmin[2] = lo[2]; n.r_desc = desc;
        n.rmax[0] = hi[0]; n.rmax[1] = hi[1]; n.rmax[2] = hi[2]; n.r_end = end;
    }
}

// ---- four-wide collapse ---------------------------------------------------------------------------------------------
// A four-wide record starts with the two children of a child-pair record and, while it has a free slot, replaces the inner
// child with the largest box by that child's own two children.  Boxes, leaf ranges and therefore the set of primitives
// reached are those of nodes2; only the fan-out changes.
struct TreeChild4 { float lo[3], hi[3]; uint32_t desc, end; };
WRT_HD float tb_area4(const TreeChild4& c) {
    const float dx = c.hi[0] - c.lo[0], dy = c.hi[1] - c.lo[1], dz = c.hi[2] - c.lo[2];
    if (dx < 0 || dy < 0 || dz < 0) return 0.0f;
    const float a = dx * dy, b = dy * dz, d = dz * dx;
    return (a + b) + d;
}
WRT_HD int tb_children_of(const Node2& n, TreeChild4* out) {
    int m = 0;
    if (n.l_desc != WRT_NONE) {
        for (int k = 0; k < 3; ++k) { out[m].lo[k] = n.lmin[k]; out[m].hi[k] = n.lmax[k]; }
        out[m].desc = n.l_desc; out[m].end = n.l_end;
        ++m;
    }
    if (n.r_desc != WRT_NONE) {
        for (int k = 0; k < 3; ++k) { out[m].lo[k] = n.rmin[k]; out[m].hi[k] = n.rmax[k]; }
        out[m].desc = n.r_desc; out[m].end = n.r_end;
        ++m;
    }
    return m;
}
// the (up to four) children of the four-wide record that stands for child-pair record `rec2`; returns their number
WRT_HD int tb_widen(const Node2* nodes2, uint32_t rec2, TreeChild4* ch) {
    int n = tb_children_of(nodes2[rec2], ch);
    for (;;) {
        if (n >= 4) break;
        int best = -1;
        float best_area = -1.0f;
        for (int i = 0; i < n; ++i)
            if ((ch[i].desc & 0x80000000u) && tb_area4(ch[i]) > best_area) { best = i; best_area = tb_area4(ch[i]); }
        if (best < 0) break;
        TreeChild4 sub[2];
        const int m = tb_children_of(nodes2[ch[best].desc & 0x7FFFFFFFu], sub);
        if (n - 1 + m > 4) break;
        for (int i = best; i + 1 < n; ++i) ch[i] = ch[i + 1];  // erase, keep the order
        --n;
        for (int i = 0; i < m; ++i) ch[n++] = sub[i];
    }
    return n;
}
// the record itself; inner children get the record indices first_child, first_child + 1, ... in slot order.  `ops` is the
// traversal program: a leaf that is one sphere / quad op carries the primitive's record index instead of the range end,
// so the traversal tests it without fetching the op first (one dependent load less per leaf).
WRT_HD Node4 tb_node4(const TreeChild4* ch, int n, uint32_t first_child, const uint4* ops) {
    Node4 r;
    for (int i = 0; i < 4; ++i) {  // empty slot: a box nothing can hit, no child
        r.lox[i] = r.loy[i] = r.loz[i] = 1.0f;
        r.hix[i] = r.hiy[i] = r.hiz[i] = -1.0f;
        r.desc[i] = WRT_NONE;
        r.end[i] = 0;
    }
    uint32_t next = first_child;
    for (int i = 0; i < n; ++i) {
        r.lox[i] = ch[i].lo[0]; r.loy[i] = ch[i].lo[1]; r.loz[i] = ch[i].lo[2];
        r.hix[i] = ch[i].hi[0]; r.hiy[i] = ch[i].hi[1]; r.hiz[i] = ch[i].hi[2];
        if (ch[i].desc & 0x80000000u) {
            r.desc[i] = 0x80000000u | next++;
            r.end[i] = 0;
        } else {
            r.desc[i] = ch[i].desc;
            r.end[i] = ch[i].end;
            const uint4 op = ops[ch[i].desc];
            if (ch[i].end == ch[i].desc + 1 && (op.x == OP_SPHERE || op.x == OP_QUAD) && op.y <= WRT_LEAF_INDEX)
                r.end[i] = WRT_LEAF_PRIM | (op.x == OP_QUAD ? WRT_LEAF_QUAD : 0u) | op.y;
        }
    }
    return r;
}

// ---- quantised four-wide record (Node4Q) -----------------------------------------------------------------------------------
// From a Node4 of a compact_ok scene.  The decoded box of every child contains its binary32 box: lo steps are rounded down,
// hi steps up, and both are checked against the decoded value.  Returns false when an axis cannot be expressed (never for
// finite boxes of sane size); the scene then keeps walking Node4.
WRT_HD bool tb_node4q(const Node4& n, Node4Q& q) {
    double corner[3] = {INFINITY, INFINITY, INFINITY}, far_[3] = {-INFINITY, -INFINITY, -INFINITY};
    const float* lo[3] = {n.lox, n.loy, n.loz};
    const float* hi[3] = {n.hix, n.hiy, n.hiz};
    bool any = false;
    for (int i = 0; i < 4; ++i) {
        if (n.desc[i] == WRT_NONE) continue;
        any = true;
        for (int k = 0; k < 3; ++k) { corner[k] = fmin(corner[k], (double)lo[k][i]); far_[k] = fmax(far_[k], (double)hi[k][i]); }
    }
    q.exps = 0;
    q._pad[0] = q._pad[1] = 0;
    float o[3] = {0.0f, 0.0f, 0.0f};
    for (int k = 0; k < 3; ++k) { q.qlo[k] = 0; q.qhi[k] = 0; }
    for (int k = 0; k < 3 && any; ++k) {
        if (!(far_[k] >= corner[k]) || !(fabs(corner[k]) < 1e30) || !(fabs(far_[k]) < 1e30)) return false;
        o[k] = (float)corner[k];  // exact: the corner is one of the binary32 lo values
        int e = 0;
        const double ext = far_[k] - corner[k];
        if (ext > 0.0) { (void)frexp(ext / 255.0, &e); }  // ext / 255 = m * 2^e, m in [0.5, 1): step 2^e >= ext / 255
        else e = -120;
        if (e < -120) e = -120;
        for (;;) {
            if (e + 127 > 254) return false;
            const double step = ldexp(1.0, e);
            bool ok = true;
            uint32_t wlo = 0, whi = 0;
            for (int i = 0; i < 4 && ok; ++i) {
                if (n.desc[i] == WRT_NONE) { wlo |= 255u << (8 * i); continue; }  // empty slot: inverted (lo 255, hi 0)
                double a = floor(((double)lo[k][i] - corner[k]) / step);
                if (corner[k] + a * step > (double)lo[k][i]) a -= 1.0;
                if (a < 0.0) a = 0.0;  // (the corner itself: decoded == lo)
                double b = ceil(((double)hi[k][i] - corner[k]) / step);
                if (corner[k] + b * step < (double)hi[k][i]) b += 1.0;
                if (b > 255.0 || a > 255.0) { ok = false; break; }
                wlo |= (uint32_t)a << (8 * i);
                whi |= (uint32_t)b << (8 * i);
            }
            if (ok) { q.qlo[k] = wlo; q.qhi[k] = whi; q.exps |= (uint32_t)(e + 127) << (8 * k); break; }
            ++e;
        }
    }
    q.ox = o[0]; q.oy = o[1]; q.oz = o[2];
    for (int i = 0; i < 4; ++i) {
        if (n.desc[i] == WRT_NONE) q.word[i] = WRT_NONE;
        else if (n.desc[i] & 0x80000000u) q.word[i] = n.desc[i];
        else q.word[i] = n.end[i] & 0x7FFFFFFFu;  // single-primitive leaf (compact_ok): kind + record index
    }
    return true;
}

}  // namespace wrt
