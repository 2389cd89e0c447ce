// wrt_device.cuh — device-side data layout and arithmetic of the B200 render back end.
//
// Everything here follows the behavioural spec of the reference's hot path; each
// function cites the reference lines it replaces.  The translation unit is compiled with -fmad=false so that
// binary64 operations stay unfused in source order (the reference emits no FMA); the places
// where exactness is irrelevant (conservative culling) call fma() explicitly.
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <type_traits>

#include "../../include/wrt.h"
#include "wrt_kernels.h"

namespace wrt {


// 256-bit loads (sm_100: LDG.E.256).  The per-lane traversals read 64- and 128-byte records with every lane on its own
// line, so the L1 pipeline spends one tag cycle per lane and INSTRUCTION whatever the width: half the instructions per
// record is half that load.  Addresses must be 32-byte aligned (all records here are).
__device__ __forceinline__ void ldg256(const void* p, float4& a, float4& b) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
__device__ __forceinline__ void ldg256(const void* p, double2& a, double2& b) {
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
}
// streaming form (evict-first) for the path pool
__device__ __forceinline__ void ldcs256(const void* p, double2& a, double2& b) {
    asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// Device scene layout.  All records are 16-byte aligned and read with 128-bit loads.
// ---------------------------------------------------------------------------------------------------------

// Traversal program: the entity tree in DFS pre-order (= the reference's visiting order: bvh_node.hit tests
// left then right, entity.zig:286-303; collections scan in insertion order, :351-367).  Stack-less: a culled
// node jumps to `skip`, the first op after its subtree.
enum OpKind : uint32_t {
    OP_NODE = 0,        // x=kind, y=box index, z=skip pc
    OP_SPHERE = 1,      // y=sphere index, z=material, w=prim id
    OP_QUAD = 2,        // y=quad index,   z=material, w=prim id
    OP_PUSH_TRANSLATE = 3,  // y=xform index
    OP_PUSH_ROTATE_Y = 4,   // y=xform index
    OP_POP = 5,         // y=parent xform index (WRT_NONE = world)
    OP_END = 6,
    OP_NODE_TIGHT_ONLY = 7  // like OP_NODE but inserted by the compiler (instance bounds); REFERENCE culling ignores it
};

struct __align__(16) BoxRef {   // the reference's cached AABB.min/max, x and y only (aabb.zig:80-101 as executed)
    double min_x, min_y, max_x, max_y;
};
struct __align__(16) BoxTight { // recomputed conservative box in binary32, rounded outwards and padded (wrt_program.cu)
    float min_x, min_y, min_z, _p0;
    float max_x, max_y, max_z, _p1;
};
// Child-pair record of a bvh_node (one 64-byte line): both children's binary32 boxes plus where each child lives in the
// program.  desc: bit 31 set = the child is itself a bvh_node, low bits = its record index; otherwise desc = first op of the
// child's op range and *_end = one past its last op.  r_desc == WRT_NONE: single-child node (span == 1, entity.zig:231-233).
struct __align__(16) Node2 {
    float lmin[3]; uint32_t l_desc;
    float lmax[3]; uint32_t l_end;
    float rmin[3]; uint32_t r_desc;
    float rmax[3]; uint32_t r_end;
};
// Four-wide record of the ordered traversal (one 128-byte line): the binary SAH tree collapsed two levels at a time
// (wrt_program.cu: build_nodes4), children's binary32 boxes as structure of arrays.  desc: bit 31 set = inner record index,
// otherwise first op of a leaf range ending at `end`; WRT_NONE = empty slot.
// `end` of a leaf that is ONE sphere / quad op: flag + kind + index into spheres[] / quads[] (the range is [desc, desc + 1))
#define WRT_LEAF_PRIM 0x80000000u
#define WRT_LEAF_QUAD 0x40000000u
#define WRT_LEAF_INDEX 0x3FFFFFFFu
struct __align__(16) Node4 {
    float lox[4], loy[4], loz[4], hix[4], hiy[4], hiz[4];
    uint32_t desc[4];
    uint32_t end[4];
};
// Quantised four-wide record (64 bytes, two 256-bit loads): the children's boxes as 8-bit offsets from the record's own
// box, lo rounded down and hi rounded up so that the decoded box contains the binary32 box of Node4; child words as in the
// compact stack entries (inner: bit 31 + record index, leaf: kind bit 30 + primitive record index, WRT_NONE: empty).  Only
// for scenes with DeviceScene::compact_ok.  Half the L1 tag cycles per step and half the record bytes in L1 / L2.
struct __align__(32) Node4Q {
    float ox, oy, oz;      // min corner of the union of the children's boxes
    uint32_t exps;         // byte k: biased binary32 exponent of the step of axis k (step = 2^(e - 127))
    uint32_t qlo[3];       // per axis: byte i = child i's lo, in steps from the corner (rounded down)
    uint32_t qhi[3];       // ... hi (rounded up)
    uint32_t word[4];
    uint32_t _pad[2];
};
static_assert(sizeof(Node4Q) == 64, "Node4Q is two 256-bit loads");

struct __align__(16) SphereGeom { // entity.zig:536-537
    double cx, cy, cz, radius;
};
struct __align__(16) SphereAux {  // moving spheres only (entity.zig:541-542)
    double mx, my, mz;
    uint32_t is_moving, _pad;
};
struct __align__(16) QuadGeom {   // entity.zig:431-441; 256 bytes = two 128-byte lines
    // line 0 — everything a traversal reads: the plane, the start point and the two functionals of the interior PRE-test,
    // alpha = w . (planar x v) = planar . (v x w) and beta = w . (u x planar) = planar . (w x u) (wrt_program.cu forms them in
    // binary64).  One line per quad in L1 / L2 instead of three (the 2^20-primitive scene holds 2^19 quads).
    double nx, ny, nz, offset;    // unit normal, D
    double sx, sy, sz, area;      // start point, area
    double ax, ay, az, bx, by, bz;
    double _l0[2];
    // line 1 — the basis: hit records with texture coordinates, light sampling, and the exact interior test within 2e-6 of an edge
    double ux, uy, uz, _p0;       // basis.u
    double vx, vy, vz, _p1;       // basis.v
    double wx, wy, wz, _p2;       // basis.w
    double _l1[4];
};
static_assert(sizeof(QuadGeom) == 256, "QuadGeom must stay two cache lines");
struct __align__(16) Xform {      // Translate / RotateY chain (entity.zig:68-205)
    double a, b, c;               // translate: offset xyz; rotate_y: sin, cos, 0
    uint32_t kind;                // OP_PUSH_TRANSLATE / OP_PUSH_ROTATE_Y
    uint32_t parent;              // enclosing xform or WRT_NONE
};
struct __align__(16) Material {   // material.zig:79-226
    double ar, ag, ab, param;     // metal albedo; fuzz / refraction index
    uint32_t kind, texture, _p0, _p1;
};
struct __align__(16) Texture {    // texture.zig:33-119
    double r, g, b, inv_scale;
    uint32_t kind, even, odd, image;
};
struct ImageDesc {                // image.zig:23-36
    cudaTextureObject_t tex;      // uchar4 texels, point sampled, unnormalised coordinates
    uint32_t width, height;
};
struct Light {                    // Scene.lights children (entity.zig:371-386)
    uint32_t kind;                // WRT_ENT_SPHERE / WRT_ENT_QUAD / other (pdf 0, direction (1,0,0))
    uint32_t index;
};

struct DeviceScene {
    const uint4* ops;
    const BoxRef* boxes_ref;
    const BoxTight* boxes_tight;
    const Node2* nodes2;
    const Node4* nodes4;           // four-wide records (ordered traversal)
    const uint32_t* root4;         // per box index: the Node4 record of the tree rooted there (WRT_NONE if it is not a root)
    const SphereGeom* spheres;
    const SphereAux* sphere_aux;
    const QuadGeom* quads;
    const Xform* xforms;
    const uint32_t* xform_chains;  // WRT_MAX_XFORM_DEPTH ids per xform, root -> leaf, WRT_NONE padded
    const Material* materials;
    const Texture* textures;
    const ImageDesc* images;
    const Light* lights;
    const BoxTight* light_boxes;  // parallel to lights
    uint32_t n_ops, n_lights, has_lights, has_moving;
    const Node4Q* nodes4q;      // compact_ok scenes: nodes4 in quantised form (same indices)
    const uint32_t* sphere_pc;  // per sphere / quad record: the op that tests it (compact stack entries, see TravCompactStack)
    const uint32_t* quad_pc;
    uint32_t compact_ok;   // one tree of single-primitive leaves, no transforms, every primitive tested by exactly one op
    uint32_t use_ordered;  // ordered traversal allowed (its worst-case stack use fits WRT_STACK_DEPTH)
    uint32_t use_wide;     // the ordered traversal walks the four-wide records (large trees) instead of the child-pair records
    uint32_t _pad1, _pad2;
};

struct SobolLut {                  // byte-indexed folds of the matrices below (global memory, L1 resident, 23 KB)
    uint64_t vdc[2][256];          // XOR of VdC[m-1][8k + j] over the set bits j of byte k of the sample index
    uint64_t vdc_inv[7][256];      // same for VdCInv[m-1] over the bytes of b
    uint32_t dim1[7][256];         // same for SobolMatrices32[52 + ...] over the bytes of the Sobol index
};
struct SobolTables {               // live part of sobolmatrices.zig for one resolution
    uint64_t vdc[52];              // VdCSobolMatrices[m-1]
    uint64_t vdc_inv[52];          // VdCSobolMatricesInv[m-1]
    uint32_t dim0[52];             // SobolMatrices32[0*52 ..]
    uint32_t dim1[52];             // SobolMatrices32[1*52 ..]
    uint32_t log2_scale, scale;
    uint32_t b_bytes, _pad;        // bytes of b = (px << m | py) ^ delta that can be non-zero
    const SobolLut* lut;
    // Sample s -> s + 1 of one pixel: index and sample bits are GF(2)-linear in (s, pixel), and s ^ (s + 1) = 2^(k+1) - 1
    // with k = number of trailing ones of s, so the bits of dims 0/1 advance by one XOR with inc[k] (built at upload).
    uint32_t inc0[32], inc1[32];
};
// What every launch that renders or samples carries as ONE __grid_constant__ argument (1.9 KB of the 4 KB parameter space):
// the launch's own constant bank, so nothing is shared between contexts, streams or host threads on a device.
struct LaunchParams {
    RenderConstants rc;
    SobolTables sobol;
};

// ---------------------------------------------------------------------------------------------------------
// Vector arithmetic (math.zig)
// ---------------------------------------------------------------------------------------------------------
struct d3 { double x, y, z; };

__device__ __forceinline__ d3 mk(double x, double y, double z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ d3 operator+(d3 a, d3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ d3 operator-(d3 a, d3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ d3 operator*(d3 a, d3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ d3 operator*(d3 a, double s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ d3 operator/(d3 a, double s) { return mk(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ d3 operator-(d3 a) { return mk(-a.x, -a.y, -a.z); }
// math.zig:243-246: (x0*y0 + x1*y1) + x2*y2
__device__ __forceinline__ double dot(d3 u, d3 v) { return (u.x * v.x + u.y * v.y) + u.z * v.z; }
// math.zig:214-229
__device__ __forceinline__ d3 cross(d3 u, d3 v) {
    return mk(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
}
__device__ __forceinline__ double length(d3 u) { return sqrt(dot(u, u)); }              // math.zig:254-256
__device__ __forceinline__ d3 normalize(d3 u) { return u * (1.0 / length(u)); }          // math.zig:262-264
__device__ __forceinline__ d3 reflect(d3 v, d3 n) { return v - n * (2.0 * dot(v, n)); }  // math.zig:270-272
// math.zig:274-279
__device__ __forceinline__ d3 refract(d3 vn, d3 n, double index) {
    double cos_theta = fmin(dot(-vn, n), 1.0);
    d3 perp = (vn + n * cos_theta) * index;
    d3 par = n * (-sqrt(fabs(1.0 - dot(perp, perp))));
    return perp + par;
}
__device__ __forceinline__ double clamp01(double v) { return fmax(0.0, fmin(v, 1.0)); }

struct Onb { d3 u, v, w; };  // math.zig:58-96
__device__ __forceinline__ Onb onb_init(d3 n) {  // math.zig:65-73
    Onb b;
    b.w = normalize(n);
    d3 a = (fabs(b.w.y) > 0.9) ? mk(1, 0, 0) : mk(0, 1, 0);
    b.u = normalize(cross(b.w, a));
    b.v = cross(b.w, b.u);
    return b;
}
__device__ __forceinline__ d3 onb_transform(const Onb& b, d3 p) {  // math.zig:89-95
    return (b.u * p.x + b.v * p.y) + b.w * p.z;
}

#define WRT_PI 3.14159265358979323846264338327950288

// ---------------------------------------------------------------------------------------------------------
// Counter-based RNG: Philox4x32-10 keyed by the render seed, counted by (pixel, sample, draw).  The oracle
// carries the same definition (oracle/wro_rng.h) so CPU and device paths consume identical numbers.
// ---------------------------------------------------------------------------------------------------------
// Draw layout: draws 0,1 = lens sample, 2 = ray time; bounce b owns draws 4+4b .. 4+4b+3 =
// {mixture choice | Fresnel uniform, light pick, u1, u2}.  One Philox block yields the two draws of an even/odd pair.
// With `sobol` set (WRT_FLAG_SAMPLER_SOBOL) the same draw slots are filled from the pixel sample's Sobol point instead:
// draw j is sampleDimension(2 + j mod 1022) of the sample's global Sobol index — the reference's get1D / get2D counter that
// starts at dimension 2 and wraps at 1024 (sampler.zig:203-220) — Owen-scrambled per dimension (sampler.zig:236-247).
struct Rng {
    uint32_t k0, k1, pixel, sample;
    const uint32_t* sobol = nullptr;  // SobolMatrices32, 1024 x 52 (global memory); nullptr = Philox
    uint64_t sobol_index = 0;         // global Sobol index of (pixel, sample): sobolIntervalToIndex
};

__device__ __forceinline__ void philox_block(const Rng& r, uint32_t block, uint64_t& lo, uint64_t& hi) {
    uint32_t c0 = r.pixel, c1 = r.sample, c2 = block, c3 = 0u;
    uint32_t k0 = r.k0, k1 = r.k1;
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    lo = (uint64_t)c0 | ((uint64_t)c1 << 32);
    hi = (uint64_t)c2 | ((uint64_t)c3 << 32);
}
__device__ __forceinline__ double bits_to_unit(uint64_t bits) { return (double)(bits >> 11) * 0x1p-53; }  // 53 bits in [0,1)
__device__ __forceinline__ void rng_pair(const Rng& r, uint32_t block, double& a, double& b);  // draws (2*block, 2*block + 1)
__device__ __forceinline__ uint32_t pick_index(double u, uint32_t n) {  // intRangeAtMost(0, n-1) stand-in
    uint32_t i = (uint32_t)(u * (double)n);
    return i < n ? i : n - 1;
}

// ---------------------------------------------------------------------------------------------------------
// Sobol pixel sampling (sampler.zig:197-201, 222-234, 249-264, 267-298) — bit exact
// ---------------------------------------------------------------------------------------------------------
// The XOR-of-matrix-columns loops of the reference (one iteration per index bit) are GF(2) matrix-vector products;
// SobolLut holds the same matrices folded into byte-indexed tables (built per resolution at upload), so a product is
// one lookup per index byte instead of eight predicated XORs.  Results are identical bit for bit (XOR is linear).
__device__ __forceinline__ uint64_t sobol_interval_to_index(const SobolTables& T, uint64_t sample_idx, uint32_t px, uint32_t py) {
    const uint32_t m = T.log2_scale;
    if (m == 0) return sample_idx;
    const SobolLut* __restrict__ L = T.lut;
    uint64_t index = sample_idx << (m << 1);
    uint64_t delta = __ldg(&L->vdc[0][sample_idx & 255u]) ^ __ldg(&L->vdc[1][(sample_idx >> 8) & 255u]);
    uint64_t rest = sample_idx >> 16;  // more than 65535 samples per pixel: finish bit by bit (sampler.zig:279-286)
    for (uint32_t c = 16; rest > 0; rest >>= 1, ++c)
        if (rest & 1) delta ^= T.vdc[c];
    const uint64_t b = ((((uint64_t)px) << m) | (uint64_t)py) ^ delta;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        if (k < (int)T.b_bytes) index ^= __ldg(&L->vdc_inv[k][(b >> (8 * k)) & 255u]);  // T.b_bytes is warp-uniform
    }
    return index;
}
__device__ __forceinline__ float sobol_sample_bits_to_float(uint32_t v) {
    float vf = __uint2float_rn(v);                       // @floatFromInt, round to nearest even
    return fminf(__fmul_rn(vf, 0x1p-32f), 0x1.fffffep-1f);  // sampler.zig:262-263
}
// the sample bits of dimensions 0 and 1 for a Sobol index (sobolSample with the noop randomiser, before the float conversion)
__device__ __forceinline__ void sobol_pixel_bits(const SobolTables& T, uint64_t index, uint32_t& v0, uint32_t& v1) {
    // dimension 0 is the van der Corput sequence: its matrix is the bit reversal of the low 32 index bits
    // (SobolMatrices32[i] = 0x80000000 >> i for i < 32, 0 above; checked when the tables are loaded)
    v0 = __brev((uint32_t)index);
    const SobolLut* __restrict__ L = T.lut;
    v1 = 0;
#pragma unroll
    for (int k = 0; k < 7; ++k) v1 ^= __ldg(&L->dim1[k][(index >> (8 * k)) & 255u]);
}
// advance the bits from sample s_prev to s_prev + 1 of the same pixel
__device__ __forceinline__ void sobol_pixel_bits_next(const SobolTables& T, uint32_t s_prev, uint32_t& v0, uint32_t& v1) {
    const int k = __ffs(~s_prev) - 1;  // trailing ones of s_prev
    v0 ^= T.inc0[k & 31];
    v1 ^= T.inc1[k & 31];
}
__device__ __forceinline__ void sobol_bits_to_offsets(const SobolTables& T, uint32_t v0, uint32_t v1, uint32_t px, uint32_t py, double& ox, double& oy) {
    const double one_minus_eps = (double)0x1.fffffep-1f;
    double rx = (double)sobol_sample_bits_to_float(v0) * (double)T.scale - (double)px;
    double ry = (double)sobol_sample_bits_to_float(v1) * (double)T.scale - (double)py;
    ox = fmax(0.0, fmin(rx, one_minus_eps));  // std.math.clamp, sampler.zig:229-231
    oy = fmax(0.0, fmin(ry, one_minus_eps));
}
__device__ __forceinline__ void sobol_pixel_2d(const SobolTables& T, uint64_t index, uint32_t px, uint32_t py, double& ox, double& oy) {
    uint32_t v0, v1;
    sobol_pixel_bits(T, index, v0, v1);
    const double one_minus_eps = (double)0x1.fffffep-1f;
    double rx = (double)sobol_sample_bits_to_float(v0) * (double)T.scale - (double)px;
    double ry = (double)sobol_sample_bits_to_float(v1) * (double)T.scale - (double)py;
    ox = fmax(0.0, fmin(rx, one_minus_eps));  // std.math.clamp, sampler.zig:229-231
    oy = fmax(0.0, fmin(ry, one_minus_eps));
}

__device__ __forceinline__ uint32_t owen_fast_apply(uint32_t seed, uint32_t v) {  // sampler.zig:39-53
    v = __brev(v);
    v ^= v * 0x3d20adeau;
    v += seed;
    v *= (seed >> 16) | 1u;
    v ^= v * 0x05526c56u;
    v ^= v * 0x53a22864u;
    return __brev(v);
}
__device__ __forceinline__ uint32_t murmur2_u32(uint32_t v, uint32_t seed) {  // std.hash.Murmur2_32.hashUint32WithSeed
    const uint32_t m = 0x5bd1e995u;
    uint32_t h1 = seed ^ 4u;
    uint32_t k1 = v * m;
    k1 ^= k1 >> 24;
    k1 *= m;
    h1 *= m;
    h1 ^= k1;
    h1 ^= h1 >> 13;
    h1 *= m;
    h1 ^= h1 >> 15;
    return h1;
}

// sampleDimension (sampler.zig:236-247) with the owen_fast randomiser: XOR of the dimension's matrix columns over the set
// bits of the index, then the Laine-Karras hash seeded per dimension through Murmur2 (seed = low word of the render seed).
__device__ __forceinline__ double sobol_dimension_unit(const Rng& r, uint32_t dimension) {
    uint32_t v = 0;
    uint64_t a = r.sobol_index;
    for (uint32_t k = dimension * 52u; a != 0; a >>= 1, ++k)
        if (a & 1) v ^= __ldg(r.sobol + k);
    v = owen_fast_apply(murmur2_u32(dimension, r.k0), v);
    return (double)sobol_sample_bits_to_float(v);
}
// draws (2*block, 2*block + 1) as uniforms in [0, 1)
__device__ __forceinline__ void rng_pair(const Rng& r, uint32_t block, double& a, double& b) {
    if (r.sobol) {  // folds away in the kernels that never set it (the megakernels)
        a = sobol_dimension_unit(r, 2u + (2u * block) % 1022u);
        b = sobol_dimension_unit(r, 2u + (2u * block + 1u) % 1022u);
        return;
    }
    uint64_t lo, hi;
    philox_block(r, block, lo, hi);
    a = bits_to_unit(lo);
    b = bits_to_unit(hi);
}

// ---------------------------------------------------------------------------------------------------------
// Closest hit
// ---------------------------------------------------------------------------------------------------------
struct Ray { d3 o, d; double time; };

struct HitRecord {  // hitrecord.zig:6-14
    d3 point, normal;
    double t, u, v;
    uint32_t material, prim_id, front_face, is_sphere;
    d3 sphere_outward;  // outward normal in object space, kept so the sphere UV can be computed lazily
};

// Object-space ray of transform context `xf` (chain root -> leaf), exactly as Translate.hit / RotateY.hit build
// it (entity.zig:93-109, 169-205).  Chains are at most WRT_MAX_XFORM_DEPTH deep (checked at upload).
#define WRT_MAX_XFORM_DEPTH 8
__device__ __forceinline__ void apply_xform(const Xform& X, d3& o, d3& d) {
    if (X.kind == OP_PUSH_TRANSLATE) {
        o = o - mk(X.a, X.b, X.c);
    } else {
        const double sn = X.a, cs = X.b;
        o = mk(cs * o.x - sn * o.z, o.y, sn * o.x + cs * o.z);
        d = mk(cs * d.x - sn * d.z, d.y, sn * d.x + cs * d.z);
    }
}
__device__ __forceinline__ void ray_in_xform(const DeviceScene& S, uint32_t xf, d3 wo, d3 wd, d3& o, d3& d) {
    o = wo; d = wd;
    if (xf == WRT_NONE) return;
    const uint32_t* chain = S.xform_chains + (size_t)xf * WRT_MAX_XFORM_DEPTH;
    for (int k = 0; k < WRT_MAX_XFORM_DEPTH; ++k) {
        const uint32_t id = __ldg(chain + k);
        if (id == WRT_NONE) break;
        apply_xform(S.xforms[id], o, d);
    }
}

template <int CULL>
struct Culler;

// The reference's AABB.hit as executed (aabb.zig:80-101, math.zig:186-190): x and y only, each on its own,
// true divisions, MaxMult on tmax.
template <>
struct Culler<WRT_CULL_REFERENCE> {
    d3 o, d;
    template <bool FAST = false>
    __device__ __forceinline__ void set_ray(d3 ro, d3 rd) { o = ro; d = rd; }
    __device__ __forceinline__ bool pass(const DeviceScene& S, uint32_t box, double tmin, double tmax) const {
        const double2* p = reinterpret_cast<const double2*>(S.boxes_ref + box);
        double2 mn = __ldg(p), mx = __ldg(p + 1);
        double t0x = (mn.x - o.x) / d.x, t1x = (mx.x - o.x) / d.x;
        double t0y = (mn.y - o.y) / d.y, t1y = (mx.y - o.y) / d.y;
        double lox = fmax(fmin(t0x, t1x), tmin), hix = fmin(fmax(t0x, t1x), tmax) * 1.0000000000000004;
        double loy = fmax(fmin(t0y, t1y), tmin), hiy = fmin(fmax(t0y, t1y), tmax) * 1.0000000000000004;
        return (hix > lox) && (hiy > loy);
    }
};

// Fast path: proper 3-axis slab intersection on boxes recomputed from the primitives.  Culling only has to be
// conservative, never exact, so it runs in binary32 on the full-rate FP32 pipe (the FP64 pipe has half the lanes and no
// single-instruction min/max): boxes are rounded outwards and padded at upload, the ray carries an absolute error bound
// per axis, and the exit distance gets a relative slack.  Error budget: inv = rcp(fl(d)) is off
// by <= 3 * 2^-24 relative (2^-24 for the rounding of d, up to 2^-23 for the approximate reciprocal of the packet scan), o*inv and the FMA round once each, so a
// slab distance is off by <= (|o*inv| + |b*inv|) * 2^-22 + |t| * 2^-21; the margins used are 2x (origin term, `err`), 16x
// (box term, the 4e-6 padding) and 4x (relative slack on the exit distance) of that.
template <>
struct Culler<WRT_CULL_TIGHT> {
    float inv_x, inv_y, inv_z;   // 1 / d
    float oi_x, oi_y, oi_z;      // o / d
    float err;                   // max_k |o_k / d_k| * 2^-21: absolute slab-distance error of this ray (the largest axis' bound serves all three)
    // FAST: one MUFU.RCP (relative error <= 2^-23; together with the rounding of f, 3 * 2^-24, inside the budget above)
    // instead of the IEEE-rounded reciprocal and its denormal slow path.  Measured: +2.4 % in the packet kernel (which sits
    // at its register cap and runs the set-up with all lanes), -9 % in the per-lane kernel — so only the packet scan asks
    // for it.
    template <bool FAST>
    static __device__ __forceinline__ float safe_inv(double v) {
        float f = (float)v;
        // a zero / denormal component would make inv infinite and b*inv - o*inv an inf - inf NaN
        if (!(fabsf(f) >= 1e-20f)) f = copysignf(1e-20f, __double2hiint(v) < 0 ? -1.0f : 1.0f);
        if (FAST) {
            float r;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(f));
            return r;
        }
        return __frcp_rn(f);
    }
    template <bool FAST = false>
    __device__ __forceinline__ void set_ray(d3 ro, d3 rd) {
        inv_x = safe_inv<FAST>(rd.x); inv_y = safe_inv<FAST>(rd.y); inv_z = safe_inv<FAST>(rd.z);
        oi_x = (float)ro.x * inv_x; oi_y = (float)ro.y * inv_y; oi_z = (float)ro.z * inv_z;
        err = fmaxf(fmaxf(fabsf(oi_x), fabsf(oi_y)), fabsf(oi_z)) * 4.8e-7f;
    }
    __device__ __forceinline__ bool pass(const DeviceScene& S, uint32_t box, double tmin, double tmax) const {
        return pass_box(S.boxes_tight + box, tmin, tmax);
    }
    __device__ __forceinline__ bool pass_box(const BoxTight* __restrict__ b, double tmin, double tmax) const {
        const float4* p = reinterpret_cast<const float4*>(b);
        const float4 lo4 = __ldg(p), hi4 = __ldg(p + 1);  // (min.x, min.y, min.z, -), (max.x, max.y, max.z, -)
        const float ax = fmaf(lo4.x, inv_x, -oi_x), bx = fmaf(hi4.x, inv_x, -oi_x);
        const float ay = fmaf(lo4.y, inv_y, -oi_y), by = fmaf(hi4.y, inv_y, -oi_y);
        const float az = fmaf(lo4.z, inv_z, -oi_z), bz = fmaf(hi4.z, inv_z, -oi_z);
        const float t_lo = __double2float_rd(tmin), t_hi = __double2float_ru(tmax);
        const float lo = fmaxf(fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)) - err, t_lo);
        const float hi = fminf(fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)) + err, t_hi);
        return hi * 1.000002f + 1e-30f >= lo;  // relative slack for the rounding of inv (|t| * 2^-21 on either side)
    }
    // same test on an explicit box; also returns the (conservative) entry distance for near-first ordering
    __device__ __forceinline__ bool entry(float mnx, float mny, float mnz, float mxx, float mxy, float mxz, float t_lo, float t_hi,
                                          float& t_entry) const {
        const float ax = fmaf(mnx, inv_x, -oi_x), bx = fmaf(mxx, inv_x, -oi_x);
        const float ay = fmaf(mny, inv_y, -oi_y), by = fmaf(mxy, inv_y, -oi_y);
        const float az = fmaf(mnz, inv_z, -oi_z), bz = fmaf(mxz, inv_z, -oi_z);
        const float lo = fmaxf(fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)) - err, t_lo);
        const float hi = fminf(fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)) + err, t_hi);
        t_entry = lo;
        return hi * 1.000002f + 1e-30f >= lo;
    }
};

// x / y where x == 0 is frequent: a bounce ray leaving an axis-aligned wall has a numerator of exactly 0 against that
// wall's plane, and a light-sampled direction below the horizon zeroes the throughput (the reference keeps tracing such
// paths).  IEEE gives +-0 for finite non-zero y, but nvcc's division sends a zero numerator down its ~100-instruction
// slow path (6 % of all executed instructions before this).  Same result bits, none of the cost.
__device__ __forceinline__ double div_zero_aware(double x, double y) {
    if (x == 0.0 && y == y && y != 0.0 && fabs(y) != CUDART_INF) return ((__double_as_longlong(x) ^ __double_as_longlong(y)) < 0) ? -0.0 : 0.0;
    return x / y;
}

// Plane distance t = num / denom of QuadEntity.hit (entity.zig:482-486), formed only when it can lie in [tmin, tmax].
// The reference always divides and then rejects t outside the interval; most quads a ray meets are behind it or beyond the
// closest hit so far, and an IEEE binary64 division is ~20 instructions here (10 % of the packet kernel's instructions were
// this division).  Rejections that need no quotient, both safe against the reference's rounded compare:
//   tmin > 0 and (num == 0 or sign(num) != sign(denom))      =>  t <= 0 < tmin
//   |num| > tmax * |denom| * (1 + 1e-12)                      =>  |t| > tmax by far more than the quotient's half ulp
// Otherwise the quotient is formed (num != 0 when tmin > 0, so the zero-numerator slow path is never taken) and the
// reference's own compare decides.  Returns false when the quad is rejected; t is valid only when it returns true.
__device__ __forceinline__ bool quad_plane_t(double num, double denom, double tmin, double tmax, double& t) {
    if (tmin > 0.0) {
        if (num == 0.0 || ((__double2hiint(num) ^ __double2hiint(denom)) < 0)) return false;
        if (fabs(num) > tmax * fabs(denom) * 1.000000000001) return false;
        t = num / denom;
    } else {
        t = div_zero_aware(num, denom);
    }
    return (tmin <= t) && (t <= tmax);
}

// isInteriorPoint (entity.zig:527-541) decides on alpha = w . (planar x v) and beta = w . (u x planar), 28 binary64
// operations after 6 loads.  Both are linear in `planar`, so alpha' = planar . (v x w) and beta' = planar . (w x u) — 10
// operations, 3 loads — equal them up to rounding (a few 1e-16 times the size of the terms).  The traversals decide on
// (alpha', beta') whenever they are at least 2e-6 away from 0 and 1 — many orders of magnitude above that rounding — and
// evaluate the reference's expressions only inside that band, so every decision is the reference's.
// `g` = the quad's record, `planar` = hit point - start.
__device__ __forceinline__ bool quad_interior(const double2* __restrict__ g, d3 planar) {
    double2 c0, c1;
    ldg256(g + 4, c0, c1);
    const double2 c2 = __ldg(g + 6);
    const double a1 = dot(planar, mk(c0.x, c0.y, c1.x));
    const double b1 = dot(planar, mk(c1.y, c2.x, c2.y));
    // constant bands: if neither coordinate is flagged outside, |planar| is bounded by the quad's size and the rounding is
    // ~1e-13; if |planar| is so large that the rounding of one coordinate could reach 1e-6, the other coordinate is far
    // outside [0,1] for the exact expressions too, so "outside" is the exact answer either way
    const bool inside = (a1 >= 2e-6) && (a1 <= 1.0 - 2e-6) && (b1 >= 2e-6) && (b1 <= 1.0 - 2e-6);
    const bool outside = (a1 < -2e-6) || (a1 > 1.0 + 2e-6) || (b1 < -2e-6) || (b1 > 1.0 + 2e-6);
    if (inside || outside) return inside;
    const double2 u0 = __ldg(g + 8), u1 = __ldg(g + 9), v0 = __ldg(g + 10), v1 = __ldg(g + 11), w0 = __ldg(g + 12), w1 = __ldg(g + 13);
    const d3 bu = mk(u0.x, u0.y, u1.x), bv = mk(v0.x, v0.y, v1.x), bw = mk(w0.x, w0.y, w1.x);
    const double alpha = dot(bw, cross(planar, bv));
    const double beta = dot(bw, cross(bu, planar));
    return (0.0 <= alpha) && (alpha <= 1.0) && (0.0 <= beta) && (beta <= 1.0);
}

struct ClosestHit {
    double t;
    uint32_t pc;     // op index of the winning primitive, WRT_NONE = miss
    uint32_t xform;  // transform context it was hit in
};

// Sequential closest-hit scan in DFS order with the running tmax, i.e. exactly what the reference's recursion
// computes: spheres accept tmin < t < tmax (entity.zig:608), quads tmin <= t <= tmax (:485, interval.zig:26-33),
// so a later coincident quad replaces an earlier hit.
template <int CULL>
__device__ inline ClosestHit closest_hit(const DeviceScene& S, d3 wo, d3 wd, double time, double tmin, double tmax) {
    ClosestHit best;
    best.t = tmax; best.pc = WRT_NONE; best.xform = WRT_NONE;
    d3 o = wo, d = wd;
    uint32_t xf = WRT_NONE;
    Culler<CULL> cull;
    cull.set_ray(o, d);
    uint32_t pc = 0;
    for (;;) {
        const uint4 op = __ldg(S.ops + pc);
        if (op.x == OP_NODE || op.x == OP_NODE_TIGHT_ONLY) {
            bool pass = (CULL == WRT_CULL_REFERENCE && op.x == OP_NODE_TIGHT_ONLY) ? true : cull.pass(S, op.y, tmin, best.t);
            pc = pass ? pc + 1 : op.z;
        } else if (op.x == OP_SPHERE) {
            // SphereEntity.hit, entity.zig:585-623
            const double2* g = reinterpret_cast<const double2*>(S.spheres + op.y);
            double2 g0 = __ldg(g), g1 = __ldg(g + 1);
            d3 center = mk(g0.x, g0.y, g1.x);
            const double radius = g1.y;
            if (S.has_moving) {
                const SphereAux ax = S.sphere_aux[op.y];
                if (ax.is_moving) center = center + mk(ax.mx, ax.my, ax.mz) * time;  // entity.zig:653-656
            }
            d3 oc = center - o;
            double a = dot(d, d);
            double h = dot(d, oc);
            double c = dot(oc, oc) - radius * radius;
            double disc = h * h - a * c;
            if (!(disc < 0.0)) {
                double sq = sqrt(disc);
                double root = (h - sq) / a;
                bool ok = (tmin < root) && (root < best.t);
                if (!ok) {
                    root = (h + sq) / a;
                    ok = (tmin < root) && (root < best.t);
                }
                if (ok) { best.t = root; best.pc = pc; best.xform = xf; }
            }
            ++pc;
        } else if (op.x == OP_QUAD) {
            // QuadEntity.hit, entity.zig:477-501
            const double2* g = reinterpret_cast<const double2*>(S.quads + op.y);
            double2 n0 = __ldg(g), n1 = __ldg(g + 1);
            d3 n = mk(n0.x, n0.y, n1.x);
            double denom = dot(n, d);
            if (!(fabs(denom) < 1e-8)) {
                double t;
                if (quad_plane_t(n1.y - dot(n, o), denom, tmin, best.t, t)) {
                    double2 s0 = __ldg(g + 2), s1 = __ldg(g + 3), u0 = __ldg(g + 8), u1 = __ldg(g + 9);
                    double2 v0 = __ldg(g + 10), v1 = __ldg(g + 11), w0 = __ldg(g + 12), w1 = __ldg(g + 13);
                    d3 p = o + d * t;
                    d3 planar = p - mk(s0.x, s0.y, s1.x);
                    d3 bu = mk(u0.x, u0.y, u1.x), bv = mk(v0.x, v0.y, v1.x), bw = mk(w0.x, w0.y, w1.x);
                    double alpha = dot(bw, cross(planar, bv));
                    double beta = dot(bw, cross(bu, planar));
                    if ((0.0 <= alpha) && (alpha <= 1.0) && (0.0 <= beta) && (beta <= 1.0)) {
                        best.t = t; best.pc = pc; best.xform = xf;
                    }
                }
            }
            ++pc;
        } else if (op.x == OP_PUSH_TRANSLATE || op.x == OP_PUSH_ROTATE_Y) {
            apply_xform(S.xforms[op.y], o, d);
            xf = op.y;
            cull.set_ray(o, d);
            ++pc;
        } else if (op.x == OP_POP) {
            xf = op.y;
            ray_in_xform(S, xf, wo, wd, o, d);
            cull.set_ray(o, d);
            ++pc;
        } else {  // OP_END
            break;
        }
    }
    return best;
}

// Ordered form of the scan for large programs under WRT_CULL_TIGHT: near-child-first descent over the child-pair
// records with a short per-thread stack, so the running tmax shrinks early and far subtrees are never opened (the fixed
// DFS order visits ~3 500 nodes per ray on the 2^20-primitive scene, this visits a few hundred).
// The visiting order differs from the reference's, the RESULT does not: among hits with the minimal t the reference's
// sequential rule (sphere accepts t < tmax, quad accepts t <= tmax, in DFS order) keeps the first of them in DFS order
// unless a later one is a quad, in which case the last such quad wins.  Ops are numbered in DFS order, so tracking
// (first op, last quad op) of the minimal-t set reproduces that rule under any visiting order.
#define WRT_STACK_DEPTH 48
// Trees with at least this many child-pair records are walked through their four-wide form.  Measured: the 2^20-primitive
// scene (latency-bound: its records miss L1) gains 22 % from half the dependent fetches, the 484-sphere scene (cache-resident,
// issue-bound) loses 7 % to the extra box tests of records opened two levels at a time.
#ifndef WRT_WIDE_TREE_MIN_RECORDS
#define WRT_WIDE_TREE_MIN_RECORDS 16384u
#endif
// The traversal state.
// The per-lane traversal stack: entries {desc | first op, end op, xform, entry distance bits}.  TravLocalStack lives in local
// memory (lane-interleaved: an entry of a lane whose neighbours stand at other depths costs four 32-byte sectors per access,
// and the stacks of the resident warps fill most of L1).  TravHybridStack keeps the bottom D entries — where nearly all
// pushes and pops happen — in shared memory, one column per thread ([entry][thread]: the bank depends on the thread only, so
// lanes at different depths do not conflict), the rest in local memory.
struct TravLocalStack {
    static constexpr bool compact = false;
    static constexpr bool quant = false;
    uint4 e[WRT_STACK_DEPTH];
    __device__ __forceinline__ void put(int i, uint4 v) { e[i] = v; }
    __device__ __forceinline__ uint4 get(int i) const { return e[i]; }
};
// COMPACT entries, 8 bytes: {child word, sort key}.  For scenes that are ONE tree of single-primitive leaves without
// transforms (DeviceScene::compact_ok — the 2^20-primitive scene): a deferred child is its record (bit 31 set) or its
// primitive (kind bit 30 + record index), the key is the entry distance with the low two mantissa bits cleared — never
// above the exact distance, so the pop-time cull stays conservative.  There is no transform context to restore and no
// range to resume.  The program position of a primitive (the tie rule compares positions) is not carried: it is looked
// up in sphere_pc / quad_pc when the primitive is actually hit.  Half the local-memory sectors per push / pop, half the
// L1 the lane stacks occupy (L1 capacity is what the tree records want), and the sort moves (key, word) pairs, so no
// child field is picked by index afterwards.
#define WRT_PC_UNKNOWN 0x7FFFFFF0u
struct TravCompactStack {
    static constexpr bool compact = true;
    static constexpr bool quant = false;
    uint2 e[WRT_STACK_DEPTH];
    __device__ __forceinline__ void put(int i, uint2 v) { e[i] = v; }
    __device__ __forceinline__ uint2 get(int i) const { return e[i]; }
};
template <int D, int THREADS>
struct TravHybridStack {
    static constexpr bool compact = false;
    static constexpr bool quant = false;
    uint4 e[WRT_STACK_DEPTH - D];
    uint4* column;  // shared memory: this thread's entry 0; entry i is column[i * THREADS]
    __device__ __forceinline__ void put(int i, uint4 v) { if (i < D) column[i * THREADS] = v; else e[i - D] = v; }
    __device__ __forceinline__ uint4 get(int i) const { return (i < D) ? column[i * THREADS] : e[i - D]; }
};

struct TravCompactStackQ : TravCompactStack {  // compact entries, and the records are read in their quantised form (Node4Q)
    static constexpr bool quant = true;
};

// LEAN: the binary64 ray in the current transform context is NOT kept in the state (12 registers that are dead weight while
// a lane walks box records, which only need the culler's binary32 copy): the leaf ops re-form it from the world-space ray the
// caller can re-read (the persistent extend kernel: 48 bytes of its path record).  Buys the occupancy step of that kernel.
struct TravRay { d3 o, d; };  // ray in the current transform context
struct TravNoRay {};
template <bool LEAN>
struct TravT : std::conditional<LEAN, TravNoRay, TravRay>::type {
    static constexpr bool lean = LEAN;
    double best_t;
    uint32_t first_pc, first_xf, lastq_pc, lastq_xf;
    uint32_t xf, pc, end, node;
    int sp;
    float t_lo;
    Culler<WRT_CULL_TIGHT> cull;
};
using Trav = TravT<false>;
using TravLean = TravT<true>;

// the ray in the state's current transform context
template <class TR, typename WORLD>
__device__ __forceinline__ void trav_local_ray(const DeviceScene& S, const TR& T, WORLD&& world, d3& o, d3& d) {
    if constexpr (TR::lean) {
        d3 wo, wd;
        double time;
        world(wo, wd, time);
        if (T.xf == WRT_NONE) { o = wo; d = wd; }
        else ray_in_xform(S, T.xf, wo, wd, o, d);
    } else {
        o = T.o; d = T.d;
    }
}

template <class TR>
__device__ __forceinline__ void trav_init(const DeviceScene& S, TR& T, d3 wo, d3 wd, double time, double tmin, double tmax) {
    (void)time;
    T.best_t = tmax;
    if constexpr (!TR::lean) { T.o = wo; T.d = wd; }
    T.first_pc = WRT_NONE; T.first_xf = WRT_NONE; T.lastq_pc = WRT_NONE; T.lastq_xf = WRT_NONE;
    T.xf = WRT_NONE;
    T.cull.set_ray(wo, wd);
    T.sp = 0;
    T.t_lo = __double2float_rd(tmin);
    T.pc = 0; T.end = S.n_ops - 1;  // current op range (excludes OP_END)
    T.node = WRT_NONE;              // != NONE: descend from this child-pair record instead
}

// One child-pair record: test both children, go to the nearer one that is hit, defer the other.  Leaves T.node set (next
// record), or a leaf op range in (T.pc, T.end), or an empty range (nothing hit: the caller pops).
template <class TR, class STK>
__device__ __forceinline__ void trav_node_step(const DeviceScene& S, TR& T, STK& stack) {
    {
        const float4* p = reinterpret_cast<const float4*>(S.nodes2 + T.node);
        float4 a0, a1, b0, b1;
        ldg256(p, a0, a1); ldg256(p + 2, b0, b1);
        const float t_hi = __double2float_ru(T.best_t);
        float el, er = 0.0f;
        const bool hl = T.cull.entry(a0.x, a0.y, a0.z, a1.x, a1.y, a1.z, T.t_lo, t_hi, el);
        const uint32_t r_desc = __float_as_uint(b0.w);
        const bool hr = (r_desc != WRT_NONE) && T.cull.entry(b0.x, b0.y, b0.z, b1.x, b1.y, b1.z, T.t_lo, t_hi, er);
        // select-based (no three-way branch: the lanes of a warp stand on different records and would serialise it)
        const uint32_t l_desc = __float_as_uint(a0.w), l_end = __float_as_uint(a1.w), r_end = __float_as_uint(b1.w);
        const bool both = hl && hr;
        const bool go_left = both ? (el <= er) : hl;
        if (both && T.sp < WRT_STACK_DEPTH)
            stack.put(T.sp++, make_uint4(go_left ? r_desc : l_desc, go_left ? r_end : l_end, T.xf, __float_as_uint(go_left ? er : el)));
        if (hl || hr) {
            const uint32_t go_desc = go_left ? l_desc : r_desc;
            if (go_desc & 0x80000000u) { T.node = go_desc & 0x7FFFFFFFu; }
            else { T.node = WRT_NONE; T.pc = go_desc; T.end = go_left ? l_end : r_end; }
            return;
        }
        T.node = WRT_NONE; T.pc = 0; T.end = 0;  // nothing hit: the caller pops
    }
}

// One four-wide record: test the (up to) four children, go to the nearest one that is hit, defer the others so that the
// nearer ones are popped first.  Order only steers the search (closest hit and tie rule do not depend on it), so the sort
// runs on truncated keys: entry distance bits with the child index in the low two mantissa bits.
template <class TR, class STK>
__device__ __forceinline__ void trav_node4_step(const DeviceScene& S, TR& T, STK& stack) {
    const float4* p = reinterpret_cast<const float4*>(S.nodes4 + T.node);
    float4 lox, loy, loz, hix, hiy, hiz;
    uint4 desc, end;
    ldg256(p, lox, loy); ldg256(p + 2, loz, hix); ldg256(p + 4, hiy, hiz); ldg256(p + 6, desc, end);
    const float t_hi = __double2float_ru(T.best_t);
    float e0, e1, e2, e3;
    const bool h0 = T.cull.entry(lox.x, loy.x, loz.x, hix.x, hiy.x, hiz.x, T.t_lo, t_hi, e0) && desc.x != WRT_NONE;
    const bool h1 = T.cull.entry(lox.y, loy.y, loz.y, hix.y, hiy.y, hiz.y, T.t_lo, t_hi, e1) && desc.y != WRT_NONE;
    const bool h2 = T.cull.entry(lox.z, loy.z, loz.z, hix.z, hiy.z, hiz.z, T.t_lo, t_hi, e2) && desc.z != WRT_NONE;
    const bool h3 = T.cull.entry(lox.w, loy.w, loz.w, hix.w, hiy.w, hiz.w, T.t_lo, t_hi, e3) && desc.w != WRT_NONE;
    const uint32_t miss = 0x7F800000u;  // +inf: sorts last
    uint32_t k0 = h0 ? ((__float_as_uint(fmaxf(e0, 0.0f)) & ~3u) | 0u) : (miss | 0u);
    uint32_t k1 = h1 ? ((__float_as_uint(fmaxf(e1, 0.0f)) & ~3u) | 1u) : (miss | 1u);
    uint32_t k2 = h2 ? ((__float_as_uint(fmaxf(e2, 0.0f)) & ~3u) | 2u) : (miss | 2u);
    uint32_t k3 = h3 ? ((__float_as_uint(fmaxf(e3, 0.0f)) & ~3u) | 3u) : (miss | 3u);
#define WRT_CSWAP(a, b) { const uint32_t lo_ = min(a, b), hi_ = max(a, b); a = lo_; b = hi_; }
    WRT_CSWAP(k0, k1) WRT_CSWAP(k2, k3) WRT_CSWAP(k0, k2) WRT_CSWAP(k1, k3) WRT_CSWAP(k1, k2)
#undef WRT_CSWAP
    auto pick_u = [](const uint4& v, uint32_t i) { return (i & 2u) ? ((i & 1u) ? v.w : v.z) : ((i & 1u) ? v.y : v.x); };
    auto pick_e = [&](uint32_t i) { return (i & 2u) ? ((i & 1u) ? e3 : e2) : ((i & 1u) ? e1 : e0); };
    // defer the far ones, farthest first (k3 >= k2 >= k1): the entry distance kept for the pop-time cull is the exact one
    if (k3 < miss && T.sp < WRT_STACK_DEPTH) { const uint32_t i = k3 & 3u; const uint32_t dd = pick_u(desc, i), ee = pick_u(end, i); stack.put(T.sp++, make_uint4(dd, ee, T.xf, __float_as_uint(pick_e(i)))); }
    if (k2 < miss && T.sp < WRT_STACK_DEPTH) { const uint32_t i = k2 & 3u; const uint32_t dd = pick_u(desc, i), ee = pick_u(end, i); stack.put(T.sp++, make_uint4(dd, ee, T.xf, __float_as_uint(pick_e(i)))); }
    if (k1 < miss && T.sp < WRT_STACK_DEPTH) { const uint32_t i = k1 & 3u; const uint32_t dd = pick_u(desc, i), ee = pick_u(end, i); stack.put(T.sp++, make_uint4(dd, ee, T.xf, __float_as_uint(pick_e(i)))); }
    if (k0 < miss) {
        const uint32_t i = k0 & 3u, go = pick_u(desc, i);
        if (go & 0x80000000u) { T.node = go & 0x7FFFFFFFu; }
        else { T.node = WRT_NONE; T.pc = go; T.end = pick_u(end, i); }
        return;
    }
    T.node = WRT_NONE; T.pc = 0; T.end = 0;  // nothing hit: the caller pops
}

// The same step with compact stack entries (TravCompactStack): the sort carries (key, child word) pairs.
template <class TR, class STK>
__device__ __forceinline__ void trav_node4_step_compact(const DeviceScene& S, TR& T, STK& stack) {
    const float4* p = reinterpret_cast<const float4*>(S.nodes4 + T.node);
    float4 lox, loy, loz, hix, hiy, hiz;
    uint4 desc, end;
    ldg256(p, lox, loy); ldg256(p + 2, loz, hix); ldg256(p + 4, hiy, hiz); ldg256(p + 6, desc, end);
    const float t_hi = __double2float_ru(T.best_t);
    float e0, e1, e2, e3;
    const bool h0 = T.cull.entry(lox.x, loy.x, loz.x, hix.x, hiy.x, hiz.x, T.t_lo, t_hi, e0) && desc.x != WRT_NONE;
    const bool h1 = T.cull.entry(lox.y, loy.y, loz.y, hix.y, hiy.y, hiz.y, T.t_lo, t_hi, e1) && desc.y != WRT_NONE;
    const bool h2 = T.cull.entry(lox.z, loy.z, loz.z, hix.z, hiy.z, hiz.z, T.t_lo, t_hi, e2) && desc.z != WRT_NONE;
    const bool h3 = T.cull.entry(lox.w, loy.w, loz.w, hix.w, hiy.w, hiz.w, T.t_lo, t_hi, e3) && desc.w != WRT_NONE;
    const uint32_t miss = 0x7F800000u;  // +inf: sorts last
    uint32_t k0 = h0 ? ((__float_as_uint(fmaxf(e0, 0.0f)) & ~3u) | 0u) : (miss | 0u);
    uint32_t k1 = h1 ? ((__float_as_uint(fmaxf(e1, 0.0f)) & ~3u) | 1u) : (miss | 1u);
    uint32_t k2 = h2 ? ((__float_as_uint(fmaxf(e2, 0.0f)) & ~3u) | 2u) : (miss | 2u);
    uint32_t k3 = h3 ? ((__float_as_uint(fmaxf(e3, 0.0f)) & ~3u) | 3u) : (miss | 3u);
    // child word: the record of an inner child, the primitive (kind + index, from `end`) of a leaf
    uint32_t m0 = (desc.x & 0x80000000u) ? desc.x : (end.x & 0x7FFFFFFFu);
    uint32_t m1 = (desc.y & 0x80000000u) ? desc.y : (end.y & 0x7FFFFFFFu);
    uint32_t m2 = (desc.z & 0x80000000u) ? desc.z : (end.z & 0x7FFFFFFFu);
    uint32_t m3 = (desc.w & 0x80000000u) ? desc.w : (end.w & 0x7FFFFFFFu);
#define WRT_CSWAP2(ka, ma, kb, mb) { const bool sw_ = ka > kb; const uint32_t klo_ = sw_ ? kb : ka, khi_ = sw_ ? ka : kb, mlo_ = sw_ ? mb : ma, mhi_ = sw_ ? ma : mb; ka = klo_; kb = khi_; ma = mlo_; mb = mhi_; }
    WRT_CSWAP2(k0, m0, k1, m1) WRT_CSWAP2(k2, m2, k3, m3) WRT_CSWAP2(k0, m0, k2, m2) WRT_CSWAP2(k1, m1, k3, m3) WRT_CSWAP2(k1, m1, k2, m2)
#undef WRT_CSWAP2
    if (k3 < miss && T.sp < WRT_STACK_DEPTH) stack.put(T.sp++, make_uint2(m3, k3));
    if (k2 < miss && T.sp < WRT_STACK_DEPTH) stack.put(T.sp++, make_uint2(m2, k2));
    if (k1 < miss && T.sp < WRT_STACK_DEPTH) stack.put(T.sp++, make_uint2(m1, k1));
    if (k0 < miss) {
        if (m0 & 0x80000000u) { T.node = m0 & 0x7FFFFFFFu; }
        else { T.node = WRT_NONE; T.pc = WRT_PC_UNKNOWN; T.end = WRT_LEAF_PRIM | m0; }
        return;
    }
    T.node = WRT_NONE; T.pc = 0; T.end = 0;  // nothing hit: the caller pops
}

// The compact step over quantised records.  A slab distance is (corner + q * step) * inv - o * inv = q * (step * inv) +
// (corner * inv - o * inv): `step` is a power of two, so step * inv is exact, and the bracket is one FMA — one rounding more
// than Culler::entry's, of size (|corner * inv| + |o * inv|) * 2^-24, a quarter of the slack its margins leave (see there).
template <class TR, class STK>
__device__ __forceinline__ void trav_node4q_step_compact(const DeviceScene& S, TR& T, STK& stack) {
    const float4* p = reinterpret_cast<const float4*>(S.nodes4q + T.node);
    uint4 a, b, c, w;
    ldg256(p, a, b); ldg256(p + 2, c, w);  // a = (ox, oy, oz, exps)  b = (qlo x, y, z, qhi x)  c = (qhi y, z, word0, word1)  w = (word2, word3, -, -)
    const auto& cu = T.cull;
    const float sx = __uint_as_float((a.w & 0xFFu) << 23) * cu.inv_x, sy = __uint_as_float(((a.w >> 8) & 0xFFu) << 23) * cu.inv_y,
                sz = __uint_as_float(((a.w >> 16) & 0xFFu) << 23) * cu.inv_z;
    const float bx = fmaf(__uint_as_float(a.x), cu.inv_x, -cu.oi_x), by = fmaf(__uint_as_float(a.y), cu.inv_y, -cu.oi_y),
                bz = fmaf(__uint_as_float(a.z), cu.inv_z, -cu.oi_z);
    const float t_hi = __double2float_ru(T.best_t);
    const uint32_t word[4] = {c.z, c.w, w.x, w.y};
    const uint32_t miss = 0x7F800000u;  // +inf: sorts last
    uint32_t k[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float lx = fmaf((float)((b.x >> (8 * i)) & 0xFFu), sx, bx), hx = fmaf((float)((b.w >> (8 * i)) & 0xFFu), sx, bx);
        const float ly = fmaf((float)((b.y >> (8 * i)) & 0xFFu), sy, by), hy = fmaf((float)((c.x >> (8 * i)) & 0xFFu), sy, by);
        const float lz = fmaf((float)((b.z >> (8 * i)) & 0xFFu), sz, bz), hz = fmaf((float)((c.y >> (8 * i)) & 0xFFu), sz, bz);
        const float lo = fmaxf(fmaxf(fmaxf(fminf(lx, hx), fminf(ly, hy)), fminf(lz, hz)) - cu.err, T.t_lo);
        const float hi = fminf(fminf(fminf(fmaxf(lx, hx), fmaxf(ly, hy)), fmaxf(lz, hz)) + cu.err, t_hi);
        const bool hit = (hi * 1.000002f + 1e-30f >= lo) && word[i] != WRT_NONE;
        k[i] = hit ? ((__float_as_uint(fmaxf(lo, 0.0f)) & ~3u) | (uint32_t)i) : (miss | (uint32_t)i);
    }
    uint32_t k0 = k[0], k1 = k[1], k2 = k[2], k3 = k[3], m0 = word[0], m1 = word[1], m2 = word[2], m3 = word[3];
#define WRT_CSWAP2(ka, ma, kb, mb) { const bool sw_ = ka > kb; const uint32_t klo_ = sw_ ? kb : ka, khi_ = sw_ ? ka : kb, mlo_ = sw_ ? mb : ma, mhi_ = sw_ ? ma : mb; ka = klo_; kb = khi_; ma = mlo_; mb = mhi_; }
    WRT_CSWAP2(k0, m0, k1, m1) WRT_CSWAP2(k2, m2, k3, m3) WRT_CSWAP2(k0, m0, k2, m2) WRT_CSWAP2(k1, m1, k3, m3) WRT_CSWAP2(k1, m1, k2, m2)
#undef WRT_CSWAP2
    if (k3 < miss && T.sp < WRT_STACK_DEPTH) stack.put(T.sp++, make_uint2(m3, k3));
    if (k2 < miss && T.sp < WRT_STACK_DEPTH) stack.put(T.sp++, make_uint2(m2, k2));
    if (k1 < miss && T.sp < WRT_STACK_DEPTH) stack.put(T.sp++, make_uint2(m1, k1));
    if (k0 < miss) {
        if (m0 & 0x80000000u) { T.node = m0 & 0x7FFFFFFFu; }
        else { T.node = WRT_NONE; T.pc = WRT_PC_UNKNOWN; T.end = WRT_LEAF_PRIM | m0; }
        return;
    }
    T.node = WRT_NONE; T.pc = 0; T.end = 0;  // nothing hit: the caller pops
}

// WIDE: 0 = child-pair records, 1 = four-wide records (the hot kernels are instantiated for the form the scene uses: a run-time
// switch inside the per-lane megakernel's record loop cost the 484-sphere scene 7 %), 2 = ask the scene (gates, diagnostics)
template <int WIDE = 2, class TR, class STK>
__device__ __forceinline__ void trav_record_step(const DeviceScene& S, TR& T, STK& stack) {
    if constexpr (STK::compact && STK::quant) {
        trav_node4q_step_compact(S, T, stack);
    } else if constexpr (STK::compact) {
        trav_node4_step_compact(S, T, stack);
    } else {
        if (WIDE == 1 || (WIDE == 2 && S.use_wide)) trav_node4_step(S, T, stack);
        else trav_node_step(S, T, stack);
    }
}

// One op of the current leaf range (T.pc < T.end).
template <class TR, class STK, typename WORLD>
__device__ __forceinline__ void trav_leaf_op_lazy(const DeviceScene& S, TR& T, STK& stack, WORLD&& world, double tmin, double tmax) {
    const uint32_t pc = T.pc;
    uint4 op;
    if (T.end & WRT_LEAF_PRIM) {  // single-primitive leaf of a four-wide record: kind and record index travel in `end`, so the
        op = make_uint4((T.end & WRT_LEAF_QUAD) ? OP_QUAD : OP_SPHERE, T.end & WRT_LEAF_INDEX, 0u, 0u);  // op fetch (a dependent
        T.end = pc + 1;                                                                                   // load) is skipped
    } else {
        op = __ldg(S.ops + pc);
        if constexpr (STK::compact) {  // a compact-stack scene is one tree: the only op ever read is its root, at pc 0
            T.node = __ldg(S.root4 + op.y);
            return;
        }
    }
    d3 o, d;
    trav_local_ray(S, T, world, o, d);
    if constexpr (STK::compact) {  // only the two primitive kinds reach this point, and there is no transform context
        if (op.x != OP_SPHERE) op.x = OP_QUAD;
    }
    if (op.x == OP_NODE) {  // a bvh subtree inside this range: descend it ordered, come back for the rest of the range
        if constexpr (!STK::compact) {  // (a compact-stack scene is one tree: nothing follows the root in its range)
            if (op.z < T.end && T.sp < WRT_STACK_DEPTH) stack.put(T.sp++, make_uint4(op.z, T.end, T.xf, 0u));
        }
        T.node = S.use_wide ? __ldg(S.root4 + op.y) : op.y;
    } else if (op.x == OP_NODE_TIGHT_ONLY) {
        T.pc = T.cull.pass(S, op.y, tmin, T.best_t) ? pc + 1 : op.z;
    } else if (op.x == OP_SPHERE) {
        const double2* g = reinterpret_cast<const double2*>(S.spheres + op.y);
        double2 g0, g1;
        ldg256(g, g0, g1);
        d3 center = mk(g0.x, g0.y, g1.x);
        const double radius = g1.y;
        if (S.has_moving) {
            const SphereAux ax = S.sphere_aux[op.y];
            if (ax.is_moving) { d3 wo_, wd_; double time; world(wo_, wd_, time); center = center + mk(ax.mx, ax.my, ax.mz) * time; }
        }
        d3 oc = center - o;
        double a = dot(d, d);
        double h = dot(d, oc);
        double c = dot(oc, oc) - radius * radius;
        double disc = h * h - a * c;
        if (!(disc < 0.0)) {
            double sq = sqrt(disc);
            double root = (h - sq) / a;
            if (!(tmin < root)) root = (h + sq) / a;  // the near root is unusable: the reference then tries the far one
            if ((tmin < root) && (root <= T.best_t) && (root < tmax)) {
                uint32_t hp = pc;  // program position of the primitive (compact entries do not carry it)
                if constexpr (STK::compact) { if (pc == WRT_PC_UNKNOWN) hp = __ldg(S.sphere_pc + op.y); }
                if (root < T.best_t) { T.best_t = root; T.first_pc = hp; T.first_xf = T.xf; T.lastq_pc = WRT_NONE; }
                else if (hp < T.first_pc) { T.first_pc = hp; T.first_xf = T.xf; }
            }
        }
        T.pc = pc + 1;
    } else if (op.x == OP_QUAD) {
        const double2* g = reinterpret_cast<const double2*>(S.quads + op.y);
        double2 n0, n1;
        ldg256(g, n0, n1);
        d3 n = mk(n0.x, n0.y, n1.x);
        double denom = dot(n, d);
        if (!(fabs(denom) < 1e-8)) {
            double t;
            if (quad_plane_t(n1.y - dot(n, o), denom, tmin, T.best_t, t)) {
                double2 s0, s1;
                ldg256(g + 2, s0, s1);
                d3 p = o + d * t;
                d3 planar = p - mk(s0.x, s0.y, s1.x);
                if (quad_interior(g, planar)) {
                    uint32_t hp = pc;
                    if constexpr (STK::compact) { if (pc == WRT_PC_UNKNOWN) hp = __ldg(S.quad_pc + op.y); }
                    if (t < T.best_t) { T.best_t = t; T.first_pc = hp; T.first_xf = T.xf; T.lastq_pc = hp; T.lastq_xf = T.xf; }
                    else {
                        if (hp < T.first_pc) { T.first_pc = hp; T.first_xf = T.xf; }
                        if (T.lastq_pc == WRT_NONE || hp > T.lastq_pc) { T.lastq_pc = hp; T.lastq_xf = T.xf; }
                    }
                }
            }
        }
        T.pc = pc + 1;
    } else if (op.x == OP_PUSH_TRANSLATE || op.x == OP_PUSH_ROTATE_Y) {
        apply_xform(S.xforms[op.y], o, d);
        if constexpr (!TR::lean) { T.o = o; T.d = d; }
        T.xf = op.y;
        T.cull.set_ray(o, d);
        T.pc = pc + 1;
    } else if (op.x == OP_POP) {
        T.xf = op.y;
        { d3 wo, wd; double time; world(wo, wd, time); ray_in_xform(S, T.xf, wo, wd, o, d); }
        if constexpr (!TR::lean) { T.o = o; T.d = d; }
        T.cull.set_ray(o, d);
        T.pc = pc + 1;
    } else {
        T.pc = T.end;  // OP_END
    }
}

// Range exhausted: resume the nearest deferred subtree that can still hold a closer hit.  Returns true when the stack is
// empty (the traversal is complete).
template <class TR, class STK, typename WORLD>
__device__ __forceinline__ bool trav_pop_lazy(const DeviceScene& S, TR& T, STK& stack, WORLD&& world) {
    if constexpr (STK::compact) {
        (void)world;
        while (T.sp > 0) {
            const uint2 e = stack.get(--T.sp);
            if (__uint_as_float(e.y & ~3u) > __double2float_ru(T.best_t)) continue;  // its box now starts beyond the closest hit
            if (e.x & 0x80000000u) { T.node = e.x & 0x7FFFFFFFu; }
            else { T.node = WRT_NONE; T.pc = WRT_PC_UNKNOWN; T.end = WRT_LEAF_PRIM | e.x; }
            return false;
        }
        return true;
    } else {
    while (T.sp > 0) {
        const uint4 e = stack.get(--T.sp);
        if (__uint_as_float(e.w) > __double2float_ru(T.best_t)) continue;  // its box now starts beyond the closest hit
        if (e.z != T.xf) {
            T.xf = e.z;
            d3 wo, wd, o, d;
            double time;
            world(wo, wd, time);
            ray_in_xform(S, T.xf, wo, wd, o, d);
            if constexpr (!TR::lean) { T.o = o; T.d = d; }
            T.cull.set_ray(o, d);
        }
        if (e.x & 0x80000000u) { T.node = e.x & 0x7FFFFFFFu; }
        else { T.node = WRT_NONE; T.pc = e.x; T.end = e.y; }
        return false;
    }
    return true;
    }
}

// One op of the current leaf range (if any is left), then — when that exhausted the range — the pop, so a single-primitive
// leaf costs one step and leaves the lane on its next record.  Returns true when the traversal is complete.
template <class TR, class STK, typename WORLD>
__device__ __forceinline__ bool trav_leaf_step_lazy(const DeviceScene& S, TR& T, STK& stack, WORLD&& world, double tmin, double tmax) {
    if (T.pc < T.end) trav_leaf_op_lazy(S, T, stack, world, tmin, tmax);
    if (T.node == WRT_NONE && T.pc >= T.end) return trav_pop_lazy(S, T, stack, world);
    return false;
}
// the same with the world-space ray at hand (megakernels, gates)
template <class STK>
__device__ __forceinline__ bool trav_leaf_step(const DeviceScene& S, Trav& T, STK& stack, const d3& wo, const d3& wd, double time,
                                               double tmin, double tmax) {
    return trav_leaf_step_lazy(S, T, stack, [&](d3& o_, d3& d_, double& t_) { o_ = wo; d_ = wd; t_ = time; }, tmin, tmax);
}

template <class TR>
__device__ __forceinline__ ClosestHit trav_result(const TR& T) {
    ClosestHit best;
    best.t = T.best_t;
    const bool quad_wins = (T.lastq_pc != WRT_NONE) && (T.first_pc != WRT_NONE) && (T.lastq_pc > T.first_pc);
    best.pc = quad_wins ? T.lastq_pc : T.first_pc;
    best.xform = quad_wins ? T.lastq_xf : T.first_xf;
    return best;
}

template <int WIDE = 2>
__device__ inline ClosestHit closest_hit_ordered(const DeviceScene& S, d3 wo, d3 wd, double time, double tmin, double tmax) {
    Trav T;
    TravLocalStack stack;
    trav_init(S, T, wo, wd, time, tmin, tmax);
    // "while-while": every lane first descends through box records until it stands on a leaf range (cheap binary32 steps),
    // then the lanes of the warp run their binary64 primitive tests together
    for (;;) {
        while (T.node != WRT_NONE) trav_record_step<WIDE>(S, T, stack);
        if (trav_leaf_step(S, T, stack, wo, wd, time, tmin, tmax)) break;
    }
    return trav_result(T);
}

// per-lane scan: ordered descent where it applies (tight culling, stack deep enough for the tree), else the DFS-order scan
template <int CULL, int WIDE = 2>
__device__ __forceinline__ ClosestHit closest_hit_lane(const DeviceScene& S, d3 wo, d3 wd, double time, double tmin, double tmax) {
    if (CULL == WRT_CULL_TIGHT && S.use_ordered) return closest_hit_ordered<WIDE>(S, wo, wd, time, tmin, tmax);
    return closest_hit<CULL>(S, wo, wd, time, tmin, tmax);
}

// Packet form of the same scan for small programs (a few dozen ops: Cornell box, emissive, ...): the program counter
// is WARP-UNIFORM, every lane looks at the same op, and a node's subtree is skipped only when no lane of the warp
// needs it.  There is no op-kind divergence and every scene load is a broadcast.  Each lane still sees exactly the
// sequential scan of closest_hit(): a lane whose own box test failed is parked (`resume`) until the op after that
// subtree, which matters for WRT_CULL_REFERENCE where the reference's boxes are not always conservative.
// Must be called by all 32 lanes; `active` = the lane carries a ray.
template <int CULL>
__device__ inline ClosestHit closest_hit_packet(const DeviceScene& S, bool active, d3 wo, d3 wd, double time, double tmin, double tmax) {
    ClosestHit best;
    best.t = tmax; best.pc = WRT_NONE; best.xform = WRT_NONE;
    d3 o = wo, d = wd;
    uint32_t xf = WRT_NONE;
    Culler<CULL> cull;
    cull.template set_ray<true>(o, d);
    uint32_t pc = 0;      // uniform
    uint32_t resume = 0;  // per lane: first op this lane takes part in again
    for (;;) {
        const uint4 op = __ldg(S.ops + pc);
        const bool live = active && pc >= resume;
        if (op.x == OP_NODE || op.x == OP_NODE_TIGHT_ONLY) {
            bool pass = live;
            if (live && !(CULL == WRT_CULL_REFERENCE && op.x == OP_NODE_TIGHT_ONLY)) {
                pass = cull.pass(S, op.y, tmin, best.t);
                if (!pass) resume = op.z;
            }
            pc = __any_sync(0xffffffffu, pass) ? pc + 1 : op.z;
        } else if (op.x == OP_SPHERE) {
            if (live) {
                const double2* g = reinterpret_cast<const double2*>(S.spheres + op.y);
                double2 g0 = __ldg(g), g1 = __ldg(g + 1);
                d3 center = mk(g0.x, g0.y, g1.x);
                const double radius = g1.y;
                if (S.has_moving) {
                    const SphereAux ax = S.sphere_aux[op.y];
                    if (ax.is_moving) center = center + mk(ax.mx, ax.my, ax.mz) * time;
                }
                d3 oc = center - o;
                double a = dot(d, d);
                double h = dot(d, oc);
                double c = dot(oc, oc) - radius * radius;
                double disc = h * h - a * c;
                if (!(disc < 0.0)) {
                    double sq = sqrt(disc);
                    double root = (h - sq) / a;
                    bool ok = (tmin < root) && (root < best.t);
                    if (!ok) {
                        root = (h + sq) / a;
                        ok = (tmin < root) && (root < best.t);
                    }
                    if (ok) { best.t = root; best.pc = pc; best.xform = xf; }
                }
            }
            ++pc;
        } else if (op.x == OP_QUAD) {
            if (live) {
                const double2* g = reinterpret_cast<const double2*>(S.quads + op.y);
                double2 n0 = __ldg(g), n1 = __ldg(g + 1);
                d3 n = mk(n0.x, n0.y, n1.x);
                double denom = dot(n, d);
                if (!(fabs(denom) < 1e-8)) {
                    double t;
                    if (quad_plane_t(n1.y - dot(n, o), denom, tmin, best.t, t)) {
                        double2 s0 = __ldg(g + 2), s1 = __ldg(g + 3);
                        d3 p = o + d * t;
                        d3 planar = p - mk(s0.x, s0.y, s1.x);
                        if (quad_interior(g, planar)) { best.t = t; best.pc = pc; best.xform = xf; }
                    }
                }
            }
            ++pc;
        } else if (op.x == OP_PUSH_TRANSLATE || op.x == OP_PUSH_ROTATE_Y) {
            apply_xform(S.xforms[op.y], o, d);
            xf = op.y;
            if (!op.z) cull.template set_ray<true>(o, d);  // z = 1 (pruned program): no box test before the next transform op
            ++pc;
        } else if (op.x == OP_POP) {
            xf = op.y;
            ray_in_xform(S, xf, wo, wd, o, d);
            if (!op.z) cull.template set_ray<true>(o, d);
            ++pc;
        } else {  // OP_END
            break;
        }
    }
    return best;
}

// entity.zig:659-666 getSphereUv
__device__ __forceinline__ void sphere_uv(d3 v, double& u_out, double& v_out) {
    double theta = acos(-v.y);
    double phi = atan2(-v.z, v.x) + WRT_PI;
    u_out = phi / (2 * WRT_PI);
    v_out = theta / WRT_PI;
}

// Rebuild the HitRecord of the winning primitive with the reference's arithmetic (entity.zig:613-620, 490-498,
// hitrecord.zig:16-21), then carry point and normal back out through the transform chain (entity.zig:105-107,
// 184-186).  `want_uv` = the sphere UV (acos + atan2) is only evaluated when a texture will read it; `want_quad_uv` likewise
// for the quad's (alpha, beta).
__device__ inline void resolve_hit(const DeviceScene& S, const ClosestHit& ch, d3 wo, d3 wd, double time, bool want_uv, HitRecord& rec,
                                   bool want_quad_uv = true) {
    const uint4 op = __ldg(S.ops + ch.pc);
    d3 o, d;
    ray_in_xform(S, ch.xform, wo, wd, o, d);
    rec.t = ch.t;
    rec.material = op.z;
    rec.prim_id = op.w;
    rec.point = o + d * ch.t;
    d3 outward;
    if (op.x == OP_SPHERE) {
        const SphereGeom g = S.spheres[op.y];
        d3 center = mk(g.cx, g.cy, g.cz);
        if (S.has_moving) {
            const SphereAux ax = S.sphere_aux[op.y];
            if (ax.is_moving) center = center + mk(ax.mx, ax.my, ax.mz) * time;
        }
        outward = (rec.point - center) / g.radius;
        rec.is_sphere = 1;
        rec.sphere_outward = outward;
        rec.u = 0; rec.v = 0;
        if (want_uv) sphere_uv(outward, rec.u, rec.v);
    } else {
        const QuadGeom q = S.quads[op.y];
        d3 planar = rec.point - mk(q.sx, q.sy, q.sz);
        rec.u = 0; rec.v = 0;
        if (want_quad_uv) {  // alpha / beta again, as the texture coordinates (entity.zig:490-498); skipped for solid colours
            rec.u = dot(mk(q.wx, q.wy, q.wz), cross(planar, mk(q.vx, q.vy, q.vz)));
            rec.v = dot(mk(q.wx, q.wy, q.wz), cross(mk(q.ux, q.uy, q.uz), planar));
        }
        outward = mk(q.nx, q.ny, q.nz);
        rec.is_sphere = 0;
        rec.sphere_outward = outward;
    }
    rec.front_face = dot(d, outward) < 0.0;
    rec.normal = rec.front_face ? outward : -outward;
    for (uint32_t k = ch.xform; k != WRT_NONE; k = S.xforms[k].parent) {
        const Xform X = S.xforms[k];
        if (X.kind == OP_PUSH_TRANSLATE) {
            rec.point = rec.point + mk(X.a, X.b, X.c);
        } else {
            const double sn = X.a, cs = X.b;
            d3 p = rec.point, n = rec.normal;
            rec.point = mk(cs * p.x + sn * p.z, p.y, -sn * p.x + cs * p.z);
            rec.normal = mk(cs * n.x + sn * n.z, n.y, -sn * n.x + cs * n.z);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Textures (texture.zig)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool texture_needs_uv(const DeviceScene& S, uint32_t tex) {
    const Texture T = S.textures[tex];
    if (T.kind == WRT_TEX_IMAGE) return true;
    if (T.kind == WRT_TEX_CHECKER) return S.textures[T.even].kind != WRT_TEX_SOLID || S.textures[T.odd].kind != WRT_TEX_SOLID;
    return false;
}

__device__ inline d3 texture_value(const DeviceScene& S, uint32_t tex, double u, double v, d3 point) {
    // checker textures nest (texture.zig:111-118); solid / image terminate.  Depth is bounded at upload.
    for (int guard = 0; guard < 16; ++guard) {
        const Texture T = S.textures[tex];
        if (T.kind == WRT_TEX_SOLID) return mk(T.r, T.g, T.b);  // texture.zig:89-93
        if (T.kind == WRT_TEX_CHECKER) {
            int xi = (int)floor(T.inv_scale * point.x);
            int yi = (int)floor(T.inv_scale * point.y);
            int zi = (int)floor(T.inv_scale * point.z);
            int m = (xi + yi + zi) % 2;
            if (m < 0) m += 2;  // @mod
            tex = (m == 0) ? T.even : T.odd;
            continue;
        }
        // ImageTexture.value, texture.zig:49-68 + Image.getPixel, image.zig:23-36 + pixelToColor, texture.zig:70-77
        const ImageDesc im = S.images[T.image];
        double r, g, b;
        if (im.height == 0) {
            r = 255.0; g = 0.0; b = 255.0;  // ERR_COLOR, image.zig:5
        } else {
            double uu = clamp01(u);
            double vv = 1.0 - clamp01(v);
            unsigned long long xi = (unsigned long long)(uu * (double)im.width);
            unsigned long long yi = (unsigned long long)(vv * (double)im.height);
            uint32_t cx = xi > (unsigned long long)(im.width - 1) ? im.width - 1 : (uint32_t)xi;
            uint32_t cy = yi > (unsigned long long)(im.height - 1) ? im.height - 1 : (uint32_t)yi;
            uchar4 px = tex2D<uchar4>(im.tex, (float)cx + 0.5f, (float)cy + 0.5f);
            r = (double)px.x; g = (double)px.y; b = (double)px.z;
        }
        const double scale = 1.0 / 255.0;
        d3 c = mk(scale * r, scale * g, scale * b);
        return c * c;  // linearizeColorSpace, math.zig:172-174
    }
    return mk(0, 0, 0);
}

// ---------------------------------------------------------------------------------------------------------
// Light sampling hooks (entity.zig:371-386, 503-525, 626-651, 668-679) and PDFs (pdf.zig)
// ---------------------------------------------------------------------------------------------------------
__device__ inline double light_pdf_value_one(const SphereGeom* __restrict__ spheres, const QuadGeom* __restrict__ quads, const Light L,
                                             d3 origin, d3 direction) {
    if (L.kind == WRT_ENT_QUAD) {
        const double2* g = reinterpret_cast<const double2*>(quads + L.index);
        const double2 n0 = __ldg(g), n1 = __ldg(g + 1);
        d3 n = mk(n0.x, n0.y, n1.x);
        double denom = dot(n, direction);
        if (fabs(denom) < 1e-8) return 0.0;
        double t;
        if (!quad_plane_t(n1.y - dot(n, origin), denom, 1e-3, CUDART_INF, t)) return 0.0;
        const double2 s0 = __ldg(g + 2), s1 = __ldg(g + 3);  // start, area
        d3 p = origin + direction * t;
        d3 planar = p - mk(s0.x, s0.y, s1.x);
        if (!quad_interior(reinterpret_cast<const double2*>(quads + L.index), planar)) return 0.0;
        // record.normal is the face-forwarded normal; only |dot| is used (entity.zig:515)
        double dir_length_sq = dot(direction, direction);
        double dist_sq = t * t * dir_length_sq;
        bool front = dot(direction, n) < 0.0;
        d3 nn = front ? n : -n;
        double cosine = fabs(dot(direction, nn)) / sqrt(dir_length_sq);
        return dist_sq / (cosine * s1.y);
    }
    if (L.kind == WRT_ENT_SPHERE) {
        const SphereGeom g = spheres[L.index];
        d3 center = mk(g.cx, g.cy, g.cz);
        d3 oc = center - origin;
        double a = dot(direction, direction);
        double h = dot(direction, oc);
        double c = dot(oc, oc) - g.radius * g.radius;
        double disc = h * h - a * c;
        if (disc < 0.0) return 0.0;
        double sq = sqrt(disc);
        // SphereEntity.pdfValue only asks WHETHER a root lies in (1e-3, inf) (entity.zig:630-633): decided on the numerator
        // against 1e-3 * a (a > 0) unless the two are within 1e-12 of each other, where the reference's quotient is formed
        auto root_in_range = [a](double numr) {
            const double lim = 1e-3 * a;
            if (numr > lim * 1.000000000001 && numr <= 1.7e308) return true;
            if (numr < lim * 0.999999999999) return false;
            const double root = numr / a;
            return (1e-3 < root) && (root < CUDART_INF);
        };
        if (!root_in_range(h - sq) && !root_in_range(h + sq)) return 0.0;
        d3 diff = center - origin;
        double dist_sq = dot(diff, diff);
        double cos_theta_max = sqrt(1.0 - g.radius * g.radius / dist_sq);
        double solid_angle = 2.0 * WRT_PI * (1.0 - cos_theta_max);
        return 1.0 / solid_angle;
    }
    return 0.0;  // entity.zig:47-55
}

// entity.zig:371-378.  MANY = the kernel may meet long light lists (the per-lane kernels of large scenes): a light the ray
// certainly misses contributes weight * 0.0 = +0, which leaves the sum's bits alone, so it is skipped on the strength of
// the conservative binary32 box test (both hit tests start at tmin = 1e-3).  The packet kernels (small programs, a
// handful of lights, 80-register budget) are compiled without that path.
template <bool MANY>
__device__ __forceinline__ double lights_pdf_value(const DeviceScene& S, d3 origin, d3 direction) {
    const double weight = 1.0 / (double)S.n_lights;
    double sum = 0.0;
    if (MANY && S.n_lights > 4) {
        Culler<WRT_CULL_TIGHT> cull;
        cull.set_ray(origin, direction);
        for (uint32_t i = 0; i < S.n_lights; ++i) {
            if (!cull.pass_box(S.light_boxes + i, 1e-3, CUDART_INF)) continue;
            sum += weight * light_pdf_value_one(S.spheres, S.quads, S.lights[i], origin, direction);
        }
        return sum;
    }
    for (uint32_t i = 0; i < S.n_lights; ++i) sum += weight * light_pdf_value_one(S.spheres, S.quads, S.lights[i], origin, direction);
    return sum;
}

// writer.zig:68-94 encodeColor: NaN -> 0, sqrt, clamp [0, 0.999], * 256, truncate
__device__ __forceinline__ uint8_t encode_channel(double c) {
    if (c != c) c = 0.0;
    c = sqrt(c);
    c = fmax(0.0, fmin(c, 0.999));
    return (uint8_t)(256.0 * c);
}

}  // namespace wrt
