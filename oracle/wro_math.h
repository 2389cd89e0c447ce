/*
 * wro_math.h — ORACLE (test infrastructure, not product code).
 *
 * CPU restatement of the reference's vector / interval / AABB arithmetic:
 *   src/math/math.zig, src/math/interval.zig, src/math/aabb.zig, src/math/ray.zig
 * Every function cites the reference lines it follows.  Build with -ffp-contract=off: the reference
 * has no @mulAdd / optimized float mode, so results are plain IEEE-754 binary64 in source order.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use
 * anything under oracle/.
 */
#ifndef WRO_MATH_H
#define WRO_MATH_H

#include <math.h>
#include <stdbool.h>
#include <stdint.h>

#define WRO_PI 3.14159265358979323846264338327950288 /* std.math.pi */

typedef struct { double x, y, z; } v3;

static inline v3 v3_make(double x, double y, double z) { v3 r = {x, y, z}; return r; }  /* math.zig:156-162 */
static inline v3 v3_splat(double s) { v3 r = {s, s, s}; return r; }                      /* math.zig:144-146 */
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_mul(v3 a, v3 b) { return v3_make(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 v3_div(v3 a, v3 b) { return v3_make(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline v3 v3_scale(v3 a, double s) { return v3_make(a.x * s, a.y * s, a.z * s); }
static inline v3 v3_neg(v3 a) { return v3_make(-a.x, -a.y, -a.z); }
static inline double v3_get(v3 a, int axis) { return axis == 0 ? a.x : (axis == 1 ? a.y : a.z); }

/* Zig @min/@max on floats return the non-NaN operand (LLVM minnum/maxnum); C fmin/fmax agree. */
static inline v3 v3_min(v3 a, v3 b) { return v3_make(fmin(a.x, b.x), fmin(a.y, b.y), fmin(a.z, b.z)); }
static inline v3 v3_max(v3 a, v3 b) { return v3_make(fmax(a.x, b.x), fmax(a.y, b.y), fmax(a.z, b.z)); }

/* math.zig:243-246: lanes 0,1,2 of u*v added left to right */
static inline double v3_dot(v3 u, v3 v) {
    double x = u.x * v.x, y = u.y * v.y, z = u.z * v.z;
    return (x + y) + z;
}
/* math.zig:214-229: (u.yzx * v.zxy) - (u.zxy * v.yzx) */
static inline v3 v3_cross(v3 u, v3 v) {
    return v3_make(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
}
static inline double v3_length(v3 u) { return sqrt(v3_dot(u, u)); }              /* math.zig:254-256 */
static inline v3 v3_normalize(v3 u) { return v3_scale(u, 1.0 / v3_length(u)); }   /* math.zig:262-264 */
/* math.zig:270-272 */
static inline v3 v3_reflect(v3 v, v3 n) { return v3_sub(v, v3_scale(n, 2.0 * v3_dot(v, n))); }
/* math.zig:274-279 */
static inline v3 v3_refract(v3 vn, v3 n, double index) {
    double cos_theta = fmin(v3_dot(v3_neg(vn), n), 1.0);
    v3 r_out_perp = v3_scale(v3_add(vn, v3_scale(n, cos_theta)), index);
    v3 r_out_parallel = v3_scale(n, -sqrt(fabs(1.0 - v3_dot(r_out_perp, r_out_perp))));
    return v3_add(r_out_perp, r_out_parallel);
}

/* std.math.clamp(v, lo, hi) = @max(lo, @min(v, hi)) */
static inline double wro_clamp(double v, double lo, double hi) { return fmax(lo, fmin(v, hi)); }

/* OrthoBasis (math.zig:58-96) */
typedef struct { v3 u, v, w; } onb;
static inline onb onb_from_vectors(v3 u, v3 v, v3 w) { onb b = {u, v, w}; return b; }  /* math.zig:75-83 */
static inline onb onb_init(v3 n) {                                                       /* math.zig:65-73 */
    v3 w = v3_normalize(n);
    v3 a = (fabs(w.y) > 0.9) ? v3_make(1, 0, 0) : v3_make(0, 1, 0);
    v3 u = v3_normalize(v3_cross(w, a));
    v3 v = v3_cross(w, u);
    return onb_from_vectors(u, v, w);
}
static inline v3 onb_transform(const onb* b, v3 p) {                                    /* math.zig:89-95 */
    return v3_add(v3_add(v3_scale(b->u, p.x), v3_scale(b->v, p.y)), v3_scale(b->w, p.z));
}

/* Interval (interval.zig:3-43) */
typedef struct { double min, max; } ival;
static inline ival ival_union(ival a, ival b) { ival r = {fmin(a.min, b.min), fmax(a.max, b.max)}; return r; }
static inline ival ival_offset(ival a, double d) { ival r = {a.min + d, a.max + d}; return r; }
static inline double ival_size(ival a) { return a.max - a.min; }
static inline bool ival_contains(ival a, double t) { return (a.min <= t) && (t <= a.max); }   /* :26-28 closed */
static inline bool ival_surrounds(ival a, double t) { return (a.min < t) && (t < a.max); }    /* :31-33 open */
static inline ival ival_expand(ival a, double delta) {                                          /* :39-42 */
    double padding = delta / 2;
    ival r = {a.min - padding, a.max + padding};
    return r;
}

/* Ray (ray.zig:5-14) */
typedef struct { v3 origin, direction; double time; } ray;
static inline v3 ray_at(const ray* r, double t) { return v3_add(r->origin, v3_scale(r->direction, t)); }

/* AABB (aabb.zig:15-24).  Default = intervals [0,0]; min/max "undefined" in the reference, modelled as 0
 * (fresh MemoryPool pages are zero). */
typedef struct { ival x, y, z; v3 min, max; } aabb;

static inline aabb aabb_default(void) {
    aabb b = {{0, 0}, {0, 0}, {0, 0}, {0, 0, 0}, {0, 0, 0}};
    return b;
}
/* aabb.zig:103-122: interval padded by delta/2 per side, cached min/max by delta per side */
static inline void aabb_pad_to_minimum(aabb* b) {
    const double delta = 0.0001;
    v3 off = v3_splat(0);
    if (ival_size(b->x) < delta) { b->x = ival_expand(b->x, delta); off.x = delta; }
    if (ival_size(b->y) < delta) { b->y = ival_expand(b->y, delta); off.y = delta; }
    if (ival_size(b->z) < delta) { b->z = ival_expand(b->z, delta); off.z = delta; }
    b->min = v3_sub(b->min, off);
    b->max = v3_add(b->max, off);
}
/* aabb.zig:26-40 */
static inline aabb aabb_init(v3 a, v3 c) {
    aabb b;
    v3 mn = v3_min(a, c), mx = v3_max(a, c);
    b.x.min = mn.x; b.x.max = mx.x;
    b.y.min = mn.y; b.y.max = mx.y;
    b.z.min = mn.z; b.z.max = mx.z;
    b.min = mn; b.max = mx;
    aabb_pad_to_minimum(&b);
    return b;
}
/* aabb.zig:42-50 */
static inline aabb aabb_union(const aabb* a, const aabb* o) {
    aabb b;
    b.x = ival_union(a->x, o->x);
    b.y = ival_union(a->y, o->y);
    b.z = ival_union(a->z, o->z);
    b.min = v3_min(a->min, o->min);
    b.max = v3_max(a->max, o->max);
    return b;
}
/* aabb.zig:52-60: intervals shift, cached min moves by -d and max by +d (quirk A.9-4) */
static inline aabb aabb_offset(const aabb* a, v3 d) {
    aabb b;
    b.x = ival_offset(a->x, d.x);
    b.y = ival_offset(a->y, d.y);
    b.z = ival_offset(a->z, d.z);
    b.min = v3_sub(a->min, d);
    b.max = v3_add(a->max, d);
    return b;
}
static inline ival aabb_axis(const aabb* a, int axis) { return axis == 0 ? a->x : (axis == 1 ? a->y : a->z); }
/* aabb.zig:70-78 */
static inline int aabb_longest_axis(const aabb* a) {
    double lx = ival_size(a->x), ly = ival_size(a->y), lz = ival_size(a->z);
    if (lx > ly) return (lx > lz) ? 0 : 2;
    return (ly > lz) ? 1 : 2;
}
/* aabb.zig:80-101 as executed: rightPad (math.zig:186-190) starts at lane vecCapacity-1 = 2, so the z lane is
 * overwritten with tmin=0,tmax=1 and only x and y take part; each axis is compared on its own. */
static inline bool aabb_hit(const aabb* b, const ray* r, ival ray_t) {
    const double max_mult = 1.0000000000000004; /* math.zig:101-107 */
    double t0x = (b->min.x - r->origin.x) / r->direction.x;
    double t1x = (b->max.x - r->origin.x) / r->direction.x;
    double t0y = (b->min.y - r->origin.y) / r->direction.y;
    double t1y = (b->max.y - r->origin.y) / r->direction.y;
    double tminx = fmax(fmin(t0x, t1x), ray_t.min);
    double tmaxx = fmin(fmax(t0x, t1x), ray_t.max);
    double tminy = fmax(fmin(t0y, t1y), ray_t.min);
    double tmaxy = fmin(fmax(t0y, t1y), ray_t.max);
    tmaxx *= max_mult;
    tmaxy *= max_mult;
    return (tmaxx > tminx) && (tmaxy > tminy);
}

#endif /* WRO_MATH_H */
