// wrh_math.hpp — host-side mirror of the reference's math types that scene construction needs
// (src/math/math.zig, interval.zig, aabb.zig).  Host code only builds scenes and cameras; every ray is traced on
// the device.  Compile with -ffp-contract=off (the reference emits no fused multiply-adds).
#pragma once

#include <algorithm>
#include <cmath>

namespace wrh {

using Real = double;  // math.zig:40

struct Vec3 {  // math.zig:42 (lanes 0..2 of @Vector(4|8, f64))
    Real x = 0, y = 0, z = 0;
    constexpr Vec3() = default;
    constexpr Vec3(Real x_, Real y_, Real z_) : x(x_), y(y_), z(z_) {}
    static constexpr Vec3 splat(Real s) { return {s, s, s}; }  // vec3s, math.zig:144
    Real operator[](int axis) const { return axis == 0 ? x : (axis == 1 ? y : z); }
};
using Point3 = Vec3;
using Color = Vec3;

inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(Vec3 a, Vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline Vec3 operator/(Vec3 a, Vec3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline Vec3 operator-(Vec3 a) { return {-a.x, -a.y, -a.z}; }
// Zig @min/@max ignore a NaN operand; so do fmin/fmax
inline Vec3 vmin(Vec3 a, Vec3 b) { return {std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)}; }
inline Vec3 vmax(Vec3 a, Vec3 b) { return {std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)}; }

inline Real dot(Vec3 u, Vec3 v) {  // math.zig:243-246
    const Real a = u.x * v.x, b = u.y * v.y, c = u.z * v.z;
    return (a + b) + c;
}
inline Vec3 cross(Vec3 u, Vec3 v) {  // math.zig:214-229
    return {u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x};
}
inline Real length(Vec3 u) { return std::sqrt(dot(u, u)); }                      // math.zig:254-256
inline Vec3 normalize(Vec3 u) { return u * Vec3::splat(1.0 / length(u)); }       // math.zig:262-264

constexpr Real kPi = 3.14159265358979323846264338327950288;
inline Real degreesToRadians(Real d) { return d * (kPi / 180.0); }  // std.math.degreesToRadians

struct Interval {  // interval.zig:3-43
    Real min = 0, max = 0;
    Interval unionWith(Interval o) const { return {std::fmin(min, o.min), std::fmax(max, o.max)}; }
    Interval offset(Real d) const { return {min + d, max + d}; }
    Real size() const { return max - min; }
    bool contains(Real t) const { return min <= t && t <= max; }
    bool surrounds(Real t) const { return min < t && t < max; }
    Real clamp(Real t) const { return std::fmax(min, std::fmin(t, max)); }
    Interval expand(Real delta) const {
        const Real padding = delta / 2;
        return {min - padding, max + padding};
    }
};

enum class Axis { x = 0, y = 1, z = 2 };  // math.zig:49-56

struct AABB {  // aabb.zig:15-24; the default box is all zeros
    Interval x, y, z;
    Vec3 min, max;  // cached bounds: what the reference's hit test reads

    static AABB init(Vec3 a, Vec3 b) {  // aabb.zig:26-40
        AABB r;
        const Vec3 mn = vmin(a, b), mx = vmax(a, b);
        r.x = {mn.x, mx.x};
        r.y = {mn.y, mx.y};
        r.z = {mn.z, mx.z};
        r.min = mn;
        r.max = mx;
        r.padToMinimum();
        return r;
    }
    AABB unionWith(const AABB& o) const {  // aabb.zig:42-50
        AABB r;
        r.x = x.unionWith(o.x);
        r.y = y.unionWith(o.y);
        r.z = z.unionWith(o.z);
        r.min = vmin(min, o.min);
        r.max = vmax(max, o.max);
        return r;
    }
    AABB offset(Vec3 d) const {  // aabb.zig:52-60: cached min moves by -d, max by +d (quirk A.9-4)
        AABB r;
        r.x = x.offset(d.x);
        r.y = y.offset(d.y);
        r.z = z.offset(d.z);
        r.min = min - d;
        r.max = max + d;
        return r;
    }
    const Interval& axisInterval(Axis a) const { return a == Axis::x ? x : (a == Axis::y ? y : z); }
    Axis longestAxis() const {  // aabb.zig:70-78
        const Real lx = x.size(), ly = y.size(), lz = z.size();
        if (lx > ly) return lx > lz ? Axis::x : Axis::z;
        return ly > lz ? Axis::y : Axis::z;
    }

   private:
    void padToMinimum() {  // aabb.zig:103-122
        const Real delta = 0.0001;
        Vec3 off;
        if (x.size() < delta) { x = x.expand(delta); off.x = delta; }
        if (y.size() < delta) { y = y.expand(delta); off.y = delta; }
        if (z.size() < delta) { z = z.expand(delta); off.z = delta; }
        min = min - off;
        max = max + off;
    }
};

struct OrthoBasis {  // math.zig:58-96 (only initFromVectors is needed on the host: QuadEntity.basis)
    Vec3 u, v, w;
};

}  // namespace wrh
