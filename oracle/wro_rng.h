/*
 * wro_rng.h — ORACLE (test infrastructure, not product code).
 *
 * Random sources behind the reference's `std.Random` call sites (src/math/rng.zig).
 *
 * Two modes:
 *
 *  WRO_RNG_REFERENCE  the reference's generator as far as it can be restated without its toolchain:
 *      std.Random.DefaultPrng = Xoshiro256++ seeded through SplitMix64 (rng.zig:6,17), Random.float(f64)
 *      (52 mantissa bits + geometric exponent), Random.floatNorm (256-layer ziggurat) and
 *      Random.intRangeAtMost (Lemire).  Those live in the Zig standard library
 *      (0.14.0-dev.1827+e1e151df0, README.md:22), NOT under /root/reference: they are restated from the
 *      published algorithms; "parity unpinned" for the exact streams.  The reference seeds from getrandom
 *      (rng.zig:16-26) so no output of it is reproducible anyway; only the distributions matter.
 *
 *  WRO_RNG_COUNTER    the stream the CUDA back end uses: Philox4x32-10 keyed by the render seed and
 *      counted by (pixel index, sample index, draw index), so that every path consumes the same numbers on
 *      the CPU and on the device whatever the tiling.  Derived samplers use direct (non-Gaussian) forms that
 *      have the same distribution as the reference's (unit sphere, unit circle, integer pick); DESIGN.md §5.
 */
#ifndef WRO_RNG_H
#define WRO_RNG_H

#include <math.h>
#include <stdint.h>

#include "wro_math.h"

enum { WRO_RNG_REFERENCE = 0, WRO_RNG_COUNTER = 1 };

typedef struct wro_rng {
    int mode;
    uint64_t s[4]; /* Xoshiro256++ state */
    uint64_t seed; /* counter mode key */
    uint32_t pixel, sample, draw, base;
} wro_rng;

/* ---- Xoshiro256++ / SplitMix64 (Zig std.Random.Xoshiro256 / SplitMix64) ------------------------- */
static inline uint64_t wro_rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

static inline uint64_t wro_splitmix64(uint64_t* s) {
    *s += 0x9e3779b97f4a7c15ull;
    uint64_t z = *s;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

static inline void wro_rng_seed_reference(wro_rng* r, uint64_t seed) {
    r->mode = WRO_RNG_REFERENCE;
    uint64_t sm = seed;
    for (int i = 0; i < 4; ++i) r->s[i] = wro_splitmix64(&sm);
}

static inline uint64_t wro_xoshiro_next(wro_rng* r) {
    uint64_t* s = r->s;
    const uint64_t result = wro_rotl64(s[0] + s[3], 23) + s[0];
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = wro_rotl64(s[3], 45);
    return result;
}

/* ---- Philox4x32-10 (Salmon et al., SC'11) ---------------------------------------------------------- */
static inline void wro_philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

/* 64 random bits of draw `draw` of sample `sample` of pixel `pixel` under `seed`. */
static inline uint64_t wro_counter_bits(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t draw) {
    uint32_t c[4] = {pixel, sample, draw >> 1, 0u};
    wro_philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    return (draw & 1u) ? ((uint64_t)c[2] | ((uint64_t)c[3] << 32)) : ((uint64_t)c[0] | ((uint64_t)c[1] << 32));
}

static inline void wro_rng_start_counter(wro_rng* r, uint64_t seed, uint32_t pixel, uint32_t sample) {
    r->mode = WRO_RNG_COUNTER;
    r->seed = seed;
    r->pixel = pixel;
    r->sample = sample;
    r->draw = 0;
    r->base = 0;
}

/* Counter mode draw layout (shared with the device): draws 0,1 = lens sample, 2 = ray time; bounce b
 * owns draws 4+4b .. 4+4b+3 = {mixture choice | Fresnel uniform, light pick, u1, u2}.  Fixed slots keep one Philox
 * block per draw pair on the device.  No-ops for the reference-like generator. */
static inline void wro_rng_set_base(wro_rng* r, uint32_t base) {
    if (r->mode == WRO_RNG_COUNTER) { r->base = base; r->draw = base; }
}
static inline void wro_rng_slot(wro_rng* r, uint32_t slot) {
    if (r->mode == WRO_RNG_COUNTER) r->draw = r->base + slot;
}

static inline uint64_t wro_rng_u64(wro_rng* r) {
    if (r->mode == WRO_RNG_COUNTER) return wro_counter_bits(r->seed, r->pixel, r->sample, r->draw++);
    return wro_xoshiro_next(r);
}

/* ---- Random.float(f64) ------------------------------------------------------------------------------ */
static inline double wro_rng_float(wro_rng* r) {
    if (r->mode == WRO_RNG_COUNTER) {
        return (double)(wro_rng_u64(r) >> 11) * 0x1p-53;
    }
    /* Zig std.Random.float(f64): 52 random mantissa bits; exponent from the leading-zero count of the rest. */
    uint64_t rand = wro_xoshiro_next(r);
    unsigned rand_lz = rand ? (unsigned)__builtin_clzll(rand) : 64u;
    if (rand_lz >= 12) {
        rand_lz = 12;
        for (;;) {
            uint64_t more = wro_xoshiro_next(r);
            unsigned lz = more ? (unsigned)__builtin_clzll(more) : 64u;
            rand_lz += lz;
            if (lz != 64) break;
            if (rand_lz >= 1022) { rand_lz = 1022; break; }
        }
    }
    uint64_t mantissa = rand & 0xFFFFFFFFFFFFFull;
    uint64_t exponent = (uint64_t)(1022 - rand_lz) << 52;
    uint64_t bits = exponent | mantissa;
    double out;
    __builtin_memcpy(&out, &bits, sizeof out);
    return out;
}

/* ---- Random.intRangeAtMost(usize, 0, n-1) ------------------------------------------------------------ */
static inline uint32_t wro_rng_pick(wro_rng* r, uint32_t n) {
    if (r->mode == WRO_RNG_COUNTER) {
        uint32_t i = (uint32_t)(wro_rng_float(r) * (double)n);
        return i < n ? i : n - 1;
    }
    /* Zig uintLessThan: Lemire's nearly-divisionless method on 64-bit words */
    uint64_t less_than = n;
    uint64_t x = wro_xoshiro_next(r);
    __uint128_t m = (__uint128_t)x * less_than;
    uint64_t l = (uint64_t)m;
    if (l < less_than) {
        uint64_t t = (0 - less_than) % less_than;
        while (l < t) {
            x = wro_xoshiro_next(r);
            m = (__uint128_t)x * less_than;
            l = (uint64_t)m;
        }
    }
    return (uint32_t)(m >> 64);
}

/* ---- Random.floatNorm(f64): Zig std.Random.ziggurat, NormDist tables -------------------------------- */
typedef struct { double x[257]; double f[257]; int ready; } wro_zig_table;
static wro_zig_table wro_norm_table;
#define WRO_NORM_R 3.6541528853610088
#define WRO_NORM_V 0.00492867323399

static inline double wro_norm_f(double x) { return exp(-x * x / 2.0); }
static inline double wro_norm_f_inv(double y) { return sqrt(-2.0 * log(y)); }

static inline void wro_norm_table_init(void) {
    wro_zig_table* t = &wro_norm_table;
    if (t->ready) return;
    t->x[0] = WRO_NORM_V / wro_norm_f(WRO_NORM_R);
    t->x[1] = WRO_NORM_R;
    for (int i = 2; i < 256; ++i) {
        double last = t->x[i - 1];
        t->x[i] = wro_norm_f_inv(WRO_NORM_V / last + wro_norm_f(last));
    }
    t->x[256] = 0;
    for (int i = 0; i < 257; ++i) t->f[i] = wro_norm_f(t->x[i]);
    __atomic_store_n(&t->ready, 1, __ATOMIC_RELEASE);
}

static inline double wro_rng_float_norm_reference(wro_rng* r) {
    const wro_zig_table* t = &wro_norm_table;
    for (;;) {
        uint64_t bits = wro_xoshiro_next(r);
        unsigned i = (unsigned)(bits & 0xff);
        uint64_t repr = ((uint64_t)(0x3ff + 1) << 52) | (bits >> 12);
        double u;
        __builtin_memcpy(&u, &repr, sizeof u);
        u -= 3.0; /* [2,4) -> [-1,1) */
        double x = u * t->x[i];
        if (fabs(x) < t->x[i + 1]) return x;
        if (i == 0) {
            double xx = 1, yy = 0;
            while (-2.0 * yy < xx * xx) {
                xx = log(wro_rng_float(r)) / WRO_NORM_R;
                yy = log(wro_rng_float(r));
            }
            return (u < 0) ? xx - WRO_NORM_R : WRO_NORM_R - xx;
        }
        if (t->f[i + 1] + (t->f[i] - t->f[i + 1]) * wro_rng_float(r) < wro_norm_f(x)) return x;
    }
}

/* ---- samplers of src/math/rng.zig ------------------------------------------------------------------- */
/* rng.zig:87-95 sampleUnitSphere: normalise(N(0,1)^3).  Counter mode: uniform on the sphere directly. */
static inline v3 wro_sample_unit_sphere(wro_rng* r) {
    if (r->mode == WRO_RNG_COUNTER) {
        double u1 = wro_rng_float(r), u2 = wro_rng_float(r);
        double z = 1.0 - 2.0 * u1;
        double s = sqrt(fmax(0.0, 1.0 - z * z));
        double phi = 2.0 * WRO_PI * u2;
        return v3_make(cos(phi) * s, sin(phi) * s, z);
    }
    double a = wro_rng_float_norm_reference(r);
    double b = wro_rng_float_norm_reference(r);
    double c = wro_rng_float_norm_reference(r);
    return v3_normalize(v3_make(a, b, c));
}
/* rng.zig:71-73 sampleUnitCircleXY: normalise(N,N,0).  Counter mode: uniform angle. */
static inline v3 wro_sample_unit_circle_xy(wro_rng* r) {
    if (r->mode == WRO_RNG_COUNTER) {
        double phi = 2.0 * WRO_PI * wro_rng_float(r);
        return v3_make(cos(phi), sin(phi), 0.0);
    }
    double a = wro_rng_float_norm_reference(r);
    double b = wro_rng_float_norm_reference(r);
    return v3_normalize(v3_make(a, b, 0));
}
/* rng.zig:76-78 sampleUnitDiskXY: radius*float is evaluated before the circle sample (linear radius, A.9-9) */
static inline v3 wro_sample_unit_disk_xy(wro_rng* r, double radius) {
    double rr = radius * wro_rng_float(r);
    v3 c = wro_sample_unit_circle_xy(r);
    return v3_scale(c, rr);
}
/* rng.zig:104-114 sampleCosineDirectionZ */
static inline v3 wro_sample_cosine_direction_z(wro_rng* r) {
    double r1 = wro_rng_float(r);
    double r2 = wro_rng_float(r);
    double phi = 2.0 * WRO_PI * r1;
    double x = cos(phi) * sqrt(r2);
    double y = sin(phi) * sqrt(r2);
    double z = sqrt(1.0 - r2);
    return v3_make(x, y, z);
}
/* rng.zig:35-41 sampleVec3 */
static inline v3 wro_sample_vec3(wro_rng* r) {
    double a = wro_rng_float(r), b = wro_rng_float(r), c = wro_rng_float(r);
    return v3_make(a, b, c);
}
/* rng.zig:43-49 sampleVec3Interval */
static inline v3 wro_sample_vec3_interval(wro_rng* r, double lo, double hi) {
    double size = hi - lo;
    double a = wro_rng_float(r) * size + lo;
    double b = wro_rng_float(r) * size + lo;
    double c = wro_rng_float(r) * size + lo;
    return v3_make(a, b, c);
}

#endif /* WRO_RNG_H */
