/* ORACLE (test infrastructure, not product code) — see wro_sobol.h. */
#include "wro_sobol.h"

#include <math.h>
#include <string.h>

/* The blob layout is documented in tools/gen_sobol_tables.py. */
#ifndef WRO_SOBOL_BLOB
#define WRO_SOBOL_BLOB "../zig-weekend-raytracer_b200/data/sobol_tables.bin"
#endif
__asm__(".section .rodata\n"
        ".balign 16\n"
        ".global wro_sobol_blob\n"
        "wro_sobol_blob:\n"
        ".incbin \"" WRO_SOBOL_BLOB "\"\n"
        ".global wro_sobol_blob_end\n"
        "wro_sobol_blob_end:\n"
        ".previous\n");
extern const unsigned char wro_sobol_blob[];

#define BLOB_HEADER 24u
const uint32_t* wro_sobol_matrices32(void) { return (const uint32_t*)(wro_sobol_blob + BLOB_HEADER); }
const uint64_t* wro_vdc_sobol_matrices(void) {
    return (const uint64_t*)(wro_sobol_blob + BLOB_HEADER + 4u * WRO_SOBOL_DIMS * WRO_SOBOL_MATRIX_SIZE);
}
const uint64_t* wro_vdc_sobol_matrices_inv(void) { return wro_vdc_sobol_matrices() + 25u * WRO_SOBOL_MATRIX_SIZE; }

uint32_t wro_ceil_pow2_u32(uint32_t v) { /* std.math.ceilPowerOfTwo */
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}
uint32_t wro_log2_u32(uint32_t v) { /* std.math.log2_int */
    uint32_t l = 0;
    while (v >>= 1) ++l;
    return l;
}

static uint32_t bit_reverse32(uint32_t v) {
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
    return (v >> 16) | (v << 16);
}

/* sampler.zig:39-53 OwenFastRandomizer.apply (Laine-Karras hash between two bit reversals, wrapping u32) */
uint32_t wro_owen_fast_apply(uint32_t seed, uint32_t v) {
    v = bit_reverse32(v);
    v ^= v * 0x3d20adeau;
    v += seed;
    v *= (seed >> 16) | 1u;
    v ^= v * 0x05526c56u;
    v ^= v * 0x53a22864u;
    return bit_reverse32(v);
}

/* std.hash.Murmur2_32.hashUint32WithSeed (Zig std, not under /root/reference; MurmurHash2 of one 4-byte word).
 * Only reached from the scrambled dimensions >= 2, which the render path never requests. */
uint32_t wro_murmur2_hash_u32_with_seed(uint32_t v, uint32_t seed) {
    const uint32_t m = 0x5bd1e995u;
    const uint32_t len = 4;
    uint32_t h1 = seed ^ len;
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

/* sampler.zig:249-264 sobolSample */
float wro_sobol_sample(uint64_t a, uint32_t dimension, int owen_fast, uint32_t randomizer_seed) {
    const uint32_t* mat = wro_sobol_matrices32();
    uint32_t v = 0;
    uint32_t i = dimension * WRO_SOBOL_MATRIX_SIZE;
    for (; a != 0; a >>= 1, ++i) {
        if (a & 1) v ^= mat[i];
    }
    if (owen_fast) v = wro_owen_fast_apply(randomizer_seed, v);
    float vf = (float)v; /* @floatFromInt, round to nearest even */
    return fminf(vf * 0x1p-32f, WRO_FLOAT32_ONE_MINUS_EPSILON);
}

/* sampler.zig:267-298 sobolIntervalToIndex */
uint64_t wro_sobol_interval_to_index(uint32_t log2_scale, uint64_t sample_idx, uint64_t px, uint64_t py) {
    if (log2_scale == 0) return sample_idx;
    const uint64_t* vdc = wro_vdc_sobol_matrices() + (uint64_t)(log2_scale - 1) * WRO_SOBOL_MATRIX_SIZE;
    const uint64_t* inv = wro_vdc_sobol_matrices_inv() + (uint64_t)(log2_scale - 1) * WRO_SOBOL_MATRIX_SIZE;
    const uint32_t scale2 = log2_scale << 1;
    uint64_t index = sample_idx << scale2;
    uint64_t delta = 0;
    for (uint32_t c = 0; sample_idx > 0; sample_idx >>= 1, ++c) {
        if (sample_idx & 1) delta ^= vdc[c];
    }
    uint64_t b = ((px << log2_scale) | py) ^ delta;
    for (uint32_t c = 0; b > 0; b >>= 1, ++c) {
        if (b & 1) index ^= inv[c];
    }
    return index;
}

/* sampler.zig:177-195 initSampler */
void wro_sobol_init(wro_sobol_sampler* s, uint32_t spp, uint32_t width, uint32_t height, int owen_fast,
                    uint32_t seed) {
    memset(s, 0, sizeof *s);
    s->samples_per_pixel = spp;
    s->scale = wro_ceil_pow2_u32(width > height ? width : height);
    s->owen_fast = owen_fast;
    s->seed = seed;
}

/* sampler.zig:197-201 */
void wro_sobol_start_pixel_sample(wro_sobol_sampler* s, uint64_t col, uint64_t row, uint64_t sample_idx) {
    s->pixel[0] = col;
    s->pixel[1] = row;
    s->dimension = 2;
    s->sobol_idx = wro_sobol_interval_to_index(wro_log2_u32(s->scale), sample_idx, col, row);
}

/* sampler.zig:236-247 sampleDimension */
float wro_sobol_sample_dimension(const wro_sobol_sampler* s, uint32_t dimension) {
    if (!s->owen_fast) return wro_sobol_sample(s->sobol_idx, dimension, 0, 0);
    uint32_t hash = wro_murmur2_hash_u32_with_seed(dimension, s->seed);
    return wro_sobol_sample(s->sobol_idx, dimension, 1, hash);
}

/* sampler.zig:203-209 get1D */
double wro_sobol_get_1d(wro_sobol_sampler* s) {
    if (s->dimension >= WRO_SOBOL_DIMS) s->dimension = 2;
    double r = (double)wro_sobol_sample_dimension(s, s->dimension);
    s->dimension += 1;
    return r;
}

/* sampler.zig:211-220 get2D */
void wro_sobol_get_2d(wro_sobol_sampler* s, double out[2]) {
    if (s->dimension + 1 >= WRO_SOBOL_DIMS) s->dimension = 2;
    out[0] = (double)wro_sobol_sample_dimension(s, s->dimension);
    out[1] = (double)wro_sobol_sample_dimension(s, s->dimension + 1);
    s->dimension += 2;
}

/* sampler.zig:222-234 getPixel2D: dims 0/1, never scrambled, remapped into the pixel and clamped */
void wro_sobol_get_pixel_2d(const wro_sobol_sampler* s, double out[2]) {
    for (uint32_t dim = 0; dim < 2; ++dim) {
        double r = (double)wro_sobol_sample(s->sobol_idx, dim, 0, 0);
        r = r * (double)s->scale - (double)s->pixel[dim];
        /* std.math.clamp(v, 0, FLOAT32_ONE_MINUS_EPSILON) */
        out[dim] = fmax(0.0, fmin(r, (double)WRO_FLOAT32_ONE_MINUS_EPSILON));
    }
}
