/*
 * wro_sobol.h — ORACLE (test infrastructure, not product code).
 * CPU restatement of the reference's Sobol sampler, src/math/sampler.zig:162-299, over the tables of
 * src/math/sobolmatrices.zig (embedded from zig-weekend-raytracer_b200/data/sobol_tables.bin, which
 * tools/gen_sobol_tables.py extracts mechanically from the reference file).
 */
#ifndef WRO_SOBOL_H
#define WRO_SOBOL_H
#include <stdint.h>

#define WRO_SOBOL_DIMS 1024u       /* sobolmatrices.zig:39 */
#define WRO_SOBOL_MATRIX_SIZE 52u  /* sobolmatrices.zig:40 */
#define WRO_FLOAT32_ONE_MINUS_EPSILON 0x1.fffffep-1f /* sampler.zig:7 */

const uint32_t* wro_sobol_matrices32(void);      /* [1024*52] */
const uint64_t* wro_vdc_sobol_matrices(void);    /* [25][52]  */
const uint64_t* wro_vdc_sobol_matrices_inv(void);/* [26][52]  */

typedef struct wro_sobol_sampler {
    uint32_t samples_per_pixel;
    uint32_t scale;          /* ceilPowerOfTwo(max(W,H)), sampler.zig:188 */
    int owen_fast;           /* RandomizerStrategy, sampler.zig:9-12 */
    uint32_t seed;
    uint64_t pixel[2];
    uint32_t dimension;
    uint64_t sobol_idx;
} wro_sobol_sampler;

uint32_t wro_ceil_pow2_u32(uint32_t v);
uint32_t wro_log2_u32(uint32_t v);
void wro_sobol_init(wro_sobol_sampler* s, uint32_t spp, uint32_t width, uint32_t height, int owen_fast, uint32_t seed);
void wro_sobol_start_pixel_sample(wro_sobol_sampler* s, uint64_t col, uint64_t row, uint64_t sample_idx);
void wro_sobol_get_pixel_2d(const wro_sobol_sampler* s, double out[2]);
double wro_sobol_get_1d(wro_sobol_sampler* s);
void wro_sobol_get_2d(wro_sobol_sampler* s, double out[2]);
float wro_sobol_sample_dimension(const wro_sobol_sampler* s, uint32_t dimension);
float wro_sobol_sample(uint64_t a, uint32_t dimension, int owen_fast, uint32_t randomizer_seed);
uint64_t wro_sobol_interval_to_index(uint32_t log2_scale, uint64_t sample_idx, uint64_t px, uint64_t py);
uint32_t wro_owen_fast_apply(uint32_t seed, uint32_t v);
uint32_t wro_murmur2_hash_u32_with_seed(uint32_t v, uint32_t seed);
#endif
