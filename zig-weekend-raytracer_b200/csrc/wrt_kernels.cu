// wrt_kernels.cu — sm_100a kernels of the render back end.
//
//   render_kernel<CULL,TRAV> persistent warps run the iterative form of rayColor (render.zig:188-289).
//                            TRAV = packet: small programs, warp-uniform scan, work drawn per LANE as (pixel, sample chunk);
//                            TRAV = lane: large programs, ordered per-lane traversal, work drawn per warp as the reference's
//                            (row, 32-column block) jobs (render.zig:55-73) times a sample chunk.  A lane regenerates a
//                            camera sample as soon as its path ends; per-(pixel, chunk) sums go to a private slot.
//   render_kernel_sync / _regroup, wf_* kernels   alternative schedules of the same functions.
//   resolve_kernel           fused final pass: clear colour + ordered sum of the chunk slots -> caller's f64
//                            framebuffer layout, and encodeColor (writer.zig:68-94) into RGB8.
//   ppm_*_kernel             the PPM writer's body: block sizes, scan, formatting (writer.zig:16-123).
//   primary_hits_kernel / trace_rays_kernel / sobol_*_kernel / fp*_peak_kernel   gates and diagnostics (include/wrt.h).
#include <math_constants.h>

#include <type_traits>

#include <cstdlib>

#include "wrt_device.cuh"
#include "wrt_kernels.h"

namespace wrt {

// Per-launch constants travel as a __grid_constant__ kernel argument (LaunchParams: RenderConstants + the Sobol rows of
// the frame's resolution, 1.8 KB): the constant bank of the launch itself, so contexts that share a device — or host
// threads with a context each — never see one another's tables (there is no module-global __constant__ state).

__device__ __forceinline__ d3 ld3(const double* p) { return mk(p[0], p[1], p[2]); }

// sampleRay, render.zig:144-174 (+ sampleDefocusDisk :182-185, rng.sampleUnitDiskXY rng.zig:76-78).
// rng == nullptr: gate-1 dump (pinhole, time 0).  Draws: block 0 = (lens radius, lens angle), block 1.lo = time.
__device__ __forceinline__ Ray sample_ray_at(const RenderConstants& rc, uint32_t col, uint32_t row, double ox, double oy, bool dof,
                                             bool need_time, const Rng* rng);
__device__ __forceinline__ Ray sample_ray(const RenderConstants& rc, const SobolTables& sob, uint32_t col, uint32_t row, uint32_t s, bool dof,
                                          bool need_time, const Rng* rng) {
    uint64_t idx = sobol_interval_to_index(sob, s, col, row);
    double ox, oy;
    sobol_pixel_2d(sob, idx, col, row, ox, oy);
    return sample_ray_at(rc, col, row, ox, oy, dof, need_time, rng);
}
// The same for a lane that walks the samples of one pixel in order: the Sobol bits of sample s are those of s - 1 advanced
// by one XOR (SobolTables::inc*), only the first sample of a job pays for the index inversion.  Bit-identical jitter.
struct PixelSobol {
    uint32_t v0, v1;
    bool primed;
};
__device__ __forceinline__ Ray sample_ray_seq(const RenderConstants& rc, const SobolTables& sob, uint32_t col, uint32_t row, uint32_t s,
                                              PixelSobol& ps, bool dof, bool need_time, const Rng* rng) {
    if (ps.primed) {
        sobol_pixel_bits_next(sob, s - 1u, ps.v0, ps.v1);
    } else {
        sobol_pixel_bits(sob, sobol_interval_to_index(sob, s, col, row), ps.v0, ps.v1);
        ps.primed = true;
    }
    double ox, oy;
    sobol_bits_to_offsets(sob, ps.v0, ps.v1, col, row, ox, oy);
    return sample_ray_at(rc, col, row, ox, oy, dof, need_time, rng);
}
__device__ __forceinline__ Ray sample_ray_at(const RenderConstants& rc, uint32_t col, uint32_t row, double ox, double oy, bool dof,
                                             bool need_time, const Rng* rng) {
    d3 sample = (ld3(rc.cam.pixel00_loc) + ld3(rc.cam.pixel_delta_u) * ((double)col + ox)) +
                ld3(rc.cam.pixel_delta_v) * ((double)row + oy);
    d3 origin = ld3(rc.cam.position);
    if (dof) {
        double ur, ua;
        rng_pair(*rng, 0, ur, ua);
        double rr = 1.0 * ur;  // radius * float, evaluated before the circle sample (linear radius, A.9-9)
        double sn, cs;
        sincos(2.0 * WRT_PI * ua, &sn, &cs);
        d3 p = mk(cs, sn, 0.0) * rr;
        origin = (ld3(rc.cam.position) + ld3(rc.cam.defocus_disk_u) * p.x) + ld3(rc.cam.defocus_disk_v) * p.y;
    }
    Ray r;
    r.o = origin;
    r.d = sample - origin;
    r.time = 0.0;
    if (need_time) {  // time = rand.float (render.zig:167); only moving spheres read it
        double ut, unused;
        rng_pair(*rng, 1, ut, unused);
        r.time = ut;
    }
    return r;
}

// material.zig:221-225 (x^5 by multiplication instead of std.math.pow)
__device__ __forceinline__ double reflectance(double ir, double cosine) {
    double r0 = (1 - ir) / (1 + ir);
    r0 *= r0;
    double x = 1 - cosine;
    double x2 = x * x;
    return r0 + (1 - r0) * (x2 * x2 * x);
}

// texture.value with the common case (solid colour) inline and everything else out of line
__device__ __noinline__ d3 texture_value_general(const DeviceScene& S, uint32_t tex, double u, double v, d3 point, d3 outward, bool is_sphere) {
    if (is_sphere && texture_needs_uv(S, tex)) sphere_uv(outward, u, v);  // lazily: acos + atan2 (entity.zig:659-666)
    return texture_value(S, tex, u, v, point);
}
__device__ __forceinline__ d3 texture_color(const DeviceScene& S, uint32_t tex, const HitRecord& rec) {
    const Texture T = S.textures[tex];
    if (T.kind == WRT_TEX_SOLID) return mk(T.r, T.g, T.b);
    return texture_value_general(S, tex, rec.u, rec.v, rec.point, rec.sphere_outward, rec.is_sphere != 0);
}

enum SampleMode { SM_FIXED = 0, SM_LIGHT_QUAD, SM_COSINE, SM_SPHERE_UNIFORM, SM_LIGHT_SPHERE };

// ---- one bounce of rayColor in throughput form (render.zig:188-289), split by material class -------------
//   L    += beta * emitted                     (render.zig:234,288)
//   beta *= attenuation [* scatteringPdf/pdf]  (render.zig:245, 282-285)
// Each function returns false when the path ends.  Draw slots of bounce b: Philox blocks 2+2b (choice | Fresnel, pick)
// and 3+2b (u1, u2).

// DiffuseLightEmissiveMaterial: no scatter => return emission (material.zig:88-96, render.zig:238-240).  Back faces emit 0;
// the product is still formed so that a NaN/inf throughput poisons the sample as it does in the reference's recursion
// (0 * NaN), which the writer later zeroes (writer.zig:72-94).
__device__ __forceinline__ bool shade_emissive(const DeviceScene& S, const HitRecord& rec, const Material& M, const d3& beta, d3& L) {
    d3 e = rec.front_face ? texture_color(S, M.texture, rec) : mk(0, 0, 0);
    L = L + beta * e;
    return false;
}

// DielectricMaterial.scatter, material.zig:190-218 (attenuation (1,1,1))
__device__ __forceinline__ bool shade_dielectric(const HitRecord& rec, const Material& M, Ray& ray, const Rng& rng, uint32_t bounce) {
    double index = rec.front_face ? 1.0 / M.param : M.param;
    d3 in_unit = normalize(ray.d);
    double cos_theta = fmin(dot(-in_unit, rec.normal), 1.0);
    double sin_theta = sqrt(1 - cos_theta * cos_theta);
    double u0, unused;
    rng_pair(rng, 2u + 2u * bounce, u0, unused);
    d3 dir;
    if (index * sin_theta > 1.0 || reflectance(M.param, cos_theta) > u0) dir = reflect(in_unit, rec.normal);
    else dir = refract(in_unit, rec.normal, index);
    ray.o = rec.point;
    ray.d = dir;
    return true;
}

// Metal, lambertian and isotropic surfaces.  All direction sampling (cosine lobe, uniform sphere, cone towards a sphere
// light) funnels through ONE sincos + ONE orthonormal-basis site, selected per lane, so the lanes of a warp stay together
// whatever they sample.
template <bool MANY_LIGHTS = false>
__device__ __forceinline__ bool shade_surface(const DeviceScene& S, const HitRecord& rec, const Material& M, Ray& ray, d3& beta, d3& L,
                                              const Rng& rng, uint32_t bounce) {
    const uint32_t block_a = 2u + 2u * bounce, block_b = block_a + 1u;
    const bool diffuse = (M.kind != WRT_MAT_METAL);
    int mode = SM_SPHERE_UNIFORM;  // metal fuzz (material.zig:167-168) and the isotropic SpherePdf (pdf.zig:40-42)
    d3 attenuation = mk(M.ar, M.ag, M.ab);
    d3 w_n = mk(0, 0, 0);      // normalised shading normal = CosinePdf.basis.w
    d3 axis_w = mk(0, 0, 1);   // normalised axis of the sampling frame
    bool cosine_pdf = false;
    double cone = 0.0;         // sqrt(1 - r^2/dist^2) of the picked sphere light
    uint32_t light_quad = 0;   // picked quad light
    if (diffuse) {
        attenuation = texture_color(S, M.texture, rec);
        cosine_pdf = (M.kind == WRT_MAT_LAMBERTIAN) || !S.has_lights;  // render.zig:264-269 forces cosine without lights
        if (cosine_pdf) { w_n = normalize(rec.normal); mode = SM_COSINE; axis_w = w_n; }
        if (S.has_lights) {  // MixturePdf.generate, pdf.zig:113-117
            double p, upick;
            rng_pair(rng, block_a, p, upick);
            if (p < 0.5) {  // EntityCollection.sampleDirectionToSurface, entity.zig:381-386
                const Light Lt = S.lights[pick_index(upick, S.n_lights)];
                if (Lt.kind == WRT_ENT_QUAD) {
                    mode = SM_LIGHT_QUAD;
                    light_quad = Lt.index;
                } else if (Lt.kind == WRT_ENT_SPHERE) {  // entity.zig:646-651
                    mode = SM_LIGHT_SPHERE;
                    const SphereGeom g = S.spheres[Lt.index];
                    d3 to_light = mk(g.cx, g.cy, g.cz) - rec.point;
                    double dist_sq = dot(to_light, to_light);
                    axis_w = normalize(to_light);
                    cone = sqrt(1.0 - g.radius * g.radius / dist_sq);
                } else {
                    mode = SM_FIXED;  // entity.zig:58-65
                }
            }
        }
    }

    // ---- the one sampling site ----
    double u1, u2;
    rng_pair(rng, block_b, u1, u2);
    d3 dir;
    if (mode >= SM_COSINE) {
        double phi_u, s, z;
        if (mode == SM_COSINE) {  // rng.sampleCosineDirectionZ, rng.zig:104-114
            phi_u = u1; s = sqrt(u2); z = sqrt(1.0 - u2);
        } else if (mode == SM_SPHERE_UNIFORM) {  // rng.sampleUnitSphere in direct form
            z = 1.0 - 2.0 * u1; s = sqrt(fmax(0.0, 1.0 - z * z)); phi_u = u2;
        } else {  // randomToSphere, entity.zig:668-679
            z = 1.0 + u2 * (cone - 1.0); s = sqrt(1.0 - z * z); phi_u = u1;
        }
        double sn, cs;
        sincos(2.0 * WRT_PI * phi_u, &sn, &cs);
        d3 local = mk(cs * s, sn * s, z);
        if (mode == SM_SPHERE_UNIFORM) {
            dir = local;
        } else {  // OrthoBasis.init + transform, math.zig:65-73, 89-95
            d3 a = (fabs(axis_w.y) > 0.9) ? mk(1, 0, 0) : mk(0, 1, 0);
            d3 bu = normalize(cross(axis_w, a));
            d3 bv = cross(axis_w, bu);
            dir = (bu * local.x + bv * local.y) + axis_w * local.z;
        }
    } else if (mode == SM_LIGHT_QUAD) {  // QuadEntity.sampleDirectionToSurface, entity.zig:520-525
        const QuadGeom lq = S.quads[light_quad];
        d3 pu = mk(lq.ux, lq.uy, lq.uz) * u1;
        d3 pv = mk(lq.vx, lq.vy, lq.vz) * u2;
        dir = ((mk(lq.sx, lq.sy, lq.sz) + pu) + pv) - rec.point;
    } else {
        dir = mk(1, 0, 0);
    }

    if (!diffuse) {  // MetalMaterial.scatter, material.zig:163-178 (reflects the unnormalised direction, A.9-8)
        double blur = clamp01(M.param);
        d3 out = reflect(ray.d, rec.normal) + dir * blur;
        if (!(dot(out, rec.normal) > 0.0)) {  // scatter false => return emission (0), still weighted
            L = L + beta * 0.0;
            return false;
        }
        beta = beta * attenuation;
        ray.o = rec.point;
        ray.d = out;
        return true;
    }

    // ---- diffuse weights: attenuation * scatteringPdf / pdf (render.zig:280-285) ----
    const d3 dir_unit = normalize(dir);
    const double surface_pdf = cosine_pdf ? fmax(0.0, dot(dir_unit, w_n) / WRT_PI) : 1.0 / (4.0 * WRT_PI);  // pdf.zig:58-61, 36-38
    double pdf_value = surface_pdf;
    if (S.has_lights) pdf_value = 0.5 * lights_pdf_value<MANY_LIGHTS>(S, rec.point, dir) + 0.5 * surface_pdf;  // MixturePdf.value, pdf.zig:106-111
    double sp;  // material.scatteringPdf, material.zig:118-125 / 145-150
    if (M.kind == WRT_MAT_LAMBERTIAN) sp = fmax(0.0, dot(rec.normal, dir_unit) / WRT_PI);
    else sp = 1.0 / (4.0 * WRT_PI);
    d3 w = beta * (attenuation * sp);
    beta = mk(div_zero_aware(w.x, pdf_value), div_zero_aware(w.y, pdf_value), div_zero_aware(w.z, pdf_value));
    ray.o = rec.point;
    ray.d = dir;
    return true;
}

// All classes in one call (megakernel).
template <bool MANY_LIGHTS = false>
__device__ __forceinline__ bool shade(const DeviceScene& S, const RenderConstants& rc, const ClosestHit& ch, Ray& ray, d3& beta, d3& L,
                                      const Rng& rng, uint32_t bounce) {
    if (ch.pc == WRT_NONE) {  // render.zig:215-217
        L = L + beta * ld3(rc.background);
        return false;
    }
    const Material M = S.materials[__ldg(&S.ops[ch.pc].z)];
    // texture coordinates are only formed when a non-solid texture will read them
    const bool reads_texture = M.kind == WRT_MAT_LAMBERTIAN || M.kind == WRT_MAT_ISOTROPIC || M.kind == WRT_MAT_DIFFUSE_EMISSIVE;
    const bool textured = reads_texture && (S.textures[M.texture].kind != WRT_TEX_SOLID);
    HitRecord rec;
    resolve_hit(S, ch, ray.o, ray.d, ray.time, false, rec, textured);
    if (M.kind == WRT_MAT_DIFFUSE_EMISSIVE) return shade_emissive(S, rec, M, beta, L);
    if (M.kind == WRT_MAT_DIELECTRIC) return shade_dielectric(rec, M, ray, rng, bounce);
    return shade_surface<MANY_LIGHTS>(S, rec, M, ray, beta, L, rng, bounce);
}

enum { TRAV_LANE = 0, TRAV_PACKET = 1, TRAV_LANE_WIDE = 2 };  // LANE_WIDE: per-lane ordered traversal over the four-wide records

template <int CULL, int TRAV>
__global__ void __launch_bounds__(WRT_RENDER_BLOCK, TRAV == 1 ? WRT_PACKET_MIN_BLOCKS : WRT_RENDER_MIN_BLOCKS) render_kernel(const __grid_constant__ LaunchParams LP, DeviceScene S, double* __restrict__ accum,
                                                                                         unsigned long long* __restrict__ counters) {
    const RenderConstants& rc = LP.rc;
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long n_rays = 0, n_paths = 0;
    const double scale = 1.0 / (double)rc.spp;  // pixel_color_scale, render.zig:123
    const bool dof = rc.dof != 0;
    const bool need_time = S.has_moving != 0;
    Rng rng;
    rng.k0 = (uint32_t)rc.seed; rng.k1 = (uint32_t)(rc.seed >> 32);
    rng.pixel = 0; rng.sample = 0;

    if (TRAV == TRAV_PACKET) {
        // Packet scan: work is handed out per LANE.  A lane job is (pixel of the shard, sample chunk), and a lane that finishes its job takes
        // the next one while its neighbours are still mid-chunk, so no lane waits for the longest path sum of its warp and
        // the kernel's tail is one lane job.  The warp draws its lanes' jobs with one atomic (ballot + prefix count).
        uint32_t job = 0;               // this lane's job (the host keeps lane_jobs below 2^32)
        bool have_job = false, drained = false;
        uint32_t col = 0, row = 0, s = 0, s_last = 0;
        uint32_t acc_rays = 0;  // 32-bit tally, flushed to the 64-bit counter every ~64 K rays and at the end; paths are not
                                // counted: every job runs all its samples, the host knows the total
        // Few values live across the traversal (the kernel runs at an 80-register cap): the job's colour sum sits in its
        // accumulator slot and is updated once per path, the radiance of a path exists only in the iteration that ends it
        // (only terminal events add radiance: emission, background, the depth cut).
        bool alive = false;
        uint32_t depth_left = 0;
        Ray ray;
        ray.o = mk(0, 0, 0); ray.d = mk(0, 0, 1); ray.time = 0.0;
        d3 beta = mk(1, 1, 1);

        for (;;) {
            const bool want = !have_job && !drained;
            const uint32_t wanting = __ballot_sync(0xffffffffu, want);
            if (wanting) {
                const int leader = __ffs(wanting) - 1;
                unsigned long long first = 0;
                if ((int)lane == leader) first = atomicAdd(&counters[0], (unsigned long long)__popc(wanting));
                first = __shfl_sync(0xffffffffu, first, leader);
                if (want) {
                    const unsigned long long mine = first + __popc(wanting & ((1u << lane) - 1u));
                    job = (uint32_t)mine;
                    if (mine < rc.lane_jobs) {  // job -> (chunk, pixel of the shard); chunks outermost
                        const uint32_t chunk = (uint32_t)(job / rc.n_pixels_local);
                        const uint32_t pix = (uint32_t)(job % rc.n_pixels_local);
                        const uint32_t local_row = pix / rc.width;
                        col = pix - local_row * rc.width;
                        row = rc.row_shard_index + local_row * rc.row_shard_count;
                        s = rc.sample_begin + chunk * rc.chunk_size;
                        s_last = min(s + rc.chunk_size, rc.sample_end);
                        rng.pixel = row * rc.width + col;
                        have_job = true;
                        double* slot = accum + (size_t)job * 3;
                        slot[0] = 0.0; slot[1] = 0.0; slot[2] = 0.0;
                        if (rc.max_depth == 0) have_job = false;  // depth 0 returns black for every sample (render.zig:199)
                    } else {
                        drained = true;
                    }
                }
            }
            if (!alive && have_job) {
                rng.sample = s;
                                // (the incremental Sobol form of the lane kernel below costs this kernel 2 %: two more live registers at its 80-register cap)
                ray = sample_ray(rc, LP.sobol, col, row, s, dof, need_time, &rng);
                beta = mk(1, 1, 1);
                depth_left = rc.max_depth;
                alive = true;
            }
            if (!__any_sync(0xffffffffu, alive)) {
                if (__all_sync(0xffffffffu, drained)) break;  // no lane has a job left and the queue is empty
                continue;
            }
            const ClosestHit ch = closest_hit_packet<CULL>(S, alive, ray.o, ray.d, ray.time, 1e-4, CUDART_INF);
            if (alive) {
                ++acc_rays;
                d3 L = mk(0, 0, 0);
                const bool cont = shade(S, rc, ch, ray, beta, L, rng, rc.max_depth - depth_left);
                --depth_left;
                if (!cont || depth_left == 0) {
                    if (cont) L = L + beta * 0.0;  // depth exhausted: the tail returns 0 (render.zig:199), times the weight
                    double* slot = accum + (size_t)job * 3;  // (chunk, pixel) slot == job index
                    slot[0] = slot[0] + L.x * scale;    // color += L * scale, render.zig:129-135
                    slot[1] = slot[1] + L.y * scale;
                    slot[2] = slot[2] + L.z * scale;
                    alive = false;
                    if (++s == s_last) have_job = false;  // chunk complete
                    if (acc_rays >= (1u << 16)) {
                        atomicAdd(&counters[1], (unsigned long long)acc_rays);
                        acc_rays = 0;
                    }
                }
            }
        }
        n_rays = acc_rays;  // the remainder joins the warp reduction below
    } else {
        // Per-lane scan: warp jobs (row, 32-column block, sample chunk) keep the lanes of a warp on neighbouring pixels, whose
        // traversals are of similar length; lanes still regenerate their own samples inside the job.
        for (;;) {
            unsigned long long job = 0;
            if (lane == 0) job = atomicAdd(&counters[0], 1ull);
            job = __shfl_sync(0xffffffffu, job, 0);
            if (job >= rc.total_jobs) break;
            // job -> (chunk, local row, column block); chunks outermost
            const uint32_t blocks_per_chunk = rc.n_rows_local * rc.n_col_blocks;
            const uint32_t chunk = (uint32_t)(job / blocks_per_chunk);
            const uint32_t rem = (uint32_t)(job % blocks_per_chunk);
            const uint32_t local_row = rem / rc.n_col_blocks;
            const uint32_t col = (rem % rc.n_col_blocks) * 32u + lane;
            const uint32_t row = rc.row_shard_index + local_row * rc.row_shard_count;
            const uint32_t s_first = rc.sample_begin + chunk * rc.chunk_size;
            const uint32_t s_last = min(s_first + rc.chunk_size, rc.sample_end);
            const bool lane_active = col < rc.width;

            d3 color = mk(0, 0, 0);
            uint32_t s = s_first;
            bool alive = false;
            uint32_t depth_left = 0;
            Ray ray;
            ray.o = mk(0, 0, 0); ray.d = mk(0, 0, 1); ray.time = 0.0;
            d3 beta = mk(1, 1, 1), L = mk(0, 0, 0);
            rng.pixel = row * rc.width + col;
            rng.sample = 0;
            PixelSobol sob;
            sob.v0 = 0; sob.v1 = 0; sob.primed = false;

            for (;;) {
                if (!alive && lane_active && s < s_last) {
                    rng.sample = s;
                    ray = sample_ray_seq(rc, LP.sobol, col, row, s, sob, dof, need_time, &rng);
                    beta = mk(1, 1, 1); L = mk(0, 0, 0);
                    depth_left = rc.max_depth;
                    alive = depth_left > 0;  // depth == 0 returns black (render.zig:199)
                    ++n_paths;
                    if (!alive) ++s;
                }
                if (!__any_sync(0xffffffffu, alive)) {
                    if (!__any_sync(0xffffffffu, lane_active && s < s_last)) break;
                    continue;
                }
                ClosestHit ch;
                ch.pc = WRT_NONE;
                if (alive) ch = closest_hit_lane<CULL, TRAV == TRAV_LANE_WIDE ? 1 : 0>(S, ray.o, ray.d, ray.time, 1e-4, CUDART_INF);
                if (alive) {
                    ++n_rays;
                    const bool cont = shade<true>(S, rc, ch, ray, beta, L, rng, rc.max_depth - depth_left);
                    --depth_left;
                    if (!cont || depth_left == 0) {
                        if (cont) L = L + beta * 0.0;  // depth exhausted: the tail returns 0 (render.zig:199), times the weight
                        color = color + L * scale;     // render.zig:129-135
                        alive = false;
                        ++s;
                    }
                }
            }
            if (lane_active) {
                double* slot = accum + ((size_t)chunk * rc.n_rows_local * rc.width + (size_t)local_row * rc.width + col) * 3;
                slot[0] = color.x; slot[1] = color.y; slot[2] = color.z;
            }
        }
    }
    // ray / path accounting: one atomic per warp
    for (int off = 16; off > 0; off >>= 1) {
        n_rays += __shfl_down_sync(0xffffffffu, n_rays, off);
        n_paths += __shfl_down_sync(0xffffffffu, n_paths, off);
    }
    if (lane == 0) {
        atomicAdd(&counters[1], n_rays);
        atomicAdd(&counters[2], n_paths);
    }
}

// Phase-synchronous variant of render_kernel: ONE block of 16 warps per SM, and all 16 warps move through the three phases
// of an iteration (regenerate | closest hit | shade) together, separated by block barriers.  The integrator is the same,
// the arithmetic is the same, the frame is bit-identical; what changes is the instruction working set: at any time the
// whole SM executes one phase (~1-2 K instructions) instead of 16 warps spread over the ~60 KB loop body, which is what
// made instruction fetch the top stall reason of render_kernel (profiles/README.md).  A warp that finishes its job
// fetches the next one inside the same loop, so the barrier count per warp stays uniform.
template <int CULL, int TRAV>
__global__ void __launch_bounds__(WRT_SYNC_BLOCK, 1) render_kernel_sync(const __grid_constant__ LaunchParams LP, DeviceScene S, double* __restrict__ accum,
                                                                        unsigned long long* __restrict__ counters) {
    const RenderConstants& rc = LP.rc;
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long n_rays = 0, n_paths = 0;
    const double scale = 1.0 / (double)rc.spp;
    const bool dof = rc.dof != 0;
    const bool need_time = S.has_moving != 0;
    Rng rng;
    rng.k0 = (uint32_t)rc.seed; rng.k1 = (uint32_t)(rc.seed >> 32);
    rng.pixel = 0; rng.sample = 0;

    bool have_job = false, out_of_jobs = false;
    uint32_t chunk = 0, local_row = 0, col = 0, row = 0, s = 0, s_last = 0;
    bool lane_active = false, alive = false;
    uint32_t depth_left = 0;
    d3 color = mk(0, 0, 0);
    Ray ray;
    ray.o = mk(0, 0, 0); ray.d = mk(0, 0, 1); ray.time = 0.0;
    d3 beta = mk(1, 1, 1), L = mk(0, 0, 0);

    for (;;) {
        // ---- job management + phase A: regenerate ----
        if (!have_job && !out_of_jobs) {
            unsigned long long job = 0;
            if (lane == 0) job = atomicAdd(&counters[0], 1ull);
            job = __shfl_sync(0xffffffffu, job, 0);
            if (job >= rc.total_jobs) {
                out_of_jobs = true;
            } else {
                const uint32_t blocks_per_chunk = rc.n_rows_local * rc.n_col_blocks;
                chunk = (uint32_t)(job / blocks_per_chunk);
                const uint32_t rem = (uint32_t)(job % blocks_per_chunk);
                local_row = rem / rc.n_col_blocks;
                col = (rem % rc.n_col_blocks) * 32u + lane;
                row = rc.row_shard_index + local_row * rc.row_shard_count;
                s = rc.sample_begin + chunk * rc.chunk_size;
                s_last = min(s + rc.chunk_size, rc.sample_end);
                lane_active = col < rc.width;
                color = mk(0, 0, 0);
                alive = false;
                rng.pixel = row * rc.width + col;
                have_job = true;
            }
        }
        if (have_job && !alive && lane_active && s < s_last) {
            rng.sample = s;
            ray = sample_ray(rc, LP.sobol, col, row, s, dof, need_time, &rng);
            beta = mk(1, 1, 1); L = mk(0, 0, 0);
            depth_left = rc.max_depth;
            alive = depth_left > 0;
            ++n_paths;
            if (!alive) ++s;
        }
        __syncthreads();
        // ---- phase B: closest hit ----
        ClosestHit ch;
        ch.pc = WRT_NONE; ch.t = CUDART_INF; ch.xform = WRT_NONE;
        if (__any_sync(0xffffffffu, alive)) {
            if (TRAV == TRAV_PACKET) ch = closest_hit_packet<CULL>(S, alive, ray.o, ray.d, ray.time, 1e-4, CUDART_INF);
            else if (alive) ch = closest_hit_lane<CULL>(S, ray.o, ray.d, ray.time, 1e-4, CUDART_INF);
        }
        __syncthreads();
        // ---- phase C: shade, path end, job end ----
        if (alive) {
            ++n_rays;
            const bool cont = shade(S, rc, ch, ray, beta, L, rng, rc.max_depth - depth_left);
            --depth_left;
            if (!cont || depth_left == 0) {
                if (cont) L = L + beta * 0.0;
                color = color + L * scale;
                alive = false;
                ++s;
            }
        }
        if (have_job && !__any_sync(0xffffffffu, alive || (lane_active && s < s_last))) {
            if (lane_active) {
                double* slot = accum + ((size_t)chunk * rc.n_rows_local * rc.width + (size_t)local_row * rc.width + col) * 3;
                slot[0] = color.x; slot[1] = color.y; slot[2] = color.z;
            }
            have_job = false;
        }
        if (__syncthreads_and(out_of_jobs && !have_job)) break;
    }
    for (int off = 16; off > 0; off >>= 1) {
        n_rays += __shfl_down_sync(0xffffffffu, n_rays, off);
        n_paths += __shfl_down_sync(0xffffffffu, n_paths, off);
    }
    if (lane == 0) {
        atomicAdd(&counters[1], n_rays);
        atomicAdd(&counters[2], n_paths);
    }
}

// Phase-synchronous kernel with per-material regrouping in shared memory (packet traversal only).  After the closest-hit
// phase the block's paths are bucketed by what they hit — nothing to shade / terminal (miss, emissive) / dielectric /
// metal / diffuse — with a block-wide counting sort built from warp ballots, and the shading phase walks the sorted
// order: thread k shades the k-th path, whoever owns it.  Warps then run ONE material's code instead of all of them in
// sequence (the per-warp megakernel spends its shading instructions at ~10 of 32 active threads on the Cornell box).
// Inputs and results travel through a structure-of-arrays staging area in shared memory; each path is still owned by
// its lane (pixel sums, sample order), so the frame is bit-identical to render_kernel's.
struct RegroupSmem {
    double f[12][WRT_REGROUP_BLOCK];      // ray origin, ray direction, throughput, gathered radiance
    double t[WRT_REGROUP_BLOCK];          // closest-hit distance
    uint32_t w[5][WRT_REGROUP_BLOCK];     // hit op, hit xform, rng pixel, rng sample, bounce -> (after shading) w[0] = continue flag
    uint16_t order[WRT_REGROUP_BLOCK];    // sorted position -> owning thread
    uint32_t counts[WRT_REGROUP_BLOCK / 32][5];
    uint32_t prefix[WRT_REGROUP_BLOCK / 32][5];
    uint32_t total[8];
};

template <int CULL>
__global__ void __launch_bounds__(WRT_REGROUP_BLOCK, 1) render_kernel_regroup(const __grid_constant__ LaunchParams LP, DeviceScene S, double* __restrict__ accum,
                                                                               unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RegroupSmem& sm = *reinterpret_cast<RegroupSmem*>(smem_raw);
    const RenderConstants& rc = LP.rc;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    constexpr uint32_t NW = WRT_REGROUP_BLOCK / 32;
    unsigned long long n_rays = 0, n_paths = 0;
    const double scale = 1.0 / (double)rc.spp;
    const bool dof = rc.dof != 0;
    Rng rng;
    rng.k0 = (uint32_t)rc.seed; rng.k1 = (uint32_t)(rc.seed >> 32);
    rng.pixel = 0; rng.sample = 0;

    bool have_job = false, out_of_jobs = false;
    uint32_t chunk = 0, local_row = 0, col = 0, row = 0, s = 0, s_last = 0;
    bool lane_active = false, alive = false;
    uint32_t depth_left = 0;
    d3 color = mk(0, 0, 0);
    Ray ray;
    ray.o = mk(0, 0, 0); ray.d = mk(0, 0, 1); ray.time = 0.0;
    d3 beta = mk(1, 1, 1), L = mk(0, 0, 0);

    for (;;) {
        // ---- job management + regenerate ----
        if (!have_job && !out_of_jobs) {
            unsigned long long job = 0;
            if (lane == 0) job = atomicAdd(&counters[0], 1ull);
            job = __shfl_sync(0xffffffffu, job, 0);
            if (job >= rc.total_jobs) {
                out_of_jobs = true;
            } else {
                const uint32_t blocks_per_chunk = rc.n_rows_local * rc.n_col_blocks;
                chunk = (uint32_t)(job / blocks_per_chunk);
                const uint32_t rem = (uint32_t)(job % blocks_per_chunk);
                local_row = rem / rc.n_col_blocks;
                col = (rem % rc.n_col_blocks) * 32u + lane;
                row = rc.row_shard_index + local_row * rc.row_shard_count;
                s = rc.sample_begin + chunk * rc.chunk_size;
                s_last = min(s + rc.chunk_size, rc.sample_end);
                lane_active = col < rc.width;
                color = mk(0, 0, 0);
                alive = false;
                rng.pixel = row * rc.width + col;
                have_job = true;
            }
        }
        if (have_job && !alive && lane_active && s < s_last) {
            rng.sample = s;
            ray = sample_ray(rc, LP.sobol, col, row, s, dof, false, &rng);
            beta = mk(1, 1, 1); L = mk(0, 0, 0);
            depth_left = rc.max_depth;
            alive = depth_left > 0;
            ++n_paths;
            if (!alive) ++s;
        }
        // ---- closest hit (warp-uniform packet scan) ----
        ClosestHit ch;
        ch.pc = WRT_NONE; ch.t = CUDART_INF; ch.xform = WRT_NONE;
        if (__any_sync(0xffffffffu, alive)) ch = closest_hit_packet<CULL>(S, alive, ray.o, ray.d, 0.0, 1e-4, CUDART_INF);

        // ---- classify + stage ----
        uint32_t cls = 0;  // 0 nothing, 1 terminal (miss / emissive), 2 dielectric, 3 metal, 4 diffuse
        if (alive) {
            cls = 1;
            if (ch.pc != WRT_NONE) {
                const uint32_t kind = S.materials[__ldg(&S.ops[ch.pc].z)].kind;
                cls = (kind == WRT_MAT_DIFFUSE_EMISSIVE) ? 1u : (kind == WRT_MAT_DIELECTRIC) ? 2u : (kind == WRT_MAT_METAL) ? 3u : 4u;
            }
            sm.f[0][tid] = ray.o.x; sm.f[1][tid] = ray.o.y; sm.f[2][tid] = ray.o.z;
            sm.f[3][tid] = ray.d.x; sm.f[4][tid] = ray.d.y; sm.f[5][tid] = ray.d.z;
            sm.f[6][tid] = beta.x; sm.f[7][tid] = beta.y; sm.f[8][tid] = beta.z;
            sm.f[9][tid] = L.x; sm.f[10][tid] = L.y; sm.f[11][tid] = L.z;
            sm.t[tid] = ch.t;
            sm.w[0][tid] = ch.pc; sm.w[1][tid] = ch.xform; sm.w[2][tid] = rng.pixel; sm.w[3][tid] = rng.sample;
            sm.w[4][tid] = rc.max_depth - depth_left;
        }
        uint32_t rank_in_warp = 0;
#pragma unroll
        for (uint32_t c = 1; c < 5; ++c) {
            const unsigned m = __ballot_sync(0xffffffffu, cls == c);
            if (cls == c) rank_in_warp = __popc(m & ((1u << lane) - 1u));
            if (lane == 0) sm.counts[warp][c] = __popc(m);
        }
        __syncthreads();
        // exclusive scan of the per-warp counts of each class: warp c-1 scans class c (NW <= 32 values, one per lane)
        if (warp < 4) {
            const uint32_t c = warp + 1;
            const uint32_t mine = lane < NW ? sm.counts[lane][c] : 0u;
            uint32_t incl = mine;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, incl, off);
                if ((int)lane >= off) incl += up;
            }
            if (lane < NW) sm.prefix[lane][c] = incl - mine;
            if (lane == 31) sm.total[c] = incl;
        }
        __syncthreads();
        uint32_t n_work = 0, class_base = 0;
#pragma unroll
        for (uint32_t c = 1; c < 5; ++c) {
            const uint32_t tc = sm.total[c];
            if (c < cls) class_base += tc;
            n_work += tc;
        }
        if (cls != 0) sm.order[class_base + sm.prefix[warp][cls] + rank_in_warp] = (uint16_t)tid;
        __syncthreads();

        // ---- shade in sorted order ----
        if (tid < n_work) {
            const uint32_t src = sm.order[tid];
            Ray r2;
            r2.o = mk(sm.f[0][src], sm.f[1][src], sm.f[2][src]);
            r2.d = mk(sm.f[3][src], sm.f[4][src], sm.f[5][src]);
            r2.time = 0.0;
            d3 b2 = mk(sm.f[6][src], sm.f[7][src], sm.f[8][src]);
            d3 l2 = mk(sm.f[9][src], sm.f[10][src], sm.f[11][src]);
            ClosestHit c2;
            c2.t = sm.t[src]; c2.pc = sm.w[0][src]; c2.xform = sm.w[1][src];
            Rng g2;
            g2.k0 = rng.k0; g2.k1 = rng.k1; g2.pixel = sm.w[2][src]; g2.sample = sm.w[3][src];
            const bool cont = shade(S, rc, c2, r2, b2, l2, g2, sm.w[4][src]);
            sm.f[0][src] = r2.o.x; sm.f[1][src] = r2.o.y; sm.f[2][src] = r2.o.z;
            sm.f[3][src] = r2.d.x; sm.f[4][src] = r2.d.y; sm.f[5][src] = r2.d.z;
            sm.f[6][src] = b2.x; sm.f[7][src] = b2.y; sm.f[8][src] = b2.z;
            sm.f[9][src] = l2.x; sm.f[10][src] = l2.y; sm.f[11][src] = l2.z;
            sm.w[0][src] = cont ? 1u : 0u;
        }
        __syncthreads();

        // ---- owners take their results back ----
        if (alive) {
            ++n_rays;
            ray.o = mk(sm.f[0][tid], sm.f[1][tid], sm.f[2][tid]);
            ray.d = mk(sm.f[3][tid], sm.f[4][tid], sm.f[5][tid]);
            beta = mk(sm.f[6][tid], sm.f[7][tid], sm.f[8][tid]);
            L = mk(sm.f[9][tid], sm.f[10][tid], sm.f[11][tid]);
            const bool cont = sm.w[0][tid] != 0u;
            --depth_left;
            if (!cont || depth_left == 0) {
                if (cont) L = L + beta * 0.0;
                color = color + L * scale;
                alive = false;
                ++s;
            }
        }
        if (have_job && !__any_sync(0xffffffffu, alive || (lane_active && s < s_last))) {
            if (lane_active) {
                double* slot = accum + ((size_t)chunk * rc.n_rows_local * rc.width + (size_t)local_row * rc.width + col) * 3;
                slot[0] = color.x; slot[1] = color.y; slot[2] = color.z;
            }
            have_job = false;
        }
        if (__syncthreads_and(out_of_jobs && !have_job)) break;
    }
    for (int off = 16; off > 0; off >>= 1) {
        n_rays += __shfl_down_sync(0xffffffffu, n_rays, off);
        n_paths += __shfl_down_sync(0xffffffffu, n_paths, off);
    }
    if (lane == 0) {
        atomicAdd(&counters[1], n_rays);
        atomicAdd(&counters[2], n_paths);
    }
}

// Final pass: framebuffer[pixel] = (clear | previous contents) + sum over chunks in order; optional RGB8.
__global__ void resolve_kernel(const double* __restrict__ accum, uint32_t n_chunks, uint32_t n_pixels, double clear_r,
                               double clear_g, double clear_b, int no_clear, double* __restrict__ fb, uint32_t stride_doubles,
                               uint8_t* __restrict__ rgb8) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pixels) return;
    double* px = fb + (size_t)i * stride_doubles;
    double r = no_clear ? px[0] : clear_r, g = no_clear ? px[1] : clear_g, b = no_clear ? px[2] : clear_b;
    d3 sum = mk(0, 0, 0);
    for (uint32_t c = 0; c < n_chunks; ++c) {
        const double* slot = accum + ((size_t)c * n_pixels + i) * 3;
        sum = sum + mk(slot[0], slot[1], slot[2]);
    }
    r += sum.x; g += sum.y; b += sum.z;  // framebuffer.buffer[..] += color, render.zig:139
    px[0] = r; px[1] = g; px[2] = b;
    for (uint32_t k = 3; k < stride_doubles; ++k) px[k] = 0.0;
    if (rgb8) {
        rgb8[3 * (size_t)i + 0] = encode_channel(r);
        rgb8[3 * (size_t)i + 1] = encode_channel(g);
        rgb8[3 * (size_t)i + 2] = encode_channel(b);
    }
}

__global__ void encode_kernel(const double* __restrict__ fb, uint32_t stride_doubles, uint32_t n_pixels, uint8_t* __restrict__ rgb8) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pixels) return;
    const double* px = fb + (size_t)i * stride_doubles;
    rgb8[3 * (size_t)i + 0] = encode_channel(px[0]);
    rgb8[3 * (size_t)i + 1] = encode_channel(px[1]);
    rgb8[3 * (size_t)i + 2] = encode_channel(px[2]);
}

// ---- PPM body on the device (writer.zig:16-123) ------------------------------------------------------------------
// The reference sizes every 1024-pixel chunk of "{r} {g} {b}\n" lines in a serial pre-pass on the main thread and then
// formats the chunks on its thread pool (writer.zig:33-66).  Here: per-block byte counts, one exclusive scan, then every
// block formats its 1024 pixels into shared memory and copies them to their final file offset with coalesced stores.
#define WRT_PPM_BLOCK 256
#define WRT_PPM_PIXELS_PER_THREAD 4
#define WRT_PPM_BLOCK_PIXELS (WRT_PPM_BLOCK * WRT_PPM_PIXELS_PER_THREAD)
__device__ __forceinline__ uint32_t ppm_digits(uint32_t v) { return v > 99u ? 3u : (v > 9u ? 2u : 1u); }  // sizeOfDigit, writer.zig:107-114
__device__ __forceinline__ uint32_t ppm_line_bytes(const uint8_t* px) {  // sizeOfLine, writer.zig:96-100
    return 3u + ppm_digits(px[0]) + ppm_digits(px[1]) + ppm_digits(px[2]);
}
__device__ __forceinline__ uint32_t ppm_thread_bytes(const uint8_t* __restrict__ rgb, uint32_t first, uint32_t n_pixels, uint32_t len[WRT_PPM_PIXELS_PER_THREAD]) {
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < WRT_PPM_PIXELS_PER_THREAD; ++k) {
        const uint32_t i = first + k;
        len[k] = (i < n_pixels) ? ppm_line_bytes(rgb + 3 * (size_t)i) : 0u;
        sum += len[k];
    }
    return sum;
}
// exclusive prefix of `v` over the block (256 threads); `total` = block sum
__device__ __forceinline__ uint32_t ppm_block_scan(uint32_t v, uint32_t* warp_sums, uint32_t& total) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= (uint32_t)off) inc += up;
    }
    if (lane == 31u) warp_sums[warp] = inc;
    __syncthreads();
    uint32_t base = 0;
    total = 0;
#pragma unroll
    for (uint32_t w = 0; w < WRT_PPM_BLOCK / 32; ++w) {
        const uint32_t ws = warp_sums[w];
        if (w < warp) base += ws;
        total += ws;
    }
    __syncthreads();
    return base + inc - v;
}
__global__ void __launch_bounds__(WRT_PPM_BLOCK) ppm_block_bytes_kernel(const uint8_t* __restrict__ rgb, uint32_t n_pixels, uint32_t* __restrict__ block_bytes) {
    __shared__ uint32_t warp_sums[WRT_PPM_BLOCK / 32];
    uint32_t len[WRT_PPM_PIXELS_PER_THREAD];
    const uint32_t first = blockIdx.x * WRT_PPM_BLOCK_PIXELS + threadIdx.x * WRT_PPM_PIXELS_PER_THREAD;
    uint32_t total;
    ppm_block_scan(ppm_thread_bytes(rgb, first, n_pixels, len), warp_sums, total);
    if (threadIdx.x == 0) block_bytes[blockIdx.x] = total;
}
// one block: exclusive scan of the block byte counts into 64-bit file offsets; offsets[n_blocks] = body size
__global__ void __launch_bounds__(1024) ppm_scan_kernel(const uint32_t* __restrict__ block_bytes, uint32_t n_blocks, unsigned long long* __restrict__ offsets) {
    __shared__ unsigned long long part[1024];
    const uint32_t per = (n_blocks + 1023u) / 1024u;
    const uint32_t lo = min(threadIdx.x * per, n_blocks), hi = min(lo + per, n_blocks);
    unsigned long long sum = 0;
    for (uint32_t i = lo; i < hi; ++i) sum += block_bytes[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int t = 0; t < 1024; ++t) { const unsigned long long v = part[t]; part[t] = run; run += v; }
        offsets[n_blocks] = run;
    }
    __syncthreads();
    unsigned long long run = part[threadIdx.x];
    for (uint32_t i = lo; i < hi; ++i) { offsets[i] = run; run += block_bytes[i]; }
}
__device__ __forceinline__ uint32_t ppm_put(uint8_t* out, uint32_t v, uint8_t sep) {  // "{d}" + separator, writer.zig:62
    uint32_t n = 0;
    if (v > 99u) out[n++] = (uint8_t)('0' + v / 100u);
    if (v > 9u) out[n++] = (uint8_t)('0' + (v / 10u) % 10u);
    out[n++] = (uint8_t)('0' + v % 10u);
    out[n++] = sep;
    return n;
}
__global__ void __launch_bounds__(WRT_PPM_BLOCK) ppm_format_kernel(const uint8_t* __restrict__ rgb, uint32_t n_pixels,
                                                                   const unsigned long long* __restrict__ offsets, uint8_t* __restrict__ body) {
    __shared__ uint32_t warp_sums[WRT_PPM_BLOCK / 32];
    __shared__ uint8_t text[WRT_PPM_BLOCK_PIXELS * 12];
    uint32_t len[WRT_PPM_PIXELS_PER_THREAD];
    const uint32_t first = blockIdx.x * WRT_PPM_BLOCK_PIXELS + threadIdx.x * WRT_PPM_PIXELS_PER_THREAD;
    uint32_t total;
    uint32_t at = ppm_block_scan(ppm_thread_bytes(rgb, first, n_pixels, len), warp_sums, total);
#pragma unroll
    for (int k = 0; k < WRT_PPM_PIXELS_PER_THREAD; ++k) {
        const uint32_t i = first + k;
        if (i < n_pixels) {
            const uint8_t* px = rgb + 3 * (size_t)i;
            at += ppm_put(text + at, px[0], ' ');
            at += ppm_put(text + at, px[1], ' ');
            at += ppm_put(text + at, px[2], '\n');
        }
    }
    __syncthreads();
    uint8_t* dst = body + offsets[blockIdx.x];
    for (uint32_t b = threadIdx.x; b < total; b += WRT_PPM_BLOCK) dst[b] = text[b];
}
cudaError_t launch_format_ppm(const uint8_t* rgb, uint32_t n_pixels, uint32_t* block_bytes, unsigned long long* offsets, uint8_t* body,
                              cudaStream_t stream) {
    if (n_pixels == 0) return cudaSuccess;
    const uint32_t n_blocks = (n_pixels + WRT_PPM_BLOCK_PIXELS - 1) / WRT_PPM_BLOCK_PIXELS;
    ppm_block_bytes_kernel<<<n_blocks, WRT_PPM_BLOCK, 0, stream>>>(rgb, n_pixels, block_bytes);
    ppm_scan_kernel<<<1, 1024, 0, stream>>>(block_bytes, n_blocks, offsets);
    ppm_format_kernel<<<n_blocks, WRT_PPM_BLOCK, 0, stream>>>(rgb, n_pixels, offsets, body);
    return cudaGetLastError();
}
uint32_t ppm_block_count(uint32_t n_pixels) { return (n_pixels + WRT_PPM_BLOCK_PIXELS - 1) / WRT_PPM_BLOCK_PIXELS; }

template <int CULL>
__global__ void primary_hits_kernel(const __grid_constant__ LaunchParams LP, DeviceScene S, uint32_t n_samples, uint32_t* __restrict__ ids, double* __restrict__ ts) {
    const RenderConstants& rc = LP.rc;
    const uint64_t total = (uint64_t)rc.width * rc.height * n_samples;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t s = (uint32_t)(i % n_samples);
        uint64_t pix = i / n_samples;
        uint32_t col = (uint32_t)(pix % rc.width), row = (uint32_t)(pix / rc.width);
        Ray r = sample_ray(rc, LP.sobol, col, row, s, false, false, nullptr);
        ClosestHit ch = closest_hit_lane<CULL>(S, r.o, r.d, 0.0, 1e-4, CUDART_INF);
        if (ids) ids[i] = (ch.pc == WRT_NONE) ? WRT_NONE : __ldg(&S.ops[ch.pc].w);
        if (ts) ts[i] = (ch.pc == WRT_NONE) ? CUDART_INF : ch.t;
    }
}

template <int CULL, int TRAV>
__global__ void trace_rays_kernel(DeviceScene S, const double* __restrict__ origins, const double* __restrict__ dirs, uint64_t n,
                                  double tmin, uint32_t* ids, double* ts, double* point, double* normal, double* uv,
                                  uint32_t* front_face) {
    // whole warps iterate together so that the packet scan can be exercised through this entry point too
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_round = (n + 31) / 32 * 32;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const bool active = i < n;
        d3 o = mk(0, 0, 0), d = mk(0, 0, 1);
        if (active) { o = ld3(origins + 3 * i); d = ld3(dirs + 3 * i); }
        ClosestHit ch;
        if (TRAV == TRAV_PACKET) {
            ch = closest_hit_packet<CULL>(S, active, o, d, 0.0, tmin, CUDART_INF);
        } else {
            ch.pc = WRT_NONE;
            if (active) ch = closest_hit_lane<CULL>(S, o, d, 0.0, tmin, CUDART_INF);
        }
        if (!active) continue;
        const bool hit = ch.pc != WRT_NONE;
        HitRecord rec;
        if (hit) resolve_hit(S, ch, o, d, 0.0, true, rec);
        if (ids) ids[i] = hit ? rec.prim_id : WRT_NONE;
        if (ts) ts[i] = hit ? rec.t : CUDART_INF;
        if (point) { point[3 * i] = hit ? rec.point.x : 0; point[3 * i + 1] = hit ? rec.point.y : 0; point[3 * i + 2] = hit ? rec.point.z : 0; }
        if (normal) { normal[3 * i] = hit ? rec.normal.x : 0; normal[3 * i + 1] = hit ? rec.normal.y : 0; normal[3 * i + 2] = hit ? rec.normal.z : 0; }
        if (uv) { uv[2 * i] = hit ? rec.u : 0; uv[2 * i + 1] = hit ? rec.v : 0; }
        if (front_face) front_face[i] = hit ? rec.front_face : 0;
    }
}

__global__ void sobol_pixel_kernel(const __grid_constant__ LaunchParams LP, const uint32_t* cols, const uint32_t* rows, const uint32_t* sidx, uint64_t n, uint64_t* index_out,
                                   double* offsets) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t idx = sobol_interval_to_index(LP.sobol, sidx[i], cols[i], rows[i]);
        if (index_out) index_out[i] = idx;
        if (offsets) sobol_pixel_2d(LP.sobol, idx, cols[i], rows[i], offsets[2 * i], offsets[2 * i + 1]);
    }
}

// sampleDimension (sampler.zig:236-247) over the full 1024 x 52 table in global memory
__global__ void sobol_dimension_kernel(const uint32_t* __restrict__ matrices, const uint64_t* index, const uint32_t* dimension,
                                       uint64_t n, uint32_t owen_fast, uint32_t seed, float* out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t a = index[i];
        uint32_t dim = dimension[i];
        uint32_t v = 0;
        for (uint32_t k = dim * 52u; a != 0; a >>= 1, ++k)
            if (a & 1) v ^= __ldg(matrices + k);
        if (owen_fast) v = owen_fast_apply(murmur2_u32(dim, seed), v);
        out[i] = sobol_sample_bits_to_float(v);
    }
}

// FP64 issue-rate probe: 8 independent DFMA chains per thread (roofline denominator for the issue-bound configs,
// BASELINE.md §4).  The result is stored so the chains cannot be optimised away.
__global__ void fp64_peak_kernel(double* out, uint32_t iters, double seed) {
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
    const double m = 1.0000001, c = 1e-9;
    for (uint32_t i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}
cudaError_t launch_fp64_peak(double* out, uint32_t grid, uint32_t block, uint32_t iters, cudaStream_t stream) {
    fp64_peak_kernel<<<grid, block, 0, stream>>>(out, iters, 0.5);
    return cudaGetLastError();
}
// the same probe on the binary32 pipe (the culler's arithmetic)
__global__ void fp32_peak_kernel(double* out, uint32_t iters, float seed) {
    float a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
    const float m = 1.0000001f, c = 1e-9f;
    for (uint32_t i = 0; i < iters; ++i) {
        a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
        a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
    }
    out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = (double)(((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7)));
}
cudaError_t launch_fp32_peak(double* out, uint32_t grid, uint32_t block, uint32_t iters, cudaStream_t stream) {
    fp32_peak_kernel<<<grid, block, 0, stream>>>(out, iters, 0.5f);
    return cudaGetLastError();
}

// =============================================================================================================
// Wavefront engine: the same path integrator as render_kernel, cut into small kernels that meet at
// queues in HBM.  A pool of path slots — slot = (sample chunk, pixel), the accumulation slot of the resolve pass —
// each works through its samples one path at a time.  Per iteration:
//     wf_generate   dead slots with samples left : Sobol jitter + camera ray            -> extend queue
//     wf_extend     closest hit (packet or per-lane scan); miss => background, finish    -> surface / metal / other queue
//     wf_shade<Q>   one material class per launch, so every lane of a warp runs the same code
//                   (surface = lambertian|isotropic, metal, other = dielectric|emissive) -> next extend queue | regenerate queue
// Queue appends are warp-aggregated (one atomic per warp).  A slot is owned by exactly one thread at any time, so
// path state and the per-slot colour sums need no atomics and every slot adds its samples in sample order: the
// frame is bit-identical to render_kernel's.
// =============================================================================================================
__device__ __forceinline__ uint32_t* wf_queue(const WavefrontArgs& A, int q) { return A.queues + (size_t)q * A.capacity; }

// The path pool (128 B per slot, gigabytes) and the queues stream through L2 once per iteration; the tree records and the
// primitives are what must stay there.  Pool accesses therefore carry the evict-first hint (ld.global.cs / st.global.cs).
__device__ __forceinline__ PathState load_path(const PathState* p) {
    union { PathState s; double2 v[8]; } u;
    const double2* q = reinterpret_cast<const double2*>(p);
#pragma unroll
    for (int k = 0; k < 8; k += 2) ldcs256(q + k, u.v[k], u.v[k + 1]);
    return u.s;
}

// warp-aggregated append: one atomicAdd per warp and queue
__device__ __forceinline__ void wf_push(const WavefrontArgs& A, int q, bool pred, uint32_t slot) {
    const unsigned mask = __ballot_sync(0xffffffffu, pred);
    if (mask == 0) return;
    const uint32_t lane = threadIdx.x & 31u;
    const int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == (uint32_t)leader) base = atomicAdd(&A.counters[q], (unsigned long long)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred) wf_queue(A, q)[base + __popc(mask & ((1u << lane) - 1u))] = slot;
}

// the same, with the reordering key of the ray written beside the queue entry (extend queues only)
__device__ __forceinline__ void wf_push_keyed(const WavefrontArgs& A, int q, bool pred, uint32_t slot, uint32_t key) {
    const unsigned mask = __ballot_sync(0xffffffffu, pred);
    if (mask == 0) return;
    const uint32_t lane = threadIdx.x & 31u;
    const int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == (uint32_t)leader) base = atomicAdd(&A.counters[q], (unsigned long long)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred) {
        const size_t at = base + __popc(mask & ((1u << lane) - 1u));
        wf_queue(A, q)[at] = slot;
        A.keys[(size_t)(q - WQ_EXTEND0) * A.capacity + at] = key;
    }
}

__device__ __forceinline__ void wf_slot_pixel(const RenderConstants& rc, const WavefrontArgs& A, uint32_t slot, uint32_t& chunk,
                                              uint32_t& col, uint32_t& row) {
    const uint32_t job = A.slot_job[slot];
    chunk = job / A.n_pixels;
    const uint32_t local = job % A.n_pixels;
    const uint32_t local_row = local / rc.width;
    col = local % rc.width;
    row = rc.row_shard_index + local_row * rc.row_shard_count;
}

// A path ended: add it to its job's colour sum (render.zig:129-135).  If the job has samples left the slot regenerates; if
// not it draws the next job (its first sample regenerates); with no job left the slot retires.  Returns whether to push the
// slot to the regenerate queue (the caller does the warp-wide push).  Jobs are summed by one slot each, in sample order, so
// the frame does not depend on which slot ran which job.
__device__ __forceinline__ bool wf_finish_path(const RenderConstants& rc, const WavefrontArgs& A, uint32_t slot, PathState& P, d3 L) {
    const double scale = 1.0 / (double)rc.spp;
    const uint32_t job = A.slot_job[slot];
    double* acc = A.accum + (size_t)job * 3;
    acc[0] += L.x * scale; acc[1] += L.y * scale; acc[2] += L.z * scale;
    const uint32_t chunk = job / A.n_pixels;
    const uint32_t s_first = rc.sample_begin + chunk * rc.chunk_size;
    const uint32_t s_last = min(s_first + rc.chunk_size, rc.sample_end);
    const uint32_t next = P.sample + 1;
    if (next < s_last) {
        A.paths[slot].sample = next;
        return true;
    }
    atomicAdd(&A.shared[WS_JOBS_DONE], 1ull);  // this job is done
    const unsigned long long next_job = atomicAdd(&A.shared[WS_JOB_CURSOR], 1ull);
    if (next_job >= A.n_jobs) return false;
    A.slot_job[slot] = (uint32_t)next_job;
    double* nacc = A.accum + (size_t)next_job * 3;
    nacc[0] = 0.0; nacc[1] = 0.0; nacc[2] = 0.0;
    A.paths[slot].sample = rc.sample_begin + (uint32_t)(next_job / A.n_pixels) * rc.chunk_size;
    return true;
}

// iteration parity selects the extend / regenerate queue pair: kernels of iteration `it` read E[it&1], R[it&1] and write
// E[(it+1)&1], R[(it+1)&1] (wf_generate appends to E[it&1] before wf_extend drains it).
__global__ void __launch_bounds__(256) wf_init_kernel(const __grid_constant__ LaunchParams LP, WavefrontArgs A) {
    const RenderConstants& rc = LP.rc;
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < A.capacity; slot += gridDim.x * blockDim.x) {
        const uint32_t job = A.job_base + slot;  // the pipelines' slots start with the first jobs (sum of capacities <= n_jobs)
        A.slot_job[slot] = job;
        A.paths[slot].sample = rc.sample_begin + (job / A.n_pixels) * rc.chunk_size;  // first sample of the job (not yet generated)
        A.accum[(size_t)job * 3 + 0] = 0.0; A.accum[(size_t)job * 3 + 1] = 0.0; A.accum[(size_t)job * 3 + 2] = 0.0;
        wf_queue(A, WQ_REGEN0)[slot] = slot;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) A.counters[WQ_REGEN0] = A.capacity;  // (the host sets the shared job cursor)
}

__global__ void __launch_bounds__(256) wf_generate_kernel(const __grid_constant__ LaunchParams LP, WavefrontArgs A, DeviceScene S, uint32_t parity) {
    const RenderConstants& rc = LP.rc;
    const int q_in = WQ_REGEN0 + (int)parity, q_out = WQ_EXTEND0 + (int)parity;
    const uint32_t n = (uint32_t)A.counters[q_in];
    const uint32_t n_round = (n + 31u) & ~31u;
    Rng rng;
    rng.k0 = (uint32_t)rc.seed; rng.k1 = (uint32_t)(rc.seed >> 32);
    unsigned long long started = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        const bool valid = i < n;
        uint32_t slot = 0;
        if (valid) {
            slot = wf_queue(A, q_in)[i];
            uint32_t chunk, col, row;
            wf_slot_pixel(rc, A, slot, chunk, col, row);
            const uint32_t s = A.paths[slot].sample;
            rng.pixel = row * rc.width + col;
            rng.sample = s;
            rng.sobol = A.sobol_matrices;
            if (rng.sobol) rng.sobol_index = sobol_interval_to_index(LP.sobol, s, col, row);
            const Ray r = sample_ray(rc, LP.sobol, col, row, s, rc.dof != 0, S.has_moving != 0, &rng);
            PathState P;
            P.ox = r.o.x; P.oy = r.o.y; P.oz = r.o.z; P.dx = r.d.x; P.dy = r.d.y; P.dz = r.d.z;
            P.bx = 1.0; P.by = 1.0; P.bz = 1.0; P.lx = 0.0; P.ly = 0.0; P.lz = 0.0;
            P.t = 0.0; P.hit_pc = WRT_NONE; P.hit_xf = WRT_NONE;
            P.depth_left = rc.max_depth; P.sample = s; P.time = r.time;
            A.paths[slot] = P;
            ++started;
        }
        wf_push(A, q_out, valid, slot);
    }
    for (int off = 16; off > 0; off >>= 1) started += __shfl_down_sync(0xffffffffu, started, off);
    if ((threadIdx.x & 31u) == 0 && started) atomicAdd(&A.shared[WS_PATHS], started);
}

template <int CULL, int TRAV>
__global__ void __launch_bounds__(128) wf_extend_kernel(const __grid_constant__ LaunchParams LP, WavefrontArgs A, DeviceScene S, uint32_t parity) {
    const RenderConstants& rc = LP.rc;
    const int q_in = WQ_EXTEND0 + (int)parity, q_regen = WQ_REGEN0 + (int)(parity ^ 1u);
    const uint32_t n = (uint32_t)A.counters[q_in];
    const uint32_t n_round = (n + 31u) & ~31u;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&A.shared[WS_RAYS], (unsigned long long)n);  // rays = closest-hit queries
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        const bool valid = i < n;
        uint32_t slot = 0;
        d3 o = mk(0, 0, 0), d = mk(0, 0, 1);
        double time = 0.0;
        if (valid) {
            slot = wf_queue(A, q_in)[i];
            const double2* p = reinterpret_cast<const double2*>(A.paths + slot);
            const double2 a = p[0], b = p[1], c = p[2];
            o = mk(a.x, a.y, b.x); d = mk(b.y, c.x, c.y);
            if (S.has_moving) time = A.paths[slot].time;
        }
        ClosestHit ch;
        if (TRAV == TRAV_PACKET) {
            ch = closest_hit_packet<CULL>(S, valid, o, d, time, 1e-4, CUDART_INF);
        } else {
            ch.pc = WRT_NONE; ch.t = CUDART_INF; ch.xform = WRT_NONE;
            if (valid) ch = closest_hit_lane<CULL>(S, o, d, time, 1e-4, CUDART_INF);
        }
        int route = -1;  // shade queue
        bool regen = false;
        if (valid) {
            if (ch.pc == WRT_NONE) {  // miss: L += beta * background, path ends (render.zig:215-217)
                PathState P = A.paths[slot];
                d3 L = mk(P.lx, P.ly, P.lz) + mk(P.bx, P.by, P.bz) * ld3(rc.background);
                regen = wf_finish_path(rc, A, slot, P, L);
            } else {
                PathState* P = A.paths + slot;
                P->t = ch.t; P->hit_pc = ch.pc; P->hit_xf = ch.xform;
                const uint32_t kind = S.materials[__ldg(&S.ops[ch.pc].z)].kind;
                route = (kind == WRT_MAT_METAL) ? WQ_METAL : ((kind == WRT_MAT_LAMBERTIAN || kind == WRT_MAT_ISOTROPIC) ? WQ_SURFACE : WQ_OTHER);
            }
        }
        wf_push(A, WQ_SURFACE, route == WQ_SURFACE, slot);
        wf_push(A, WQ_METAL, route == WQ_METAL, slot);
        wf_push(A, WQ_OTHER, route == WQ_OTHER, slot);
        wf_push(A, q_regen, regen, slot);
    }
}

// Persistent form of wf_extend for the ordered per-lane traversal (large programs): every LANE owns one ray at a time and
// takes the next ray of the extend queue the moment its traversal ends, so no lane waits for the longest traversal of its
// warp (in render_kernel<.,lane> the node loop of the 2^20-primitive scene runs with 4.4 of 32 lanes: traversal lengths are
// heavy-tailed and a warp is as slow as its slowest ray).  Fetching a ray is cheap here — 48 bytes from the path pool — which
// is what a megakernel cannot offer (there a refill is a whole shading step).  One outer iteration: refill | child-pair records
// for the lanes that stand on one, while they are at least half of the live lanes | leaf ops + pops for the others | retire.
// The traversal functions are closest_hit_ordered's, the visiting order per ray is identical, so are the results.
#ifndef WRT_WF_NODE_BURST
#define WRT_WF_NODE_BURST 32
#endif
#ifndef WRT_WF_LEAF_BURST
#define WRT_WF_LEAF_BURST 4
#endif
#ifndef WRT_WF_NODE_SHIFT
#define WRT_WF_NODE_SHIFT 2  // the record phase runs while (lanes on a record) << shift >= lanes with a ray: 2 = at least a quarter
                            // (measured on C5: half / quarter, leaf burst 2 / 4, 5 / 6 / 8 blocks per SM: 499 ... 509 Mrays/s; 8 blocks 452)
#endif
#ifndef WRT_WF_EXTEND_MIN_BLOCKS
#define WRT_WF_EXTEND_MIN_BLOCKS 6  // <= 85 registers: 24 warps per SM (the kernel is bound by memory latency, not by issue)
#endif
#ifndef WRT_WF_EXTEND_DEFAULT_BLOCKS
#define WRT_WF_EXTEND_DEFAULT_BLOCKS 8  // 64 registers, 32 warps per SM (measured 6 / 7 / 8 blocks: 698 / 731 / 732 Mrays/s on C5); what the wide (four-wide records) instantiation runs with unless WRT_WF_BLOCKS says otherwise
#endif
#ifndef WRT_WF_CURSOR_CHUNK
#define WRT_WF_CURSOR_CHUNK 256u    // most queue entries a warp draws per atomic
#endif
// Output queues are fed through per-warp staging rows in shared memory: a retiring lane drops its slot there and the warp
// appends a full row of 32 with ONE atomic (the plain wf_push costs one atomic per warp per retire event, and with lanes
// retiring one or two at a time four counters were taking 10^8 atomics a second).
// A row is appended when it holds WRT_WF_STAGE_FLUSH entries; with 16 the four rows of a warp take 768 bytes, a block 3 KB (+ 1 KB
// the system reserves), and eight resident blocks fit the 32 KB shared-memory configuration — the next one (64 KB) would take 32 KB
// of L1 away from the tree records.
#ifndef WRT_WF_STAGE_FLUSH
#define WRT_WF_STAGE_FLUSH 16u
#endif
#define WRT_WF_STAGE (WRT_WF_STAGE_FLUSH + 32u)  // entries per (warp, queue) row: < FLUSH before a retire event adds <= 32
struct WfStage {
    uint32_t slot[4][WRT_WF_STAGE];
};
__device__ __forceinline__ void wf_stage_push(const WavefrontArgs& A, WfStage& st, uint32_t& count, int row, int queue, bool pred, uint32_t slot, uint32_t lane) {
    const unsigned mask = __ballot_sync(0xffffffffu, pred);
    if (mask == 0) return;
    if (pred) st.slot[row][count + __popc(mask & ((1u << lane) - 1u))] = slot;
    count += __popc(mask);
    __syncwarp();
    if (count >= WRT_WF_STAGE_FLUSH) {  // append the oldest min(count, 32), keep the rest
        const uint32_t out = min(count, 32u);
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&A.counters[queue], (unsigned long long)out);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane < out) wf_queue(A, queue)[base + lane] = st.slot[row][lane];
        __syncwarp();
        const uint32_t rest = count - out;
        uint32_t moved = 0;
        if (lane < rest) moved = st.slot[row][out + lane];
        __syncwarp();
        if (lane < rest) st.slot[row][lane] = moved;
        count = rest;
        __syncwarp();
    }
}
__device__ __forceinline__ void wf_stage_flush(const WavefrontArgs& A, WfStage& st, uint32_t count, int row, int queue, uint32_t lane) {
    if (count == 0) return;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&A.counters[queue], (unsigned long long)count);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (lane < count) wf_queue(A, queue)[base + lane] = st.slot[row][lane];
    if (lane + 32u < count) wf_queue(A, queue)[base + 32u + lane] = st.slot[row][lane + 32u];
}

template <int N> struct WfStackColumns { __device__ __forceinline__ static uint4* get() { __shared__ uint4 a[N * 128]; return a; } };
template <> struct WfStackColumns<0> { __device__ __forceinline__ static uint4* get() { return nullptr; } };
template <int WIDE, int MINB = WRT_WF_EXTEND_MIN_BLOCKS, int SMSTACK = 0, bool COMPACT = false, bool QUANT = false>
__global__ void __launch_bounds__(128, MINB) wf_extend_ordered_kernel(const __grid_constant__ LaunchParams LP, WavefrontArgs A, DeviceScene S, uint32_t parity) {
    __shared__ WfStage stage[4];  // one per warp of the block
    // SMSTACK > 0 keeps the bottom entries of every thread's stack in shared memory.  Measured on C5 (8 blocks / SM): 0 / 4 / 6 / 8
    // entries = 755 / 686 / 676 / 661 Mrays/s — the shared memory comes out of L1, and the records' L1 hits are worth more than
    // the local-memory sectors saved.  Not instantiated.
    uint4* const stack_columns = WfStackColumns<SMSTACK>::get();
    const RenderConstants& rc = LP.rc;
    const int q_in = WQ_EXTEND0 + (int)parity, q_regen = WQ_REGEN0 + (int)(parity ^ 1u);
    const uint32_t n = (uint32_t)A.counters[q_in];
    const uint32_t lane = threadIdx.x & 31u;
    WfStage& st = stage[threadIdx.x >> 5];
    uint32_t n_surface = 0, n_metal = 0, n_other = 0, n_regen = 0;  // warp-uniform fill of the staging rows
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&A.shared[WS_RAYS], (unsigned long long)n);  // rays = closest-hit queries
    const uint32_t* __restrict__ queue = wf_queue(A, q_in);
    unsigned long long* cursor = &A.counters[WF_CURSOR];
    uint32_t w_next = 0, w_end = 0;  // warp-uniform: the queue range this warp hands out to its lanes
    // chunk: a few draws per warp on a full queue, but never so large that a short queue lands on a handful of warps
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    const uint32_t chunk = max(32u, min((uint32_t)WRT_WF_CURSOR_CHUNK, (n / (2u * n_warps)) & ~31u));
    const int node_burst = A.node_burst ? (int)A.node_burst : WRT_WF_NODE_BURST, leaf_burst = A.leaf_burst ? (int)A.leaf_burst : WRT_WF_LEAF_BURST;
    const int node_shift = A.node_shift ? (int)A.node_shift - 1 : WRT_WF_NODE_SHIFT;  // (stored + 1 so that 0 means default)
    TravLean T;  // the local ray is re-formed from the path record where a leaf op needs it
    typename std::conditional<COMPACT, typename std::conditional<QUANT, TravCompactStackQ, TravCompactStack>::type,
                              typename std::conditional<SMSTACK != 0, TravHybridStack<(SMSTACK ? SMSTACK : 1), 128>, TravLocalStack>::type>::type stack;
    if constexpr (SMSTACK != 0) stack.column = stack_columns + threadIdx.x;
    uint32_t slot = 0;
    bool has = false, drained = false;
    T.node = WRT_NONE; T.pc = 0; T.end = 0; T.sp = 0;
    unsigned long long steps = 0;
    // the world-space ray is only needed again when a pop or a POP op changes the transform context (rare; never in scenes
    // without instances): it is re-read from the path pool there instead of living in 12 registers
    auto world_ray = [&](d3& wo, d3& wd, double& time) {
        const double2* p = reinterpret_cast<const double2*>(A.paths + slot);
        double2 a, b;
        ldcs256(p, a, b);
        const double2 c = __ldcs(p + 2);
        wo = mk(a.x, a.y, b.x); wd = mk(b.y, c.x, c.y);
        time = S.has_moving ? A.paths[slot].time : 0.0;
    };
    for (;;) {
        // ---- refill: the warp draws WRT_WF_CURSOR_CHUNK queue entries at a time and hands them to the lanes without a ray ----
        const bool want = !has && !drained;
        const unsigned wanting = __ballot_sync(0xffffffffu, want);
        if (wanting) {
            if (w_next >= w_end) {
                unsigned long long first = 0;
                if (lane == 0) first = atomicAdd(cursor, (unsigned long long)chunk);
                first = __shfl_sync(0xffffffffu, first, 0);
                w_next = (uint32_t)min(first, (unsigned long long)n);
                w_end = (uint32_t)min(first + chunk, (unsigned long long)n);
            }
            if (want) {
                const uint32_t mine = w_next + __popc(wanting & ((1u << lane) - 1u));
                if (mine < w_end) {
                    slot = __ldg(queue + mine);
                    d3 wo, wd;
                    double time;
                    world_ray(wo, wd, time);
                    trav_init(S, T, wo, wd, time, 1e-4, CUDART_INF);
                    has = true;
                } else if (w_end >= n) {
                    drained = true;  // the queue is exhausted (lanes beyond a chunk's end simply ask again next round)
                }
            }
            w_next = min(w_next + (uint32_t)__popc(wanting), w_end);
        }
        if (!__any_sync(0xffffffffu, has)) {
            if (__all_sync(0xffffffffu, drained)) break;
            continue;
        }
        // ---- node phase: box records, while at least half of the lanes that hold a ray stand on one ----
        const int n_has = __popc(__ballot_sync(0xffffffffu, has));
#pragma unroll 1
        for (int k = 0; k < node_burst; ++k) {
            const bool in_node = has && T.node != WRT_NONE;
            const int n_node = __popc(__ballot_sync(0xffffffffu, in_node));
            if (n_node == 0 || (k > 0 && (n_node << node_shift) < n_has)) break;
            if (in_node) { trav_record_step<WIDE>(S, T, stack); ++steps; }
        }
        // ---- leaf phase: ops of leaf ranges (binary64 primitive tests, transforms, nested roots) and pops, for the others ----
        bool done = false;
#pragma unroll 1
        for (int k = 0; k < leaf_burst; ++k) {
            const bool in_leaf = has && !done && T.node == WRT_NONE;
            if (!__any_sync(0xffffffffu, in_leaf)) break;
            if (in_leaf) { done = trav_leaf_step_lazy(S, T, stack, world_ray, 1e-4, CUDART_INF); ++steps; }
        }
        // ---- retire finished rays (warp-uniform: the staging appends are ballots) ----
        const bool fin = has && done;
        if (__any_sync(0xffffffffu, fin)) {
            int route = -1;
            bool regen = false;
            if (fin) {
                const ClosestHit ch = trav_result(T);
                if (ch.pc == WRT_NONE) {  // miss: L += beta * background, path ends (render.zig:215-217)
                    PathState P = load_path(A.paths + slot);
                    d3 L = mk(P.lx, P.ly, P.lz) + mk(P.bx, P.by, P.bz) * ld3(rc.background);
                    regen = wf_finish_path(rc, A, slot, P, L);
                } else {
                    // t | hit_pc, hit_xf: bytes 96..111 of the record, one 16-byte streaming store
                    __stcs(reinterpret_cast<double2*>(A.paths + slot) + 6,
                           make_double2(ch.t, __longlong_as_double((long long)(((unsigned long long)ch.xform << 32) | ch.pc))));
                    const uint32_t kind = S.materials[__ldg(&S.ops[ch.pc].z)].kind;
                    route = (kind == WRT_MAT_METAL) ? WQ_METAL : ((kind == WRT_MAT_LAMBERTIAN || kind == WRT_MAT_ISOTROPIC) ? WQ_SURFACE : WQ_OTHER);
                }
                has = false;
            }
            wf_stage_push(A, st, n_surface, 0, WQ_SURFACE, route == WQ_SURFACE, slot, lane);
            wf_stage_push(A, st, n_metal, 1, WQ_METAL, route == WQ_METAL, slot, lane);
            wf_stage_push(A, st, n_other, 2, WQ_OTHER, route == WQ_OTHER, slot, lane);
            wf_stage_push(A, st, n_regen, 3, q_regen, regen, slot, lane);
        }
    }
    wf_stage_flush(A, st, n_surface, 0, WQ_SURFACE, lane);
    wf_stage_flush(A, st, n_metal, 1, WQ_METAL, lane);
    wf_stage_flush(A, st, n_other, 2, WQ_OTHER, lane);
    wf_stage_flush(A, st, n_regen, 3, q_regen, lane);
    for (int off = 16; off > 0; off >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, off);
    if (lane == 0 && steps) atomicAdd(&A.shared[WS_STEPS], steps);
}

template <int QUEUE, bool MANY_LIGHTS = false>
__global__ void __launch_bounds__(128) wf_shade_kernel(const __grid_constant__ LaunchParams LP, WavefrontArgs A, DeviceScene S, uint32_t parity) {
    const RenderConstants& rc = LP.rc;
    const int q_extend = WQ_EXTEND0 + (int)(parity ^ 1u), q_regen = WQ_REGEN0 + (int)(parity ^ 1u);
    const uint32_t n = (uint32_t)A.counters[QUEUE];
    const uint32_t n_round = (n + 31u) & ~31u;
    Rng rng;
    rng.k0 = (uint32_t)rc.seed; rng.k1 = (uint32_t)(rc.seed >> 32);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        const bool valid = i < n;
        uint32_t slot = 0, key = 0;
        bool extend = false, regen = false;
        if (valid) {
            slot = wf_queue(A, QUEUE)[i];
            PathState P = load_path(A.paths + slot);
            uint32_t chunk, col, row;
            wf_slot_pixel(rc, A, slot, chunk, col, row);
            rng.pixel = row * rc.width + col;
            rng.sample = P.sample;
            rng.sobol = A.sobol_matrices;
            if (rng.sobol) rng.sobol_index = sobol_interval_to_index(LP.sobol, P.sample, col, row);
            Ray ray;
            ray.o = mk(P.ox, P.oy, P.oz); ray.d = mk(P.dx, P.dy, P.dz); ray.time = P.time;
            d3 beta = mk(P.bx, P.by, P.bz), L = mk(P.lx, P.ly, P.lz);
            ClosestHit ch;
            ch.t = P.t; ch.pc = P.hit_pc; ch.xform = P.hit_xf;
            HitRecord rec;
            resolve_hit(S, ch, ray.o, ray.d, ray.time, false, rec);
            const Material M = S.materials[rec.material];
            const uint32_t bounce = rc.max_depth - P.depth_left;
            bool cont;
            if (QUEUE == WQ_OTHER) {
                if (M.kind == WRT_MAT_DIELECTRIC) cont = shade_dielectric(rec, M, ray, rng, bounce);
                else cont = shade_emissive(S, rec, M, beta, L);
            } else {
                cont = shade_surface<MANY_LIGHTS>(S, rec, M, ray, beta, L, rng, bounce);
            }
            const uint32_t depth_left = P.depth_left - 1;
            if (!cont || depth_left == 0) {
                if (cont) L = L + beta * 0.0;  // depth exhausted: the tail returns 0 (render.zig:199), times the weight
                regen = wf_finish_path(rc, A, slot, P, L);
            } else {
                PathState* out = A.paths + slot;
                double2* w = reinterpret_cast<double2*>(out);
                __stcs(w + 0, make_double2(ray.o.x, ray.o.y)); __stcs(w + 1, make_double2(ray.o.z, ray.d.x)); __stcs(w + 2, make_double2(ray.d.y, ray.d.z));
                if (QUEUE != WQ_OTHER) {  // dielectric attenuation is (1,1,1) and it gathers nothing
                    __stcs(w + 3, make_double2(beta.x, beta.y)); __stcs(w + 4, make_double2(beta.z, L.x)); __stcs(w + 5, make_double2(L.y, L.z));
                }
                out->depth_left = depth_left;
                extend = true;
                if (A.sort_buckets) {
                    const uint32_t oct = (ray.d.x < 0.0 ? 1u : 0u) | (ray.d.y < 0.0 ? 2u : 0u) | (ray.d.z < 0.0 ? 4u : 0u);
                    key = min(((ch.pc >> A.sort_shift) << 3) | oct, A.sort_buckets - 1u);
                }
            }
        }
        if (A.sort_buckets) wf_push_keyed(A, q_extend, extend, slot, key);
        else wf_push(A, q_extend, extend, slot);
        wf_push(A, q_regen, regen, slot);
    }
}

// ---- ray reordering: one counting-sort pass over the extend queue the shade kernels filled (WavefrontArgs::keys) ----------
__global__ void __launch_bounds__(256) wf_sort_hist_kernel(WavefrontArgs A, uint32_t parity) {
    const uint32_t n = (uint32_t)A.counters[WQ_EXTEND0 + parity];
    const uint32_t* __restrict__ keys = A.keys + (size_t)parity * A.capacity;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) atomicAdd(&A.sort_hist[keys[i]], 1u);
}
__global__ void __launch_bounds__(1024) wf_sort_scan_kernel(WavefrontArgs A) {
    __shared__ uint32_t warp_sums[32];
    const uint32_t nb = A.sort_buckets;
    const uint32_t per = (nb + 1023u) / 1024u;
    const uint32_t lo = min(threadIdx.x * per, nb), hi = min(lo + per, nb);
    uint32_t sum = 0;
    for (uint32_t b = lo; b < hi; ++b) sum += A.sort_hist[b];
    uint32_t incl = sum;
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    for (int m = 1; m < 32; m <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, m); if (lane >= (uint32_t)m) incl += v; }
    if (lane == 31u) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
        uint32_t ws = warp_sums[lane];
        for (int m = 1; m < 32; m <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, ws, m); if (lane >= (uint32_t)m) ws += v; }
        warp_sums[lane] = ws;
    }
    __syncthreads();
    uint32_t run = incl - sum + (w ? warp_sums[w - 1] : 0u);
    for (uint32_t b = lo; b < hi; ++b) { const uint32_t c = A.sort_hist[b]; A.sort_hist[b] = run; run += c; }
}
__global__ void __launch_bounds__(256) wf_sort_scatter_kernel(WavefrontArgs A, uint32_t parity) {
    const uint32_t n = (uint32_t)A.counters[WQ_EXTEND0 + parity];
    const uint32_t* __restrict__ keys = A.keys + (size_t)parity * A.capacity;
    const uint32_t* __restrict__ queue = wf_queue(A, WQ_EXTEND0 + (int)parity);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        A.sort_tmp[atomicAdd(&A.sort_hist[keys[i]], 1u)] = queue[i];
}
__global__ void __launch_bounds__(256) wf_sort_copy_kernel(WavefrontArgs A, uint32_t parity) {
    const uint32_t n = (uint32_t)A.counters[WQ_EXTEND0 + parity];
    uint32_t* queue = wf_queue(A, WQ_EXTEND0 + (int)parity);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) queue[i] = A.sort_tmp[i];
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < A.sort_buckets; b += gridDim.x * blockDim.x) A.sort_hist[b] = 0;
}

// reset the queues consumed by iteration `parity` so the next iteration can append to them
__global__ void wf_reset_kernel(const __grid_constant__ LaunchParams LP, WavefrontArgs A, uint32_t parity) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        A.counters[WQ_EXTEND0 + parity] = 0;
        A.counters[WQ_REGEN0 + parity] = 0;
        A.counters[WQ_SURFACE] = 0;
        A.counters[WQ_METAL] = 0;
        A.counters[WQ_OTHER] = 0;
        A.counters[WF_CURSOR] = 0;
    }
}

// ---- launchers -------------------------------------------------------------------------------------------
template <typename F>
static cudaError_t dispatch(uint32_t cull_mode, bool packet, F&& f) {
    if (cull_mode == WRT_CULL_REFERENCE) {
        if (packet) f(std::integral_constant<int, WRT_CULL_REFERENCE>(), std::integral_constant<int, TRAV_PACKET>());
        else f(std::integral_constant<int, WRT_CULL_REFERENCE>(), std::integral_constant<int, TRAV_LANE>());
    } else {
        if (packet) f(std::integral_constant<int, WRT_CULL_TIGHT>(), std::integral_constant<int, TRAV_PACKET>());
        else f(std::integral_constant<int, WRT_CULL_TIGHT>(), std::integral_constant<int, TRAV_LANE>());
    }
    return cudaGetLastError();
}

cudaError_t launch_render(const LaunchParams& lp, const DeviceScene& S, uint32_t cull_mode, bool packet, uint32_t grid, double* accum, unsigned long long* counters,
                          cudaStream_t stream) {
    if (!packet && cull_mode != WRT_CULL_REFERENCE && S.use_wide) {
        render_kernel<WRT_CULL_TIGHT, TRAV_LANE_WIDE><<<grid, WRT_RENDER_BLOCK, 0, stream>>>(lp, S, accum, counters);
        return cudaGetLastError();
    }
    return dispatch(cull_mode, packet, [&](auto c, auto t) {
        render_kernel<decltype(c)::value, decltype(t)::value><<<grid, WRT_RENDER_BLOCK, 0, stream>>>(lp, S, accum, counters);
    });
}
cudaError_t launch_render_regroup(const LaunchParams& lp, const DeviceScene& S, uint32_t cull_mode, uint32_t grid, double* accum, unsigned long long* counters,
                                  cudaStream_t stream) {
    const size_t smem = sizeof(RegroupSmem);
    cudaError_t e;
    if (cull_mode == WRT_CULL_REFERENCE) {
        e = cudaFuncSetAttribute(render_kernel_regroup<WRT_CULL_REFERENCE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        render_kernel_regroup<WRT_CULL_REFERENCE><<<grid, WRT_REGROUP_BLOCK, smem, stream>>>(lp, S, accum, counters);
    } else {
        e = cudaFuncSetAttribute(render_kernel_regroup<WRT_CULL_TIGHT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        render_kernel_regroup<WRT_CULL_TIGHT><<<grid, WRT_REGROUP_BLOCK, smem, stream>>>(lp, S, accum, counters);
    }
    return cudaGetLastError();
}
cudaError_t launch_render_sync(const LaunchParams& lp, const DeviceScene& S, uint32_t cull_mode, bool packet, uint32_t grid, double* accum,
                               unsigned long long* counters, cudaStream_t stream) {
    return dispatch(cull_mode, packet, [&](auto c, auto t) {
        render_kernel_sync<decltype(c)::value, decltype(t)::value><<<grid, WRT_SYNC_BLOCK, 0, stream>>>(lp, S, accum, counters);
    });
}
cudaError_t render_occupancy(const DeviceScene& S, uint32_t cull_mode, bool packet, int* blocks_per_sm) {
    cudaError_t err = cudaSuccess;
    if (!packet && cull_mode != WRT_CULL_REFERENCE && S.use_wide)
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, render_kernel<WRT_CULL_TIGHT, TRAV_LANE_WIDE>, WRT_RENDER_BLOCK, 0);
    dispatch(cull_mode, packet, [&](auto c, auto t) {
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, render_kernel<decltype(c)::value, decltype(t)::value>,
                                                            WRT_RENDER_BLOCK, 0);
    });
    return err;
}
cudaError_t launch_resolve(const double* accum, uint32_t n_chunks, uint32_t n_pixels, const double clear[3], int no_clear, double* fb,
                           uint32_t stride_doubles, uint8_t* rgb8, cudaStream_t stream) {
    if (n_pixels == 0) return cudaSuccess;
    resolve_kernel<<<(n_pixels + 255) / 256, 256, 0, stream>>>(accum, n_chunks, n_pixels, clear[0], clear[1], clear[2], no_clear, fb,
                                                                 stride_doubles, rgb8);
    return cudaGetLastError();
}
cudaError_t launch_encode(const double* fb, uint32_t stride_doubles, uint32_t n_pixels, uint8_t* rgb8, cudaStream_t stream) {
    if (n_pixels == 0) return cudaSuccess;
    encode_kernel<<<(n_pixels + 255) / 256, 256, 0, stream>>>(fb, stride_doubles, n_pixels, rgb8);
    return cudaGetLastError();
}
cudaError_t launch_primary_hits(const LaunchParams& lp, const DeviceScene& S, uint32_t cull_mode, uint32_t n_samples, uint32_t* ids, double* ts, uint32_t grid,
                                cudaStream_t stream) {
    if (cull_mode == WRT_CULL_REFERENCE) primary_hits_kernel<WRT_CULL_REFERENCE><<<grid, 128, 0, stream>>>(lp, S, n_samples, ids, ts);
    else primary_hits_kernel<WRT_CULL_TIGHT><<<grid, 128, 0, stream>>>(lp, S, n_samples, ids, ts);
    return cudaGetLastError();
}
cudaError_t launch_trace_rays(const DeviceScene& S, uint32_t cull_mode, bool packet, const double* origins, const double* dirs, uint64_t n,
                              double tmin, uint32_t* ids, double* ts, double* point, double* normal, double* uv, uint32_t* front_face,
                              uint32_t grid, cudaStream_t stream) {
    return dispatch(cull_mode, packet, [&](auto c, auto t) {
        trace_rays_kernel<decltype(c)::value, decltype(t)::value><<<grid, 128, 0, stream>>>(S, origins, dirs, n, tmin, ids, ts, point, normal,
                                                                                              uv, front_face);
    });
}
cudaError_t launch_sobol_pixel(const LaunchParams& lp, const uint32_t* cols, const uint32_t* rows, const uint32_t* sidx, uint64_t n, uint64_t* index_out,
                               double* offsets, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    uint32_t grid = (uint32_t)((n + 127) / 128 < 4096 ? (n + 127) / 128 : 4096);
    sobol_pixel_kernel<<<grid, 128, 0, stream>>>(lp, cols, rows, sidx, n, index_out, offsets);
    return cudaGetLastError();
}
cudaError_t launch_sobol_dimension(const uint32_t* matrices, const uint64_t* index, const uint32_t* dimension, uint64_t n,
                                   uint32_t owen_fast, uint32_t seed, float* out, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    uint32_t grid = (uint32_t)((n + 127) / 128 < 4096 ? (n + 127) / 128 : 4096);
    sobol_dimension_kernel<<<grid, 128, 0, stream>>>(matrices, index, dimension, n, owen_fast, seed, out);
    return cudaGetLastError();
}

cudaError_t wf_launch_init(const LaunchParams& lp, const WavefrontArgs& A, uint32_t grid, cudaStream_t stream) {
    wf_init_kernel<<<grid, 256, 0, stream>>>(lp, A);
    return cudaGetLastError();
}
// One wavefront iteration: generate -> extend -> shade (surface, metal, other) -> reset of the consumed queues.
// resident blocks per SM the wide extend kernel is compiled for: WRT_WF_BLOCKS = 6 (80 registers), 7 (72) or 8 (64)
static int wf_extend_min_blocks() {
    static const int v = [] {
        const char* env = std::getenv("WRT_WF_BLOCKS");
        const int b = env ? std::atoi(env) : WRT_WF_EXTEND_DEFAULT_BLOCKS;
        return (b == 6 || b == 7 || b == 8) ? b : WRT_WF_EXTEND_DEFAULT_BLOCKS;
    }();
    return v;
}
cudaError_t wf_extend_occupancy(int* blocks_per_sm) {
    const int b = wf_extend_min_blocks();
    if (b == 8) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, wf_extend_ordered_kernel<1, 8>, 128, 0);
    if (b == 7) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, wf_extend_ordered_kernel<1, 7>, 128, 0);
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, wf_extend_ordered_kernel<1>, 128, 0);
}
cudaError_t wf_launch_iteration(const LaunchParams& lp, const WavefrontArgs& A, const DeviceScene& S, uint32_t cull_mode, bool packet, uint32_t parity,
                                uint32_t grid, uint32_t persist_grid, cudaStream_t stream) {
    if (A.sort_buckets) {  // reorder what the last iteration's shade kernels queued; wf_generate appends the camera rays behind it
        wf_sort_hist_kernel<<<grid, 256, 0, stream>>>(A, parity);
        wf_sort_scan_kernel<<<1, 1024, 0, stream>>>(A);
        wf_sort_scatter_kernel<<<grid, 256, 0, stream>>>(A, parity);
        wf_sort_copy_kernel<<<grid, 256, 0, stream>>>(A, parity);
    }
    wf_generate_kernel<<<grid, 256, 0, stream>>>(lp, A, S, parity);
    const bool ordered = !packet && cull_mode != WRT_CULL_REFERENCE && S.use_ordered;
    if (ordered) {  // persistent lanes with ray replacement (one wave of resident blocks)
        if (S.use_wide) {
            const int b = wf_extend_min_blocks();
            if (S.compact_ok && S.nodes4q && b >= 7) wf_extend_ordered_kernel<1, 8, 0, true, true><<<persist_grid, 128, 0, stream>>>(lp, A, S, parity);
            else if (S.compact_ok && b >= 7) wf_extend_ordered_kernel<1, 8, 0, true><<<persist_grid, 128, 0, stream>>>(lp, A, S, parity);
            else if (S.compact_ok) wf_extend_ordered_kernel<1, 6, 0, true><<<persist_grid, 128, 0, stream>>>(lp, A, S, parity);
            else if (b == 8) wf_extend_ordered_kernel<1, 8><<<persist_grid, 128, 0, stream>>>(lp, A, S, parity);
            else if (b == 7) wf_extend_ordered_kernel<1, 7><<<persist_grid, 128, 0, stream>>>(lp, A, S, parity);
            else wf_extend_ordered_kernel<1><<<persist_grid, 128, 0, stream>>>(lp, A, S, parity);
        }
        else wf_extend_ordered_kernel<0><<<persist_grid, 128, 0, stream>>>(lp, A, S, parity);
    } else {
        dispatch(cull_mode, packet, [&](auto c, auto t) {
            wf_extend_kernel<decltype(c)::value, decltype(t)::value><<<grid * 2, 128, 0, stream>>>(lp, A, S, parity);
        });
    }
    if (packet) {
        wf_shade_kernel<WQ_SURFACE, false><<<grid * 2, 128, 0, stream>>>(lp, A, S, parity);
        wf_shade_kernel<WQ_METAL, false><<<grid * 2, 128, 0, stream>>>(lp, A, S, parity);
    } else {  // large scenes: the light-list pdf skips lights whose box the ray misses (same sums)
        wf_shade_kernel<WQ_SURFACE, true><<<grid * 2, 128, 0, stream>>>(lp, A, S, parity);
        wf_shade_kernel<WQ_METAL, true><<<grid * 2, 128, 0, stream>>>(lp, A, S, parity);
    }
    wf_shade_kernel<WQ_OTHER><<<grid * 2, 128, 0, stream>>>(lp, A, S, parity);
    wf_reset_kernel<<<1, 32, 0, stream>>>(lp, A, parity);
    return cudaGetLastError();
}

}  // namespace wrt
