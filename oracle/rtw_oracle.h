/*
 * rtw_oracle.h -- CPU ORACLE for the render() -> ray_color() -> hit()/scatter() hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library, and only as the
 * checker / the timed CPU baseline.  The product path (raytracingweekend.jl_b200/csrc) never
 * links, loads or calls anything in oracle/.
 *
 * It is an independent plain-C restatement of the reference's algorithm
 * (claforte/RayTracingWeekend.jl @ fe20135d), written from the Julia sources:
 *   src/render.jl:8-44, src/ray_color.jl:1-38, src/hit.jl:3-50, src/material.jl:13-53,
 *   src/light.jl:6-25, src/rand.jl:15-38, src/camera.jl:43-48, src/vec.jl:19-22.
 * The reference itself (Julia >= 1.6 + registry packages) cannot run in this image, so the
 * restatement is pinned against every known answer the reference's own tests / notebook hold
 * (tests/test_oracle_known_answers.py):
 *   reflect  test/runtests.jl:180, src/pluto_RayTracingWeekend.jl:382
 *   refract  src/pluto_RayTracingWeekend.jl:603-615 (x3)
 *   near_zero test/runtests.jl:131
 *   hit_sphere2 closed form test/runtests.jl:99-111
 * IMAGE PARITY IS UNPINNED BY THE REFERENCE: it holds no golden image and no RNG known answer
 * (SURVEY.md section 8c).  The RandomNumbers.jl Xoroshiro128Plus stream (rng_mode 1) is restated
 * from the published xoroshiro128+ algorithm and is UNVERIFIED against Julia.
 *
 * Floating-point contract ("one legal @fastmath evaluation of the reference expressions",
 * shared with the CUDA path so that images are comparable path-for-path):
 *   dot(a,b)      = fma(a.z,b.z, fma(a.y,b.y, a.x*b.x))
 *   a - b*c       = fma(-b,c,a)      a + b*c = fma(b,c,a)         (single rounding)
 *   sqrt, /       = IEEE-754 correctly rounded
 *   normalize(v)  = v * (1/sqrt(dot(v,v)))           (StaticArrays 1.2.13: inv(norm(v))*v)
 *   x^5           = x2=x*x; x4=x2*x2; x4*x           (llvm powi expansion)
 *   colour math   = Float64 even for Float32 scenes, because skycolor's constants are Float64
 *                   literals (src/ray_color.jl:2-3) and promote everything above them;
 *                   only the final img[i,j] store rounds to T (src/render.jl:40).
 * Two instantiations: T=float (the hot path) and T=double (used for the Float64 known answers).
 */
#ifndef RTW_ORACLE_H
#define RTW_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* RNG stream selectors */
#define RTWO_RNG_PHILOX 0     /* production stream: Philox4x32-7 addressed by (pixel, sample, event, draw) */
#define RTWO_PHILOX_ROUNDS 7  /* rounds of the production stream: the smallest Crush-resistant count (Salmon et al. 2011, table 2) */
#define RTWO_RNG_XOROSHIRO 1  /* reference-shaped stream: one sequential xoroshiro128+ per thread (UNVERIFIED vs Julia) */

/* material kinds (flattened Material{T} subtypes, src/material.jl:3,25,37) */
#define RTWO_LAMBERTIAN 0u
#define RTWO_METAL 1u
#define RTWO_DIELECTRIC 2u

/* Camera{Float32} field order, src/camera.jl:1-10 (22 x f32 = 88 bytes) */
typedef struct {
    float origin[3];
    float lower_left_corner[3];
    float horizontal[3];
    float vertical[3];
    float u[3];
    float v[3];
    float w[3];
    float lens_radius;
} rtwo_camera_f32;

typedef struct {
    double origin[3];
    double lower_left_corner[3];
    double horizontal[3];
    double vertical[3];
    double u[3];
    double v[3];
    double w[3];
    double lens_radius;
} rtwo_camera_f64;

typedef struct {
    uint64_t paths;         /* W*H*spp                                             */
    uint64_t ray_segments;  /* executions of hit(world, r, ...), src/ray_color.jl:19 */
    uint64_t sphere_tests;  /* executions of hit(::Sphere), src/hit.jl:12          */
    double seconds;         /* wall time of the render loop                        */
    int threads;            /* worker threads actually used                        */
} rtwo_stats;

/* image height of render(): image_width div (16//9), src/render.jl:11-12 */
int rtwo_image_height(int image_width);

/* ---- scalar building blocks exposed for the known-answer tests (f64 and f32) ---- */
void rtwo_reflect_f64(const double v[3], const double n[3], double out[3]);      /* src/light.jl:6 */
void rtwo_reflect_f32(const float v[3], const float n[3], float out[3]);
void rtwo_refract_f64(const double d[3], const double n[3], double ratio, double out[3]); /* src/light.jl:12-17 */
void rtwo_refract_f32(const float d[3], const float n[3], float ratio, float out[3]);
double rtwo_reflectance_f64(double cos_theta, double ratio);                       /* src/light.jl:19-25 */
float rtwo_reflectance_f32(float cos_theta, float ratio);
int rtwo_near_zero_f64(const double v[3]);                                         /* src/vec.jl:20 */
int rtwo_near_zero_f32(const float v[3]);
/* hit(::Sphere): returns 1 and writes t, p, n (face-corrected), front_face; 0 on miss. src/hit.jl:12-35, 6-10 */
int rtwo_hit_sphere_f64(const double center[3], double radius, const double o[3], const double d[3],
                        double tmin, double tmax, double* t, double p[3], double n[3], int* front_face);
int rtwo_hit_sphere_f32(const float center[3], float radius, const float o[3], const float d[3],
                        float tmin, float tmax, float* t, float p[3], float n[3], int* front_face);
void rtwo_skycolor_f32(const float dir[3], double out[3]);                         /* src/ray_color.jl:1-6 */
void rtwo_skycolor_f64(const double dir[3], double out[3]);

/* Philox4x32-R (Salmon et al. 2011), one block: R rounds; _10 = Random123's default (known-answer vectors), _7 = the
 * production stream (RTWO_PHILOX_ROUNDS; known-answer vectors too) */
void rtwo_philox4x32(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4]);
void rtwo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void rtwo_philox4x32_7(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* draws 0..n-1 (uniforms in [0,1)) of event `event` of path (pixel, sample) in the production stream */
void rtwo_path_stream_f32(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t event, int n, float* out);
/* first n outputs of xoroshiro128+ seeded like RandomNumbers.jl Xoroshiro128Plus(seed) (UNVERIFIED) */
void rtwo_xoroshiro_u64(uint64_t seed, int n, uint64_t* out);
void rtwo_xoroshiro_f32(uint64_t seed, int n, float* out);

/* one path, by explicit (row i0, col j0, sample s0), 0-based; linear colour (Float64) out */
/* debugging aid: the same path, with every segment recorded -- 8 doubles each: origin, direction, index of the
 * closest sphere (-1 = miss), its t */
void rtwo_path_trace_f32(const float* geom4, const float* mat4, const uint32_t* kind, uint32_t n_spheres,
                         const rtwo_camera_f32* cam, int image_width, int max_depth, uint64_t seed, int i0, int j0, int s0,
                         double rgb[3], double* trace, int trace_cap, int* trace_n);
void rtwo_path_f32(const float* geom4, const float* mat4, const uint32_t* kind, uint32_t n_spheres,
                   const rtwo_camera_f32* cam, int image_width, int max_depth, uint64_t seed,
                   int i0, int j0, int s0, double rgb[3], uint32_t* segments);

/*
 * render(scene, cam, image_width, n_samples), src/render.jl:8-44.
 *   geom4 : n x {cx,cy,cz,radius}      mat4 : n x {albedo r,g,b, fuzz|ir|0}     kind : n x u32
 *   out_rgb      : H*W*3 T, Julia column-major Matrix{RGB{T}}: pixel (i0,j0) at ((j0*H)+i0)*3, post-gamma
 *   out_linear   : optional H*W*3 doubles (same layout): accum_color / n_samples before sqrt
 *   max_depth    : ray_color's `depth` (reference default 16, src/ray_color.jl:14)
 *   n_threads    : <=0 -> all online cores.  With RTWO_RNG_PHILOX the image is thread-count independent.
 *   row_start/row_stride : render only rows i0 = row_start, row_start+row_stride, ... (others left untouched)
 * Returns 0 on success, <0 on bad arguments.
 */
int rtwo_render_f32(const float* geom4, const float* mat4, const uint32_t* kind, uint32_t n_spheres,
                    const rtwo_camera_f32* cam, int image_width, int n_samples, int max_depth,
                    uint64_t seed, int rng_mode, int n_threads, int row_start, int row_stride,
                    float* out_rgb, double* out_linear, rtwo_stats* stats);

int rtwo_render_f64(const double* geom4, const double* mat4, const uint32_t* kind, uint32_t n_spheres,
                    const rtwo_camera_f64* cam, int image_width, int n_samples, int max_depth,
                    uint64_t seed, int rng_mode, int n_threads, int row_start, int row_stride,
                    double* out_rgb, double* out_linear, rtwo_stats* stats);

#ifdef __cplusplus
}
#endif
#endif
