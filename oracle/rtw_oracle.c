/*
 * rtw_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE; see the header comment of rtw_oracle.h).
 * Plain C11 + pthreads.  Build: see oracle/Makefile (-ffp-contract=off: every fused multiply-add
 * in the FP contract is an explicit fma()/fmaf(); nothing else may be contracted).
 */
#define _GNU_SOURCE
#include "rtw_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

/* ------------------------------------------------------------------ RNG streams */

typedef struct {
    int mode;
    /* Philox4x32-7, addressed by (pixel, sample, event, word position) */
    uint32_t key[2];
    uint32_t pixel, sample, event;
    uint32_t pos;       /* next word of the current event's word sequence: block = pos / 4, word = pos % 4 */
    uint32_t buf[4];
    uint32_t buf_block;
    int buf_valid;
    /* xoroshiro128+ */
    uint64_t x, y;
} rtwo_rng;

#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void rtwo_philox4x32(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < rounds; ++r) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
void rtwo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { rtwo_philox4x32(ctr, key, 10, out); }
void rtwo_philox4x32_7(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { rtwo_philox4x32(ctr, key, RTWO_PHILOX_ROUNDS, out); }

/*
 * Production stream (documented in DESIGN.md "RNG stream"): every uniform is ADDRESSED, not drawn from a running
 * sequence, so a GPU lane can produce any of them without carrying generator state:
 *   key     = (seed lo32, seed hi32)
 *   counter = (block, sample s0, pixel i0*W+j0, event)
 *   event 0 = the primary ray (src/render.jl:30-37):  draws 0,1 = jitter du, dv (unused for the first sample);
 *             disk attempt k (src/rand.jl:31-38) = draws 2+2k, 3+2k
 *   event e >= 1 = scatter() at the e-th hit of the path (src/material.jl):
 *             ball attempt a (src/rand.jl:15-22) = draws 4a, 4a+1, 4a+2  (x, y, z);  draw 3 = the dielectric coin
 *   draw n of an event is word (n mod 4) of block (n div 4) for Float32 (Float64 takes two words per draw);
 *   f32 = (word >> 9) * 2^-23.
 */
static inline void rtwo_rng_begin_path(rtwo_rng* g, uint64_t seed, uint32_t pixel, uint32_t sample) {
    g->key[0] = (uint32_t)seed;
    g->key[1] = (uint32_t)(seed >> 32);
    g->pixel = pixel;
    g->sample = sample;
    g->event = 0;
    g->pos = 0;
    g->buf_valid = 0;
}

/* start the next event (no-op for the sequential xoroshiro stream) */
static inline void rtwo_rng_next_event(rtwo_rng* g) {
    g->event += 1;
    g->pos = 0;
    g->buf_valid = 0;
}

/* position the stream at draw `draw` of the current event (no-op for the sequential xoroshiro stream) */
static inline void rtwo_rng_seek(rtwo_rng* g, uint32_t draw, uint32_t words_per_draw) { g->pos = draw * words_per_draw; }

static inline uint64_t rotl64(uint64_t v, int k) { return (v << k) | (v >> (64 - k)); }

/* xoroshiro128+ (Blackman & Vigna 2016 constants 55,14,36 as used by RandomNumbers.jl 1.5.x) -- UNVERIFIED vs Julia */
static inline uint64_t xoroshiro_next(rtwo_rng* g) {
    uint64_t s0 = g->x, s1 = g->y;
    uint64_t r = s0 + s1;
    s1 ^= s0;
    g->x = rotl64(s0, 55) ^ s1 ^ (s1 << 14);
    g->y = rotl64(s1, 36);
    return r;
}

static inline uint64_t splitmix64(uint64_t* s) {
    uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

/* Xoroshiro128Plus(seed::Integer): SplitMix64 expansion to two words, then one warm-up step -- UNVERIFIED vs Julia */
static inline void rtwo_rng_seed_xoroshiro(rtwo_rng* g, uint64_t seed) {
    memset(g, 0, sizeof *g);
    g->mode = RTWO_RNG_XOROSHIRO;
    uint64_t s = seed;
    g->x = splitmix64(&s);
    g->y = splitmix64(&s);
    (void)xoroshiro_next(g);
}

static inline uint32_t rtwo_next_u32(rtwo_rng* g) {
    if (g->mode == RTWO_RNG_XOROSHIRO) return (uint32_t)xoroshiro_next(g); /* rand(rng, UInt64) % UInt32 */
    uint32_t blk = g->pos >> 2;
    if (!g->buf_valid || g->buf_block != blk) {
        uint32_t ctr[4] = {blk, g->sample, g->pixel, g->event};
        rtwo_philox4x32(ctr, g->key, RTWO_PHILOX_ROUNDS, g->buf);
        g->buf_block = blk;
        g->buf_valid = 1;
    }
    return g->buf[(g->pos++) & 3u];
}

static inline uint64_t rtwo_next_u64(rtwo_rng* g) {
    if (g->mode == RTWO_RNG_XOROSHIRO) return xoroshiro_next(g);
    uint64_t lo = rtwo_next_u32(g);
    uint64_t hi = rtwo_next_u32(g);
    return (hi << 32) | lo;
}

/* ------------------------------------------------------------------ two instantiations of the generic body */

#define RT float
#define RT_IS_F32 1
#define RT_WORDS_PER_DRAW 1u
#define SFX(x) x##_f32
#define FMA fmaf
#define SQRT sqrtf
#define FABS fabsf
#define FMIN fminf
#include "rtw_oracle_impl.h"
#undef RT
#undef RT_IS_F32
#undef RT_WORDS_PER_DRAW
#undef SFX
#undef FMA
#undef SQRT
#undef FABS
#undef FMIN

#define RT double
#define RT_IS_F32 0
#define RT_WORDS_PER_DRAW 2u
#define SFX(x) x##_f64
#define FMA fma
#define SQRT sqrt
#define FABS fabs
#define FMIN fmin
#include "rtw_oracle_impl.h"
#undef RT
#undef RT_IS_F32
#undef RT_WORDS_PER_DRAW
#undef SFX
#undef FMA
#undef SQRT
#undef FABS
#undef FMIN

/* ------------------------------------------------------------------ exported scalar blocks */

int rtwo_image_height(int image_width) {
    /* image_width div (16//9) == floor(image_width*9/16), src/render.jl:11-12 */
    return (int)(((long long)image_width * 9) / 16);
}

void rtwo_reflect_f64(const double v[3], const double n[3], double out[3]) {
    v3_f64 r = reflect_f64(mk_f64(v[0], v[1], v[2]), mk_f64(n[0], n[1], n[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void rtwo_reflect_f32(const float v[3], const float n[3], float out[3]) {
    v3_f32 r = reflect_f32(mk_f32(v[0], v[1], v[2]), mk_f32(n[0], n[1], n[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void rtwo_refract_f64(const double d[3], const double n[3], double ratio, double out[3]) {
    v3_f64 r = refract_f64(mk_f64(d[0], d[1], d[2]), mk_f64(n[0], n[1], n[2]), ratio);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void rtwo_refract_f32(const float d[3], const float n[3], float ratio, float out[3]) {
    v3_f32 r = refract_f32(mk_f32(d[0], d[1], d[2]), mk_f32(n[0], n[1], n[2]), ratio);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
double rtwo_reflectance_f64(double c, double ratio) { return reflectance_f64(c, ratio); }
float rtwo_reflectance_f32(float c, float ratio) { return reflectance_f32(c, ratio); }
int rtwo_near_zero_f64(const double v[3]) { return near_zero_f64(mk_f64(v[0], v[1], v[2])); }
int rtwo_near_zero_f32(const float v[3]) { return near_zero_f32(mk_f32(v[0], v[1], v[2])); }

int rtwo_hit_sphere_f64(const double c[3], double radius, const double o[3], const double d[3], double tmin,
                        double tmax, double* t, double p[3], double n[3], int* front_face) {
    hitrec_f64 rec;
    if (!hit_sphere_f64(mk_f64(c[0], c[1], c[2]), radius, mk_f64(o[0], o[1], o[2]), mk_f64(d[0], d[1], d[2]), tmin, tmax, &rec))
        return 0;
    *t = rec.t;
    p[0] = rec.p.x; p[1] = rec.p.y; p[2] = rec.p.z;
    n[0] = rec.n.x; n[1] = rec.n.y; n[2] = rec.n.z;
    *front_face = rec.front_face;
    return 1;
}
int rtwo_hit_sphere_f32(const float c[3], float radius, const float o[3], const float d[3], float tmin, float tmax,
                        float* t, float p[3], float n[3], int* front_face) {
    hitrec_f32 rec;
    if (!hit_sphere_f32(mk_f32(c[0], c[1], c[2]), radius, mk_f32(o[0], o[1], o[2]), mk_f32(d[0], d[1], d[2]), tmin, tmax, &rec))
        return 0;
    *t = rec.t;
    p[0] = rec.p.x; p[1] = rec.p.y; p[2] = rec.p.z;
    n[0] = rec.n.x; n[1] = rec.n.y; n[2] = rec.n.z;
    *front_face = rec.front_face;
    return 1;
}
void rtwo_skycolor_f32(const float dir[3], double out[3]) { skycolor_f32(mk_f32(dir[0], dir[1], dir[2]), out); }
void rtwo_skycolor_f64(const double dir[3], double out[3]) { skycolor_f64(mk_f64(dir[0], dir[1], dir[2]), out); }

void rtwo_path_stream_f32(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t event, int n, float* out) {
    rtwo_rng g;
    memset(&g, 0, sizeof g);
    g.mode = RTWO_RNG_PHILOX;
    rtwo_rng_begin_path(&g, seed, pixel, sample);
    g.event = event;
    for (int i = 0; i < n; ++i) out[i] = trand_f32(&g);
}
void rtwo_xoroshiro_u64(uint64_t seed, int n, uint64_t* out) {
    rtwo_rng g;
    rtwo_rng_seed_xoroshiro(&g, seed);
    for (int i = 0; i < n; ++i) out[i] = xoroshiro_next(&g);
}
void rtwo_xoroshiro_f32(uint64_t seed, int n, float* out) {
    rtwo_rng g;
    rtwo_rng_seed_xoroshiro(&g, seed);
    for (int i = 0; i < n; ++i) out[i] = trand_f32(&g);
}

static cam_f32 load_cam_f32(const rtwo_camera_f32* c) {
    cam_f32 k;
    k.origin = mk_f32(c->origin[0], c->origin[1], c->origin[2]);
    k.llc = mk_f32(c->lower_left_corner[0], c->lower_left_corner[1], c->lower_left_corner[2]);
    k.horizontal = mk_f32(c->horizontal[0], c->horizontal[1], c->horizontal[2]);
    k.vertical = mk_f32(c->vertical[0], c->vertical[1], c->vertical[2]);
    k.u = mk_f32(c->u[0], c->u[1], c->u[2]);
    k.v = mk_f32(c->v[0], c->v[1], c->v[2]);
    k.w = mk_f32(c->w[0], c->w[1], c->w[2]);
    k.lens_radius = c->lens_radius;
    return k;
}
static cam_f64 load_cam_f64(const rtwo_camera_f64* c) {
    cam_f64 k;
    k.origin = mk_f64(c->origin[0], c->origin[1], c->origin[2]);
    k.llc = mk_f64(c->lower_left_corner[0], c->lower_left_corner[1], c->lower_left_corner[2]);
    k.horizontal = mk_f64(c->horizontal[0], c->horizontal[1], c->horizontal[2]);
    k.vertical = mk_f64(c->vertical[0], c->vertical[1], c->vertical[2]);
    k.u = mk_f64(c->u[0], c->u[1], c->u[2]);
    k.v = mk_f64(c->v[0], c->v[1], c->v[2]);
    k.w = mk_f64(c->w[0], c->w[1], c->w[2]);
    k.lens_radius = c->lens_radius;
    return k;
}

void rtwo_path_trace_f32(const float* geom4, const float* mat4, const uint32_t* kind, uint32_t n_spheres,
                         const rtwo_camera_f32* cam, int image_width, int max_depth, uint64_t seed, int i0, int j0, int s0,
                         double rgb[3], double* trace, int trace_cap, int* trace_n) {
    world_f32 w = {geom4, mat4, kind, n_spheres, 0, 0, trace, trace_cap, 0};
    cam_f32 c = load_cam_f32(cam);
    int H = rtwo_image_height(image_width);
    rtwo_rng g;
    memset(&g, 0, sizeof g);
    g.mode = RTWO_RNG_PHILOX;
    rtwo_rng_begin_path(&g, seed, (uint32_t)(i0 * image_width + j0), (uint32_t)s0);
    sample_path_f32(&w, &g, &c, image_width, H, max_depth, i0 + 1, j0 + 1, s0 + 1, rgb);
    if (trace_n) *trace_n = w.trace_n;
}

void rtwo_path_f32(const float* geom4, const float* mat4, const uint32_t* kind, uint32_t n_spheres,
                   const rtwo_camera_f32* cam, int image_width, int max_depth, uint64_t seed, int i0, int j0, int s0,
                   double rgb[3], uint32_t* segments) {
    world_f32 w = {geom4, mat4, kind, n_spheres, 0, 0, NULL, 0, 0};
    cam_f32 c = load_cam_f32(cam);
    int H = rtwo_image_height(image_width);
    rtwo_rng g;
    memset(&g, 0, sizeof g);
    g.mode = RTWO_RNG_PHILOX;
    rtwo_rng_begin_path(&g, seed, (uint32_t)(i0 * image_width + j0), (uint32_t)s0);
    sample_path_f32(&w, &g, &c, image_width, H, max_depth, i0 + 1, j0 + 1, s0 + 1, rgb);
    if (segments) *segments = (uint32_t)w.segments;
}

/* ------------------------------------------------------------------ threaded render drivers */

static double now_seconds(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static int pick_threads(int n_threads) {
    if (n_threads > 0) return n_threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

#define RTWO_DEFINE_RENDER(SUF, RTYPE, CAMTYPE)                                                                       \
    int rtwo_render_##SUF(const RTYPE* geom4, const RTYPE* mat4, const uint32_t* kind, uint32_t n_spheres,            \
                          const CAMTYPE* cam, int image_width, int n_samples, int max_depth, uint64_t seed,           \
                          int rng_mode, int n_threads, int row_start, int row_stride, RTYPE* out_rgb,                 \
                          double* out_linear, rtwo_stats* stats) {                                                    \
        if (!geom4 || !mat4 || !kind || !cam || !out_rgb) return -1;                                                  \
        if (image_width < 1 || n_samples < 1 || max_depth < 0) return -2;                                             \
        if (rng_mode != RTWO_RNG_PHILOX && rng_mode != RTWO_RNG_XOROSHIRO) return -3;                                 \
        if (row_stride < 1 || row_start < 0) return -4;                                                               \
        int H = rtwo_image_height(image_width);                                                                       \
        int nt = pick_threads(n_threads);                                                                             \
        if (nt > 1024) nt = 1024;                                                                                     \
        job_##SUF* jobs = (job_##SUF*)calloc((size_t)nt, sizeof(job_##SUF));                                          \
        pthread_t* th = (pthread_t*)calloc((size_t)nt, sizeof(pthread_t));                                            \
        if (!jobs || !th) { free(jobs); free(th); return -5; }                                                        \
        double t0 = now_seconds();                                                                                    \
        for (int t = 0; t < nt; ++t) {                                                                                \
            job_##SUF* jb = &jobs[t];                                                                                 \
            jb->geom4 = geom4; jb->mat4 = mat4; jb->kind = kind; jb->n = n_spheres;                                   \
            jb->cam = load_cam_##SUF(cam);                                                                            \
            jb->W = image_width; jb->H = H; jb->spp = n_samples; jb->max_depth = max_depth;                           \
            jb->seed = seed; jb->rng_mode = rng_mode;                                                                 \
            jb->row_start = row_start; jb->row_stride = row_stride;                                                   \
            jb->out_rgb = out_rgb; jb->out_linear = out_linear;                                                       \
            jb->tid = t; jb->nthreads = nt;                                                                           \
            jb->joinable = (nt > 1 && pthread_create(&th[t], NULL, worker_##SUF, jb) == 0);                          \
            if (!jb->joinable) worker_##SUF(jb); /* single thread, or thread creation failed: run inline */          \
        }                                                                                                             \
        uint64_t seg = 0, tests = 0;                                                                                  \
        for (int t = 0; t < nt; ++t) {                                                                                \
            if (jobs[t].joinable) pthread_join(th[t], NULL);                                                          \
            seg += jobs[t].segments; tests += jobs[t].tests;                                                          \
        }                                                                                                             \
        double t1 = now_seconds();                                                                                    \
        if (stats) {                                                                                                  \
            int nrows = (H - row_start + row_stride - 1) / row_stride;                                                \
            if (nrows < 0) nrows = 0;                                                                                 \
            stats->paths = (uint64_t)nrows * (uint64_t)image_width * (uint64_t)n_samples;                             \
            stats->ray_segments = seg; stats->sphere_tests = tests;                                                   \
            stats->seconds = t1 - t0; stats->threads = nt;                                                         \
        }                                                                                                             \
        free(jobs); free(th);                                                                                         \
        return 0;                                                                                                     \
    }

RTWO_DEFINE_RENDER(f32, float, rtwo_camera_f32)
RTWO_DEFINE_RENDER(f64, double, rtwo_camera_f64)
