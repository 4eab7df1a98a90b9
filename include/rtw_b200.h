/*
 * rtw_b200.h -- C-ABI of librtw_b200.so: the B200 (sm_100a) implementation of
 * RayTracingWeekend.jl's render() -> ray_color() -> hit()/scatter() hot path.
 *
 * The reference (claforte/RayTracingWeekend.jl @ fe20135d) has no FFI of its own; the boundary this
 * library sits behind is the Julia call
 *     render(scene::HittableList, cam::Camera{T}, image_width=400, n_samples=1)   src/render.jl:8-9
 * Julia host code (Camera/default_camera src/camera.jl:1-41, scene builders src/scenes.jl, structs
 * src/structs.jl) stays as it is; a thin shim (raytracingweekend.jl_b200/julia/RayTracingWeekendB200.jl,
 * shown in INTEGRATION.md) flattens the scene and `ccall`s the entry points below.  Everything is plain C:
 * pointers, sizes, POD structs.  No C++ exception crosses this boundary; nothing here aborts the process.
 *
 * There is NO CPU fallback: every compute entry point needs a CUDA device and fails with an error code
 * otherwise.
 *
 * Every function returns an int status: 0 = RTW_OK, < 0 = RTW_E_* below, > 0 = a cudaError_t value.
 * rtw_last_error(ctx) returns a human-readable message for the last failure on that context.
 */
#ifndef RTW_B200_H
#define RTW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* v3: rtw_stats grew (n_devices, grid counters), rtw_scene_random_spheres, rtw_has_variants, RTW_OPT_GATHER,
 *     RTW_OPT_SMALL_RENDER; RTW_OPT_STRIP removed; kernel variants need RTW_BUILD_VARIANTS=1 */
#define RTW_ABI_VERSION 3

#if defined(__GNUC__)
#define RTW_API __attribute__((visibility("default")))
#else
#define RTW_API
#endif

#define RTW_OK 0
#define RTW_E_INVALID_ARG (-1)   /* null pointer, non-positive size, bad device index ...            */
#define RTW_E_NO_DEVICE (-2)     /* no CUDA device / driver: the hot path has no CPU fallback        */
#define RTW_E_NO_SCENE (-3)      /* render called before rtw_set_scene                               */
#define RTW_E_UNSUPPORTED (-4)   /* e.g. unknown material kind, unknown option                       */
#define RTW_E_INTERNAL (-5)
#define RTW_E_IO (-6)            /* a file could not be opened / read / written                       */
#define RTW_E_FORMAT (-7)        /* not a .rtwscene file, or its checksum does not match              */

/* Material{T} subtypes, flattened (src/material.jl:3-5 Lambertian, :25-29 Metal, :37-39 Dielectric) */
#define RTW_LAMBERTIAN 0u
#define RTW_METAL 1u
#define RTW_DIELECTRIC 2u

/* ray_color's default `depth` (src/ray_color.jl:14) -- render() never overrides it (src/render.jl:38) */
#define RTW_DEFAULT_MAX_DEPTH 16
/* reseed!() at the top of every render (src/render.jl:21, src/rand.jl:2): same seed => same image (the stream is an
 * addressed Philox4x32-7 keyed by the seed) */
#define RTW_DEFAULT_SEED 1ull

/*
 * Camera{Float32}: identical field order and layout to the isbits Julia struct, src/camera.jl:1-10
 * (7 x Vec3{Float32} + lens_radius = 22 floats = 88 bytes), so Julia passes it with Ref(cam).
 */
typedef struct rtw_camera {
    float origin[3];
    float lower_left_corner[3];
    float horizontal[3];
    float vertical[3];
    float u[3];
    float v[3];
    float w[3];
    float lens_radius;
} rtw_camera;

/* Work counters and device timings of the last render (what Mrays/s and the roofline are computed from). */
typedef struct rtw_stats {
    uint64_t paths;          /* (pixel, sample) pairs traced = rows * W * n_samples                      */
    uint64_t ray_segments;   /* executions of hit(world, r, ...) (src/ray_color.jl:19): primary + bounces */
    uint64_t sphere_tests;   /* ray_segments * n_spheres = executions of hit(::Sphere) (src/hit.jl:12)   */
    uint32_t n_spheres;
    int32_t image_width;
    int32_t image_height;
    int32_t rows_rendered;
    int32_t kernel_launches; /* kernels this library launched for the call                                */
    float ms_total;          /* CUDA-event time of the whole call on the device (incl. copies if any)     */
    float ms_trace;          /* the trace kernel(s): raygen + intersect + shade + accumulate              */
    float ms_resolve;        /* accumulator -> gamma-2 RGB (src/render.jl:40, src/vec.jl:22)              */
    float ms_h2d;            /* scene / camera upload inside the call                                     */
    float ms_d2h;            /* image download inside the call                                            */
    /* ABI v3 */
    int32_t n_devices;       /* devices the call ran on (rows r -> device r mod n_devices)                 */
    int32_t reserved0;
    uint64_t grid_fallback_rays; /* RTW_MODE_GRID: ray segments no registration margin covers (non-unit direction after
                                    a glass reflection, flying far), resolved by the exact whole-list sweep     */
    uint64_t grid_loose_cells;   /* RTW_MODE_GRID: cells walked with the loose registration (far part of long flights) */
    uint64_t grid_cells;         /* RTW_MODE_GRID work model: cells walked by all rays ...                        */
    uint64_t grid_tests;         /* ... and ray-sphere tests made in them and on the big spheres (the exact sweeps of
                                    grid_fallback_rays are not included)                                          */
} rtw_stats;

typedef struct rtw_ctx rtw_ctx;

/* ---- options for rtw_set_option -------------------------------------------------------------- */
#define RTW_OPT_MODE 1            /* RTW_MODE_*                                                        */
#define RTW_OPT_STRIP 2           /* removed in ABI v3: RTW_E_UNSUPPORTED                              */
#define RTW_OPT_BLOCKS_PER_SM 3   /* persistent CTAs per SM (0 = library default)                      */
#define RTW_OPT_COLLECT_TIMING 4  /* 1 = record per-stage CUDA-event timings into rtw_stats (default 1) */
#define RTW_OPT_RAYS_PER_LANE 5   /* paths traced concurrently by one lane: 1, 2 or 4 (0 = library default) */
#define RTW_OPT_SWEEP 6           /* RTW_SWEEP_*: inner-loop variant of the sphere-list sweep          */
#define RTW_OPT_COOP 7            /* lanes sharing sphere loads in the packed sweep: 1, 2 or 4 (0 = default) */
#define RTW_OPT_TAIL 8            /* RTW_TAIL_*: layout of the per-bounce work after the sweep (RTW_MODE_FUSED) */

#define RTW_OPT_WALK 9            /* RTW_WALK_*: candidate resolution after the sweep (RTW_TAIL_UNIFIED, lists <= 1024) */

#define RTW_OPT_GATHER 10         /* RTW_GATHER_*: how a multi-device context collects the row tiles on device 0 */

#define RTW_OPT_SMALL_RENDER 11    /* 1 (default): a small render() on a one-device context (<= 2^17 paths, <= 1024 spheres,
                                     default kernel options) runs as ONE kernel launch with the image written to mapped
                                     host memory -- the latency path for the reference's own 96x54 smoke/benchmark sizes;
                                     0: always the persistent kernel.  Same image bits either way. */

#define RTW_GATHER_PEER 0         /* cudaMemcpyPeerAsync on each producer's stream (default; measured fastest)  */
#define RTW_GATHER_NCCL 1         /* one grouped ncclSend/ncclRecv (single-process ncclCommInitAll); libnccl.so.2 is
                                     bound with dlopen on first use -- RTW_E_UNSUPPORTED when it is not installed   */

#define RTW_WALK_DEFAULT 0        /* library default (the fastest measured)                           */
#define RTW_WALK_SLOTS 1          /* a lane resolves, slot by slot, the candidates it found; partial hits merged by shuffle */
#define RTW_WALK_OWN_RAY 2        /* a lane resolves all candidates of its own ray, reading its partners' masks */

#define RTW_TAIL_DEFAULT 0        /* library default (the fastest measured)                           */
#define RTW_TAIL_SPLIT 1          /* regenerate / shade each run by the lanes in that state            */
#define RTW_TAIL_UNIFIED 2        /* one Philox block, cooperative rejection sampling and shared normalize for
                                     continuing and new paths (packed sweep, 1 path per lane, coop 2 or 4) */

#define RTW_SWEEP_DEFAULT 0       /* library default (the fastest measured)                           */
#define RTW_SWEEP_BRANCH 1        /* test + immediate root selection under a branch                    */
#define RTW_SWEEP_MASK 2          /* sign-bit candidate masks, roots resolved after the sweep          */
#define RTW_SWEEP_PACKED 3        /* RTW_SWEEP_MASK on packed FP32x2 instructions (two spheres / instr) */

#define RTW_MODE_FUSED 0          /* one persistent kernel: raygen -> {intersect, shade} loop -> accumulate */
#define RTW_MODE_WAVEFRONT 1      /* separate raygen / intersect / shade / accumulate kernels + compaction  */
#define RTW_MODE_GRID 3           /* the fused kernel with a uniform-grid traversal in place of the linear sweep: the same
                                     closest hit and image bits from far fewer sphere tests, for EVERY list size (the
                                     reference is brute force by design, README.md:30; acceleration structures are its
                                     long-term goal, README.md:175).  Exact by construction: rays the grid cannot answer
                                     with certainty -- the reference's a = 1 shortcut and the rounding of its discriminant
                                     let far spheres grow -- are resolved by an exact cooperative sweep and counted in
                                     rtw_stats.grid_fallback_rays.  Not the benchmarked path;
                                     rtw_stats.sphere_tests still reports ray_segments * n_spheres, the tests the linear
                                     sweep would have made */
#define RTW_MODE_CTA_WAVEFRONT 2  /* the same stages inside persistent CTAs: path pool + work lists in shared memory
                                     (lists <= 1024 spheres); a measured comparison, only in libraries built with
                                     RTW_BUILD_VARIANTS=1 (rtw_has_variants)                                    */

/* ---- life cycle ------------------------------------------------------------------------------ */

RTW_API int rtw_abi_version(void);

/* 1 when the library was built with RTW_BUILD_VARIANTS=1 (csrc/build.sh): the kernel families kept only as measured
 * comparisons -- RTW_TAIL_SPLIT, RTW_SWEEP_BRANCH/MASK, rays per lane 2/4, RTW_OPT_COOP 1/4, RTW_WALK_SLOTS,
 * RTW_MODE_CTA_WAVEFRONT -- are then selectable; the default build returns RTW_E_UNSUPPORTED for them. */
RTW_API int rtw_has_variants(void);

/* number of visible CUDA devices (0 and RTW_E_NO_DEVICE when there is none) */
RTW_API int rtw_device_count(int* count);

/* image height render() derives from the width: image_width div (16//9), src/render.jl:11-12 */
RTW_API int rtw_image_height(int image_width);

/*
 * Create a context that renders on the given CUDA devices (device_ids == NULL => devices 0..n_devices-1;
 * n_devices <= 0 => device 0 only).  Owns streams, device buffers and events; buffers grow lazily.
 * Replaces the module-load set-up of the reference (src/init.jl:4-12).  One context is not re-entrant
 * (calls are serialised by an internal mutex); distinct contexts are independent.
 */
RTW_API int rtw_create(const int* device_ids, int n_devices, rtw_ctx** out_ctx);
RTW_API int rtw_destroy(rtw_ctx* ctx);
RTW_API const char* rtw_last_error(const rtw_ctx* ctx);
RTW_API int rtw_set_option(rtw_ctx* ctx, int option, int64_t value);

/* ---- scene ----------------------------------------------------------------------------------- */

/*
 * Upload a flattened HittableList (src/structs.jl:10, Sphere src/structs.jl:31-35) to every device of ctx.
 *   geom4 : n x {center.x, center.y, center.z, radius}   (radius keeps its sign: hollow glass, src/scenes.jl:35-36)
 *   mat4  : n x {albedo.r, albedo.g, albedo.b, param}    param = fuzz (Metal) | ir (Dielectric) | 0 (Lambertian)
 *   kind  : n x RTW_LAMBERTIAN | RTW_METAL | RTW_DIELECTRIC
 * List order is preserved (ties in t go to the later sphere, src/hit.jl:24-26,44-46).  Host pointers;
 * the library copies and never retains them.  Failure-atomic: after an error the context holds no scene.  Passing the
 * arrays of the scene that is already resident is a no-op (rtw_render_scene does it on every call).
 * Albedo components must be finite and >= 0 (RTW_E_UNSUPPORTED otherwise).  The per-pixel sums are 64-bit fixed point
 * with 6 bits (64x) of head-room per path: every albedo <= 1 -- all of the reference's scenes -- can never exceed it; a
 * scene that uses albedo > 1 as emission is rendered as long as max_albedo^(max_depth - 1) <= 64 and the render call is
 * refused (RTW_E_UNSUPPORTED) beyond that, instead of saturating silently.
 */
RTW_API int rtw_set_scene(rtw_ctx* ctx, const float* geom4, const float* mat4, const uint32_t* kind, uint32_t n_spheres);

/*
 * scene_random_spheres(; elem_type = Float32), src/scenes.jl:49-84, built ON THE DEVICE -- no host loop, so the ~100k-sphere
 * list of the large configuration takes milliseconds instead of seconds.  The list is the one the reference's sequential
 * loop produces, bit for bit: `rng_state` is the state (s0, s1) of the calling thread's Xoroshiro128Plus (src/rand.jl:7),
 * from which the builder draws in whatever state it is; on return it holds the state after the builder's last draw, so
 * host code that keeps using the same generator continues as after the reference's loop.  half_extent generalises the
 * `-11:10` grid (reference value 11; 158 gives BASELINE's ~100k spheres).
 *   install != 0 : the list becomes the scene of the context (as rtw_set_scene);
 *   geom4 / mat4 / kind (may all be NULL) : host arrays of `capacity` >= n_spheres entries that receive the list.
 * *n_spheres is always written (4 * half_extent^2 + 4 is an upper bound).
 */
RTW_API int rtw_scene_random_spheres(rtw_ctx* ctx, uint64_t rng_state[2], int half_extent, int install, float* geom4,
                                     float* mat4, uint32_t* kind, uint32_t capacity, uint32_t* n_spheres);

/* ---- the hot path ---------------------------------------------------------------------------- */

/*
 * render(scene, cam, image_width, n_samples), src/render.jl:8-44, on the scene last given to rtw_set_scene.
 *   out_rgb : HOST buffer of H*W*3 floats in the memory layout of Julia's Matrix{RGB{Float32}}(H, W)
 *             (column-major): pixel (row i0, col j0), 0-based, at ((j0*H)+i0)*3.  Row 0 is the top row.
 *             Values are post-gamma (sqrt, src/vec.jl:22) and unclamped, as in the reference.
 *   max_depth : ray_color depth (reference: 16).   seed : stream seed (reference semantics: constant).
 * Synchronous: returns after the image is in out_rgb.  Rows are split over all devices of the context
 * (row r -> device r mod n_devices), tiles are collected on device 0.
 */
RTW_API int rtw_render(rtw_ctx* ctx, const rtw_camera* cam, int image_width, int n_samples, int max_depth, uint64_t seed,
               float* out_rgb, rtw_stats* stats);

/* rtw_set_scene + rtw_render in one call: exactly what the Julia method render(scene, cam, W, spp) binds to. */
RTW_API int rtw_render_scene(rtw_ctx* ctx, const float* geom4, const float* mat4, const uint32_t* kind, uint32_t n_spheres,
                     const rtw_camera* cam, int image_width, int n_samples, int max_depth, uint64_t seed,
                     float* out_rgb, rtw_stats* stats);

/*
 * Device-resident variant (inputs already in HBM, output stays in HBM; used by the multi-process
 * driver, where each rank owns one GPU, and for kernel-only timing).  Renders rows
 *   i0 = row_start, row_start + row_stride, ...  (< H)
 * of the image on device `device_slot` (index into the ctx's device list) and writes a compact tile
 *   d_tile[(k*W + j0)*3 + c],  k = 0 .. n_rows-1   (row-major, post-gamma Float32)
 * into DEVICE memory.  `stream` is a cudaStream_t (NULL = the context's own stream; to target the legacy
 * default stream pass cudaStreamLegacy, i.e. (void*)1); the call only
 * enqueues work on it and returns -- synchronise the stream before reading d_tile or stats.
 * With row_start = 0, row_stride = 1 and column_major != 0 the tile is written in the Julia layout of
 * rtw_render instead.
 */
RTW_API int rtw_render_rows_device(rtw_ctx* ctx, int device_slot, const rtw_camera* cam, int image_width, int n_samples,
                           int max_depth, uint64_t seed, int row_start, int row_stride, int column_major,
                           float* d_tile, void* stream);

/* Work counters / timings of the last rtw_render_rows_device on that device (synchronises its stream). */
RTW_API int rtw_last_stats(rtw_ctx* ctx, int device_slot, rtw_stats* stats);

/*
 * Un-interleave `n_tiles` gathered row tiles (tile g holds rows g, g+n_tiles, ...; tiles are stored
 * back to back, each padded to ceil(H/n_tiles) rows) into the Julia column-major image.  Device
 * pointers on device `device_slot`; enqueued on `stream`.
 */
RTW_API int rtw_assemble_tiles_device(rtw_ctx* ctx, int device_slot, const float* d_tiles, int n_tiles, int image_width,
                              float* d_out_rgb, void* stream);

/* ---- Float64 (the reference is generic over T; its own test and published timings use Float64) ----------- */

/* Camera{Float64}: same field order as rtw_camera, 22 doubles (src/camera.jl:1-10) */
typedef struct rtw_camera_f64 {
    double origin[3];
    double lower_left_corner[3];
    double horizontal[3];
    double vertical[3];
    double u[3];
    double v[3];
    double w[3];
    double lens_radius;
} rtw_camera_f64;

/*
 * The Float64 instantiation of rtw_set_scene / rtw_render / rtw_render_scene: scene arrays and image are doubles
 * (out_rgb = memory of Matrix{RGB{Float64}}(H, W)), the path is traced in Float64 on the FP64 pipe with the same
 * addressed stream (a Float64 draw takes two words: 52 random mantissa bits).  The Float64 scene is held separately
 * from the Float32 one.
 */
RTW_API int rtw_set_scene_f64(rtw_ctx* ctx, const double* geom4, const double* mat4, const uint32_t* kind, uint32_t n_spheres);
RTW_API int rtw_render_f64(rtw_ctx* ctx, const rtw_camera_f64* cam, int image_width, int n_samples, int max_depth,
                           uint64_t seed, double* out_rgb, rtw_stats* stats);
RTW_API int rtw_render_scene_f64(rtw_ctx* ctx, const double* geom4, const double* mat4, const uint32_t* kind,
                                 uint32_t n_spheres, const rtw_camera_f64* cam, int image_width, int n_samples,
                                 int max_depth, uint64_t seed, double* out_rgb, rtw_stats* stats);

/* ---- progressive rendering ------------------------------------------------------------------- */

/*
 * render() split into passes over the samples.  Every random draw is addressed by (pixel, sample, event, draw) and
 * the accumulator is an integer sum, so the image after rtw_accumulate(0, a) + rtw_accumulate(a, b) is bit-identical
 * to rtw_render with a + b samples (src/render.jl:29-40 run in one go).  Uses: previews of a 1000-spp render,
 * checkpoint / resume (rtw_accumulator_read / _write), time-sliced rendering.
 *
 * rtw_accumulate adds samples sample_first .. sample_first + sample_count - 1 of every pixel to the accumulators held
 * by the context.  n_samples_total = the number of samples the finished image will hold: it fixes the fixed-point
 * scale, so it must be the same in every call of one image.  sample_first == 0 starts a new image (accumulators
 * zeroed); otherwise it must equal the number of samples accumulated so far, with the same image_width.
 * rtw_render* discards a progressive image.
 */
RTW_API int rtw_accumulate(rtw_ctx* ctx, const rtw_camera* cam, int image_width, int sample_first, int sample_count,
                           int n_samples_total, int max_depth, uint64_t seed, rtw_stats* stats);

/* sqrt(accumulated sum / samples accumulated so far) into a HOST buffer, layout as rtw_render (src/render.jl:40). */
RTW_API int rtw_resolve(rtw_ctx* ctx, float* out_rgb);

/*
 * The same image as 8-bit RGB, row-major, top row first (H*W*3 bytes, HOST buffer): what Images.jl writes for
 * save("x.png", img) -- every channel through clamp01nan, then N0f8: round(x * 255).  (The reference itself never
 * saves an image: README.md:138,170.)
 */
RTW_API int rtw_resolve_rgb8(rtw_ctx* ctx, uint8_t* out_rgb8);

/* state of the progressive image: width, samples accumulated, samples planned (all 0 when there is none) */
RTW_API int rtw_progress(rtw_ctx* ctx, int* image_width, int* samples_done, int* samples_total);

/*
 * Checkpoint / resume: the raw accumulators of the progressive image, H*W*4 int64 values, row-major
 * [row][col][r,g,b,unused] fixed-point sums (HOST buffers; n_values must be H*W*4).  rtw_accumulator_write installs
 * a saved state (in a context with the same or a different device count) so that rtw_accumulate can continue at
 * sample `samples_done`.
 */
RTW_API int rtw_accumulator_read(rtw_ctx* ctx, int64_t* out, uint64_t n_values);
RTW_API int rtw_accumulator_write(rtw_ctx* ctx, const int64_t* in, uint64_t n_values, int image_width, int samples_done,
                                  int samples_total);

/*
 * The same checkpoint as a self-describing file ("RTWCKPT2": width, samples done / planned, fixed-point scale, and the
 * seed, max_depth, camera and scene the samples were traced with, CRC-32).  rtw_checkpoint_load needs the scene to be
 * set already and refuses a file that belongs to another scene (RTW_E_INVALID_ARG) or is damaged (RTW_E_FORMAT); after
 * it, rtw_accumulate continues at the saved sample and -- unlike after rtw_accumulator_write, whose raw sums carry no
 * metadata -- must be given the seed, max_depth and camera of the file.
 */
RTW_API int rtw_checkpoint_save(rtw_ctx* ctx, const char* path);
RTW_API int rtw_checkpoint_load(rtw_ctx* ctx, const char* path);

/* ---- image and scene files (host side; no device needed) ---------------------------------------- */

/* binary PPM (P6) / PNG (8-bit RGB, stored deflate blocks) of a row-major 8-bit image, top row first */
RTW_API int rtw_write_ppm(const char* path, const uint8_t* rgb8, int width, int height);
RTW_API int rtw_write_png(const char* path, const uint8_t* rgb8, int width, int height);

/*
 * .rtwscene: the flattened HittableList exactly as it crosses rtw_set_scene -- "RTWSCN01", u32 n, u32 0 (Float32),
 * geom4[n][4], mat4[n][4], kind[n], CRC-32 -- little endian.  One fixture format for the Julia shim, the Python
 * harness, the oracle and the kernels.  rtw_scene_load with capacity 0 only reports n_spheres.
 */
RTW_API int rtw_scene_save(const char* path, const float* geom4, const float* mat4, const uint32_t* kind, uint32_t n_spheres);
RTW_API int rtw_scene_load(const char* path, float* geom4, float* mat4, uint32_t* kind, uint32_t capacity,
                           uint32_t* n_spheres);

/* ---- roofline denominator --------------------------------------------------------------------- */

/*
 * FP32 issue microbenchmark on device `device_slot`: independent FFMA chains on every SM.
 *   variant 0: pure FFMA;  variant 1: the scalar mask sweep's own mix (3 FADD, 2 FMUL, 6 FFMA + 1 SHF, 1 LDS.128 per test);
 *   variant 2: the packed FP32x2 mask sweep (the same 11 lane-ops per test, two tests per instruction);
 *   variants 3, 4: variant 2 with 2 / 4 cooperating lanes per sphere load (RTW_OPT_COOP)
 *   variant 9: independent Float64 FMA chains (DFMA) -- the measured denominator of the Float64 kernel's fraction
 * Writes achieved FP32 instructions/s (lane-instructions, i.e. warp instructions x 32) and the kernel time.
 */
RTW_API int rtw_measure_fp32_peak(rtw_ctx* ctx, int device_slot, int variant, double* fp32_instr_per_s, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* RTW_B200_H */
