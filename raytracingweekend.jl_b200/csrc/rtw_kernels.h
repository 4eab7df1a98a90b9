// rtw_kernels.h -- host-visible launch interface of the sm_100a kernels (internal; the public ABI is include/rtw_b200.h)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rtw_device.cuh"

namespace rtw {

// HBM layout of one render on one device:
//   geom   : n x float4 {cx,cy,cz,r}        (16 B aligned; staged to shared memory by 1-D bulk TMA, 1024-sphere tiles)
//   mat    : n x float4 {albedo rgb, fuzz|ir}
//   kind   : n x u32
//   accum  : n_rows*W x 4 x i64  fixed-point radiance sums (r,g,b,pad), scale 2^fx_bits, order-independent atomics
//   counters[0] = path ticket counter, counters[1] = ray segments traced
// unsigned division by an invariant through multiply + shifts (make_magic_div); exact for every 32-bit dividend
struct MagicDiv {
    uint32_t m, sh1, sh2;
};

// RTW_MODE_GRID: a uniform grid over the small spheres + a list of big ones (rtw_grid.cuh); nx == 0: no grid.
// The cells carry TWO registrations of the small spheres (CSR lists of the spheres whose AABB, inflated by a margin,
// touches the cell): a tight one (5 % of a cell) for the near part of a flight and a loose one (up to one cell) for
// the far part, where the reference's arithmetic lets spheres grow (rtw_grid.cuh).  On small scenes both are the same.
struct GridCull {  // by-centre binning into coarse cells: lets the exact sweep of an unsafe ray skip what it cannot reach
    float ox, oy, oz, h;
    int nx, ny, nz;              // nx == 0: no culling, the sweep visits the whole list
    float r_max;                 // largest small radius
    const uint32_t* cell_start;  // nx*ny*nz + 1 offsets into items
    const uint32_t* items;       // every small sphere exactly once, ascending within a cell
};
struct GridParams {
    float ox, oy, oz;  // minimum corner: the extent of the small spheres padded by `pad`
    float h, inv_h;    // cell edge
    int nx, ny, nz;
    float safe2_tight, safe2_loose;  // ((r_min + margin)^2 - r_min^2) / 2: the growth of r^2 each registration covers
    float reach;       // largest small radius + loose margin: how far from its centre a registered sphere can matter
    float pad;         // how far the box extends beyond the real spheres (loose margin + half a cell)
    float r_min;       // smallest small radius (the sphere whose apparent size grows fastest)
    float ball_r;      // half the diagonal of the box
    const uint32_t* cell_start_tight;  // nx*ny*nz + 1 offsets into items_tight
    const uint32_t* items_tight;       // sphere indices per cell, ascending
    const uint32_t* cell_start_loose;
    const uint32_t* items_loose;
    const uint32_t* big;  // spheres every ray tests
    uint32_t n_big;
    GridCull cull;
};

struct TraceParams {
    DevCamera cam;
    const float4* geom;
    const float4* geom_pairs;  // pair layout for the packed sweep: {xa,xb,ya,yb}{za,zb,ra,rb} per 2 spheres
    // RTW_TAIL_UNIFIED only: AoS copy of the list in the order the lanes of a cooperating group meet the spheres
    // (entry (c*coop + h)*32 + j = list index c*32*coop + 2*((j>>1)*coop + h) + (j&1)), zero padded to whole
    // super-chunks of 32*coop spheres; matches the `coop` the kernel is launched with
    const float4* geom_perm;
    const float* u_tab;  // W entries: T((col+1)/W), src/render.jl:26
    const float* v_tab;  // H entries: T((H-1-i0)/H), src/render.jl:27
    MagicDiv div_spp, div_w;
    uint32_t rk[20];  // Philox round keys: rk[2r] = key0 + r*W0, rk[2r+1] = key1 + r*W1 (the first 2*kPhiloxRounds are used)
    GridParams grid;  // RTW_MODE_GRID only
    const float4* mat;
    const uint32_t* kind;
    uint32_t n_spheres;
    int W, H, spp, max_depth;  // spp = samples per pixel traced by THIS launch
    int sample_first;          // they are samples sample_first .. sample_first + spp - 1 of the image (progressive passes)
    uint32_t key0, key1;  // Philox key = seed lo/hi
    int row_start, row_stride, n_rows;
    unsigned long long n_paths;  // n_rows * W * spp ; path ticket t -> pixel t / spp, sample t % spp
    unsigned long long* accum;
    double fx_scale;  // 2^fx_bits
    unsigned long long* counters;
};

// Float64 instantiation (rtw_f64.cu): Camera{Float64} fields get_ray reads, scene as double4 arrays
struct DevCamera64 {
    double origin[3], llc[3], horizontal[3], vertical[3], u[3], v[3];
    double lens_radius;
};

struct TraceParams64 {
    DevCamera64 cam;
    const double4* geom;  // n x {cx,cy,cz,r}
    const double4* mat;   // n x {albedo rgb, fuzz|ir}
    const uint32_t* kind;
    uint32_t n_spheres;
    int W, H, spp, max_depth, sample_first;
    uint32_t key0, key1;
    int row_start, row_stride, n_rows;
    unsigned long long n_paths;
    unsigned long long* accum;
    double fx_scale;
    unsigned long long* counters;
};

// RTW_MODE_WAVEFRONT: the path pool in HBM (structure of arrays over `capacity` slots) and its work lists
struct WavefrontBuffers {
    uint32_t capacity;
    float4* ray_o;        // origin (xyz)
    float4* ray_d;        // unit direction (xyz)
    double* thr;          // 3 x capacity: product of attenuations so far (Float64)
    uint32_t* pix_local;  // accumulator pixel of the path
    uint32_t* sample;     // Philox counter word 1
    uint32_t* pixel;      // Philox counter word 2 (global pixel index)
    int* depth_left;
    float* hit_t;
    int* hit_k;           // closest sphere; -1 miss (sky), -2 depth exhausted (black), -3 slot never used
    uint32_t* alive;
    uint32_t* list[3];    // 0: path ended (sky / depth) -> accumulate + regenerate; 1: Lambertian/Metal; 2: Dielectric
    unsigned int* counts; // [0..2] list lengths, [3] rays traced in the current step
};

struct LaunchInfo {
    int grid, block, smem_bytes, blocks_per_sm, launches, rays_per_lane, sweep;
};

constexpr int kSweepBranch = 1;  // RTW_SWEEP_BRANCH
constexpr int kSweepMask = 2;    // RTW_SWEEP_MASK
constexpr int kSweepPacked = 3;  // RTW_SWEEP_PACKED

// spheres per shared-memory tile of the sweep (32 chunks of 32): 16 KB of geometry per buffer
constexpr uint32_t kTileSpheres = 1024;

cudaError_t launch_fused_trace(const TraceParams& p, int num_sms, int blocks_per_sm_override, int rays_per_lane,
                               int sweep, int coop, cudaStream_t stream, LaunchInfo* info);
// RTW_TAIL_UNIFIED (rtw_fused2.cu): packed sweep, one path per lane, coop = 2 or 4; needs geom_perm / u_tab / v_tab /
// div_* of TraceParams
// walk: 1 = per-slot walks + merge, 2 = every lane resolves the candidates of its own ray (transposed), 0 = default
cudaError_t launch_fused_trace2(const TraceParams& p, int num_sms, int blocks_per_sm_override, int coop, int walk,
                                cudaStream_t stream, LaunchInfo* info);
// RTW_MODE_GRID: the unified-tail kernel with the grid traversal in place of the sweep (needs TraceParams.grid)
cudaError_t launch_fused_trace2_grid(const TraceParams& p, int num_sms, int blocks_per_sm_override, cudaStream_t stream,
                                     LaunchInfo* info);
cudaError_t launch_uv_tables(int W, int H, float* u_tab, float* v_tab, cudaStream_t stream);
MagicDiv make_magic_div(uint32_t d);
// RTW_MODE_WAVEFRONT (rtw_wavefront.cu); synchronises `stream` internally (host-driven step loop)
size_t wavefront_bytes(uint32_t capacity);
cudaError_t launch_wavefront_trace(const TraceParams& p, const WavefrontBuffers& b, int num_sms, unsigned int* h_traced,
                                   cudaStream_t stream, LaunchInfo* info);
// RTW_MODE_CTA_WAVEFRONT (rtw_cta_wavefront.cu): persistent CTAs, path pool and work lists in shared memory
cudaError_t launch_cta_wavefront_trace(const TraceParams& p, int num_sms, int blocks_per_sm_override,
                                       cudaStream_t stream, LaunchInfo* info);
cudaError_t launch_resolve(const unsigned long long* accum, int W, int H, int n_rows, int row_start, int row_stride,
                           int spp, double inv_scale, int column_major, float* out, cudaStream_t stream);
cudaError_t launch_assemble(const float* tiles, int n_tiles, int W, int H, float* out, cudaStream_t stream);
// Julia column-major Float32 image -> row-major 8-bit RGB, clamp01nan + N0f8 rounding (rtw_image.cu)
cudaError_t launch_quantize_rgb8(const float* img, int W, int H, unsigned char* out, cudaStream_t stream);
// latency path (rtw_small.cu): a whole small render in one launch; both device counters must be zero on entry
cudaError_t launch_small_render(const TraceParams& p, double inv_scale, float* out_img, unsigned long long* host_totals,
                                cudaStream_t stream, LaunchInfo* info);
// device-side scene_random_spheres (rtw_scenegen.cu)
size_t scenegen_max_spheres(int half);
size_t scenegen_workspace_bytes(int half);
cudaError_t launch_scenegen(unsigned long long s0, unsigned long long s1, int half, void* workspace, float4* geom, float4* mat,
                            uint32_t* kind, unsigned long long* out_host, cudaStream_t stream);
// Float64 path (rtw_f64.cu)
cudaError_t launch_trace_f64(const TraceParams64& p, int num_sms, cudaStream_t stream, LaunchInfo* info);
cudaError_t launch_small_render_f64(const TraceParams64& p, double inv_scale, double* out_img, unsigned long long* host_totals,
                                    cudaStream_t stream, LaunchInfo* info);
cudaError_t launch_resolve_f64(const unsigned long long* accum, int W, int H, int n_rows, int row_start, int row_stride,
                               int spp, double inv_scale, int column_major, double* out, cudaStream_t stream);
cudaError_t launch_assemble_f64(const double* tiles, int n_tiles, int W, int H, double* out, cudaStream_t stream);
// FP32 issue microbenchmarks; returns lane-instructions executed through *fp32_instr
cudaError_t launch_fp32_peak(int variant, int num_sms, float* scratch, cudaStream_t stream, double* fp32_instr);

}  // namespace rtw
