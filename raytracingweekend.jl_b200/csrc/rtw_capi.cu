// rtw_capi.cu -- the C-ABI of include/rtw_b200.h: context, buffers, streams, multi-device row split.
// No C++ exception leaves this file; every entry point returns a status code.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types and enums only: the library is bound with dlopen when RTW_GATHER_NCCL is selected

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/rtw_b200.h"
#include "rtw_kernels.h"

namespace {

struct DeviceState {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_tile = nullptr;  // tile ready (multi-device gather)
    // scene
    float4* d_geom = nullptr;
    float4* d_geom_pairs = nullptr;
    float4* d_geom_perm[2] = {nullptr, nullptr};  // RTW_TAIL_UNIFIED: lane-order AoS copies for coop = 2, 4 (lists <= 1024)
    float* d_uv = nullptr;                         // u_tab (W) then v_tab (H)
    size_t uv_cap = 0;
    float4* d_mat = nullptr;
    uint32_t* d_kind = nullptr;
    size_t scene_cap = 0;
    // RTW_MODE_GRID: uniform grid over the scene (same content on every device)
    uint32_t* d_grid_start[2] = {nullptr, nullptr};  // [0] tight, [1] loose registration
    uint32_t* d_grid_items[2] = {nullptr, nullptr};
    size_t grid_start_cap[2] = {0, 0}, grid_items_cap[2] = {0, 0};
    uint32_t* d_grid_big = nullptr;
    size_t grid_big_cap = 0;
    uint32_t* d_cull_start = nullptr;
    uint32_t* d_cull_items = nullptr;
    size_t cull_start_cap = 0, cull_items_cap = 0;
    // Float64 scene and image buffers (rtw_*_f64)
    double4* d_geom64 = nullptr;
    double4* d_mat64 = nullptr;
    uint32_t* d_kind64 = nullptr;
    size_t scene64_cap = 0;
    double* d_tile64 = nullptr;
    size_t tile64_cap = 0;
    double* d_gather64 = nullptr;
    size_t gather64_cap = 0;
    double* d_image64 = nullptr;
    size_t image64_cap = 0;
    // render buffers
    unsigned long long* d_accum = nullptr;
    size_t accum_cap = 0;  // in pixels
    unsigned long long* d_counters = nullptr;
    unsigned long long* h_counters = nullptr;  // pinned
    float* d_tile = nullptr;
    size_t tile_cap = 0;  // floats
    float* d_gather = nullptr;
    size_t gather_cap = 0;
    float* d_image = nullptr;
    size_t image_cap = 0;
    unsigned char* d_rgb8 = nullptr;
    size_t rgb8_cap = 0;
    float* d_scratch = nullptr;
    unsigned char* d_wf = nullptr;  // RTW_MODE_WAVEFRONT path pool
    size_t wf_cap = 0;              // bytes
    // device-side scene generator (rtw_scenegen.cu): work space + the generated list
    unsigned char* d_gen_ws = nullptr;
    size_t gen_ws_cap = 0;
    // latency path (rtw_small.cu): image and totals in mapped pinned host memory, device counters kept zero between calls
    void* h_small_img = nullptr;
    size_t small_img_cap = 0;  // bytes
    unsigned long long* h_small_tot = nullptr;
    bool counters_clean = false;
    // last resident render
    rtw_stats last = {};
    cudaStream_t last_stream = nullptr;
    bool last_valid = false;
    bool last_resolved = false;  // the last enqueue included a resolve (ev[2] is meaningful)
};

}  // namespace

// progressive image held in the accumulators of the context's devices (rtw_accumulate / rtw_resolve)
struct ProgressiveState {
    bool valid = false;
    int W = 0, s_total = 0, s_done = 0;
    // what the samples accumulated so far were traced with: a continuation pass must use the same
    int max_depth = 0;
    uint64_t seed = 0;
    uint64_t cam_hash = 0;
    bool have_inputs = false;  // false after rtw_accumulator_write (a raw checkpoint carries no render inputs)
};

struct rtw_ctx {
    std::mutex mu;
    ProgressiveState prog;
    std::string err;
    std::vector<DeviceState> dev;
    uint32_t n_spheres = 0;
    bool have_scene = false;
    uint32_t n_spheres64 = 0;
    bool have_scene64 = false;
    rtw::GridParams grid = {};  // host copy of the grid header (device pointers are per device)
    bool grid_two = false;      // the loose registration has lists of its own (large scenes)
    bool grid_valid = false;    // the grid of the current scene has been built and uploaded (lazily: RTW_MODE_GRID only)
    std::vector<float> h_geom;  // host copy of geom4, kept for the lazy grid build
    std::vector<float> h_mat;   // host copies of mat4 / kind: rtw_set_scene with an unchanged scene is a no-op
    std::vector<uint32_t> h_kind;
    uint64_t scene_hash = 0;    // FNV-1a of the flattened scene (checkpoint files are tied to it)
    int small_render = 1;       // RTW_OPT_SMALL_RENDER: small renders take the single-launch latency path
    float max_albedo = 0.f;     // largest albedo component of the scene (fixed-point head-room check)
    double max_albedo64 = 0.0;  // the same for the Float64 scene
    std::vector<double> h_geom64, h_mat64;  // host copies: rtw_set_scene_f64 with an unchanged scene is a no-op
    std::vector<uint32_t> h_kind64;
    int mode = RTW_MODE_FUSED;
    int rays_per_lane = 0;  // 0 = default
    int sweep = 0;          // 0 = default
    int coop = 0;           // 0 = default
    int tail = 0;           // RTW_TAIL_*; 0 = default
    int walk = 0;           // RTW_WALK_*; 0 = default
    int blocks_per_sm = 0;
    int collect_timing = 1;
    // framebuffer gather of the multi-device render: peer copies (default) or one grouped NCCL send/recv
    int gather = RTW_GATHER_PEER;
    void* nccl_lib = nullptr;
    std::vector<ncclComm_t> nccl_comm;  // one communicator per device of the context (ncclCommInitAll), created lazily
    ncclResult_t (*nccl_CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*nccl_CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*nccl_GroupStart)() = nullptr;
    ncclResult_t (*nccl_GroupEnd)() = nullptr;
    ncclResult_t (*nccl_Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*nccl_Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*nccl_GetErrorString)(ncclResult_t) = nullptr;
};

namespace {

// defaults chosen by measurement on B200 (profiles/): see DESIGN.md "Kernel variants"
// device counters: [0] path tickets, [1] ray segments, [2] rays RTW_MODE_GRID resolved by the exact fallback sweep, [3] spare;
// the pinned host mirror has 2 * kCounters words (the upper half is scratch of RTW_MODE_WAVEFRONT)
constexpr int kCounters = 8;  // [4] cells walked, [5] sphere tests made in RTW_MODE_GRID; [6], [7] spare
constexpr int kDefaultRaysPerLane = 1;
constexpr int kDefaultSweep = RTW_SWEEP_PACKED;
constexpr int kDefaultCoop = 2;
constexpr int kDefaultTail = RTW_TAIL_UNIFIED;

uint64_t fnv1a(const void* data, size_t bytes, uint64_t h) {
    const unsigned char* p = (const unsigned char*)data;
    for (size_t i = 0; i < bytes; ++i) h = (h ^ p[i]) * 1099511628211ull;
    return h;
}

uint32_t crc32_of(const void* data, size_t bytes, uint32_t crc) {  // CRC-32 (zlib polynomial), continuing from `crc`
    struct Table {
        uint32_t v[256];
        Table() {
            for (uint32_t n = 0; n < 256; ++n) {
                uint32_t c = n;
                for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
                v[n] = c;
            }
        }
    };
    static const Table tab;  // initialised once, thread-safe (contexts on different threads share it)
    const uint32_t* table = tab.v;
    const unsigned char* p = (const unsigned char*)data;
    uint32_t c = crc ^ 0xFFFFFFFFu;
    for (size_t i = 0; i < bytes; ++i) c = table[(c ^ p[i]) & 0xFFu] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

#ifdef RTW_BUILD_VARIANTS
constexpr bool kHaveVariants = true;
#else
constexpr bool kHaveVariants = false;
#endif

// the default build ships one configuration of the fused kernel: unified tail, packed sweep, one path per lane, two
// cooperating lanes, own-ray candidate walk (per-slot walk for streamed lists)
bool variant_is_built(bool streamed, int rays, int sweep, int coop, int tail, int walk) {
    if (kHaveVariants) return true;
    if (tail != RTW_TAIL_UNIFIED || rays != 1 || sweep != RTW_SWEEP_PACKED || coop != 2) return false;
    return streamed || walk == RTW_WALK_DEFAULT || walk == RTW_WALK_OWN_RAY;
}

int fail(rtw_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}

int cuda_fail(rtw_ctx* c, cudaError_t e, const char* what) {
    std::string m = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    if (c) c->err = m;
    (void)cudaGetLastError();
    return (int)e > 0 ? (int)e : RTW_E_INTERNAL;
}

#define RTW_CUDA(ctx, call)                                   \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call); \
    } while (0)

// RTW_GATHER_NCCL: bind libnccl at run time (the process may already hold one, e.g. torch's bundled copy -- the
// loader then returns that one) and create one communicator per device of the context.
int ensure_nccl(rtw_ctx* ctx) {
    if (!ctx->nccl_comm.empty()) return RTW_OK;
    if (!ctx->nccl_lib) {
        ctx->nccl_lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!ctx->nccl_lib) ctx->nccl_lib = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!ctx->nccl_lib) return fail(ctx, RTW_E_UNSUPPORTED, "RTW_GATHER_NCCL: libnccl.so.2 cannot be loaded");
        void* h = ctx->nccl_lib;
        ctx->nccl_CommInitAll = (decltype(ctx->nccl_CommInitAll))dlsym(h, "ncclCommInitAll");
        ctx->nccl_CommDestroy = (decltype(ctx->nccl_CommDestroy))dlsym(h, "ncclCommDestroy");
        ctx->nccl_GroupStart = (decltype(ctx->nccl_GroupStart))dlsym(h, "ncclGroupStart");
        ctx->nccl_GroupEnd = (decltype(ctx->nccl_GroupEnd))dlsym(h, "ncclGroupEnd");
        ctx->nccl_Send = (decltype(ctx->nccl_Send))dlsym(h, "ncclSend");
        ctx->nccl_Recv = (decltype(ctx->nccl_Recv))dlsym(h, "ncclRecv");
        ctx->nccl_GetErrorString = (decltype(ctx->nccl_GetErrorString))dlsym(h, "ncclGetErrorString");
        if (!ctx->nccl_CommInitAll || !ctx->nccl_CommDestroy || !ctx->nccl_GroupStart || !ctx->nccl_GroupEnd ||
            !ctx->nccl_Send || !ctx->nccl_Recv || !ctx->nccl_GetErrorString)
            return fail(ctx, RTW_E_UNSUPPORTED, "RTW_GATHER_NCCL: libnccl lacks a required symbol");
    }
    const int G = (int)ctx->dev.size();
    std::vector<int> devs(G);
    for (int g = 0; g < G; ++g) devs[g] = ctx->dev[g].device;
    std::vector<ncclComm_t> comms(G, nullptr);
    const ncclResult_t r = ctx->nccl_CommInitAll(comms.data(), G, devs.data());
    if (r != ncclSuccess) return fail(ctx, RTW_E_INTERNAL, std::string("ncclCommInitAll: ") + ctx->nccl_GetErrorString(r));
    ctx->nccl_comm = comms;
    return RTW_OK;
}

#define RTW_NCCL(ctx, call)                                                                                   \
    do {                                                                                                      \
        const ncclResult_t r__ = (call);                                                                      \
        if (r__ != ncclSuccess) return fail(ctx, RTW_E_INTERNAL, std::string(#call ": ") + ctx->nccl_GetErrorString(r__)); \
    } while (0)

template <typename T>
int grow(rtw_ctx* c, T** p, size_t* cap, size_t need) {
    if (need <= *cap && *p) return RTW_OK;
    if (*p) {
        RTW_CUDA(c, cudaFree(*p));
        *p = nullptr;
        *cap = 0;
    }
    size_t n = need ? need : 1;
    RTW_CUDA(c, cudaMalloc((void**)p, n * sizeof(T)));
    *cap = n;
    return RTW_OK;
}

rtw::DevCamera to_dev_camera(const rtw_camera* c) {
    rtw::DevCamera k;
    k.origin = rtw::f3{c->origin[0], c->origin[1], c->origin[2]};
    k.llc = rtw::f3{c->lower_left_corner[0], c->lower_left_corner[1], c->lower_left_corner[2]};
    k.horizontal = rtw::f3{c->horizontal[0], c->horizontal[1], c->horizontal[2]};
    k.vertical = rtw::f3{c->vertical[0], c->vertical[1], c->vertical[2]};
    k.u = rtw::f3{c->u[0], c->u[1], c->u[2]};
    k.v = rtw::f3{c->v[0], c->v[1], c->v[2]};
    k.lens_radius = c->lens_radius;
    return k;
}

int rows_of(int H, int row_start, int row_stride) {
    if (row_start >= H) return 0;
    return (H - row_start + row_stride - 1) / row_stride;
}

// fixed-point fraction bits of the accumulator: 62 bits total, minus ceil(log2(spp)), minus 6 bits (64x) of
// head-room for per-path radiance above 1 (the reference's scenes never exceed 1: albedo and sky are <= 1).
int fx_bits_for(int spp) {
    int b = 0;
    while ((1ll << b) < (long long)spp) ++b;
    return 62 - b - 6;
}

// The fixed-point accumulator keeps 6 bits (64x) of head-room per path above radiance 1.  A path's radiance is at most
// max_albedo^(max_depth - 1) (sky <= 1): with every albedo <= 1 -- all of the reference's scenes -- it never exceeds 1.
// A scene that uses albedo > 1 as emission is accepted as long as that bound stays below 64; beyond it the sums could
// saturate silently, so the call is refused instead.
bool radiance_fits_headroom(double max_albedo, int max_depth) {
    if (!(max_albedo > 1.0) || max_depth <= 1) return true;
    return (double)(max_depth - 1) * std::log2(max_albedo) <= 6.0;
}

int check_render_args(rtw_ctx* ctx, const rtw_camera* cam, int W, int spp, int max_depth) {
    if (!cam) return fail(ctx, RTW_E_INVALID_ARG, "camera is NULL");
    if (W < 1 || W > 65536) return fail(ctx, RTW_E_INVALID_ARG, "image_width must be in 1..65536");
    if (spp < 1 || spp > (1 << 24)) return fail(ctx, RTW_E_INVALID_ARG, "n_samples must be in 1..2^24");
    if (max_depth < 0 || max_depth > (1 << 20)) return fail(ctx, RTW_E_INVALID_ARG, "max_depth must be in 0..2^20");
    if (!ctx->have_scene) return fail(ctx, RTW_E_NO_SCENE, "rtw_set_scene has not been called");
    if (!radiance_fits_headroom(ctx->max_albedo, max_depth))
        return fail(ctx, RTW_E_UNSUPPORTED, "albedo > 1 with this max_depth can exceed the 64x head-room of the fixed-point accumulator");
    return RTW_OK;
}

// RTW_MODE_WAVEFRONT: carve the path pool (structure of arrays) out of one device allocation.  The pool holds a
// whole number of intersect-kernel waves (3 CTAs of 256 lanes per SM), at most ~1 M paths.
int wavefront_buffers(rtw_ctx* ctx, DeviceState& ds, unsigned long long n_paths, rtw::WavefrontBuffers* out) {
    const unsigned long long wave = (unsigned long long)ds.num_sms * 3ull * 256ull;
    unsigned long long cap_max = ((1ull << 20) / wave) * wave;
    if (cap_max == 0) cap_max = wave;
    const unsigned long long want = (n_paths + 255ull) & ~255ull;
    uint32_t capacity = (uint32_t)(want < cap_max ? want : cap_max);
    if (capacity < 256u) capacity = 256u;
    int rc = grow(ctx, &ds.d_wf, &ds.wf_cap, rtw::wavefront_bytes(capacity));
    if (rc) return rc;
    unsigned char* q = ds.d_wf;
    auto take = [&q](size_t bytes) {
        unsigned char* r = q;
        q += (bytes + 63) & ~(size_t)63;
        return r;
    };
    rtw::WavefrontBuffers& b = *out;
    b.capacity = capacity;
    b.ray_o = (float4*)take((size_t)capacity * 16);
    b.ray_d = (float4*)take((size_t)capacity * 16);
    b.thr = (double*)take((size_t)capacity * 24);
    b.pix_local = (uint32_t*)take((size_t)capacity * 4);
    b.sample = (uint32_t*)take((size_t)capacity * 4);
    b.pixel = (uint32_t*)take((size_t)capacity * 4);
    b.depth_left = (int*)take((size_t)capacity * 4);
    b.hit_t = (float*)take((size_t)capacity * 4);
    b.hit_k = (int*)take((size_t)capacity * 4);
    b.alive = (uint32_t*)take((size_t)capacity * 4);
    for (int l = 0; l < 3; ++l) b.list[l] = (uint32_t*)take((size_t)capacity * 4);
    b.counts = (unsigned int*)take(64);
    return RTW_OK;
}

// Which samples of the image a trace launch adds to the accumulator.  render(): all of them at once.  Progressive
// passes: samples [s_first, s_first + s_count) of an image that will hold s_total samples per pixel; the fixed-point
// scale depends on s_total only, so any split of the samples into passes sums to the same integers.
struct PassSpec {
    int s_first, s_count, s_total;
    bool reset;  // zero the accumulator first
};

int ensure_grid_locked(rtw_ctx* ctx);

// Enqueue the trace of one pass for a row subset on one device (accumulates into ds.d_accum).
int enqueue_trace(rtw_ctx* ctx, DeviceState& ds, const rtw_camera* cam, int W, int max_depth, uint64_t seed,
                  int row_start, int row_stride, const PassSpec& ps, cudaStream_t stream, bool timing) {
    const int H = rtw_image_height(W);
    const int n_rows = rows_of(H, row_start, row_stride);
    const int spp = ps.s_count;
    RTW_CUDA(ctx, cudaSetDevice(ds.device));
    ds.last = rtw_stats{};
    ds.last.n_spheres = ctx->n_spheres;
    ds.last.image_width = W;
    ds.last.image_height = H;
    ds.last.rows_rendered = n_rows;
    ds.last.paths = (uint64_t)n_rows * (uint64_t)W * (uint64_t)spp;
    ds.last_stream = stream;
    ds.last_valid = true;
    if (n_rows == 0 || H == 0) return RTW_OK;

    const size_t npix = (size_t)n_rows * (size_t)W;
    if (!ps.reset && npix * 4 > ds.accum_cap) return fail(ctx, RTW_E_INVALID_ARG, "no accumulator of this size to add samples to");
    int rc = grow(ctx, &ds.d_accum, &ds.accum_cap, npix * 4);
    if (rc) return rc;
    if (timing) RTW_CUDA(ctx, cudaEventRecord(ds.ev[0], stream));
    if (ps.reset) RTW_CUDA(ctx, cudaMemsetAsync(ds.d_accum, 0, npix * 4 * sizeof(unsigned long long), stream));
    RTW_CUDA(ctx, cudaMemsetAsync(ds.d_counters, 0, kCounters * sizeof(unsigned long long), stream));
    ds.counters_clean = false;

    const int fx_bits = fx_bits_for(ps.s_total);
    int launches = 0;
    if (max_depth > 0 && spp > 0) {
        rtw::TraceParams p;
        p.cam = to_dev_camera(cam);
        p.geom = ds.d_geom;
        p.geom_pairs = ds.d_geom_pairs;
        p.geom_perm = nullptr;
        p.u_tab = p.v_tab = nullptr;
        p.div_spp = p.div_w = rtw::MagicDiv{0u, 0u, 0u};
        std::memset(p.rk, 0, sizeof p.rk);
        p.mat = ds.d_mat;
        p.kind = ds.d_kind;
        p.n_spheres = ctx->n_spheres;
        p.W = W; p.H = H; p.spp = spp; p.max_depth = max_depth;
        p.sample_first = ps.s_first;
        p.key0 = (uint32_t)seed; p.key1 = (uint32_t)(seed >> 32);
        p.row_start = row_start; p.row_stride = row_stride; p.n_rows = n_rows;
        p.n_paths = (unsigned long long)npix * (unsigned long long)spp;
        p.accum = ds.d_accum;
        p.fx_scale = std::ldexp(1.0, fx_bits);
        p.counters = ds.d_counters;
        rtw::LaunchInfo li{};
        p.grid = rtw::GridParams{};
        if (ctx->mode == RTW_MODE_GRID) {
            rc = ensure_grid_locked(ctx);  // no-op unless this is the first grid-mode trace of the scene
            if (rc) return rc;
            RTW_CUDA(ctx, cudaSetDevice(ds.device));
            p.grid = ctx->grid;
            p.grid.big = ds.d_grid_big;
            p.grid.cull.cell_start = ds.d_cull_start;
            p.grid.cull.items = ds.d_cull_items;
            p.grid.cell_start_tight = ds.d_grid_start[0];
            p.grid.items_tight = ds.d_grid_items[0];
            p.grid.cell_start_loose = ds.d_grid_start[ctx->grid_two ? 1 : 0];
            p.grid.items_loose = ds.d_grid_items[ctx->grid_two ? 1 : 0];
            rc = grow(ctx, &ds.d_uv, &ds.uv_cap, (size_t)W + (size_t)H);
            if (rc) return rc;
            RTW_CUDA(ctx, rtw::launch_uv_tables(W, H, ds.d_uv, ds.d_uv + W, stream));
            launches += 1;
            p.u_tab = ds.d_uv;
            p.v_tab = ds.d_uv + W;
            p.div_spp = rtw::make_magic_div((uint32_t)spp);
            p.div_w = rtw::make_magic_div((uint32_t)W);
            for (uint32_t r = 0; r < (uint32_t)rtw::kPhiloxRounds; ++r) {
                p.rk[2 * r] = p.key0 + r * rtw::kPhiloxW0;
                p.rk[2 * r + 1] = p.key1 + r * rtw::kPhiloxW1;
            }
            RTW_CUDA(ctx, rtw::launch_fused_trace2_grid(p, ds.num_sms, ctx->blocks_per_sm, stream, &li));
#ifdef RTW_BUILD_VARIANTS
        } else if (ctx->mode == RTW_MODE_CTA_WAVEFRONT && ctx->n_spheres <= rtw::kTileSpheres) {
            RTW_CUDA(ctx, rtw::launch_cta_wavefront_trace(p, ds.num_sms, ctx->blocks_per_sm, stream, &li));
#endif
        } else if (ctx->mode == RTW_MODE_WAVEFRONT) {
            if (ctx->n_spheres > rtw::kTileSpheres)
                return fail(ctx, RTW_E_UNSUPPORTED, "RTW_MODE_WAVEFRONT supports at most 1024 spheres; use RTW_MODE_FUSED");
            rtw::WavefrontBuffers b;
            rc = wavefront_buffers(ctx, ds, p.n_paths, &b);
            if (rc) return rc;
            RTW_CUDA(ctx, rtw::launch_wavefront_trace(p, b, ds.num_sms, (unsigned int*)(ds.h_counters + kCounters), stream, &li));
        } else {
            const int rays = ctx->rays_per_lane > 0 ? ctx->rays_per_lane : kDefaultRaysPerLane;
            const int sweep = ctx->sweep > 0 ? ctx->sweep : kDefaultSweep;
            const int coop = ctx->coop > 0 ? ctx->coop : kDefaultCoop;
            const int tail = ctx->tail > 0 ? ctx->tail : kDefaultTail;
            // the unified tail exists for the default sweep family: packed, one path per lane, 2 or 4 cooperating lanes
            if (!variant_is_built(ctx->n_spheres > rtw::kTileSpheres, rays, sweep, coop, tail, ctx->walk))
                return fail(ctx, RTW_E_UNSUPPORTED, "this kernel variant needs a library built with RTW_BUILD_VARIANTS=1");
            if (tail == RTW_TAIL_UNIFIED && rays == 1 && sweep == RTW_SWEEP_PACKED && (coop == 2 || coop == 4)) {
                rc = grow(ctx, &ds.d_uv, &ds.uv_cap, (size_t)W + (size_t)H);
                if (rc) return rc;
                RTW_CUDA(ctx, rtw::launch_uv_tables(W, H, ds.d_uv, ds.d_uv + W, stream));
                launches += 1;
                p.geom_perm = ds.d_geom_perm[coop == 2 ? 0 : 1];
                p.u_tab = ds.d_uv;
                p.v_tab = ds.d_uv + W;
                p.div_spp = rtw::make_magic_div((uint32_t)spp);
                p.div_w = rtw::make_magic_div((uint32_t)W);
                for (uint32_t r = 0; r < (uint32_t)rtw::kPhiloxRounds; ++r) {
                    p.rk[2 * r] = p.key0 + r * rtw::kPhiloxW0;
                    p.rk[2 * r + 1] = p.key1 + r * rtw::kPhiloxW1;
                }
                RTW_CUDA(ctx, rtw::launch_fused_trace2(p, ds.num_sms, ctx->blocks_per_sm, coop, ctx->walk, stream, &li));
            } else {
                RTW_CUDA(ctx, rtw::launch_fused_trace(p, ds.num_sms, ctx->blocks_per_sm, rays, sweep, coop, stream, &li));
            }
        }
        launches += li.launches;
    }
    if (timing) RTW_CUDA(ctx, cudaEventRecord(ds.ev[1], stream));
    RTW_CUDA(ctx, cudaMemcpyAsync(ds.h_counters, ds.d_counters, kCounters * sizeof(unsigned long long),
                                  cudaMemcpyDeviceToHost, stream));
    ds.last.kernel_launches = launches;
    ds.last_resolved = false;
    return RTW_OK;
}

// Enqueue the resolve of a device's accumulator rows: accum / divisor -> sqrt -> Float32 (src/render.jl:40), into a
// row-major tile or the Julia column-major image; then the counter read-back of the preceding trace.
int enqueue_resolve(rtw_ctx* ctx, DeviceState& ds, int W, int row_start, int row_stride, int column_major, int divisor,
                    int s_total, float* d_out, cudaStream_t stream, bool timing) {
    const int H = rtw_image_height(W);
    const int n_rows = rows_of(H, row_start, row_stride);
    RTW_CUDA(ctx, cudaSetDevice(ds.device));
    if (n_rows == 0 || H == 0) return RTW_OK;
    RTW_CUDA(ctx, rtw::launch_resolve(ds.d_accum, W, H, n_rows, row_start, row_stride, divisor,
                                      std::ldexp(1.0, -fx_bits_for(s_total)), column_major, d_out, stream));
    ds.last.kernel_launches += 1;
    if (timing) RTW_CUDA(ctx, cudaEventRecord(ds.ev[2], stream));
    ds.last_resolved = true;
    return RTW_OK;
}

// trace + resolve of a whole render() for a row subset on one device
int enqueue_rows(rtw_ctx* ctx, DeviceState& ds, const rtw_camera* cam, int W, int spp, int max_depth, uint64_t seed,
                 int row_start, int row_stride, int column_major, float* d_out, cudaStream_t stream, bool timing) {
    int rc = enqueue_trace(ctx, ds, cam, W, max_depth, seed, row_start, row_stride, PassSpec{0, spp, spp, true}, stream,
                           timing);
    if (rc) return rc;
    return enqueue_resolve(ctx, ds, W, row_start, row_stride, column_major, spp, spp, d_out, stream, timing);
}

// after the stream has been synchronised: fill counters / timings
int finish_stats(rtw_ctx* ctx, DeviceState& ds, bool timing) {
    ds.last.ray_segments = ds.h_counters[1];
    ds.last.grid_fallback_rays = ds.h_counters[2];
    ds.last.grid_loose_cells = ds.h_counters[3];
    ds.last.grid_cells = ds.h_counters[4];
    ds.last.grid_tests = ds.h_counters[5];
    ds.last.sphere_tests = ds.last.ray_segments * (uint64_t)ctx->n_spheres;
    if (timing && ds.last.rows_rendered > 0) {
        float a = 0.f, b = 0.f;
        RTW_CUDA(ctx, cudaEventElapsedTime(&a, ds.ev[0], ds.ev[1]));
        if (ds.last_resolved) RTW_CUDA(ctx, cudaEventElapsedTime(&b, ds.ev[1], ds.ev[2]));
        ds.last.ms_trace = a;
        ds.last.ms_resolve = b;
        ds.last.ms_total = a + b;
    }
    return RTW_OK;
}

// RTW_MODE_GRID: bin the spheres (rtw_grid.cuh describes the structure and why the result is unchanged)
struct HostGrid {
    rtw::GridParams hdr = {};
    std::vector<uint32_t> cell_start[2], items[2], big;  // [0] tight, [1] loose (empty: the same as tight)
    std::vector<uint32_t> cull_start, cull_items;  // GridCull (long lists only)
};

void build_grid(const float* geom4, uint32_t n, HostGrid* out) {
    HostGrid& g = *out;
    g = HostGrid{};
    std::vector<float> radii(n);
    for (uint32_t i = 0; i < n; ++i) radii[i] = std::fabs(geom4[4 * (size_t)i + 3]);
    float r_med = 0.f;
    if (n > 0) {
        std::vector<float> tmp(radii);
        std::nth_element(tmp.begin(), tmp.begin() + n / 2, tmp.end());
        r_med = tmp[n / 2];
    }
    std::vector<uint32_t> small;
    for (uint32_t i = 0; i < n; ++i) {
        const float* s = geom4 + 4 * (size_t)i;
        const bool finite = std::isfinite(s[0]) && std::isfinite(s[1]) && std::isfinite(s[2]) && std::isfinite(s[3]);
        if (n <= 8 || !finite || !(r_med > 0.f) || radii[i] > 4.0f * r_med) g.big.push_back(i);  // every ray tests these
        else small.push_back(i);
    }
    if (small.empty()) return;  // no grid: hdr.nx == 0
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (uint32_t i : small) {
        const float* s = geom4 + 4 * (size_t)i;
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], (double)s[a] - radii[i]);
            hi[a] = std::max(hi[a], (double)s[a] + radii[i]);
        }
    }
    double h = 2.5 * r_med;  // a small sphere (|r| <= 4 r_med) spans at most 5 cells per axis, typically 2
    auto dims = [&](double hh, int* nd) {
        double cells = 1.0;
        for (int a = 0; a < 3; ++a) {
            // one cell of slack around the inflated boxes; formed in double and clamped before the conversion (an
            // extreme extent / radius ratio must not overflow the int)
            const double want = std::floor((hi[a] - lo[a]) / hh) + 2.0;
            nd[a] = want < 1.0e6 ? (int)want : 1000000;
            cells *= want;
        }
        return cells;
    };
    int nd[3];
    int grow_steps = 0;
    while (dims(h, nd) > 2.0 * (double)small.size() + 64.0 || nd[0] > 1024 || nd[1] > 1024 || nd[2] > 1024) {
        h *= 1.25;
        if (++grow_steps > 4096 || !std::isfinite(h)) {  // degenerate extent: no grid, every sphere is tested by every ray
            g.big.clear();
            for (uint32_t i = 0; i < n; ++i) g.big.push_back(i);
            return;
        }
    }
    double r_min = 1e300, r_max = 0.0;
    for (uint32_t i : small) {
        r_min = std::min(r_min, (double)radii[i]);
        r_max = std::max(r_max, (double)radii[i]);
    }
    const double h0 = (double)(float)h;  // the cell edge as the device sees it
    const double diag = std::sqrt((hi[0] - lo[0]) * (hi[0] - lo[0]) + (hi[1] - lo[1]) * (hi[1] - lo[1]) + (hi[2] - lo[2]) * (hi[2] - lo[2]));
    // tight margin = 5 % of a cell: rounding of the traversal, and the reference's a = 1 shortcut over moderate flights.
    // loose margin: sized so that a unit-length direction (eps = the rounding floor the device uses, kGridEpsFloor =
    // 1e-6, plus as much again for |d|^2 - 1) stays covered over 1.5 diagonals of the box, at most one cell.  Small
    // scenes (the reference's 22 x 22 field) do not need it: there the loose registration is the tight one.
    const double margin0 = 0.05 * h0;
    const double far = 1.5 * diag;
    double margin1 = std::min(std::sqrt(r_min * r_min + 2.0 * 2e-6 * far * far) - r_min, h0);
    const bool two = margin1 > 1.2 * margin0;
    if (!two) margin1 = margin0;
    rtw::GridParams& G = g.hdr;
    G.h = (float)h0;
    G.inv_h = 1.0f / G.h;
    G.safe2_tight = (float)(0.5 * ((r_min + margin0) * (r_min + margin0) - r_min * r_min));  // half of it: slack for rounding
    G.safe2_loose = (float)(0.5 * ((r_min + margin1) * (r_min + margin1) - r_min * r_min));
    G.reach = (float)(r_max + margin1);
    G.r_min = (float)r_min;
    // the box: [lo, hi] holds every small sphere; pad it by the loose margin + half a cell, so that a ray point within
    // r + margin of a centre lies inside (the walk then visits a cell the sphere is registered in)
    const double pad = margin1 + 0.5 * h0;
    G.pad = (float)(0.999 * pad);
    G.ox = (float)(lo[0] - pad);
    G.oy = (float)(lo[1] - pad);
    G.oz = (float)(lo[2] - pad);
    const double org[3] = {(double)G.ox, (double)G.oy, (double)G.oz};
    int ndl[3];
    for (int a = 0; a < 3; ++a) ndl[a] = std::max(1, (int)std::floor((hi[a] + pad - org[a]) / h0) + 1);
    G.nx = ndl[0]; G.ny = ndl[1]; G.nz = ndl[2];
    G.ball_r = (float)(0.5 * h0 * std::sqrt((double)ndl[0] * ndl[0] + (double)ndl[1] * ndl[1] + (double)ndl[2] * ndl[2]) * 1.001);
    const size_t ncell = (size_t)ndl[0] * ndl[1] * ndl[2];
    // CSR lists of the spheres whose AABB, inflated by `margin`, touches each cell
    auto bin = [&](double margin, std::vector<uint32_t>& cell_start, std::vector<uint32_t>& items) {
        auto range = [&](uint32_t i, int a, int* c0, int* c1) {
            const float* s = geom4 + 4 * (size_t)i;
            *c0 = std::min(std::max((int)std::floor(((double)s[a] - radii[i] - margin - org[a]) / h0), 0), ndl[a] - 1);
            *c1 = std::min(std::max((int)std::floor(((double)s[a] + radii[i] + margin - org[a]) / h0), 0), ndl[a] - 1);
        };
        cell_start.assign(ncell + 1, 0u);
        for (int pass = 0; pass < 2; ++pass) {
            std::vector<uint32_t> fill;
            if (pass == 1) {
                for (size_t c = 0, acc = 0; c <= ncell; ++c) {  // counts -> offsets
                    const uint32_t cnt = c < ncell ? cell_start[c] : 0u;
                    cell_start[c] = (uint32_t)acc;
                    acc += cnt;
                }
                items.assign(cell_start[ncell], 0u);
                fill.assign(cell_start.begin(), cell_start.end() - 1);
            }
            for (uint32_t i : small) {  // ascending sphere index, so the items of a cell are ascending too
                int x0, x1, y0, y1, z0, z1;
                range(i, 0, &x0, &x1);
                range(i, 1, &y0, &y1);
                range(i, 2, &z0, &z1);
                for (int z = z0; z <= z1; ++z)
                    for (int y = y0; y <= y1; ++y)
                        for (int x = x0; x <= x1; ++x) {
                            const size_t c = (size_t)x + (size_t)ndl[0] * ((size_t)y + (size_t)ndl[1] * z);
                            if (pass == 0) cell_start[c] += 1u;
                            else items[fill[c]++] = i;
                        }
            }
        }
    };
    bin(margin0, g.cell_start[0], g.items[0]);
    if (two) bin(margin1, g.cell_start[1], g.items[1]);
    // (Coarser levels with a margin of one coarse cell were measured and dropped: a lane walking 100-sphere cells alone
    // holds its warp for ~30 ordinary segments.  What the loose level cannot cover goes to the cooperative sweep.)
    if (small.size() > 4096) {
        // GridCull: every small sphere once, binned by its centre into cells of 16 fine cells; the whole-list sweep of an
        // unsafe ray visits only the cells its cone of acceptance can reach
        rtw::GridCull& C = g.hdr.cull;
        C.h = (float)(16.0 * h0);
        C.ox = (float)lo[0]; C.oy = (float)lo[1]; C.oz = (float)lo[2];
        C.r_max = (float)r_max;
        int nc[3];
        for (int a = 0; a < 3; ++a) nc[a] = std::max(1, (int)std::floor((hi[a] - (double)(&C.ox)[a]) / (double)C.h) + 1);
        C.nx = nc[0]; C.ny = nc[1]; C.nz = nc[2];
        const size_t ncell = (size_t)nc[0] * nc[1] * nc[2];
        auto cell_of = [&](uint32_t i) {
            const float* s = geom4 + 4 * (size_t)i;
            size_t c[3];
            for (int a = 0; a < 3; ++a)
                c[a] = (size_t)std::min(std::max((int)std::floor(((double)s[a] - (double)(&C.ox)[a]) / (double)C.h), 0), nc[a] - 1);
            return c[0] + (size_t)nc[0] * (c[1] + (size_t)nc[1] * c[2]);
        };
        g.cull_start.assign(ncell + 1, 0u);
        for (uint32_t i : small) g.cull_start[cell_of(i) + 1] += 1u;
        for (size_t c = 0; c < ncell; ++c) g.cull_start[c + 1] += g.cull_start[c];
        g.cull_items.assign(small.size(), 0u);
        std::vector<uint32_t> fill(g.cull_start.begin(), g.cull_start.end() - 1);
        for (uint32_t i : small) g.cull_items[fill[cell_of(i)]++] = i;
    }
}

// RTW_MODE_GRID only, on first use after rtw_set_scene: build the grid from the host copy of the geometry and upload
// it to every device of the context (the default mode never pays for it)
int ensure_grid_locked(rtw_ctx* ctx) {
    if (ctx->grid_valid) return RTW_OK;
    HostGrid hg;
    build_grid(ctx->h_geom.data(), ctx->n_spheres, &hg);
    for (auto& ds : ctx->dev) {
        RTW_CUDA(ctx, cudaSetDevice(ds.device));
        int grc = grow(ctx, &ds.d_grid_big, &ds.grid_big_cap, hg.big.size());
        for (int l = 0; l < 2 && !grc; ++l) {
            grc = grow(ctx, &ds.d_grid_start[l], &ds.grid_start_cap[l], hg.cell_start[l].size());
            if (!grc) grc = grow(ctx, &ds.d_grid_items[l], &ds.grid_items_cap[l], hg.items[l].size());
        }
        if (!grc) grc = grow(ctx, &ds.d_cull_start, &ds.cull_start_cap, hg.cull_start.size());
        if (!grc) grc = grow(ctx, &ds.d_cull_items, &ds.cull_items_cap, hg.cull_items.size());
        if (grc) return grc;
        if (!hg.cull_start.empty()) {
            RTW_CUDA(ctx, cudaMemcpyAsync(ds.d_cull_start, hg.cull_start.data(), hg.cull_start.size() * 4, cudaMemcpyHostToDevice, ds.stream));
            RTW_CUDA(ctx, cudaMemcpyAsync(ds.d_cull_items, hg.cull_items.data(), hg.cull_items.size() * 4, cudaMemcpyHostToDevice, ds.stream));
        }
        if (!hg.big.empty())
            RTW_CUDA(ctx, cudaMemcpyAsync(ds.d_grid_big, hg.big.data(), hg.big.size() * 4, cudaMemcpyHostToDevice, ds.stream));
        for (int l = 0; l < 2; ++l) {
            if (hg.cell_start[l].empty()) continue;
            RTW_CUDA(ctx, cudaMemcpyAsync(ds.d_grid_start[l], hg.cell_start[l].data(), hg.cell_start[l].size() * 4, cudaMemcpyHostToDevice, ds.stream));
            if (!hg.items[l].empty())
                RTW_CUDA(ctx, cudaMemcpyAsync(ds.d_grid_items[l], hg.items[l].data(), hg.items[l].size() * 4, cudaMemcpyHostToDevice, ds.stream));
        }
    }
    for (auto& ds : ctx->dev) {  // the host vectors die with this call
        RTW_CUDA(ctx, cudaSetDevice(ds.device));
        RTW_CUDA(ctx, cudaStreamSynchronize(ds.stream));
    }
    ctx->grid = hg.hdr;
    ctx->grid.n_big = (uint32_t)hg.big.size();
    ctx->grid_two = !hg.cell_start[1].empty();
    ctx->grid_valid = true;
    return RTW_OK;
}

int set_scene_locked(rtw_ctx* ctx, const float* geom4, const float* mat4, const uint32_t* kind, uint32_t n) {
    if (n > 0 && (!geom4 || !mat4 || !kind)) return fail(ctx, RTW_E_INVALID_ARG, "scene arrays are NULL");
    if (n > (1u << 26)) return fail(ctx, RTW_E_INVALID_ARG, "too many spheres");
    float max_albedo = 0.f;
    for (uint32_t i = 0; i < n; ++i) {
        if (kind[i] > RTW_DIELECTRIC) return fail(ctx, RTW_E_UNSUPPORTED, "unknown material kind (only Lambertian/Metal/Dielectric)");
        if (kind[i] != RTW_DIELECTRIC)
            for (int c = 0; c < 3; ++c) {
                const float a = mat4[4 * (size_t)i + c];
                if (!std::isfinite(a) || a < 0.f)
                    return fail(ctx, RTW_E_UNSUPPORTED, "albedo components must be finite and >= 0 (fixed-point accumulator)");
                max_albedo = std::max(max_albedo, a);
            }
    }
    // The same scene again (the drop-in render(scene, cam, ...) passes it with every call): nothing to upload.
    if (ctx->have_scene && n == ctx->n_spheres && ctx->h_geom.size() == 4 * (size_t)n &&
        (n == 0 || (std::memcmp(ctx->h_geom.data(), geom4, 16 * (size_t)n) == 0 &&
                    std::memcmp(ctx->h_mat.data(), mat4, 16 * (size_t)n) == 0 &&
                    std::memcmp(ctx->h_kind.data(), kind, 4 * (size_t)n) == 0)))
        return RTW_OK;
    // From here on the previous scene is gone: a failure below must not leave a context that claims to hold one
    // (device arrays freed / half uploaded).  Re-validated only after every device has synchronised.
    ctx->have_scene = false;
    ctx->n_spheres = 0;
    ctx->grid_valid = false;
    ctx->grid = rtw::GridParams{};
    ctx->prog = ProgressiveState{};  // a progressive image belongs to the scene it was traced on
    ctx->h_geom.assign(geom4, geom4 + 4 * (size_t)n);  // for the lazy grid build (RTW_MODE_GRID only)
    ctx->h_mat.assign(mat4, mat4 + 4 * (size_t)n);
    ctx->h_kind.assign(kind, kind + (size_t)n);
    ctx->scene_hash = fnv1a(kind, 4 * (size_t)n, fnv1a(mat4, 16 * (size_t)n, fnv1a(geom4, 16 * (size_t)n, 1469598103934665603ull)));
    ctx->max_albedo = max_albedo;
    // pair layout of the geometry for the packed sweep: spheres (2p, 2p+1) -> {xa,xb,ya,yb}{za,zb,ra,rb}
    std::vector<float> pairs((size_t)((n + 1u) / 2u) * 8u, 0.0f);
    for (uint32_t i = 0; i < n; ++i) {
        float* dst = pairs.data() + (size_t)(i >> 1) * 8u + (i & 1u);
        dst[0] = geom4[4 * (size_t)i + 0];
        dst[2] = geom4[4 * (size_t)i + 1];
        dst[4] = geom4[4 * (size_t)i + 2];
        dst[6] = geom4[4 * (size_t)i + 3];
    }
    // RTW_TAIL_UNIFIED: copies of the list in the order the lanes of a cooperating group meet the spheres
    // (single-tile lists only): entry (c*coop + h)*32 + j = sphere c*32*coop + 2*((j>>1)*coop + h) + (j&1)
    std::vector<float> perm[2];
    size_t perm_entries[2] = {0, 0};
    if (n > 0 && n <= rtw::kTileSpheres) {
        for (int v = 0; v < 2; ++v) {
            const uint32_t coop = v == 0 ? 2u : 4u, super = 32u * coop;
            const uint32_t nsc = (n + super - 1u) / super;
            perm_entries[v] = (size_t)nsc * super;
            perm[v].assign(perm_entries[v] * 4u, 0.0f);
            for (uint32_t kl = 0; kl < n; ++kl) {
                const uint32_t c = kl / super, within = kl % super, pr = within >> 1, half = within & 1u;
                const uint32_t h = pr % coop, i = pr / coop, j = 2u * i + half;
                std::memcpy(perm[v].data() + ((size_t)(c * coop + h) * 32u + j) * 4u, geom4 + 4 * (size_t)kl, 16);
            }
        }
    }
    for (auto& ds : ctx->dev) {
        RTW_CUDA(ctx, cudaSetDevice(ds.device));
        if (n > ds.scene_cap || !ds.d_geom) {
            if (ds.d_geom) cudaFree(ds.d_geom);
            if (ds.d_geom_pairs) cudaFree(ds.d_geom_pairs);
            for (auto& q : ds.d_geom_perm) { if (q) cudaFree(q); q = nullptr; }
            if (ds.d_mat) cudaFree(ds.d_mat);
            if (ds.d_kind) cudaFree(ds.d_kind);
            ds.d_geom = nullptr; ds.d_geom_pairs = nullptr; ds.d_mat = nullptr; ds.d_kind = nullptr; ds.scene_cap = 0;
            size_t cap = n ? n : 1;
            RTW_CUDA(ctx, cudaMalloc((void**)&ds.d_geom, cap * sizeof(float4)));
            RTW_CUDA(ctx, cudaMalloc((void**)&ds.d_geom_pairs, (cap + 1) * sizeof(float4)));
            for (auto& q : ds.d_geom_perm) RTW_CUDA(ctx, cudaMalloc((void**)&q, (cap + 128) * sizeof(float4)));
            RTW_CUDA(ctx, cudaMalloc((void**)&ds.d_mat, cap * sizeof(float4)));
            RTW_CUDA(ctx, cudaMalloc((void**)&ds.d_kind, cap * sizeof(uint32_t)));
            ds.scene_cap = cap;
        }
        if (n) {
            RTW_CUDA(ctx, cudaMemcpyAsync(ds.d_geom, geom4, (size_t)n * 16, cudaMemcpyHostToDevice, ds.stream));
            RTW_CUDA(ctx, cudaMemcpyAsync(ds.d_geom_pairs, pairs.data(), pairs.size() * sizeof(float), cudaMemcpyHostToDevice, ds.stream));
            RTW_CUDA(ctx, cudaMemcpyAsync(ds.d_mat, mat4, (size_t)n * 16, cudaMemcpyHostToDevice, ds.stream));
            RTW_CUDA(ctx, cudaMemcpyAsync(ds.d_kind, kind, (size_t)n * 4, cudaMemcpyHostToDevice, ds.stream));
            for (int v = 0; v < 2; ++v)
                if (perm_entries[v])
                    RTW_CUDA(ctx, cudaMemcpyAsync(ds.d_geom_perm[v], perm[v].data(), perm_entries[v] * 16, cudaMemcpyHostToDevice, ds.stream));
        }
    }
    for (auto& ds : ctx->dev) {
        RTW_CUDA(ctx, cudaSetDevice(ds.device));
        RTW_CUDA(ctx, cudaStreamSynchronize(ds.stream));  // host arrays may be freed by the caller after return
    }
    ctx->n_spheres = n;
    ctx->have_scene = true;
    return RTW_OK;
}

// One pass over all devices of the context: trace (rows r -> device r mod G) and / or resolve + collect on device 0
// + download.  render() = trace and resolve of all samples; progressive rendering = trace-only passes followed by
// resolve-only calls (rtw_accumulate / rtw_resolve / rtw_resolve_rgb8).
//   out_rgb  : host, Julia column-major Float32 image (or NULL)
//   out_rgb8 : host, row-major 8-bit image, top row first (or NULL)
int pass_locked(rtw_ctx* ctx, const rtw_camera* cam, int W, int max_depth, uint64_t seed, const PassSpec& ps, bool do_trace,
                bool do_resolve, int divisor, float* out_rgb, uint8_t* out_rgb8, rtw_stats* stats,
                bool scene_uploaded_in_call) {
    const int H = rtw_image_height(W);
    const int G = (int)ctx->dev.size();
    const size_t img_floats = (size_t)W * (size_t)H * 3;
    DeviceState& d0 = ctx->dev[0];
    const bool timing = ctx->collect_timing != 0;
    int rc;
    RTW_CUDA(ctx, cudaSetDevice(d0.device));
    // ev[3] = start of the call (recorded by rtw_render_scene before the upload when the scene travels with it)
    if (!scene_uploaded_in_call) RTW_CUDA(ctx, cudaEventRecord(d0.ev[3], d0.stream));
    RTW_CUDA(ctx, cudaEventRecord(d0.ev[6], d0.stream));
    if (do_trace) {
        for (int g = 0; g < G; ++g) {
            DeviceState& ds = ctx->dev[g];
            rc = enqueue_trace(ctx, ds, cam, W, max_depth, seed, g, G, ps, ds.stream, timing);
            if (rc) return rc;
        }
    } else {
        for (auto& ds : ctx->dev) {  // stats of this call: nothing traced
            ds.last = rtw_stats{};
            ds.last.n_spheres = ctx->n_spheres;
            ds.last.image_width = W;
            ds.last.image_height = H;
            ds.last_resolved = false;
            ds.h_counters[1] = 0;
            for (int i = 2; i < kCounters; ++i) ds.h_counters[i] = 0;
        }
    }
    if (do_resolve) {
        RTW_CUDA(ctx, cudaSetDevice(d0.device));
        rc = grow(ctx, &d0.d_image, &d0.image_cap, img_floats);
        if (rc) return rc;
        if (G == 1) {
            rc = enqueue_resolve(ctx, d0, W, 0, 1, 1, divisor, ps.s_total, d0.d_image, d0.stream, timing);
            if (rc) return rc;
        } else {
            const int rows_pad = (H + G - 1) / G;
            const size_t tile_floats = (size_t)rows_pad * W * 3;
            rc = grow(ctx, &d0.d_gather, &d0.gather_cap, tile_floats * G);
            if (rc) return rc;
            const bool use_nccl = ctx->gather == RTW_GATHER_NCCL;
            if (use_nccl) {
                rc = ensure_nccl(ctx);
                if (rc) return rc;
            }
            for (int g = 0; g < G; ++g) {
                DeviceState& ds = ctx->dev[g];
                RTW_CUDA(ctx, cudaSetDevice(ds.device));
                float* dst = d0.d_gather + tile_floats * g;
                float* tile = dst;
                if (g != 0) {
                    rc = grow(ctx, &ds.d_tile, &ds.tile_cap, tile_floats);
                    if (rc) return rc;
                    tile = ds.d_tile;
                }
                rc = enqueue_resolve(ctx, ds, W, g, G, 0, divisor, ps.s_total, tile, ds.stream, timing);
                if (rc) return rc;
                if (g != 0 && !use_nccl) {
                    // framebuffer gather: tile -> device 0 over NVLink (peer copy on the producer's stream)
                    size_t bytes = (size_t)rows_of(H, g, G) * W * 3 * sizeof(float);
                    if (bytes) RTW_CUDA(ctx, cudaMemcpyPeerAsync(dst, d0.device, tile, ds.device, bytes, ds.stream));
                    RTW_CUDA(ctx, cudaEventRecord(ds.ev_tile, ds.stream));
                }
            }
            RTW_CUDA(ctx, cudaSetDevice(d0.device));
            if (use_nccl) {
                // the single NCCL gather: every device sends its tile to device 0, which receives them behind its own
                // resolve -- one grouped call, the streams of the context carry the ordering
                RTW_NCCL(ctx, ctx->nccl_GroupStart());
                for (int g = 1; g < G; ++g) {
                    const size_t count = (size_t)rows_of(H, g, G) * W * 3;
                    if (!count) continue;
                    RTW_NCCL(ctx, ctx->nccl_Send(ctx->dev[g].d_tile, count, ncclFloat, 0, ctx->nccl_comm[g], ctx->dev[g].stream));
                    RTW_NCCL(ctx, ctx->nccl_Recv(d0.d_gather + tile_floats * g, count, ncclFloat, g, ctx->nccl_comm[0], d0.stream));
                }
                RTW_NCCL(ctx, ctx->nccl_GroupEnd());
                RTW_CUDA(ctx, cudaSetDevice(d0.device));
            } else {
                for (int g = 1; g < G; ++g) RTW_CUDA(ctx, cudaStreamWaitEvent(d0.stream, ctx->dev[g].ev_tile, 0));
            }
            RTW_CUDA(ctx, rtw::launch_assemble(d0.d_gather, G, W, H, d0.d_image, d0.stream));
        }
    }
    RTW_CUDA(ctx, cudaSetDevice(d0.device));
    RTW_CUDA(ctx, cudaEventRecord(d0.ev[4], d0.stream));
    int extra_launches = (do_resolve && G > 1) ? 1 : 0;
    if (do_resolve && img_floats) {
        if (out_rgb)
            RTW_CUDA(ctx, cudaMemcpyAsync(out_rgb, d0.d_image, img_floats * sizeof(float), cudaMemcpyDeviceToHost, d0.stream));
        if (out_rgb8) {
            // 8-bit image: quantised on the device into its own buffer, then downloaded
            rc = grow(ctx, &d0.d_rgb8, &d0.rgb8_cap, img_floats);
            if (rc) return rc;
            RTW_CUDA(ctx, rtw::launch_quantize_rgb8(d0.d_image, W, H, d0.d_rgb8, d0.stream));
            extra_launches += 1;
            RTW_CUDA(ctx, cudaMemcpyAsync(out_rgb8, d0.d_rgb8, img_floats, cudaMemcpyDeviceToHost, d0.stream));
        }
    }
    RTW_CUDA(ctx, cudaEventRecord(d0.ev[5], d0.stream));
    for (int g = G - 1; g >= 0; --g) {
        RTW_CUDA(ctx, cudaSetDevice(ctx->dev[g].device));
        RTW_CUDA(ctx, cudaStreamSynchronize(ctx->dev[g].stream));
    }
    rtw_stats total = {};
    total.n_spheres = ctx->n_spheres;
    total.image_width = W;
    total.image_height = H;
    for (int g = 0; g < G; ++g) {
        DeviceState& ds = ctx->dev[g];
        RTW_CUDA(ctx, cudaSetDevice(ds.device));
        if (do_trace) {
            rc = finish_stats(ctx, ds, timing);
            if (rc) return rc;
        }
        total.paths += ds.last.paths;
        total.ray_segments += ds.last.ray_segments;
        total.grid_fallback_rays += ds.last.grid_fallback_rays;
        total.grid_loose_cells += ds.last.grid_loose_cells;
        total.grid_cells += ds.last.grid_cells;
        total.grid_tests += ds.last.grid_tests;
        total.sphere_tests += ds.last.sphere_tests;
        total.rows_rendered += ds.last.rows_rendered;
        total.kernel_launches += ds.last.kernel_launches;
        total.ms_trace = std::fmax(total.ms_trace, ds.last.ms_trace);
        total.ms_resolve = std::fmax(total.ms_resolve, ds.last.ms_resolve);
    }
    total.kernel_launches += extra_launches;
    total.n_devices = G;
    RTW_CUDA(ctx, cudaSetDevice(d0.device));
    float ms = 0.f;
    RTW_CUDA(ctx, cudaEventElapsedTime(&ms, d0.ev[3], d0.ev[5]));
    total.ms_total = ms;
    RTW_CUDA(ctx, cudaEventElapsedTime(&ms, d0.ev[4], d0.ev[5]));
    total.ms_d2h = ms;
    RTW_CUDA(ctx, cudaEventElapsedTime(&ms, d0.ev[3], d0.ev[6]));
    total.ms_h2d = ms;
    if (stats) *stats = total;
    return RTW_OK;
}

// The latency path: a small render() on one device in ONE kernel launch (rtw_small.cu) -- no memsets, no u/v tables, no
// resolve kernel, no enqueued copies; the image lands in mapped pinned memory and is memcpy'd to the caller's buffer.
// mapped pinned image / totals of the latency path, device counters zeroed if a persistent kernel left them dirty
int small_render_buffers(rtw_ctx* ctx, DeviceState& ds, size_t img_bytes) {
    if (img_bytes > ds.small_img_cap || !ds.h_small_img) {
        if (ds.h_small_img) RTW_CUDA(ctx, cudaFreeHost(ds.h_small_img));
        ds.h_small_img = nullptr;
        ds.small_img_cap = 0;
        const size_t cap = std::max<size_t>(img_bytes, 1 << 18);
        RTW_CUDA(ctx, cudaHostAlloc(&ds.h_small_img, cap, cudaHostAllocMapped));
        ds.small_img_cap = cap;
    }
    if (!ds.h_small_tot) RTW_CUDA(ctx, cudaHostAlloc((void**)&ds.h_small_tot, 64, cudaHostAllocMapped));
    if (!ds.counters_clean) {
        RTW_CUDA(ctx, cudaMemsetAsync(ds.d_counters, 0, kCounters * sizeof(unsigned long long), ds.stream));
        ds.counters_clean = true;  // the kernel leaves them zero
    }
    ds.h_small_tot[0] = 0;
    return RTW_OK;
}

bool small_render_eligible(const rtw_ctx* ctx, int W, int spp, int max_depth) {
    if (!ctx->small_render || ctx->dev.size() != 1 || ctx->mode != RTW_MODE_FUSED) return false;
    if (ctx->rays_per_lane || ctx->sweep || ctx->coop || ctx->tail || ctx->walk || ctx->blocks_per_sm) return false;  // a variant was asked for
    if (ctx->n_spheres > rtw::kTileSpheres || max_depth < 1) return false;
    const long long H = rtw_image_height(W);
    const long long paths = (long long)W * H * spp;
    return H > 0 && paths <= (1 << 17) && paths * (long long)std::max<uint32_t>(ctx->n_spheres, 8u) <= (1ll << 22);
}

int small_render_locked(rtw_ctx* ctx, const rtw_camera* cam, int W, int spp, int max_depth, uint64_t seed, float* out_rgb,
                        rtw_stats* stats) {
    DeviceState& ds = ctx->dev[0];
    const int H = rtw_image_height(W);
    const size_t img_floats = (size_t)W * (size_t)H * 3;
    const bool timing = ctx->collect_timing != 0;
    RTW_CUDA(ctx, cudaSetDevice(ds.device));
    int rc0 = small_render_buffers(ctx, ds, img_floats * sizeof(float));
    if (rc0) return rc0;
    rtw::TraceParams p{};
    p.cam = to_dev_camera(cam);
    p.geom = ds.d_geom;
    p.mat = ds.d_mat;
    p.kind = ds.d_kind;
    p.n_spheres = ctx->n_spheres;
    p.W = W; p.H = H; p.spp = spp; p.max_depth = max_depth;
    p.sample_first = 0;
    p.key0 = (uint32_t)seed; p.key1 = (uint32_t)(seed >> 32);
    p.row_start = 0; p.row_stride = 1; p.n_rows = H;
    p.n_paths = (unsigned long long)W * H * spp;
    const int fx_bits = fx_bits_for(spp);
    p.fx_scale = std::ldexp(1.0, fx_bits);
    p.counters = ds.d_counters;
    float* d_img = nullptr;
    unsigned long long* d_tot = nullptr;
    RTW_CUDA(ctx, cudaHostGetDevicePointer((void**)&d_img, ds.h_small_img, 0));
    RTW_CUDA(ctx, cudaHostGetDevicePointer((void**)&d_tot, ds.h_small_tot, 0));
    rtw::LaunchInfo li{};
    if (timing) RTW_CUDA(ctx, cudaEventRecord(ds.ev[0], ds.stream));
    RTW_CUDA(ctx, rtw::launch_small_render(p, std::ldexp(1.0, -fx_bits), d_img, d_tot, ds.stream, &li));
    if (timing) RTW_CUDA(ctx, cudaEventRecord(ds.ev[1], ds.stream));
    RTW_CUDA(ctx, cudaStreamSynchronize(ds.stream));
    std::memcpy(out_rgb, ds.h_small_img, img_floats * sizeof(float));
    rtw_stats st = {};
    st.n_spheres = ctx->n_spheres;
    st.image_width = W;
    st.image_height = H;
    st.rows_rendered = H;
    st.paths = p.n_paths;
    st.ray_segments = ds.h_small_tot[0];
    st.sphere_tests = st.ray_segments * (uint64_t)ctx->n_spheres;
    st.kernel_launches = 1;
    st.n_devices = 1;
    if (timing) {
        float ms = 0.f;
        RTW_CUDA(ctx, cudaEventElapsedTime(&ms, ds.ev[0], ds.ev[1]));
        st.ms_trace = st.ms_total = ms;
    }
    ds.last = st;
    ds.last_valid = true;
    ds.last_stream = ds.stream;
    ds.last_resolved = true;
    ds.h_counters[1] = st.ray_segments;  // rtw_last_stats re-reads the pinned mirror
    for (int i = 2; i < kCounters; ++i) ds.h_counters[i] = 0;
    if (stats) *stats = st;
    return RTW_OK;
}

int render_locked(rtw_ctx* ctx, const rtw_camera* cam, int W, int spp, int max_depth, uint64_t seed, float* out_rgb,
                  rtw_stats* stats, bool scene_uploaded_in_call) {
    int rc = check_render_args(ctx, cam, W, spp, max_depth);
    if (rc) return rc;
    if (!out_rgb) return fail(ctx, RTW_E_INVALID_ARG, "out_rgb is NULL");
    ctx->prog = ProgressiveState{};  // a plain render() owns the accumulators: any progressive image is gone
    if (small_render_eligible(ctx, W, spp, max_depth)) return small_render_locked(ctx, cam, W, spp, max_depth, seed, out_rgb, stats);
    return pass_locked(ctx, cam, W, max_depth, seed, PassSpec{0, spp, spp, true}, true, true, spp, out_rgb, nullptr, stats,
                       scene_uploaded_in_call);
}

// ---- Float64 path ----------------------------------------------------------------------------------------------
int set_scene_f64_locked(rtw_ctx* ctx, const double* geom4, const double* mat4, const uint32_t* kind, uint32_t n) {
    if (n > 0 && (!geom4 || !mat4 || !kind)) return fail(ctx, RTW_E_INVALID_ARG, "scene arrays are NULL");
    if (n > (1u << 26)) return fail(ctx, RTW_E_INVALID_ARG, "too many spheres");
    for (uint32_t i = 0; i < n; ++i)
        if (kind[i] > RTW_DIELECTRIC) return fail(ctx, RTW_E_UNSUPPORTED, "unknown material kind (only Lambertian/Metal/Dielectric)");
    double max_albedo = 0.0;
    for (uint32_t i = 0; i < n; ++i)
        if (kind[i] != RTW_DIELECTRIC)
            for (int c = 0; c < 3; ++c) {
                const double a = mat4[4 * (size_t)i + c];
                if (!std::isfinite(a) || a < 0.0)
                    return fail(ctx, RTW_E_UNSUPPORTED, "albedo components must be finite and >= 0 (fixed-point accumulator)");
                max_albedo = std::max(max_albedo, a);
            }
    if (ctx->have_scene64 && n == ctx->n_spheres64 && ctx->h_geom64.size() == 4 * (size_t)n &&
        (n == 0 || (std::memcmp(ctx->h_geom64.data(), geom4, 32 * (size_t)n) == 0 &&
                    std::memcmp(ctx->h_mat64.data(), mat4, 32 * (size_t)n) == 0 &&
                    std::memcmp(ctx->h_kind64.data(), kind, 4 * (size_t)n) == 0)))
        return RTW_OK;  // the same scene again: nothing to upload
    ctx->have_scene64 = false;  // failure-atomic: see set_scene_locked
    ctx->n_spheres64 = 0;
    ctx->max_albedo64 = max_albedo;
    ctx->h_geom64.assign(geom4, geom4 + 4 * (size_t)n);
    ctx->h_mat64.assign(mat4, mat4 + 4 * (size_t)n);
    ctx->h_kind64.assign(kind, kind + (size_t)n);
    for (auto& ds : ctx->dev) {
        RTW_CUDA(ctx, cudaSetDevice(ds.device));
        if (n > ds.scene64_cap || !ds.d_geom64) {
            cudaFree(ds.d_geom64); cudaFree(ds.d_mat64); cudaFree(ds.d_kind64);
            ds.d_geom64 = nullptr; ds.d_mat64 = nullptr; ds.d_kind64 = nullptr; ds.scene64_cap = 0;
            const size_t cap = n ? n : 1;
            RTW_CUDA(ctx, cudaMalloc((void**)&ds.d_geom64, cap * sizeof(double4)));
            RTW_CUDA(ctx, cudaMalloc((void**)&ds.d_mat64, cap * sizeof(double4)));
            RTW_CUDA(ctx, cudaMalloc((void**)&ds.d_kind64, cap * sizeof(uint32_t)));
            ds.scene64_cap = cap;
        }
        if (n) {
            RTW_CUDA(ctx, cudaMemcpyAsync(ds.d_geom64, geom4, (size_t)n * 32, cudaMemcpyHostToDevice, ds.stream));
            RTW_CUDA(ctx, cudaMemcpyAsync(ds.d_mat64, mat4, (size_t)n * 32, cudaMemcpyHostToDevice, ds.stream));
            RTW_CUDA(ctx, cudaMemcpyAsync(ds.d_kind64, kind, (size_t)n * 4, cudaMemcpyHostToDevice, ds.stream));
        }
    }
    for (auto& ds : ctx->dev) {
        RTW_CUDA(ctx, cudaSetDevice(ds.device));
        RTW_CUDA(ctx, cudaStreamSynchronize(ds.stream));
    }
    ctx->n_spheres64 = n;
    ctx->have_scene64 = true;
    return RTW_OK;
}

int render_f64_locked(rtw_ctx* ctx, const rtw_camera_f64* cam, int W, int spp, int max_depth, uint64_t seed, double* out_rgb,
                      rtw_stats* stats, bool scene_uploaded_in_call) {
    if (!cam) return fail(ctx, RTW_E_INVALID_ARG, "camera is NULL");
    if (W < 1 || W > 65536) return fail(ctx, RTW_E_INVALID_ARG, "image_width must be in 1..65536");
    if (spp < 1 || spp > (1 << 24)) return fail(ctx, RTW_E_INVALID_ARG, "n_samples must be in 1..2^24");
    if (max_depth < 0 || max_depth > (1 << 20)) return fail(ctx, RTW_E_INVALID_ARG, "max_depth must be in 0..2^20");
    if (!ctx->have_scene64) return fail(ctx, RTW_E_NO_SCENE, "rtw_set_scene_f64 has not been called");
    if (!radiance_fits_headroom(ctx->max_albedo64, max_depth))
        return fail(ctx, RTW_E_UNSUPPORTED, "albedo > 1 with this max_depth can exceed the 64x head-room of the fixed-point accumulator");
    if (!out_rgb) return fail(ctx, RTW_E_INVALID_ARG, "out_rgb is NULL");
    ctx->prog = ProgressiveState{};  // the accumulators are reused
    const int H = rtw_image_height(W);
    const int G = (int)ctx->dev.size();
    const size_t img_vals = (size_t)W * (size_t)H * 3;
    const int fx_bits = fx_bits_for(spp);
    const bool timing = ctx->collect_timing != 0;
    DeviceState& d0 = ctx->dev[0];
    int rc;
    // the latency path (small_render_f64_kernel): the reference's own smoke test is a 96x54x16 Float64 render
    if (ctx->small_render && G == 1 && ctx->mode == RTW_MODE_FUSED && ctx->n_spheres64 <= rtw::kTileSpheres && max_depth >= 1 &&
        H > 0 && (long long)W * H * spp <= (1 << 17) &&
        (long long)W * H * spp * (long long)std::max<uint32_t>(ctx->n_spheres64, 8u) <= (1ll << 22)) {
        RTW_CUDA(ctx, cudaSetDevice(d0.device));
        rc = small_render_buffers(ctx, d0, img_vals * sizeof(double));
        if (rc) return rc;
        rtw::TraceParams64 p{};
        for (int k = 0; k < 3; ++k) {
            p.cam.origin[k] = cam->origin[k];
            p.cam.llc[k] = cam->lower_left_corner[k];
            p.cam.horizontal[k] = cam->horizontal[k];
            p.cam.vertical[k] = cam->vertical[k];
            p.cam.u[k] = cam->u[k];
            p.cam.v[k] = cam->v[k];
        }
        p.cam.lens_radius = cam->lens_radius;
        p.geom = d0.d_geom64; p.mat = d0.d_mat64; p.kind = d0.d_kind64;
        p.n_spheres = ctx->n_spheres64;
        p.W = W; p.H = H; p.spp = spp; p.max_depth = max_depth; p.sample_first = 0;
        p.key0 = (uint32_t)seed; p.key1 = (uint32_t)(seed >> 32);
        p.row_start = 0; p.row_stride = 1; p.n_rows = H;
        p.n_paths = (unsigned long long)W * H * spp;
        p.fx_scale = std::ldexp(1.0, fx_bits);
        p.counters = d0.d_counters;
        double* d_img = nullptr;
        unsigned long long* d_tot = nullptr;
        RTW_CUDA(ctx, cudaHostGetDevicePointer((void**)&d_img, d0.h_small_img, 0));
        RTW_CUDA(ctx, cudaHostGetDevicePointer((void**)&d_tot, d0.h_small_tot, 0));
        rtw::LaunchInfo li{};
        if (timing) RTW_CUDA(ctx, cudaEventRecord(d0.ev[0], d0.stream));
        RTW_CUDA(ctx, rtw::launch_small_render_f64(p, std::ldexp(1.0, -fx_bits), d_img, d_tot, d0.stream, &li));
        if (timing) RTW_CUDA(ctx, cudaEventRecord(d0.ev[1], d0.stream));
        RTW_CUDA(ctx, cudaStreamSynchronize(d0.stream));
        std::memcpy(out_rgb, d0.h_small_img, img_vals * sizeof(double));
        rtw_stats st = {};
        st.n_spheres = ctx->n_spheres64;
        st.image_width = W; st.image_height = H; st.rows_rendered = H;
        st.paths = p.n_paths;
        st.ray_segments = d0.h_small_tot[0];
        st.sphere_tests = st.ray_segments * (uint64_t)ctx->n_spheres64;
        st.kernel_launches = 1;
        st.n_devices = 1;
        if (timing) {
            float ms = 0.f;
            RTW_CUDA(ctx, cudaEventElapsedTime(&ms, d0.ev[0], d0.ev[1]));
            st.ms_trace = st.ms_total = ms;
        }
        d0.last = st;
        d0.last_valid = true;
        d0.last_stream = d0.stream;
        d0.last_resolved = true;
        d0.h_counters[1] = st.ray_segments;
        for (int i = 2; i < kCounters; ++i) d0.h_counters[i] = 0;
        if (stats) *stats = st;
        return RTW_OK;
    }
    RTW_CUDA(ctx, cudaSetDevice(d0.device));
    rc = grow(ctx, &d0.d_image64, &d0.image64_cap, img_vals);
    if (rc) return rc;
    if (!scene_uploaded_in_call) RTW_CUDA(ctx, cudaEventRecord(d0.ev[3], d0.stream));
    RTW_CUDA(ctx, cudaEventRecord(d0.ev[6], d0.stream));
    const int rows_pad = (H + G - 1) / G;
    const size_t tile_vals = (size_t)rows_pad * W * 3;
    if (G > 1) {
        rc = grow(ctx, &d0.d_gather64, &d0.gather64_cap, tile_vals * G);
        if (rc) return rc;
    }
    for (int g = 0; g < G; ++g) {
        DeviceState& ds = ctx->dev[g];
        RTW_CUDA(ctx, cudaSetDevice(ds.device));
        const int n_rows = rows_of(H, g, G);
        ds.last = rtw_stats{};
        ds.last.n_spheres = ctx->n_spheres64;
        ds.last.image_width = W;
        ds.last.image_height = H;
        ds.last.rows_rendered = n_rows;
        ds.last.paths = (uint64_t)n_rows * (uint64_t)W * (uint64_t)spp;
        ds.last_stream = ds.stream;
        ds.last_valid = true;
        ds.last_resolved = false;
        ds.h_counters[1] = 0;
        if (n_rows == 0) continue;
        const size_t npix = (size_t)n_rows * (size_t)W;
        rc = grow(ctx, &ds.d_accum, &ds.accum_cap, npix * 4);
        if (rc) return rc;
        if (timing) RTW_CUDA(ctx, cudaEventRecord(ds.ev[0], ds.stream));
        RTW_CUDA(ctx, cudaMemsetAsync(ds.d_accum, 0, npix * 4 * sizeof(unsigned long long), ds.stream));
        RTW_CUDA(ctx, cudaMemsetAsync(ds.d_counters, 0, kCounters * sizeof(unsigned long long), ds.stream));
        ds.counters_clean = false;
        int launches = 0;
        if (max_depth > 0) {
            rtw::TraceParams64 p;
            for (int k = 0; k < 3; ++k) {
                p.cam.origin[k] = cam->origin[k];
                p.cam.llc[k] = cam->lower_left_corner[k];
                p.cam.horizontal[k] = cam->horizontal[k];
                p.cam.vertical[k] = cam->vertical[k];
                p.cam.u[k] = cam->u[k];
                p.cam.v[k] = cam->v[k];
            }
            p.cam.lens_radius = cam->lens_radius;
            p.geom = ds.d_geom64;
            p.mat = ds.d_mat64;
            p.kind = ds.d_kind64;
            p.n_spheres = ctx->n_spheres64;
            p.W = W; p.H = H; p.spp = spp; p.max_depth = max_depth; p.sample_first = 0;
            p.key0 = (uint32_t)seed; p.key1 = (uint32_t)(seed >> 32);
            p.row_start = g; p.row_stride = G; p.n_rows = n_rows;
            p.n_paths = (unsigned long long)npix * (unsigned long long)spp;
            p.accum = ds.d_accum;
            p.fx_scale = std::ldexp(1.0, fx_bits);
            p.counters = ds.d_counters;
            rtw::LaunchInfo li{};
            RTW_CUDA(ctx, rtw::launch_trace_f64(p, ds.num_sms, ds.stream, &li));
            launches += li.launches;
        }
        if (timing) RTW_CUDA(ctx, cudaEventRecord(ds.ev[1], ds.stream));
        RTW_CUDA(ctx, cudaMemcpyAsync(ds.h_counters, ds.d_counters, kCounters * sizeof(unsigned long long),
                                      cudaMemcpyDeviceToHost, ds.stream));
        double* dst = G == 1 ? d0.d_image64 : d0.d_gather64 + tile_vals * g;
        double* tile = dst;
        if (g != 0) {
            rc = grow(ctx, &ds.d_tile64, &ds.tile64_cap, tile_vals);
            if (rc) return rc;
            tile = ds.d_tile64;
        }
        RTW_CUDA(ctx, rtw::launch_resolve_f64(ds.d_accum, W, H, n_rows, g, G, spp, std::ldexp(1.0, -fx_bits), G == 1 ? 1 : 0,
                                              tile, ds.stream));
        launches += 1;
        if (timing) RTW_CUDA(ctx, cudaEventRecord(ds.ev[2], ds.stream));
        ds.last_resolved = true;
        ds.last.kernel_launches = launches;
        if (g != 0) {
            const size_t bytes = (size_t)n_rows * W * 3 * sizeof(double);
            RTW_CUDA(ctx, cudaMemcpyPeerAsync(dst, d0.device, tile, ds.device, bytes, ds.stream));
            RTW_CUDA(ctx, cudaEventRecord(ds.ev_tile, ds.stream));
        }
    }
    RTW_CUDA(ctx, cudaSetDevice(d0.device));
    if (G > 1) {
        for (int g = 1; g < G; ++g) RTW_CUDA(ctx, cudaStreamWaitEvent(d0.stream, ctx->dev[g].ev_tile, 0));
        RTW_CUDA(ctx, rtw::launch_assemble_f64(d0.d_gather64, G, W, H, d0.d_image64, d0.stream));
    }
    RTW_CUDA(ctx, cudaEventRecord(d0.ev[4], d0.stream));
    if (img_vals)
        RTW_CUDA(ctx, cudaMemcpyAsync(out_rgb, d0.d_image64, img_vals * sizeof(double), cudaMemcpyDeviceToHost, d0.stream));
    RTW_CUDA(ctx, cudaEventRecord(d0.ev[5], d0.stream));
    for (int g = G - 1; g >= 0; --g) {
        RTW_CUDA(ctx, cudaSetDevice(ctx->dev[g].device));
        RTW_CUDA(ctx, cudaStreamSynchronize(ctx->dev[g].stream));
    }
    rtw_stats total = {};
    total.n_spheres = ctx->n_spheres64;
    total.image_width = W;
    total.image_height = H;
    for (int g = 0; g < G; ++g) {
        DeviceState& ds = ctx->dev[g];
        RTW_CUDA(ctx, cudaSetDevice(ds.device));
        ds.last.ray_segments = ds.h_counters[1];
        ds.last.sphere_tests = ds.last.ray_segments * (uint64_t)ctx->n_spheres64;
        if (timing && ds.last.rows_rendered > 0) {
            float a = 0.f, b = 0.f;
            RTW_CUDA(ctx, cudaEventElapsedTime(&a, ds.ev[0], ds.ev[1]));
            RTW_CUDA(ctx, cudaEventElapsedTime(&b, ds.ev[1], ds.ev[2]));
            ds.last.ms_trace = a;
            ds.last.ms_resolve = b;
        }
        total.paths += ds.last.paths;
        total.ray_segments += ds.last.ray_segments;
        total.sphere_tests += ds.last.sphere_tests;
        total.rows_rendered += ds.last.rows_rendered;
        total.kernel_launches += ds.last.kernel_launches;
        total.ms_trace = std::fmax(total.ms_trace, ds.last.ms_trace);
        total.ms_resolve = std::fmax(total.ms_resolve, ds.last.ms_resolve);
    }
    if (G > 1) total.kernel_launches += 1;
    total.n_devices = G;
    RTW_CUDA(ctx, cudaSetDevice(d0.device));
    float ms = 0.f;
    RTW_CUDA(ctx, cudaEventElapsedTime(&ms, d0.ev[3], d0.ev[5]));
    total.ms_total = ms;
    RTW_CUDA(ctx, cudaEventElapsedTime(&ms, d0.ev[4], d0.ev[5]));
    total.ms_d2h = ms;
    RTW_CUDA(ctx, cudaEventElapsedTime(&ms, d0.ev[3], d0.ev[6]));
    total.ms_h2d = ms;
    if (stats) *stats = total;
    return RTW_OK;
}

}  // namespace

// ================================================================================================ C-ABI

extern "C" {

int rtw_abi_version(void) { return RTW_ABI_VERSION; }

int rtw_has_variants(void) { return kHaveVariants ? 1 : 0; }

int rtw_image_height(int image_width) {
    if (image_width < 0) return 0;
    return (int)(((long long)image_width * 9) / 16);  // image_width div (16//9), src/render.jl:11-12
}

int rtw_device_count(int* count) {
    if (!count) return RTW_E_INVALID_ARG;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        *count = 0;
        return RTW_E_NO_DEVICE;
    }
    *count = n;
    return RTW_OK;
}

int rtw_create(const int* device_ids, int n_devices, rtw_ctx** out_ctx) {
    if (!out_ctx) return RTW_E_INVALID_ARG;
    *out_ctx = nullptr;
    int visible = 0;
    int rc = rtw_device_count(&visible);
    if (rc) return rc;
    if (n_devices <= 0) n_devices = 1;
    if (n_devices > visible) return RTW_E_INVALID_ARG;
    rtw_ctx* ctx = new (std::nothrow) rtw_ctx();
    if (!ctx) return RTW_E_INTERNAL;
    try {
        ctx->dev.resize((size_t)n_devices);
    } catch (...) {
        delete ctx;
        return RTW_E_INTERNAL;
    }
    for (int g = 0; g < n_devices; ++g) {
        DeviceState& ds = ctx->dev[g];
        ds.device = device_ids ? device_ids[g] : g;
        if (ds.device < 0 || ds.device >= visible) {
            rtw_destroy(ctx);
            return RTW_E_INVALID_ARG;
        }
        cudaError_t e = cudaSetDevice(ds.device);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ds.num_sms, cudaDevAttrMultiProcessorCount, ds.device);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ds.stream, cudaStreamNonBlocking);
        for (int i = 0; i < 8 && e == cudaSuccess; ++i) e = cudaEventCreate(&ds.ev[i]);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ds.ev_tile, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaMalloc((void**)&ds.d_counters, kCounters * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMallocHost((void**)&ds.h_counters, 2 * kCounters * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMalloc((void**)&ds.d_scratch, 4u << 20);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            rtw_destroy(ctx);
            return (int)e;
        }
        for (int i = 0; i < 2 * kCounters; ++i) ds.h_counters[i] = 0;
    }
    // enable peer access towards device 0 for the tile gather (ignored when unavailable: the copy is then staged)
    for (int g = 1; g < n_devices; ++g) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, ctx->dev[g].device, ctx->dev[0].device) == cudaSuccess && can) {
            cudaSetDevice(ctx->dev[g].device);
            cudaError_t e = cudaDeviceEnablePeerAccess(ctx->dev[0].device, 0);
            if (e != cudaSuccess) (void)cudaGetLastError();
        }
    }
    *out_ctx = ctx;
    return RTW_OK;
}

int rtw_destroy(rtw_ctx* ctx) {
    if (!ctx) return RTW_OK;
    for (size_t g = 0; g < ctx->nccl_comm.size(); ++g) {
        if (!ctx->nccl_comm[g] || g >= ctx->dev.size()) continue;
        if (cudaSetDevice(ctx->dev[g].device) == cudaSuccess) {
            cudaStreamSynchronize(ctx->dev[g].stream);
            ctx->nccl_CommDestroy(ctx->nccl_comm[g]);
        }
    }
    ctx->nccl_comm.clear();
    for (auto& ds : ctx->dev) {
        if (cudaSetDevice(ds.device) != cudaSuccess) continue;
        if (ds.stream) cudaStreamSynchronize(ds.stream);
        cudaFree(ds.d_geom); cudaFree(ds.d_geom_pairs); cudaFree(ds.d_mat); cudaFree(ds.d_kind);
        cudaFree(ds.d_geom_perm[0]); cudaFree(ds.d_geom_perm[1]); cudaFree(ds.d_uv);
        cudaFree(ds.d_geom64); cudaFree(ds.d_mat64); cudaFree(ds.d_kind64);
        cudaFree(ds.d_tile64); cudaFree(ds.d_gather64); cudaFree(ds.d_image64);
        cudaFree(ds.d_grid_big); cudaFree(ds.d_cull_start); cudaFree(ds.d_cull_items);
        for (int l = 0; l < 2; ++l) { cudaFree(ds.d_grid_start[l]); cudaFree(ds.d_grid_items[l]); }
        cudaFree(ds.d_accum); cudaFree(ds.d_counters); cudaFree(ds.d_tile);
        cudaFree(ds.d_gather); cudaFree(ds.d_image); cudaFree(ds.d_rgb8); cudaFree(ds.d_scratch); cudaFree(ds.d_wf);
        if (ds.h_counters) cudaFreeHost(ds.h_counters);
        cudaFree(ds.d_gen_ws);
        if (ds.h_small_img) cudaFreeHost(ds.h_small_img);
        if (ds.h_small_tot) cudaFreeHost(ds.h_small_tot);
        for (auto& e : ds.ev) if (e) cudaEventDestroy(e);
        if (ds.ev_tile) cudaEventDestroy(ds.ev_tile);
        if (ds.stream) cudaStreamDestroy(ds.stream);
    }
    (void)cudaGetLastError();
    delete ctx;
    return RTW_OK;
}

const char* rtw_last_error(const rtw_ctx* ctx) { return ctx ? ctx->err.c_str() : "rtw_ctx is NULL"; }

int rtw_set_option(rtw_ctx* ctx, int option, int64_t value) {
    if (!ctx) return RTW_E_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    switch (option) {
        case RTW_OPT_MODE:
            if (value != RTW_MODE_FUSED && value != RTW_MODE_WAVEFRONT && value != RTW_MODE_CTA_WAVEFRONT && value != RTW_MODE_GRID)
                return fail(ctx, RTW_E_INVALID_ARG, "unknown mode");
            if (value == RTW_MODE_CTA_WAVEFRONT && !kHaveVariants)
                return fail(ctx, RTW_E_UNSUPPORTED, "RTW_MODE_CTA_WAVEFRONT needs a library built with RTW_BUILD_VARIANTS=1");
            ctx->mode = (int)value;
            return RTW_OK;
        case RTW_OPT_STRIP:
            return fail(ctx, RTW_E_UNSUPPORTED, "RTW_OPT_STRIP was removed in ABI v3");
        case RTW_OPT_RAYS_PER_LANE:
            if (value != 0 && value != 1 && value != 2 && value != 4)
                return fail(ctx, RTW_E_INVALID_ARG, "rays_per_lane must be 0 (default), 1, 2 or 4");
            ctx->rays_per_lane = (int)value;
            return RTW_OK;
        case RTW_OPT_COOP:
            if (value != 0 && value != 1 && value != 2 && value != 4)
                return fail(ctx, RTW_E_INVALID_ARG, "coop must be 0 (default), 1, 2 or 4");
            ctx->coop = (int)value;
            return RTW_OK;
        case RTW_OPT_SWEEP:
            if (value < 0 || value > RTW_SWEEP_PACKED) return fail(ctx, RTW_E_INVALID_ARG, "unknown sweep variant");
            ctx->sweep = (int)value;
            return RTW_OK;
        case RTW_OPT_TAIL:
            if (value != RTW_TAIL_DEFAULT && value != RTW_TAIL_SPLIT && value != RTW_TAIL_UNIFIED)
                return fail(ctx, RTW_E_INVALID_ARG, "unknown tail variant");
            ctx->tail = (int)value;
            return RTW_OK;
        case RTW_OPT_WALK:
            if (value != RTW_WALK_DEFAULT && value != RTW_WALK_SLOTS && value != RTW_WALK_OWN_RAY)
                return fail(ctx, RTW_E_INVALID_ARG, "unknown walk variant");
            ctx->walk = (int)value;
            return RTW_OK;
        case RTW_OPT_BLOCKS_PER_SM:
            if (value < 0 || value > 32) return fail(ctx, RTW_E_INVALID_ARG, "blocks_per_sm must be in 0..32");
            ctx->blocks_per_sm = (int)value;
            return RTW_OK;
        case RTW_OPT_COLLECT_TIMING:
            ctx->collect_timing = value != 0;
            return RTW_OK;
        case RTW_OPT_SMALL_RENDER:
            ctx->small_render = value != 0;
            return RTW_OK;
        case RTW_OPT_GATHER:
            if (value != RTW_GATHER_PEER && value != RTW_GATHER_NCCL) return fail(ctx, RTW_E_INVALID_ARG, "unknown gather");
            ctx->gather = (int)value;
            return RTW_OK;
        default:
            return fail(ctx, RTW_E_UNSUPPORTED, "unknown option");
    }
}

int rtw_set_scene(rtw_ctx* ctx, const float* geom4, const float* mat4, const uint32_t* kind, uint32_t n_spheres) {
    if (!ctx) return RTW_E_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    try {
        return set_scene_locked(ctx, geom4, mat4, kind, n_spheres);
    } catch (...) {
        return fail(ctx, RTW_E_INTERNAL, "unexpected C++ exception in rtw_set_scene");
    }
}

int rtw_scene_random_spheres(rtw_ctx* ctx, uint64_t rng_state[2], int half_extent, int install, float* geom4, float* mat4,
                             uint32_t* kind, uint32_t capacity, uint32_t* n_spheres) {
    if (!ctx) return RTW_E_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    try {
        if (!rng_state || !n_spheres) return fail(ctx, RTW_E_INVALID_ARG, "rng_state / n_spheres is NULL");
        if (half_extent < 1 || half_extent > 512) return fail(ctx, RTW_E_INVALID_ARG, "half_extent must be in 1..512");
        DeviceState& ds = ctx->dev[0];
        RTW_CUDA(ctx, cudaSetDevice(ds.device));
        const size_t cap = rtw::scenegen_max_spheres(half_extent);
        const size_t ws = rtw::scenegen_workspace_bytes(half_extent);
        const size_t list_bytes = cap * (16 + 16 + 4) + 1024;
        int rc = grow(ctx, &ds.d_gen_ws, &ds.gen_ws_cap, ws + list_bytes);
        if (rc) return rc;
        float4* d_g = (float4*)(ds.d_gen_ws + ws);
        float4* d_m = d_g + cap;
        uint32_t* d_k = (uint32_t*)(d_m + cap);
        unsigned long long* h_out = ds.h_counters + kCounters;  // pinned scratch
        RTW_CUDA(ctx, rtw::launch_scenegen(rng_state[0], rng_state[1], half_extent, ds.d_gen_ws, d_g, d_m, d_k, h_out, ds.stream));
        RTW_CUDA(ctx, cudaStreamSynchronize(ds.stream));
        const uint32_t n = (uint32_t)(h_out[2] & 0xffffffffull);
        *n_spheres = n;
        if (n > cap) return fail(ctx, RTW_E_INTERNAL, "scene generator produced more spheres than its bound");
        const bool want_arrays = geom4 || mat4 || kind;
        if (want_arrays && (!geom4 || !mat4 || !kind || capacity < n))
            return fail(ctx, RTW_E_INVALID_ARG, "geom4 / mat4 / kind must all be given with capacity >= n_spheres");
        std::vector<float> hg, hm;
        std::vector<uint32_t> hk;
        float* pg = geom4;
        float* pm = mat4;
        uint32_t* pk = kind;
        if (!want_arrays) {
            hg.resize(4 * (size_t)n); hm.resize(4 * (size_t)n); hk.resize(n);
            pg = hg.data(); pm = hm.data(); pk = hk.data();
        }
        if (want_arrays || install) {
            RTW_CUDA(ctx, cudaMemcpy(pg, d_g, 16 * (size_t)n, cudaMemcpyDeviceToHost));
            RTW_CUDA(ctx, cudaMemcpy(pm, d_m, 16 * (size_t)n, cudaMemcpyDeviceToHost));
            RTW_CUDA(ctx, cudaMemcpy(pk, d_k, 4 * (size_t)n, cudaMemcpyDeviceToHost));
        }
        rng_state[0] = h_out[0];
        rng_state[1] = h_out[1];
        if (install) return set_scene_locked(ctx, pg, pm, pk, n);
        return RTW_OK;
    } catch (...) {
        return fail(ctx, RTW_E_INTERNAL, "unexpected C++ exception in rtw_scene_random_spheres");
    }
}

int rtw_render(rtw_ctx* ctx, const rtw_camera* cam, int image_width, int n_samples, int max_depth, uint64_t seed,
               float* out_rgb, rtw_stats* stats) {
    if (!ctx) return RTW_E_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    try {
        return render_locked(ctx, cam, image_width, n_samples, max_depth, seed, out_rgb, stats, false);
    } catch (...) {
        return fail(ctx, RTW_E_INTERNAL, "unexpected C++ exception in rtw_render");
    }
}

int rtw_render_scene(rtw_ctx* ctx, const float* geom4, const float* mat4, const uint32_t* kind, uint32_t n_spheres,
                     const rtw_camera* cam, int image_width, int n_samples, int max_depth, uint64_t seed,
                     float* out_rgb, rtw_stats* stats) {
    if (!ctx) return RTW_E_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    try {
        DeviceState& d0 = ctx->dev[0];
        RTW_CUDA(ctx, cudaSetDevice(d0.device));
        RTW_CUDA(ctx, cudaEventRecord(d0.ev[3], d0.stream));
        int rc = set_scene_locked(ctx, geom4, mat4, kind, n_spheres);
        if (rc) return rc;
        return render_locked(ctx, cam, image_width, n_samples, max_depth, seed, out_rgb, stats, true);
    } catch (...) {
        return fail(ctx, RTW_E_INTERNAL, "unexpected C++ exception in rtw_render_scene");
    }
}

int rtw_render_rows_device(rtw_ctx* ctx, int device_slot, const rtw_camera* cam, int image_width, int n_samples,
                           int max_depth, uint64_t seed, int row_start, int row_stride, int column_major,
                           float* d_tile, void* stream) {
    if (!ctx) return RTW_E_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    try {
        if (device_slot < 0 || device_slot >= (int)ctx->dev.size()) return fail(ctx, RTW_E_INVALID_ARG, "bad device_slot");
        int rc = check_render_args(ctx, cam, image_width, n_samples, max_depth);
        if (rc) return rc;
        if (!d_tile) return fail(ctx, RTW_E_INVALID_ARG, "d_tile is NULL");
        if (row_start < 0 || row_stride < 1) return fail(ctx, RTW_E_INVALID_ARG, "bad row_start/row_stride");
        if (column_major && (row_start != 0 || row_stride != 1))
            return fail(ctx, RTW_E_INVALID_ARG, "column_major output needs row_start=0,row_stride=1");
        DeviceState& ds = ctx->dev[device_slot];
        cudaStream_t s = stream ? (cudaStream_t)stream : ds.stream;
        ctx->prog = ProgressiveState{};  // this device's accumulator is reused
        return enqueue_rows(ctx, ds, cam, image_width, n_samples, max_depth, seed, row_start, row_stride, column_major,
                            d_tile, s, ctx->collect_timing != 0);
    } catch (...) {
        return fail(ctx, RTW_E_INTERNAL, "unexpected C++ exception in rtw_render_rows_device");
    }
}

int rtw_last_stats(rtw_ctx* ctx, int device_slot, rtw_stats* stats) {
    if (!ctx || !stats) return RTW_E_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (device_slot < 0 || device_slot >= (int)ctx->dev.size()) return fail(ctx, RTW_E_INVALID_ARG, "bad device_slot");
    DeviceState& ds = ctx->dev[device_slot];
    if (!ds.last_valid) return fail(ctx, RTW_E_INVALID_ARG, "no render has been enqueued on this device");
    RTW_CUDA(ctx, cudaSetDevice(ds.device));
    RTW_CUDA(ctx, cudaStreamSynchronize(ds.last_stream ? ds.last_stream : ds.stream));
    int rc = finish_stats(ctx, ds, ctx->collect_timing != 0);
    if (rc) return rc;
    *stats = ds.last;
    return RTW_OK;
}

int rtw_assemble_tiles_device(rtw_ctx* ctx, int device_slot, const float* d_tiles, int n_tiles, int image_width,
                              float* d_out_rgb, void* stream) {
    if (!ctx) return RTW_E_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (device_slot < 0 || device_slot >= (int)ctx->dev.size()) return fail(ctx, RTW_E_INVALID_ARG, "bad device_slot");
    if (!d_tiles || !d_out_rgb || n_tiles < 1 || image_width < 1) return fail(ctx, RTW_E_INVALID_ARG, "bad arguments");
    DeviceState& ds = ctx->dev[device_slot];
    RTW_CUDA(ctx, cudaSetDevice(ds.device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ds.stream;
    RTW_CUDA(ctx, rtw::launch_assemble(d_tiles, n_tiles, image_width, rtw_image_height(image_width), d_out_rgb, s));
    return RTW_OK;
}

int rtw_set_scene_f64(rtw_ctx* ctx, const double* geom4, const double* mat4, const uint32_t* kind, uint32_t n_spheres) {
    if (!ctx) return RTW_E_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    try {
        return set_scene_f64_locked(ctx, geom4, mat4, kind, n_spheres);
    } catch (...) {
        return fail(ctx, RTW_E_INTERNAL, "unexpected C++ exception in rtw_set_scene_f64");
    }
}

int rtw_render_f64(rtw_ctx* ctx, const rtw_camera_f64* cam, int image_width, int n_samples, int max_depth, uint64_t seed,
                   double* out_rgb, rtw_stats* stats) {
    if (!ctx) return RTW_E_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    try {
        return render_f64_locked(ctx, cam, image_width, n_samples, max_depth, seed, out_rgb, stats, false);
    } catch (...) {
        return fail(ctx, RTW_E_INTERNAL, "unexpected C++ exception in rtw_render_f64");
    }
}

int rtw_render_scene_f64(rtw_ctx* ctx, const double* geom4, const double* mat4, const uint32_t* kind, uint32_t n_spheres,
                         const rtw_camera_f64* cam, int image_width, int n_samples, int max_depth, uint64_t seed,
                         double* out_rgb, rtw_stats* stats) {
    if (!ctx) return RTW_E_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    try {
        DeviceState& d0 = ctx->dev[0];
        RTW_CUDA(ctx, cudaSetDevice(d0.device));
        RTW_CUDA(ctx, cudaEventRecord(d0.ev[3], d0.stream));
        int rc = set_scene_f64_locked(ctx, geom4, mat4, kind, n_spheres);
        if (rc) return rc;
        return render_f64_locked(ctx, cam, image_width, n_samples, max_depth, seed, out_rgb, stats, true);
    } catch (...) {
        return fail(ctx, RTW_E_INTERNAL, "unexpected C++ exception in rtw_render_scene_f64");
    }
}

int rtw_accumulate(rtw_ctx* ctx, const rtw_camera* cam, int image_width, int sample_first, int sample_count,
                   int n_samples_total, int max_depth, uint64_t seed, rtw_stats* stats) {
    if (!ctx) return RTW_E_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    try {
        int rc = check_render_args(ctx, cam, image_width, n_samples_total, max_depth);
        if (rc) return rc;
        if (sample_first < 0 || sample_count < 1 || sample_first > n_samples_total - sample_count)
            return fail(ctx, RTW_E_INVALID_ARG, "samples must satisfy 0 <= first, 1 <= count, first + count <= total");
        ProgressiveState& pg = ctx->prog;
        const uint64_t cam_hash = fnv1a(cam, sizeof(rtw_camera), 1469598103934665603ull);  // the 88 bytes of the camera
        if (sample_first != 0) {
            if (!pg.valid || pg.W != image_width || pg.s_total != n_samples_total || pg.s_done != sample_first)
                return fail(ctx, RTW_E_INVALID_ARG, "sample_first must continue the progressive image held by the context "
                                                    "(same width and total, first == samples accumulated so far)");
            if (pg.have_inputs && (pg.seed != seed || pg.max_depth != max_depth || pg.cam_hash != cam_hash))
                return fail(ctx, RTW_E_INVALID_ARG, "a continuation pass must use the seed, max_depth and camera the "
                                                    "progressive image was started with");
        }
        const bool had_inputs = sample_first != 0 ? pg.have_inputs : true;
        const PassSpec ps{sample_first, sample_count, n_samples_total, sample_first == 0};
        pg.valid = false;  // stays invalid if the pass fails half-way
        rc = pass_locked(ctx, cam, image_width, max_depth, seed, ps, true, false, 0, nullptr, nullptr, stats, false);
        if (rc) return rc;
        pg.valid = true;
        pg.W = image_width;
        pg.s_total = n_samples_total;
        pg.s_done = sample_first + sample_count;
        pg.have_inputs = had_inputs;
        if (sample_first == 0 || !had_inputs) {
            pg.seed = seed;
            pg.max_depth = max_depth;
            pg.cam_hash = cam_hash;
            pg.have_inputs = true;  // from here on the inputs are pinned (also after a raw-checkpoint resume)
        }
        return RTW_OK;
    } catch (...) {
        return fail(ctx, RTW_E_INTERNAL, "unexpected C++ exception in rtw_accumulate");
    }
}

static int resolve_common(rtw_ctx* ctx, float* out_rgb, uint8_t* out_rgb8) {
    if (!ctx) return RTW_E_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    try {
        if (!out_rgb && !out_rgb8) return fail(ctx, RTW_E_INVALID_ARG, "output buffer is NULL");
        const ProgressiveState& pg = ctx->prog;
        if (!pg.valid || pg.s_done < 1) return fail(ctx, RTW_E_INVALID_ARG, "no progressive image: call rtw_accumulate first");
        const PassSpec ps{0, 0, pg.s_total, false};
        return pass_locked(ctx, nullptr, pg.W, 0, 0, ps, false, true, pg.s_done, out_rgb, out_rgb8, nullptr, false);
    } catch (...) {
        return fail(ctx, RTW_E_INTERNAL, "unexpected C++ exception in rtw_resolve");
    }
}

int rtw_resolve(rtw_ctx* ctx, float* out_rgb) { return resolve_common(ctx, out_rgb, nullptr); }

int rtw_resolve_rgb8(rtw_ctx* ctx, uint8_t* out_rgb8) { return resolve_common(ctx, nullptr, out_rgb8); }

int rtw_progress(rtw_ctx* ctx, int* image_width, int* samples_done, int* samples_total) {
    if (!ctx) return RTW_E_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    const ProgressiveState& pg = ctx->prog;
    if (image_width) *image_width = pg.valid ? pg.W : 0;
    if (samples_done) *samples_done = pg.valid ? pg.s_done : 0;
    if (samples_total) *samples_total = pg.valid ? pg.s_total : 0;
    return RTW_OK;
}

// rows r = g, g + G, ... of the image live, compacted, in device g's accumulator
static int accumulator_copy(rtw_ctx* ctx, int W, int64_t* host_out, const int64_t* host_in) {
    const int H = rtw_image_height(W);
    const int G = (int)ctx->dev.size();
    std::vector<int64_t> tmp;
    for (int g = 0; g < G; ++g) {
        DeviceState& ds = ctx->dev[g];
        const int n_rows = rows_of(H, g, G);
        const size_t vals = (size_t)n_rows * (size_t)W * 4;
        if (vals == 0) continue;
        RTW_CUDA(ctx, cudaSetDevice(ds.device));
        tmp.resize(vals);
        if (host_in) {
            for (int k = 0; k < n_rows; ++k)
                std::memcpy(tmp.data() + (size_t)k * W * 4, host_in + (size_t)(g + k * G) * W * 4, (size_t)W * 4 * sizeof(int64_t));
            int rc = grow(ctx, &ds.d_accum, &ds.accum_cap, vals);
            if (rc) return rc;
            RTW_CUDA(ctx, cudaMemcpyAsync(ds.d_accum, tmp.data(), vals * sizeof(int64_t), cudaMemcpyHostToDevice, ds.stream));
            RTW_CUDA(ctx, cudaStreamSynchronize(ds.stream));
        } else {
            if (vals > ds.accum_cap) return fail(ctx, RTW_E_INTERNAL, "accumulator smaller than the progressive image");
            RTW_CUDA(ctx, cudaMemcpyAsync(tmp.data(), ds.d_accum, vals * sizeof(int64_t), cudaMemcpyDeviceToHost, ds.stream));
            RTW_CUDA(ctx, cudaStreamSynchronize(ds.stream));
            for (int k = 0; k < n_rows; ++k)
                std::memcpy(host_out + (size_t)(g + k * G) * W * 4, tmp.data() + (size_t)k * W * 4, (size_t)W * 4 * sizeof(int64_t));
        }
    }
    return RTW_OK;
}

int rtw_accumulator_read(rtw_ctx* ctx, int64_t* out, uint64_t n_values) {
    if (!ctx) return RTW_E_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    try {
        const ProgressiveState& pg = ctx->prog;
        if (!pg.valid) return fail(ctx, RTW_E_INVALID_ARG, "no progressive image: call rtw_accumulate first");
        const uint64_t need = (uint64_t)pg.W * (uint64_t)rtw_image_height(pg.W) * 4ull;
        if (!out || n_values != need) return fail(ctx, RTW_E_INVALID_ARG, "out must hold H*W*4 int64 values");
        return accumulator_copy(ctx, pg.W, out, nullptr);
    } catch (...) {
        return fail(ctx, RTW_E_INTERNAL, "unexpected C++ exception in rtw_accumulator_read");
    }
}

int rtw_accumulator_write(rtw_ctx* ctx, const int64_t* in, uint64_t n_values, int image_width, int samples_done,
                          int samples_total) {
    if (!ctx) return RTW_E_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    try {
        if (image_width < 1 || image_width > 65536 || samples_total < 1 || samples_total > (1 << 24) || samples_done < 0 ||
            samples_done > samples_total)
            return fail(ctx, RTW_E_INVALID_ARG, "bad image_width / samples_done / samples_total");
        const uint64_t need = (uint64_t)image_width * (uint64_t)rtw_image_height(image_width) * 4ull;
        if (!in || n_values != need) return fail(ctx, RTW_E_INVALID_ARG, "in must hold H*W*4 int64 values");
        ctx->prog = ProgressiveState{};
        int rc = accumulator_copy(ctx, image_width, nullptr, in);
        if (rc) return rc;
        ctx->prog.valid = true;
        ctx->prog.W = image_width;
        ctx->prog.s_total = samples_total;
        ctx->prog.s_done = samples_done;
        return RTW_OK;
    } catch (...) {
        return fail(ctx, RTW_E_INTERNAL, "unexpected C++ exception in rtw_accumulator_write");
    }
}

/* checkpoint file, little endian ("RTWCKPT2": the sums of the Philox4x32-7 stream; "RTWCKPT1" files hold 10-round sums and are
 * refused -- continuing them would mix two streams):
 *    0 char[8] "RTWCKPT2"      8 i32 W   12 i32 H   16 i32 samples_done   20 i32 samples_total   24 i32 fx_bits
 *   28 i32 max_depth   32 u32 have_inputs   36 u32 n_spheres   40 u64 seed   48 u64 camera hash   56 u64 scene hash
 *   64 i64[H*W*4] fixed-point sums, row-major [row][col][r,g,b,unused]      then u32 CRC-32 of all preceding bytes */
struct CkptHeader {
    char magic[8];
    int32_t W, H, s_done, s_total, fx_bits, max_depth;
    uint32_t have_inputs, n_spheres;
    uint64_t seed, cam_hash, scene_hash;
};
static_assert(sizeof(CkptHeader) == 64, "checkpoint header layout");

int rtw_checkpoint_save(rtw_ctx* ctx, const char* path) {
    if (!ctx) return RTW_E_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    try {
        const ProgressiveState& pg = ctx->prog;
        if (!path) return fail(ctx, RTW_E_INVALID_ARG, "path is NULL");
        if (!pg.valid) return fail(ctx, RTW_E_INVALID_ARG, "no progressive image: call rtw_accumulate first");
        CkptHeader h{};
        std::memcpy(h.magic, "RTWCKPT2", 8);
        h.W = pg.W; h.H = rtw_image_height(pg.W); h.s_done = pg.s_done; h.s_total = pg.s_total;
        h.fx_bits = fx_bits_for(pg.s_total);
        h.max_depth = pg.max_depth; h.have_inputs = pg.have_inputs ? 1u : 0u; h.n_spheres = ctx->n_spheres;
        h.seed = pg.seed; h.cam_hash = pg.cam_hash; h.scene_hash = ctx->scene_hash;
        std::vector<int64_t> acc((size_t)h.W * (size_t)h.H * 4);
        int rc = accumulator_copy(ctx, pg.W, acc.data(), nullptr);
        if (rc) return rc;
        uint32_t crc = crc32_of(&h, sizeof h, 0u);
        crc = crc32_of(acc.data(), acc.size() * sizeof(int64_t), crc);
        FILE* f = fopen(path, "wb");
        if (!f) return fail(ctx, RTW_E_IO, "cannot open the checkpoint file for writing");
        bool ok = fwrite(&h, sizeof h, 1, f) == 1;
        ok = ok && (acc.empty() || fwrite(acc.data(), sizeof(int64_t), acc.size(), f) == acc.size());
        ok = ok && fwrite(&crc, 4, 1, f) == 1;
        ok = (fclose(f) == 0) && ok;
        return ok ? RTW_OK : fail(ctx, RTW_E_IO, "short write to the checkpoint file");
    } catch (...) {
        return fail(ctx, RTW_E_INTERNAL, "unexpected C++ exception in rtw_checkpoint_save");
    }
}

int rtw_checkpoint_load(rtw_ctx* ctx, const char* path) {
    if (!ctx) return RTW_E_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    try {
        if (!path) return fail(ctx, RTW_E_INVALID_ARG, "path is NULL");
        if (!ctx->have_scene) return fail(ctx, RTW_E_NO_SCENE, "set the scene the checkpoint was rendered from first");
        FILE* f = fopen(path, "rb");
        if (!f) return fail(ctx, RTW_E_IO, "cannot open the checkpoint file");
        CkptHeader h{};
        bool ok = fread(&h, sizeof h, 1, f) == 1 && std::memcmp(h.magic, "RTWCKPT2", 8) == 0;
        ok = ok && h.W >= 1 && h.W <= 65536 && h.H == rtw_image_height(h.W) && h.s_total >= 1 && h.s_total <= (1 << 24) &&
             h.s_done >= 0 && h.s_done <= h.s_total && h.fx_bits == fx_bits_for(h.s_total);
        std::vector<int64_t> acc;
        uint32_t crc_file = 0;
        if (ok) {
            acc.resize((size_t)h.W * (size_t)h.H * 4);
            ok = (acc.empty() || fread(acc.data(), sizeof(int64_t), acc.size(), f) == acc.size()) && fread(&crc_file, 4, 1, f) == 1;
        }
        fclose(f);
        if (!ok) return fail(ctx, RTW_E_FORMAT, "not a checkpoint file of this library (magic, sizes or fixed-point scale)");
        uint32_t crc = crc32_of(&h, sizeof h, 0u);
        crc = crc32_of(acc.data(), acc.size() * sizeof(int64_t), crc);
        if (crc != crc_file) return fail(ctx, RTW_E_FORMAT, "checkpoint checksum does not match");
        if (h.scene_hash != ctx->scene_hash || h.n_spheres != ctx->n_spheres)
            return fail(ctx, RTW_E_INVALID_ARG, "the checkpoint belongs to another scene");
        ctx->prog = ProgressiveState{};
        int rc = accumulator_copy(ctx, h.W, nullptr, acc.data());
        if (rc) return rc;
        ProgressiveState& pg = ctx->prog;
        pg.valid = true; pg.W = h.W; pg.s_total = h.s_total; pg.s_done = h.s_done;
        pg.max_depth = h.max_depth; pg.seed = h.seed; pg.cam_hash = h.cam_hash; pg.have_inputs = h.have_inputs != 0;
        return RTW_OK;
    } catch (...) {
        return fail(ctx, RTW_E_INTERNAL, "unexpected C++ exception in rtw_checkpoint_load");
    }
}

int rtw_measure_fp32_peak(rtw_ctx* ctx, int device_slot, int variant, double* fp32_instr_per_s, float* ms_out) {
    if (!ctx || !fp32_instr_per_s) return RTW_E_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (device_slot < 0 || device_slot >= (int)ctx->dev.size()) return fail(ctx, RTW_E_INVALID_ARG, "bad device_slot");
    if (variant < 0 || variant / 10 > 8) return fail(ctx, RTW_E_INVALID_ARG, "variant must be 0..9 (+10*L)");
    DeviceState& ds = ctx->dev[device_slot];
    RTW_CUDA(ctx, cudaSetDevice(ds.device));
    double instr = 0.0, best = 0.0;
    float best_ms = 0.f;
    const float ray[6] = {0.125f, 0.25f, 0.5f, 0.6f, 0.0f, 0.8f};  // origin, unit direction
    RTW_CUDA(ctx, cudaMemcpyAsync(ds.d_scratch, ray, sizeof ray, cudaMemcpyHostToDevice, ds.stream));
    for (int rep = 0; rep < 5; ++rep) {  // rep 0 is the warm-up
        RTW_CUDA(ctx, cudaEventRecord(ds.ev[0], ds.stream));
        RTW_CUDA(ctx, rtw::launch_fp32_peak(variant, ds.num_sms, ds.d_scratch, ds.stream, &instr));
        RTW_CUDA(ctx, cudaEventRecord(ds.ev[1], ds.stream));
        RTW_CUDA(ctx, cudaStreamSynchronize(ds.stream));
        float ms = 0.f;
        RTW_CUDA(ctx, cudaEventElapsedTime(&ms, ds.ev[0], ds.ev[1]));
        if (rep > 0 && ms > 0.f) {
            double r = instr / (ms * 1e-3);
            if (r > best) { best = r; best_ms = ms; }
        }
    }
    *fp32_instr_per_s = best;
    if (ms_out) *ms_out = best_ms;
    return RTW_OK;
}

}  // extern "C"
