// rtw_kernels.cu -- hand-written sm_100a kernels of the render hot path.
//
//   fused_trace_kernel : persistent-threads wavefront.  Every lane owns one path; the CTA-resident loop is
//                        regenerate (raygen, src/render.jl:26-37 + src/camera.jl:43-48)
//                        -> intersect (closest-hit sweep over the sphere list, src/hit.jl:38-50, 12-35)
//                        -> shade/scatter (src/material.jl, src/light.jl, src/ray_color.jl:1-6, 14-37)
//                        -> accumulate (src/render.jl:38).  Finished lanes are refilled from a ticket counter
//                        (warp ballot + prefix), so the sweep always runs on dense warps.
//   resolve_kernel     : accum/n_samples, gamma-2, store in Julia column-major (src/render.jl:40, src/vec.jl:22)
//   assemble_kernel    : un-interleave gathered row tiles (multi-GPU)
//   fp32_peak_kernel   : roofline denominators (FP32 issue rate)
//
// Compiled with -fmad=false; see rtw_device.cuh for the FP contract.
#include "rtw_kernels.h"
#include "rtw_sweep.cuh"

namespace rtw {

namespace {

#ifdef RTW_BUILD_VARIANTS  // the first fused kernel family (RTW_TAIL_SPLIT): measured comparison, not shipped by default
// ---- the persistent fused kernel ------------------------------------------------------------------------------
// kMulti = false: the whole list (<= kTileSpheres) is staged once; warps then run free of CTA barriers.
// kMulti = true : the list is streamed per bounce through two 16 KB TMA buffers, CTA-synchronously.
template <int R, int SWEEP, bool kMulti, int kCoop>
__global__ void __launch_bounds__(kTraceBlock, (R == 1 ? 3 : (R == 2 ? 2 : 1)))
    fused_trace_kernel(const __grid_constant__ TraceParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar[2];
    const uint32_t n = P.n_spheres;
    // the packed sweep stages the pair layout (even sphere count, zero padded); the others the plain AoS list
    const float4* __restrict__ g_src = SWEEP == kSweepPacked ? P.geom_pairs : P.geom;
    const uint32_t n_stage = SWEEP == kSweepPacked ? ((n + 1u) & ~1u) : n;  // spheres worth of bytes to copy
    const uint32_t n_tiles = kMulti ? (n + kTileSpheres - 1u) / kTileSpheres : 1u;
    constexpr uint32_t kGran = 32u * kCoop;  // buffers hold whole (super-)chunks
    const uint32_t tile_cap = kMulti ? kTileSpheres : ((n + kGran - 1u) / kGran) * kGran;
    float4* s_tile0 = reinterpret_cast<float4*>(smem_raw);
    float4* s_tile1 = s_tile0 + tile_cap;  // kMulti: second streaming buffer; else (packed sweep): AoS copy of the list
    constexpr bool kAosCopy = !kMulti && SWEEP == kSweepPacked;
    const float4* s_aos = kAosCopy ? s_tile1 : nullptr;
    uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_tile0 + ((kMulti || kAosCopy) ? 2u : 1u) * tile_cap) + threadIdx.x;

    // zero the tile buffers once (padding entries are read by the unrolled sweep and then masked off)
    for (uint32_t i = threadIdx.x; i < ((kMulti || kAosCopy) ? 2u : 1u) * tile_cap; i += kTraceBlock)
        s_tile0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        fence_mbar_init();
    }
    // generic-proxy writes (the zero fill) must be ordered before the async-proxy (TMA) writes to the same bytes
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    uint32_t bar_phase0 = 0u, bar_phase1 = 0u;
    if (!kMulti) {
        if (threadIdx.x == 0 && n > 0u) {
            mbar_arrive_expect_tx(&s_bar[0], n_stage * 16u + (kAosCopy ? n * 16u : 0u));
            tma_bulk_g2s(s_tile0, g_src, n_stage * 16u, &s_bar[0]);
            if (kAosCopy) tma_bulk_g2s(s_tile1, P.geom, n * 16u, &s_bar[0]);  // candidate resolution reads AoS
        }
        if (n > 0u) mbar_wait(&s_bar[0], 0u);
    }

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t k0 = P.key0, k1 = P.key1;

    // per-slot path state
    f3 o[R], d[R];
    double thr_r[R], thr_g[R], thr_b[R];  // product of attenuations so far (Float64, as the reference promotes)
    uint32_t pix_local[R];
    int depth_left[R];
    bool alive[R];
    PathRng rng[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        o[r] = mk3(0.f, 0.f, 0.f);
        d[r] = mk3(0.f, 1.f, 0.f);
        thr_r[r] = thr_g[r] = thr_b[r] = 1.0;
        pix_local[r] = 0u;
        depth_left[r] = 0;
        alive[r] = false;
        rng[r].sample = 0u;
        rng[r].pixel = 0u;
    }
    bool done = false;  // lane-level: no more tickets
    uint32_t seg_count = 0;
    // warp-level ticket pool (uniform)
    unsigned long long pool_next = 0, pool_end = 0;
    bool exhausted = false;

    for (;;) {
        // ------------------------------------------------------------ regenerate: idle slots take the next path
#pragma unroll
        for (int r = 0; r < R; ++r) {
            bool want = !alive[r] && !done;
            unsigned pending = __ballot_sync(kFullMask, want);
            if (pending == 0u) continue;
            unsigned long long ticket = 0;
            bool got = false;
            while (pending) {
                if (exhausted) {
                    if (want) { done = true; want = false; }
                    break;
                }
                if (pool_next >= pool_end) {
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(P.counters, (unsigned long long)kPoolChunk);
                    base = __shfl_sync(kFullMask, base, 0);
                    if (base >= P.n_paths) { exhausted = true; continue; }
                    pool_next = base;
                    pool_end = base + kPoolChunk < P.n_paths ? base + kPoolChunk : P.n_paths;
                }
                unsigned avail = (unsigned)(pool_end - pool_next);
                unsigned rank = __popc(pending & lt_mask);
                if (want && rank < avail) { ticket = pool_next + rank; want = false; got = true; }
                unsigned npend = __popc(pending);
                pool_next += npend < avail ? npend : avail;
                pending = __ballot_sync(kFullMask, want);
            }
            if (got) {
                // ticket -> (pixel, sample): consecutive tickets are consecutive samples of one pixel
                uint32_t pl, s0;
                if ((P.n_paths >> 32) == 0ull) {
                    pl = (uint32_t)ticket / (uint32_t)P.spp;
                    s0 = (uint32_t)ticket - pl * (uint32_t)P.spp;
                } else {
                    unsigned long long q = ticket / (unsigned)P.spp;
                    pl = (uint32_t)q;
                    s0 = (uint32_t)(ticket - q * (unsigned)P.spp);
                }
                s0 += (uint32_t)P.sample_first;
                uint32_t row_local = pl / (uint32_t)P.W;
                uint32_t col = pl - row_local * (uint32_t)P.W;
                uint32_t i0 = (uint32_t)P.row_start + row_local * (uint32_t)P.row_stride;
                // u = T(j/W), v = T((H-i)/H): quotient in Float64, rounded to Float32 (src/render.jl:26-27).
                // Both operands are integers < 2^24, so the correctly rounded Float32 quotient is the same value:
                // rounding through Float64 (53 >= 2*24+2 bits) is innocuous for division.
                float su = __fdiv_rn((float)(col + 1u), (float)P.W);
                float sv = __fdiv_rn((float)((uint32_t)P.H - 1u - i0), (float)P.H);
                rng[r].pixel = i0 * (uint32_t)P.W + col;
                rng[r].sample = s0;
                primary_ray(P.cam, rng[r], k0, k1, s0, su, sv, (float)P.W, (float)P.H, o[r], d[r]);
                thr_r[r] = thr_g[r] = thr_b[r] = 1.0;
                depth_left[r] = P.max_depth;
                pix_local[r] = pl;
                alive[r] = true;
            }
        }
        bool any_alive = false;
#pragma unroll
        for (int r = 0; r < R; ++r) any_alive |= alive[r];
        if (kMulti) {
            if (__syncthreads_or(any_alive ? 1 : 0) == 0) break;  // also: everyone is done with both tile buffers
        } else {
            if (__ballot_sync(kFullMask, any_alive) == 0u) break;
        }

        // ------------------------------------------------------------ intersect: closest hit over the list
        float best_t[R];
        int best_k[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            best_t[r] = __int_as_float(0x7f800000);  // typemax(T) = Inf, src/ray_color.jl:19
            best_k[r] = -1;
        }
        if (!kMulti) {
            sweep_tile<R, SWEEP, kCoop, kTraceBlock>(s_tile0, s_aos, n, 0u, s_mask, o, d, alive, best_t, best_k);
        } else {
            if (threadIdx.x == 0) {  // prologue: tile 0 -> buffer 0
                uint32_t cnt = n_stage < kTileSpheres ? n_stage : kTileSpheres;
                mbar_arrive_expect_tx(&s_bar[0], cnt * 16u);
                tma_bulk_g2s(s_tile0, g_src, cnt * 16u, &s_bar[0]);
            }
            for (uint32_t t = 0; t < n_tiles; ++t) {
                const uint32_t base = t * kTileSpheres;
                const uint32_t cnt = n - base < kTileSpheres ? n - base : kTileSpheres;
                if (threadIdx.x == 0 && t + 1u < n_tiles) {  // prefetch tile t+1 into the other buffer
                    const uint32_t nb = base + kTileSpheres;
                    const uint32_t ncnt = n_stage - nb < kTileSpheres ? n_stage - nb : kTileSpheres;
                    unsigned long long* bar = &s_bar[(t + 1u) & 1u];
                    mbar_arrive_expect_tx(bar, ncnt * 16u);
                    tma_bulk_g2s((t & 1u) ? s_tile0 : s_tile1, g_src + nb, ncnt * 16u, bar);
                }
                const float4* tile = (t & 1u) ? s_tile1 : s_tile0;
                if (t & 1u) { mbar_wait(&s_bar[1], bar_phase1); bar_phase1 ^= 1u; }
                else { mbar_wait(&s_bar[0], bar_phase0); bar_phase0 ^= 1u; }
                sweep_tile<R, SWEEP, kCoop, kTraceBlock>(tile, nullptr, cnt, base, s_mask, o, d, alive, best_t, best_k);
                __syncthreads();  // the buffer may be overwritten by the prefetch issued in the next iteration
            }
        }

        // ------------------------------------------------------------ shade / scatter / accumulate
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (!alive[r]) continue;
            seg_count += 1;
            bool finished = false;
            double cr = 0.0, cg = 0.0, cb = 0.0;
            if (best_k[r] < 0) {  // miss: sky (src/ray_color.jl:36)
                double sr, sg, sb;
                skycolor(d[r], sr, sg, sb);
                cr = __dmul_rn(thr_r[r], sr);
                cg = __dmul_rn(thr_g[r], sg);
                cb = __dmul_rn(thr_b[r], sb);
                finished = true;
            } else if (--depth_left[r] == 0) {
                finished = true;  // the next ray_color call returns black (src/ray_color.jl:15-17)
            } else {
                float4 g = __ldg(P.geom + best_k[r]);
                float4 m = __ldg(P.mat + best_k[r]);
                uint32_t kind = __ldg(P.kind + best_k[r]);
                f3 att;
                shade_hit(o[r], d[r], best_t[r], g, m, kind, rng[r], (uint32_t)(P.max_depth - depth_left[r]), k0, k1, att);
                thr_r[r] = __dmul_rn(thr_r[r], (double)att.x);
                thr_g[r] = __dmul_rn(thr_g[r], (double)att.y);
                thr_b[r] = __dmul_rn(thr_b[r], (double)att.z);
            }
            if (finished) {
                // accumulate (src/render.jl:38): order-independent fixed-point atomics => the image is
                // bit-identical for any schedule, rays-per-lane setting and GPU count
                unsigned long long* a = P.accum + (unsigned long long)pix_local[r] * 4ull;
                atomicAdd(a + 0, (unsigned long long)__double2ll_rn(cr * P.fx_scale));
                atomicAdd(a + 1, (unsigned long long)__double2ll_rn(cg * P.fx_scale));
                atomicAdd(a + 2, (unsigned long long)__double2ll_rn(cb * P.fx_scale));
                alive[r] = false;
            }
        }
    }
    // ray-segment statistics: one atomic per warp
    for (int off = 16; off > 0; off >>= 1) seg_count += __shfl_xor_sync(kFullMask, seg_count, off);
    if (lane == 0 && seg_count) atomicAdd(P.counters + 1, (unsigned long long)seg_count);
}

#endif  // RTW_BUILD_VARIANTS

// ---- resolve: accum / n_samples -> sqrt -> Float32 (src/render.jl:40, src/vec.jl:22) ------------------------
__global__ void __launch_bounds__(256) resolve_kernel(const unsigned long long* __restrict__ accum, int W, int H,
                                                      int n_rows, int row_start, int row_stride, int spp,
                                                      double inv_scale, int column_major, float* __restrict__ out) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)n_rows * W;
    if (t >= total) return;
    int k, col;
    if (column_major) { col = (int)(t / n_rows); k = (int)(t - (long long)col * n_rows); }
    else { k = (int)(t / W); col = (int)(t - (long long)k * W); }
    const unsigned long long* a = accum + ((long long)k * W + col) * 4;
    float rgb[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double sum = (double)(long long)a[c] * inv_scale;
        double lin = sum / (double)spp;
        rgb[c] = (float)sqrt(lin);
    }
    long long at;
    if (column_major) at = ((long long)col * H + (row_start + (long long)k * row_stride)) * 3;
    else at = ((long long)k * W + col) * 3;
    out[at + 0] = rgb[0];
    out[at + 1] = rgb[1];
    out[at + 2] = rgb[2];
}

// ---- assemble: tiles[g][k][col][3] (row i0 = g + k*G) -> Julia column-major image ---------------------------
__global__ void __launch_bounds__(256) assemble_kernel(const float* __restrict__ tiles, int G, int W, int H,
                                                       int rows_pad, float* __restrict__ out) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)W * H;
    if (t >= total) return;
    int col = (int)(t / H), i0 = (int)(t - (long long)col * H);
    int g = i0 % G, k = i0 / G;
    const float* src = tiles + (((long long)g * rows_pad + k) * W + col) * 3;
    float* dst = out + t * 3;
    dst[0] = src[0];
    dst[1] = src[1];
    dst[2] = src[2];
}

// ---- FP32 issue microbenchmarks ------------------------------------------------------------------------------
constexpr int kPeakBlock = 256;
constexpr int kPeakIters = 4096;
constexpr int kPeakChains = 16;
constexpr int kPeakSpheres = 512;
constexpr int kPeakSweeps = 64;

__global__ void __launch_bounds__(kPeakBlock) fp32_peak_ffma_kernel(float* out, float b, float c) {
    float a[kPeakChains];
#pragma unroll
    for (int i = 0; i < kPeakChains; ++i) a[i] = (float)(threadIdx.x + i) * 1e-3f;
    for (int it = 0; it < kPeakIters; ++it) {
#pragma unroll
        for (int i = 0; i < kPeakChains; ++i) a[i] = fmaf(a[i], b, c);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kPeakChains; ++i) s += a[i];
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true: keeps the chains alive
}

// variant 9: the same with Float64 chains -- the measured DFMA rate, denominator of the Float64 kernel's fraction
__global__ void __launch_bounds__(kPeakBlock) fp64_peak_dfma_kernel(float* out, double b, double c) {
    double a[kPeakChains];
#pragma unroll
    for (int i = 0; i < kPeakChains; ++i) a[i] = (double)(threadIdx.x + i) * 1e-3;
    for (int it = 0; it < kPeakIters / 4; ++it) {
#pragma unroll
        for (int i = 0; i < kPeakChains; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kPeakChains; ++i) s += a[i];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;  // never true: keeps the chains alive
}

// the sweep's own instruction mix (mask variant: 11 FP32 + 1 SHF per test, 1 LDS.128 per R tests) with no
// candidate ever resolved; ray data comes from memory so nothing is constant-folded
template <int R, bool kPacked, int kCoopPeak>
__global__ void __launch_bounds__(kPeakBlock) fp32_peak_sweep_kernel(float* out, const float* __restrict__ rays) {
    __shared__ float4 s_geom[kPeakSpheres];
    __shared__ uint32_t s_mask_peak[(kPeakSpheres / 32) * R * kPeakBlock];  // (n/(32 kCoop)) super-chunks x kCoop slots
    for (int i = threadIdx.x; i < kPeakSpheres; i += blockDim.x)
        s_geom[i] = kPacked ? ((i & 1) ? make_float4(-3000.f, -3000.f, 0.5f, 0.5f) : make_float4(1000.f + (float)i, 1001.f + (float)i, 2000.f, 2000.f))
                            : make_float4(1000.f + (float)i, 2000.f, -3000.f, 0.5f);  // far off-axis: disc < 0 always
    __syncthreads();
    f3 o[R], d[R];
    bool alive[R];
    float best_t[R];
    int best_k[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        // every component differs per lane, so nothing is uniform or shared between cooperating slots
        const float t = (float)threadIdx.x;
        o[r] = mk3(rays[0] + t * 1e-3f, rays[1] + (float)r + t * 2e-3f, rays[2] - t * 1e-3f);
        d[r] = mk3(rays[3] + t * 1e-4f, rays[4] - t * 1e-4f, rays[5] + t * 2e-4f);
        alive[r] = true;
        best_t[r] = __int_as_float(0x7f800000);
        best_k[r] = -1;
    }
    int hits = 0;
    for (int it = 0; it < kPeakSweeps; ++it) {
        if (kPacked) sweep_tile<R, kSweepPacked, kCoopPeak, kPeakBlock>(s_geom, nullptr, kPeakSpheres, 0u, s_mask_peak + threadIdx.x, o, d, alive, best_t, best_k);
        else sweep_tile_mask<R, kPeakBlock>(s_geom, kPeakSpheres, 0u, s_mask_peak + threadIdx.x, o, d, alive, best_t, best_k);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            hits += best_k[r] >= 0;
            o[r].x += 1e-3f;
        }
    }
    if (hits) out[blockIdx.x * blockDim.x + threadIdx.x] = (float)hits;
}

// Warp-specialisation experiment: warps [0, kSweepWarps) run the packed sweep, the remaining warps run a
// shade-like integer/SFU mix (Philox blocks, IEEE sqrt/rcp, lane-divergent retries) until the sweep warps are done.
// Reports only the sweep warps' FP32 work: does divergent integer work in other warps slow the sweep down?
constexpr int kMixSweepWarps = 8;
template <int kCoopPeak, int kOtherWarps>
__global__ void __launch_bounds__((kMixSweepWarps + kOtherWarps) * 32)
    fp32_peak_mixed_kernel(float* out, const float* __restrict__ rays) {
    constexpr int kSweepThreads = kMixSweepWarps * 32;
    __shared__ float4 s_geom[kPeakSpheres];
    __shared__ uint32_t s_mask_peak[(kPeakSpheres / 32) * kSweepThreads];
    __shared__ volatile int s_done;
    for (int i = threadIdx.x; i < kPeakSpheres; i += blockDim.x)
        s_geom[i] = (i & 1) ? make_float4(-3000.f, -3000.f, 0.5f, 0.5f) : make_float4(1000.f + (float)i, 1001.f + (float)i, 2000.f, 2000.f);
    if (threadIdx.x == 0) s_done = 0;
    __syncthreads();
    const float t = (float)threadIdx.x;
    if (threadIdx.x < kSweepThreads) {
        f3 o[1], d[1];
        bool alive[1] = {true};
        float best_t[1] = {__int_as_float(0x7f800000)};
        int best_k[1] = {-1};
        o[0] = mk3(rays[0] + t * 1e-3f, rays[1] + t * 2e-3f, rays[2] - t * 1e-3f);
        d[0] = mk3(rays[3] + t * 1e-4f, rays[4] - t * 1e-4f, rays[5] + t * 2e-4f);
        int hits = 0;
        for (int it = 0; it < kPeakSweeps; ++it) {
            sweep_tile<1, kSweepPacked, kCoopPeak, kSweepThreads>(s_geom, nullptr, kPeakSpheres, 0u, s_mask_peak + threadIdx.x, o, d, alive, best_t, best_k);
            hits += best_k[0] >= 0;
            o[0].x += 1e-3f;
        }
        if (hits) out[blockIdx.x * blockDim.x + threadIdx.x] = (float)hits;
        if (threadIdx.x == 0) s_done = 1;  // warp 0 finishes with the others (identical work)
    } else {
        PathRng g{threadIdx.x, blockIdx.x};
        uint32_t ev = 1, acc = 0;
        float facc = 0.f;
        while (!s_done) {
            u32x4 b0 = philox_block(g, ev, 0u, 0x1234u, 0x5678u);
            f3 v = rng_unit_vector(g, ev, b0, 0x1234u, 0x5678u);  // divergent rejection retries
            f3 n = normalize3(mk3(v.x + 0.5f, v.y - 0.25f, v.z + 0.125f));
            facc += __fdiv_rn(n.x, 0.2f) + n.y * n.z;
            acc ^= b0.w3;
            ++ev;
        }
        if (acc == 0x12345u && facc == 1.5f) out[blockIdx.x * blockDim.x + threadIdx.x] = facc;
    }
}

}  // namespace

// ---- host launchers --------------------------------------------------------------------------------------------

namespace {

#ifdef RTW_BUILD_VARIANTS
template <int R, int SWEEP, bool kMulti, int kCoop>
cudaError_t launch_trace_variant(const TraceParams& p, int num_sms, int blocks_per_sm_override, cudaStream_t stream,
                                 LaunchInfo* info) {
    auto kern = fused_trace_kernel<R, SWEEP, kMulti, kCoop>;
    constexpr uint32_t kGran = 32u * kCoop;
    const uint32_t tile_cap = kMulti ? kTileSpheres : ((p.n_spheres + kGran - 1u) / kGran) * kGran;
    const uint32_t chunks = tile_cap / 32u;  // mask words per lane and slot-set: (tile_cap / kGran) * kCoop
    constexpr bool kAosCopy = !kMulti && SWEEP == kSweepPacked;
    int smem = (int)(((kMulti || kAosCopy) ? 2u : 1u) * tile_cap * 16u);
    if (SWEEP != kSweepBranch) smem += (int)(chunks * R * kTraceBlock * 4u);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTraceBlock, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (blocks_per_sm_override > 0 && blocks_per_sm_override < per_sm) per_sm = blocks_per_sm_override;
    // persistent grid: every CTA is resident; never launch more lanes than there are paths
    long long grid = (long long)num_sms * per_sm;
    long long max_useful = (long long)((p.n_paths + (unsigned long long)(kTraceBlock * R) - 1ull) /
                                       (unsigned long long)(kTraceBlock * R));
    if (grid > max_useful) grid = max_useful;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, kTraceBlock, smem, stream>>>(p);
    if (info) {
        info->grid = (int)grid;
        info->block = kTraceBlock;
        info->smem_bytes = smem;
        info->blocks_per_sm = per_sm;
        info->launches = 1;
        info->rays_per_lane = R;
        info->sweep = SWEEP;
    }
    return cudaGetLastError();
}

template <int R, int SWEEP, int kCoop = 1>
cudaError_t launch_trace_tiles(const TraceParams& p, int num_sms, int bps, cudaStream_t stream, LaunchInfo* info) {
    if (p.n_spheres <= kTileSpheres) return launch_trace_variant<R, SWEEP, false, kCoop>(p, num_sms, bps, stream, info);
    return launch_trace_variant<R, SWEEP, true, kCoop>(p, num_sms, bps, stream, info);
}

template <int SWEEP>
cudaError_t launch_trace_rays(const TraceParams& p, int num_sms, int bps, int R, cudaStream_t stream, LaunchInfo* info) {
    switch (R) {
        case 1: return launch_trace_tiles<1, SWEEP>(p, num_sms, bps, stream, info);
        case 4: return launch_trace_tiles<4, SWEEP>(p, num_sms, bps, stream, info);
        default: return launch_trace_tiles<2, SWEEP>(p, num_sms, bps, stream, info);
    }
}

#endif  // RTW_BUILD_VARIANTS

}  // namespace

cudaError_t launch_fused_trace(const TraceParams& p, int num_sms, int blocks_per_sm_override, int rays_per_lane,
                               int sweep, int coop, cudaStream_t stream, LaunchInfo* info) {
#ifndef RTW_BUILD_VARIANTS
    (void)p; (void)num_sms; (void)blocks_per_sm_override; (void)rays_per_lane; (void)sweep; (void)coop; (void)stream; (void)info;
    return cudaErrorNotSupported;  // built without RTW_BUILD_VARIANTS=1
#else
    if (sweep != kSweepBranch && sweep != kSweepMask && rays_per_lane == 1 && coop > 1) {
        if (coop == 2) return launch_trace_tiles<1, kSweepPacked, 2>(p, num_sms, blocks_per_sm_override, stream, info);
        return launch_trace_tiles<1, kSweepPacked, 4>(p, num_sms, blocks_per_sm_override, stream, info);
    }
    if (sweep == kSweepBranch) return launch_trace_rays<kSweepBranch>(p, num_sms, blocks_per_sm_override, rays_per_lane, stream, info);
    if (sweep == kSweepMask) return launch_trace_rays<kSweepMask>(p, num_sms, blocks_per_sm_override, rays_per_lane, stream, info);
    return launch_trace_rays<kSweepPacked>(p, num_sms, blocks_per_sm_override, rays_per_lane, stream, info);
#endif
}

cudaError_t launch_resolve(const unsigned long long* accum, int W, int H, int n_rows, int row_start, int row_stride,
                           int spp, double inv_scale, int column_major, float* out, cudaStream_t stream) {
    long long total = (long long)n_rows * W;
    if (total <= 0) return cudaSuccess;
    unsigned grid = (unsigned)((total + 255) / 256);
    resolve_kernel<<<grid, 256, 0, stream>>>(accum, W, H, n_rows, row_start, row_stride, spp, inv_scale, column_major,
                                             out);
    return cudaGetLastError();
}

cudaError_t launch_assemble(const float* tiles, int n_tiles, int W, int H, float* out, cudaStream_t stream) {
    long long total = (long long)W * H;
    if (total <= 0) return cudaSuccess;
    int rows_pad = (H + n_tiles - 1) / n_tiles;
    unsigned grid = (unsigned)((total + 255) / 256);
    assemble_kernel<<<grid, 256, 0, stream>>>(tiles, n_tiles, W, H, rows_pad, out);
    return cudaGetLastError();
}

cudaError_t launch_fp32_peak(int variant, int num_sms, float* scratch, cudaStream_t stream, double* fp32_instr) {
    // variant = base + 10 * L: L > 0 limits residency to L CTAs per SM (by padding dynamic shared memory), to
    // measure the sweep at the occupancy of the real kernel
    const int limit = variant / 10;
    variant %= 10;
    const int grid = num_sms * 8;
    size_t dyn = 0;
    if (limit > 0) dyn = (size_t)(227 * 1024) / (size_t)limit - 42 * 1024;  // static smem of the sweep kernels: ~40 KB
    auto launch = [&](auto kern) -> cudaError_t {
        if (dyn > 0) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
            if (e != cudaSuccess) return e;
        }
        kern<<<grid, kPeakBlock, dyn, stream>>>(scratch + 64, scratch);
        return cudaGetLastError();
    };
    *fp32_instr = (double)grid * kPeakBlock * 1.0 * (double)kPeakSweeps * kPeakSpheres * 11.0;
    switch (variant) {
        case 0:
            fp32_peak_ffma_kernel<<<grid, kPeakBlock, 0, stream>>>(scratch, 0.999f, 1e-4f);
            *fp32_instr = (double)grid * kPeakBlock * (double)kPeakIters * kPeakChains;
            return cudaGetLastError();
        case 9:
            fp64_peak_dfma_kernel<<<grid, kPeakBlock, 0, stream>>>(scratch, 0.999, 1e-4);
            *fp32_instr = (double)grid * kPeakBlock * (double)(kPeakIters / 4) * kPeakChains;
            return cudaGetLastError();
        case 1: return launch(fp32_peak_sweep_kernel<1, false, 1>);
        case 2: return launch(fp32_peak_sweep_kernel<1, true, 1>);
        case 3: return launch(fp32_peak_sweep_kernel<1, true, 2>);
        case 4: return launch(fp32_peak_sweep_kernel<1, true, 4>);
        default: break;
    }
    // variants 5..8: warp-specialisation experiment (8 sweep warps + 0 / 4 / 8 "shade-like" warps per CTA)
    *fp32_instr = (double)grid * (kMixSweepWarps * 32) * (double)kPeakSweeps * kPeakSpheres * 11.0;
    auto launch_mixed = [&](auto kern, int threads) -> cudaError_t {
        if (dyn > 0) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
            if (e != cudaSuccess) return e;
        }
        kern<<<grid, threads, dyn, stream>>>(scratch + 64, scratch);
        return cudaGetLastError();
    };
    switch (variant) {
        case 5: return launch_mixed(fp32_peak_mixed_kernel<2, 0>, kMixSweepWarps * 32);
        case 6: return launch_mixed(fp32_peak_mixed_kernel<2, 4>, (kMixSweepWarps + 4) * 32);
        case 7: return launch_mixed(fp32_peak_mixed_kernel<2, 8>, (kMixSweepWarps + 8) * 32);
        default: return launch_mixed(fp32_peak_mixed_kernel<4, 8>, (kMixSweepWarps + 8) * 32);
    }
}

}  // namespace rtw
