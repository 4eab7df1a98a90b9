// rtw_wavefront.cu -- RTW_MODE_WAVEFRONT: the same path tracer as separate kernels over a pool of paths in HBM.
//
//   K1  wf_terminate_regenerate_kernel : accumulate the paths that ended (sky or depth limit, src/ray_color.jl:15-17,36;
//                                        src/render.jl:38) and re-launch their slots with the next path ticket
//                                        (raygen, src/render.jl:26-37 + src/camera.jl:43-48)
//   K2  wf_intersect_kernel            : closest hit of every live ray over the sphere list (src/hit.jl:38-50), the
//                                        same packed FP32x2 sweep as the fused kernel; its epilogue sorts the slots
//                                        into per-class work lists with warp ballots (the compaction step)
//   K3  wf_scatter_kernel<CLASS>       : scatter() for one material class per launch (src/material.jl): every warp
//                                        runs one code path
//   K4  resolve_kernel (rtw_kernels.cu): accum / n_samples, gamma-2
//
// Slots are refilled in place, so the pool stays dense until the tickets run out.  Results are bit-identical to the
// fused kernel: same arithmetic, same addressed Philox stream, order-independent fixed-point accumulation.
// This mode exists to measure the design the fused kernel is compared against; the fused kernel is the
// default because ray state never leaves registers there.
#include "rtw_sweep.cuh"

namespace rtw {

namespace {

constexpr int kWfBlock = 256;
constexpr int kWfCoop = 2;

__device__ __forceinline__ void list_append(uint32_t* __restrict__ list, unsigned int* __restrict__ count, bool pred,
                                            uint32_t value) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned m = __ballot_sync(kFullMask, pred);
    if (m == 0u) return;
    unsigned base = 0;
    if (lane == (unsigned)(__ffs((int)m) - 1)) base = atomicAdd(count, (unsigned)__popc(m));
    base = __shfl_sync(kFullMask, base, __ffs((int)m) - 1);
    if (pred) list[base + __popc(m & ((1u << lane) - 1u))] = value;
}

// ---- K1: end-of-path accumulation + in-place regeneration --------------------------------------------------------
__global__ void __launch_bounds__(kWfBlock) wf_terminate_regenerate_kernel(const __grid_constant__ TraceParams P,
                                                                           const __grid_constant__ WavefrontBuffers B) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t n = B.counts[0];  // list 0: slots whose path ended in the previous intersect step
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t n_round = (n + 31u) & ~31u;  // whole warps stay together for the ballots
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n_round; e += stride) {
        const bool valid = e < n;
        uint32_t slot = 0;
        if (valid) {
            slot = B.list[0][e];
            const int hk = B.hit_k[slot];
            if (hk == -1) {  // miss: the sky colour times the path throughput (src/ray_color.jl:36)
                const float4 dd = B.ray_d[slot];
                double sr, sg, sb;
                skycolor(mk3(dd.x, dd.y, dd.z), sr, sg, sb);
                const double cr = __dmul_rn(B.thr[slot], sr);
                const double cg = __dmul_rn(B.thr[B.capacity + slot], sg);
                const double cb = __dmul_rn(B.thr[2ull * B.capacity + slot], sb);
                unsigned long long* a = P.accum + (unsigned long long)B.pix_local[slot] * 4ull;
                atomicAdd(a + 0, (unsigned long long)__double2ll_rn(cr * P.fx_scale));
                atomicAdd(a + 1, (unsigned long long)__double2ll_rn(cg * P.fx_scale));
                atomicAdd(a + 2, (unsigned long long)__double2ll_rn(cb * P.fx_scale));
            }
            // hk == -2: depth exhausted -> black (src/ray_color.jl:15-17); hk == -3: the slot was never used
        }
        // next path ticket for every slot of this warp that needs one
        const unsigned m = __ballot_sync(kFullMask, valid);
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(P.counters, (unsigned long long)__popc(m));
        base = __shfl_sync(kFullMask, base, 0);
        if (!valid) continue;
        const unsigned long long ticket = base + __popc(m & ((1u << lane) - 1u));
        if (ticket >= P.n_paths) {
            B.alive[slot] = 0u;
            continue;
        }
        const unsigned long long q = ticket / (unsigned)P.spp;
        const uint32_t pl = (uint32_t)q, s0 = (uint32_t)(ticket - q * (unsigned)P.spp) + (uint32_t)P.sample_first;
        const uint32_t row_local = pl / (uint32_t)P.W, col = pl - row_local * (uint32_t)P.W;
        const uint32_t i0 = (uint32_t)P.row_start + row_local * (uint32_t)P.row_stride;
        const float su = __fdiv_rn((float)(col + 1u), (float)P.W);                       // src/render.jl:26
        const float sv = __fdiv_rn((float)((uint32_t)P.H - 1u - i0), (float)P.H);        // src/render.jl:27
        PathRng rng;
        rng.pixel = i0 * (uint32_t)P.W + col;
        rng.sample = s0;
        f3 o, d;
        primary_ray(P.cam, rng, P.key0, P.key1, s0, su, sv, (float)P.W, (float)P.H, o, d);
        B.ray_o[slot] = make_float4(o.x, o.y, o.z, 0.f);
        B.ray_d[slot] = make_float4(d.x, d.y, d.z, 0.f);
        B.thr[slot] = 1.0;
        B.thr[B.capacity + slot] = 1.0;
        B.thr[2ull * B.capacity + slot] = 1.0;
        B.pix_local[slot] = pl;
        B.sample[slot] = s0;
        B.pixel[slot] = rng.pixel;
        B.depth_left[slot] = P.max_depth;
        B.alive[slot] = 1u;
    }
}

// ---- K2: intersect + classification --------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWfBlock, 3) wf_intersect_kernel(const __grid_constant__ TraceParams P,
                                                                   const __grid_constant__ WavefrontBuffers B) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar;
    const uint32_t n = P.n_spheres;
    constexpr uint32_t kGran = 32u * kWfCoop;
    const uint32_t tile_cap = ((n + kGran - 1u) / kGran) * kGran;
    float4* s_tile = reinterpret_cast<float4*>(smem_raw);
    float4* s_aos = s_tile + tile_cap;
    uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_tile + 2u * tile_cap) + threadIdx.x;
    for (uint32_t i = threadIdx.x; i < 2u * tile_cap; i += kWfBlock) s_tile[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0 && n > 0u) {
        const uint32_t n_stage = (n + 1u) & ~1u;
        mbar_arrive_expect_tx(&s_bar, n_stage * 16u + n * 16u);
        tma_bulk_g2s(s_tile, P.geom_pairs, n_stage * 16u, &s_bar);
        tma_bulk_g2s(s_aos, P.geom, n * 16u, &s_bar);
    }
    if (n > 0u) mbar_wait(&s_bar, 0u);

    uint32_t segs = 0;
    const uint32_t n_blocks = (B.capacity + kWfBlock - 1u) / kWfBlock;
    for (uint32_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const uint32_t slot = blk * kWfBlock + threadIdx.x;
        bool alive[1] = {slot < B.capacity && B.alive[slot] != 0u};
        if (__ballot_sync(kFullMask, alive[0]) == 0u) continue;
        f3 o[1] = {mk3(0.f, 0.f, 0.f)}, d[1] = {mk3(0.f, 1.f, 0.f)};
        if (alive[0]) {
            const float4 oo = B.ray_o[slot], dd = B.ray_d[slot];
            o[0] = mk3(oo.x, oo.y, oo.z);
            d[0] = mk3(dd.x, dd.y, dd.z);
        }
        float best_t[1] = {__int_as_float(0x7f800000)};
        int best_k[1] = {-1};
        sweep_tile<1, kSweepPacked, kWfCoop, kWfBlock>(s_tile, s_aos, n, 0u, s_mask, o, d, alive, best_t, best_k);
        // classification = the compaction step: every live slot goes to exactly one work list
        int cls = -1;
        if (alive[0]) {
            segs += 1;
            int hk = best_k[0];
            if (hk < 0) cls = 0;                                   // miss -> sky
            else if (B.depth_left[slot] == 1) { cls = 0; hk = -2; }  // the next ray_color call returns black
            else cls = __ldg(P.kind + hk) == 2u ? 2 : 1;            // dielectric | Lambertian/Metal
            B.hit_t[slot] = best_t[0];
            B.hit_k[slot] = hk;
        }
        list_append(B.list[0], B.counts + 0, cls == 0, slot);
        list_append(B.list[1], B.counts + 1, cls == 1, slot);
        list_append(B.list[2], B.counts + 2, cls == 2, slot);
    }
    for (int off = 16; off > 0; off >>= 1) segs += __shfl_xor_sync(kFullMask, segs, off);
    if ((threadIdx.x & 31u) == 0u && segs) {
        atomicAdd(P.counters + 1, (unsigned long long)segs);
        atomicAdd(B.counts + 3, segs);  // rays traced in this step (termination test on the host)
    }
}

// ---- K3: scatter for one material class ------------------------------------------------------------------------------
template <int CLASS>
__global__ void __launch_bounds__(kWfBlock) wf_scatter_kernel(const __grid_constant__ TraceParams P,
                                                              const __grid_constant__ WavefrontBuffers B) {
    const uint32_t n = B.counts[CLASS];
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const uint32_t slot = B.list[CLASS][e];
        const float4 oo = B.ray_o[slot], dd = B.ray_d[slot];
        f3 o = mk3(oo.x, oo.y, oo.z), d = mk3(dd.x, dd.y, dd.z);
        const int hk = B.hit_k[slot];
        const int depth_left = B.depth_left[slot] - 1;
        PathRng rng;
        rng.pixel = B.pixel[slot];
        rng.sample = B.sample[slot];
        const float4 g = __ldg(P.geom + hk);
        const float4 m = __ldg(P.mat + hk);
        const uint32_t kind = __ldg(P.kind + hk);
        f3 att;
        shade_hit(o, d, B.hit_t[slot], g, m, kind, rng, (uint32_t)(P.max_depth - depth_left), P.key0, P.key1, att);
        B.ray_o[slot] = make_float4(o.x, o.y, o.z, 0.f);
        B.ray_d[slot] = make_float4(d.x, d.y, d.z, 0.f);
        if (CLASS == 1) {  // dielectric attenuation is exactly (1,1,1): the throughput is unchanged
            B.thr[slot] = __dmul_rn(B.thr[slot], (double)att.x);
            B.thr[B.capacity + slot] = __dmul_rn(B.thr[B.capacity + slot], (double)att.y);
            B.thr[2ull * B.capacity + slot] = __dmul_rn(B.thr[2ull * B.capacity + slot], (double)att.z);
        }
        B.depth_left[slot] = depth_left;
    }
}

__global__ void wf_init_kernel(const __grid_constant__ WavefrontBuffers B) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B.capacity) {
        B.list[0][i] = i;
        B.hit_k[i] = -3;
        B.alive[i] = 0u;
    }
    if (i == 0) {
        B.counts[0] = B.capacity;
        B.counts[1] = B.counts[2] = B.counts[3] = 0u;
    }
}

__global__ void wf_reset_counts_kernel(const __grid_constant__ WavefrontBuffers B, unsigned int* traced_out) {
    if (threadIdx.x == 0) {
        if (traced_out) *traced_out = B.counts[3];
        B.counts[0] = B.counts[1] = B.counts[2] = B.counts[3] = 0u;
    }
}

}  // namespace

size_t wavefront_bytes(uint32_t capacity) {
    // 16 + 16 + 24 (rays, throughput) + 7 x 4 (pix, sample, pixel, depth, hit_t, hit_k, alive) + 3 x 4 (lists) per slot,
    // every array padded to 64 B
    return (size_t)capacity * (16 + 16 + 24 + 7 * 4 + 3 * 4) + 16 * 64;
}

cudaError_t launch_wavefront_trace(const TraceParams& p, const WavefrontBuffers& b, int num_sms, unsigned int* h_traced,
                                   cudaStream_t stream, LaunchInfo* info) {
    if (p.n_spheres > kTileSpheres) return cudaErrorNotSupported;
    constexpr uint32_t kGran = 32u * kWfCoop;
    const uint32_t tile_cap = ((p.n_spheres + kGran - 1u) / kGran) * kGran;
    const int smem = (int)(2u * tile_cap * 16u + (tile_cap / 32u) * kWfBlock * 4u);
    cudaError_t e = cudaFuncSetAttribute(wf_intersect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    const unsigned list_grid = (unsigned)((b.capacity + kWfBlock - 1u) / kWfBlock);
    const unsigned small_grid = list_grid < (unsigned)num_sms * 8u ? list_grid : (unsigned)num_sms * 8u;
    unsigned isect_grid = (unsigned)num_sms * 3u;
    if (isect_grid > list_grid) isect_grid = list_grid;
    int launches = 0;
    wf_init_kernel<<<list_grid, kWfBlock, 0, stream>>>(b);
    ++launches;
    const unsigned long long max_steps = 4ull + (unsigned long long)p.max_depth *
                                         ((p.n_paths + b.capacity - 1ull) / b.capacity + 1ull);
    unsigned long long steps = 0;
    for (;;) {
        // a chunk of steps is enqueued without host synchronisation; the ray count of the LAST step of the chunk
        // tells whether every path has ended (slots are refilled while tickets remain, so 0 means done)
        for (int it = 0; it < 16; ++it) {
            wf_terminate_regenerate_kernel<<<small_grid, kWfBlock, 0, stream>>>(p, b);
            wf_reset_counts_kernel<<<1, 32, 0, stream>>>(b, nullptr);
            wf_intersect_kernel<<<isect_grid, kWfBlock, smem, stream>>>(p, b);
            wf_scatter_kernel<1><<<small_grid, kWfBlock, 0, stream>>>(p, b);
            wf_scatter_kernel<2><<<small_grid, kWfBlock, 0, stream>>>(p, b);
            launches += 5;
            ++steps;
        }
        e = cudaMemcpyAsync(h_traced, b.counts + 3, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return e;
        if (*h_traced == 0u) break;
        if (steps > max_steps) return cudaErrorLaunchTimeout;  // cannot happen: every step ends or advances each path
    }
    // the last chunk's final K2 left list 0 empty (nothing traced); all contributions are in the accumulator
    if (info) {
        info->grid = (int)isect_grid;
        info->block = kWfBlock;
        info->smem_bytes = smem;
        info->blocks_per_sm = 3;
        info->launches = launches;
        info->rays_per_lane = 1;
        info->sweep = kSweepPacked;
    }
    return cudaGetLastError();
}

}  // namespace rtw
