// rtw_cta_wavefront.cu -- RTW_MODE_CTA_WAVEFRONT: the wavefront of the north-star kept entirely inside a persistent
// CTA.  256 path slots live in shared memory; every bounce the CTA runs the stages
//
//   regenerate (K1)  ended slots are refilled with the next path tickets (raygen, src/render.jl:26-37, camera.jl:43-48)
//   intersect  (K2)  the packed FP32x2 sweep of the whole sphere list (src/hit.jl:38-50), branch free: masks only
//   compact    (K0)  all candidates of all 256 rays are enumerated into ONE list (warp scan + one atomic per warp) and
//                    resolved by the whole CTA, list entry by list entry: balanced, no lane waits for a busy neighbour;
//                    closest hit per ray by a 64-bit atomicMin on (t, ~k) -- smaller t wins, equal t goes to the later
//                    sphere (src/hit.jl:24-26,44-46)
//   sort + shade (K3) slots are counting-sorted by what happens next (ended / Lambertian+Metal / Dielectric) so that a
//                    warp shades one class; ended slots are exactly the list regenerate refills
//   accumulate (K4)  fixed-point atomics, then resolve_kernel
//
// Same arithmetic, same addressed Philox stream and the same order-independent accumulation as the other modes, so the
// image is bit-identical (tests/test_gpu_parity.py).  Lists stay in shared memory: no HBM traffic for rays.
#include "rtw_sweep.cuh"

namespace rtw {

namespace {

constexpr int kCtaBlock = 256;
constexpr int kCtaWarps = kCtaBlock / 32;
constexpr int kCtaCoop = 2;
constexpr uint32_t kListCap = 2048;  // candidate entries per bounce (mean ~3.4 per ray); overflow is resolved in place
constexpr unsigned long long kNoHit = ~0ull;

struct CtaShared {  // fixed-size part; the geometry tile, AoS copy and masks follow in dynamic shared memory
    float ox[kCtaBlock], oy[kCtaBlock], oz[kCtaBlock], dx[kCtaBlock], dy[kCtaBlock], dz[kCtaBlock];
    double thr[3][kCtaBlock];
    unsigned long long key[kCtaBlock];  // closest hit of the slot's ray: (t bits << 32) | ~k
    uint32_t pix[kCtaBlock], sample[kCtaBlock], pixel[kCtaBlock];
    int depth[kCtaBlock];  // bounces left; 0 = slot idle
    uint32_t list[kListCap];
    uint32_t perm[kCtaBlock];
    uint32_t warp_count[kCtaWarps][4];
    uint32_t list_count;
    unsigned long long bar;
};

__device__ __forceinline__ unsigned long long pack_hit(float t, uint32_t k) {
    return ((unsigned long long)__float_as_uint(t) << 32) | (unsigned long long)(0xffffffffu - k);
}

// hit(::Sphere) for one (ray, sphere) candidate against tmin only (src/hit.jl:12-29 with tmax = Inf); the closest hit
// of the ray is the minimum over its candidates, ties to the later sphere -- identical to the sequential sweep
__device__ __forceinline__ void resolve_candidate(CtaShared& S, const float4* __restrict__ aos, uint32_t rs, uint32_t kl) {
    const float tmin = 1e-4f;
    const float4 s = aos[kl];
    const f3 o = mk3(S.ox[rs], S.oy[rs], S.oz[rs]), d = mk3(S.dx[rs], S.dy[rs], S.dz[rs]);
    const f3 oc = mk3(o.x - s.x, o.y - s.y, o.z - s.z);
    const float hb = dot3(oc, d);
    const float cq = fmaf(-s.w, s.w, dot3(oc, oc));
    if (hb > 0.0f && cq > 0.0f) return;  // wholly behind the origin: both roots <= 0 < tmin
    const float disc = fmaf(hb, hb, -cq);
    const float sq = __fsqrt_rn(disc);
    float root = -hb - sq;
    if (root < tmin) {
        root = -hb + sq;
        if (root < tmin) return;
    }
    atomicMin(&S.key[rs], pack_hit(root, kl));
}

// list entry = slot << 16 | coop parity << 15 | super-chunk << 5 | mask bit (test j of the lane's super-chunk)
__device__ __forceinline__ void resolve_entry(CtaShared& S, const float4* __restrict__ aos, uint32_t n, uint32_t ent) {
    const uint32_t j = ent & 31u, c = (ent >> 5) & 1023u, h = (ent >> 15) & 1u;
    const uint32_t kl = c * (32u * kCtaCoop) + 2u * ((j >> 1) * kCtaCoop + h) + (j & 1u);
    if (kl < n) resolve_candidate(S, aos, ent >> 16, kl);  // kl == n: the zero pad partner of an odd last sphere
}

__global__ void __launch_bounds__(kCtaBlock, 3) cta_wavefront_kernel(const __grid_constant__ TraceParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t n = P.n_spheres;
    constexpr uint32_t kGran = 32u * kCtaCoop;
    constexpr uint32_t kSuper = kGran;
    const uint32_t tile_cap = ((n + kGran - 1u) / kGran) * kGran;
    const uint32_t nsc = tile_cap / kSuper;
    float4* s_tile = reinterpret_cast<float4*>(smem_raw);
    float4* s_aos = s_tile + tile_cap;
    uint32_t* s_mask_base = reinterpret_cast<uint32_t*>(s_tile + 2u * tile_cap);
    CtaShared& S = *reinterpret_cast<CtaShared*>(s_mask_base + nsc * kCtaCoop * kCtaBlock);

    const uint32_t tid = threadIdx.x;
    const unsigned lane = tid & 31u, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t k0 = P.key0, k1 = P.key1;

    for (uint32_t i = tid; i < 2u * tile_cap; i += kCtaBlock) s_tile[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    S.depth[tid] = 0;
    S.perm[tid] = tid;  // every slot starts "ended": the first regenerate fills all of them
    S.key[tid] = pack_hit(0.f, 0u);
    if (tid == 0) {
        S.list_count = 0u;
        mbar_init(&S.bar, 1);
        fence_mbar_init();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0 && n > 0u) {
        const uint32_t n_stage = (n + 1u) & ~1u;
        mbar_arrive_expect_tx(&S.bar, n_stage * 16u + n * 16u);
        tma_bulk_g2s(s_tile, P.geom_pairs, n_stage * 16u, &S.bar);
        tma_bulk_g2s(s_aos, P.geom, n * 16u, &S.bar);
    }
    if (n > 0u) mbar_wait(&S.bar, 0u);

    uint32_t n_end = kCtaBlock;  // slots perm[0 .. n_end) ended in the previous bounce
    uint32_t seg_count = 0;
    bool did_shade = false;

    for (;;) {
        // ------------------------------------------------------------ K1 regenerate: refill the ended slots
        bool live_work = did_shade;  // this thread knows of a live ray: one it just shaded, or one it regenerates now
        {
            const bool mine = tid < n_end;
            const uint32_t slot = mine ? S.perm[tid] : 0u;
            const unsigned m = __ballot_sync(kFullMask, mine);
            if (m) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(P.counters, (unsigned long long)__popc(m));
                base = __shfl_sync(kFullMask, base, 0);
                if (mine) {
                    const unsigned long long ticket = base + __popc(m & lt_mask);
                    if (ticket >= P.n_paths) {
                        S.depth[slot] = 0;
                    } else {
                        uint32_t pl, s0;
                        if ((P.n_paths >> 32) == 0ull) {
                            pl = (uint32_t)ticket / (uint32_t)P.spp;
                            s0 = (uint32_t)ticket - pl * (uint32_t)P.spp;
                        } else {
                            const unsigned long long q = ticket / (unsigned)P.spp;
                            pl = (uint32_t)q;
                            s0 = (uint32_t)(ticket - q * (unsigned)P.spp);
                        }
                        s0 += (uint32_t)P.sample_first;
                        const uint32_t row_local = pl / (uint32_t)P.W, col = pl - row_local * (uint32_t)P.W;
                        const uint32_t i0 = (uint32_t)P.row_start + row_local * (uint32_t)P.row_stride;
                        const float su = __fdiv_rn((float)(col + 1u), (float)P.W);                  // src/render.jl:26
                        const float sv = __fdiv_rn((float)((uint32_t)P.H - 1u - i0), (float)P.H);   // src/render.jl:27
                        PathRng rng;
                        rng.pixel = i0 * (uint32_t)P.W + col;
                        rng.sample = s0;
                        f3 o, d;
                        primary_ray(P.cam, rng, k0, k1, s0, su, sv, (float)P.W, (float)P.H, o, d);
                        S.ox[slot] = o.x; S.oy[slot] = o.y; S.oz[slot] = o.z;
                        S.dx[slot] = d.x; S.dy[slot] = d.y; S.dz[slot] = d.z;
                        S.thr[0][slot] = 1.0; S.thr[1][slot] = 1.0; S.thr[2][slot] = 1.0;
                        S.pix[slot] = pl; S.sample[slot] = s0; S.pixel[slot] = rng.pixel;
                        S.depth[slot] = P.max_depth;
                        live_work = true;
                    }
                }
            }
        }
        // barrier: every ray of the next bounce is in shared memory; exit when no thread knows of a live ray
        if (__syncthreads_or(live_work ? 1 : 0) == 0) break;

        // ------------------------------------------------------------ K2 intersect: masks only
        const uint32_t h = tid & 1u;
        const uint32_t slots[2] = {tid, tid ^ 1u};  // own ray and the neighbour's (2-lane cooperation through smem)
        f3 o[2], d[2];
        bool alive[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            o[r] = mk3(S.ox[slots[r]], S.oy[slots[r]], S.oz[slots[r]]);
            d[r] = mk3(S.dx[slots[r]], S.dy[slots[r]], S.dz[slots[r]]);
            alive[r] = S.depth[slots[r]] > 0;
        }
        S.key[tid] = kNoHit;
        __syncwarp();  // the neighbour lane may resolve overflow candidates of this slot in place
        seg_count += alive[0] ? 1u : 0u;
        uint32_t* s_mask = s_mask_base + tid;
        uint32_t summary[2];
        sweep_masks_packed<2, kCtaCoop, kCtaBlock>(s_tile, n, h, s_mask, o, d, summary);

        // ------------------------------------------------------------ K0 compaction: one candidate list per CTA
        // entries are (slot << 16 | super-chunk << 5 | bit): decoding to a sphere index is left to the balanced
        // resolve loop; the per-lane candidate count comes from popcounts of the mask words
        uint32_t cnt = 0;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (!alive[r]) summary[r] = 0u;
            for (uint32_t sum = summary[r]; sum; sum &= sum - 1u) {
                const uint32_t c = (uint32_t)__ffs((int)sum) - 1u;
                cnt += (uint32_t)__popc(~s_mask[(c * 2u + r) * kCtaBlock]);
            }
        }
        uint32_t incl = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t v = __shfl_up_sync(kFullMask, incl, off);
            if (lane >= (unsigned)off) incl += v;
        }
        uint32_t base = 0;
        if (lane == 31u && incl) base = atomicAdd(&S.list_count, incl);
        base = __shfl_sync(kFullMask, base, 31);
        uint32_t pos = base + incl - cnt;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const uint32_t tag = (slots[r] << 16) | (h << 15);
            for (uint32_t sum = summary[r]; sum; sum &= sum - 1u) {
                const uint32_t c = (uint32_t)__ffs((int)sum) - 1u;
                for (uint32_t cand = ~s_mask[(c * 2u + r) * kCtaBlock]; cand; ++pos) {
                    const uint32_t j = (uint32_t)__clz((int)cand);
                    cand &= ~(0x80000000u >> j);
                    const uint32_t ent = tag | (c << 5) | j;
                    if (pos < kListCap) S.list[pos] = ent;
                    else resolve_entry(S, s_aos, n, ent);  // list overflow: resolve in place
                }
            }
        }
        __syncthreads();
        const uint32_t total = S.list_count < kListCap ? S.list_count : kListCap;
        for (uint32_t e = tid; e < total; e += kCtaBlock) resolve_entry(S, s_aos, n, S.list[e]);
        __syncthreads();

        // ------------------------------------------------------------ sort the slots by what happens next
        // class 0: path ends (sky, or depth exhausted: the next ray_color call returns black); 1: Lambertian/Metal;
        // 2: Dielectric; 3: idle slot
        int cls = 3;
        const unsigned long long mykey = S.key[tid];
        if (S.depth[tid] > 0) {
            if (mykey == kNoHit) {
                // K4: the path left the scene: sky colour times throughput (src/ray_color.jl:36, src/render.jl:38)
                cls = 0;
                double sr, sg, sb;
                skycolor(mk3(S.dx[tid], S.dy[tid], S.dz[tid]), sr, sg, sb);
                unsigned long long* a = P.accum + (unsigned long long)S.pix[tid] * 4ull;
                atomicAdd(a + 0, (unsigned long long)__double2ll_rn(__dmul_rn(S.thr[0][tid], sr) * P.fx_scale));
                atomicAdd(a + 1, (unsigned long long)__double2ll_rn(__dmul_rn(S.thr[1][tid], sg) * P.fx_scale));
                atomicAdd(a + 2, (unsigned long long)__double2ll_rn(__dmul_rn(S.thr[2][tid], sb) * P.fx_scale));
            } else if (S.depth[tid] == 1) {
                cls = 0;  // depth exhausted: the next ray_color call returns black (src/ray_color.jl:15-17)
            } else {
                cls = __ldg(P.kind + (0xffffffffu - (uint32_t)mykey)) == 2u ? 2 : 1;
            }
        }
        uint32_t rank = 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const unsigned mc = __ballot_sync(kFullMask, cls == c);
            if (cls == c) rank = (uint32_t)__popc(mc & lt_mask);
            if (lane == 0) S.warp_count[warp][c] = (uint32_t)__popc(mc);
        }
        if (tid == 0) S.list_count = 0u;
        __syncthreads();
        uint32_t tot[3] = {0u, 0u, 0u}, before[3] = {0u, 0u, 0u};
#pragma unroll
        for (int w = 0; w < kCtaWarps; ++w) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const uint32_t v = S.warp_count[w][c];
                tot[c] += v;
                if ((unsigned)w < warp) before[c] += v;
            }
        }
        if (cls == 0) S.perm[before[0] + rank] = tid;
        else if (cls == 1) S.perm[tot[0] + before[1] + rank] = tid;
        else if (cls == 2) S.perm[tot[0] + tot[1] + before[2] + rank] = tid;
        n_end = tot[0];
        __syncthreads();

        // ------------------------------------------------------------ K3 shade: one material class per warp (mostly)
        did_shade = tid >= n_end && tid < tot[0] + tot[1] + tot[2];
        if (did_shade) {
            const uint32_t slot = S.perm[tid];
            const unsigned long long key = S.key[slot];
            const uint32_t hk = 0xffffffffu - (uint32_t)key;
            const float t = __uint_as_float((uint32_t)(key >> 32));
            f3 ro = mk3(S.ox[slot], S.oy[slot], S.oz[slot]), rd = mk3(S.dx[slot], S.dy[slot], S.dz[slot]);
            const int depth_left = S.depth[slot] - 1;
            PathRng rng;
            rng.pixel = S.pixel[slot];
            rng.sample = S.sample[slot];
            const float4 g = s_aos[hk];
            const float4 mm = __ldg(P.mat + hk);
            const uint32_t kind = __ldg(P.kind + hk);
            f3 att;
            shade_hit(ro, rd, t, g, mm, kind, rng, (uint32_t)(P.max_depth - depth_left), k0, k1, att);
            S.ox[slot] = ro.x; S.oy[slot] = ro.y; S.oz[slot] = ro.z;
            S.dx[slot] = rd.x; S.dy[slot] = rd.y; S.dz[slot] = rd.z;
            if (kind != 2u) {  // dielectric attenuation is exactly (1,1,1)
                S.thr[0][slot] = __dmul_rn(S.thr[0][slot], (double)att.x);
                S.thr[1][slot] = __dmul_rn(S.thr[1][slot], (double)att.y);
                S.thr[2][slot] = __dmul_rn(S.thr[2][slot], (double)att.z);
            }
            S.depth[slot] = depth_left;
        }
        // no barrier here: thread tid < n_end regenerates exactly the slot it would have shaded (perm[tid]); the
        // barrier after regenerate orders every smem write before the next sweep
    }
    for (int off = 16; off > 0; off >>= 1) seg_count += __shfl_xor_sync(kFullMask, seg_count, off);
    if (lane == 0 && seg_count) atomicAdd(P.counters + 1, (unsigned long long)seg_count);
}

}  // namespace

cudaError_t launch_cta_wavefront_trace(const TraceParams& p, int num_sms, int blocks_per_sm_override,
                                       cudaStream_t stream, LaunchInfo* info) {
    if (p.n_spheres > kTileSpheres) return cudaErrorNotSupported;
    constexpr uint32_t kGran = 32u * kCtaCoop;
    const uint32_t tile_cap = ((p.n_spheres + kGran - 1u) / kGran) * kGran;
    const int smem = (int)(2u * tile_cap * 16u + (tile_cap / 32u) * kCtaBlock * 4u + sizeof(CtaShared) + 128);
    cudaError_t e = cudaFuncSetAttribute(cta_wavefront_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cta_wavefront_kernel, kCtaBlock, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (blocks_per_sm_override > 0 && blocks_per_sm_override < per_sm) per_sm = blocks_per_sm_override;
    long long grid = (long long)num_sms * per_sm;
    const long long max_useful = (long long)((p.n_paths + kCtaBlock - 1ull) / kCtaBlock);
    if (grid > max_useful) grid = max_useful;
    if (grid < 1) grid = 1;
    cta_wavefront_kernel<<<(unsigned)grid, kCtaBlock, smem, stream>>>(p);
    if (info) {
        info->grid = (int)grid;
        info->block = kCtaBlock;
        info->smem_bytes = smem;
        info->blocks_per_sm = per_sm;
        info->launches = 1;
        info->rays_per_lane = 1;
        info->sweep = kSweepPacked;
    }
    return cudaGetLastError();
}

}  // namespace rtw
