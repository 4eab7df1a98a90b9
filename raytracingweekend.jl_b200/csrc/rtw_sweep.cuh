// rtw_sweep.cuh -- the closest-hit sweep over one shared-memory tile of the sphere list (src/hit.jl:38-50, 12-35)
// and the bulk-TMA / mbarrier helpers that stage the tiles.  Shared by the fused persistent kernel
// (rtw_kernels.cu) and the split wavefront kernels (rtw_wavefront.cu).  Everything is static (internal linkage).
#pragma once
#include "rtw_kernels.h"

namespace rtw {
namespace {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kTraceBlock = 256;
constexpr unsigned kPoolChunk = 128;  // path tickets a warp takes from the global counter at a time

// ---- 1-D bulk TMA (cp.async.bulk, SASS UBLKCP) + mbarrier helpers ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes,
                                             unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---- closest-hit sweep over one shared-memory tile of the sphere list (src/hit.jl:38-50) -------------------
// Every lane traces R independent paths ("slots"); one broadcast LDS.128 of a sphere feeds R tests.

// RTW_SWEEP_BRANCH: test, and select the root at once under a (rare, divergent) branch.
template <int R>
__device__ __forceinline__ void sweep_tile_branch(const float4* __restrict__ tile, uint32_t count, uint32_t k_base,
                                                  const f3 (&o)[R], const f3 (&d)[R], const bool (&alive)[R],
                                                  float (&best_t)[R], int (&best_k)[R]) {
    const float tmin = 1e-4f;  // T(1e-4), src/ray_color.jl:19
#pragma unroll 4
    for (uint32_t k = 0; k < count; ++k) {
        float4 s = tile[k];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float hb;
            float disc = sphere_disc(s, o[r], d[r], hb);
            if (!(disc < 0.0f) && alive[r]) {  // src/hit.jl:19
                if (sphere_accept(disc, hb, tmin, best_t[r])) best_k[r] = (int)(k_base + k);
            }
        }
    }
}

// RTW_SWEEP_MASK: the inner loop is branch-free -- per test 11 FP32 instructions + one funnel shift that
// pushes the sign bit of the discriminant (set = miss, src/hit.jl:19) into a per-slot 32-test mask.  Masks go
// to shared memory once per 32 tests; after the tile each lane walks only its own candidates (in list order,
// so ties still go to the later sphere) and redoes the identical arithmetic to select the root.
// (A NaN discriminant -- only reachable with non-finite scene/camera values -- counts as a miss when its
// sign bit is set; the reference would treat it as a hit.)
template <int R, int kBlock>
__device__ __forceinline__ void sweep_tile_mask(const float4* __restrict__ tile, uint32_t count, uint32_t k_base,
                                                uint32_t* __restrict__ s_mask, const f3 (&o)[R], const f3 (&d)[R],
                                                const bool (&alive)[R], float (&best_t)[R], int (&best_k)[R]) {
    const float tmin = 1e-4f;
    const uint32_t nchunks = (count + 31u) >> 5;
    uint32_t summary[R];
#pragma unroll
    for (int r = 0; r < R; ++r) summary[r] = 0u;
    for (uint32_t c = 0; c < nchunks; ++c) {
        const float4* ch = tile + c * 32u;
        uint32_t m[R];
#pragma unroll
        for (int r = 0; r < R; ++r) m[r] = 0u;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float4 s = ch[j];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float hb;
                float disc = sphere_disc(s, o[r], d[r], hb);
                m[r] = __funnelshift_l(__float_as_uint(disc), m[r], 1);  // test j of the chunk ends at bit 31-j
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            s_mask[(c * R + r) * kBlock] = m[r];
            summary[r] |= (m[r] != 0xffffffffu ? 1u : 0u) << c;
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (!alive[r]) continue;
        uint32_t sum = summary[r], cand = 0u, c = 0u;
        for (;;) {
            if (cand == 0u) {
                if (sum == 0u) break;
                c = (uint32_t)__ffs((int)sum) - 1u;
                sum &= sum - 1u;
                cand = ~s_mask[(c * R + r) * kBlock];
                uint32_t valid = count - c * 32u;  // entries of the last chunk beyond `count` are padding
                if (valid < 32u) cand &= 0xffffffffu << (32u - valid);
                if (cand == 0u) continue;
            }
            uint32_t j = (uint32_t)__clz((int)cand);
            cand &= ~(0x80000000u >> j);
            uint32_t kl = c * 32u + j;
            float4 s = tile[kl];
            float hb;
            float disc = sphere_disc(s, o[r], d[r], hb);  // bit-identical to the value computed in the sweep
            if (sphere_accept(disc, hb, tmin, best_t[r])) best_k[r] = (int)(k_base + kl);
        }
    }
}

// RTW_SWEEP_PACKED: the mask sweep on Blackwell's packed FP32x2 pipe.  Two spheres are tested per instruction
// (FADD2/FMUL2/FFMA2, IEEE rn per half => bit-identical to the scalar form); the ray components are broadcast
// operands.  A packed instruction keeps the FP32 pipe busy for two cycles but takes one issue slot, so the
// LDS.128 / funnel-shift / loop instructions issue in the shadow of the arithmetic: the loop is bound by the
// FP32 pipe (11 lane-ops per test), not by instruction issue.
// Shared-memory layout ("pair layout"): for spheres (a,b) = (2p, 2p+1): {xa,xb,ya,yb} {za,zb,ra,rb}.
__device__ __forceinline__ float2 dup2(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 neg2(float a, float b) { return make_float2(-a, -b); }

__device__ __forceinline__ float4 pair_layout_fetch(const float4* __restrict__ tile, uint32_t kl) {
    const float* f = reinterpret_cast<const float*>(tile + (kl >> 1) * 2u) + (kl & 1u);
    return make_float4(f[0], f[2], f[4], f[6]);
}

// Lane cooperation (kCoop = 1, 2 or 4): an LDS.128 delivers 512 B to the warp and the shared-memory pipe moves
// 128 B/clk/SM, so one broadcast load per 2 tests caps the loop near 70 % of the FP32 pipe.  With kCoop > 1 the
// kCoop lanes of an aligned group exchange their rays by shuffle; each lane then tests ALL kCoop rays of its group
// (slots) against every kCoop-th sphere pair, so one loaded pair feeds 2*kCoop tests.  Results are merged back by
// shuffle with the list-order tie rule (equal t => later sphere, src/hit.jl:24-26,44-46).
// Sphere pair P of a "super-chunk" (16*kCoop pairs = 32*kCoop spheres) is tested by lane h = P mod kCoop as its
// i-th pair, i = P div kCoop; bit (31 - 2i - half) of the lane's mask word for that super-chunk.
// one sphere pair against the NS slots of a lane: 11 packed FP32 instructions + 2 funnel shifts per slot
template <int NS>
__device__ __forceinline__ void test_pair_packed(const float4 A, const float4 B, const f3 (&o)[NS], const f3 (&d)[NS],
                                                 uint32_t (&m)[NS]) {
#pragma unroll
    for (int r = 0; r < NS; ++r) {
        // oc = o - c (src/hit.jl:13) for both spheres of the pair
        const float2 ocx = __fadd2_rn(dup2(o[r].x), neg2(A.x, A.y));
        const float2 ocy = __fadd2_rn(dup2(o[r].y), neg2(A.z, A.w));
        const float2 ocz = __fadd2_rn(dup2(o[r].z), neg2(B.x, B.y));
        // half_b = oc . d (src/hit.jl:16), dot = fma(z,z, fma(y,y, x*x))
        const float2 hb = __ffma2_rn(ocz, dup2(d[r].z), __ffma2_rn(ocy, dup2(d[r].y), __fmul2_rn(ocx, dup2(d[r].x))));
        // c = oc . oc - radius^2 (src/hit.jl:17)
        const float2 q = __ffma2_rn(ocz, ocz, __ffma2_rn(ocy, ocy, __fmul2_rn(ocx, ocx)));
        const float2 rr = make_float2(B.z, B.w);
        const float2 cq = __ffma2_rn(neg2(rr.x, rr.y), rr, q);
        // discriminant = half_b^2 - c (src/hit.jl:18)
        const float2 disc = __ffma2_rn(hb, hb, neg2(cq.x, cq.y));
        m[r] = __funnelshift_l(__float_as_uint(disc.x), m[r], 1);  // pair i of the chunk ends at bits 31-2i, 30-2i
        m[r] = __funnelshift_l(__float_as_uint(disc.y), m[r], 1);
    }
}

// The branch-free part of the packed sweep: fills the per-slot candidate masks of this lane in shared memory
// (s_mask[(c*NS + r)*kBlock], super-chunk c, slot r) and returns one summary bit per non-empty mask word.
template <int NS, int kCoop, int kBlock>
__device__ __forceinline__ void sweep_masks_packed(const float4* __restrict__ tile, uint32_t count, uint32_t coop_h,
                                                   uint32_t* __restrict__ s_mask, const f3 (&o)[NS], const f3 (&d)[NS],
                                                   uint32_t (&summary)[NS]) {
    constexpr uint32_t kSuper = 32u * kCoop;   // spheres per super-chunk
    constexpr uint32_t kSuperPairs = 16u * kCoop;
    const uint32_t npairs = (count + 1u) >> 1;
    const uint32_t nsc = (count + kSuper - 1u) / kSuper;
#pragma unroll
    for (int r = 0; r < NS; ++r) summary[r] = 0u;
    const float4* lane_base = tile + coop_h * 2u;  // this lane's first pair of each super-chunk
    for (uint32_t c = 0; c < nsc; ++c) {
        const float4* ch = lane_base + c * kSuper;  // kSuper spheres = kSuper float4 of pair layout
        uint32_t m[NS];
#pragma unroll
        for (int r = 0; r < NS; ++r) m[r] = 0u;
        const uint32_t pairs_here = npairs - c * kSuperPairs;  // pairs left from this super-chunk on
        if (pairs_here > kSuperPairs || (pairs_here == kSuperPairs && (count & 1u) == 0u)) {
#pragma unroll
            for (int i = 0; i < 16; ++i) test_pair_packed<NS>(ch[2 * kCoop * i], ch[2 * kCoop * i + 1], o, d, m);
        } else {
            // last super-chunk, ragged or ending in the zero pad partner of an odd last sphere: only the pairs that
            // exist (no padded arithmetic); left-align the mask and mark the missing tests as misses
            const uint32_t mine = pairs_here > coop_h ? (pairs_here - coop_h + kCoop - 1u) / kCoop : 0u;
#pragma unroll 1
            for (uint32_t i = 0; i < mine; ++i) test_pair_packed<NS>(ch[2 * kCoop * i], ch[2 * kCoop * i + 1], o, d, m);
            const uint32_t sh = 32u - 2u * mine;  // 0..32
            // the pad partner is the second half of the last pair: the last test of the lane that holds that pair
            const uint32_t pad = ((count & 1u) != 0u && coop_h == (pairs_here - 1u) % kCoop) ? (1u << sh) : 0u;
#pragma unroll
            for (int r = 0; r < NS; ++r) m[r] = sh >= 32u ? 0xffffffffu : ((m[r] << sh) | ((1u << sh) - 1u) | pad);
        }
#pragma unroll
        for (int r = 0; r < NS; ++r) {
            s_mask[(c * NS + r) * kBlock] = m[r];
            summary[r] |= (m[r] != 0xffffffffu ? 1u : 0u) << c;
        }
    }
}

template <int NS, int kCoop, int kBlock>
__device__ __forceinline__ void sweep_tile_packed(const float4* __restrict__ tile, const float4* __restrict__ aos,
                                                  uint32_t count, uint32_t k_base, uint32_t coop_h,
                                                  uint32_t* __restrict__ s_mask, const f3 (&o)[NS], const f3 (&d)[NS],
                                                  const bool (&alive)[NS], float (&best_t)[NS], int (&best_k)[NS]) {
    const float tmin = 1e-4f;
    constexpr uint32_t kSuper = 32u * kCoop;
    uint32_t summary[NS];
    sweep_masks_packed<NS, kCoop, kBlock>(tile, count, coop_h, s_mask, o, d, summary);
    // ---- candidate resolution: each lane walks its own candidates, slot by slot, in list order so that ties still
    // go to the later sphere (src/hit.jl:44-46)
#pragma unroll
    for (int r = 0; r < NS; ++r) {
        if (!alive[r]) continue;
        uint32_t sum = summary[r], cand = 0u, c = 0u;
        for (;;) {
            if (cand == 0u) {
                if (sum == 0u) break;
                c = (uint32_t)__ffs((int)sum) - 1u;
                sum &= sum - 1u;
                cand = ~s_mask[(c * NS + r) * kBlock];
            }
            const uint32_t j = (uint32_t)__clz((int)cand);
            cand &= ~(0x80000000u >> j);
            const uint32_t kl = c * kSuper + 2u * ((j >> 1) * kCoop + coop_h) + (j & 1u);
            if (kl >= count) continue;  // the zero pad partner of an odd last sphere
            const float4 s = aos ? aos[kl] : pair_layout_fetch(tile, kl);
            // scalar redo of src/hit.jl:13-18: bit-identical to the packed values
            const f3 oc = mk3(o[r].x - s.x, o[r].y - s.y, o[r].z - s.z);
            const float hb = dot3(oc, d[r]);
            const float cq = fmaf(-s.w, s.w, dot3(oc, oc));
            // Sphere entirely behind the origin (half_b > 0 and origin outside): sqrt(disc) <= half_b in IEEE
            // arithmetic, so both roots are <= 0 < tmin and src/hit.jl:24-28 rejects them -- skip the square root.
            if (hb > 0.0f && cq > 0.0f) continue;
            const float disc = fmaf(hb, hb, -cq);
            if (sphere_accept(disc, hb, tmin, best_t[r])) best_k[r] = (int)(k_base + kl);
        }
    }
}

// One tile for the R slots of a lane, through the sweep variant SWEEP; with kCoop > 1 (packed sweep, R == 1) the
// rays of the lane group are exchanged first and the per-lane partial results merged afterwards.
template <int R, int SWEEP, int kCoop, int kBlock>
__device__ __forceinline__ void sweep_tile(const float4* __restrict__ tile, const float4* __restrict__ aos,
                                           uint32_t count, uint32_t k_base, uint32_t* __restrict__ s_mask,
                                           const f3 (&o)[R], const f3 (&d)[R], const bool (&alive)[R],
                                           float (&best_t)[R], int (&best_k)[R]) {
    if constexpr (SWEEP == kSweepBranch) {
        sweep_tile_branch<R>(tile, count, k_base, o, d, alive, best_t, best_k);
    } else if constexpr (SWEEP == kSweepMask) {
        sweep_tile_mask<R, kBlock>(tile, count, k_base, s_mask, o, d, alive, best_t, best_k);
    } else if constexpr (kCoop == 1) {
        sweep_tile_packed<R, 1, kBlock>(tile, aos, count, k_base, 0u, s_mask, o, d, alive, best_t, best_k);
    } else {
        static_assert(R == 1 || kCoop == 1, "lane cooperation is implemented for one path per lane");
        const uint32_t h = threadIdx.x & (kCoop - 1);
        f3 so[kCoop], sd[kCoop];
        bool sa[kCoop];
        float bt[kCoop];
        int bk[kCoop];
#pragma unroll
        for (int q = 0; q < kCoop; ++q) {  // slot q holds the ray of lane (lane ^ q)
            so[q] = mk3(__shfl_xor_sync(kFullMask, o[0].x, q), __shfl_xor_sync(kFullMask, o[0].y, q),
                        __shfl_xor_sync(kFullMask, o[0].z, q));
            sd[q] = mk3(__shfl_xor_sync(kFullMask, d[0].x, q), __shfl_xor_sync(kFullMask, d[0].y, q),
                        __shfl_xor_sync(kFullMask, d[0].z, q));
            sa[q] = __shfl_xor_sync(kFullMask, alive[0] ? 1 : 0, q) != 0;
            bt[q] = __int_as_float(0x7f800000);
            bk[q] = -1;
        }
        sweep_tile_packed<kCoop, kCoop, kBlock>(tile, aos, count, k_base, h, s_mask, so, sd, sa, bt, bk);
#pragma unroll
        for (int q = 0; q < kCoop; ++q) {  // lane ^ q holds, in ITS slot q, the partial result for my ray
            const float pt = __shfl_xor_sync(kFullMask, bt[q], q);
            const int pk = __shfl_xor_sync(kFullMask, bk[q], q);
            if (pk >= 0 && (best_k[0] < 0 || pt < best_t[0] || (pt == best_t[0] && pk > best_k[0]))) {
                best_t[0] = pt;
                best_k[0] = pk;
            }
        }
    }
}

}  // namespace
}  // namespace rtw
