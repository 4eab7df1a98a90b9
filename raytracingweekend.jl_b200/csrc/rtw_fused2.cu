// rtw_fused2.cu -- RTW_TAIL_UNIFIED: the persistent fused trace kernel with a warp-converged tail.
//
// Same path, same bits as fused_trace_kernel (rtw_kernels.cu); what differs is how the work AFTER the sphere-list
// sweep is laid out.  The roofline of this kernel is FP32 issue slots, so every warp instruction that is not one
// of the 11 FP32 lane-ops of a ray-sphere test costs wall time at the rate of one slot per instruction, whatever
// the number of lanes it serves.  After the sweep a warp holds lanes in different states (path ended / scatters
// on Lambertian, Metal, Dielectric / needs a new primary ray), and the first kernel ran each state's code at
// 8-10 of 32 lanes.  Here the states share their expensive pieces:
//   * the lanes whose path ended take their next path ticket BEFORE the random draws, so that ONE Philox block 0
//     serves the scatter of the continuing lanes (src/material.jl) and the primary ray of the new lanes
//     (src/render.jl:30-37, src/camera.jl:43-48);
//   * the rejection loops (src/rand.jl:15-22 ball, :31-38 disk) run warp-cooperatively: the lanes that still need
//     a sample publish their stream address in shared memory and ALL 32 lanes evaluate the following attempts for
//     them (the stream is addressed, not sequential, so attempt a of a path can be computed by any lane); a lane
//     takes the first accepted attempt in stream order, so the result is the same as the sequential loop's;
//   * one normalize() instance serves unit(ball sample) and the primary-ray direction, a second one serves the
//     scattered directions of all three materials;
//   * u = T(j/W), v = T((H-i)/H) (src/render.jl:26-27) come from tables filled once per render, ticket -> (pixel,
//     sample) uses multiply-shift division;
//   * candidate resolution after the sweep reads the spheres from a copy laid out in the order a lane meets them
//     (one LEA + LDS.128 per candidate), keeps (t, position) only and decodes the list index once.
#include "rtw_kernels.h"
#include "rtw_sweep.cuh"
#include "rtw_grid.cuh"

namespace rtw {

namespace {

__device__ __forceinline__ uint32_t magic_div(uint32_t n, const MagicDiv k) {
    const uint32_t t = __umulhi(k.m, n);
    return (t + ((n - t) >> k.sh1)) >> k.sh2;
}

__device__ __forceinline__ uint32_t bfind_u32(uint32_t x) {  // position of the highest set bit (x != 0)
    uint32_t r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
    return r;
}

// ((1 << width) - 1), all ones for width >= 32
__device__ __forceinline__ uint32_t low_mask(uint32_t width) {
    uint32_t r;
    asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(r) : "r"(0u), "r"(width));
    return r;
}

// Philox4x32-7 (kPhiloxRounds, rtw_device.cuh) with the round keys read from the kernel parameter block (constant bank operands) instead of
// being re-derived (k += W per round) by every evaluation; same function as philox_block (rtw_device.cuh)
__device__ __forceinline__ u32x4 philox_block_rk(const TraceParams& P, uint32_t sample, uint32_t pixel, uint32_t event,
                                                 uint32_t block) {
    uint32_t c0 = block, c1 = sample, c2 = pixel, c3 = event;
#pragma unroll
    for (int r = 0; r < kPhiloxRounds; ++r) {
        const uint32_t hi0 = __umulhi(kPhiloxM0, c0), lo0 = kPhiloxM0 * c0;
        const uint32_t hi1 = __umulhi(kPhiloxM1, c2), lo1 = kPhiloxM1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ P.rk[2 * r];
        const uint32_t n2 = hi0 ^ c3 ^ P.rk[2 * r + 1];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    return u32x4{c0, c1, c2, c3};
}

// cooperative rejection sampling: with `cnt` lanes in need, each gets per = 32 / cnt helper lanes;
// entry = per | (ceil(256 / per) << 8), so that lane / per = (lane * (entry >> 8)) >> 8 for lane < 32
__constant__ uint32_t c_coop_tab[33] = {
    0u,
    32u | (8u << 8), 16u | (16u << 8), 10u | (26u << 8), 8u | (32u << 8), 6u | (43u << 8), 5u | (52u << 8),
    4u | (64u << 8), 4u | (64u << 8), 3u | (86u << 8), 3u | (86u << 8), 2u | (128u << 8), 2u | (128u << 8),
    2u | (128u << 8), 2u | (128u << 8), 2u | (128u << 8), 2u | (128u << 8), 1u | (256u << 8), 1u | (256u << 8),
    1u | (256u << 8), 1u | (256u << 8), 1u | (256u << 8), 1u | (256u << 8), 1u | (256u << 8), 1u | (256u << 8),
    1u | (256u << 8), 1u | (256u << 8), 1u | (256u << 8), 1u | (256u << 8), 1u | (256u << 8), 1u | (256u << 8),
    1u | (256u << 8), 1u | (256u << 8)};

// random_between(-1, 1) of the word's uniform (src/rand.jl:24): 2*(k*2^-23) - 1 = k*2^-22 - 1, exact either way
// Two instructions instead of shift + convert + fma: the 23 bits become the mantissa of a float in [2, 4), 2 + k*2^-22, and
// subtracting 3 is exact (the result is a multiple of 2^-22 below 1 in magnitude).  u01x: the same for k*2^-23 via [1, 2).
__device__ __forceinline__ float pm1x(uint32_t w) { return __uint_as_float(__funnelshift_r(w, 0x80u, 9)) - 3.0f; }
__device__ __forceinline__ float u01x(uint32_t w) { return __uint_as_float(__funnelshift_r(w, 0x7fu, 9)) - 1.0f; }

// ---- candidate resolution over the permuted AoS copy ----------------------------------------------------------
// aos_perm[(c*kCoop + h)*32 + j] = the sphere that lane h of a group tests as its j-th test of super-chunk c, i.e.
// list index c*32*kCoop + 2*((j>>1)*kCoop + h) + (j&1).  Mask bit (31 - j) of word (c, slot) clear = candidate.
// Candidates are visited in list order per slot, so ties in t go to the later sphere (src/hit.jl:24-26,44-46).
__device__ __forceinline__ float4 lds128(uint32_t addr) {  // 32-bit shared-window address
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");  // ordered after this thread's mask stores
    return v;
}

template <int NS, int kCoop, int kBlock>
__device__ __forceinline__ void walk_candidates_perm(const float4* __restrict__ aos_perm, uint32_t coop_h,
                                                     const uint32_t* __restrict__ s_mask, const f3 (&o)[NS],
                                                     const f3 (&d)[NS], const bool (&alive)[NS],
                                                     const uint32_t (&summary)[NS], float (&best_t)[NS],
                                                     int (&best_k)[NS]) {
    const float tmin = 1e-4f;  // T(1e-4), src/ray_color.jl:19
    // 32-bit shared addresses: of entry 31 of this lane's run in super-chunk 0, and of this lane's mask word 0
    const uint32_t a_lane = smem_u32(aos_perm) + (coop_h * 32u + 31u) * 16u;
    const uint32_t m_lane = smem_u32(s_mask);
#pragma unroll
    for (int r = 0; r < NS; ++r) {
        uint32_t sum = alive[r] ? summary[r] : 0u;
        uint32_t cand = 0u;
        uint32_t a31 = 0u;   // address of test j = 0's partner at bit 31: test j = 31 - p sits at a31 - 16*p
        uint32_t code31 = 0u;  // (c*kCoop + h)*32 + 31: position code of bit p is code31 - p
        float bt = __int_as_float(0x7f800000);
        uint32_t bcode = 0xffffffffu;
        const float ox = o[r].x, oy = o[r].y, oz = o[r].z, dx = d[r].x, dy = d[r].y, dz = d[r].z;
        for (;;) {
            if (cand == 0u) {
                if (sum == 0u) break;
                const uint32_t c = (uint32_t)__ffs((int)sum) - 1u;
                sum &= sum - 1u;
                cand = ~lds32(m_lane + (c * NS + r) * (kBlock * 4u));  // != 0: summary bits mark words with a candidate
                a31 = a_lane + c * (kCoop * 32u * 16u);
                code31 = (c * kCoop + coop_h) * 32u + 31u;
            }
            const uint32_t p = bfind_u32(cand);
            cand &= low_mask(p);  // p is the highest set bit
            const float4 s = lds128(a31 - 16u * p);
            // scalar redo of src/hit.jl:13-18: bit-identical to the packed values of the sweep
            const float ocx = ox - s.x, ocy = oy - s.y, ocz = oz - s.z;
            const float hb = fmaf(ocz, dz, fmaf(ocy, dy, ocx * dx));
            const float cq = fmaf(-s.w, s.w, fmaf(ocz, ocz, fmaf(ocy, ocy, ocx * ocx)));
            // Sphere entirely behind the origin (half_b > 0 and origin outside): sqrt(disc) <= half_b in IEEE
            // arithmetic, so both roots are <= 0 < tmin and src/hit.jl:24-28 rejects them -- skip the square root.
            if (hb > 0.0f && cq > 0.0f) continue;
            const float sq = __fsqrt_rn(fmaf(hb, hb, -cq));
            const float r1 = -hb - sq, r2 = -hb + sq;            // src/hit.jl:23, 25
            const bool bad1 = r1 < tmin || bt < r1;              // src/hit.jl:24
            const bool bad2 = r2 < tmin || bt < r2;              // src/hit.jl:26
            if (!(bad1 && bad2)) {
                bt = bad1 ? r2 : r1;
                bcode = code31 - p;  // (c*kCoop + h)*32 + j
            }
        }
        best_t[r] = bt;
        if (bcode != 0xffffffffu) {
            const uint32_t j = bcode & 31u, cq2 = bcode >> 5;  // cq2 = c*kCoop + h
            const uint32_t c = cq2 / (uint32_t)kCoop;
            best_k[r] = (int)(c * (32u * kCoop) + 2u * ((j >> 1) * kCoop + coop_h) + (j & 1u));
        } else {
            best_k[r] = -1;
        }
    }
}

// The branch-free part of the packed sweep (same tests, same mask words as sweep_masks_packed in rtw_sweep.cuh),
// addressed through 32-bit shared-window pointers that advance per super-chunk: `tile_lane` = address of this
// lane's first pair of super-chunk 0, `mask_lane` = address of this lane's mask word 0.
template <int NS, int kCoop, int kBlock>
__device__ __forceinline__ void sweep_masks_packed2(uint32_t tile_lane, uint32_t count, uint32_t coop_h,
                                                    uint32_t mask_lane, const f3 (&o)[NS], const f3 (&d)[NS],
                                                    uint32_t (&summary)[NS]) {
    constexpr uint32_t kSuper = 32u * kCoop;  // spheres per super-chunk = float4 entries of pair layout
    constexpr uint32_t kSuperPairs = 16u * kCoop;
    const uint32_t npairs = (count + 1u) >> 1;
    // super-chunks that are complete and do not end in the zero pad partner of an odd last sphere
    const uint32_t nfull = (count & 1u) ? (npairs - 1u) / kSuperPairs : npairs / kSuperPairs;
#pragma unroll
    for (int r = 0; r < NS; ++r) summary[r] = 0u;
    uint32_t a = tile_lane, ma = mask_lane, bit = 1u;
    for (uint32_t c = 0; c < nfull; ++c) {
        uint32_t m[NS];
#pragma unroll
        for (int r = 0; r < NS; ++r) m[r] = 0u;
#pragma unroll
        for (int i = 0; i < 16; ++i)
            test_pair_packed<NS>(lds128(a + (uint32_t)i * (kCoop * 32u)), lds128(a + (uint32_t)i * (kCoop * 32u) + 16u), o, d, m);
#pragma unroll
        for (int r = 0; r < NS; ++r) {
            sts32(ma + (uint32_t)r * (kBlock * 4u), m[r]);
            summary[r] |= m[r] != 0xffffffffu ? bit : 0u;
        }
        a += kSuper * 16u;
        ma += NS * kBlock * 4u;
        bit <<= 1;
    }
    const uint32_t pairs_here = npairs - nfull * kSuperPairs;  // 0 .. kSuperPairs
    if (pairs_here != 0u) {
        // last super-chunk, ragged or ending in the pad: only the pairs that exist (no padded arithmetic);
        // left-align the mask and mark the missing tests (and the pad) as misses
        uint32_t m[NS];
#pragma unroll
        for (int r = 0; r < NS; ++r) m[r] = 0u;
        const uint32_t mine = pairs_here > coop_h ? (pairs_here - coop_h + kCoop - 1u) / kCoop : 0u;
#pragma unroll 1
        for (uint32_t i = 0; i < mine; ++i) {
            test_pair_packed<NS>(lds128(a), lds128(a + 16u), o, d, m);
            a += kCoop * 32u;
        }
        const uint32_t sh = 32u - 2u * mine;  // 0..32
        const uint32_t pad = ((count & 1u) != 0u && coop_h == (pairs_here - 1u) % kCoop) ? (1u << sh) : 0u;
        const uint32_t fill = low_mask(sh) | pad;
#pragma unroll
        for (int r = 0; r < NS; ++r) {
            m[r] = sh >= 32u ? 0xffffffffu : ((m[r] << sh) | fill);
            sts32(ma + (uint32_t)r * (kBlock * 4u), m[r]);
            summary[r] |= m[r] != 0xffffffffu ? bit : 0u;
        }
    }
}

// The sweep of the own-ray walk below: same tests and same mask words as sweep_masks_packed2, but
//   * ONE summary word per lane, bit w = c*kCoop + q set when mask word (super-chunk c, slot q) holds a candidate -- the
//     order in which the words sit in shared memory, so the walk derives both addresses from w with shifts;
//   * the ragged last super-chunk runs the same number of pairs on every lane of a group (ceil(pairs / kCoop), the
//     tile is zero-padded to a whole super-chunk), unrolled by 4; tests of pairs that do not exist are forced to "miss"
//     when the mask word is stored.  Lists <= 32 * 32 spheres (kCoop * n_super_chunks <= 32 summary bits).
// what the sweep of a list needs besides the tile: loop-invariant over the bounces, computed once per kernel
struct SweepPlan {
    uint32_t nfull;  // whole super-chunks
    uint32_t run;    // pairs every lane runs in the ragged last super-chunk (0: there is none)
    uint32_t sh;     // left-alignment of its mask words
    uint32_t fill;   // bits of its mask words that are forced to "miss" for this lane
};

template <int kCoop>
__device__ __forceinline__ SweepPlan make_sweep_plan(uint32_t count, uint32_t coop_h) {
    constexpr uint32_t kSuperPairs = 16u * kCoop;
    const uint32_t npairs = (count + 1u) >> 1;
    SweepPlan P;
    // super-chunks that are complete and do not end in the zero pad partner of an odd last sphere
    P.nfull = (count & 1u) ? (npairs - 1u) / kSuperPairs : npairs / kSuperPairs;
    const uint32_t pairs_here = npairs - P.nfull * kSuperPairs;  // 0 .. kSuperPairs
    P.run = (pairs_here + kCoop - 1u) / kCoop;                  // uniform over the lanes of a group
    // this lane's real tests: pairs coop_h, coop_h + kCoop, ... < pairs_here; the last of them may end in the pad
    // partner of an odd last sphere.  Everything after them (zero pairs of the padding, the pad partner) = miss.
    const uint32_t mine = pairs_here > coop_h ? (pairs_here - coop_h + kCoop - 1u) / kCoop : 0u;
    const uint32_t real = 2u * mine - (((count & 1u) != 0u && mine != 0u && coop_h == (pairs_here - 1u) % kCoop) ? 1u : 0u);
    P.sh = 32u - 2u * P.run;           // left-align the `2 * run` tests that were executed (run >= 1 where it is used)
    P.fill = low_mask(32u - real);     // bits below the `real` top ones
    return P;
}

template <int kCoop, int kBlock>
__device__ __forceinline__ uint32_t sweep_masks_packed3(uint32_t tile_lane, const SweepPlan plan, uint32_t mask_lane,
                                                        const f3 (&o)[kCoop], const f3 (&d)[kCoop]) {
    const uint32_t nfull = plan.nfull;
    uint32_t summary = 0u;
    uint32_t a = tile_lane, ma = mask_lane, bit = 1u;
    for (uint32_t c = 0; c < nfull; ++c) {
        uint32_t m[kCoop];
#pragma unroll
        for (int r = 0; r < kCoop; ++r) m[r] = 0u;
        // coop 4: 8 pairs per unrolled body (the 16-pair body of 4 rays is ~15 KB of code), coop 2: all 16
        constexpr int kBody = kCoop == 4 ? 8 : 16;
#pragma unroll 1
        for (int half = 0; half < 16 / kBody; ++half) {
#pragma unroll
            for (int i = 0; i < kBody; ++i)
                test_pair_packed<kCoop>(lds128(a + (uint32_t)i * (kCoop * 32u)), lds128(a + (uint32_t)i * (kCoop * 32u) + 16u), o, d, m);
            a += kBody * kCoop * 32u;
        }
#pragma unroll
        for (int r = 0; r < kCoop; ++r) {
            sts32(ma + (uint32_t)r * (kBlock * 4u), m[r]);
            summary |= m[r] != 0xffffffffu ? (bit << r) : 0u;
        }
        ma += kCoop * kBlock * 4u;
        bit <<= kCoop;
    }
    const uint32_t run = plan.run;  // pairs every lane runs in the ragged last super-chunk (uniform), 0 .. 16
    if (run != 0u) {
        uint32_t m[kCoop];
#pragma unroll
        for (int r = 0; r < kCoop; ++r) m[r] = 0u;
        uint32_t i = 0;
        for (; i + 4u <= run; i += 4u) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                test_pair_packed<kCoop>(lds128(a + (uint32_t)u * (kCoop * 32u)), lds128(a + (uint32_t)u * (kCoop * 32u) + 16u), o, d, m);
            a += 4u * kCoop * 32u;
        }
        for (; i < run; ++i) {
            test_pair_packed<kCoop>(lds128(a), lds128(a + 16u), o, d, m);
            a += kCoop * 32u;
        }
#pragma unroll
        for (int r = 0; r < kCoop; ++r) {
            m[r] = (m[r] << plan.sh) | plan.fill;
            sts32(ma + (uint32_t)r * (kBlock * 4u), m[r]);
            summary |= m[r] != 0xffffffffu ? (bit << r) : 0u;
        }
    }
    return summary;
}

// Candidate resolution, transposed: every lane walks ALL candidates of ITS OWN ray -- its own slot-0 mask words and
// the slot-q words of its group partners lane^q, read straight from their shared-memory slices -- in one loop with
// the ray in registers; no per-slot loops, no merge of partial results.  The candidates of a ray then arrive out of
// list order (the partners' sphere subsets interleave), so the closest hit is taken in its order-independent form:
// the root a sphere offers is the first one >= tmin (src/hit.jl:23-28: near root, else far root -- which one does
// not depend on the running closest t, because far >= near), the winner is the smallest such root and equal roots go
// to the larger list index (src/hit.jl:24,26 are inclusive, so the sequential sweep lets the later sphere win).
__device__ __forceinline__ uint32_t perm_code_to_index(uint32_t code, uint32_t coop) {
    // code = (c*coop + h)*32 + j  ->  list index c*32*coop + 2*((j>>1)*coop + h) + (j&1)
    const uint32_t j = code & 31u, ch = code >> 5;
    const uint32_t c = ch / coop, h = ch - c * coop;
    return c * (32u * coop) + 2u * ((j >> 1) * coop + h) + (j & 1u);
}

template <int kCoop, int kBlock>
__device__ __forceinline__ void walk_own_ray_perm(const float4* __restrict__ aos_perm, const uint32_t* s_mask_base,
                                                  const f3 o, const f3 d, const bool alive, const uint32_t summary,
                                                  float& best_t, int& best_k) {
    const float tmin = 1e-4f;  // T(1e-4), src/ray_color.jl:19
    const uint32_t tid = threadIdx.x, h = tid & (kCoop - 1);
    __syncwarp();  // the partners' mask words are visible
    // summary bit w = c*kCoop + q of lane L says: L's mask word (c, slot q) -- the tests of the ray of lane L^q -- holds a
    // candidate.  The words about MY ray are word (c, q) of lane L^q for q = 0 .. kCoop-1: pick, from each partner's
    // summary, the bits of its slot q.
    constexpr uint32_t kSlot0 = kCoop == 2 ? 0x55555555u : 0x11111111u;  // bits with w mod kCoop == 0
    uint32_t sum = summary & kSlot0;
#pragma unroll
    for (int q = 1; q < kCoop; ++q) sum |= __shfl_xor_sync(kFullMask, summary, q) & (kSlot0 << q);
    if (!alive) sum = 0u;
    const uint32_t a_base = smem_u32(aos_perm) + 31u * 16u, m_base = smem_u32(s_mask_base) + tid * 4u;
    uint32_t cand = 0u, a31 = 0u, code31 = 0u;
    float bt = __int_as_float(0x7f800000);
    uint32_t bcode = 0xffffffffu;
    for (;;) {
        if (cand == 0u) {
            if (sum == 0u) break;
            const uint32_t w = (uint32_t)__ffs((int)sum) - 1u;  // = c*kCoop + q
            sum &= sum - 1u;
            // lane^q tested my ray in its slot q: word index w of thread tid ^ q, q = w mod kCoop
            cand = ~lds32((m_base + w * (kBlock * 4u)) ^ ((w & (kCoop - 1u)) << 2));
            // it tested the spheres of ITS run: position code (c*kCoop + (h^q))*32 + j = ((w ^ h) << 5) + j
            code31 = ((w ^ h) << 5) + 31u;
            a31 = a_base + ((w ^ h) << 9);  // address of entry j = 31 of that run; test j = 31 - p sits 16*p below
        }
        const uint32_t p = bfind_u32(cand);
        cand &= low_mask(p);
        const float4 s = lds128(a31 - 16u * p);
        // scalar redo of src/hit.jl:13-18: bit-identical to the packed values of the sweep
        const float ocx = o.x - s.x, ocy = o.y - s.y, ocz = o.z - s.z;
        const float hb = fmaf(ocz, d.z, fmaf(ocy, d.y, ocx * d.x));
        const float cq = fmaf(-s.w, s.w, fmaf(ocz, ocz, fmaf(ocy, ocy, ocx * ocx)));
        // branch-free: a sphere wholly behind the origin has both roots <= 0 < tmin and falls out at `t >= tmin`
        const float sq = __fsqrt_rn(fmaf(hb, hb, -cq));
        const float r1 = -hb - sq, r2 = -hb + sq;  // src/hit.jl:23, 25
        const float t = r1 < tmin ? r2 : r1;       // the first root >= tmin, if any
        const uint32_t code = code31 - p;
        if (t == bt && t >= tmin && bcode != 0xffffffffu) {  // rare: coincident surfaces -- the larger list index wins
            if (perm_code_to_index(code, kCoop) > perm_code_to_index(bcode, kCoop)) bcode = code;
        }
        const bool closer = t >= tmin && t < bt;
        bt = closer ? t : bt;
        bcode = closer ? code : bcode;
    }
    __syncwarp();  // all reads of the partners' mask words are done before the next sweep overwrites them
    best_t = bt;
    best_k = bcode != 0xffffffffu ? (int)perm_code_to_index(bcode, kCoop) : -1;
}

// Candidate resolution straight from a pair-layout tile (streamed lists: no lane-order copy in shared memory).
// `tile_lane` = address of this lane's first pair of super-chunk 0, k_base = list index of the tile's first sphere;
// best_t / best_k carry the closest hit over the tiles walked so far (tiles come in list order, so ties still go to
// the later sphere).
__device__ __forceinline__ float lds32f(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

template <int NS, int kCoop, int kBlock>
__device__ __forceinline__ void walk_candidates_pairs(uint32_t tile_lane, uint32_t coop_h, uint32_t mask_lane,
                                                      uint32_t k_base, const f3 (&o)[NS], const f3 (&d)[NS],
                                                      const bool (&alive)[NS], const uint32_t (&summary)[NS],
                                                      float (&best_t)[NS], int (&best_k)[NS]) {
    const float tmin = 1e-4f;
#pragma unroll
    for (int r = 0; r < NS; ++r) {
        uint32_t sum = alive[r] ? summary[r] : 0u;
        uint32_t cand = 0u, pbase = 0u, klbase = 0u;
        float bt = best_t[r];
        int bk = best_k[r];
        const float ox = o[r].x, oy = o[r].y, oz = o[r].z, dx = d[r].x, dy = d[r].y, dz = d[r].z;
        for (;;) {
            if (cand == 0u) {
                if (sum == 0u) break;
                const uint32_t c = (uint32_t)__ffs((int)sum) - 1u;
                sum &= sum - 1u;
                cand = ~lds32(mask_lane + (c * NS + r) * (kBlock * 4u));
                pbase = tile_lane + c * (kCoop * 32u * 16u);
                klbase = k_base + c * (32u * kCoop) + 2u * coop_h;
            }
            const uint32_t p = bfind_u32(cand);
            cand &= low_mask(p);
            const uint32_t j = 31u - p;  // this lane's j-th test of the super-chunk: pair j>>1, half j&1
            const uint32_t addr = pbase + (j >> 1) * (kCoop * 32u) + (j & 1u) * 4u;
            const float sx = lds32f(addr), sy = lds32f(addr + 8u), sz = lds32f(addr + 16u), sr = lds32f(addr + 24u);
            const float ocx = ox - sx, ocy = oy - sy, ocz = oz - sz;
            const float hb = fmaf(ocz, dz, fmaf(ocy, dy, ocx * dx));
            const float cq = fmaf(-sr, sr, fmaf(ocz, ocz, fmaf(ocy, ocy, ocx * ocx)));
            if (hb > 0.0f && cq > 0.0f) continue;  // wholly behind the origin: both roots < tmin
            const float sq = __fsqrt_rn(fmaf(hb, hb, -cq));
            const float r1 = -hb - sq, r2 = -hb + sq;
            const bool bad1 = r1 < tmin || bt < r1;
            const bool bad2 = r2 < tmin || bt < r2;
            if (!(bad1 && bad2)) {
                bt = bad1 ? r2 : r1;
                bk = (int)(klbase + (j >> 1) * (2u * kCoop) + (j & 1u));
            }
        }
        best_t[r] = bt;
        best_k[r] = bk;
    }
}

// slot q of a lane holds the ray of lane (lane ^ q) of its cooperating group
template <int kCoop>
__device__ __forceinline__ void exchange_rays(const f3 o, const f3 d, const bool alive, f3 (&so)[kCoop],
                                              f3 (&sd)[kCoop], bool (&sa)[kCoop]) {
    so[0] = o;
    sd[0] = d;
    sa[0] = alive;
#pragma unroll
    for (int q = 1; q < kCoop; ++q) {
        so[q] = mk3(__shfl_xor_sync(kFullMask, o.x, q), __shfl_xor_sync(kFullMask, o.y, q),
                    __shfl_xor_sync(kFullMask, o.z, q));
        sd[q] = mk3(__shfl_xor_sync(kFullMask, d.x, q), __shfl_xor_sync(kFullMask, d.y, q),
                    __shfl_xor_sync(kFullMask, d.z, q));
        sa[q] = __shfl_xor_sync(kFullMask, alive ? 1 : 0, q) != 0;
    }
}

// lane ^ q holds, in ITS slot q, the partial closest hit for my ray: merge with the list-order tie rule
// (equal t => later sphere, src/hit.jl:24-26,44-46)
template <int kCoop>
__device__ __forceinline__ void merge_partial_hits(const float (&bt)[kCoop], const int (&bk)[kCoop], float& best_t,
                                                   int& best_k) {
    best_t = bt[0];
    best_k = bk[0];
#pragma unroll
    for (int q = 1; q < kCoop; ++q) {
        const float pt = __shfl_xor_sync(kFullMask, bt[q], q);
        const int pk = __shfl_xor_sync(kFullMask, bk[q], q);
        if (pk >= 0 && (best_k < 0 || pt < best_t || (pt == best_t && pk > best_k))) {
            best_t = pt;
            best_k = pk;
        }
    }
}

// ---- the kernel ---------------------------------------------------------------------------------------------
template <bool kMulti, int kCoop, bool kOwnWalk, bool kGrid = false, int kMinBlocks = 3, int kBlock = kTraceBlock>
__global__ void __launch_bounds__(kBlock, kMinBlocks) fused_trace2_kernel(const __grid_constant__ TraceParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ __align__(16) uint4 s_coop[kBlock / 32][32];  // rejection-sampling requests of a warp
    const uint32_t n = P.n_spheres;
    const float4* __restrict__ g_src = P.geom_pairs;
    const uint32_t n_stage = (n + 1u) & ~1u;
    const uint32_t n_tiles = kMulti ? (n + kTileSpheres - 1u) / kTileSpheres : 1u;
    constexpr uint32_t kGran = 32u * kCoop;
    const uint32_t tile_cap = kGrid ? 0u : (kMulti ? kTileSpheres : ((n + kGran - 1u) / kGran) * kGran);
    float4* s_tile0 = reinterpret_cast<float4*>(smem_raw);
    float4* s_tile1 = s_tile0 + tile_cap;  // kMulti: second streaming buffer; else: permuted AoS copy of the list
    uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_tile0 + 2u * tile_cap) + threadIdx.x;

    for (uint32_t i = threadIdx.x; i < 2u * tile_cap; i += kBlock) s_tile0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        fence_mbar_init();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    uint32_t bar_phase0 = 0u, bar_phase1 = 0u;
    if (!kMulti && !kGrid) {
        if (threadIdx.x == 0 && n > 0u) {
            mbar_arrive_expect_tx(&s_bar[0], n_stage * 16u + tile_cap * 16u);
            tma_bulk_g2s(s_tile0, g_src, n_stage * 16u, &s_bar[0]);
            tma_bulk_g2s(s_tile1, P.geom_perm, tile_cap * 16u, &s_bar[0]);
        }
        if (n > 0u) mbar_wait(&s_bar[0], 0u);
    }

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    uint4* const coop_slot = s_coop[threadIdx.x >> 5];
    const SweepPlan plan = make_sweep_plan<kCoop>(n, threadIdx.x & (kCoop - 1));

    // path state of the lane
    f3 o = mk3(0.f, 0.f, 0.f), d = mk3(0.f, 1.f, 0.f);
    double thr_r = 1.0, thr_g = 1.0, thr_b = 1.0;  // product of attenuations so far (Float64, as the reference promotes)
    uint32_t pix_local = 0u, pixel = 0u, sample = 0u, nhits = 0u;
    bool alive = false, done = false;
    float best_t = __int_as_float(0x7f800000);
    int best_k = -1;
    uint32_t seg_count = 0;
    uint32_t fb_count = 0;  // kGrid: rays of this warp resolved by the exact fallback sweep (uniform)
    uint32_t loose_count = 0;  // kGrid: cells this lane walked with the loose registration
    uint32_t cell_count = 0, test_count = 0;  // kGrid: cells walked / sphere tests made by this lane (the mode's work model)
    unsigned long long pool_next = 0, pool_end = 0;  // warp-level ticket pool (uniform)
    bool exhausted = false;

    for (;;) {
        // ------------------------------------------------------------ classify + accumulate the ended paths
        bool cont = false;
        if (alive) {
            seg_count += 1;
            if (best_k < 0) {  // miss: sky (src/ray_color.jl:36), path ends
                // skycolor, src/ray_color.jl:1-6: (1-t)*white + t*skyblue with Float64 literals; x*1.0 is exact and omitted
                const float t = 0.5f * (d.y + 1.0f);
                const double a1 = (double)(1.0f - t), w1 = (double)t;
                const double sr = __dadd_rn(a1, __dmul_rn(w1, 0.5));
                const double sg = __dadd_rn(a1, __dmul_rn(w1, 0.7));
                const double sb = __dadd_rn(a1, w1);
                // accumulate (src/render.jl:38): order-independent fixed-point atomics
                unsigned long long* a = P.accum + (unsigned long long)pix_local * 4ull;
                atomicAdd(a + 0, (unsigned long long)__double2ll_rn(__dmul_rn(thr_r, sr) * P.fx_scale));
                atomicAdd(a + 1, (unsigned long long)__double2ll_rn(__dmul_rn(thr_g, sg) * P.fx_scale));
                atomicAdd(a + 2, (unsigned long long)__double2ll_rn(__dmul_rn(thr_b, sb) * P.fx_scale));
                alive = false;
            } else if (++nhits == (uint32_t)P.max_depth) {
                alive = false;  // the next ray_color call returns black (src/ray_color.jl:15-17): adds nothing
            } else {
                cont = true;
            }
        }
        uint32_t kind = 0u;
        if (cont) kind = __ldg(P.kind + best_k);

        // ------------------------------------------------------------ regenerate: idle lanes take the next path
        bool newp = false;
        unsigned long long ticket = 0;
        {
            bool want = !alive && !done;
            unsigned pending = __ballot_sync(kFullMask, want);
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
                const unsigned avail = (unsigned)(pool_end - pool_next);
                const unsigned rank = __popc(pending & lt_mask);
                if (want && rank < avail) { ticket = pool_next + rank; want = false; newp = true; }
                const unsigned npend = __popc(pending);
                pool_next += npend < avail ? npend : avail;
                pending = __ballot_sync(kFullMask, want);
            }
        }
        float su = 0.f, sv = 0.f;
        uint32_t s0 = 0u, pl = 0u;
        if (newp) {
            // ticket -> (pixel, sample): consecutive tickets are consecutive samples of one pixel
            if ((P.n_paths >> 32) == 0ull) {
                pl = magic_div((uint32_t)ticket, P.div_spp);
                s0 = (uint32_t)ticket - pl * (uint32_t)P.spp;
            } else {
                const unsigned long long q = ticket / (unsigned)P.spp;
                pl = (uint32_t)q;
                s0 = (uint32_t)(ticket - q * (unsigned)P.spp);
            }
            s0 += (uint32_t)P.sample_first;
            const uint32_t row_local = magic_div(pl, P.div_w);
            const uint32_t col = pl - row_local * (uint32_t)P.W;
            const uint32_t i0 = (uint32_t)P.row_start + row_local * (uint32_t)P.row_stride;
            su = __ldg(P.u_tab + col);  // T(j/W), src/render.jl:26
            sv = __ldg(P.v_tab + i0);   // T((H-i)/H), src/render.jl:27
            pixel = i0 * (uint32_t)P.W + col;
            sample = s0;
        }
        if (kMulti) {
            if (__syncthreads_or((cont || newp) ? 1 : 0) == 0) break;  // also: everyone is done with both tile buffers
        } else {
            if (__ballot_sync(kFullMask, cont || newp) == 0u) break;
        }

        // ------------------------------------------------------------ block 0 of the event, for every lane at once
        // continuing lanes: event = index of this hit (scatter draws); new lanes: event 0 (primary-ray draws)
        const uint32_t ev = cont ? nhits : 0u;
        const u32x4 b0 = philox_block_rk(P, sample, pixel, ev, 0u);
        float px = pm1x(b0.w0), py = pm1x(b0.w1), pz = pm1x(b0.w2);
        const float pw = pm1x(b0.w3);
        const bool ball = cont && kind != 2u;  // Lambertian and Metal draw a unit vector (Metal even when fuzz == 0)
        float js = su, jt = sv;                // new lanes: jittered screen coordinates (src/render.jl:30-36)
        float coin = 0.f;                      // dielectric coin, draw 3 of the event (src/material.jl:47)
        bool need;
        {
            const float q01 = fmaf(py, py, px * px);
            const float qball = fmaf(pz, pz, q01);
            const float q23 = fmaf(pw, pw, pz * pz);
            if (newp) {
                if (s0 != 0u) {  // du = draw 0, dv = draw 1; the first sample is centred
                    js = su + __fdiv_rn(u01x(b0.w0), (float)P.W);
                    jt = sv + __fdiv_rn(u01x(b0.w1), (float)P.H);
                }
                px = pz;  // disk attempt 0 = draws 2, 3 (src/rand.jl:31-38; always drawn, src/camera.jl:44)
                py = pw;
            } else {
                coin = u01x(b0.w3);
            }
            need = ball ? !(qball <= 1.0f) : (newp ? !(q23 <= 1.0f) : false);
        }
        // ------------------------------------------------------------ cooperative rejection sampling
        // attempt a >= 1 of the ball  = words 0..2 of block a of the event (src/rand.jl:15-22)
        // attempts 2a-1, 2a of the disk = words (0,1), (2,3) of block a of event 0 (src/rand.jl:31-38)
        {
            unsigned needm = __ballot_sync(kFullMask, need);
            uint32_t blk = 1u;  // next unevaluated block; uniform: every needy lane has failed the same attempts
            if (needm) {
                // block 1 by the lane itself: after attempt 0 about 40 % of the lanes are in need, so a cooperative pass
                // would give each of them two helper lanes at the price of the request exchange; one more own block
                // halves their number for less, and the cooperative passes take over from block 2 with 5+ helpers each
                const u32x4 hb = philox_block_rk(P, sample, pixel, ev, 1u);
                float hx = pm1x(hb.w0), hy = pm1x(hb.w1);
                const float hz = pm1x(hb.w2), hw = pm1x(hb.w3);
                const float q01 = fmaf(hy, hy, hx * hx);
                const float qball = fmaf(hz, hz, q01);
                const float q23 = fmaf(hw, hw, hz * hz);
                const bool ok01 = q01 <= 1.0f;
                const float qsel = newp ? fminf(q01, q23) : qball;
                if (newp & !ok01) { hx = hz; hy = hw; }
                if (need && qsel <= 1.0f) { px = hx; py = hy; pz = hz; need = false; }
                blk = 2u;
                needm = __ballot_sync(kFullMask, need);
            }
            while (needm) {
                const uint32_t cnt = (uint32_t)__popc(needm);
                const uint32_t tab = c_coop_tab[cnt];
                const uint32_t per = tab & 0xffu;  // helper lanes (= attempts evaluated) per needy lane: 32 / cnt
                const uint32_t rank = (uint32_t)__popc(needm & lt_mask);
                if (need) coop_slot[rank] = make_uint4(sample, pixel, ev | (newp ? 0x80000000u : 0u), 0u);
                __syncwarp();
                const uint32_t hq = (lane * (tab >> 8)) >> 8;  // lane / per: the request this lane helps
                const uint32_t ha = lane - hq * per;           // and which of its attempts
                const uint4 tsk = coop_slot[hq < cnt ? hq : 0u];
                const u32x4 hb = philox_block_rk(P, tsk.x, tsk.y, tsk.z & 0x7fffffffu, blk + ha);
                float hx = pm1x(hb.w0), hy = pm1x(hb.w1);
                const float hz = pm1x(hb.w2), hw = pm1x(hb.w3);
                const float q01 = fmaf(hy, hy, hx * hx);
                const float qball = fmaf(hz, hz, q01);
                const float q23 = fmaf(hw, hw, hz * hz);
                const bool is_disk = (int)tsk.z < 0;
                const bool ok01 = q01 <= 1.0f;
                const float qsel = is_disk ? fminf(q01, q23) : qball;
                const bool ok = (hq < cnt) & (qsel <= 1.0f);
                if (is_disk & !ok01) { hx = hz; hy = hw; }
                const unsigned okm = __ballot_sync(kFullMask, ok);
                const uint32_t first = rank * per;
                const uint32_t mine = need ? ((okm >> first) & low_mask(per)) : 0u;
                const uint32_t src = mine ? first + (uint32_t)__ffs((int)mine) - 1u : lane;
                const float gx = __shfl_sync(kFullMask, hx, src);
                const float gy = __shfl_sync(kFullMask, hy, src);
                const float gz = __shfl_sync(kFullMask, hz, src);
                if (mine) { px = gx; py = gy; pz = gz; need = false; }
                blk += per;
                needm = __ballot_sync(kFullMask, need);
                __syncwarp();  // every lane has read its request before the next pass overwrites the slots
            }
        }

        // ------------------------------------------------------------ new lanes: get_ray (src/camera.jl:43-48)
        f3 v1 = mk3(px, py, pz);  // continuing Lambertian/Metal lanes: the point in the unit ball
        f3 o_new = o;
        if (newp) {
            const DevCamera& c = P.cam;
            const float rx = c.lens_radius * px, ry = c.lens_radius * py;
            const f3 off = mk3(fmaf(c.v.x, ry, c.u.x * rx), fmaf(c.v.y, ry, c.u.y * rx), fmaf(c.v.z, ry, c.u.z * rx));
            o_new = mk3(c.origin.x + off.x, c.origin.y + off.y, c.origin.z + off.z);
            v1.x = fmaf(jt, c.vertical.x, fmaf(js, c.horizontal.x, c.llc.x)) - c.origin.x - off.x;
            v1.y = fmaf(jt, c.vertical.y, fmaf(js, c.horizontal.y, c.llc.y)) - c.origin.y - off.y;
            v1.z = fmaf(jt, c.vertical.z, fmaf(js, c.horizontal.z, c.llc.z)) - c.origin.z - off.z;
        }
        const f3 n1 = normalize3(v1);  // unit(ball sample) | primary direction: one instance for both

        // ------------------------------------------------------------ continuing lanes: HitRecord + scatter
        f3 v2 = mk3(1.f, 0.f, 0.f);  // direction before the final normalize
        f3 alt = v2;                 // direction used as is (near-zero Lambertian, reflecting Dielectric)
        bool use_alt = false;
        if (cont) {
            const float4 g = __ldg(P.geom + best_k);
            const float4 m = __ldg(P.mat + best_k);
            const f3 p = mk3(fmaf(best_t, d.x, o.x), fmaf(best_t, d.y, o.y), fmaf(best_t, d.z, o.z));   // hit.jl:3
            const f3 on = mk3(__fdiv_rn(p.x - g.x, g.w), __fdiv_rn(p.y - g.y, g.w), __fdiv_rn(p.z - g.z, g.w));  // hit.jl:33
            const bool front = dot3(d, on) < 0.0f;                                                    // hit.jl:7
            const f3 nn = front ? on : mk3(-on.x, -on.y, -on.z);                                      // hit.jl:8
            if (kind == 0u) {  // Lambertian, material.jl:13-23
                v2 = mk3(nn.x + n1.x, nn.y + n1.y, nn.z + n1.z);
                // near_zero: squared length (Float32) promoted and compared with the Float64 literal 1e-5, vec.jl:20
                use_alt = (double)dot3(v2, v2) < 1e-5;
                alt = nn;
            } else if (kind == 1u) {  // Metal, material.jl:31-34; never absorbs (structs.jl:43)
                const f3 refl = reflect3(d, nn);
                v2 = mk3(fmaf(m.w, n1.x, refl.x), fmaf(m.w, n1.y, refl.y), fmaf(m.w, n1.z, refl.z));
            } else {  // Dielectric, material.jl:41-53
                const float ratio = front ? __fdiv_rn(1.0f, m.w) : m.w;
                const float cos_t = fminf(-dot3(d, nn), 1.0f);
                const float sin_t = __fsqrt_rn(fmaf(-cos_t, cos_t, 1.0f));
                // `||` short-circuits (material.jl:47): the coin is ignored on total internal reflection
                if (ratio * sin_t > 1.0f || reflectance(cos_t, ratio) > coin) {
                    use_alt = true;
                    alt = reflect3(d, nn);  // not re-normalised, material.jl:48
                } else {  // refract, light.jl:12-17
                    const f3 perp = mk3(ratio * fmaf(cos_t, nn.x, d.x), ratio * fmaf(cos_t, nn.y, d.y),
                                        ratio * fmaf(cos_t, nn.z, d.z));
                    const float s = __fsqrt_rn(fabsf(1.0f - dot3(perp, perp)));
                    v2 = mk3(fmaf(-s, nn.x, perp.x), fmaf(-s, nn.y, perp.y), fmaf(-s, nn.z, perp.z));
                }
            }
            if (kind != 2u) {  // attenuation = albedo (Dielectric: (1,1,1), an exact no-op)
                thr_r = __dmul_rn(thr_r, (double)m.x);
                thr_g = __dmul_rn(thr_g, (double)m.y);
                thr_b = __dmul_rn(thr_b, (double)m.z);
            }
            o = p;
        }
        const f3 n2 = normalize3(v2);  // one instance for the three materials
        if (cont) {
            d = use_alt ? alt : n2;
        } else if (newp) {
            o = o_new;
            d = n1;
            thr_r = thr_g = thr_b = 1.0;
            nhits = 0u;
            pix_local = pl;
            alive = true;
        }

        // ------------------------------------------------------------ intersect: closest hit over the list
        best_t = __int_as_float(0x7f800000);  // typemax(T) = Inf, src/ray_color.jl:19
        best_k = -1;
        if (kGrid) {
            const bool unsafe = closest_hit_grid(P.grid, P.geom, o, d, alive, best_t, best_k, loose_count, cell_count, test_count);
            // rays the grid cannot answer exactly (non-unit direction after a glass reflection, very long flights):
            // warp-cooperative sweep of the whole list, for every list size -- the mode is exact by construction
            fb_count += grid_fallback_sweep(P.grid.cull, P.geom, n, o, d, unsafe, best_t, best_k);
        } else {
            const uint32_t h = threadIdx.x & (kCoop - 1);
            f3 so[kCoop], sd[kCoop];
            bool sa[kCoop];
            float bt[kCoop];
            int bk[kCoop];
            uint32_t summary[kCoop];
            exchange_rays<kCoop>(o, d, alive, so, sd, sa);
            if (!kMulti && kOwnWalk) {
                const uint32_t sum1 = sweep_masks_packed3<kCoop, kBlock>(smem_u32(s_tile0) + h * 32u, plan,
                                                                         smem_u32(s_mask), so, sd);
                walk_own_ray_perm<kCoop, kBlock>(s_tile1, s_mask - threadIdx.x, o, d, alive, sum1, best_t, best_k);
            } else if (!kMulti) {
                sweep_masks_packed2<kCoop, kCoop, kBlock>(smem_u32(s_tile0) + h * 32u, n, h, smem_u32(s_mask), so, sd,
                                                               summary);
                walk_candidates_perm<kCoop, kCoop, kBlock>(s_tile1, h, s_mask, so, sd, sa, summary, bt, bk);
            } else {
#pragma unroll
                for (int q = 0; q < kCoop; ++q) {
                    bt[q] = __int_as_float(0x7f800000);
                    bk[q] = -1;
                }
                if (threadIdx.x == 0) {  // prologue: tile 0 -> buffer 0
                    const uint32_t cnt = n_stage < kTileSpheres ? n_stage : kTileSpheres;
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
                    const uint32_t tile_lane = smem_u32((t & 1u) ? s_tile1 : s_tile0) + h * 32u;
                    if (t & 1u) { mbar_wait(&s_bar[1], bar_phase1); bar_phase1 ^= 1u; }
                    else { mbar_wait(&s_bar[0], bar_phase0); bar_phase0 ^= 1u; }
                    sweep_masks_packed2<kCoop, kCoop, kBlock>(tile_lane, cnt, h, smem_u32(s_mask), so, sd, summary);
                    walk_candidates_pairs<kCoop, kCoop, kBlock>(tile_lane, h, smem_u32(s_mask), base, so, sd, sa,
                                                                     summary, bt, bk);
                    __syncthreads();  // the buffer may be overwritten by the prefetch issued in the next iteration
                }
            }
            if (kMulti || !kOwnWalk) merge_partial_hits<kCoop>(bt, bk, best_t, best_k);
        }
    }
    // ray-segment statistics: one atomic per warp
    for (int off = 16; off > 0; off >>= 1) seg_count += __shfl_xor_sync(kFullMask, seg_count, off);
    if (lane == 0 && seg_count) atomicAdd(P.counters + 1, (unsigned long long)seg_count);
    if (kGrid && lane == 0 && fb_count) atomicAdd(P.counters + 2, (unsigned long long)fb_count);
    if (kGrid) {
        for (int off = 16; off > 0; off >>= 1) loose_count += __shfl_xor_sync(kFullMask, loose_count, off);
        if (lane == 0 && loose_count) atomicAdd(P.counters + 3, (unsigned long long)loose_count);
        unsigned long long cc = cell_count, tc = test_count;  // 64-bit: a lane can pass 2^32 tests in a long render
        for (int off = 16; off > 0; off >>= 1) {
            cc += __shfl_xor_sync(kFullMask, cc, off);
            tc += __shfl_xor_sync(kFullMask, tc, off);
        }
        if (lane == 0) {
            atomicAdd(P.counters + 4, cc);
            atomicAdd(P.counters + 5, tc);
        }
    }
}

// u_tab[col] = T((col+1)/W), v_tab[i0] = T((H-1-i0)/H) (src/render.jl:26-27): the quotient is formed in Float64 and
// rounded to Float32 by the reference; both operands are integers < 2^24, so the correctly rounded Float32 quotient
// is the same value (rounding through a format with >= 2*24+2 bits is innocuous for division).
__global__ void __launch_bounds__(256) uv_table_kernel(int W, int H, float* __restrict__ u_tab,
                                                       float* __restrict__ v_tab) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < W) u_tab[t] = __fdiv_rn((float)(t + 1), (float)W);
    if (t < H) v_tab[t] = __fdiv_rn((float)(H - 1 - t), (float)H);
}

template <bool kMulti, int kCoop, bool kOwnWalk, int kMinBlocks = 3, int kBlock = kTraceBlock>
cudaError_t launch_variant2(const TraceParams& p, int num_sms, int blocks_per_sm_override, cudaStream_t stream,
                            LaunchInfo* info) {
    auto kern = fused_trace2_kernel<kMulti, kCoop, kOwnWalk, false, kMinBlocks, kBlock>;
    constexpr uint32_t kGran = 32u * kCoop;
    const uint32_t tile_cap = kMulti ? kTileSpheres : ((p.n_spheres + kGran - 1u) / kGran) * kGran;
    const uint32_t chunks = tile_cap / 32u;  // mask words per lane: (tile_cap / kGran) super-chunks x kCoop slots
    const int smem = (int)(2u * tile_cap * 16u + chunks * kBlock * 4u);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBlock, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (blocks_per_sm_override > 0 && blocks_per_sm_override < per_sm) per_sm = blocks_per_sm_override;
    long long grid = (long long)num_sms * per_sm;  // persistent grid: every CTA is resident
    const long long max_useful = (long long)((p.n_paths + (unsigned long long)kBlock - 1ull) /
                                             (unsigned long long)kBlock);
    if (grid > max_useful) grid = max_useful;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, kBlock, smem, stream>>>(p);
    if (info) {
        info->grid = (int)grid;
        info->block = kBlock;
        info->smem_bytes = smem;
        info->blocks_per_sm = per_sm;
        info->launches = 1;
        info->rays_per_lane = 1;
        info->sweep = kSweepPacked;
    }
    return cudaGetLastError();
}

}  // namespace

MagicDiv make_magic_div(uint32_t d) {
    // Granlund & Montgomery, "Division by invariant integers using multiplication", fig. 4.1 (N = 32)
    MagicDiv k;
    uint32_t L = 0;
    while ((1ull << L) < (unsigned long long)d) ++L;
    k.m = (uint32_t)(((1ull << 32) * ((1ull << L) - (unsigned long long)d)) / (unsigned long long)d + 1ull);
    k.sh1 = L < 1u ? L : 1u;
    k.sh2 = L > 0u ? L - 1u : 0u;
    return k;
}

cudaError_t launch_uv_tables(int W, int H, float* u_tab, float* v_tab, cudaStream_t stream) {
    const int n = W > H ? W : H;
    if (n <= 0) return cudaSuccess;
    uv_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(W, H, u_tab, v_tab);
    return cudaGetLastError();
}

cudaError_t launch_fused_trace2_grid(const TraceParams& p, int num_sms, int blocks_per_sm_override, cudaStream_t stream,
                                     LaunchInfo* info) {
    // 64 registers, 4 CTAs per SM: the traversal is latency-bound and wants the warps
    auto kern = fused_trace2_kernel<false, 2, false, true, 4>;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTraceBlock, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (blocks_per_sm_override > 0 && blocks_per_sm_override < per_sm) per_sm = blocks_per_sm_override;
    long long grid = (long long)num_sms * per_sm;
    const long long max_useful = (long long)((p.n_paths + (unsigned long long)kTraceBlock - 1ull) /
                                             (unsigned long long)kTraceBlock);
    if (grid > max_useful) grid = max_useful;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, kTraceBlock, 0, stream>>>(p);
    if (info) {
        info->grid = (int)grid;
        info->block = kTraceBlock;
        info->smem_bytes = 0;
        info->blocks_per_sm = per_sm;
        info->launches = 1;
        info->rays_per_lane = 1;
        info->sweep = 0;
    }
    return cudaGetLastError();
}

cudaError_t launch_fused_trace2(const TraceParams& p, int num_sms, int blocks_per_sm_override, int coop, int walk,
                                cudaStream_t stream, LaunchInfo* info) {
    const bool multi = p.n_spheres > kTileSpheres;
    // the shipped configuration: 2 cooperating lanes; single-tile lists resolve candidates with the own-ray walk,
    // streamed lists (> 1024 spheres) tile by tile with the per-slot walk
    if (coop == 0) coop = 2;
    if (walk == 0) walk = 2;  // measured (profiles/r02_summary.md): own-ray walk wins with 2 and with 4 lanes
    if (multi && coop == 2) return launch_variant2<true, 2, false>(p, num_sms, blocks_per_sm_override, stream, info);
    if (!multi && coop == 2 && walk == 2) return launch_variant2<false, 2, true>(p, num_sms, blocks_per_sm_override, stream, info);
#ifdef RTW_BUILD_VARIANTS
    if (multi) return launch_variant2<true, 4, false>(p, num_sms, blocks_per_sm_override, stream, info);
    if (walk == 1) {  // per-slot candidate walks + merge
        return coop == 4 ? launch_variant2<false, 4, false>(p, num_sms, blocks_per_sm_override, stream, info)
                         : launch_variant2<false, 2, false>(p, num_sms, blocks_per_sm_override, stream, info);
    }
    if (blocks_per_sm_override == 2)  // 128 registers per thread, 2 CTAs per SM
        return launch_variant2<false, 4, true, 2>(p, num_sms, blocks_per_sm_override, stream, info);
    return launch_variant2<false, 4, true>(p, num_sms, blocks_per_sm_override, stream, info);
#else
    return cudaErrorNotSupported;  // built without RTW_BUILD_VARIANTS=1
#endif
}

}  // namespace rtw
