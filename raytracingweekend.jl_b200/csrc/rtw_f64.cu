// rtw_f64.cu -- the Float64 instantiation of the hot path.
//
// The reference is generic over T (src/scenes.jl:2 `elem_type`, src/camera.jl:38) and its own test and every
// timing it publishes use Float64 (test/runtests.jl:190-194, README.md:86-122).  This file is the same persistent
// fused kernel -- regenerate -> closest-hit sweep -> shade/scatter -> accumulate -- written for T = Float64 on the
// FP64 pipe (DADD/DMUL/DFMA; B200 issues them at half the FP32 rate), so that a Camera{Float64} / Float64 scene is
// served by the GPU instead of being refused.  It follows the reference functions line for line like the Float32
// path (rtw_device.cuh) and the same floating-point contract: dot = fma(z,z, fma(y,y, x*x)), a +- b*c is one fma,
// sqrt and / are IEEE, normalize(v) = v * (1/sqrt(v.v)); -fmad=false, nothing else is contracted.
//
// Stream: the same addressed Philox4x32-7, a Float64 draw takes two words (u64 = hi:lo, f64 = (u64 >> 12) * 2^-52,
// RandomNumbers-style 52 random mantissa bits):  draw n of an event = words (2(n&1), 2(n&1)+1) of block n >> 1.
//   event 0: draws 0,1 = jitter; disk attempt k = draws 2+2k, 3+2k (block 1+k)
//   event e: ball attempt a = draws 4a, 4a+1, 4a+2 (blocks 2a, 2a+1); draw 3 = dielectric coin (block 1, words 2,3)
// The sweep works on chunks of 32 spheres: branch-free discriminants whose sign bits are collected into a mask by
// funnel shifts (as in the Float32 kernel, un-packed: there is no packed FP64), then each lane resolves its own
// candidates of the chunk in list order.  The first version tested and branched per sphere: 26 instructions per
// test for 11 of arithmetic; this one issues 14.
#include "rtw_kernels.h"
#include "rtw_sweep.cuh"

namespace rtw {

namespace {

struct d3 {
    double x, y, z;
};
__device__ __forceinline__ d3 mkd(double x, double y, double z) { return d3{x, y, z}; }
__device__ __forceinline__ double dotd(d3 a, d3 b) { return fma(a.z, b.z, fma(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ d3 normalized(d3 a) {
    const double inv = 1.0 / sqrt(dotd(a, a));  // StaticArrays 1.2.13: inv(norm(v)) * v
    return mkd(a.x * inv, a.y * inv, a.z * inv);
}
// trand(Float64), src/rand.jl:10-13: 52 random mantissa bits
__device__ __forceinline__ double u01d(uint32_t lo, uint32_t hi) {
    return (double)((((unsigned long long)hi << 32) | lo) >> 12) * 2.220446049250313e-16;  // 2^-52
}
__device__ __forceinline__ double pm1d(uint32_t lo, uint32_t hi) { return fma(u01d(lo, hi), 2.0, -1.0); }  // rand.jl:24

__device__ __forceinline__ d3 reflectd(d3 v, d3 n) {  // src/light.jl:6
    const double k = 2.0 * dotd(v, n);
    return mkd(fma(-k, n.x, v.x), fma(-k, n.y, v.y), fma(-k, n.z, v.z));
}

// cooperative rejection sampling: with `cnt` lanes in need, each gets per = 32 / cnt helper lanes;
// entry = per | (ceil(256 / per) << 8), so that lane / per = (lane * (entry >> 8)) >> 8 for lane < 32
__constant__ uint32_t c_coop_tab64[33] = {
    0u,
    32u | (8u << 8), 16u | (16u << 8), 10u | (26u << 8), 8u | (32u << 8), 6u | (43u << 8), 5u | (52u << 8),
    4u | (64u << 8), 4u | (64u << 8), 3u | (86u << 8), 3u | (86u << 8), 2u | (128u << 8), 2u | (128u << 8),
    2u | (128u << 8), 2u | (128u << 8), 2u | (128u << 8), 2u | (128u << 8), 1u | (256u << 8), 1u | (256u << 8),
    1u | (256u << 8), 1u | (256u << 8), 1u | (256u << 8), 1u | (256u << 8), 1u | (256u << 8), 1u | (256u << 8),
    1u | (256u << 8), 1u | (256u << 8), 1u | (256u << 8), 1u | (256u << 8), 1u | (256u << 8), 1u | (256u << 8),
    1u | (256u << 8), 1u | (256u << 8)};

__device__ __forceinline__ double shfl_d(double v, uint32_t src) { return __shfl_sync(kFullMask, v, (int)src); }

// The loop is laid out like the Float32 kernel's (rtw_fused2.cu): classify + accumulate the paths that ended ->
// regenerate (idle lanes take the next ticket) -> blocks 0 and 1 of the event's stream for EVERY lane at once (scatter
// draws of the continuing lanes, primary-ray draws of the new ones) -> warp-cooperative rejection sampling (the stream is
// addressed, so attempt a of a path can be evaluated by any lane; a lane takes its first accepted attempt in stream
// order) -> one normalize() for unit(ball sample) | primary direction, one for the scattered direction of all three
// materials -> closest-hit sweep.  The first Float64 kernel ran each state's code by itself: the two rejection loops
// alone were 650 of its 8500 issued instructions per bounce, at 3-6 active lanes.
template <bool kShared>
__global__ void __launch_bounds__(kTraceBlock, 2) trace_f64_kernel(const __grid_constant__ TraceParams64 P) {
    extern __shared__ __align__(16) unsigned char smem_raw64[];
    __shared__ __align__(16) uint4 s_coop[kTraceBlock / 32][32];  // rejection-sampling requests of a warp
    double4* s_geom = reinterpret_cast<double4*>(smem_raw64);
    const uint32_t n = P.n_spheres;
    // kShared: one 32-test candidate mask per chunk and lane, behind the list (word c of a lane at (c * block + tid))
    uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_geom + (kShared ? n : 0u)) + threadIdx.x;
    if (kShared) {
        for (uint32_t i = threadIdx.x; i < n; i += kTraceBlock) s_geom[i] = P.geom[i];
        __syncthreads();
    }
    const double4* __restrict__ list = kShared ? s_geom : P.geom;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t k0 = P.key0, k1 = P.key1;
    const double tmin = 1e-4;  // T(1e-4), src/ray_color.jl:19
    uint4* const coop_slot = s_coop[threadIdx.x >> 5];

    d3 o = mkd(0, 0, 0), d = mkd(0, 1, 0);
    double thr_r = 1.0, thr_g = 1.0, thr_b = 1.0;
    uint32_t pix_local = 0u, nhits = 0u;
    PathRng rng{0u, 0u};
    bool alive = false, done = false;
    uint32_t seg_count = 0;
    unsigned long long pool_next = 0, pool_end = 0;
    bool exhausted = false;
    double best_t = __longlong_as_double(0x7ff0000000000000ll);  // typemax(T)
    int best_k = -1;

    for (;;) {
        // ---- classify + accumulate the paths that ended
        bool cont = false;
        uint32_t kind = 0u;
        if (alive) {
            seg_count += 1;
            if (best_k < 0) {  // miss: skycolor, src/ray_color.jl:1-6 (no contraction)
                const double t = 0.5 * (d.y + 1.0);
                const double a = 1.0 - t;
                const double sr = a + t * 0.5, sg = a + t * 0.7, sb = a + t;
                unsigned long long* acc = P.accum + (unsigned long long)pix_local * 4ull;
                atomicAdd(acc + 0, (unsigned long long)__double2ll_rn(thr_r * sr * P.fx_scale));
                atomicAdd(acc + 1, (unsigned long long)__double2ll_rn(thr_g * sg * P.fx_scale));
                atomicAdd(acc + 2, (unsigned long long)__double2ll_rn(thr_b * sb * P.fx_scale));
                alive = false;
            } else if (++nhits == (uint32_t)P.max_depth) {
                alive = false;  // the next ray_color call returns black, src/ray_color.jl:15-17
            } else {
                cont = true;
                kind = __ldg(P.kind + best_k);
            }
        }
        // ---- regenerate: idle lanes take the next path ticket (src/render.jl:24-37)
        bool newp = false;
        double su = 0.0, sv = 0.0;
        uint32_t s0 = 0u, pl = 0u;
        {
            bool want = !alive && !done;
            unsigned pending = __ballot_sync(kFullMask, want);
            unsigned long long ticket = 0;
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
            if (newp) {
                const unsigned long long q = ticket / (unsigned)P.spp;
                pl = (uint32_t)q;
                s0 = (uint32_t)(ticket - q * (unsigned)P.spp) + (uint32_t)P.sample_first;
                const uint32_t row_local = pl / (uint32_t)P.W, col = pl - row_local * (uint32_t)P.W;
                const uint32_t i0 = (uint32_t)P.row_start + row_local * (uint32_t)P.row_stride;
                su = (double)(col + 1u) / (double)P.W;                  // u = T(j/W), src/render.jl:26
                sv = (double)((uint32_t)P.H - 1u - i0) / (double)P.H;   // v = T((H-i)/H), src/render.jl:27
                rng.pixel = i0 * (uint32_t)P.W + col;
                rng.sample = s0;
            }
        }
        if (__ballot_sync(kFullMask, cont || newp) == 0u) break;

        // ---- blocks 0 and 1 of the event for every lane at once
        // continuing lanes: event = index of this hit -- ball attempt 0 = draws 0, 1, 2 (block 0, block 1 words 0-1), the
        // dielectric coin = draw 3 (block 1 words 2-3); new lanes: event 0 -- jitter = draws 0, 1 (block 0), disk attempt 0 =
        // draws 2, 3 (block 1)
        const uint32_t ev = cont ? nhits : 0u;
        double px, py, pz = 0.0, coin = 0.0;
        bool need = false;
        {
            const u32x4 b0 = philox_block(rng, ev, 0u, k0, k1);
            const u32x4 b1 = philox_block(rng, ev, 1u, k0, k1);
            if (newp) {
                if (s0 != 0u) {  // src/render.jl:30-36: the first sample is centred; du = draw 0, dv = draw 1
                    su += u01d(b0.w0, b0.w1) / (double)(float)P.W;
                    sv += u01d(b0.w2, b0.w3) / (double)(float)P.H;
                }
                px = pm1d(b1.w0, b1.w1);  // random_vec2_in_disk (src/rand.jl:31-38) is always drawn, src/camera.jl:44
                py = pm1d(b1.w2, b1.w3);
                need = !(fma(py, py, px * px) <= 1.0);
            } else {
                px = pm1d(b0.w0, b0.w1);
                py = pm1d(b0.w2, b0.w3);
                pz = pm1d(b1.w0, b1.w1);
                coin = u01d(b1.w2, b1.w3);
                need = cont && kind != 2u && !(dotd(mkd(px, py, pz), mkd(px, py, pz)) <= 1.0);  // src/rand.jl:15-22
            }
        }
        // ---- cooperative rejection sampling: ball attempt a = blocks 2a, 2a+1 of the event; disk attempt a = block 1+a
        {
            unsigned needm = __ballot_sync(kFullMask, need);
            uint32_t tried = 1u;  // attempts evaluated so far; uniform: every needy lane has failed the same attempts
            while (needm) {
                const uint32_t cnt = (uint32_t)__popc(needm);
                const uint32_t tab = c_coop_tab64[cnt];
                const uint32_t per = tab & 0xffu;  // helper lanes (= attempts evaluated) per needy lane: 32 / cnt
                const uint32_t rank = (uint32_t)__popc(needm & lt_mask);
                if (need) coop_slot[rank] = make_uint4(rng.sample, rng.pixel, ev | (newp ? 0x80000000u : 0u), 0u);
                __syncwarp();
                const uint32_t hq = (lane * (tab >> 8)) >> 8;  // lane / per: the request this lane helps
                const uint32_t ha = lane - hq * per;           // and which of its attempts
                const uint4 tsk = coop_slot[hq < cnt ? hq : 0u];
                const bool is_disk = (int)tsk.z < 0;
                const uint32_t att = tried + ha;
                const PathRng hr{tsk.x, tsk.y};
                const uint32_t hev = tsk.z & 0x7fffffffu;
                const u32x4 ba = philox_block(hr, hev, is_disk ? 1u + att : 2u * att, k0, k1);
                const u32x4 bb = philox_block(hr, hev, 2u * att + 1u, k0, k1);  // third coordinate of a ball attempt
                const double hx = pm1d(ba.w0, ba.w1), hy = pm1d(ba.w2, ba.w3);
                const double hz = is_disk ? 0.0 : pm1d(bb.w0, bb.w1);
                // disk: fma(y, y, x*x) (src/rand.jl:31-38); ball: dot = fma(z, z, fma(y, y, x*x)) -- with z = 0 the same value
                const bool ok = (hq < cnt) & (fma(hz, hz, fma(hy, hy, hx * hx)) <= 1.0);
                const unsigned okm = __ballot_sync(kFullMask, ok);
                const uint32_t first = rank * per;
                const uint32_t mine = need ? ((okm >> first) & (per >= 32u ? 0xffffffffu : ((1u << per) - 1u))) : 0u;
                const uint32_t src = mine ? first + (uint32_t)__ffs((int)mine) - 1u : lane;
                const double gx = shfl_d(hx, src), gy = shfl_d(hy, src), gz = shfl_d(hz, src);
                if (mine) { px = gx; py = gy; pz = gz; need = false; }
                tried += per;
                needm = __ballot_sync(kFullMask, need);
                __syncwarp();  // every lane has read its request before the next pass overwrites the slots
            }
        }
        // ---- new lanes: get_ray, src/camera.jl:43-48
        d3 v1 = mkd(px, py, pz);  // continuing Lambertian / Metal lanes: the point in the unit ball
        d3 o_new = o;
        if (newp) {
            const DevCamera64& c = P.cam;
            const double rx = c.lens_radius * px, ry = c.lens_radius * py;
            const d3 off = mkd(fma(c.v[0], ry, c.u[0] * rx), fma(c.v[1], ry, c.u[1] * rx), fma(c.v[2], ry, c.u[2] * rx));
            o_new = mkd(c.origin[0] + off.x, c.origin[1] + off.y, c.origin[2] + off.z);
            v1.x = fma(sv, c.vertical[0], fma(su, c.horizontal[0], c.llc[0])) - c.origin[0] - off.x;
            v1.y = fma(sv, c.vertical[1], fma(su, c.horizontal[1], c.llc[1])) - c.origin[1] - off.y;
            v1.z = fma(sv, c.vertical[2], fma(su, c.horizontal[2], c.llc[2])) - c.origin[2] - off.z;
        }
        const d3 n1 = normalized(v1);  // unit(ball sample) | primary direction: one instance for both
        // ---- continuing lanes: HitRecord + scatter
        d3 v2 = mkd(1.0, 0.0, 0.0);  // direction before the final normalize
        d3 alt = v2;                 // direction used as is (near-zero Lambertian, reflecting Dielectric)
        bool use_alt = false;
        if (cont) {
            const double4 g = P.geom[best_k];
            const double4 m = P.mat[best_k];
            const d3 p = mkd(fma(best_t, d.x, o.x), fma(best_t, d.y, o.y), fma(best_t, d.z, o.z));  // hit.jl:3
            const d3 on = mkd((p.x - g.x) / g.w, (p.y - g.y) / g.w, (p.z - g.z) / g.w);             // hit.jl:33
            const bool front = dotd(d, on) < 0.0;                                                  // hit.jl:7
            const d3 nn = front ? on : mkd(-on.x, -on.y, -on.z);
            if (kind == 0u) {  // Lambertian, src/material.jl:13-23
                v2 = mkd(nn.x + n1.x, nn.y + n1.y, nn.z + n1.z);
                use_alt = dotd(v2, v2) < 1e-5;  // near_zero, src/vec.jl:20
                alt = nn;
            } else if (kind == 1u) {  // Metal, src/material.jl:31-34; never absorbs (src/structs.jl:43)
                const d3 refl = reflectd(d, nn);
                v2 = mkd(fma(m.w, n1.x, refl.x), fma(m.w, n1.y, refl.y), fma(m.w, n1.z, refl.z));
            } else {  // Dielectric, src/material.jl:41-53
                const double ratio = front ? 1.0 / m.w : m.w;
                const double cos_t = fmin(-dotd(d, nn), 1.0);
                const double sin_t = sqrt(fma(-cos_t, cos_t, 1.0));
                double r0 = (1.0 - ratio) / (1.0 + ratio);  // Schlick, src/light.jl:19-25
                r0 = r0 * r0;
                const double x = 1.0 - cos_t, x2 = x * x, x4 = x2 * x2;
                // `||` short-circuits (material.jl:47): the coin is ignored on total internal reflection
                if (ratio * sin_t > 1.0 || fma(1.0 - r0, x4 * x, r0) > coin) {
                    use_alt = true;
                    alt = reflectd(d, nn);  // not re-normalised, src/material.jl:48
                } else {  // refract, src/light.jl:12-17
                    const d3 perp = mkd(ratio * fma(cos_t, nn.x, d.x), ratio * fma(cos_t, nn.y, d.y),
                                        ratio * fma(cos_t, nn.z, d.z));
                    const double sp = sqrt(fabs(1.0 - dotd(perp, perp)));
                    v2 = mkd(fma(-sp, nn.x, perp.x), fma(-sp, nn.y, perp.y), fma(-sp, nn.z, perp.z));
                }
            }
            if (kind != 2u) {  // attenuation = albedo (Dielectric: ones)
                thr_r *= m.x;
                thr_g *= m.y;
                thr_b *= m.z;
            }
            o = p;
        }
        const d3 n2 = normalized(v2);  // one instance for the three materials
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

        // ---- intersect: hit(::HittableList), src/hit.jl:38-50, hit(::Sphere) src/hit.jl:12-35
        best_t = __longlong_as_double(0x7ff0000000000000ll);  // typemax(T)
        best_k = -1;
        // one ray-sphere test with root selection against the running closest t (list order: ties go to the later sphere)
        auto resolve = [&](uint32_t k) {
            const double4 s = list[k];
            const d3 oc = mkd(o.x - s.x, o.y - s.y, o.z - s.z);
            const double hb = dotd(oc, d);
            const double cq = fma(-s.w, s.w, dotd(oc, oc));
            const double disc = fma(hb, hb, -cq);
            if (disc < 0.0) return;  // src/hit.jl:19
            const double sq = sqrt(disc);
            double root = -hb - sq;
            if (root < tmin || best_t < root) {
                root = -hb + sq;
                if (root < tmin || best_t < root) return;
            }
            best_t = root;
            best_k = (int)k;
        };
        if constexpr (kShared) {
            // ---- lists <= 1024 spheres, in shared memory.  Two lanes cooperate: a broadcast LDS.128 delivers 512 B to the
            // warp and the shared-memory pipe moves 128 B/clk/SM, so with two loads per test (a double4) the sweep of four
            // sub-partitions would need 32 LSU cycles per 22 FP64 cycles -- the first version of this kernel ran at exactly
            // that ratio (72 % of the FP64 pipe).  Lane h of a pair tests BOTH rays of the pair against the spheres
            // 64c + 2j + h (j = 0..31) of super-chunk c: one load per test.  Per test still 3 DADD + 2 DMUL + 6 DFMA; the
            // sign bit of each discriminant goes into a 32-test mask by one funnel shift, masks go to shared memory, and
            // after the sweep every lane resolves the candidates of ITS OWN ray (its slot-0 words and its partner's slot-1
            // words) in one loop -- closest hit in the order-independent form: min over the spheres of the first root
            // >= tmin, ties to the larger list index (the sequential sweep lets the later sphere win, src/hit.jl:24-26,44-46).
            const uint32_t h = threadIdx.x & 1u;
            const d3 o1 = mkd(__shfl_xor_sync(kFullMask, o.x, 1), __shfl_xor_sync(kFullMask, o.y, 1), __shfl_xor_sync(kFullMask, o.z, 1));
            const d3 d1 = mkd(__shfl_xor_sync(kFullMask, d.x, 1), __shfl_xor_sync(kFullMask, d.y, 1), __shfl_xor_sync(kFullMask, d.z, 1));
            auto disc_hi = [&](const double4 s, const d3 ro, const d3 rd) {
                const d3 oc = mkd(ro.x - s.x, ro.y - s.y, ro.z - s.z);
                const double hb = dotd(oc, rd);
                const double cq = fma(-s.w, s.w, dotd(oc, oc));
                return (uint32_t)__double2hiint(fma(hb, hb, -cq));
            };
            uint32_t summary = 0u;  // bit 2c + q: mask word (super-chunk c, slot q) holds a candidate
            const uint32_t nfull = n >> 6;
            const double4* lp = list + h;
            uint32_t* mp = s_mask;
            for (uint32_t c = 0; c < nfull; ++c) {
                uint32_t m0 = 0u, m1 = 0u;
                // 16 spheres per unrolled body: the 32-sphere body (847 instructions) overflowed the instruction cache
#pragma unroll 1
                for (int jb = 0; jb < 64; jb += 32) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const double4 s = lp[jb + 2 * j];
                        m0 = __funnelshift_l(disc_hi(s, o, d), m0, 1);  // test j ends at bit 31 - j; set = miss
                        m1 = __funnelshift_l(disc_hi(s, o1, d1), m1, 1);
                    }
                }
                mp[0] = m0;
                mp[kTraceBlock] = m1;
                summary |= ((m0 != 0xffffffffu ? 1u : 0u) | (m1 != 0xffffffffu ? 2u : 0u)) << (2u * c);
                lp += 64;
                mp += 2 * kTraceBlock;
            }
            const uint32_t rem = n - (nfull << 6);  // 0 .. 63 spheres in the ragged last super-chunk
            if (rem != 0u) {
                uint32_t m0 = 0u, m1 = 0u;
                const uint32_t run = (rem + 1u) >> 1;  // tests every lane runs (uniform); lane h owns (rem - h + 1) / 2 of them
                for (uint32_t j = 0; j < run; ++j) {
                    const uint32_t idx = 2u * j + h;
                    const double4 s = lp[idx < rem ? 2u * j : 0u];
                    m0 = __funnelshift_l(disc_hi(s, o, d), m0, 1);
                    m1 = __funnelshift_l(disc_hi(s, o1, d1), m1, 1);
                }
                const uint32_t mine = (rem + 1u - h) >> 1;
                const uint32_t sh = 32u - run;                                 // left-align the tests that were executed
                const uint32_t fill = mine >= 32u ? 0u : (0xffffffffu >> mine);  // everything below this lane's real tests = miss
                m0 = (m0 << sh) | fill;
                m1 = (m1 << sh) | fill;
                mp[0] = m0;
                mp[kTraceBlock] = m1;
                summary |= ((m0 != 0xffffffffu ? 1u : 0u) | (m1 != 0xffffffffu ? 2u : 0u)) << (2u * nfull);
            }
            __syncwarp();  // the partner's mask words are visible
            uint32_t sum = (summary & 0x55555555u) | (__shfl_xor_sync(kFullMask, summary, 1) & 0xaaaaaaaau);
            if (!alive) sum = 0u;
            const uint32_t* mbase = s_mask - threadIdx.x;
            while (sum) {
                const uint32_t w = (uint32_t)__ffs((int)sum) - 1u;  // = 2c + q: word of lane tid ^ q, which tests spheres 64c + 2j + (h ^ q)
                sum &= sum - 1u;
                const uint32_t q = w & 1u;
                uint32_t cand = ~mbase[w * kTraceBlock + (threadIdx.x ^ q)];
                const uint32_t kb = (w >> 1) * 64u + (h ^ q);
                while (cand) {
                    const uint32_t j = (uint32_t)__clz((int)cand);
                    cand &= ~(0x80000000u >> j);
                    const uint32_t k = kb + 2u * j;
                    const double4 s = list[k];
                    const d3 oc = mkd(o.x - s.x, o.y - s.y, o.z - s.z);
                    const double hb = dotd(oc, d);
                    const double cq = fma(-s.w, s.w, dotd(oc, oc));
                    const double sq = sqrt(fma(hb, hb, -cq));
                    const double r1 = -hb - sq, r2 = -hb + sq;  // src/hit.jl:23, 25
                    const double t = r1 < tmin ? r2 : r1;       // the first root >= tmin, if any
                    if (t >= tmin && (t < best_t || (t == best_t && (int)k > best_k))) {
                        best_t = t;
                        best_k = (int)k;
                    }
                }
            }
            __syncwarp();  // all reads of the partner's words are done before the next sweep overwrites them
        } else {
        uint32_t k = 0;
        // streamed lists (> 1024 spheres, read through L1/L2): whole chunks of 32 spheres, branch-free discriminants,
        // candidates resolved chunk by chunk in list order
        for (; k + 32u <= n; k += 32u) {
            uint32_t m = 0u;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const double4 s = list[k + j];
                const d3 oc = mkd(o.x - s.x, o.y - s.y, o.z - s.z);
                const double hb = dotd(oc, d);
                const double cq = fma(-s.w, s.w, dotd(oc, oc));
                const double disc = fma(hb, hb, -cq);
                m = __funnelshift_l((uint32_t)__double2hiint(disc), m, 1);  // test j ends at bit 31 - j; set = miss
            }
            uint32_t cand = alive ? ~m : 0u;
            while (cand) {
                const uint32_t j = (uint32_t)__clz((int)cand);
                cand &= ~(0x80000000u >> j);
                resolve(k + j);
            }
        }
        if (alive)
            for (; k < n; ++k) resolve(k);  // ragged tail
        }
    }
    for (int off = 16; off > 0; off >>= 1) seg_count += __shfl_xor_sync(kFullMask, seg_count, off);
    if (lane == 0 && seg_count) atomicAdd(P.counters + 1, (unsigned long long)seg_count);
}

// ---- the latency path in Float64 (the reference's own smoke test renders Float64: test/runtests.jl:188-194) ------------
// One launch = one whole small render, as small_render_kernel (rtw_small.cu) does for Float32: a pixel is owned by a
// group of 2^group_log2 adjacent lanes, lane j traces samples j, j + g, ... one path at a time against the list in shared
// memory (list order, sequential rejection loops -- the code of the reference, line by line), the group adds the
// fixed-point integers of its paths by shuffle and lane 0 writes sqrt(sum / spp) into mapped host memory; the last CTA
// publishes the ray-segment count and re-zeroes the device counters.
constexpr int kSmall64Block = 64;

__global__ void __launch_bounds__(kSmall64Block) small_render_f64_kernel(const __grid_constant__ TraceParams64 P, int group_log2,
                                                                         double inv_scale, double* __restrict__ out_img,
                                                                         unsigned long long* __restrict__ host_totals) {
    extern __shared__ __align__(16) unsigned char smem_small64[];
    double4* s_list = reinterpret_cast<double4*>(smem_small64);
    const uint32_t n = P.n_spheres;
    for (uint32_t i = threadIdx.x; i < n; i += kSmall64Block) s_list[i] = P.geom[i];
    __syncthreads();
    const uint32_t g = 1u << group_log2;
    const unsigned long long gtid = (unsigned long long)blockIdx.x * kSmall64Block + threadIdx.x;
    const unsigned long long npix = (unsigned long long)P.n_rows * (unsigned long long)P.W;
    const unsigned long long pl = gtid >> group_log2;
    const uint32_t j = (uint32_t)gtid & (g - 1u);
    const bool active = pl < npix;
    const uint32_t k0 = P.key0, k1 = P.key1;
    const double tmin = 1e-4;  // T(1e-4), src/ray_color.jl:19
    long long acc_r = 0, acc_g = 0, acc_b = 0;
    uint32_t seg_count = 0, i0 = 0, col = 0;
    if (active) {
        const uint32_t row_local = (uint32_t)(pl / (unsigned)P.W);
        col = (uint32_t)(pl - (unsigned long long)row_local * (unsigned)P.W);
        i0 = (uint32_t)P.row_start + row_local * (uint32_t)P.row_stride;
        PathRng rng;
        rng.pixel = i0 * (uint32_t)P.W + col;
        for (uint32_t s0 = j + (uint32_t)P.sample_first; s0 < (uint32_t)(P.sample_first + P.spp); s0 += g) {  // src/render.jl:29
            rng.sample = s0;
            double su = (double)(col + 1u) / (double)P.W;                  // u = T(j/W), src/render.jl:26
            double sv = (double)((uint32_t)P.H - 1u - i0) / (double)P.H;   // v = T((H-i)/H), src/render.jl:27
            if (s0 != 0u) {  // src/render.jl:30-36: the first sample is centred; du = draw 0, dv = draw 1
                const u32x4 b = philox_block(rng, 0u, 0u, k0, k1);
                su += u01d(b.w0, b.w1) / (double)(float)P.W;
                sv += u01d(b.w2, b.w3) / (double)(float)P.H;
            }
            double px, py;  // get_ray, src/camera.jl:43-48; random_vec2_in_disk (src/rand.jl:31-38) is always drawn
            for (uint32_t k = 0;; ++k) {
                const u32x4 b = philox_block(rng, 0u, 1u + k, k0, k1);
                px = pm1d(b.w0, b.w1);
                py = pm1d(b.w2, b.w3);
                if (fma(py, py, px * px) <= 1.0) break;
            }
            const DevCamera64& c = P.cam;
            const double rx = c.lens_radius * px, ry = c.lens_radius * py;
            const d3 off = mkd(fma(c.v[0], ry, c.u[0] * rx), fma(c.v[1], ry, c.u[1] * rx), fma(c.v[2], ry, c.u[2] * rx));
            d3 o = mkd(c.origin[0] + off.x, c.origin[1] + off.y, c.origin[2] + off.z);
            d3 q3;
            q3.x = fma(sv, c.vertical[0], fma(su, c.horizontal[0], c.llc[0])) - c.origin[0] - off.x;
            q3.y = fma(sv, c.vertical[1], fma(su, c.horizontal[1], c.llc[1])) - c.origin[1] - off.y;
            q3.z = fma(sv, c.vertical[2], fma(su, c.horizontal[2], c.llc[2])) - c.origin[2] - off.z;
            d3 d = normalized(q3);
            double thr_r = 1.0, thr_g = 1.0, thr_b = 1.0;
            for (uint32_t nhits = 0;;) {  // ray_color, src/ray_color.jl:14-38, as a loop
                if ((int)nhits >= P.max_depth) break;
                double best_t = __longlong_as_double(0x7ff0000000000000ll);  // typemax(T)
                int best_k = -1;
                for (uint32_t k = 0; k < n; ++k) {  // hit(::HittableList), src/hit.jl:38-50 + hit(::Sphere), :12-35
                    const double4 s = s_list[k];
                    const d3 oc = mkd(o.x - s.x, o.y - s.y, o.z - s.z);
                    const double hb = dotd(oc, d);
                    const double cq = fma(-s.w, s.w, dotd(oc, oc));
                    const double disc = fma(hb, hb, -cq);
                    if (disc < 0.0) continue;
                    const double sq = sqrt(disc);
                    double root = -hb - sq;
                    if (root < tmin || best_t < root) {
                        root = -hb + sq;
                        if (root < tmin || best_t < root) continue;
                    }
                    best_t = root;
                    best_k = (int)k;
                }
                seg_count += 1;
                if (best_k < 0) {  // skycolor, src/ray_color.jl:1-6
                    const double t = 0.5 * (d.y + 1.0);
                    const double a = 1.0 - t;
                    const double sr = a + t * 0.5, sg = a + t * 0.7, sb = a + t;
                    acc_r += __double2ll_rn(thr_r * sr * P.fx_scale);
                    acc_g += __double2ll_rn(thr_g * sg * P.fx_scale);
                    acc_b += __double2ll_rn(thr_b * sb * P.fx_scale);
                    break;
                }
                nhits += 1;
                if ((int)nhits >= P.max_depth) break;  // the next ray_color call returns black
                const double4 gs = s_list[best_k];
                const double4 m = P.mat[best_k];
                const uint32_t kind = __ldg(P.kind + best_k);
                const d3 p = mkd(fma(best_t, d.x, o.x), fma(best_t, d.y, o.y), fma(best_t, d.z, o.z));  // hit.jl:3
                const d3 on = mkd((p.x - gs.x) / gs.w, (p.y - gs.y) / gs.w, (p.z - gs.z) / gs.w);       // hit.jl:33
                const bool front = dotd(d, on) < 0.0;                                                  // hit.jl:7
                const d3 nn = front ? on : mkd(-on.x, -on.y, -on.z);
                d3 nd;
                if (kind != 2u) {  // Lambertian / Metal: normalize(random_vec3_in_sphere), src/rand.jl:15-22,29
                    d3 rv;
                    for (uint32_t a = 0;; ++a) {
                        const u32x4 b0 = philox_block(rng, nhits, 2u * a, k0, k1);
                        const u32x4 b1 = philox_block(rng, nhits, 2u * a + 1u, k0, k1);
                        const d3 q = mkd(pm1d(b0.w0, b0.w1), pm1d(b0.w2, b0.w3), pm1d(b1.w0, b1.w1));
                        if (dotd(q, q) <= 1.0) { rv = normalized(q); break; }
                    }
                    if (kind == 0u) {  // src/material.jl:13-23
                        const d3 sd = mkd(nn.x + rv.x, nn.y + rv.y, nn.z + rv.z);
                        nd = dotd(sd, sd) < 1e-5 ? nn : normalized(sd);  // near_zero, src/vec.jl:20
                    } else {  // src/material.jl:31-34
                        const d3 refl = reflectd(d, nn);
                        nd = normalized(mkd(fma(m.w, rv.x, refl.x), fma(m.w, rv.y, refl.y), fma(m.w, rv.z, refl.z)));
                    }
                    thr_r *= m.x;
                    thr_g *= m.y;
                    thr_b *= m.z;
                } else {  // Dielectric, src/material.jl:41-53
                    const double ratio = front ? 1.0 / m.w : m.w;
                    const double cos_t = fmin(-dotd(d, nn), 1.0);
                    const double sin_t = sqrt(fma(-cos_t, cos_t, 1.0));
                    bool reflects = ratio * sin_t > 1.0;
                    if (!reflects) {  // `||` short-circuits: the coin (draw 3 of the event) only when refraction is possible
                        const u32x4 b = philox_block(rng, nhits, 1u, k0, k1);
                        double r0 = (1.0 - ratio) / (1.0 + ratio);  // Schlick, src/light.jl:19-25
                        r0 = r0 * r0;
                        const double x = 1.0 - cos_t, x2 = x * x, x4 = x2 * x2;
                        reflects = fma(1.0 - r0, x4 * x, r0) > u01d(b.w2, b.w3);
                    }
                    if (reflects) {
                        nd = reflectd(d, nn);  // not re-normalised, src/material.jl:48
                    } else {  // refract, src/light.jl:12-17
                        const d3 perp = mkd(ratio * fma(cos_t, nn.x, d.x), ratio * fma(cos_t, nn.y, d.y),
                                            ratio * fma(cos_t, nn.z, d.z));
                        const double sp = sqrt(fabs(1.0 - dotd(perp, perp)));
                        nd = normalized(mkd(fma(-sp, nn.x, perp.x), fma(-sp, nn.y, perp.y), fma(-sp, nn.z, perp.z)));
                    }
                }
                o = p;
                d = nd;
            }
        }
    }
    for (uint32_t off = g >> 1; off > 0; off >>= 1) {  // integer sums: order-independent
        acc_r += __shfl_xor_sync(0xffffffffu, acc_r, off);
        acc_g += __shfl_xor_sync(0xffffffffu, acc_g, off);
        acc_b += __shfl_xor_sync(0xffffffffu, acc_b, off);
    }
    if (active && j == 0u) {
        const long long at = ((long long)col * P.H + (long long)i0) * 3;  // Julia column-major Matrix{RGB{Float64}}(H, W)
        const long long a[3] = {acc_r, acc_g, acc_b};
#pragma unroll
        for (int c = 0; c < 3; ++c) out_img[at + c] = sqrt((double)a[c] * inv_scale / (double)P.spp);  // render.jl:40, vec.jl:22
    }
    for (int off = 16; off > 0; off >>= 1) seg_count += __shfl_xor_sync(0xffffffffu, seg_count, off);
    if ((threadIdx.x & 31) == 0 && seg_count) atomicAdd(P.counters + 1, (unsigned long long)seg_count);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long done = atomicAdd(P.counters, 1ull) + 1ull;
        if (done == gridDim.x) {
            __threadfence();
            host_totals[0] = atomicExch(P.counters + 1, 0ull);
            atomicExch(P.counters, 0ull);
            __threadfence_system();
        }
    }
}

// accum / n_samples -> sqrt (src/render.jl:40, src/vec.jl:22), Float64 out: tile row-major or Julia column-major
__global__ void __launch_bounds__(256) resolve_f64_kernel(const unsigned long long* __restrict__ accum, int W, int H,
                                                          int n_rows, int row_start, int row_stride, int spp,
                                                          double inv_scale, int column_major, double* __restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_rows * W) return;
    int k, col;
    if (column_major) { col = (int)(t / n_rows); k = (int)(t - (long long)col * n_rows); }
    else { k = (int)(t / W); col = (int)(t - (long long)k * W); }
    const unsigned long long* a = accum + ((long long)k * W + col) * 4;
    const long long at = column_major ? ((long long)col * H + (row_start + (long long)k * row_stride)) * 3
                                      : ((long long)k * W + col) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) out[at + c] = sqrt((double)(long long)a[c] * inv_scale / (double)spp);
}

__global__ void __launch_bounds__(256) assemble_f64_kernel(const double* __restrict__ tiles, int G, int W, int H,
                                                           int rows_pad, double* __restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)W * H) return;
    const int col = (int)(t / H), i0 = (int)(t - (long long)col * H);
    const int g = i0 % G, k = i0 / G;
    const double* src = tiles + (((long long)g * rows_pad + k) * W + col) * 3;
    out[t * 3 + 0] = src[0];
    out[t * 3 + 1] = src[1];
    out[t * 3 + 2] = src[2];
}

}  // namespace

cudaError_t launch_trace_f64(const TraceParams64& p, int num_sms, cudaStream_t stream, LaunchInfo* info) {
    const bool shared = p.n_spheres <= kTileSpheres;
    // shared: the list + (per lane) two mask words per super-chunk of 64 spheres
    const int smem = shared ? (int)(p.n_spheres * sizeof(double4) + 2u * (p.n_spheres / 64u + 1u) * kTraceBlock * 4u) : 0;
    cudaError_t e = cudaSuccess;
    int per_sm = 0;
    if (shared) {
        e = cudaFuncSetAttribute(trace_f64_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trace_f64_kernel<true>, kTraceBlock, smem);
    } else {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trace_f64_kernel<false>, kTraceBlock, 0);
    }
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)num_sms * per_sm;
    const long long max_useful = (long long)((p.n_paths + (unsigned long long)kTraceBlock - 1ull) / (unsigned long long)kTraceBlock);
    if (grid > max_useful) grid = max_useful;
    if (grid < 1) grid = 1;
    if (shared) trace_f64_kernel<true><<<(unsigned)grid, kTraceBlock, smem, stream>>>(p);
    else trace_f64_kernel<false><<<(unsigned)grid, kTraceBlock, 0, stream>>>(p);
    if (info) {
        info->grid = (int)grid;
        info->block = kTraceBlock;
        info->smem_bytes = smem;
        info->blocks_per_sm = per_sm;
        info->launches = 1;
        info->rays_per_lane = 1;
        info->sweep = kSweepBranch;
    }
    return cudaGetLastError();
}

cudaError_t launch_small_render_f64(const TraceParams64& p, double inv_scale, double* out_img, unsigned long long* host_totals,
                                    cudaStream_t stream, LaunchInfo* info) {
    int group_log2 = 0;
    while ((1 << group_log2) < p.spp && group_log2 < 5) ++group_log2;
    const unsigned long long threads = ((unsigned long long)p.n_rows * (unsigned long long)p.W) << group_log2;
    const unsigned long long blocks = (threads + kSmall64Block - 1) / kSmall64Block;
    if (blocks == 0 || blocks > 0x7fffffffull) return cudaErrorInvalidValue;
    const int smem = (int)(p.n_spheres * sizeof(double4));
    small_render_f64_kernel<<<(unsigned)blocks, kSmall64Block, smem, stream>>>(p, group_log2, inv_scale, out_img, host_totals);
    if (info) {
        info->grid = (int)blocks;
        info->block = kSmall64Block;
        info->smem_bytes = smem;
        info->blocks_per_sm = 0;
        info->launches = 1;
        info->rays_per_lane = 1;
        info->sweep = kSweepBranch;
    }
    return cudaGetLastError();
}

cudaError_t launch_resolve_f64(const unsigned long long* accum, int W, int H, int n_rows, int row_start, int row_stride,
                               int spp, double inv_scale, int column_major, double* out, cudaStream_t stream) {
    const long long total = (long long)n_rows * W;
    if (total <= 0) return cudaSuccess;
    resolve_f64_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(accum, W, H, n_rows, row_start, row_stride, spp,
                                                                             inv_scale, column_major, out);
    return cudaGetLastError();
}

cudaError_t launch_assemble_f64(const double* tiles, int n_tiles, int W, int H, double* out, cudaStream_t stream) {
    const long long total = (long long)W * H;
    if (total <= 0) return cudaSuccess;
    const int rows_pad = (H + n_tiles - 1) / n_tiles;
    assemble_f64_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(tiles, n_tiles, W, H, rows_pad, out);
    return cudaGetLastError();
}

}  // namespace rtw
