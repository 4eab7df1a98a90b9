// rtw_grid.cuh -- RTW_MODE_GRID: closest hit through a uniform grid instead of the linear sweep.
//
// The reference is brute force by design (README.md:30) and lists acceleration structures as a long-term goal
// (README.md:175); the linear sweep stays the default and the benchmarked path.  This mode returns THE SAME closest
// hit -- and therefore the same image bits -- by testing fewer spheres: hit(::HittableList) (src/hit.jl:38-50) is
// the minimum over the list of each sphere's first root >= tmin, ties to the larger list index (inclusive range
// tests, src/hit.jl:24,26), and a minimum does not care which non-minimal elements are skipped.  Every sphere that
// IS tested goes through the identical arithmetic (src/hit.jl:13-29, DESIGN.md FP contract).
//
// Structure (built on the host, csrc/rtw_capi.cu build_grid): spheres much larger than the typical one ("big":
// |r| > 4 x median, e.g. the r = 1000 ground) are kept in a list that every ray tests; the others are binned into a
// uniform grid over their bounding box by their AABB inflated by 5 % of a cell edge, so a sphere is registered in
// every cell its surface can reach, with slack far above the rounding of the traversal.  A ray walks the cells it
// crosses front to back (3-D DDA) and stops as soon as the closest hit so far lies before the exit of the current
// cell: any sphere registered only in later cells is at least the inflation margin beyond that exit.
//
// One subtlety makes "the same closest hit" more than geometry.  hit(::Sphere) assumes a unit direction (a = 1,
// src/hit.jl:14-15), but scatter(::Dielectric) returns reflect(...) un-normalised (src/material.jl:48) and in Float32
// the normal (p - c)/r of a small far sphere is off by up to ~1e-4, so |d|^2 = 1 + eps with eps up to ~1e-3 after a
// glass reflection (and ~3e-7 by plain rounding otherwise).  With a = 1 + eps the reference's discriminant
// hb^2 - c accepts spheres whose distance rho from the ray line satisfies rho^2 <= r^2 + eps*m^2 (m = distance along
// the ray): far spheres grow.  The linear sweep reproduces that arithmetic by construction; the grid only sees
// spheres near the geometric ray.  So the traversal is used only while eps*t_exit^2 stays below what the
// registration margin covers ((r_min + inflate)^2 - r_min^2, GridParams::safe2), where t_exit is the closest hit
// the traversal found (no unseen sphere beyond it can win) or the exit of the grid; a ray beyond that is reported
// as unsafe and the caller resolves it by a warp-cooperative sweep of the whole list (lists <= kGridFallbackMax).
#pragma once
#include "rtw_kernels.h"

namespace rtw {
namespace {

__device__ __forceinline__ void grid_test_sphere(const float4 s, uint32_t k, const f3 o, const f3 d, float& bt, int& bk) {
    const float tmin = 1e-4f;  // T(1e-4), src/ray_color.jl:19
    const float ocx = o.x - s.x, ocy = o.y - s.y, ocz = o.z - s.z;                                  // src/hit.jl:13
    const float hb = fmaf(ocz, d.z, fmaf(ocy, d.y, ocx * d.x));                                      // :16
    const float cq = fmaf(-s.w, s.w, fmaf(ocz, ocz, fmaf(ocy, ocy, ocx * ocx)));                     // :17
    const float disc = fmaf(hb, hb, -cq);                                                            // :18
    if (disc < 0.0f) return;                                                                         // :19
    if (hb > 0.0f && cq > 0.0f) return;  // wholly behind the origin: both roots <= 0 < tmin (sqrt(disc) <= hb)
    const float sq = __fsqrt_rn(disc);
    const float r1 = -hb - sq, r2 = -hb + sq;                                                        // :23, :25
    const float t = r1 < tmin ? r2 : r1;  // the first root >= tmin, if any
    if (t < tmin) return;
    if (t < bt || (t == bt && (int)k > bk)) {  // ties: the later sphere wins, as in the sequential sweep
        bt = t;
        bk = (int)k;
    }
}

// returns true when the ray is "unsafe" for the grid (see above): the result is then NOT final and the caller must
// run grid_fallback_sweep
__device__ __forceinline__ bool closest_hit_grid(const GridParams& G, const float4* __restrict__ geom, const f3 o,
                                                 const f3 d, const bool alive, float& best_t, int& best_k) {
    float bt = __int_as_float(0x7f800000);  // typemax(T) = Inf, src/ray_color.jl:19
    int bk = -1;
    bool unsafe = false;
    if (alive) {
        for (uint32_t i = 0; i < G.n_big; ++i) {
            const uint32_t k = __ldg(G.big + i);
            grid_test_sphere(__ldg(geom + k), k, o, d, bt, bk);
        }
        if (G.nx > 0) {
            // ray against the grid box (slabs); a zero direction component gives +-Inf (or NaN when the origin lies
            // on the plane, which the min/max below ignore)
            const float ix = __fdiv_rn(1.0f, d.x), iy = __fdiv_rn(1.0f, d.y), iz = __fdiv_rn(1.0f, d.z);
            const float hx = G.h * (float)G.nx, hy = G.h * (float)G.ny, hz = G.h * (float)G.nz;
            const float ax = (G.ox - o.x) * ix, bx = (G.ox + hx - o.x) * ix;
            const float ay = (G.oy - o.y) * iy, by = (G.oy + hy - o.y) * iy;
            const float az = (G.oz - o.z) * iz, bz = (G.oz + hz - o.z) * iz;
            float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
            const float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
            // enter a little early / accept a little late: the cells are clamped, so slack only costs a cell
            // |d|^2 - 1, with a floor for the rounding of a normalised Float32 vector
            const float eps = fmaxf(fabsf(fmaf(d.z, d.z, fmaf(d.y, d.y, d.x * d.x)) - 1.0f), 4e-7f);
            const bool enters = t0 <= t1 * 1.0001f + 1e-4f && t0 < bt;
            if (enters) {
                const float px = fmaf(t0, d.x, o.x), py = fmaf(t0, d.y, o.y), pz = fmaf(t0, d.z, o.z);
                int cx = min(max((int)floorf((px - G.ox) * G.inv_h), 0), G.nx - 1);
                int cy = min(max((int)floorf((py - G.oy) * G.inv_h), 0), G.ny - 1);
                int cz = min(max((int)floorf((pz - G.oz) * G.inv_h), 0), G.nz - 1);
                const int sx = d.x > 0.0f ? 1 : -1, sy = d.y > 0.0f ? 1 : -1, sz = d.z > 0.0f ? 1 : -1;
                const float inf = __int_as_float(0x7f800000);
                // parameter at which the ray leaves the current cell along each axis
                float tx = d.x != 0.0f ? (G.ox + G.h * (float)(cx + (sx > 0 ? 1 : 0)) - o.x) * ix : inf;
                float ty = d.y != 0.0f ? (G.oy + G.h * (float)(cy + (sy > 0 ? 1 : 0)) - o.y) * iy : inf;
                float tz = d.z != 0.0f ? (G.oz + G.h * (float)(cz + (sz > 0 ? 1 : 0)) - o.z) * iz : inf;
                const float dtx = d.x != 0.0f ? G.h * fabsf(ix) : inf;
                const float dty = d.y != 0.0f ? G.h * fabsf(iy) : inf;
                const float dtz = d.z != 0.0f ? G.h * fabsf(iz) : inf;
                for (;;) {
                    const uint32_t cell = (uint32_t)cx + (uint32_t)G.nx * ((uint32_t)cy + (uint32_t)G.ny * (uint32_t)cz);
                    const uint32_t first = __ldg(G.cell_start + cell), last = __ldg(G.cell_start + cell + 1u);
                    for (uint32_t i = first; i < last; ++i) {
                        const uint32_t k = __ldg(G.items + i);
                        grid_test_sphere(__ldg(geom + k), k, o, d, bt, bk);
                    }
                    const float texit = fminf(tx, fminf(ty, tz));
                    if (bt < texit) break;  // nothing registered only in later cells can be closer (registration margin)
                    if (tx <= ty && tx <= tz) {
                        cx += sx;
                        if ((unsigned)cx >= (unsigned)G.nx) break;
                        tx += dtx;
                    } else if (ty <= tz) {
                        cy += sy;
                        if ((unsigned)cy >= (unsigned)G.ny) break;
                        ty += dty;
                    } else {
                        cz += sz;
                        if ((unsigned)cz >= (unsigned)G.nz) break;
                        tz += dtz;
                    }
                }
            }
            // Safety of the answer (see the header): how far along the ray a sphere the traversal did not see could
            // still matter -- up to the closest hit found (plus the reach of a small sphere), at most the exit of the
            // box; for a ray that misses the box with a visibly non-unit direction, the far side of the box's
            // bounding ball.
            float t_far = t1;
            if (!enters) {
                const float mx = G.ox + 0.5f * hx - o.x, my = G.oy + 0.5f * hy - o.y, mz = G.oz + 0.5f * hz - o.z;
                t_far = eps > 1e-6f ? sqrtf(mx * mx + my * my + mz * mz) + 0.5f * sqrtf(hx * hx + hy * hy + hz * hz) : 0.0f;
            }
            t_far = fminf(t_far, bt + 2.0f * G.h);
            unsafe = eps * t_far * t_far > G.safe2;
        }
    }
    best_t = bt;
    best_k = bk;
    return unsafe;
}

// Exact closest hit for the unsafe rays of a warp: one ray at a time, the 32 lanes split the list, the partial
// results are reduced with the tie rule (equal t: larger index).  Called by all lanes of the warp.
// Returns the number of rays of the warp it resolved (uniform).
__device__ __forceinline__ uint32_t grid_fallback_sweep(const float4* __restrict__ geom, uint32_t n, const f3 o, const f3 d,
                                                        bool unsafe, float& best_t, int& best_k) {
    const unsigned lane = threadIdx.x & 31u;
    unsigned pending = __ballot_sync(0xffffffffu, unsafe);
    const uint32_t resolved = (uint32_t)__popc(pending);
    while (pending) {
        const int src = __ffs((int)pending) - 1;
        pending &= pending - 1u;
        const f3 ro = mk3(__shfl_sync(0xffffffffu, o.x, src), __shfl_sync(0xffffffffu, o.y, src), __shfl_sync(0xffffffffu, o.z, src));
        const f3 rd = mk3(__shfl_sync(0xffffffffu, d.x, src), __shfl_sync(0xffffffffu, d.y, src), __shfl_sync(0xffffffffu, d.z, src));
        float bt = __int_as_float(0x7f800000);
        int bk = -1;
        for (uint32_t k = lane; k < n; k += 32u) grid_test_sphere(__ldg(geom + k), k, ro, rd, bt, bk);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float pt = __shfl_xor_sync(0xffffffffu, bt, off);
            const int pk = __shfl_xor_sync(0xffffffffu, bk, off);
            if (pk >= 0 && (bk < 0 || pt < bt || (pt == bt && pk > bk))) {
                bt = pt;
                bk = pk;
            }
        }
        if ((int)lane == src) {
            best_t = bt;
            best_k = bk;
        }
    }
    return resolved;
}

}  // namespace
}  // namespace rtw
