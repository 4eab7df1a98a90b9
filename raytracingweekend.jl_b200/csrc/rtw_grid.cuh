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
// the traversal found (no unseen sphere beyond it can win) or the exit of the grid; along a long flight the walk
// switches to a second, looser registration of the same cells, and a ray that even that cannot cover is resolved by a
// warp-cooperative sweep of the list -- for every list size, so the mode is exact by construction.
#pragma once
#include "rtw_kernels.h"

namespace rtw {
namespace {

constexpr float kGridEpsFloor = 1e-6f;  // see grid_walk_cells; the host sizes the loose margin with the same value

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

// Closest hit through the grid.  Returns true when the ray is "unsafe": the answer is then NOT final and the caller
// must run grid_fallback_sweep (which starts from the hit found so far -- every hit found is a legitimate candidate).
//
// The walk is front to back (3-D DDA) and stops when the closest hit so far lies before the exit of the current cell.
// A sphere that could still beat that hit would have its hit point inside a cell already visited, so what has to
// hold is: every sphere whose APPARENT extent (radius sqrt(r^2 + eps m^2) at distance m, header) reaches a visited cell
// was registered there.  The margin that takes grows with the distance, so the registration is chosen per cell from
// the parameter at which the ray leaves it: tight while eps t^2 stays below safe2_tight, loose while below
// safe2_loose, and beyond that the ray is unsafe.  Work counters of this lane: `cells` walked (`loose_cells` of them with
// the loose lists) and sphere `tests` made (big spheres + the lists of the cells).
__device__ __forceinline__ bool closest_hit_grid(const GridParams& G, const float4* __restrict__ geom, const f3 o,
                                                 const f3 d, const bool alive, float& best_t, int& best_k,
                                                 uint32_t& loose_cells, uint32_t& cells, uint32_t& tests) {
    float bt = __int_as_float(0x7f800000);  // typemax(T) = Inf, src/ray_color.jl:19
    int bk = -1;
    bool unsafe = false;
    if (alive) {
        for (uint32_t i = 0; i < G.n_big; ++i) {
            const uint32_t k = __ldg(G.big + i);
            grid_test_sphere(__ldg(geom + k), k, o, d, bt, bk);
        }
        tests += G.n_big;
    }
    if (alive && G.nx > 0) {
        // ray against the grid box (slabs); a zero direction component gives +-Inf (or NaN when the origin lies
        // on the plane, which the min/max below ignore)
        const float ix = __fdiv_rn(1.0f, d.x), iy = __fdiv_rn(1.0f, d.y), iz = __fdiv_rn(1.0f, d.z);
        const float hx = G.h * (float)G.nx, hy = G.h * (float)G.ny, hz = G.h * (float)G.nz;
        const float ax = (G.ox - o.x) * ix, bx = (G.ox + hx - o.x) * ix;
        const float ay = (G.oy - o.y) * iy, by = (G.oy + hy - o.y) * iy;
        const float az = (G.oz - o.z) * iz, bz = (G.oz + hz - o.z) * iz;
        const float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
        const float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
        // eps = | |d|^2 - 1 | + kGridEpsFloor: the floor stands for the ROUNDING of the reference's discriminant, which
        // for a sphere m away is computed from terms of size m^2 (half_b^2 and oc.oc, three roundings each, on a
        // rounded oc): up to ~8e-7 m^2 of absolute error in hb^2 - c, the growth of r^2 of a direction with
        // |d|^2 = 1 + 8e-7
        const float dd = fmaf(d.z, d.z, fmaf(d.y, d.y, d.x * d.x));
        const float eps = fabsf(dd - 1.0f) + kGridEpsFloor;
        // a sphere whose centre projects to m along the ray offers no root below m (sqrt(1 + eps) - sqrt(eps)) - r:
        // nothing beyond t (1 + 3 sqrt(eps)) + reach can beat a hit at t
        const float seps = sqrtf(eps);
        const float stretch = fmaf(3.0f, seps, 1.0f);
        // no centre projects beyond the far side of the box's bounding ball; over that range a sphere's apparent radius
        // exceeds its real one by at most grow_far
        const float mx = G.ox + 0.5f * hx - o.x, my = G.oy + 0.5f * hy - o.y, mz = G.oz + 0.5f * hz - o.z;
        const float ball_far = (sqrtf(mx * mx + my * my + mz * mz) + G.ball_r) * 1.001f;
        // enter a little early / accept a little late: the cells are clamped, so slack only costs a cell
        const bool enters = t0 <= t1 * 1.0001f + 1e-4f && t0 < bt;
        if (!(eps <= 0.01f)) {
            unsafe = true;  // far from unit length: no margin argument holds
        } else if (enters) {
            const float px = fmaf(t0, d.x, o.x), py = fmaf(t0, d.y, o.y), pz = fmaf(t0, d.z, o.z);
            int cx = min(max((int)floorf((px - G.ox) * G.inv_h), 0), G.nx - 1);
            int cy = min(max((int)floorf((py - G.oy) * G.inv_h), 0), G.ny - 1);
            int cz = min(max((int)floorf((pz - G.oz) * G.inv_h), 0), G.nz - 1);
            const int sx = d.x > 0.0f ? 1 : -1, sy = d.y > 0.0f ? 1 : -1, sz = d.z > 0.0f ? 1 : -1;
            const float inf = __int_as_float(0x7f800000);
            // parameter at which the ray leaves the current cell along each axis (recomputed from the cell index at
            // every step, not accumulated: over the hundreds of cells of a long flight an accumulated parameter drifts
            // by a visible fraction of the tight margin)
            const int ux = sx > 0 ? 1 : 0, uy = sy > 0 ? 1 : 0, uz = sz > 0 ? 1 : 0;
            float tx = d.x != 0.0f ? (G.ox + G.h * (float)(cx + ux) - o.x) * ix : inf;
            float ty = d.y != 0.0f ? (G.oy + G.h * (float)(cy + uy) - o.y) * iy : inf;
            float tz = d.z != 0.0f ? (G.oz + G.h * (float)(cz + uz) - o.z) * iz : inf;
            // largest exit parameter for which the tight / the loose registration is enough: eps (t stretch + 2 reach)^2
            // <= safe2  <=>  t <= (sqrt(safe2 / eps) - 2 reach) / stretch
            const float inv_seps = 1.0f / seps, inv_stretch = 1.0f / stretch;
            const float t_tight = (sqrtf(G.safe2_tight) * inv_seps - 2.0f * G.reach) * inv_stretch;
            const float t_loose = (sqrtf(G.safe2_loose) * inv_seps - 2.0f * G.reach) * inv_stretch;
            float exit_slope = 1.0f;  // component of the unit direction along the normal of the face the ray left through
            bool left_box = false;
            for (;;) {
                const float texit = fminf(tx, fminf(ty, tz));
                const float tneed = fminf(texit, bt);  // nothing beyond the hit found matters
                const uint32_t cell = (uint32_t)cx + (uint32_t)G.nx * ((uint32_t)cy + (uint32_t)G.ny * (uint32_t)cz);
                const bool tight = tneed <= t_tight;
                if (!tight && !(tneed <= t_loose)) {
                    unsafe = true;
                    break;
                }
                const uint32_t* __restrict__ cs = tight ? G.cell_start_tight : G.cell_start_loose;
                const uint32_t* __restrict__ it = tight ? G.items_tight : G.items_loose;
                loose_cells += tight ? 0u : 1u;
                const uint32_t first = __ldg(cs + cell), last = __ldg(cs + cell + 1u);
                cells += 1u;
                tests += last - first;
                for (uint32_t i = first; i < last; ++i) {
                    const uint32_t k = __ldg(it + i);
                    grid_test_sphere(__ldg(geom + k), k, o, d, bt, bk);
                }
                if (bt < texit) break;  // nothing registered only in later cells can be closer
                if (tx <= ty && tx <= tz) {
                    cx += sx;
                    if ((unsigned)cx >= (unsigned)G.nx) { left_box = true; exit_slope = fabsf(d.x); break; }
                    tx = (G.ox + G.h * (float)(cx + ux) - o.x) * ix;
                } else if (ty <= tz) {
                    cy += sy;
                    if ((unsigned)cy >= (unsigned)G.ny) { left_box = true; exit_slope = fabsf(d.y); break; }
                    ty = (G.oy + G.h * (float)(cy + uy) - o.y) * iy;
                } else {
                    cz += sz;
                    if ((unsigned)cz >= (unsigned)G.nz) { left_box = true; exit_slope = fabsf(d.z); break; }
                    tz = (G.oz + G.h * (float)(cz + uz) - o.z) * iz;
                }
            }
            if (left_box && !unsafe) {
                // Past the exit the ray recedes from the box at least at exit_slope per unit of parameter, while the
                // apparent radius of a far sphere grows like sqrt(eps) per unit: unless the ray outruns that cone (or
                // the spheres never grow out of the pad at all), a sphere beyond the exit could still be hit.
                const float grow_far = sqrtf(fmaf(eps * ball_far, ball_far, G.r_min * G.r_min)) - G.r_min;
                if (grow_far > G.pad && !(exit_slope > 2.0f * seps)) unsafe = true;
            }
        } else {
            // The ray does not walk the box (it misses it, or the hit found lies before the entry) and has seen nothing.
            // The box holds the real spheres with G.pad to spare: if their growth over the range that matters fits in
            // the pad, or the ray also misses the box grown by the difference (or reaches it only behind the hit
            // found), no sphere can be hit.  (Such a ray can owe a hit to a sphere whose apparent size reaches OUT of
            // the box -- a path that got inside the r = 1000 ground sphere and comes back up from 1900 units below.)
            const float t_far = fminf(ball_far, bt * stretch + 2.0f * G.reach);
            const float grow = sqrtf(fmaf(eps * t_far, t_far, G.r_min * G.r_min)) - G.r_min;
            const float e = grow - G.pad;
            if (e > 0.0f) {
                const float ax2 = (G.ox - e - o.x) * ix, bx2 = (G.ox + hx + e - o.x) * ix;
                const float ay2 = (G.oy - e - o.y) * iy, by2 = (G.oy + hy + e - o.y) * iy;
                const float az2 = (G.oz - e - o.z) * iz, bz2 = (G.oz + hz + e - o.z) * iz;
                const float u0 = fmaxf(fmaxf(fminf(ax2, bx2), fminf(ay2, by2)), fmaxf(fminf(az2, bz2), 0.0f));
                const float u1 = fminf(fminf(fmaxf(ax2, bx2), fmaxf(ay2, by2)), fmaxf(az2, bz2));
                if (u0 <= u1 * 1.0001f + 1e-4f && u0 < bt * stretch + G.reach) unsafe = true;
            }
        }
    }
    best_t = bt;
    best_k = bk;
    return unsafe;
}

// Exact closest hit for the unsafe rays of a warp: one ray at a time, the 32 lanes split the work, the partial results
// are reduced with the tie rule (equal t: larger index).  Called by all lanes of the warp.  Starts from the hit the
// ray already holds (the big spheres and whatever the walks found: all legitimate candidates).
//   * short lists (no GridCull): the lanes split the whole list;
//   * long lists: the lanes first split the coarse by-centre cells and keep those the ray's cone of acceptance can
//     reach -- a sphere is accepted iff rho^2 <= r^2 + eps m^2 (rho = distance of its centre from the ray line, m = its
//     projection; header), so a cell whose centre is farther from the line than sqrt(r_max^2 + eps (|m| + hd)^2) + hd
//     (hd = half its diagonal) holds none -- and then sweep only the spheres of those cells.
// Returns the number of rays of the warp it resolved (uniform).
__device__ __forceinline__ uint32_t grid_fallback_sweep(const GridCull& C, const float4* __restrict__ geom, uint32_t n,
                                                        const f3 o, const f3 d, bool unsafe, float& best_t, int& best_k) {
    const unsigned lane = threadIdx.x & 31u;
    unsigned pending = __ballot_sync(0xffffffffu, unsafe);
    const uint32_t resolved = (uint32_t)__popc(pending);
    while (pending) {
        const int src = __ffs((int)pending) - 1;
        pending &= pending - 1u;
        const f3 ro = mk3(__shfl_sync(0xffffffffu, o.x, src), __shfl_sync(0xffffffffu, o.y, src), __shfl_sync(0xffffffffu, o.z, src));
        const f3 rd = mk3(__shfl_sync(0xffffffffu, d.x, src), __shfl_sync(0xffffffffu, d.y, src), __shfl_sync(0xffffffffu, d.z, src));
        float bt = __shfl_sync(0xffffffffu, best_t, src);
        int bk = __shfl_sync(0xffffffffu, best_k, src);
        if (C.nx == 0) {
            // 8 independent 512-byte rows of the list in flight per trip: the loop is bound by the latency of the loads
            for (uint32_t k0 = lane; k0 < n; k0 += 32u * 8u) {
                float4 s[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t k = k0 + 32u * (uint32_t)u;
                    s[u] = __ldg(geom + (k < n ? k : k0));
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t k = k0 + 32u * (uint32_t)u;
                    if (k < n) grid_test_sphere(s[u], k, ro, rd, bt, bk);
                }
            }
        } else {
            const float dd = fmaf(rd.z, rd.z, fmaf(rd.y, rd.y, rd.x * rd.x));
            const float eps = fabsf(dd - 1.0f) + kGridEpsFloor;
            const float inv_dd = 1.0f / dd, len = sqrtf(dd);
            const float hd = 0.8661f * C.h;  // half the diagonal of a cell, rounded up
            const uint32_t nxy = (uint32_t)(C.nx * C.ny), ncell = nxy * (uint32_t)C.nz;
            for (uint32_t base = 0; base < ncell; base += 32u) {
                const uint32_t cell = base + lane;
                bool relevant = false;
                if (cell < ncell) {
                    const uint32_t cz = cell / nxy, rem = cell - cz * nxy, cy = rem / (uint32_t)C.nx, cx = rem - cy * (uint32_t)C.nx;
                    const float qx = C.ox + ((float)cx + 0.5f) * C.h - ro.x;
                    const float qy = C.oy + ((float)cy + 0.5f) * C.h - ro.y;
                    const float qz = C.oz + ((float)cz + 0.5f) * C.h - ro.z;
                    const float m = (qx * rd.x + qy * rd.y + qz * rd.z) * inv_dd;  // parameter of the closest approach
                    const float q2 = qx * qx + qy * qy + qz * qz;
                    const float rho2 = fmaxf(q2 - m * m * dd, 0.0f);
                    const float mlen = fabsf(m) * len + hd;
                    const float reach = (sqrtf(fmaf(eps * mlen, mlen, C.r_max * C.r_max)) + hd) * 1.001f + 1e-3f;
                    // the subtraction above cancels: allow for its rounding (a few ulp of q2)
                    relevant = rho2 <= reach * reach + 1e-6f * q2 && __ldg(C.cell_start + cell + 1u) > __ldg(C.cell_start + cell);
                }
                unsigned hits = __ballot_sync(0xffffffffu, relevant);
                while (hits) {
                    const uint32_t c = base + (uint32_t)__ffs((int)hits) - 1u;
                    hits &= hits - 1u;
                    const uint32_t first = __ldg(C.cell_start + c), last = __ldg(C.cell_start + c + 1u);
                    for (uint32_t i = first + lane; i < last; i += 32u) {
                        const uint32_t k = __ldg(C.items + i);
                        grid_test_sphere(__ldg(geom + k), k, ro, rd, bt, bk);
                    }
                }
            }
        }
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
