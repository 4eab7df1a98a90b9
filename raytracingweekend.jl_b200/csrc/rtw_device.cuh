// rtw_device.cuh -- device-side building blocks of the B200 hot path (sm_100a, FP32 CUDA cores).
//
// Restates, for one lane = one path, the reference functions (paths relative to the reference repo):
//   get_ray            src/camera.jl:43-48          hit(::Sphere)   src/hit.jl:12-35
//   ray_to_HitRecord   src/hit.jl:6-10              scatter x3      src/material.jl:13-53
//   reflect/refract/reflectance  src/light.jl:6-25  skycolor        src/ray_color.jl:1-6
//   random_vec3_in_sphere / random_vec2_in_disk     src/rand.jl:15-38
//
// Floating-point contract (DESIGN.md "FP contract"; the CPU oracle follows the same one independently):
//   dot(a,b) = fma(a.z,b.z, fma(a.y,b.y, a.x*b.x));  a +- b*c is ONE fma;  sqrt and / are IEEE (rn);
//   normalize(v) = v * (1/sqrt(dot(v,v)));  colour math is Float64 (the reference promotes, ray_color.jl:2-3).
// This file is compiled with -fmad=false: nothing is fused unless written as fmaf()/fma().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rtw {

struct f3 {
    float x, y, z;
};
__device__ __forceinline__ f3 mk3(float x, float y, float z) { return f3{x, y, z}; }
__device__ __forceinline__ float dot3(f3 a, f3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ f3 normalize3(f3 a) {
    float inv = __frcp_rn(__fsqrt_rn(dot3(a, a)));
    return mk3(a.x * inv, a.y * inv, a.z * inv);
}

// ---------------------------------------------------------------------------------------------- RNG
// Production stream: Philox4x32-7, key = (seed lo, seed hi), counter = (block, sample, pixel, event).
// Every uniform is ADDRESSED by (pixel, sample, event, draw) -- no generator state is carried in registers
// and every Philox evaluation sits at a point where the lanes of a warp are converged:
//   event 0 = primary ray: draws 0,1 = jitter du,dv; disk attempt k = draws 2+2k, 3+2k
//   event e >= 1 = scatter at the e-th hit: ball attempt a = draws 4a..4a+2 (x,y,z); draw 3 = dielectric coin
//   draw n = word (n mod 4) of block (n div 4);  f32 = (word >> 9) * 2^-23
// This replaces the reference's per-thread sequential Xoroshiro128Plus (src/init.jl:2-12, src/rand.jl:2-13).
constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;
// Rounds of the production stream: Philox4x32-7, the smallest round count Salmon et al. (SC'11, table 2) report as
// Crush-resistant (BigCrush passes); 10 is Random123's default with a safety margin.  The reference's own generator,
// xoroshiro128+, fails BigCrush's linearity tests on its low bits, so 7 rounds are no step down from it.  Three rounds less
// are 1.3 % of the headline render (the kernel is issue-bound and a round is two 64-bit multiplies).  KAT: tests/.
constexpr int kPhiloxRounds = 7;

struct PathRng {
    uint32_t sample;  // counter word 1
    uint32_t pixel;   // counter word 2 (global pixel index i0*W + j0)
};

struct u32x4 {
    uint32_t w0, w1, w2, w3;
};

__device__ __forceinline__ u32x4 philox_block(const PathRng& g, uint32_t event, uint32_t block, uint32_t k0,
                                              uint32_t k1) {
    uint32_t c0 = block, c1 = g.sample, c2 = g.pixel, c3 = event;
#pragma unroll
    for (int r = 0; r < kPhiloxRounds; ++r) {
        uint32_t hi0 = __umulhi(kPhiloxM0, c0), lo0 = kPhiloxM0 * c0;
        uint32_t hi1 = __umulhi(kPhiloxM1, c2), lo1 = kPhiloxM1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += kPhiloxW0; k1 += kPhiloxW1;
    }
    return u32x4{c0, c1, c2, c3};
}

// trand(Float32), src/rand.jl:10-13
__device__ __forceinline__ float u01(uint32_t w) { return (float)(w >> 9) * 1.1920928955078125e-07f; }  // 2^-23
// random_between(-1, 1) = trand*(max-min)+min, src/rand.jl:24
__device__ __forceinline__ float pm1(uint32_t w) { return fmaf(u01(w), 2.0f, -1.0f); }

// normalize(random_vec3_in_sphere), src/rand.jl:15-22,29: rejection in the cube; attempt 0 comes from `b0`
// (block 0 of the event, already evaluated by the caller with the warp converged), attempt a from block a.
__device__ __forceinline__ f3 rng_unit_vector(const PathRng& g, uint32_t event, u32x4 b0, uint32_t k0, uint32_t k1) {
    f3 p = mk3(pm1(b0.w0), pm1(b0.w1), pm1(b0.w2));
    uint32_t a = 1;
    while (!(dot3(p, p) <= 1.0f)) {
        u32x4 b = philox_block(g, event, a, k0, k1);
        p = mk3(pm1(b.w0), pm1(b.w1), pm1(b.w2));
        ++a;
    }
    return normalize3(p);
}

// ---------------------------------------------------------------------------------------------- camera
struct DevCamera {  // the fields of Camera{Float32} get_ray reads (src/camera.jl:1-10; w is unused by get_ray)
    f3 origin, llc, horizontal, vertical, u, v;
    float lens_radius;
};

// Primary ray of sample s0 of a pixel (src/render.jl:26-37 + get_ray, src/camera.jl:43-48).
// u_base, v_base are the un-jittered T(j/W), T((H-i)/H); the disk sample is always drawn (camera.jl:44).
__device__ __forceinline__ void primary_ray(const DevCamera& c, const PathRng& g, uint32_t k0, uint32_t k1,
                                            uint32_t s0, float u_base, float v_base, float fw, float fh, f3& o,
                                            f3& d) {
    u32x4 b = philox_block(g, 0u, 0u, k0, k1);
    float s = u_base, t = v_base;
    if (s0 != 0u) {  // first sample is centred (src/render.jl:30-36); du = draw 0, dv = draw 1
        s = u_base + __fdiv_rn(u01(b.w0), fw);
        t = v_base + __fdiv_rn(u01(b.w1), fh);
    }
    // random_vec2_in_disk, src/rand.jl:31-38: attempt k = draws 2+2k, 3+2k
    float px = pm1(b.w2), py = pm1(b.w3);
    uint32_t blk = 1;
    while (!(fmaf(py, py, px * px) <= 1.0f)) {
        b = philox_block(g, 0u, blk, k0, k1);
        px = pm1(b.w0);
        py = pm1(b.w1);
        if (!(fmaf(py, py, px * px) <= 1.0f)) {
            px = pm1(b.w2);
            py = pm1(b.w3);
        }
        ++blk;
    }
    float rx = c.lens_radius * px, ry = c.lens_radius * py;
    f3 off = mk3(fmaf(c.v.x, ry, c.u.x * rx), fmaf(c.v.y, ry, c.u.y * rx), fmaf(c.v.z, ry, c.u.z * rx));
    o = mk3(c.origin.x + off.x, c.origin.y + off.y, c.origin.z + off.z);
    f3 q;
    q.x = fmaf(t, c.vertical.x, fmaf(s, c.horizontal.x, c.llc.x)) - c.origin.x - off.x;
    q.y = fmaf(t, c.vertical.y, fmaf(s, c.horizontal.y, c.llc.y)) - c.origin.y - off.y;
    q.z = fmaf(t, c.vertical.z, fmaf(s, c.horizontal.z, c.llc.z)) - c.origin.z - off.z;
    d = normalize3(q);
}

// ---------------------------------------------------------------------------------------------- intersection
// The discriminant of hit(::Sphere), src/hit.jl:13-18: 3 FADD + 2 FMUL + 6 FFMA = 11 FP32 instructions.
__device__ __forceinline__ float sphere_disc(float4 s, f3 o, f3 d, float& half_b) {
    f3 oc = mk3(o.x - s.x, o.y - s.y, o.z - s.z);
    half_b = dot3(oc, d);
    float cq = fmaf(-s.w, s.w, dot3(oc, oc));
    return fmaf(half_b, half_b, -cq);
}

// Root selection of hit(::Sphere), src/hit.jl:20-29, against the running closest t (src/hit.jl:43-46).
// Returns true and updates best_t when the sphere is the new closest hit (ties: later sphere wins).
__device__ __forceinline__ bool sphere_accept(float disc, float half_b, float tmin, float& best_t) {
    float sq = __fsqrt_rn(disc);
    float root = -half_b - sq;
    if (root < tmin || best_t < root) {
        root = -half_b + sq;
        if (root < tmin || best_t < root) return false;
    }
    best_t = root;
    return true;
}

// ---------------------------------------------------------------------------------------------- materials
// reflect(v,n) = v - (2v.n)n, src/light.jl:6
__device__ __forceinline__ f3 reflect3(f3 v, f3 n) {
    float k = 2.0f * dot3(v, n);
    return mk3(fmaf(-k, n.x, v.x), fmaf(-k, n.y, v.y), fmaf(-k, n.z, v.z));
}

// Schlick, src/light.jl:19-25
__device__ __forceinline__ float reflectance(float cos_t, float ratio) {
    float r0 = __fdiv_rn(1.0f - ratio, 1.0f + ratio);
    r0 = r0 * r0;
    float x = 1.0f - cos_t;
    float x2 = x * x;
    float x4 = x2 * x2;
    float x5 = x4 * x;
    return fmaf(1.0f - r0, x5, r0);
}

// refract, src/light.jl:12-17
__device__ __forceinline__ f3 refract3(f3 d, f3 n, float ratio) {
    float cos_t = fminf(-dot3(d, n), 1.0f);
    f3 perp = mk3(ratio * fmaf(cos_t, n.x, d.x), ratio * fmaf(cos_t, n.y, d.y), ratio * fmaf(cos_t, n.z, d.z));
    float s = __fsqrt_rn(fabsf(1.0f - dot3(perp, perp)));
    return normalize3(mk3(fmaf(-s, n.x, perp.x), fmaf(-s, n.y, perp.y), fmaf(-s, n.z, perp.z)));
}

// HitRecord reconstruction (src/hit.jl:31-34, 6-10) + scatter (src/material.jl) for the closest sphere.
// In: ray (o,d), root t, sphere geometry g, material m (albedo + fuzz|ir), kind.
// Out: o,d replaced by the scattered ray; att = attenuation.
// `event` = index of this hit along the path (1, 2, ...): addresses the scatter's random draws.
__device__ __forceinline__ void shade_hit(f3& o, f3& d, float t, float4 g, float4 m, uint32_t kind,
                                          const PathRng& rng, uint32_t event, uint32_t k0, uint32_t k1, f3& att) {
    const u32x4 b0 = philox_block(rng, event, 0u, k0, k1);  // evaluated by every shading lane together
    f3 p = mk3(fmaf(t, d.x, o.x), fmaf(t, d.y, o.y), fmaf(t, d.z, o.z));                       // point(), hit.jl:3
    f3 on = mk3(__fdiv_rn(p.x - g.x, g.w), __fdiv_rn(p.y - g.y, g.w), __fdiv_rn(p.z - g.z, g.w));  // hit.jl:33
    bool front = dot3(d, on) < 0.0f;                                                         // hit.jl:7
    f3 n = front ? on : mk3(-on.x, -on.y, -on.z);                                            // hit.jl:8
    f3 nd;
    if (kind != 2u) {
        // Lambertian (material.jl:13-23) and Metal (material.jl:31-34) both draw one unit vector first (Metal even
        // when fuzz == 0): one shared rejection loop keeps the two materials converged
        f3 rv = rng_unit_vector(rng, event, b0, k0, k1);
        if (kind == 0u) {
            f3 sd = mk3(n.x + rv.x, n.y + rv.y, n.z + rv.z);
            // near_zero: squared length (Float32) promoted and compared with the Float64 literal 1e-5, vec.jl:20
            nd = ((double)dot3(sd, sd) < 1e-5) ? n : normalize3(sd);
        } else {  // never absorbs (Scatter.reflected is always true, structs.jl:43)
            f3 refl = reflect3(d, n);
            nd = normalize3(mk3(fmaf(m.w, rv.x, refl.x), fmaf(m.w, rv.y, refl.y), fmaf(m.w, rv.z, refl.z)));
        }
        att = mk3(m.x, m.y, m.z);
    } else {  // Dielectric, material.jl:41-53
        float ratio = front ? __fdiv_rn(1.0f, m.w) : m.w;
        float cos_t = fminf(-dot3(d, n), 1.0f);
        float sin_t = __fsqrt_rn(fmaf(-cos_t, cos_t, 1.0f));
        bool cannot_refract = ratio * sin_t > 1.0f;
        att = mk3(1.0f, 1.0f, 1.0f);
        // `||` short-circuits (material.jl:47): the coin (draw 3 of the event) is ignored on total internal reflection
        if (cannot_refract || reflectance(cos_t, ratio) > u01(b0.w3))
            nd = reflect3(d, n);  // not re-normalised, material.jl:48
        else
            nd = refract3(d, n, ratio);
    }
    o = p;
    d = nd;
}

// skycolor(ray), src/ray_color.jl:1-6: t in Float32, white/skyblue are Float64 literals, no contraction
__device__ __forceinline__ void skycolor(f3 d, double& r, double& g, double& b) {
    float t = 0.5f * (d.y + 1.0f);
    double a = (double)(1.0f - t), w = (double)t;
    r = __dadd_rn(__dmul_rn(a, 1.0), __dmul_rn(w, 0.5));
    g = __dadd_rn(__dmul_rn(a, 1.0), __dmul_rn(w, 0.7));
    b = __dadd_rn(__dmul_rn(a, 1.0), __dmul_rn(w, 1.0));
}

}  // namespace rtw
