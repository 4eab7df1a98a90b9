// rtw_scenegen.cu -- scene_random_spheres (src/scenes.jl:49-84) generated on the device.
//
// The reference's builder is a sequential loop over the cells (a, b) of a square field; every cell draws from the calling
// thread's Xoroshiro128Plus (src/rand.jl:7-13): choose_mat, then the centre (x before z), then 6 more draws for a
// diffuse sphere, 4 for a metal one, none for glass -- and a cell too close to (4, 0.2, 0) is skipped after its three
// draws (src/scenes.jl:57-61).  How many draws a cell takes therefore depends on its own first draw, and where its draws
// sit in the stream on every cell before it.  This file reproduces the host loop's list BIT FOR BIT (the Python mirror
// raytracingweekend.jl_b200/host.py::scene_random_spheres; 1.8 s for the ~100k-sphere list of BASELINE configs[4]) in a
// handful of parallel kernels:
//   1. the raw stream: xoroshiro128+ is linear over GF(2), so the state after k steps is T^k s.  The host squares the
//      128 x 128 bit matrix T twenty times once per process; thread t jumps to step 64 t by multiplying its state with
//      the T^(2^j) of the set bits of 64 t, then emits 64 draws as Float32 ((u & 0xffffffff) >> 9) * 2^-23.
//   2. cell positions: next[p] = p + 9 | 7 | 3 by the draw at p (as if a cell started at EVERY stream position), then
//      log2(cells) rounds of pointer doubling J_j[p] = J_(j-1)[J_(j-1)[p]]: the position k cells after p is a product of
//      J_j over the bits of k.  Only the four cells with a in {3, 4}, b in {-1, 0} can be skipped (|cx - 4| < 0.9 and
//      |cz| < 0.9 need exactly those); one thread walks these four and records an anchor (cell index, stream position,
//      spheres dropped so far) after each.
//   3. one thread per cell: position from its anchor, the cell's spheres in the reference's Float32 arithmetic (no
//      contraction: this file is compiled with -fmad=false like the rest), written at its final list index.
// The generator state after the last draw (T^draws s) goes back to the caller, so host code that keeps drawing from the
// same TRNG continues exactly where the reference's loop would have left it.
#include "rtw_kernels.h"

#include <vector>

namespace rtw {

namespace {

struct U128 {
    unsigned long long s0, s1;
};

__host__ __device__ inline unsigned long long rotl64(unsigned long long x, int k) { return (x << k) | (x >> (64 - k)); }

// one step of xoroshiro128+ (constants 55, 14, 36 -- RandomNumbers.jl 1.5.3 as restated in host.py): the state transition
__host__ __device__ inline U128 xoro_step(U128 s) {
    const unsigned long long a = s.s0, b = s.s1 ^ s.s0;
    return U128{rotl64(a, 55) ^ b ^ (b << 14), rotl64(b, 36)};
}

constexpr int kJumpLevels = 24;  // T^(2^j), j = 0 .. 23: jumps of up to 16 M draws

// M v over GF(2): column b of M is the image of basis vector b
__host__ __device__ inline U128 jump_apply(const U128* cols, U128 v) {
    U128 acc{0ull, 0ull};
    for (int b = 0; b < 64; ++b)
        if ((v.s0 >> b) & 1ull) { acc.s0 ^= cols[b].s0; acc.s1 ^= cols[b].s1; }
    for (int b = 0; b < 64; ++b)
        if ((v.s1 >> b) & 1ull) { acc.s0 ^= cols[64 + b].s0; acc.s1 ^= cols[64 + b].s1; }
    return acc;
}

const std::vector<U128>& jump_tables() {
    static const std::vector<U128> tab = [] {
        std::vector<U128> t((size_t)kJumpLevels * 128);
        for (int b = 0; b < 128; ++b) {
            U128 e{b < 64 ? 1ull << b : 0ull, b >= 64 ? 1ull << (b - 64) : 0ull};
            t[b] = xoro_step(e);
        }
        for (int j = 1; j < kJumpLevels; ++j)
            for (int b = 0; b < 128; ++b) t[(size_t)j * 128 + b] = jump_apply(&t[(size_t)(j - 1) * 128], t[(size_t)(j - 1) * 128 + b]);
        return t;
    }();
    return tab;
}

__device__ inline U128 jump_by(const U128* __restrict__ tables, U128 s, unsigned long long k) {
    for (int j = 0; j < kJumpLevels; ++j)
        if ((k >> j) & 1ull) s = jump_apply(tables + (size_t)j * 128, s);
    return s;
}

constexpr int kGenBlock = 64;  // draws per thread of the stream kernel

__global__ void __launch_bounds__(128) gen_stream_kernel(U128 state0, const U128* __restrict__ tables, float* __restrict__ rnd,
                                                         uint32_t n_draws) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long first = (unsigned long long)t * kGenBlock;
    if (first >= n_draws) return;
    U128 s = jump_by(tables, state0, first);
    for (int i = 0; i < kGenBlock && first + i < n_draws; ++i) {
        const unsigned long long u = s.s0 + s.s1;  // the output precedes the transition
        rnd[first + i] = (float)((uint32_t)u >> 9) * 1.1920928955078125e-07f;  // trand(Float32): 23 bits of the low word
        s = xoro_step(s);
    }
}

// draws a cell takes when it is not skipped, by its first draw (src/scenes.jl:63-74)
__device__ inline uint32_t cell_draws(float choose_mat) { return choose_mat < 0.8f ? 9u : (choose_mat < 0.95f ? 7u : 3u); }

__global__ void __launch_bounds__(256) next_kernel(const float* __restrict__ rnd, uint32_t n_draws, uint32_t* __restrict__ J0) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > n_draws) return;
    uint32_t q = n_draws;  // sentinel: past the end stays past the end
    if (p < n_draws) q = min(p + cell_draws(rnd[p]), n_draws);
    J0[p] = q;
}

__global__ void __launch_bounds__(256) double_kernel(const uint32_t* __restrict__ Jprev, uint32_t* __restrict__ Jnext, uint32_t n_draws) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > n_draws) return;
    Jnext[p] = Jprev[Jprev[p]];
}

__device__ inline uint32_t jump_cells(const uint32_t* __restrict__ J, uint32_t stride, int levels, uint32_t pos, uint32_t k) {
    for (int j = 0; j < levels; ++j)
        if ((k >> j) & 1u) pos = J[(size_t)j * stride + pos];
    return pos;
}

// the cell body of src/scenes.jl:57-61 in the reference's Float32 arithmetic: centre and the "too close" test
__device__ inline bool cell_centre(const float* __restrict__ rnd, uint32_t p, int a, int b, float& cx, float& cz) {
    cx = __fadd_rn((float)a, __fmul_rn(0.9f, rnd[p + 1u]));  // :58, x drawn before z
    cz = __fadd_rn((float)b, __fmul_rn(0.9f, rnd[p + 2u]));
    const float dx = __fadd_rn(cx, -4.0f), dy = __fadd_rn(0.2f, -0.2f), dz = cz;
    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    return nrm < 0.9f;  // :61 skip
}

struct GenAnchors {
    U128 state_after;                      // the first 32 bytes go back to the host
    uint32_t n_spheres, total_draws, n_anchors, pad;
    uint32_t idx[5], pos[5], dropped[5];  // from cell idx[k] on: stream position pos[k], dropped[k] cells skipped before
};

__global__ void anchors_kernel(const float* __restrict__ rnd, const uint32_t* __restrict__ J, uint32_t stride, int levels,
                               int half, U128 state0, const U128* __restrict__ tables, GenAnchors* __restrict__ out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const uint32_t side = 2u * (uint32_t)half, cells = side * side;
    GenAnchors A;
    A.idx[0] = 0; A.pos[0] = 0; A.dropped[0] = 0;
    uint32_t na = 1, dropped = 0;
    for (int a = 3; a <= 4; ++a)
        for (int b = -1; b <= 0; ++b) {
            if (a < -half || a > half - 1 || b < -half || b > half - 1) continue;
            const uint32_t ci = (uint32_t)(a + half) * side + (uint32_t)(b + half);
            const uint32_t p = jump_cells(J, stride, levels, A.pos[na - 1], ci - A.idx[na - 1]);
            float cx, cz;
            const bool skip = cell_centre(rnd, p, a, b, cx, cz);
            dropped += skip ? 1u : 0u;
            A.idx[na] = ci + 1u;
            A.pos[na] = p + (skip ? 3u : cell_draws(rnd[p]));
            A.dropped[na] = dropped;
            ++na;
        }
    A.n_anchors = na;
    A.n_spheres = 1u + cells - dropped + 3u;
    A.total_draws = jump_cells(J, stride, levels, A.pos[na - 1], cells - A.idx[na - 1]);
    A.pad = 0;
    A.state_after = jump_by(tables, state0, A.total_draws);
    *out = A;
}

__global__ void __launch_bounds__(256) cells_kernel(const float* __restrict__ rnd, const uint32_t* __restrict__ J, uint32_t stride,
                                                    int levels, int half, const GenAnchors* __restrict__ anchors,
                                                    float4* __restrict__ geom, float4* __restrict__ mat, uint32_t* __restrict__ kind) {
    const uint32_t side = 2u * (uint32_t)half, cells = side * side;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const GenAnchors& A = *anchors;
    if (i == 0) {
        geom[0] = make_float4(0.f, -1000.f, -1.f, 1000.f);  // ground, src/scenes.jl:53-54
        mat[0] = make_float4(0.5f, 0.5f, 0.5f, 0.f);
        kind[0] = 0u;
        const uint32_t n = A.n_spheres;
        geom[n - 3] = make_float4(0.f, 1.f, 0.f, 1.f);   mat[n - 3] = make_float4(1.f, 1.f, 1.f, 1.5f);  kind[n - 3] = 2u;  // :78
        geom[n - 2] = make_float4(-4.f, 1.f, 0.f, 1.f);  mat[n - 2] = make_float4(0.4f, 0.2f, 0.1f, 0.f); kind[n - 2] = 0u;  // :79-80
        geom[n - 1] = make_float4(4.f, 1.f, 0.f, 1.f);   mat[n - 1] = make_float4(0.7f, 0.6f, 0.5f, 0.f); kind[n - 1] = 1u;  // :81-82
    }
    if (i >= cells) return;
    uint32_t k = 0;
    for (uint32_t q = 1; q < A.n_anchors; ++q)
        if (A.idx[q] <= i) k = q;
    const uint32_t p = jump_cells(J, stride, levels, A.pos[k], i - A.idx[k]);
    const int a = (int)(i / side) - half, b = (int)(i % side) - half;  // `for a in lo:hi, b in lo:hi`: b runs fastest
    float cx, cz;
    if (cell_centre(rnd, p, a, b, cx, cz)) return;  // skipped
    const uint32_t at = 1u + i - A.dropped[k];
    const float choose_mat = rnd[p];
    geom[at] = make_float4(cx, 0.2f, cz, 0.2f);
    if (choose_mat < 0.8f) {  // diffuse: albedo = rand3 .* rand3, :63-66
        mat[at] = make_float4(__fmul_rn(rnd[p + 3u], rnd[p + 6u]), __fmul_rn(rnd[p + 4u], rnd[p + 7u]),
                              __fmul_rn(rnd[p + 5u], rnd[p + 8u]), 0.f);
        kind[at] = 0u;
    } else if (choose_mat < 0.95f) {  // metal: albedo in [0.5, 1), fuzz in [0, 5), :67-71 (random_between = r*(max-min)+min)
        mat[at] = make_float4(__fadd_rn(__fmul_rn(rnd[p + 3u], 0.5f), 0.5f), __fadd_rn(__fmul_rn(rnd[p + 4u], 0.5f), 0.5f),
                              __fadd_rn(__fmul_rn(rnd[p + 5u], 0.5f), 0.5f), __fadd_rn(__fmul_rn(rnd[p + 6u], 5.0f), 0.0f));
        kind[at] = 1u;
    } else {  // glass, :72-74
        mat[at] = make_float4(1.f, 1.f, 1.f, 1.5f);
        kind[at] = 2u;
    }
}

}  // namespace

size_t scenegen_max_spheres(int half) { return (size_t)4 * half * half + 4; }

// Work space (device): rnd[L + 16] floats, J[levels][L + 1] u32, tables, anchors.  Returns the byte count.
size_t scenegen_workspace_bytes(int half) {
    const size_t cells = (size_t)4 * half * half, L = 9 * cells;
    int levels = 1;
    while ((1ull << levels) <= cells) ++levels;
    return ((L + 16) * 4 + 255) / 256 * 256 + ((size_t)levels * (L + 1) * 4 + 255) / 256 * 256 +
           (size_t)kJumpLevels * 128 * sizeof(U128) + 256;
}

// Enqueues the generation on `stream`.  state = the generator state (s0, s1) the host builder would start from.
// geom / mat / kind: device arrays of scenegen_max_spheres(half) entries.  After the stream has been synchronised,
// out_host (4 x u64) holds {state_after.s0, state_after.s1, n_spheres | total_draws << 32, n_anchors}.
cudaError_t launch_scenegen(unsigned long long s0, unsigned long long s1, int half, void* workspace, float4* geom, float4* mat,
                            uint32_t* kind, unsigned long long* out_host, cudaStream_t stream) {
    if (half < 1 || half > 512) return cudaErrorInvalidValue;
    const uint32_t cells = 4u * (uint32_t)half * (uint32_t)half, L = 9u * cells;
    int levels = 1;
    while ((1ull << levels) <= cells) ++levels;
    unsigned char* w = (unsigned char*)workspace;
    float* rnd = (float*)w;
    w += ((size_t)(L + 16) * 4 + 255) / 256 * 256;
    uint32_t* J = (uint32_t*)w;
    w += ((size_t)levels * (L + 1) * 4 + 255) / 256 * 256;
    U128* tables = (U128*)w;
    w += (size_t)kJumpLevels * 128 * sizeof(U128);
    GenAnchors* anchors = (GenAnchors*)w;
    const std::vector<U128>& tab = jump_tables();
    cudaError_t e = cudaMemcpyAsync(tables, tab.data(), tab.size() * sizeof(U128), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    const U128 state0{s0, s1};
    const uint32_t gen_threads = (L + kGenBlock - 1) / kGenBlock;
    gen_stream_kernel<<<(gen_threads + 127) / 128, 128, 0, stream>>>(state0, tables, rnd, L);
    e = cudaMemsetAsync(rnd + L, 0, 16 * 4, stream);  // reads past a cell that starts near the end stay defined
    if (e != cudaSuccess) return e;
    const uint32_t stride = L + 1;
    next_kernel<<<(stride + 255) / 256, 256, 0, stream>>>(rnd, L, J);
    for (int j = 1; j < levels; ++j)
        double_kernel<<<(stride + 255) / 256, 256, 0, stream>>>(J + (size_t)(j - 1) * stride, J + (size_t)j * stride, L);
    anchors_kernel<<<1, 32, 0, stream>>>(rnd, J, stride, levels, half, state0, tables, anchors);
    cells_kernel<<<(cells + 255) / 256, 256, 0, stream>>>(rnd, J, stride, levels, half, anchors, geom, mat, kind);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // the state after the last draw, n_spheres and total_draws: the head of GenAnchors
    return cudaMemcpyAsync(out_host, anchors, 32, cudaMemcpyDeviceToHost, stream);
}

}  // namespace rtw
