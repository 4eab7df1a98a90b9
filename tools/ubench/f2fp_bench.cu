// f2fp_bench.cu -- semantics and issue cost of cvt.rs.satfinite.e2m1x4.f32 (two chained F2FP.E2M1.PACK_AB_MERGE_C) as a
// way to collect the sign bits of four FP32 values with two instructions (the sweep's funnel shift takes one per value).
//   nvcc -arch=sm_100a -O3 -o f2fp_bench f2fp_bench.cu && ./f2fp_bench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t cvt_e2m1x4(float a, float b, float c, float d) {
    uint32_t r;
    asm("{ .reg .b16 t; cvt.rs.satfinite.e2m1x4.f32 t, {%1, %2, %3, %4}, %5; cvt.u32.u16 %0, t; }"
        : "=r"(r) : "f"(a), "f"(b), "f"(c), "f"(d), "r"(0u));
    return r;
}

__global__ void semantics(const float* in, uint32_t* out) {
    const float* v = in + 4 * threadIdx.x;
    out[threadIdx.x] = cvt_e2m1x4(v[0], v[1], v[2], v[3]);
}

// mix: per iteration 22 FFMA2-equivalent... here: K independent FFMA chains + per 4 chains one e2m1x4 (variant 1) or four
// funnel shifts (variant 0); the elapsed time tells whether the conversion costs issue slots like an ALU instruction
template <int kVariant>
__global__ void __launch_bounds__(256) mix(float* out, float b, float c, int iters) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = (float)(threadIdx.x + i) * 1e-3f - 0.1f;
    uint32_t m = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 11; ++rep)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);
        if (kVariant == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) m = __funnelshift_l(__float_as_uint(a[i]), m, 1);
        } else if (kVariant == 1) {
            m ^= cvt_e2m1x4(a[0], a[1], a[2], a[3]);
            m ^= cvt_e2m1x4(a[4], a[5], a[6], a[7]) << 16;
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)m;
}

int main() {
    float h_in[32 * 4];
    for (int t = 0; t < 32; ++t)
        for (int i = 0; i < 4; ++i) h_in[4 * t + i] = ((t >> i) & 1) ? -0.75f : 0.75f;   // lane t: value i negative iff bit i of t
    h_in[4 * 16 + 0] = -0.0f; h_in[4 * 16 + 1] = 1e-30f; h_in[4 * 16 + 2] = -1e-30f; h_in[4 * 16 + 3] = 0.0f;  // lane 16: zeros / tiny
    h_in[4 * 17 + 0] = -1e30f; h_in[4 * 17 + 1] = 1e30f; h_in[4 * 17 + 2] = -3.0f; h_in[4 * 17 + 3] = 3.0f;
    float* d_in; uint32_t* d_out;
    cudaMalloc(&d_in, sizeof h_in); cudaMalloc(&d_out, 32 * 4);
    cudaMemcpy(d_in, h_in, sizeof h_in, cudaMemcpyHostToDevice);
    semantics<<<1, 32>>>(d_in, d_out);
    uint32_t h_out[32];
    cudaMemcpy(h_out, d_out, sizeof h_out, cudaMemcpyDeviceToHost);
    for (int t = 0; t < 18; ++t) printf("lane %2d (negative mask %x): 0x%04x\n", t, t & 15, h_out[t]);
    float* d_f; cudaMalloc(&d_f, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int variant = 0; variant < 3; ++variant) {
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            if (variant == 0) mix<0><<<148 * 8, 256>>>(d_f, 1.0001f, 1e-4f, 4000);
            else if (variant == 1) mix<1><<<148 * 8, 256>>>(d_f, 1.0001f, 1e-4f, 4000);
            else mix<2><<<148 * 8, 256>>>(d_f, 1.0001f, 1e-4f, 4000);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        const double ffma = 148.0 * 8 * 256 * 4000.0 * 88;
        printf("variant %d (%s): %.3f ms  %.2f T FFMA lane-ops/s\n", variant,
               variant == 0 ? "8 SHF per 88 FFMA" : variant == 1 ? "2 e2m1x4 (4 F2FP) per 88 FFMA" : "FFMA only", best, ffma / best / 1e9);
    }
    return 0;
}
