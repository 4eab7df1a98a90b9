// Microbenchmark: scalar FFMA vs packed FFMA2 issue/pipe throughput on sm_100a, and how many extra
// ALU / LDS instructions ride along for free next to packed FP32 work.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 2048;
constexpr int kChains = 8;

__global__ void __launch_bounds__(256) k_ffma(float* out, float b, float c) {
    float a[kChains * 2];
#pragma unroll
    for (int i = 0; i < kChains * 2; ++i) a[i] = (float)(threadIdx.x + i) * 1e-3f;
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains * 2; ++i) a[i] = fmaf(a[i], b, c);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kChains * 2; ++i) s += a[i];
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_ffma2(float* out, float b, float c) {
    float2 a[kChains];
    float2 bb = make_float2(b, b * 1.0001f), cc = make_float2(c, c * 0.999f);
#pragma unroll
    for (int i = 0; i < kChains; ++i) a[i] = make_float2((float)(threadIdx.x + i) * 1e-3f, (float)i);
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) a[i] = __ffma2_rn(a[i], bb, cc);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i].x + a[i].y;
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// packed FFMA2 plus NALU funnel shifts and NLDS shared loads per 11 FFMA2 (the sweep's ratio is 11 : 2 : 2)
template <int NALU, int NLDS>
__global__ void __launch_bounds__(256) k_mix(float* out, float b, float c) {
    __shared__ float4 sm[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) sm[i] = make_float4(b, c, b, c);
    __syncthreads();
    float2 a[11];
    float2 cc = make_float2(c, c * 0.999f);
#pragma unroll
    for (int i = 0; i < 11; ++i) a[i] = make_float2((float)(threadIdx.x + i) * 1e-3f, (float)i);
    unsigned m = threadIdx.x;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int it = 0; it < kIters; ++it) {
        float4 v[NLDS > 0 ? NLDS : 1];
#pragma unroll
        for (int j = 0; j < NLDS; ++j) v[j] = sm[(it * NLDS + j) & 511];
        float2 bb = NLDS > 0 ? make_float2(v[0].x, v[0].y) : make_float2(b, b);
#pragma unroll
        for (int i = 0; i < 11; ++i) a[i] = __ffma2_rn(a[i], bb, cc);
#pragma unroll
        for (int j = 0; j < NALU; ++j) m = __funnelshift_l(__float_as_uint(a[j].x), m, 1);
#pragma unroll
        for (int j = 1; j < NLDS; ++j) { acc.x += 0.f; m ^= __float_as_uint(v[j].z); }
    }
    float s = acc.x;
#pragma unroll
    for (int i = 0; i < 11; ++i) s += a[i].x + a[i].y;
    if (s == 123.456f || m == 0x12345u) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
double time_kernel(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    return best * 1e-3;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out;
    cudaMalloc(&out, 64 << 20);
    const int grid = sms * 8, block = 256;
    const double lanes = (double)grid * block;
    double t;
    t = time_kernel([&] { k_ffma<<<grid, block>>>(out, 0.999f, 1e-4f); });
    printf("scalar FFMA : %.2f T lane-FMA/s (%.3f ms)\n", lanes * kIters * kChains * 2 / t / 1e12, t * 1e3);
    t = time_kernel([&] { k_ffma2<<<grid, block>>>(out, 0.999f, 1e-4f); });
    printf("packed FFMA2: %.2f T lane-FMA/s (%.3f ms)\n", lanes * kIters * kChains * 2 / t / 1e12, t * 1e3);
    t = time_kernel([&] { k_mix<0, 0><<<grid, block>>>(out, 0.999f, 1e-4f); });
    printf("FFMA2 x11 + 0 SHF + 0 LDS: %.2f T lane-FMA/s\n", lanes * kIters * 22 / t / 1e12);
    t = time_kernel([&] { k_mix<2, 0><<<grid, block>>>(out, 0.999f, 1e-4f); });
    printf("FFMA2 x11 + 2 SHF + 0 LDS: %.2f T lane-FMA/s\n", lanes * kIters * 22 / t / 1e12);
    t = time_kernel([&] { k_mix<2, 2><<<grid, block>>>(out, 0.999f, 1e-4f); });
    printf("FFMA2 x11 + 2 SHF + 2 LDS.128: %.2f T lane-FMA/s\n", lanes * kIters * 22 / t / 1e12);
    t = time_kernel([&] { k_mix<4, 2><<<grid, block>>>(out, 0.999f, 1e-4f); });
    printf("FFMA2 x11 + 4 SHF + 2 LDS.128: %.2f T lane-FMA/s\n", lanes * kIters * 22 / t / 1e12);
    t = time_kernel([&] { k_mix<8, 2><<<grid, block>>>(out, 0.999f, 1e-4f); });
    printf("FFMA2 x11 + 8 SHF + 2 LDS.128: %.2f T lane-FMA/s\n", lanes * kIters * 22 / t / 1e12);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
