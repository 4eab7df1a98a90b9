// Which packed FP32x2 instructions run at full rate on sm_100a?  Independent chains, 256 threads x 8 CTAs/SM.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int kIters = 4096;
constexpr int kChains = 12;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float b, float c) {
    float2 a[kChains];
    float2 bb = make_float2(b, b * 1.0001f), cc = make_float2(c, c * 0.999f);
    float bs = b * 0.9999f;
#pragma unroll
    for (int i = 0; i < kChains; ++i) a[i] = make_float2((float)(threadIdx.x + i) * 1e-3f, (float)i);
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            if (MODE == 0) a[i] = __ffma2_rn(a[i], bb, cc);                       // FFMA2 all packed operands
            if (MODE == 1) a[i] = __fadd2_rn(a[i], cc);                           // FADD2
            if (MODE == 2) a[i] = __fmul2_rn(a[i], bb);                           // FMUL2
            if (MODE == 3) a[i] = __ffma2_rn(a[i], make_float2(bs, bs), cc);      // FFMA2 with a broadcast (.F32) operand
            if (MODE == 4) a[i] = __fadd2_rn(make_float2(bs, bs), make_float2(-a[i].x, -a[i].y));  // FADD2 bcast + neg
            if (MODE == 5) {                                                        // the sweep's 3:2:6 mix
                int m = i % 11;
                if (m < 3) a[i] = __fadd2_rn(make_float2(bs, bs), make_float2(-a[i].x, -a[i].y));
                else if (m < 5) a[i] = __fmul2_rn(a[i], make_float2(bs, bs));
                else a[i] = __ffma2_rn(a[i], bb, cc);
            }
            if (MODE == 6) a[i] = __ffma2_rn(a[i], a[i], make_float2(-cc.x, -cc.y));  // FFMA2 same reg twice + neg
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i].x + a[i].y;
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, float* out, int grid) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<grid, 256>>>(out, 0.999f, 1e-4f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    printf("%-44s %.2f T lane-ops/s\n", name, (double)grid * 256 * kIters * kChains * 2 / (best * 1e-3) / 1e12);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out;
    cudaMalloc(&out, 64 << 20);
    int grid = sms * 8;
    run<0>("FFMA2 packed operands", out, grid);
    run<1>("FADD2", out, grid);
    run<2>("FMUL2", out, grid);
    run<3>("FFMA2 with broadcast .F32 operand", out, grid);
    run<4>("FADD2 broadcast + negated packed", out, grid);
    run<5>("mix 3 FADD2 : 2 FMUL2 : 6 FFMA2", out, grid);
    run<6>("FFMA2 a*a - c", out, grid);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
