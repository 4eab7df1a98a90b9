// The packed mask sweep (11 FP32x2 instructions + 2 funnel shifts per sphere pair and ray, src/hit.jl:13-19) with the sphere
// pairs fetched (a) from shared memory by LDS.128 (one ray per lane, no lane cooperation), (b) from the constant bank
// through the uniform datapath (LDCU -> uniform registers as FP32x2 operands).  Question: do uniform loads issue for free?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o sweep_const_bench sweep_const_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int kSpheres = 512;   // 256 pairs
constexpr int kSweeps = 64;
__constant__ float4 c_pairs[kSpheres];  // pair layout: {xa,xb,ya,yb} {za,zb,ra,rb}

__device__ __forceinline__ float2 dup2(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 neg2(float a, float b) { return make_float2(-a, -b); }

__device__ __forceinline__ void test_pair(const float4 A, const float4 B, float ox, float oy, float oz, float dx, float dy, float dz, uint32_t& m) {
    const float2 ocx = __fadd2_rn(dup2(ox), neg2(A.x, A.y));
    const float2 ocy = __fadd2_rn(dup2(oy), neg2(A.z, A.w));
    const float2 ocz = __fadd2_rn(dup2(oz), neg2(B.x, B.y));
    const float2 hb = __ffma2_rn(ocz, dup2(dz), __ffma2_rn(ocy, dup2(dy), __fmul2_rn(ocx, dup2(dx))));
    const float2 q = __ffma2_rn(ocz, ocz, __ffma2_rn(ocy, ocy, __fmul2_rn(ocx, ocx)));
    const float2 rr = make_float2(B.z, B.w);
    const float2 cq = __ffma2_rn(neg2(rr.x, rr.y), rr, q);
    const float2 disc = __ffma2_rn(hb, hb, neg2(cq.x, cq.y));
    m = __funnelshift_l(__float_as_uint(disc.x), m, 1);
    m = __funnelshift_l(__float_as_uint(disc.y), m, 1);
}

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, const float* __restrict__ rays) {
    __shared__ __align__(16) float4 s_pairs[kSpheres];
    for (int i = threadIdx.x; i < kSpheres; i += blockDim.x) s_pairs[i] = c_pairs[i];
    __syncthreads();
    const float t = (float)threadIdx.x;
    float ox = rays[0] + t * 1e-3f, oy = rays[1] + t * 2e-3f, oz = rays[2] - t * 1e-3f;
    const float dx = rays[3] + t * 1e-4f, dy = rays[4] - t * 1e-4f, dz = rays[5] + t * 2e-4f;
    uint32_t acc = 0;
    for (int it = 0; it < kSweeps; ++it) {
        for (int c = 0; c < kSpheres / 32; ++c) {   // a mask word per 32 spheres
            uint32_t m = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int p = c * 16 + i;
                if (MODE == 0) test_pair(s_pairs[2 * p], s_pairs[2 * p + 1], ox, oy, oz, dx, dy, dz, m);
                else test_pair(c_pairs[2 * p], c_pairs[2 * p + 1], ox, oy, oz, dx, dy, dz, m);
            }
            acc ^= m;   // all misses: m == 0xffffffff
        }
        ox += 1e-3f;
    }
    if (acc == 0x12345u) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, uint32_t* out, const float* rays, int grid) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<grid, 256>>>(out, rays);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    printf("%-48s %7.3f ms  %.2f T FP32 lane-instr/s\n", name, (double)best,
           (double)grid * 256 * kSweeps * kSpheres * 11.0 / (best * 1e-3) / 1e12);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float4 h[kSpheres];
    for (int i = 0; i < kSpheres; ++i)
        h[i] = (i & 1) ? make_float4(-3000.f, -3000.f, 0.5f, 0.5f) : make_float4(1000.f + i, 1001.f + i, 2000.f, 2000.f);
    cudaMemcpyToSymbol(c_pairs, h, sizeof h);
    const float ray[6] = {0.125f, 0.25f, 0.5f, 0.6f, 0.0f, 0.8f};
    float* rays;
    cudaMalloc(&rays, sizeof ray);
    cudaMemcpy(rays, ray, sizeof ray, cudaMemcpyHostToDevice);
    uint32_t* out;
    cudaMalloc(&out, 64 << 20);
    const int grid = sms * 3;   // 3 CTAs of 256 per SM, as the trace kernel
    run<0>("sphere pairs by LDS.128 (1 ray/lane, no coop)", out, rays, grid);
    run<1>("sphere pairs by uniform constant loads", out, rays, grid);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
