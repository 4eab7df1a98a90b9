// rtw_small.cu -- the latency path: a whole small render() in ONE kernel launch.
//
// The reference's own smoke test and its published micro-timings are tiny images (96x54 at 1 and 16 samples,
// test/runtests.jl:194, src/proto/proto.jl:64-66, 87-89).  At that size the persistent kernel is bound by its stream
// operations (accumulator + counter memsets, u/v tables, trace, resolve, two copies), not by tracing.  This kernel does
// the same arithmetic with none of them:
//   * a pixel is owned by a group of g = min(32, 2^ceil(log2 spp)) adjacent lanes; lane j of the group traces the samples
//     j, j + g, ... of that pixel (src/render.jl:29-39), one path at a time, against the list in shared memory
//     (src/hit.jl:38-50 in list order);
//   * each path's Float64 radiance becomes the same fixed-point integers the persistent kernel adds with atomics
//     (rtw_fused2.cu), the lanes of the group add them with shuffles -- integer sums, so the pixel has the same bits --
//     and lane 0 writes sqrt(sum / spp) (src/render.jl:40, src/vec.jl:22) straight into the caller-visible image
//     (mapped pinned host memory: no device-to-host copy is enqueued);
//   * the last CTA to finish publishes the ray-segment count to mapped host memory and re-zeroes the two device counters,
//     so the next call needs no memset either.
// Same FP contract, same addressed Philox stream (rtw_device.cuh): bit-identical to the persistent kernel and the oracle.
#include "rtw_kernels.h"

namespace rtw {

namespace {

constexpr int kSmallBlock = 64;  // small CTAs: a 96x54x1 render still spreads over every SM

__global__ void __launch_bounds__(kSmallBlock) small_render_kernel(const __grid_constant__ TraceParams P, int group_log2,
                                                                   double inv_scale, float* __restrict__ out_img,
                                                                   unsigned long long* __restrict__ host_totals) {
    extern __shared__ __align__(16) float4 s_list[];
    const uint32_t n = P.n_spheres;
    for (uint32_t i = threadIdx.x; i < n; i += kSmallBlock) s_list[i] = P.geom[i];
    __syncthreads();

    const uint32_t g = 1u << group_log2;                 // lanes per pixel
    const unsigned long long gtid = (unsigned long long)blockIdx.x * kSmallBlock + threadIdx.x;
    const unsigned long long npix = (unsigned long long)P.n_rows * (unsigned long long)P.W;
    const unsigned long long pl = gtid >> group_log2;    // local pixel index (row-major over the rows of this call)
    const uint32_t j = (uint32_t)gtid & (g - 1u);        // lane of the group
    const bool active = pl < npix;
    const float tmin = 1e-4f;  // T(1e-4), src/ray_color.jl:19

    long long acc_r = 0, acc_g = 0, acc_b = 0;
    uint32_t seg_count = 0;
    uint32_t i0 = 0, col = 0;
    if (active) {
        const uint32_t row_local = (uint32_t)(pl / (unsigned)P.W);
        col = (uint32_t)(pl - (unsigned long long)row_local * (unsigned)P.W);
        i0 = (uint32_t)P.row_start + row_local * (uint32_t)P.row_stride;
        const float u_base = __fdiv_rn((float)(col + 1u), (float)P.W);                    // T(j/W), src/render.jl:26
        const float v_base = __fdiv_rn((float)((uint32_t)P.H - 1u - i0), (float)P.H);     // T((H-i)/H), src/render.jl:27
        PathRng rng;
        rng.pixel = i0 * (uint32_t)P.W + col;
        for (uint32_t s0 = j; s0 < (uint32_t)P.spp; s0 += g) {  // src/render.jl:29
            rng.sample = s0 + (uint32_t)P.sample_first;
            f3 o, d;
            primary_ray(P.cam, rng, P.key0, P.key1, rng.sample, u_base, v_base, (float)P.W, (float)P.H, o, d);
            double thr_r = 1.0, thr_g = 1.0, thr_b = 1.0;
            for (uint32_t nhits = 0;;) {  // ray_color, src/ray_color.jl:14-38, unrolled into a loop
                if ((int)nhits >= P.max_depth) break;  // depth exhausted: black (src/ray_color.jl:15-17)
                float best_t = __int_as_float(0x7f800000);
                int best_k = -1;
                // hit(::HittableList), src/hit.jl:38-50, in list order; 8 independent discriminants in flight (a lone
                // warp per scheduler is bound by the latency of the LDS -> 11 FP32 chain, not by issue)
                uint32_t k = 0;
                for (; k + 8u <= n; k += 8u) {
                    float hb[8], disc[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) disc[u] = sphere_disc(s_list[k + u], o, d, hb[u]);
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (!(disc[u] < 0.0f) && sphere_accept(disc[u], hb[u], tmin, best_t)) best_k = (int)(k + u);
                }
                for (; k < n; ++k) {
                    float hb;
                    const float disc = sphere_disc(s_list[k], o, d, hb);
                    if (!(disc < 0.0f) && sphere_accept(disc, hb, tmin, best_t)) best_k = (int)k;
                }
                seg_count += 1;
                if (best_k < 0) {  // skycolor, src/ray_color.jl:1-6 (Float64 literals)
                    const float t = 0.5f * (d.y + 1.0f);
                    const double a1 = (double)(1.0f - t), w1 = (double)t;
                    const double sr = __dadd_rn(a1, __dmul_rn(w1, 0.5));
                    const double sg = __dadd_rn(a1, __dmul_rn(w1, 0.7));
                    const double sb = __dadd_rn(a1, w1);
                    acc_r += __double2ll_rn(__dmul_rn(thr_r, sr) * P.fx_scale);
                    acc_g += __double2ll_rn(__dmul_rn(thr_g, sg) * P.fx_scale);
                    acc_b += __double2ll_rn(__dmul_rn(thr_b, sb) * P.fx_scale);
                    break;
                }
                nhits += 1;
                if ((int)nhits >= P.max_depth) break;  // the next ray_color call returns black
                const uint32_t kind = __ldg(P.kind + best_k);
                const float4 m = __ldg(P.mat + best_k);
                f3 att;
                shade_hit(o, d, best_t, s_list[best_k], m, kind, rng, nhits, P.key0, P.key1, att);
                if (kind != 2u) {  // attenuation = albedo (Dielectric: (1,1,1), an exact no-op)
                    thr_r = __dmul_rn(thr_r, (double)m.x);
                    thr_g = __dmul_rn(thr_g, (double)m.y);
                    thr_b = __dmul_rn(thr_b, (double)m.z);
                }
            }
        }
    }
    // the samples of a pixel live in g adjacent lanes of one warp: integer sums by shuffle (order-independent)
    for (uint32_t off = g >> 1; off > 0; off >>= 1) {
        acc_r += __shfl_xor_sync(0xffffffffu, acc_r, off);
        acc_g += __shfl_xor_sync(0xffffffffu, acc_g, off);
        acc_b += __shfl_xor_sync(0xffffffffu, acc_b, off);
    }
    if (active && j == 0u) {
        const long long at = ((long long)col * P.H + (long long)i0) * 3;  // Julia column-major Matrix{RGB{Float32}}(H, W)
        const long long a[3] = {acc_r, acc_g, acc_b};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double sum = (double)a[c] * inv_scale;
            const double lin = sum / (double)(P.spp);  // accum_color / n_samples, src/render.jl:40
            out_img[at + c] = (float)sqrt(lin);        // rgb_gamma2, src/vec.jl:22
        }
    }
    // ray-segment statistics + completion: the last CTA publishes the total and re-zeroes the counters
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

}  // namespace

// One launch = one whole render of the rows (row_start, row_stride) of `p`; needs p.n_spheres <= kTileSpheres and both
// device counters zero on entry (the kernel leaves them zero).  out_img / host_totals: device-visible pointers (mapped
// pinned host memory in the latency path).  p.spp samples per pixel, p.fx_scale as for the persistent kernel.
cudaError_t launch_small_render(const TraceParams& p, double inv_scale, float* out_img, unsigned long long* host_totals,
                                cudaStream_t stream, LaunchInfo* info) {
    int group_log2 = 0;
    while ((1 << group_log2) < p.spp && group_log2 < 5) ++group_log2;
    const unsigned long long threads = ((unsigned long long)p.n_rows * (unsigned long long)p.W) << group_log2;
    const unsigned long long blocks = (threads + kSmallBlock - 1) / kSmallBlock;
    if (blocks == 0 || blocks > 0x7fffffffull) return cudaErrorInvalidValue;
    const int smem = (int)(p.n_spheres * sizeof(float4));
    small_render_kernel<<<(unsigned)blocks, kSmallBlock, smem, stream>>>(p, group_log2, inv_scale, out_img, host_totals);
    if (info) {
        info->grid = (int)blocks;
        info->block = kSmallBlock;
        info->smem_bytes = smem;
        info->blocks_per_sm = 0;
        info->launches = 1;
        info->rays_per_lane = 1;
        info->sweep = kSweepBranch;
    }
    return cudaGetLastError();
}

}  // namespace rtw
