// rtw_image.cu -- the steps either side of the hot path:
//   * 8-bit image output: the reference returns Matrix{RGB{T}} and leaves saving to Images.jl (README "save image"
//     was never done, README.md:138,170); `save("x.png", img)` there maps every channel through clamp01nan and
//     N0f8 (round(x * 255)).  quantize_rgb8_kernel does that mapping on the device, PPM / PNG writers put it on disk
//     without any image library (PNG with stored deflate blocks: no zlib dependency).
//   * the flattened-scene file (.rtwscene): the SoA arrays that cross the C-ABI (rtw_set_scene), so that the Julia
//     shim, the Python harness, the oracle and the kernels share one fixture format.
// Host code only, except for the one quantisation kernel.  No C++ exception leaves this file.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/rtw_b200.h"
#include "rtw_kernels.h"

namespace rtw {

namespace {

// in: Julia column-major Float32 image (pixel (i0, j0) at ((j0*H)+i0)*3), post-gamma, unclamped
// out: row-major 8-bit RGB, top row first (the order PPM and PNG store)
__global__ void __launch_bounds__(256) quantize_rgb8_kernel(const float* __restrict__ img, int W, int H,
                                                            unsigned char* __restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)W * H) return;
    const int i0 = (int)(t / W), j0 = (int)(t - (long long)i0 * W);
    const float* src = img + ((long long)j0 * H + i0) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float x = src[c];
        x = x >= 0.0f ? x : 0.0f;  // clamp01nan: NaN and negatives -> 0
        x = x <= 1.0f ? x : 1.0f;
        out[t * 3 + c] = (unsigned char)__float2int_rn(x * 255.0f);  // N0f8(x) = round(x * 255)
    }
}

uint32_t crc_table[256];
bool crc_ready = false;

void crc_init() {
    if (crc_ready) return;
    for (uint32_t n = 0; n < 256; ++n) {
        uint32_t c = n;
        for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        crc_table[n] = c;
    }
    crc_ready = true;
}

void put_be32(std::vector<unsigned char>& v, uint32_t x) {
    v.push_back((unsigned char)(x >> 24));
    v.push_back((unsigned char)(x >> 16));
    v.push_back((unsigned char)(x >> 8));
    v.push_back((unsigned char)x);
}

bool write_chunk(FILE* f, const char type[4], const std::vector<unsigned char>& data) {
    std::vector<unsigned char> head;
    put_be32(head, (uint32_t)data.size());
    if (fwrite(head.data(), 1, 4, f) != 4) return false;
    if (fwrite(type, 1, 4, f) != 4) return false;
    if (!data.empty() && fwrite(data.data(), 1, data.size(), f) != data.size()) return false;
    crc_init();
    uint32_t c = 0xFFFFFFFFu;  // CRC-32 over type + data
    for (int i = 0; i < 4; ++i) c = crc_table[(c ^ (unsigned char)type[i]) & 0xFFu] ^ (c >> 8);
    for (size_t i = 0; i < data.size(); ++i) c = crc_table[(c ^ data[i]) & 0xFFu] ^ (c >> 8);
    const uint32_t crc = c ^ 0xFFFFFFFFu;
    std::vector<unsigned char> tail;
    put_be32(tail, crc);
    return fwrite(tail.data(), 1, 4, f) == 4;
}

constexpr char kSceneMagic[8] = {'R', 'T', 'W', 'S', 'C', 'N', '0', '1'};

}  // namespace

cudaError_t launch_quantize_rgb8(const float* img, int W, int H, unsigned char* out, cudaStream_t stream) {
    const long long total = (long long)W * H;
    if (total <= 0) return cudaSuccess;
    quantize_rgb8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(img, W, H, out);
    return cudaGetLastError();
}

}  // namespace rtw

extern "C" {

int rtw_write_ppm(const char* path, const uint8_t* rgb8, int width, int height) {
    if (!path || (!rgb8 && width > 0 && height > 0) || width < 0 || height < 0) return RTW_E_INVALID_ARG;
    FILE* f = fopen(path, "wb");
    if (!f) return RTW_E_IO;
    bool ok = fprintf(f, "P6\n%d %d\n255\n", width, height) > 0;
    const size_t n = (size_t)width * (size_t)height * 3;
    if (ok && n) ok = fwrite(rgb8, 1, n, f) == n;
    ok = (fclose(f) == 0) && ok;
    return ok ? RTW_OK : RTW_E_IO;
}

int rtw_write_png(const char* path, const uint8_t* rgb8, int width, int height) {
    if (!path || !rgb8 || width < 1 || height < 1) return RTW_E_INVALID_ARG;
    try {
        // raw scanlines: filter byte 0 + row
        const size_t row = (size_t)width * 3, raw_n = (row + 1) * (size_t)height;
        std::vector<unsigned char> raw(raw_n);
        for (int y = 0; y < height; ++y) {
            raw[(row + 1) * y] = 0;
            std::memcpy(raw.data() + (row + 1) * y + 1, rgb8 + row * y, row);
        }
        // zlib stream of stored (uncompressed) deflate blocks
        std::vector<unsigned char> z;
        z.reserve(raw_n + raw_n / 65535 * 5 + 16);
        z.push_back(0x78);
        z.push_back(0x01);
        size_t at = 0;
        do {
            const size_t len = raw_n - at < 65535 ? raw_n - at : 65535;
            z.push_back(at + len == raw_n ? 1 : 0);  // BFINAL, BTYPE = 00
            z.push_back((unsigned char)(len & 0xFF));
            z.push_back((unsigned char)(len >> 8));
            z.push_back((unsigned char)(~len & 0xFF));
            z.push_back((unsigned char)((~len >> 8) & 0xFF));
            z.insert(z.end(), raw.begin() + (long)at, raw.begin() + (long)(at + len));
            at += len;
        } while (at < raw_n);
        uint32_t a = 1, b = 0;  // adler32
        for (size_t i = 0; i < raw_n; ++i) {
            a = (a + raw[i]) % 65521u;
            b = (b + a) % 65521u;
        }
        rtw::put_be32(z, (b << 16) | a);
        FILE* f = fopen(path, "wb");
        if (!f) return RTW_E_IO;
        static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
        bool ok = fwrite(sig, 1, 8, f) == 8;
        std::vector<unsigned char> ihdr;
        rtw::put_be32(ihdr, (uint32_t)width);
        rtw::put_be32(ihdr, (uint32_t)height);
        ihdr.push_back(8);  // bit depth
        ihdr.push_back(2);  // colour type: RGB
        ihdr.push_back(0);
        ihdr.push_back(0);
        ihdr.push_back(0);
        ok = ok && rtw::write_chunk(f, "IHDR", ihdr) && rtw::write_chunk(f, "IDAT", z) &&
             rtw::write_chunk(f, "IEND", std::vector<unsigned char>());
        ok = (fclose(f) == 0) && ok;
        return ok ? RTW_OK : RTW_E_IO;
    } catch (...) {
        return RTW_E_INTERNAL;
    }
}

/* .rtwscene, little endian:
 *   0  char[8]  "RTWSCN01"
 *   8  u32      n_spheres
 *  12  u32      element type: 0 = Float32
 *  16  f32[4n]  geom4 {cx,cy,cz,radius}      (list order = HittableList order)
 *      f32[4n]  mat4  {albedo r,g,b, fuzz|ir|0}
 *      u32[n]   kind  RTW_LAMBERTIAN | RTW_METAL | RTW_DIELECTRIC
 *      u32      CRC-32 of all preceding bytes
 */
int rtw_scene_save(const char* path, const float* geom4, const float* mat4, const uint32_t* kind, uint32_t n_spheres) {
    if (!path || (n_spheres > 0 && (!geom4 || !mat4 || !kind))) return RTW_E_INVALID_ARG;
    for (uint32_t i = 0; i < n_spheres; ++i)
        if (kind[i] > RTW_DIELECTRIC) return RTW_E_UNSUPPORTED;
    FILE* f = fopen(path, "wb");
    if (!f) return RTW_E_IO;
    unsigned char head[16];
    std::memcpy(head, rtw::kSceneMagic, 8);
    const uint32_t zero = 0;
    std::memcpy(head + 8, &n_spheres, 4);
    std::memcpy(head + 12, &zero, 4);
    uint32_t c = 0xFFFFFFFFu;
    rtw::crc_init();
    auto put = [&](const void* p, size_t n) -> bool {
        const unsigned char* q = (const unsigned char*)p;
        for (size_t i = 0; i < n; ++i) c = rtw::crc_table[(c ^ q[i]) & 0xFFu] ^ (c >> 8);
        return n == 0 || fwrite(p, 1, n, f) == n;
    };
    bool ok = put(head, 16) && put(geom4, (size_t)n_spheres * 16) && put(mat4, (size_t)n_spheres * 16) &&
              put(kind, (size_t)n_spheres * 4);
    const uint32_t crc = c ^ 0xFFFFFFFFu;
    ok = ok && fwrite(&crc, 1, 4, f) == 4;
    ok = (fclose(f) == 0) && ok;
    return ok ? RTW_OK : RTW_E_IO;
}

int rtw_scene_load(const char* path, float* geom4, float* mat4, uint32_t* kind, uint32_t capacity, uint32_t* n_spheres) {
    if (!path || !n_spheres) return RTW_E_INVALID_ARG;
    *n_spheres = 0;
    FILE* f = fopen(path, "rb");
    if (!f) return RTW_E_IO;
    unsigned char head[16];
    int rc = RTW_OK;
    uint32_t n = 0, type = 0;
    if (fread(head, 1, 16, f) != 16 || std::memcmp(head, rtw::kSceneMagic, 8) != 0) {
        rc = RTW_E_FORMAT;
    } else {
        std::memcpy(&n, head + 8, 4);
        std::memcpy(&type, head + 12, 4);
        if (type != 0 || n > (1u << 26)) rc = RTW_E_FORMAT;
    }
    if (rc == RTW_OK) {
        *n_spheres = n;  // with capacity 0 (or too small) the caller learns the size and calls again
        if (capacity >= n && n > 0 && (!geom4 || !mat4 || !kind)) rc = RTW_E_INVALID_ARG;
    }
    if (rc == RTW_OK && capacity >= n) {
        rtw::crc_init();
        uint32_t c = 0xFFFFFFFFu;
        auto eat = [&](const void* p, size_t m) {
            const unsigned char* q = (const unsigned char*)p;
            for (size_t i = 0; i < m; ++i) c = rtw::crc_table[(c ^ q[i]) & 0xFFu] ^ (c >> 8);
        };
        eat(head, 16);
        uint32_t crc = 0;
        const size_t g = (size_t)n * 16, k = (size_t)n * 4;
        if ((g && fread(geom4, 1, g, f) != g) || (g && fread(mat4, 1, g, f) != g) || (k && fread(kind, 1, k, f) != k) ||
            fread(&crc, 1, 4, f) != 4) {
            rc = RTW_E_FORMAT;
        } else {
            eat(geom4, g);
            eat(mat4, g);
            eat(kind, k);
            if ((c ^ 0xFFFFFFFFu) != crc) rc = RTW_E_FORMAT;
            for (uint32_t i = 0; rc == RTW_OK && i < n; ++i)
                if (kind[i] > RTW_DIELECTRIC) rc = RTW_E_UNSUPPORTED;
        }
    } else if (rc == RTW_OK) {
        rc = capacity == 0 ? RTW_OK : RTW_E_INVALID_ARG;  // size query succeeds; a too-small buffer is an error
    }
    fclose(f);
    return rc;
}

}  // extern "C"
