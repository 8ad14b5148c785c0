// k0_rectify.cu — K0: stereo rectification of the raw camera stacks on load.
// Replaces stereoRect::doStereoRectify (Duke/stereorect.cpp:26-34) = cv::remap(img, out, map1, map2, INTER_LINEAR)
// with the CV_16SC2 fixed-point maps of cv::initUndistortRectifyMap (:42-43), applied to every image of a
// camera's stack (Duke/mfreconstruct.cpp:127-134, Duke/reconstruct.cpp:166-175).  SURVEY.md §8f row N1.
//
// One thread per FOUR output pixels of one (scan, camera): the map entries are read once and reused for all N planes.
// Rectification maps are smooth, so the four pixels' taps almost always sit in one aligned 8-byte window of two source
// rows: four 32-bit loads per plane, the taps pulled out with one PRMT per pixel and row, the bilinear blend done as
// two IDP.2A (16-bit weight x 8-bit tap dot products), one 32-bit store.  Groups that do not fit take the per-tap
// path (also the kernel for widths that are not a multiple of 4).  Arithmetic is OpenCV's fixed point exactly: weights (32-fx)(32-fy)*32 ... as 2^15-scaled shorts
// (the (0,0) entry saturates to 32767 and the table fix-up gives the missing 1 to the diagonal tap),
// (sum + 2^14) >> 15, BORDER_CONSTANT 0.  Algorithmic traffic: 2*N bytes per pixel + 6 bytes of map.
#include <stdlib.h>

#include "slr_device.cuh"

namespace {

__global__ void __launch_bounds__(256)
k0_remap_linear(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, const short2 *__restrict__ map1,
                const uint16_t *__restrict__ map2, int W, int H, int N, int views)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int view = blockIdx.z;  // scan*2 + cam
    if (x >= W) return;
    const int cam = view & 1;
    const size_t P = (size_t)W * H;
    const size_t o = (size_t)y * W + x;
    const short2 s = map1[cam * P + o];
    const int a = map2[cam * P + o] & 1023, fx = a & 31, fy = a >> 5;
    int w0 = (32 - fx) * (32 - fy) * 32, w1 = fx * (32 - fy) * 32, w2 = (32 - fx) * fy * 32, w3 = fx * fy * 32;
    if (a == 0) {
        w0 = 32767;
        w3 = 1;
    }
    const int sx = s.x, sy = s.y;
    const bool in0 = (unsigned)sx < (unsigned)W, in1 = (unsigned)(sx + 1) < (unsigned)W;
    const bool iy0 = (unsigned)sy < (unsigned)H, iy1 = (unsigned)(sy + 1) < (unsigned)H;
    const long long base = (long long)sy * W + sx;
    const uint8_t *sv = src + (size_t)view * N * P;
    uint8_t *dv = dst + (size_t)view * N * P;
    for (int n = 0; n < N; n++) {
        const uint8_t *pl = sv + (size_t)n * P;
        const int p00 = (in0 && iy0) ? __ldg(pl + base) : 0;
        const int p01 = (in1 && iy0) ? __ldg(pl + base + 1) : 0;
        const int p10 = (in0 && iy1) ? __ldg(pl + base + W) : 0;
        const int p11 = (in1 && iy1) ? __ldg(pl + base + W + 1) : 0;
        dv[(size_t)n * P + o] = (uint8_t)((p00 * w0 + p01 * w1 + p10 * w2 + p11 * w3 + (1 << 14)) >> 15);
    }
}

// Four consecutive output pixels per thread.  Rectification maps are smooth: the four pixels almost always read the
// same two source rows and source columns that fit one aligned 8-byte window, so each plane costs four aligned
// 32-bit loads instead of sixteen byte gathers, and one 32-bit store instead of four byte stores.  Groups that do not
// fit (row change inside the group, image border, strong local distortion) take the per-tap path; the arithmetic is
// the same fixed point either way.
__device__ __forceinline__ void remap_weights(int a, int &w0, int &w1, int &w2, int &w3)
{
    const int fx = a & 31, fy = a >> 5;
    w0 = (32 - fx) * (32 - fy) * 32, w1 = fx * (32 - fy) * 32, w2 = (32 - fx) * fy * 32, w3 = fx * fy * 32;
    if (a == 0) {
        w0 = 32767;
        w3 = 1;
    }
}

template <int UNROLL, int MINB>
__global__ void __launch_bounds__(128, MINB)
k0_remap_linear_x4(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, const short2 *__restrict__ map1,
                   const uint16_t *__restrict__ map2, int W, int H, int N, int views)
{
    const int x = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int y = blockIdx.y;
    const int view = blockIdx.z;  // scan*2 + cam
    if (x >= W) return;
    const int cam = view & 1;
    const size_t P = (size_t)W * H;
    const size_t o = (size_t)y * W + x;
    const uint4 m1 = *reinterpret_cast<const uint4 *>(map1 + cam * P + o);   // 4 x short2
    const uint2 m2 = *reinterpret_cast<const uint2 *>(map2 + cam * P + o);   // 4 x u16
    int sx[4], sy[4], w0[4], w1[4], w2[4], w3[4];
    const uint32_t m1w[4] = {m1.x, m1.y, m1.z, m1.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        sx[i] = (int)(short)(m1w[i] & 0xffffu);
        sy[i] = (int)(short)(m1w[i] >> 16);
        const int a = (int)(((i < 2 ? m2.x : m2.y) >> (16 * (i & 1))) & 1023u);
        remap_weights(a, w0[i], w1[i], w2[i], w3[i]);
    }
    const int bx = sx[0] & ~3;   // aligned window [bx, bx+8) x rows {sy0, sy0+1}
    bool fast = bx >= 0 && bx + 8 <= W && sy[0] >= 0 && sy[0] + 1 < H;
#pragma unroll
    for (int i = 0; i < 4; i++) fast = fast && sy[i] == sy[0] && sx[i] >= bx && sx[i] + 1 < bx + 8;
    const uint8_t *sv = src + (size_t)view * N * P;
    uint8_t *dv = dst + (size_t)view * N * P;
    if (fast) {
        const size_t wb = (size_t)sy[0] * W + bx;
        // per pixel: PRMT selector that pulls its two horizontally adjacent taps out of the 8-byte window, and the
        // four weights as two pairs of 16-bit lanes for DP2A (w0 <= 32767 after the table fix-up)
        uint32_t sel[4], w01[4], w23[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t off = (uint32_t)(sx[i] - bx);            // 0..6
            sel[i] = off | ((off + 1u) << 4) | 0x4400u;
            w01[i] = (uint32_t)w0[i] | ((uint32_t)w1[i] << 16);
            w23[i] = (uint32_t)w2[i] | ((uint32_t)w3[i] << 16);
        }
#pragma unroll UNROLL   // several planes' loads in flight per thread: the kernel is bound by L2 latency, not issue
        for (int n = 0; n < N; n++) {
            const uint32_t *q0 = reinterpret_cast<const uint32_t *>(sv + (size_t)n * P + wb);   // 4-byte aligned
            const uint32_t *q1 = reinterpret_cast<const uint32_t *>(sv + (size_t)n * P + wb + W);
            const uint32_t a0 = __ldg(q0), a1 = __ldg(q0 + 1), b0 = __ldg(q1), b1 = __ldg(q1 + 1);
            uint32_t r[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t t0 = __byte_perm(a0, a1, sel[i]), t1 = __byte_perm(b0, b1, sel[i]);   // [p00 p01 . .], [p10 p11 . .]
                uint32_t acc = __dp2a_lo(w01[i], t0, 1u << 14);     // p00*w0 + p01*w1 + 2^14
                acc = __dp2a_lo(w23[i], t1, acc);                   // + p10*w2 + p11*w3
                r[i] = acc >> 15;
            }
            const uint32_t out = __byte_perm(__byte_perm(r[0], r[1], 0x0040), __byte_perm(r[2], r[3], 0x0040), 0x5410);
            *reinterpret_cast<uint32_t *>(dv + (size_t)n * P + o) = out;
        }
    } else {
        for (int n = 0; n < N; n++) {
            const uint8_t *pl = sv + (size_t)n * P;
            uint32_t out = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const bool in0 = (unsigned)sx[i] < (unsigned)W, in1 = (unsigned)(sx[i] + 1) < (unsigned)W;
                const bool iy0 = (unsigned)sy[i] < (unsigned)H, iy1 = (unsigned)(sy[i] + 1) < (unsigned)H;
                const long long base = (long long)sy[i] * W + sx[i];
                const int p00 = (in0 && iy0) ? __ldg(pl + base) : 0;
                const int p01 = (in1 && iy0) ? __ldg(pl + base + 1) : 0;
                const int p10 = (in0 && iy1) ? __ldg(pl + base + W) : 0;
                const int p11 = (in1 && iy1) ? __ldg(pl + base + W + 1) : 0;
                out |= (uint32_t)((p00 * w0[i] + p01 * w1[i] + p10 * w2[i] + p11 * w3[i] + (1 << 14)) >> 15) << (8 * i);
            }
            *reinterpret_cast<uint32_t *>(dv + (size_t)n * P + o) = out;
        }
    }
}

}  // namespace

slr_status slr_launch_rectify(slr_engine *e, const uint8_t *d_raw, int batch, int N, uint8_t *d_out)
{
    // vector map loads (4 pixels = 16 + 8 bytes) and 4-byte aligned source windows
    if (e->W % 4 == 0 && (((uintptr_t)d_raw | (uintptr_t)d_out) % 4) == 0) {
        const int q = e->W / 4;                     // 4-pixel groups per row
        const int tx = (q % 128 == 0 || q > 512) ? 128 : (q % 64 == 0 ? 64 : (q % 32 == 0 ? 32 : 128));
        dim3 block(tx), grid((q + tx - 1) / tx, e->H, batch * 2);
        // two planes in flight per thread, 64 registers (8 CTAs of 128 threads per SM) measured best
        k0_remap_linear_x4<2, 8><<<grid, block, 0, e->stream>>>(d_raw, d_out, (const short2 *)e->d_map1, e->d_map2, e->W,
                                                               e->H, N, batch * 2);
        SLR_CHECK_LAUNCH(e);
        return SLR_OK;
    }
    dim3 block(256), grid((e->W + 255) / 256, e->H, batch * 2);
    k0_remap_linear<<<grid, block, 0, e->stream>>>(d_raw, d_out, (const short2 *)e->d_map1, e->d_map2, e->W, e->H, N,
                                                   batch * 2);
    SLR_CHECK_LAUNCH(e);
    return SLR_OK;
}
