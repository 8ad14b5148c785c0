// k0_rectify.cu — K0: stereo rectification of the raw camera stacks on load.
// Replaces stereoRect::doStereoRectify (Duke/stereorect.cpp:26-34) = cv::remap(img, out, map1, map2, INTER_LINEAR)
// with the CV_16SC2 fixed-point maps of cv::initUndistortRectifyMap (:42-43), applied to every image of a
// camera's stack (Duke/mfreconstruct.cpp:127-134, Duke/reconstruct.cpp:166-175).  SURVEY.md §8f row N1.
//
// One warp per 128 output pixels of one (scan, camera) row, four pixels per lane: the map entries are read once and
// reused for all N planes.  Rectification maps are smooth, so the four pixels' taps almost always sit in one aligned
// 8-byte window of two source rows: two 32-bit loads per plane (the other two words come from the right neighbour by
// shuffle), the taps pulled out with one PRMT per pixel and row, the bilinear blend done as two IDP.2A (16-bit weight x
// 8-bit tap dot products), one 32-bit store.  Lanes that do not fit are served by the whole warp afterwards
// (slr_rectify.cuh); widths that are not a multiple of 4 take the per-pixel kernel.  Arithmetic is OpenCV's fixed point
// exactly: weights (32-fx)(32-fy)*32 ... as 2^15-scaled shorts (the (0,0) entry saturates to 32767 and the table fix-up
// gives the missing 1 to the diagonal tap), (sum + 2^14) >> 15, BORDER_CONSTANT 0.  Algorithmic traffic: 2*N bytes per
// pixel + 6 bytes of map.
#include <stdlib.h>

#include "slr_rectify.cuh"

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

// One warp per 128 consecutive output pixels of one row of one (scan, camera), all N planes: slr::rectify_job
// (slr_rectify.cuh), the same code the fused kernel's rectify jobs run.  A lane owns four pixels whose taps almost always
// sit in one aligned 8-byte window of two source rows; it loads the first word of each row and takes the second from
// its right neighbour, four planes' loads are in flight before the first blend, the blend is two DP2A per pixel; the
// lanes that do not fit (a source-row change inside the group, the image border) are served by the warp together.
__global__ void __launch_bounds__(1024)
k0_remap_warp(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, const short2 *__restrict__ map1,
              const uint16_t *__restrict__ map2, int W, int H, int N)
{
    const int lane = threadIdx.x & 31;
    const int x = 4 * (int)threadIdx.x;       // blockDim.x = 32 * ceil(W / 128): one warp per 128-pixel segment
    const int y = blockIdx.x;
    const int view = blockIdx.y;              // scan*2 + cam
    const int cam = view & 1;
    const size_t P = (size_t)W * H;
    slr::rectify_job<0>(src + (size_t)view * N * P, map1 + (size_t)cam * P, map2 + (size_t)cam * P, W, H, N, y, x, x < W,
                        dst + (size_t)view * N * P + (size_t)y * W + x, P, lane);
}

}  // namespace

slr_status slr_launch_rectify(slr_engine *e, const uint8_t *d_raw, int batch, int N, uint8_t *d_out)
{
    // vector map loads (4 pixels = 16 + 8 bytes) and 4-byte aligned source windows; one warp per 128 pixels of a row
    if (e->W % 4 == 0 && e->W <= 4096 && (((uintptr_t)d_raw | (uintptr_t)d_out) % 4) == 0 && batch * 2 <= 65535) {
        dim3 block(32 * ((e->W + 127) / 128)), grid(e->H, batch * 2);
        k0_remap_warp<<<grid, block, 0, e->stream>>>(d_raw, d_out, (const short2 *)e->d_map1, e->d_map2, e->W, e->H, N);
        SLR_CHECK_LAUNCH(e);
        return SLR_OK;
    }
    dim3 block(256), grid((e->W + 255) / 256, e->H, batch * 2);
    k0_remap_linear<<<grid, block, 0, e->stream>>>(d_raw, d_out, (const short2 *)e->d_map1, e->d_map2, e->W, e->H, N,
                                                   batch * 2);
    SLR_CHECK_LAUNCH(e);
    return SLR_OK;
}
