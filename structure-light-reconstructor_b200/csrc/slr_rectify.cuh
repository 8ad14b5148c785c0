// slr_rectify.cuh — cv::remap(INTER_LINEAR, CV_16SC2 maps, BORDER_CONSTANT 0) of a camera row, shared by the stand-alone
// K0 (k0_rectify.cu) and the rectify jobs of the fused kernel (k_fused_flow.cu <RAW>).
// stereoRect::doStereoRectify, Duke/stereorect.cpp:26-34.
#pragma once
#include "slr_device.cuh"

namespace slr {

// OpenCV's fixed-point bilinear weights of one CV_16SC2 map entry (5 + 5 fractional bits), as two pairs of 16-bit lanes
// for DP2A: (32-fx)(32-fy)*32 ... = 2^15-scaled shorts; the (0,0) entry saturates to 32767 and cv::remap's table fix-up
// gives the missing 1 to the diagonal tap.
__device__ __forceinline__ void remap_weights(uint32_t m2, uint32_t &w01, uint32_t &w23)
{
    const int a = (int)(m2 & 1023u), fx = a & 31, fy = a >> 5;
    int w0 = (32 - fx) * (32 - fy) * 32, w1 = fx * (32 - fy) * 32, w2 = (32 - fx) * fy * 32, w3 = fx * fy * 32;
    if (a == 0) w0 = 32767, w3 = 1;
    w01 = (uint32_t)w0 | ((uint32_t)w1 << 16);
    w23 = (uint32_t)w2 | ((uint32_t)w3 << 16);
}

// A rectify job: cv::remap(INTER_LINEAR, BORDER_CONSTANT 0) of 128 consecutive output pixels of row i of one camera, all
// N planes, written to the stage rows; called by the whole warp, lane l owns the four pixels x .. x+3 (active = x < W),
// dst = stage address of pixel x of plane 0.  Same arithmetic as k0_rectify.cu (the stand-alone K0).
//
// Rectification maps are smooth: a lane's four pixels almost always read the same two source rows and source columns
// that fit one aligned 8-byte window, so a plane costs it four aligned 32-bit loads, a PRMT per pixel and row, two DP2A
// per pixel and one 32-bit shared-memory store.  The source row of a slightly rotated camera changes every hundred
// pixels or so, i.e. in about ONE lane of most warps; a per-lane fallback would make the whole warp sit through its
// 14 x 4 x 4 byte loads, so the lanes that do not fit are served by the warp together afterwards: one (plane, pixel)
// item per lane, the map entry re-read (it is in L1), four byte taps, one byte stored.
// dst = address of pixel x of plane 0 (shared or global memory), planes dst_plane_stride bytes apart.  LEAD > 0: the
// source row an output row LEAD rows further down will need is prefetched into L2.
template <int LEAD>
__device__ __forceinline__ void rectify_job(const uint8_t *__restrict__ sv /* raw [N][H][W] of this scan + camera */,
                                            const short2 *__restrict__ map1, const uint16_t *__restrict__ map2 /* this camera */,
                                            int W, int H, int N, int i, int x, bool active, unsigned char *dst,
                                            size_t dst_plane_stride, int lane)
{
    const size_t P = (size_t)W * H;
    const size_t o = (size_t)i * W + x;
    bool fast = false;
    int sx[4], sy[4], bx = 0;
    uint32_t w01[4], w23[4];
    if (active) {
        const uint4 m1 = __ldg(reinterpret_cast<const uint4 *>(map1 + o));   // 4 x short2
        const uint2 m2 = __ldg(reinterpret_cast<const uint2 *>(map2 + o));   // 4 x u16
        const uint32_t m1w[4] = {m1.x, m1.y, m1.z, m1.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            sx[k] = (int)(short)(m1w[k] & 0xffffu);
            sy[k] = (int)(short)(m1w[k] >> 16);
            remap_weights((k < 2 ? m2.x : m2.y) >> (16 * (k & 1)), w01[k], w23[k]);
        }
        bx = sx[0] & ~3;   // aligned window [bx, bx+8) x rows {sy0, sy0+1}
        fast = bx >= 0 && bx + 8 <= W && sy[0] >= 0 && sy[0] + 1 < H;
#pragma unroll
        for (int k = 0; k < 4; k++) fast = fast && sy[k] == sy[0] && sx[k] >= bx && sx[k] + 1 < bx + 8;
    }
    const unsigned fmask = __ballot_sync(0xffffffffu, fast);             // the lanes on the fast path
    {
        if (fast) {
            uint32_t sel[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t off = (uint32_t)(sx[k] - bx);            // 0..6
                sel[k] = off | ((off + 1u) << 4) | 0x4400u;
            }
            const uint8_t *q = sv + (size_t)sy[0] * W + bx;              // 4-byte aligned
            // An output row needs one source row per plane that no earlier row has touched (the lower tap row): pull the
            // one the row LEAD steps ahead will need from HBM into L2 now (one 32-byte sector per 8 lanes).
            if (LEAD > 0 && (lane & 7) == 0 && sy[0] + 1 + LEAD < H) {
                const uint8_t *pf = q + (size_t)(1 + LEAD) * W;
                for (int n = 0; n < N; n++, pf += P) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
            }
            // The job is bound by how many loads it keeps in flight (L2 latency x 14 planes), so a lane loads only the
            // FIRST word of its two window rows and takes the second from its right neighbour, whose window usually
            // starts right there; four planes' loads are issued before the first blend.
            const int nsy = __shfl_down_sync(fmask, sy[0], 1), nbx = __shfl_down_sync(fmask, bx, 1);
            const bool share = lane < 31 && ((fmask >> (lane + 1)) & 1u) && nsy == sy[0] && nbx == bx + 4;
            auto blend_store = [&](int n, uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1) {
                uint32_t r[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t t0 = __byte_perm(a0, a1, sel[k]), t1 = __byte_perm(b0, b1, sel[k]);
                    uint32_t acc = __dp2a_lo(w01[k], t0, 1u << 14);     // p00*w0 + p01*w1 + 2^14
                    acc = __dp2a_lo(w23[k], t1, acc);                   // + p10*w2 + p11*w3
                    r[k] = acc >> 15;
                }
                *reinterpret_cast<uint32_t *>(dst + (size_t)n * dst_plane_stride) =
                    __byte_perm(__byte_perm(r[0], r[1], 0x0040), __byte_perm(r[2], r[3], 0x0040), 0x5410);
            };
            constexpr int DEPTH = 4;
            for (int n0 = 0; n0 < N; n0 += DEPTH) {
                uint32_t a0[DEPTH], b0[DEPTH], a1[DEPTH], b1[DEPTH];
#pragma unroll
                for (int d = 0; d < DEPTH; d++) {
                    a0[d] = b0[d] = a1[d] = b1[d] = 0u;
                    if (n0 + d < N) {
                        const uint8_t *qd = q + (size_t)(n0 + d) * P;
                        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(a0[d]) : "l"(qd));
                        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(b0[d]) : "l"(qd + W));
                        if (!share) {
                            asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(a1[d]) : "l"(qd + 4));
                            asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(b1[d]) : "l"(qd + W + 4));
                        }
                    }
                }
#pragma unroll
                for (int d = 0; d < DEPTH; d++) {
                    const uint32_t na = __shfl_down_sync(fmask, a0[d], 1), nb = __shfl_down_sync(fmask, b0[d], 1);
                    if (n0 + d < N) blend_store(n0 + d, a0[d], share ? na : a1[d], b0[d], share ? nb : b1[d]);
                }
            }
        }
    }
    unsigned todo = __ballot_sync(0xffffffffu, active && !fast);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int xs = __shfl_sync(0xffffffffu, x, src);
        const unsigned long long dptr = (unsigned long long)dst;
        unsigned char *dsts = reinterpret_cast<unsigned char *>(
            ((unsigned long long)__shfl_sync(0xffffffffu, (unsigned)(dptr >> 32), src) << 32) | __shfl_sync(0xffffffffu, (unsigned)dptr, src));
        for (int item = lane; item < 4 * N; item += 32) {
            const int n = item >> 2, k = item & 3;
            const size_t ok = (size_t)i * W + xs + k;
            const short2 m = __ldg(map1 + ok);
            uint32_t w01, w23;
            remap_weights(__ldg(map2 + ok), w01, w23);
            const int sx = m.x, sy = m.y;
            const bool in0 = (unsigned)sx < (unsigned)W, in1 = (unsigned)(sx + 1) < (unsigned)W;
            const bool iy0 = (unsigned)sy < (unsigned)H, iy1 = (unsigned)(sy + 1) < (unsigned)H;
            const uint8_t *pl = sv + (size_t)n * P + ((long long)sy * W + sx);
            const uint32_t p00 = (in0 && iy0) ? __ldg(pl) : 0u, p01 = (in1 && iy0) ? __ldg(pl + 1) : 0u;
            const uint32_t p10 = (in0 && iy1) ? __ldg(pl + W) : 0u, p11 = (in1 && iy1) ? __ldg(pl + W + 1) : 0u;
            uint32_t acc = __dp2a_lo(w01, p00 | (p01 << 8), 1u << 14);
            acc = __dp2a_lo(w23, p10 | (p11 << 8), acc);
            dsts[(size_t)n * dst_plane_stride + k] = (unsigned char)(acc >> 15);
        }
    }
}

}  // namespace slr
