// k2_gray_decode.cu — K2: shadow mask + Gray-code bitplane decode.
// Replaces Reconstruct::computeShadows (Duke/reconstruct.cpp:210-227), decodePatterns_GE /
// getProjPixel_GE (:79-97, :381-407), decodePaterns / getProjPixel (:56-74, :325-370) and
// GrayCodes::grayToDec (Duke/graycodes.cpp:116-128).
//
// HBM-bound streaming kernel, integer only: each thread owns 16 consecutive pixels, reads the
// white/black planes and every (pattern, inverse) pair with 128-bit loads, accumulates the Gray
// bits MSB-first in a register per pixel and converts Gray -> binary with a log-step prefix XOR.
// Algorithmic traffic: N + 4 (+4) + 1 bytes per pixel, N = 2 + 2*nbits_col + 2*nbits_row.
#include "slr_device.cuh"

namespace {

constexpr int K2_THREADS = 256;

__device__ __forceinline__ int byte_at(const uint4 &v, int i)
{
    const uint32_t w = (i < 4) ? v.x : (i < 8) ? v.y : (i < 12) ? v.z : v.w;
    return (int)((w >> (8 * (i & 3))) & 0xffu);
}

// GrayCodes::grayToDec: dec = sum_i prefixXOR_i * 2^(n-1-i)  ==  prefix XOR from the MSB down.
__device__ __forceinline__ int gray_to_binary(int g)
{
    g ^= g >> 1;
    g ^= g >> 2;
    g ^= g >> 4;
    g ^= g >> 8;
    g ^= g >> 16;
    return g;
}

template <int PX>  // 16 (vector path) or 1 (scalar path)
__global__ void __launch_bounds__(K2_THREADS)
k2_gray_decode(const uint8_t *__restrict__ stack, size_t P, long long chunks_per_view, long long total_chunks,
               int nbits_col, int nbits_row, int black_thr, int white_thr, int scan_w, int scan_h,
               int32_t *__restrict__ col, int32_t *__restrict__ row, uint8_t *__restrict__ mask)
{
    const int N = 2 + 2 * nbits_col + 2 * nbits_row;
    for (long long chunk = (long long)blockIdx.x * K2_THREADS + threadIdx.x; chunk < total_chunks;
         chunk += (long long)gridDim.x * K2_THREADS) {
        const long long view = chunk / chunks_per_view;
        const long long c = chunk - view * chunks_per_view;
        const uint8_t *src = stack + (size_t)view * N * P + (size_t)c * PX;

        int gx[PX], gy[PX];
        bool err[PX], lit[PX];
        {
            uint4 w, b;
            if (PX == 16) {
                w = slr::ldg_stream_u4(src);
                b = slr::ldg_stream_u4(src + P);
            } else {
                w = make_uint4(src[0], 0, 0, 0);
                b = make_uint4(src[P], 0, 0, 0);
            }
#pragma unroll
            for (int i = 0; i < PX; i++) {
                lit[i] = (byte_at(w, i) - byte_at(b, i)) > black_thr;  // computeShadows (:219-224)
                err[i] = false;
                gx[i] = 0;
                gy[i] = 0;
            }
        }
        const int nb_total = nbits_col + nbits_row;
        for (int bit = 0; bit < nb_total; bit++) {
            const uint8_t *pp = src + (size_t)(2 + 2 * bit) * P;
            uint4 v1, v2;
            if (PX == 16) {
                v1 = slr::ldg_stream_u4(pp);
                v2 = slr::ldg_stream_u4(pp + P);
            } else {
                v1 = make_uint4(pp[0], 0, 0, 0);
                v2 = make_uint4(pp[P], 0, 0, 0);
            }
            const bool is_col = bit < nbits_col;
#pragma unroll
            for (int i = 0; i < PX; i++) {
                const int a = byte_at(v1, i), d = byte_at(v2, i);
                err[i] |= (abs(a - d) < white_thr);  // :393
                const int on = (a > d) ? 1 : 0;      // :396
                if (is_col)
                    gx[i] = (gx[i] << 1) | on;
                else
                    gy[i] = (gy[i] << 1) | on;
            }
        }
        const size_t o = (size_t)view * P + (size_t)c * PX;
        int xs[PX], ys[PX];
        uint32_t mk[(PX + 3) / 4];
#pragma unroll
        for (int i = 0; i < (PX + 3) / 4; i++) mk[i] = 0;
#pragma unroll
        for (int i = 0; i < PX; i++) {
            const int x = gray_to_binary(gx[i]);
            const int y = gray_to_binary(gy[i]);
            bool bad = err[i] || (x > scan_w);          // :403 / :365 (strict >)
            if (nbits_row > 0) bad = bad || (y > scan_h);
            const bool m = lit[i] && !bad;
            xs[i] = m ? x : -1;
            ys[i] = (m && nbits_row > 0) ? y : -1;
            mk[i >> 2] |= (m ? 1u : 0u) << (8 * (i & 3));
        }
        if (PX == 16) {
#pragma unroll
            for (int v = 0; v < 4; v++) {
                slr::stg_stream_u4(col + o + 4 * v, make_uint4(xs[4 * v], xs[4 * v + 1], xs[4 * v + 2], xs[4 * v + 3]));
                if (row)
                    slr::stg_stream_u4(row + o + 4 * v,
                                       make_uint4(ys[4 * v], ys[4 * v + 1], ys[4 * v + 2], ys[4 * v + 3]));
            }
            slr::stg_stream_u4(mask + o, make_uint4(mk[0], mk[1], mk[2], mk[3]));
        } else {
            col[o] = xs[0];
            if (row) row[o] = ys[0];
            mask[o] = (uint8_t)(mk[0] & 0xff);
        }
    }
}

// ---- byte-parallel form (the common case: 0 <= black_thr <= 255, white_thr <= 255, <= 16 bits per axis) ----------
// One 32-bit word = 4 pixels of one plane; every comparison below runs on the four bytes at once.

struct GrayAcc {
    uint32_t lo[4], hi[4];   // [word]: gray bits 0..7 and 8..15 of 4 pixels each
};

// accumulate the nbits planes pairs starting at plane index first (MSB first in memory, Duke/graycodes.cpp:63-85)
template <bool CHECK_WHITE>
__device__ __forceinline__ void gray_axis(const uint8_t *__restrict__ src, size_t P, int first, int nbits, uint32_t thr4,
                                          GrayAcc &g, uint32_t (&bad7)[4])
{
#pragma unroll
    for (int w = 0; w < 4; w++) g.lo[w] = g.hi[w] = 0;
#pragma unroll 4
    for (int k = 0; k < nbits; k++) {          // k = 0 is the LSB = the LAST pair of the axis
        const uint8_t *pp = src + (size_t)(first + 2 * (nbits - 1 - k)) * P;
        const uint4 v1 = slr::ldg_stream_u4(pp), v2 = slr::ldg_stream_u4(pp + P);
        const uint32_t a[4] = {v1.x, v1.y, v1.z, v1.w}, d[4] = {v2.x, v2.y, v2.z, v2.w};
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const uint32_t t = slr::gt7(a[w], d[w]);                                   // bit = v1 > v2   (reconstruct.cpp:396)
            if (k < 8)
                g.lo[w] = slr::push_bit(g.lo[w], t);
            else
                g.hi[w] = slr::push_bit(g.hi[w], t);
            if (CHECK_WHITE) bad7[w] |= ~slr::gt7(__vabsdiffu4(a[w], d[w]) | 0u, thr4);  // |v1-v2| < whiteThreshold (:393)
        }
    }
    const int kl = nbits < 8 ? nbits : 8, kh = nbits - kl;                         // right-align both bytes
    const uint32_t ml = 0x01010101u * ((1u << kl) - 1u), mh = 0x01010101u * ((1u << kh) - 1u);
#pragma unroll
    for (int w = 0; w < 4; w++) {
        g.lo[w] = (g.lo[w] >> (8 - kl)) & ml;
        g.hi[w] = kh ? (g.hi[w] >> (8 - kh)) & mh : 0u;
    }
}

template <bool CHECK_WHITE, bool ROWS>
__global__ void __launch_bounds__(K2_THREADS)
k2_gray_decode_x4(const uint8_t *__restrict__ stack, size_t P, long long chunks_per_view, long long total_chunks,
                  int nbits_col, int nbits_row, int black_thr, int white_thr, int scan_w, int scan_h,
                  int32_t *__restrict__ col, int32_t *__restrict__ row, uint8_t *__restrict__ mask)
{
    const int N = 2 + 2 * nbits_col + 2 * nbits_row;
    const uint32_t bthr4 = 0x01010101u * (uint32_t)black_thr;
    // |v1-v2| < thr  <=>  !(|v1-v2| > thr-1); white_thr >= 1 here
    const uint32_t wthr4 = 0x01010101u * (uint32_t)(white_thr - 1);
    for (long long chunk = (long long)blockIdx.x * K2_THREADS + threadIdx.x; chunk < total_chunks;
         chunk += (long long)gridDim.x * K2_THREADS) {
        const long long view = chunk / chunks_per_view;
        const long long c = chunk - view * chunks_per_view;
        const uint8_t *src = stack + (size_t)view * N * P + (size_t)c * 16;
        const uint4 wv = slr::ldg_stream_u4(src), bv = slr::ldg_stream_u4(src + P);
        const uint32_t wh[4] = {wv.x, wv.y, wv.z, wv.w}, bl[4] = {bv.x, bv.y, bv.z, bv.w};
        uint32_t bad7[4];
#pragma unroll
        for (int w = 0; w < 4; w++)   // white - black > blackThreshold (>= 0)  <=>  white > black and |white-black| > thr
            bad7[w] = ~(slr::gt7(wh[w], bl[w]) & slr::gt7(__vabsdiffu4(wh[w], bl[w]), bthr4));
        GrayAcc gx, gy;
        gray_axis<CHECK_WHITE>(src, P, 2, nbits_col, wthr4, gx, bad7);
        if (ROWS) gray_axis<CHECK_WHITE>(src, P, 2 + 2 * nbits_col, nbits_row, wthr4, gy, bad7);

        const size_t o = (size_t)view * P + (size_t)c * 16;
        uint32_t mk[4];
#pragma unroll
        for (int w = 0; w < 4; w++) {
            // pixels (0,1) and (2,3) of this word as 16-bit lanes: [lo0 hi0 lo1 hi1], [lo2 hi2 lo3 hi3]
            const uint32_t x01 = slr::gray_to_binary_x2(__byte_perm(gx.lo[w], gx.hi[w], 0x5140));
            const uint32_t x23 = slr::gray_to_binary_x2(__byte_perm(gx.lo[w], gx.hi[w], 0x7362));
            int xs[4] = {(int)(x01 & 0xffffu), (int)(x01 >> 16), (int)(x23 & 0xffffu), (int)(x23 >> 16)};
            int ys[4] = {0, 0, 0, 0};
            if (ROWS) {
                const uint32_t y01 = slr::gray_to_binary_x2(__byte_perm(gy.lo[w], gy.hi[w], 0x5140));
                const uint32_t y23 = slr::gray_to_binary_x2(__byte_perm(gy.lo[w], gy.hi[w], 0x7362));
                ys[0] = (int)(y01 & 0xffffu), ys[1] = (int)(y01 >> 16), ys[2] = (int)(y23 & 0xffffu), ys[3] = (int)(y23 >> 16);
            }
            uint32_t m4 = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                bool bad = ((bad7[w] >> (8 * i + 7)) & 1u) != 0 || xs[i] > scan_w;   // :403 / :365 (strict >)
                if (ROWS) bad = bad || ys[i] > scan_h;
                xs[i] = bad ? -1 : xs[i];
                ys[i] = bad ? -1 : ys[i];
                m4 |= (bad ? 0u : 1u) << (8 * i);
            }
            slr::stg_stream_u4(col + o + 4 * w, make_uint4(xs[0], xs[1], xs[2], xs[3]));
            if (ROWS) slr::stg_stream_u4(row + o + 4 * w, make_uint4(ys[0], ys[1], ys[2], ys[3]));
            mk[w] = m4;
        }
        slr::stg_stream_u4(mask + o, make_uint4(mk[0], mk[1], mk[2], mk[3]));
    }
}

}  // namespace

slr_status slr_launch_gray_decode(slr_engine *e, const uint8_t *d_stack, int views, int nbits_col, int nbits_row,
                                  int black_thr, int white_thr, int scan_w, int scan_h, int32_t *d_col,
                                  int32_t *d_row, uint8_t *d_mask)
{
    const size_t P = (size_t)e->W * e->H;
    const bool vec = (P % 16 == 0) &&
                     (((uintptr_t)d_stack | (uintptr_t)d_col | (uintptr_t)d_row | (uintptr_t)d_mask) % 16 == 0);
    const int px = vec ? 16 : 1;
    const long long cpv = (long long)(P / px);
    const long long total = cpv * views;
    long long blocks = (total + K2_THREADS - 1) / K2_THREADS;
    const long long cap = (long long)e->num_sms * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const bool rows = nbits_row > 0 && d_row != nullptr;
    const bool fast = vec && black_thr >= 0 && black_thr <= 255 && white_thr <= 256 && nbits_col >= 1 && nbits_col <= 16 &&
                      nbits_row <= 16 && (nbits_row == 0 || rows);
    if (fast) {
        void (*kern)(const uint8_t *, size_t, long long, long long, int, int, int, int, int, int, int32_t *, int32_t *,
                     uint8_t *);
        const bool cw = white_thr > 0;   // whiteThreshold <= 0 can never flag a pixel (|v1-v2| < 0 is false)
        kern = cw ? (rows ? k2_gray_decode_x4<true, true> : k2_gray_decode_x4<true, false>)
                  : (rows ? k2_gray_decode_x4<false, true> : k2_gray_decode_x4<false, false>);
        kern<<<(unsigned)blocks, K2_THREADS, 0, e->stream>>>(d_stack, P, cpv, total, nbits_col, nbits_row, black_thr,
                                                            white_thr, scan_w, scan_h, d_col, d_row, d_mask);
    } else if (vec)
        k2_gray_decode<16><<<(unsigned)blocks, K2_THREADS, 0, e->stream>>>(d_stack, P, cpv, total, nbits_col, nbits_row,
                                                                           black_thr, white_thr, scan_w, scan_h, d_col,
                                                                           d_row, d_mask);
    else
        k2_gray_decode<1><<<(unsigned)blocks, K2_THREADS, 0, e->stream>>>(d_stack, P, cpv, total, nbits_col, nbits_row,
                                                                          black_thr, white_thr, scan_w, scan_h, d_col,
                                                                          d_row, d_mask);
    SLR_CHECK_LAUNCH(e);
    return SLR_OK;
}
