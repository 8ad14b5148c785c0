// k_png.cu — the GPU half of the PNG ingest path and the PointCloudImage-layout epilogue.
//
// cv::imread(path, 0) (Duke/mfreconstruct.cpp:125, Duke/reconstruct.cpp:164) = container parse + inflate + PNG
// unfiltering + grey conversion.  Entropy decoding of a deflate stream is sequential and stays on the host cores, one
// stream per thread (facade/inflate.cpp); everything after it is data parallel and runs here, on the bytes exactly as
// the zlib stream holds them: H scanlines of [filter type][W filtered bytes] (8-bit grey), uploaded image by image
// while the other images are still being inflated.
//
//   k_png_rows   one warp per scanline: type 0 (None) copies, type 1 (Sub: x[i] += x[i-1] mod 256) is a byte-wise
//                prefix sum — 4 pixels per lane as SIMD-in-a-word adds, a warp scan of the word totals, ten words of
//                carry chain per 1280-pixel row — type 2 (Up) rows are copied and finished by
//   k_png_up     one thread per 4 columns walks down the image adding the row above to every type-2 row (launched only
//                when the image has such rows; OpenCV's encoder, which writes the reference's scan images, uses Sub
//                for every row).
// Rows of type 3 / 4 (Average / Paeth: a serial recurrence along the row) are left to the host decoder by the caller.
//
//   k_cloud_image  dense cloud [H][W] -> the storage of the reference's PointCloudImage (Duke/pointcloudimage.cpp:3-13):
//                sums float [scan_h][scan_w][3] + counts u8 [scan_h][scan_w], with MFReconstruct's addPoint(i, j, p)
//                addressing — cell (i_w = image row, j_h = image column), dropped when i_w >= scan_w or j_h >= scan_h
//                (Duke/mfreconstruct.cpp:326, Duke/pointcloudimage.cpp:86-97) — i.e. a clipped transpose; with a colour
//                plane (Reconstruct::triangulation_ge's haveColor, Duke/reconstruct.cpp:596-603) the grey value of every
//                cell goes along.
#include "slr_device.cuh"

namespace {

// per-byte a + b mod 256 of four packed bytes
__device__ __forceinline__ uint32_t add4(uint32_t a, uint32_t b)
{
    return ((a & 0x7f7f7f7fu) + (b & 0x7f7f7f7fu)) ^ ((a ^ b) & 0x80808080u);
}

__global__ void __launch_bounds__(256)
k_png_rows(const uint8_t *__restrict__ filt, uint8_t *__restrict__ dst, int W, int H)
{
    const int lane = threadIdx.x & 31;
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (y >= H) return;
    const uint8_t *src = filt + (size_t)y * (W + 1);
    const int type = src[0];
    src += 1;
    uint32_t *out = reinterpret_cast<uint32_t *>(dst + (size_t)y * W);   // W % 4 == 0, dst 4-byte aligned
    uint32_t carry = 0;                                                  // last reconstructed byte of the previous chunk
    for (int x0 = 0; x0 < W; x0 += 128) {
        const int x = x0 + 4 * lane;
        uint32_t w = 0;
        if (x < W) w = (uint32_t)src[x] | ((uint32_t)src[x + 1] << 8) | ((uint32_t)src[x + 2] << 16) | ((uint32_t)src[x + 3] << 24);
        if (type == 1) {
            w = add4(w, w << 8);            // prefix sums inside the word
            w = add4(w, w << 16);
            uint32_t tot = w >> 24;         // inclusive scan of the word totals across the warp (mod 256 at the end)
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, tot, o);
                if (lane >= o) tot += v;
            }
            const uint32_t before = (__shfl_up_sync(0xffffffffu, tot, 1) & (lane ? 0xffffffffu : 0u)) + carry;
            w = add4(w, (before & 0xffu) * 0x01010101u);
            carry = (__shfl_sync(0xffffffffu, tot, 31) + carry) & 0xffu;
        }
        if (x < W) out[x >> 2] = w;
    }
}

__global__ void __launch_bounds__(128)
k_png_up(const uint8_t *__restrict__ filt, uint8_t *__restrict__ dst, int W, int H)
{
    const int x = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (x >= W) return;
    uint32_t *col = reinterpret_cast<uint32_t *>(dst + x);
    const size_t pitch = (size_t)W >> 2;
    uint32_t above = col[0];                 // (a type-2 first row adds zeros)
    for (int y = 1; y < H; y++) {
        uint32_t w = col[(size_t)y * pitch];
        if (filt[(size_t)y * (W + 1)] == 2) {
            w = add4(w, above);
            col[(size_t)y * pitch] = w;
        }
        above = w;
    }
}

__global__ void __launch_bounds__(256)
k_cloud_image(const float *__restrict__ xyz, const uint8_t *__restrict__ valid, const uint8_t *__restrict__ gray, int W,
              int H, int scan_w, int scan_h, float *__restrict__ sum, uint8_t *__restrict__ cnt, uint8_t *__restrict__ cell_gray)
{
    // one 32 x 32 tile of the PointCloudImage per CTA, read transposed through shared memory
    __shared__ float t[3][32][33];
    __shared__ uint8_t tv[32][33], tg[32][33];
    const int iw0 = blockIdx.x * 32, jh0 = blockIdx.y * 32;   // cell (i_w, j_h) = image (row i_w, column j_h)
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 warps
    for (int r = ty; r < 32; r += 8) {                        // image row iw0 + r, columns jh0 .. jh0 + 31: coalesced
        const int i = iw0 + r, j = jh0 + tx;
        const bool in = i < H && j < W && i < scan_w && j < scan_h;
        uint8_t v = 0, g = 0;
        float px = 0.f, py = 0.f, pz = 0.f;
        if (in) {
            const size_t p = (size_t)i * W + j;
            v = valid[p];
            if (v) px = xyz[p * 3], py = xyz[p * 3 + 1], pz = xyz[p * 3 + 2];
            if (v && gray) g = gray[p];
        }
        tg[r][tx] = g;
        t[0][r][tx] = px;
        t[1][r][tx] = py;
        t[2][r][tx] = pz;
        tv[r][tx] = v ? 1 : 0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {                        // storage row j_h = jh0 + r, elements i_w = iw0 .. +31
        const int jh = jh0 + r, iw = iw0 + tx;
        if (jh < scan_h && iw < scan_w) {
            const size_t o = (size_t)jh * scan_w + iw;
            sum[o * 3] = t[0][tx][r];
            sum[o * 3 + 1] = t[1][tx][r];
            sum[o * 3 + 2] = t[2][tx][r];
            cnt[o] = tv[tx][r];
            if (cell_gray) cell_gray[o] = tg[tx][r];
        }
    }
}

}  // namespace

slr_status slr_launch_png_unfilter(slr_engine *e, cudaStream_t stream, const uint8_t *d_filtered, uint8_t *d_plane,
                                   bool has_up_rows)
{
    const int rows_per_cta = 8;
    k_png_rows<<<(e->H + rows_per_cta - 1) / rows_per_cta, rows_per_cta * 32, 0, stream>>>(d_filtered, d_plane, e->W, e->H);
    SLR_CHECK_LAUNCH(e);
    if (has_up_rows) {
        k_png_up<<<(e->W / 4 + 127) / 128, 128, 0, stream>>>(d_filtered, d_plane, e->W, e->H);
        SLR_CHECK_LAUNCH(e);
    }
    return SLR_OK;
}

slr_status slr_launch_cloud_image(slr_engine *e, const float *d_xyz, const uint8_t *d_valid, const uint8_t *d_gray, int scan_w,
                                  int scan_h, float *d_sum, uint8_t *d_cnt, uint8_t *d_cell_gray)
{
    dim3 grid((scan_w + 31) / 32, (scan_h + 31) / 32);
    k_cloud_image<<<grid, 256, 0, e->stream>>>(d_xyz, d_valid, d_gray, e->W, e->H, scan_w, scan_h, d_sum, d_cnt, d_cell_gray);
    SLR_CHECK_LAUNCH(e);
    return SLR_OK;
}
