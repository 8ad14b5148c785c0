// k_contrast.cu — Utilities::autoContrast (Duke/utilities.cpp:340-355) applied by Reconstruct::loadCamImgs to every
// loaded (and rectified) camera image when the "auto contrast" setting is on (Duke/reconstruct.cpp:182-183):
//     minMaxIdx(img, &min, &max);  min += 255*0.05;  a = 255/(max-min);  img -= min;  img *= a;
// on a CV_8U image, i.e. two saturating OpenCV 2.4 element-wise operations:
//     t   = saturate_cast<uchar>(cvRound(float(v) - float(min)))      (cv::subtract with a non-integer scalar works in CV_32F)
//     out = saturate_cast<uchar>(cvRound(float(t) * float(a)))        (Mat::operator*= -> convertTo -> cvtScale_<uchar, uchar, float>)
// Note: the reference splits the one-channel image into a vector of ONE Mat and then touches bgr[1], bgr[2]
// (undefined behaviour); what is restated here is the defined part, channel 0.  No reference pin is possible.
//
// Two passes per image: block-reduced min/max into one int2 per image, then a 256-entry look-up table built in
// shared memory from (min, max) and applied with 128-bit loads / stores (HBM bound: 1 read + 1 read + 1 write).
#include "slr_device.cuh"

namespace {

__global__ void k_contrast_minmax(const uint8_t *__restrict__ img, size_t P, int2 *__restrict__ mm)
{
    const uint8_t *src = img + (size_t)blockIdx.y * P;
    unsigned lo = 255, hi = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if ((((uintptr_t)src | P) & 15) == 0) {
        const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
        for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < P / 16; q += stride) {
            const uint4 v = slr::ldg_stream_u4(s4 + q);
            const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                lo = min(lo, min(min(w[k] & 255u, (w[k] >> 8) & 255u), min((w[k] >> 16) & 255u, w[k] >> 24)));
                hi = max(hi, max(max(w[k] & 255u, (w[k] >> 8) & 255u), max((w[k] >> 16) & 255u, w[k] >> 24)));
            }
        }
    } else {
        for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < P; q += stride) {
            lo = min(lo, (unsigned)src[q]);
            hi = max(hi, (unsigned)src[q]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&mm[blockIdx.y].x, (int)lo);
        atomicMax(&mm[blockIdx.y].y, (int)hi);
    }
}

// cv::saturate_cast<uchar>(cvRound(x)): round half to even, then clamp
__device__ __forceinline__ unsigned sat_round_u8(float x) { return (unsigned)min(max(__float2int_rn(x), 0), 255); }

__global__ void k_contrast_apply(uint8_t *__restrict__ img, size_t P, const int2 *__restrict__ mm)
{
    __shared__ uint8_t lut[256];
    uint8_t *dst = img + (size_t)blockIdx.y * P;
    {
        const int2 m = mm[blockIdx.y];
        const double mind = (double)m.x + 255 * 0.05;          // :348
        const float a = (float)(255.0 / ((double)m.y - mind));  // :350 (double), narrowed by convertTo
        const float minf = (float)mind;
        for (int v = threadIdx.x; v < 256; v += blockDim.x) {
            const unsigned t = sat_round_u8(__fsub_rn((float)v, minf));   // :351
            lut[v] = (uint8_t)sat_round_u8(__fmul_rn((float)t, a));       // :352
        }
    }
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if ((((uintptr_t)dst | P) & 15) == 0) {
        uint4 *d4 = reinterpret_cast<uint4 *>(dst);
        for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < P / 16; q += stride) {
            uint4 v = d4[q];
            unsigned *w = &v.x;
#pragma unroll
            for (int k = 0; k < 4; k++)
                w[k] = (unsigned)lut[w[k] & 255u] | ((unsigned)lut[(w[k] >> 8) & 255u] << 8) |
                       ((unsigned)lut[(w[k] >> 16) & 255u] << 16) | ((unsigned)lut[w[k] >> 24] << 24);
            d4[q] = v;
        }
    } else {
        for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < P; q += stride) dst[q] = lut[dst[q]];
    }
}

__global__ void k_contrast_init(int2 *mm, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mm[i] = make_int2(255, 0);
}

}  // namespace

// d_images = n_images images of W*H bytes each, stretched in place, every image with its own min / max
slr_status slr_launch_auto_contrast(slr_engine *e, uint8_t *d_images, int n_images)
{
    if (e->minmax_images < n_images) {
        cudaFree(e->d_minmax);
        e->d_minmax = nullptr;
        e->minmax_images = 0;
        SLR_CHECK_CUDA(cudaMalloc(&e->d_minmax, (size_t)n_images * sizeof(int2)));
        e->minmax_images = n_images;
    }
    const size_t P = (size_t)e->W * e->H;
    int2 *mm = reinterpret_cast<int2 *>(e->d_minmax);
    k_contrast_init<<<(n_images + 127) / 128, 128, 0, e->stream>>>(mm, n_images);
    SLR_CHECK_LAUNCH(e);
    int chunks = (int)((P / 16 + 255) / 256);
    const int cap = (8 * e->num_sms + n_images - 1) / n_images;   // ~8 CTAs per SM over all images
    if (chunks > cap) chunks = cap;
    if (chunks < 1) chunks = 1;
    const dim3 grid((unsigned)chunks, (unsigned)n_images);
    k_contrast_minmax<<<grid, 256, 0, e->stream>>>(d_images, P, mm);
    SLR_CHECK_LAUNCH(e);
    k_contrast_apply<<<grid, 256, 0, e->stream>>>(d_images, P, mm);
    SLR_CHECK_LAUNCH(e);
    return SLR_OK;
}
