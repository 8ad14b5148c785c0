// k3c_bucket.cu — K3c: Gray-only projector-cell bucket triangulation.
// Replaces the bucketing of Reconstruct::decodePaterns (Duke/reconstruct.cpp:56-74), Reconstruct::triangulation
// (:417-481), cam2WorldSpace (:310-322), Utilities::pixelToImageSpace / normalize / line_lineIntersection
// (Duke/utilities.cpp:47-56, 19-28, 399-425) and PointCloudImage::addPoint (Duke/pointcloudimage.cpp:86-97).
//
// Reference semantics: every camera pixel that decodes to projector cell (x, y) is pushed into that cell's
// list in camera column-major order; for every cell with both lists non-empty ALL (left, right) pixel pairs
// are intersected (ray-ray midpoint, rejected when |denom| < 0.1) and the midpoints are summed into the cell
// (float sum, u8 count that wraps and resets).  The summation order matters for the float sum, so it is kept.
//
// GPU plan per scan (a counting sort by cell, then one thread per projector cell):
//   kc_count    histogram of cells per camera (global atomics; the returned value is the pixel's slot in its slice)
//   kc_scan*    exclusive scan over the 2*ncell counters (block-local prefixes + scanned block totals; the users add
//               the two)
//   kc_scatter  pixel keys (x*H + y, i.e. column-major rank) into their cell's slice, any order
//   kc_rays     (once per calibration) the unit ray of every camera pixel: undistortPoints in fp64, image -> world,
//               normalise; 31 MB at 1280x1024, so that pairs cost loads instead of 300 fp64 instructions per ray
//   kc_cells    one thread per cell: order both slices by key (the reference's push order), walk the pairs in
//               the reference's (c1 outer, c2 inner) order with the exact fp32/fp64 operation sequence
// This path is latency bound (tiny irregular lists), not an HBM streaming kernel; it is not on the north-star
// bench.  Scans are processed one after another ("replicas only" across GPUs).
#include "slr_device.cuh"

namespace {

constexpr int KC_THREADS = 256;

struct BucketCalib {
    slr_camera cam[2];
    float pos[2][3];   // camera positions in world space: cam2WorldSpace((0,0,0)) (reconstruct.cpp:239-240)
    float rigid[12];
    int has_rigid;
};

// histogram of cells per camera; the value the atomic returns is the pixel's slot inside its cell's slice (any order
// will do: kc_cells sorts each slice), kept for kc_scatter so that the second pass needs no atomics
__global__ void kc_count(const int32_t *__restrict__ col, const int32_t *__restrict__ row,
                         const uint8_t *__restrict__ mask, int P, int scan_h, int ncell, int *__restrict__ count,
                         int *__restrict__ slot)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over 2*P (cam-major)
    if (idx >= 2 * P) return;
    int s = -1;
    if (mask[idx]) {
        const int cam = idx / P;
        const long long cell = (long long)col[idx] * scan_h + row[idx];  // ac(x, y) = x*scan_h + y (reconstruct.h:94-97)
        if (cell >= 0 && cell < ncell) s = atomicAdd(&count[cam * ncell + (int)cell], 1);  // out-of-table cells: dropped
    }
    slot[idx] = s;
}

// exclusive scan of n ints: block-local exclusive prefixes + per-block totals (scanned by kc_scan_partials)
__global__ void kc_scan_blocks(const int *__restrict__ in, int n, int *__restrict__ out, int *__restrict__ block_sums)
{
    __shared__ int s[KC_THREADS];
    const int base = blockIdx.x * KC_THREADS * 4;
    int v[4], sum = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int idx = base + threadIdx.x * 4 + i;
        v[i] = (idx < n) ? in[idx] : 0;
        sum += v[i];
    }
    s[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < KC_THREADS; o <<= 1) {
        const int t = (threadIdx.x >= o) ? s[threadIdx.x - o] : 0;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    int run = s[threadIdx.x] - sum;  // exclusive prefix of this thread within the block
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int idx = base + threadIdx.x * 4 + i;
        if (idx < n) out[idx] = run;
        run += v[i];
    }
    if (threadIdx.x == KC_THREADS - 1) block_sums[blockIdx.x] = s[threadIdx.x];
}

__global__ void kc_scan_partials(int *__restrict__ block_sums, int nblocks)
{
    // single block, sequential over chunks of KC_THREADS
    __shared__ int s[KC_THREADS];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += KC_THREADS) {
        const int idx = base + threadIdx.x;
        const int v = (idx < nblocks) ? block_sums[idx] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < KC_THREADS; o <<= 1) {
            const int t = (threadIdx.x >= o) ? s[threadIdx.x - o] : 0;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        if (idx < nblocks) block_sums[idx] = carry + s[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == KC_THREADS - 1) carry += s[threadIdx.x];
        __syncthreads();
    }
}

__global__ void kc_scatter(const int32_t *__restrict__ col, const int32_t *__restrict__ row,
                           const int *__restrict__ slot, int W, int H, int scan_h, int ncell,
                           const int *__restrict__ start, const int *__restrict__ block_sums, int *__restrict__ items)
{
    const int P = W * H;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 2 * P) return;
    const int sl = slot[idx];
    if (sl < 0) return;
    const int cam = idx / P, p = idx - cam * P;
    const long long cell = (long long)col[idx] * scan_h + row[idx];
    const int c = cam * ncell + (int)cell;
    const int pos = start[c] + block_sums[c / (KC_THREADS * 4)] + sl;   // block-local prefix + scanned block total
    const int x = p % W, y = p / W;
    items[pos] = x * H + y;  // column-major rank: the order decodePaterns visits camera pixels (:60-61)
}

// ---- exact reference arithmetic -----------------------------------------------------------------------------
__device__ __forceinline__ void undistort_d(float px, float py, const slr_camera &cam, float &ox, float &oy)
{
    // Utilities::undistortPoints, Duke/utilities.cpp:58-94 (same sequence as k_calib_synth.cu)
    const double k0 = cam.dist[0], k1 = cam.dist[1], k2 = cam.dist[2], k3 = cam.dist[3], k4 = 0.0;
    const double fx = cam.fc[0], fy = cam.fc[1];
    const double ifx = __ddiv_rn(1.0, fx), ify = __ddiv_rn(1.0, fy);
    const double cx = cam.cc[0], cy = cam.cc[1];
    double x = px, y = py;
    const double x0 = x = __dmul_rn(__dsub_rn(x, cx), ifx);
    const double y0 = y = __dmul_rn(__dsub_rn(y, cy), ify);
#pragma unroll 1
    for (int jj = 0; jj < 5; jj++) {
        const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
        double t = __dadd_rn(__dmul_rn(k4, r2), k1);
        t = __dadd_rn(__dmul_rn(t, r2), k0);
        t = __dmul_rn(t, r2);
        const double icdist = __ddiv_rn(1.0, __dadd_rn(1.0, t));
        const double dX = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, k2), x), y),
                                    __dmul_rn(k3, __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x))));
        const double dY = __dadd_rn(__dmul_rn(k2, __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y))),
                                    __dmul_rn(__dmul_rn(__dmul_rn(2.0, k3), x), y));
        x = __dmul_rn(__dsub_rn(x0, dX), icdist);
        y = __dmul_rn(__dsub_rn(y0, dY), icdist);
    }
    ox = __double2float_rn(__dadd_rn((double)__double2float_rn(__dmul_rn(x, fx)), cx));
    oy = __double2float_rn(__dadd_rn((double)__double2float_rn(__dmul_rn(y, fy)), cy));
}

// cam2WorldSpace (reconstruct.cpp:310-322): tmp = -R^T t, tmpPoint = R^T p (double accumulators, k ascending,
// narrowed to float: cv::Mat CV_32F products), p = tmp + tmpPoint in float
__device__ __forceinline__ void cam2world_d(const slr_camera &cam, float (&p)[3])
{
    float o[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double st = 0.0, sp = 0.0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            st = __dadd_rn(st, __dmul_rn((double)(-cam.R[k * 3 + i]), (double)cam.t[k]));
            sp = __dadd_rn(sp, __dmul_rn((double)cam.R[k * 3 + i], (double)p[k]));
        }
        o[i] = __fadd_rn(__double2float_rn(st), __double2float_rn(sp));
    }
    p[0] = o[0];
    p[1] = o[1];
    p[2] = o[2];
}

__device__ __forceinline__ float dot3(const float (&a)[3], const float (&b)[3])
{
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; i++) s = __fadd_rn(s, __fmul_rn(a[i], b[i]));
    return s;
}

// the unit ray of camera pixel (x, y): undistort -> image space -> world -> position - point -> normalize
__device__ __forceinline__ void pixel_ray(int x, int y, const slr_camera &cam, const float (&pos)[3], float (&ray)[3])
{
    float ux, uy;
    undistort_d((float)x, (float)y, cam, ux, uy);                                   // :441
    float pt[3] = {__fdiv_rn(__fsub_rn(ux, cam.cc[0]), cam.fc[0]),                  // utilities.cpp:47-56
                   __fdiv_rn(__fsub_rn(uy, cam.cc[1]), cam.fc[1]), 1.0f};
    cam2world_d(cam, pt);
    ray[0] = __fsub_rn(pos[0], pt[0]);                                              // :445
    ray[1] = __fsub_rn(pos[1], pt[1]);
    ray[2] = __fsub_rn(pos[2], pt[2]);
    // Utilities::normalize (utilities.cpp:19-28): float sum of squares, sqrt(float), max(1e-6, mag) in double
    const float ss = __fadd_rn(__fadd_rn(__fmul_rn(ray[0], ray[0]), __fmul_rn(ray[1], ray[1])), __fmul_rn(ray[2], ray[2]));
    const double mag = (double)__fsqrt_rn(ss);
    const float m = __double2float_rn((0.000001 > mag) ? 0.000001 : mag);
    ray[0] = __fdiv_rn(ray[0], m);
    ray[1] = __fdiv_rn(ray[1], m);
    ray[2] = __fdiv_rn(ray[2], m);
}

// unit rays of every pixel of both cameras, once per calibration: [2][H*W][3]
__global__ void kc_rays(BucketCalib cal, int W, int H, float *__restrict__ rays)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // over 2*P
    const int P = W * H;
    if (idx >= 2 * P) return;
    const int cam = idx / P, p = idx - cam * P;
    float ray[3];
    pixel_ray(p % W, p / W, cal.cam[cam], cal.pos[cam], ray);
    rays[(size_t)idx * 3 + 0] = ray[0];
    rays[(size_t)idx * 3 + 1] = ray[1];
    rays[(size_t)idx * 3 + 2] = ray[2];
}

__device__ __forceinline__ void sort_slice(int *a, int n)
{
    for (int i = 1; i < n; i++) {  // insertion sort: slices hold a handful of pixels
        const int v = a[i];
        int j = i - 1;
        while (j >= 0 && a[j] > v) {
            a[j + 1] = a[j];
            j--;
        }
        a[j + 1] = v;
    }
}

__global__ void kc_cells(const int *__restrict__ start, const int *__restrict__ block_sums, const int *__restrict__ count,
                         int *__restrict__ items, int W, int H,
                         int ncell, BucketCalib cal, const float *__restrict__ rays, float *__restrict__ sum,
                         uint8_t *__restrict__ cnt, unsigned long long *__restrict__ n_cells)
{
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned filled = 0;
    if (cell < ncell) {
        const int n1 = count[cell], n2 = count[ncell + cell];
        float acc[3] = {0.0f, 0.0f, 0.0f};
        unsigned char num = 0;
        if (n1 > 0 && n2 > 0) {                                            // :436
            int *l1 = items + start[cell] + block_sums[cell / (KC_THREADS * 4)];
            int *l2 = items + start[ncell + cell] + block_sums[(ncell + cell) / (KC_THREADS * 4)];
            sort_slice(l1, n1);
            sort_slice(l2, n2);
            for (int c1 = 0; c1 < n1; c1++) {
                // the pixel's unit ray (pixel_ray: undistort -> image space -> world -> normalise, :440-447) comes from
                // the per-calibration table; items hold the column-major rank x*H + y
                const size_t p1 = (size_t)(l1[c1] % H) * W + l1[c1] / H;
                const float ray1[3] = {__ldg(rays + p1 * 3), __ldg(rays + p1 * 3 + 1), __ldg(rays + p1 * 3 + 2)};
                for (int c2 = 0; c2 < n2; c2++) {
                    const size_t p2 = (size_t)W * H + (size_t)(l2[c2] % H) * W + l2[c2] / H;
                    const float ray2[3] = {__ldg(rays + p2 * 3), __ldg(rays + p2 * 3 + 1), __ldg(rays + p2 * 3 + 2)};
                    // Utilities::line_lineIntersection, utilities.cpp:399-425 (all fp32)
                    const float v12[3] = {__fsub_rn(cal.pos[0][0], cal.pos[1][0]), __fsub_rn(cal.pos[0][1], cal.pos[1][1]),
                                          __fsub_rn(cal.pos[0][2], cal.pos[1][2])};
                    const float d11 = dot3(ray1, ray1), d22 = dot3(ray2, ray2), d12 = dot3(ray1, ray2);
                    const float e1 = dot3(v12, ray1), e2 = dot3(v12, ray2);
                    const float denom = __fsub_rn(__fmul_rn(d11, d22), __fmul_rn(d12, d12));
                    if (fabsf(denom) < 0.1f) continue;                     // abs(denom) < 0.1 (see phase_match on 0.1 vs 0.1f)
                    const float s = __fsub_rn(__fmul_rn(__fdiv_rn(d12, denom), e2), __fmul_rn(__fdiv_rn(d22, denom), e1));
                    const float t = __fadd_rn(__fmul_rn(-__fdiv_rn(d12, denom), e1), __fmul_rn(__fdiv_rn(d11, denom), e2));
                    float ip[3];
#pragma unroll
                    for (int i = 0; i < 3; i++) {
                        const float a = __fadd_rn(cal.pos[0][i], __fmul_rn(s, ray1[i]));
                        const float b = __fadd_rn(cal.pos[1][i], __fmul_rn(t, ray2[i]));
                        ip[i] = __double2float_rn(__dmul_rn(0.5, (double)__fadd_rn(a, b)));
                    }
                    if (cal.has_rigid) {                                   // :466-473
                        float o[3];
#pragma unroll
                        for (int i = 0; i < 3; i++) {
                            double q = __dmul_rn((double)cal.rigid[4 * i + 0], (double)ip[0]);
                            q = __dadd_rn(q, __dmul_rn((double)cal.rigid[4 * i + 1], (double)ip[1]));
                            q = __dadd_rn(q, __dmul_rn((double)cal.rigid[4 * i + 2], (double)ip[2]));
                            q = __dadd_rn(q, (double)cal.rigid[4 * i + 3]);
                            o[i] = __double2float_rn(q);
                        }
                        ip[0] = o[0];
                        ip[1] = o[1];
                        ip[2] = o[2];
                    }
                    // PointCloudImage::addPoint (pointcloudimage.cpp:86-97): u8 count wraps, a zero count resets the sum
                    if (num == 0) {
                        acc[0] = ip[0];
                        acc[1] = ip[1];
                        acc[2] = ip[2];
                        num = 1;
                    } else {
                        acc[0] = __fadd_rn(ip[0], acc[0]);
                        acc[1] = __fadd_rn(ip[1], acc[1]);
                        acc[2] = __fadd_rn(ip[2], acc[2]);
                        num = (unsigned char)(num + 1);
                    }
                }
            }
        }
        // (a count that wrapped to 0 leaves the stale sum in place, as in the reference; readers use count > 0)
        sum[(size_t)cell * 3 + 0] = acc[0];
        sum[(size_t)cell * 3 + 1] = acc[1];
        sum[(size_t)cell * 3 + 2] = acc[2];
        cnt[cell] = num;
        filled = num ? 1u : 0u;
    }
    if (n_cells) {
        const unsigned long long s = slr::warp_sum_u32(filled);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(n_cells, s);
    }
}

}  // namespace

slr_status slr_launch_bucket_triangulate(slr_engine *e, const int32_t *d_col, const int32_t *d_row,
                                         const uint8_t *d_mask, int batch, int scan_w, int scan_h, float *d_sum,
                                         uint8_t *d_cnt, unsigned long long *d_n_cells)
{
    const int W = e->W, H = e->H;
    const long long P = (long long)W * H;
    const long long ncell_ll = (long long)scan_w * scan_h;
    SLR_REQUIRE(2 * P < (1LL << 31) && 2 * ncell_ll < (1LL << 31), "slr_bucket_triangulate: image or scan area too large");
    const int ncell = (int)ncell_ll;
    const int n = 2 * ncell;
    const int nblocks = (n + KC_THREADS * 4 - 1) / (KC_THREADS * 4);

    // scratch for one scan (sized on first use / growth)
    const size_t need = ((size_t)2 * n + nblocks + 4 * (size_t)P + 16) * sizeof(int);
    if (e->bucket_scratch_bytes < need) {
        if (e->d_bucket_scratch) SLR_CHECK_CUDA(cudaFree(e->d_bucket_scratch));
        e->d_bucket_scratch = nullptr;
        e->bucket_scratch_bytes = 0;
        SLR_CHECK_CUDA(cudaMalloc(&e->d_bucket_scratch, need));
        e->bucket_scratch_bytes = need;
    }
    int *count = (int *)e->d_bucket_scratch, *start = count + n, *bsum = start + n;
    int *items = bsum + nblocks, *slot = items + 2 * P;

    BucketCalib cal;
    cal.cam[0] = e->cams[0];
    cal.cam[1] = e->cams[1];
    for (int c = 0; c < 2; c++)
        for (int i = 0; i < 3; i++) {  // cam2WorldSpace((0,0,0)): (float)(-R^T t) + (float)0
            double st = 0.0;
            for (int k = 0; k < 3; k++) st += (double)(-e->cams[c].R[k * 3 + i]) * (double)e->cams[c].t[k];
            cal.pos[c][i] = (float)st + 0.0f;
        }
    memcpy(cal.rigid, e->calib.rigid, sizeof(cal.rigid));
    cal.has_rigid = e->calib.has_rigid;

    // per-pixel unit rays: a function of the calibration only (fp64 undistortion, 5 iterations per pixel)
    if (!e->d_rays || e->rays_version != e->calib_version) {
        if (!e->d_rays) SLR_CHECK_CUDA(cudaMalloc(&e->d_rays, (size_t)2 * P * 3 * sizeof(float)));
        kc_rays<<<(int)((2 * P + KC_THREADS - 1) / KC_THREADS), KC_THREADS, 0, e->stream>>>(cal, W, H, e->d_rays);
        SLR_CHECK_LAUNCH(e);
        e->rays_version = e->calib_version;
    }

    for (int b = 0; b < batch; b++) {
        const int32_t *col = d_col + (size_t)b * 2 * P, *row = d_row + (size_t)b * 2 * P;
        const uint8_t *mask = d_mask + (size_t)b * 2 * P;
        SLR_CHECK_CUDA(cudaMemsetAsync(count, 0, (size_t)n * sizeof(int), e->stream));
        const int pix_blocks = (int)((2 * P + KC_THREADS - 1) / KC_THREADS);
        kc_count<<<pix_blocks, KC_THREADS, 0, e->stream>>>(col, row, mask, (int)P, scan_h, ncell, count, slot);
        SLR_CHECK_LAUNCH(e);
        kc_scan_blocks<<<nblocks, KC_THREADS, 0, e->stream>>>(count, n, start, bsum);
        SLR_CHECK_LAUNCH(e);
        kc_scan_partials<<<1, KC_THREADS, 0, e->stream>>>(bsum, nblocks);
        SLR_CHECK_LAUNCH(e);
        kc_scatter<<<pix_blocks, KC_THREADS, 0, e->stream>>>(col, row, slot, W, H, scan_h, ncell, start, bsum, items);
        SLR_CHECK_LAUNCH(e);
        kc_cells<<<(ncell + 127) / 128, 128, 0, e->stream>>>(start, bsum, count, items, W, H, ncell, cal, e->d_rays,
                                                             d_sum + (size_t)b * ncell * 3, d_cnt + (size_t)b * ncell,
                                                             d_n_cells);
        SLR_CHECK_LAUNCH(e);
    }
    return SLR_OK;
}
