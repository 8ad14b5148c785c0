// k_merge.cu — SURVEY.md 8f row N2: rigid alignment of consecutive scans and the multi-scan merge.
//
// Reference: DotMatch::calMatrix (Duke/dotmatch.cpp:1180-1324) turns matched marker positions of two scans into a rigid
// transform with Horn's closed-form quaternion method (it calls mrpt::scanmatching::HornMethod of MRPT 1.2.2 — a third-
// party library that is NOT under /root/reference; restated here from Horn, "Closed-form solution of absolute
// orientation using unit quaternions", JOSA A 4(4), 1987, with MRPT's conventions: the first point of a pair is the
// base, the second is moved onto it; output {tx ty tz qr qx qy qz}; the translation carries the estimated scale unless
// unit scale is forced.  Parity of this helper is unpinned: no MRPT here and no reference fixture; the tests check it
// against an SVD (Kabsch) solution), converts the quaternion to a matrix (:1241-1250), chains it onto the transforms
// of the earlier scans (:1312-1317: R = R_prev R_n, T = R_prev T_n + T_prev) and writes scan/transfer_mat<sn>.txt, which
// MFReconstruct::triangulation / Reconstruct::triangulation_ge apply to every point (Duke/mfreconstruct.cpp:315-323).
//
//   slr_register_scan   the host part: marker pairs (+ the previous scan's accumulated transform) -> 3x4 transfer matrix
//   slr_merge_scans     the device part: the clouds of n scans, each with its transfer matrix, -> ONE compacted point
//                       list in the common frame, scan by scan and pixel by pixel (the order in which the reference's
//                       per-scan exports list them), with the source pixel of every point.  Three launches: valid
//                       counts per 1024-pixel tile, exclusive scan of the tile totals, ordered scatter with the 3x4
//                       product evaluated exactly as the per-scan epilogue does (double products of float operands,
//                       one rounding).
#include <math.h>
#include <string.h>

#include "slr_device.cuh"

namespace {

constexpr int TILE = 1024;

__global__ void __launch_bounds__(TILE)
k_merge_count(const uint8_t *__restrict__ valid, size_t total, unsigned *__restrict__ tiles)
{
    const size_t p = (size_t)blockIdx.x * TILE + threadIdx.x;
    const int n = __syncthreads_count(p < total && valid[p] != 0);
    if (threadIdx.x == 0) tiles[blockIdx.x] = (unsigned)n;
}

// exclusive scan of the tile totals in place (64-bit running sum kept as two outputs: offsets fit 2^32 points per call,
// checked by the launcher), total to *count
__global__ void __launch_bounds__(1024)
k_merge_scan(unsigned *__restrict__ tiles, int n_tiles, unsigned long long *__restrict__ count)
{
    __shared__ unsigned warp_sums[32];
    __shared__ unsigned carry_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned v = i < n_tiles ? tiles[i] : 0u;
        unsigned inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_sums[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            unsigned w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_sums[lane] = w;   // inclusive over warps
        }
        __syncthreads();
        const unsigned before = carry_s + (wid ? warp_sums[wid - 1] : 0u) + inc - v;
        if (i < n_tiles) tiles[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = carry_s;
}

struct MergeXf {
    float m[12];
    int apply;
};

__global__ void __launch_bounds__(TILE)
k_merge_emit(const float *__restrict__ xyz, const uint8_t *__restrict__ valid, size_t total, size_t px_per_scan,
             const MergeXf *__restrict__ xf, const unsigned *__restrict__ tiles, float *__restrict__ points,
             long long *__restrict__ source)
{
    __shared__ unsigned warp_sums[32];
    const size_t p = (size_t)blockIdx.x * TILE + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool v = p < total && valid[p] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, v);
    if (lane == 0) warp_sums[wid] = __popc(bal);
    __syncthreads();
    if (wid == 0) {
        unsigned w = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        warp_sums[lane] = w;
    }
    __syncthreads();
    if (!v) return;
    const size_t pos = (size_t)tiles[blockIdx.x] + (wid ? warp_sums[wid - 1] : 0u) + __popc(bal & ((1u << lane) - 1u));
    float x = xyz[p * 3], y = xyz[p * 3 + 1], z = xyz[p * 3 + 2];
    const MergeXf t = xf[p / px_per_scan];
    if (t.apply) {   // XYZ <- M(3x4, f32) * [XYZ, 1] as cv::Mat's product evaluates it (Duke/mfreconstruct.cpp:315-323)
        float o[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            double s = __dmul_rn((double)t.m[4 * i + 0], (double)x);
            s = __dadd_rn(s, __dmul_rn((double)t.m[4 * i + 1], (double)y));
            s = __dadd_rn(s, __dmul_rn((double)t.m[4 * i + 2], (double)z));
            s = __dadd_rn(s, (double)t.m[4 * i + 3]);
            o[i] = __double2float_rn(s);
        }
        x = o[0], y = o[1], z = o[2];
    }
    points[pos * 3] = x;
    points[pos * 3 + 1] = y;
    points[pos * 3 + 2] = z;
    if (source) source[pos] = (long long)p;
}

// ---- Horn's method on the host --------------------------------------------------------------------------------------
// largest eigenvector of a symmetric 4x4 matrix: cyclic Jacobi rotations
void sym4_largest_eigenvector(double a[4][4], double v_out[4])
{
    double v[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    for (int sweep = 0; sweep < 64; sweep++) {
        double off = 0;
        for (int p = 0; p < 4; p++)
            for (int q = p + 1; q < 4; q++) off += a[p][q] * a[p][q];
        if (off < 1e-300) break;
        for (int p = 0; p < 4; p++)
            for (int q = p + 1; q < 4; q++) {
                if (fabs(a[p][q]) < 1e-300) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 4; k++) {
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 4; k++) {
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 4; k++) {
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - s * vkq;
                    v[k][q] = s * vkp + c * vkq;
                }
            }
    }
    int best = 0;
    for (int k = 1; k < 4; k++)
        if (a[k][k] > a[best][best]) best = k;
    for (int k = 0; k < 4; k++) v_out[k] = v[k][best];
}

}  // namespace

extern "C" slr_status slr_horn_method(const double *pairs, int n, int force_unit_scale, double out7[7], double *scale_out)
{
    SLR_REQUIRE(pairs && out7 && n >= 3, "slr_horn_method: needs at least 3 point pairs");
    double cb[3] = {0, 0, 0}, cm[3] = {0, 0, 0};
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) cb[k] += pairs[6 * i + k], cm[k] += pairs[6 * i + 3 + k];
    for (int k = 0; k < 3; k++) cb[k] /= n, cm[k] /= n;
    double S[3][3] = {{0}}, nb = 0, nm = 0;   // S = sum m' b'^T (moving x base)
    for (int i = 0; i < n; i++) {
        double b[3], m[3];
        for (int k = 0; k < 3; k++) b[k] = pairs[6 * i + k] - cb[k], m[k] = pairs[6 * i + 3 + k] - cm[k];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) S[r][c] += m[r] * b[c];
        nb += b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
        nm += m[0] * m[0] + m[1] * m[1] + m[2] * m[2];
    }
    SLR_REQUIRE(nm > 0 && nb > 0, "slr_horn_method: degenerate point set");
    double N[4][4];
    N[0][0] = S[0][0] + S[1][1] + S[2][2];
    N[0][1] = S[1][2] - S[2][1];
    N[0][2] = S[2][0] - S[0][2];
    N[0][3] = S[0][1] - S[1][0];
    N[1][1] = S[0][0] - S[1][1] - S[2][2];
    N[1][2] = S[0][1] + S[1][0];
    N[1][3] = S[2][0] + S[0][2];
    N[2][2] = -S[0][0] + S[1][1] - S[2][2];
    N[2][3] = S[1][2] + S[2][1];
    N[3][3] = -S[0][0] - S[1][1] + S[2][2];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < r; c++) N[r][c] = N[c][r];
    double q[4];
    sym4_largest_eigenvector(N, q);
    double qn = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (q[0] < 0) qn = -qn;   // qr >= 0
    for (int k = 0; k < 4; k++) q[k] /= qn;
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double R[3][3] = {{w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)},
                            {2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)},
                            {2 * (z * x - w * y), 2 * (z * y + w * x), w * w - x * x - y * y + z * z}};
    const double s = force_unit_scale ? 1.0 : sqrt(nb / nm);
    for (int r = 0; r < 3; r++) out7[r] = cb[r] - s * (R[r][0] * cm[0] + R[r][1] * cm[1] + R[r][2] * cm[2]);
    for (int k = 0; k < 4; k++) out7[3 + k] = q[k];
    if (scale_out) *scale_out = sqrt(nb / nm);
    return SLR_OK;
}

extern "C" slr_status slr_register_scan(const double *pairs, int n, const double *prev3x4, double out3x4[12])
{
    SLR_REQUIRE(out3x4 != nullptr, "slr_register_scan: out3x4 is NULL");
    double h[7];
    const slr_status st = slr_horn_method(pairs, n, 0, h, nullptr);   // the reference calls HornMethod with its defaults
    if (st != SLR_OK) return st;
    const double tx = h[0], ty = h[1], tz = h[2], w = h[3], i = h[4], j = h[5], k = h[6];
    // Duke/dotmatch.cpp:1241-1250
    const double cur[12] = {pow(w, 2) + pow(i, 2) - pow(j, 2) - pow(k, 2), 2 * (i * j - w * k), 2 * (i * k + w * j), tx,
                            2 * (i * j + w * k), pow(w, 2) - pow(i, 2) + pow(j, 2) - pow(k, 2), 2 * (j * k - w * i), ty,
                            2 * (k * i - w * j), 2 * (k * j + w * i), pow(w, 2) - pow(i, 2) - pow(j, 2) + pow(k, 2), tz};
    if (!prev3x4) {   // scanSN == 1: the matrix itself (:1299-1310)
        memcpy(out3x4, cur, sizeof(cur));
        return SLR_OK;
    }
    // :1312-1317  R = R_prev R_n, T = R_prev T_n + T_prev (cv::Mat double products: sums left to right)
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++)
            out3x4[4 * r + c] = prev3x4[4 * r + 0] * cur[c] + prev3x4[4 * r + 1] * cur[4 + c] + prev3x4[4 * r + 2] * cur[8 + c];
        out3x4[4 * r + 3] = (prev3x4[4 * r + 0] * cur[3] + prev3x4[4 * r + 1] * cur[7] + prev3x4[4 * r + 2] * cur[11]) + prev3x4[4 * r + 3];
    }
    return SLR_OK;
}

slr_status slr_launch_merge(slr_engine *e, const float *d_xyz, const uint8_t *d_valid, int n_scans, const float *h_rigid,
                            const uint8_t *h_has_rigid, float *d_points, long long *d_source, unsigned long long *d_count)
{
    const size_t P = (size_t)e->W * e->H, total = P * (size_t)n_scans;
    SLR_REQUIRE(total < (1ull << 32), "slr_merge_scans: more than 2^32 pixels in one call");
    const int n_tiles = (int)((total + TILE - 1) / TILE);
    const size_t need = (size_t)n_tiles * sizeof(unsigned) + (size_t)n_scans * sizeof(MergeXf) + 256;
    if (e->merge_bytes < need) {
        if (e->d_merge) SLR_CHECK_CUDA(cudaFree(e->d_merge));
        e->d_merge = nullptr;
        e->merge_bytes = 0;
        SLR_CHECK_CUDA(cudaMalloc(&e->d_merge, need));
        e->merge_bytes = need;
    }
    unsigned *tiles = reinterpret_cast<unsigned *>(e->d_merge);
    MergeXf *xf = reinterpret_cast<MergeXf *>(reinterpret_cast<unsigned char *>(e->d_merge) + (((size_t)n_tiles * sizeof(unsigned) + 255) & ~(size_t)255));
    std::string host((size_t)n_scans * sizeof(MergeXf), '\0');
    MergeXf *hx = reinterpret_cast<MergeXf *>(&host[0]);
    for (int s = 0; s < n_scans; s++) {
        hx[s].apply = (h_rigid && (!h_has_rigid || h_has_rigid[s])) ? 1 : 0;
        if (hx[s].apply) memcpy(hx[s].m, h_rigid + 12 * (size_t)s, sizeof(hx[s].m));
    }
    SLR_CHECK_CUDA(cudaMemcpyAsync(xf, hx, host.size(), cudaMemcpyHostToDevice, e->stream));
    SLR_CHECK_CUDA(cudaStreamSynchronize(e->stream));   // `host` goes out of scope
    k_merge_count<<<n_tiles, TILE, 0, e->stream>>>(d_valid, total, tiles);
    SLR_CHECK_LAUNCH(e);
    k_merge_scan<<<1, 1024, 0, e->stream>>>(tiles, n_tiles, d_count);
    SLR_CHECK_LAUNCH(e);
    k_merge_emit<<<n_tiles, TILE, 0, e->stream>>>(d_xyz, d_valid, total, P, xf, tiles, d_points, d_source);
    SLR_CHECK_LAUNCH(e);
    return SLR_OK;
}
