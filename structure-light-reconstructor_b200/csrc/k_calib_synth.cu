// k_calib_synth.cu — (1) the per-pixel Utilities::undistortPoints maps computed once per calibration,
// (2) on-device synthetic scan renderers used by bench.py / smoke.
#include "slr_device.cuh"

namespace {

// Utilities::undistortPoints (Duke/utilities.cpp:58-94) in IEEE double, operation for operation
// (explicit _rn intrinsics: no FMA contraction), so the map equals the host evaluation bit for bit.
__device__ __forceinline__ void undistort_point(float px, float py, const slr_camera &cam, float &ox, float &oy)
{
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
        // deltaX = 2*k[2]*x*y + k[3]*(r2 + 2*x*x);  deltaY = k[2]*(r2 + 2*y*y) + 2*k[3]*x*y
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

__global__ void k_undistort_maps(slr_camera camL, slr_camera camR, int W, int H, int row0, float *__restrict__ lx,
                                 float *__restrict__ ly, float *__restrict__ rx)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= W || i >= H) return;
    float ax, ay, bx, by;
    undistort_point((float)j, (float)(i + row0), camL, ax, ay);   // row0: first image row of a row band
    undistort_point((float)j, (float)(i + row0), camR, bx, by);
    const size_t o = (size_t)i * W + j;
    lx[o] = ax;
    ly[o] = ay;
    rx[o] = bx;
}

// ------------------------------------------------------------------------------------------------
// synthetic scenes
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash_u32(uint32_t x)
{
    x ^= x >> 16;
    x *= 0x7feb352dU;
    x ^= x >> 15;
    x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}
__device__ __forceinline__ float hash_unit(uint32_t x) { return (hash_u32(x) >> 8) * (1.0f / 16777216.0f); }

struct Scene {
    float d0, d1, amp, cx, cy, inv_s2;  // disparity field
    float g_scale, g_off;               // right-view projector coordinate g(k,i) = g_scale*k + g_off + 0.02*i
    int sh_x0[3], sh_x1[3], sh_y0[3], sh_y1[3];
};

__device__ Scene make_scene(unsigned seed, int b, int W, int H, int proj_w)
{
    Scene s;
    const uint32_t h = hash_u32(seed * 0x9E3779B9u + (uint32_t)b * 0x85EBCA6Bu + 1u);
    s.d0 = 24.0f + 16.0f * hash_unit(h + 1);
    s.d1 = 8.0f * (hash_unit(h + 2) - 0.5f) / (float)H;
    s.amp = 6.0f + 10.0f * hash_unit(h + 3);
    s.cx = W * (0.35f + 0.3f * hash_unit(h + 4));
    s.cy = H * (0.35f + 0.3f * hash_unit(h + 5));
    const float sig = 0.18f * (float)W;
    s.inv_s2 = 1.0f / (sig * sig);
    s.g_scale = 0.92f * (float)proj_w / (float)W;
    s.g_off = 0.02f * (float)proj_w;
    for (int q = 0; q < 3; q++) {
        const int x0 = (int)(hash_unit(h + 10 + q) * 0.85f * W), y0 = (int)(hash_unit(h + 20 + q) * 0.85f * H);
        s.sh_x0[q] = x0;
        s.sh_x1[q] = x0 + (int)(0.06f * W + hash_unit(h + 30 + q) * 0.08f * W);
        s.sh_y0[q] = y0;
        s.sh_y1[q] = y0 + (int)(0.06f * H + hash_unit(h + 40 + q) * 0.08f * H);
    }
    return s;
}

// projector coordinate seen by pixel (x, i) of camera `cam`, and whether it is lit
__device__ __forceinline__ bool scene_coord(const Scene &s, int cam, int x, int i, int proj_w, int integer_disp,
                                            float &u)
{
    float k = (float)x;
    if (cam == 0) {
        const float dx = (float)x - s.cx, dy = (float)i - s.cy;
        float d = s.d0 + s.d1 * (float)i + s.amp * __expf(-(dx * dx + dy * dy) * s.inv_s2);
        if (integer_disp) d = rintf(d);
        k = (float)x - d;
    }
    u = s.g_scale * k + s.g_off + 0.02f * (float)i;
    bool lit = (u >= 0.0f) && (u < (float)proj_w) && (k >= 0.0f);
    for (int q = 0; q < 3; q++)  // shadow rectangles differ per camera (occlusion-like)
        lit = lit && !(x >= s.sh_x0[q] + 37 * cam && x < s.sh_x1[q] + 37 * cam && i >= s.sh_y0[q] && i < s.sh_y1[q]);
    return lit;
}

__device__ __forceinline__ uint8_t add_noise(float v, float noise_dn, uint32_t key)
{
    if (noise_dn > 0.0f) {
        // sum of two uniforms (triangular), scaled to roughly +-2*noise_dn
        const float n = (hash_unit(key) + hash_unit(key ^ 0x68bc21ebu) - 1.0f) * 2.0f * noise_dn;
        v += n;
    }
    v = fminf(fmaxf(v, 0.0f), 255.0f);
    return (uint8_t)v;
}

// one thread per (row, col) of one view; writes all N planes
__global__ void k_synth_mf(uint8_t *__restrict__ stack, int W, int H, int batch, int F, int S, int proj_w, unsigned seed,
                           int integer_disp, float noise_dn)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    const int view = blockIdx.z;  // b*2 + cam
    if (x >= W) return;
    const int b = view >> 1, cam = view & 1;
    const Scene s = make_scene(seed, b, W, H, proj_w);
    float u;
    const bool lit = scene_coord(s, cam, x, i, proj_w, integer_disp, u);
    const size_t P = (size_t)W * H;
    uint8_t *base = stack + (size_t)view * (2 + F * S) * P + (size_t)i * W + x;
    // 16 noise keys per pixel for the reference's 14-image stack (unchanged streams), 2 + F*S rounded up otherwise
    const uint32_t keys_per_px = (F * S <= 14) ? 16u : (uint32_t)((2 + F * S + 15) & ~15);
    const uint32_t key0 = hash_u32(seed ^ (uint32_t)(view * 0x01000193u)) + (uint32_t)(i * W + x) * keys_per_px;
    base[0] = add_noise(lit ? 200.0f : 30.0f, noise_dn, key0);
    base[P] = add_noise(20.0f, noise_dn, key0 + 1);
    // first three: Duke/multifrequency.cpp:3; the fourth keeps the last level of a 4-frequency cascade from beating at
    // frequency 0 ((70-64)-(64-59) = 1, (64-59)-(59-56) = 2; 55 would give 1 and 1: a constant phase along every row)
    const int freq[8] = {70, 64, 59, 56, 52, 50, 47, 45};
    for (int f = 0; f < F; f++)
        for (int sft = 0; sft < S; sft++) {
            // same form as the reference's generator (Duke/multifrequency.cpp:27) at a fractional column u; the shift
            // PI*2*s/S equals the reference's PI*s/2 for S = 4 bit for bit
            const double arg = 3.1416 * 2 * (double)u * (double)freq[f] / (double)proj_w + 3.1416 * 2 * (double)sft / S;
            float v = 135.0f + 79.0f * cosf((float)arg);
            v = truncf(v);
            if (!lit) v = 20.0f;
            base[(size_t)(2 + S * f + sft) * P] = add_noise(v, noise_dn, key0 + 2 + S * f + sft);
        }
}

__global__ void k_synth_gray(uint8_t *__restrict__ stack, int W, int H, int batch, int scan_w, int nbits,
                             unsigned seed, int integer_disp, float noise_dn)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    const int view = blockIdx.z;
    if (x >= W) return;
    const int b = view >> 1, cam = view & 1;
    const Scene s = make_scene(seed, b, W, H, scan_w);
    float u;
    const bool lit = scene_coord(s, cam, x, i, scan_w, integer_disp, u);
    const size_t P = (size_t)W * H;
    const int N = 2 + 2 * nbits;
    uint8_t *base = stack + (size_t)view * N * P + (size_t)i * W + x;
    const uint32_t key0 = hash_u32(seed ^ (uint32_t)(view * 0x01000193u)) + (uint32_t)(i * W + x) * 64u;
    base[0] = add_noise(lit ? 200.0f : 30.0f, noise_dn, key0);
    base[P] = add_noise(20.0f, noise_dn, key0 + 1);
    const int col = lit ? (int)u : 0;
    const int gray = col ^ (col >> 1);
    for (int c = 0; c < nbits; c++) {  // image 2+2c holds Gray bit nbits-1-c (MSB first), 3+2c its inverse
        const int bit = (gray >> (nbits - 1 - c)) & 1;
        const float on = lit ? 200.0f : 20.0f, off = 20.0f;
        base[(size_t)(2 + 2 * c) * P] = add_noise(bit ? on : off, noise_dn, key0 + 2 + 2 * c);
        base[(size_t)(3 + 2 * c) * P] = add_noise(bit ? off : on, noise_dn, key0 + 3 + 2 * c);
    }
}

}  // namespace

slr_status slr_launch_undistort_maps(slr_engine *e)
{
    dim3 block(128), grid((e->W + 127) / 128, e->H);
    k_undistort_maps<<<grid, block, 0, e->stream>>>(e->cams[0], e->cams[1], e->W, e->H, e->calib.row0, e->d_undist_lx,
                                                    e->d_undist_ly, e->d_undist_rx);
    SLR_CHECK_LAUNCH(e);
    return SLR_OK;
}

slr_status slr_launch_synth_mf(slr_engine *e, uint8_t *d_stack, int batch, int F, int S, int proj_w, unsigned seed,
                               int integer_disparity, float noise_dn)
{
    dim3 block(128), grid((e->W + 127) / 128, e->H, batch * 2);
    k_synth_mf<<<grid, block, 0, e->stream>>>(d_stack, e->W, e->H, batch, F, S, proj_w, seed, integer_disparity, noise_dn);
    SLR_CHECK_LAUNCH(e);
    return SLR_OK;
}

slr_status slr_launch_synth_gray(slr_engine *e, uint8_t *d_stack, int batch, int scan_w, unsigned seed,
                                 int integer_disparity, float noise_dn)
{
    const int nbits = slr_gray_num_bits(scan_w);
    dim3 block(128), grid((e->W + 127) / 128, e->H, batch * 2);
    k_synth_gray<<<grid, block, 0, e->stream>>>(d_stack, e->W, e->H, batch, scan_w, nbits, seed,
                                                integer_disparity, noise_dn);
    SLR_CHECK_LAUNCH(e);
    return SLR_OK;
}
