// slr_engine.cu — engine lifetime, calibration upload, the extern "C" entry points of
// include/slr_b200.h, host-side pattern synthesis and the host-buffer (end-to-end) pipelines.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <new>
#include <vector>

#include "slr_internal.h"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void slr_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *slr_last_error(void) { return g_err; }
extern "C" const char *slr_version(void) { return "slr_b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------------------------------------
// engine
// ------------------------------------------------------------------------------------------------
static void free_engine(slr_engine *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    cudaFree(e->d_undist_lx);
    cudaFree(e->d_undist_ly);
    cudaFree(e->d_undist_rx);
    cudaFree(e->d_atan_lut);
    cudaFree(e->d_ptab);
    cudaFree(e->d_btab);
    cudaFree(e->d_mesh_pn);
    cudaFree(e->d_mesh_tiles);
    cudaFree(e->d_mesh_sum);
    cudaFree(e->d_mesh_count);
    cudaFree(e->d_mesh_vertices);
    cudaFree(e->d_mesh_src);
    cudaFree(e->d_mesh_faces);
    cudaFree(e->d_mesh_counts);
    cudaFree(e->d_rays);
    for (int i = 0; i < 6; i++) cudaFree(e->d_pad[i]);
    if (e->child) {
        e->child->stream = e->child->own_stream;
        slr_destroy(e->child);
        e->child = nullptr;
    }
    cudaFree(e->d_phase);
    cudaFree(e->d_code);
    cudaFree(e->d_mask);
    for (int i = 0; i < 2; i++) {
        cudaFree(e->d_stage_in[i]);
        cudaFree(e->d_stage_xyz[i]);
        cudaFree(e->d_stage_valid[i]);
        cudaFree(e->d_stage_k[i]);
        cudaFree(e->d_stage_color[i]);
        if (e->ev_in[i]) cudaEventDestroy(e->ev_in[i]);
        if (e->ev_k[i]) cudaEventDestroy(e->ev_k[i]);
        if (e->ev_out[i]) cudaEventDestroy(e->ev_out[i]);
    }
    cudaFree(e->d_counter);
    cudaFree(e->d_bucket_scratch);
    cudaFree(e->d_minmax);
    for (void *g : e->d_gray) cudaFree(g);
    cudaFree(e->d_map1);
    cudaFree(e->d_map2);
    cudaFree(e->d_stage_rect[0]);
    cudaFree(e->d_stage_rect[1]);
    cudaFree(e->d_merge);
    cudaFree(e->d_ingest);
    cudaFree(e->d_cloud_sum);
    cudaFree(e->d_cloud_cnt);
    cudaFree(e->d_cloud_gray);
    if (e->ev_ingest) cudaEventDestroy(e->ev_ingest);
    if (e->h_counter) cudaFreeHost(e->h_counter);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    if (e->copy_in) cudaStreamDestroy(e->copy_in);
    if (e->copy_out) cudaStreamDestroy(e->copy_out);
    delete e;
}

extern "C" slr_status slr_create(slr_engine **out, int device, int width, int height, int max_batch)
{
    SLR_REQUIRE(out != nullptr, "slr_create: out is NULL");
    *out = nullptr;
    SLR_REQUIRE(width > 0 && height > 0 && max_batch > 0, "slr_create: bad size %dx%d batch %d", width, height, max_batch);
    SLR_REQUIRE(width <= 65535 && height <= 65535, "slr_create: image larger than 65535 not supported");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        slr_set_error("slr_create: no CUDA device (%s); libslr_b200 has no CPU fallback",
                      ce != cudaSuccess ? cudaGetErrorString(ce) : "device count 0");
        return SLR_ERR_CUDA;
    }
    SLR_REQUIRE(device >= 0 && device < ndev, "slr_create: device %d out of range (%d devices)", device, ndev);
    SLR_CHECK_CUDA(cudaSetDevice(device));
    slr_engine *e = new (std::nothrow) slr_engine();
    if (!e) {
        slr_set_error("slr_create: out of host memory");
        return SLR_ERR_NOMEM;
    }
    e->device = device;
    e->calib.row0 = 0;
    e->W = width;
    e->H = height;
    e->max_batch = max_batch;
    cudaDeviceProp prop;
    cudaError_t err = cudaGetDeviceProperties(&prop, device);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&e->copy_in, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&e->copy_out, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && err == cudaSuccess; i++) {
        err = cudaEventCreateWithFlags(&e->ev_in[i], cudaEventDisableTiming);
        if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e->ev_k[i], cudaEventDisableTiming);
        if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e->ev_out[i], cudaEventDisableTiming);
    }
    const size_t P = (size_t)width * height;
    if (err == cudaSuccess) err = cudaMalloc(&e->d_undist_lx, P * sizeof(float));
    if (err == cudaSuccess) err = cudaMalloc(&e->d_undist_ly, P * sizeof(float));
    if (err == cudaSuccess) err = cudaMalloc(&e->d_undist_rx, P * sizeof(float));
    if (err == cudaSuccess) err = cudaMalloc(&e->d_atan_lut, 512 * sizeof(float));
    if (err == cudaSuccess) err = cudaMalloc(&e->d_counter, sizeof(unsigned long long));
    if (err == cudaSuccess) err = cudaMallocHost(&e->h_counter, sizeof(unsigned long long));
    if (err == cudaSuccess) {
        // atan(float(q)), q = (G4-G2)/(G1-G3) in C++ int division (Duke/mfreconstruct.cpp:257-261):
        // only 511 arguments exist, so the table is built with the host libm and the device result
        // is bit-identical to the host evaluation.
        float lut[512];
        for (int q = -255; q <= 255; q++) lut[q + 255] = atanf((float)q);
        lut[511] = 0.0f;
        err = cudaMemcpy(e->d_atan_lut, lut, sizeof(lut), cudaMemcpyHostToDevice);
    }
    if (err != cudaSuccess) {
        slr_set_error("slr_create: %s", cudaGetErrorString(err));
        free_engine(e);
        return SLR_ERR_CUDA;
    }
    e->num_sms = prop.multiProcessorCount;
    e->stream = e->own_stream;
    if (slr_build_strict_tables(e) != SLR_OK) {
        free_engine(e);
        return SLR_ERR_CUDA;
    }
    *out = e;
    return SLR_OK;
}

extern "C" slr_status slr_destroy(slr_engine *e)
{
    if (!e) return SLR_OK;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    free_engine(e);
    return SLR_OK;
}

extern "C" slr_status slr_set_stream(slr_engine *e, void *cuda_stream)
{
    SLR_REQUIRE(e != nullptr, "slr_set_stream: engine is NULL");
    // NULL is the CUDA legacy default stream (what torch uses unless told otherwise), as in the CUDA API.
    e->stream = (cuda_stream == SLR_STREAM_OWN) ? e->own_stream : (cudaStream_t)cuda_stream;
    return SLR_OK;
}

extern "C" slr_status slr_synchronize(slr_engine *e)
{
    SLR_REQUIRE(e != nullptr, "slr_synchronize: engine is NULL");
    SLR_CHECK_CUDA(cudaSetDevice(e->device));
    SLR_CHECK_CUDA(cudaStreamSynchronize(e->stream));
    return SLR_OK;
}

extern "C" unsigned long long slr_kernel_launches(const slr_engine *e) { return e ? e->launches : 0ULL; }

extern "C" slr_status slr_set_calib(slr_engine *e, const slr_camera cams[2], const double Q[16],
                                    const float *rigid3x4)
{
    SLR_REQUIRE(e && cams && Q, "slr_set_calib: NULL argument");
    SLR_CHECK_CUDA(cudaSetDevice(e->device));
    e->cams[0] = cams[0];
    e->cams[1] = cams[1];
    memcpy(e->calib.Q, Q, sizeof(double) * 16);
    e->calib.has_rigid = rigid3x4 ? 1 : 0;
    memset(e->calib.rigid, 0, sizeof(e->calib.rigid));
    if (rigid3x4) memcpy(e->calib.rigid, rigid3x4, sizeof(float) * 12);
    {
        const double *q = e->calib.Q;
        const bool zeros = q[1] == 0 && q[2] == 0 && q[4] == 0 && q[6] == 0 && q[8] == 0 && q[9] == 0 && q[10] == 0 &&
                           q[12] == 0 && q[13] == 0;
        // the sign of a zero coefficient matters for the skipped +-0 terms only through the cases excluded in
        // reproject_q; non-zero, finite constants are required where they absorb a zero sum
        const bool consts = q[3] != 0 && q[7] != 0 && q[11] != 0 && isfinite(q[0]) && isfinite(q[3]) && isfinite(q[5]) &&
                            isfinite(q[7]) && isfinite(q[11]) && isfinite(q[14]) && isfinite(q[15]);
        e->calib.q_std = (zeros && consts) ? 1 : 0;
    }
    slr_status st = slr_launch_undistort_maps(e);
    if (st != SLR_OK) return st;
    e->calib_set = true;
    e->calib_version++;
    return SLR_OK;
}

static slr_status ensure_scratch(slr_engine *e, bool need_phase, bool need_code)
{
    const size_t n = (size_t)e->max_batch * 2 * e->W * e->H;
    if (!e->d_mask) SLR_CHECK_CUDA(cudaMalloc(&e->d_mask, n));
    if (need_phase && !e->d_phase) SLR_CHECK_CUDA(cudaMalloc(&e->d_phase, n * sizeof(float)));
    if (need_code && !e->d_code) SLR_CHECK_CUDA(cudaMalloc(&e->d_code, n * sizeof(int32_t)));
    return SLR_OK;
}

// ------------------------------------------------------------------------------------------------
// host-side pattern synthesis (a1, a2)
// ------------------------------------------------------------------------------------------------
extern "C" int slr_gray_num_bits(int n)
{
    // GrayCodes::calNumOfImgs: (int)ceil(log(double(n))/log(2.0))   Duke/graycodes.cpp:24-25
    return n > 0 ? (int)ceil(log((double)n) / log(2.0)) : 0;
}

extern "C" int slr_gray_num_imgs(int scan_w, int scan_h, int use_epi)
{
    const int nc = slr_gray_num_bits(scan_w), nr = slr_gray_num_bits(scan_h);
    return use_epi ? 2 + 2 * nc : 2 + 2 * nc + 2 * nr;
}

extern "C" slr_status slr_generate_gray_patterns(uint8_t *h_out, int W, int H, int use_epi)
{
    SLR_REQUIRE(h_out && W > 0 && H > 0, "slr_generate_gray_patterns: bad argument");
    const int nc = slr_gray_num_bits(W), nr = slr_gray_num_bits(H);
    const size_t P = (size_t)W * H;
    memset(h_out, 255, P);
    memset(h_out + P, 0, P);
    // Image 2+2c carries Gray bit (nc-1-c) of the column index (MSB first), 3+2c its inverse:
    // the index mapping of GrayCodes::generateGrays (Duke/graycodes.cpp:63-85) with g = j ^ (j >> 1).
    std::vector<uint8_t> line((size_t)W);
    for (int c = 0; c < nc; c++) {
        uint8_t *img = h_out + (size_t)(2 + 2 * c) * P, *inv = h_out + (size_t)(3 + 2 * c) * P;
        for (int j = 0; j < W; j++) line[j] = (((j ^ (j >> 1)) >> (nc - 1 - c)) & 1) ? 255 : 0;
        for (int i = 0; i < H; i++) {
            memcpy(img + (size_t)i * W, line.data(), (size_t)W);
            for (int j = 0; j < W; j++) inv[(size_t)i * W + j] = (uint8_t)(255 - line[j]);
        }
    }
    if (!use_epi) {
        for (int c = 0; c < nr; c++) {
            uint8_t *img = h_out + (size_t)(2 + 2 * nc + 2 * c) * P, *inv = h_out + (size_t)(3 + 2 * nc + 2 * c) * P;
            for (int i = 0; i < H; i++) {
                const uint8_t v = (((i ^ (i >> 1)) >> (nr - 1 - c)) & 1) ? 255 : 0;
                memset(img + (size_t)i * W, v, (size_t)W);
                memset(inv + (size_t)i * W, 255 - v, (size_t)W);
            }
        }
    }
    return SLR_OK;
}

extern "C" slr_status slr_generate_mf_patterns(uint8_t *h_out, int projW, int projH)
{
    SLR_REQUIRE(h_out && projW > 0 && projH > 0, "slr_generate_mf_patterns: bad argument");
    const size_t P = (size_t)projW * projH;
    memset(h_out, 255, P);
    memset(h_out + P, 0, P);
    static const int freq[3] = {70, 64, 59};  // Duke/multifrequency.cpp:3
    const double PI = 3.1416;                 // Duke/multifrequency.h:5
    std::vector<uint8_t> line((size_t)projW);
    for (int f = 0; f < 3; f++)
        for (int s = 0; s < 4; s++) {
            for (int w = 0; w < projW; w++) {
                // 135+79*cos(float(PI*2*w*frequency[f]/projW+PI*phi/2))   Duke/multifrequency.cpp:27
                const double arg = PI * 2 * (double)w * (double)freq[f] / (double)projW + PI * (double)s / 2;
                const float v = 135.0f + 79.0f * cosf((float)arg);
                line[w] = (uint8_t)v;
            }
            uint8_t *img = h_out + (size_t)(2 + 4 * f + s) * P;
            for (int h = 0; h < projH; h++) memcpy(img + (size_t)h * projW, line.data(), (size_t)projW);
        }
    return SLR_OK;
}

// ------------------------------------------------------------------------------------------------
// multi-GPU assembly over NVLink: peer-mapped cloud buffers, gather targets, row bands
// ------------------------------------------------------------------------------------------------
extern "C" slr_status slr_set_row_offset(slr_engine *e, int first_row)
{
    SLR_REQUIRE(e != nullptr && first_row >= 0, "slr_set_row_offset: bad argument");
    SLR_CHECK_CUDA(cudaSetDevice(e->device));
    if (e->calib.row0 == first_row) return SLR_OK;
    e->calib.row0 = first_row;
    if (!e->calib_set) return SLR_OK;
    e->calib_version++;                     // the undistort maps (and a padded child's) depend on the absolute row
    return slr_launch_undistort_maps(e);
}

extern "C" slr_status slr_peer_alloc(slr_engine *e, size_t bytes, void **d_ptr, void *ipc_handle64)
{
    SLR_REQUIRE(e && d_ptr && ipc_handle64 && bytes > 0, "slr_peer_alloc: bad argument");
    SLR_CHECK_CUDA(cudaSetDevice(e->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    SLR_CHECK_CUDA(cudaMalloc(d_ptr, bytes));
    cudaIpcMemHandle_t h;
    const cudaError_t err = cudaIpcGetMemHandle(&h, *d_ptr);
    if (err != cudaSuccess) {
        cudaFree(*d_ptr);
        *d_ptr = nullptr;
        slr_set_error("slr_peer_alloc: cudaIpcGetMemHandle: %s", cudaGetErrorString(err));
        return SLR_ERR_CUDA;
    }
    memcpy(ipc_handle64, &h, 64);
    return SLR_OK;
}

extern "C" slr_status slr_peer_open(slr_engine *e, const void *ipc_handle64, void **d_ptr)
{
    SLR_REQUIRE(e && d_ptr && ipc_handle64, "slr_peer_open: bad argument");
    SLR_CHECK_CUDA(cudaSetDevice(e->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle64, 64);
    SLR_CHECK_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SLR_OK;
}

extern "C" slr_status slr_peer_close(slr_engine *e, void *d_ptr)
{
    SLR_REQUIRE(e != nullptr, "slr_peer_close: engine is NULL");
    SLR_CHECK_CUDA(cudaSetDevice(e->device));
    if (d_ptr) SLR_CHECK_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return SLR_OK;
}

extern "C" slr_status slr_peer_free(slr_engine *e, void *d_ptr)
{
    SLR_REQUIRE(e != nullptr, "slr_peer_free: engine is NULL");
    SLR_CHECK_CUDA(cudaSetDevice(e->device));
    if (d_ptr) SLR_CHECK_CUDA(cudaFree(d_ptr));
    return SLR_OK;
}

extern "C" slr_status slr_set_gather_targets(slr_engine *e, int n_targets, float *const *d_xyz_all,
                                             uint8_t *const *d_valid_all, long long first_scan)
{
    SLR_REQUIRE(e != nullptr && n_targets >= 0 && n_targets <= SLR_MAX_TARGETS && first_scan >= 0,
                "slr_set_gather_targets: bad argument (at most %d targets)", SLR_MAX_TARGETS);
    SLR_REQUIRE(n_targets == 0 || (d_xyz_all && d_valid_all), "slr_set_gather_targets: NULL target arrays");
    for (int t = 0; t < n_targets; t++) {
        SLR_REQUIRE(d_xyz_all[t] && d_valid_all[t], "slr_set_gather_targets: target %d is NULL", t);
        e->xyz_t[t] = d_xyz_all[t];
        e->valid_t[t] = d_valid_all[t];
    }
    e->n_targets = n_targets;
    e->target_first_scan = first_scan;
    return SLR_OK;
}

// this engine's block of its own assembled cloud (target 0): where the kernels write when targets are set
static void slr_own_target_block(slr_engine *e, float **d_xyz, uint8_t **d_valid)
{
    const size_t off = (size_t)e->target_first_scan * e->W * e->H;
    *d_xyz = e->xyz_t[0] + off * 3;
    *d_valid = e->valid_t[0] + off;
}

// routes whose kernels do not store to the peers themselves: copy the finished block over NVLink (stream ordered)
static slr_status slr_copy_to_targets(slr_engine *e, int batch)
{
    if (e->n_targets <= 1 || e->targets_written) return SLR_OK;
    const size_t off = (size_t)e->target_first_scan * e->W * e->H, px = (size_t)batch * e->W * e->H;
    for (int t = 1; t < e->n_targets; t++) {
        SLR_CHECK_CUDA(cudaMemcpyAsync(e->xyz_t[t] + off * 3, e->xyz_t[0] + off * 3, px * 3 * sizeof(float), cudaMemcpyDeviceToDevice, e->stream));
        SLR_CHECK_CUDA(cudaMemcpyAsync(e->valid_t[t] + off, e->valid_t[0] + off, px, cudaMemcpyDeviceToDevice, e->stream));
    }
    return SLR_OK;
}

// ------------------------------------------------------------------------------------------------
// device-pointer entry points
// ------------------------------------------------------------------------------------------------
#define SLR_ENTER(e)                                                     \
    SLR_REQUIRE((e) != nullptr, "%s: engine is NULL", __func__);         \
    SLR_CHECK_CUDA(cudaSetDevice((e)->device))

extern "C" slr_status slr_mf_decode(slr_engine *e, const uint8_t *d_stack, int batch, int F, int S,
                                    int black_thr, int mode, float *d_phase, uint8_t *d_mask)
{
    SLR_ENTER(e);
    SLR_REQUIRE(d_stack && d_phase && d_mask && batch > 0, "slr_mf_decode: bad argument");
    return slr_launch_mf_decode(e, d_stack, batch * 2, F, S, black_thr, mode, d_phase, d_mask);
}

extern "C" slr_status slr_gray_decode(slr_engine *e, const uint8_t *d_stack, int batch, int nbits_col,
                                      int nbits_row, int black_thr, int white_thr, int scan_w, int scan_h,
                                      int32_t *d_col, int32_t *d_row, uint8_t *d_mask)
{
    SLR_ENTER(e);
    SLR_REQUIRE(d_stack && d_col && d_mask && batch > 0, "slr_gray_decode: bad argument");
    SLR_REQUIRE(nbits_col >= 1 && nbits_col <= 16 && nbits_row >= 0 && nbits_row <= 16,
                "slr_gray_decode: bit counts out of range (col %d, row %d)", nbits_col, nbits_row);
    SLR_REQUIRE(nbits_row == 0 || d_row != nullptr, "slr_gray_decode: d_row is NULL but nbits_row > 0");
    return slr_launch_gray_decode(e, d_stack, batch * 2, nbits_col, nbits_row, black_thr, white_thr, scan_w,
                                  scan_h, d_col, d_row, d_mask);
}

extern "C" slr_status slr_match_triangulate_phase(slr_engine *e, const float *d_phase, const uint8_t *d_mask,
                                                  int batch, float *d_xyz, uint8_t *d_valid, int32_t *d_match_k,
                                                  unsigned long long *d_n_points)
{
    SLR_ENTER(e);
    SLR_REQUIRE(d_phase && d_mask && d_xyz && d_valid && batch > 0, "slr_match_triangulate_phase: bad argument");
    if (!e->calib_set) {
        slr_set_error("slr_match_triangulate_phase: call slr_set_calib first");
        return SLR_ERR_STATE;
    }
    // TMA rows need 16-byte multiples and 16-byte aligned bases: anything else is staged through aligned copies
    if (e->W % 16 != 0 || slr_misaligned16(d_phase, d_mask, d_xyz, d_valid, d_match_k))
        return slr_padded_run(e, 2, d_phase, d_mask, batch, 1, 0, 0, 0, 0, 0, d_xyz, d_valid, d_match_k, nullptr, d_n_points);
    return slr_launch_match_phase(e, d_phase, d_mask, batch, d_xyz, d_valid, d_match_k, d_n_points);
}

extern "C" slr_status slr_match_triangulate_code(slr_engine *e, const int32_t *d_col, const uint8_t *d_mask,
                                                 int batch, const uint8_t *d_white, float *d_xyz,
                                                 uint8_t *d_valid, int32_t *d_match_k, uint8_t *d_color,
                                                 unsigned long long *d_n_points)
{
    SLR_ENTER(e);
    SLR_REQUIRE(d_col && d_mask && d_xyz && d_valid && batch > 0, "slr_match_triangulate_code: bad argument");
    SLR_REQUIRE((d_color == nullptr) == (d_white == nullptr), "slr_match_triangulate_code: d_white and d_color go together");
    if (!e->calib_set) {
        slr_set_error("slr_match_triangulate_code: call slr_set_calib first");
        return SLR_ERR_STATE;
    }
    if (e->W % 16 != 0 || slr_misaligned16(d_col, d_mask, d_xyz, d_valid, d_match_k, d_white, d_color)) {
        SLR_REQUIRE(d_color == nullptr, "slr_match_triangulate_code: colour output needs an image width that is a multiple "
                                        "of 16 and 16-byte aligned pointers (use slr_run_ge, which stages the whole stack); "
                                        "got width %d", e->W);
        return slr_padded_run(e, 3, d_col, d_mask, batch, 1, 0, 0, 0, 0, 0, d_xyz, d_valid, d_match_k, nullptr, d_n_points);
    }
    return slr_launch_match_code(e, d_col, d_mask, batch, d_white, (size_t)e->W * e->H, d_xyz, d_valid, d_match_k,
                                 d_color, d_n_points);
}

extern "C" slr_status slr_bucket_triangulate(slr_engine *e, const int32_t *d_col, const int32_t *d_row,
                                             const uint8_t *d_mask, int batch, int scan_w, int scan_h,
                                             float *d_sum, uint8_t *d_cnt, unsigned long long *d_n_cells)
{
    SLR_ENTER(e);
    SLR_REQUIRE(d_col && d_row && d_mask && d_sum && d_cnt && batch > 0 && scan_w > 0 && scan_h > 0,
                "slr_bucket_triangulate: bad argument");
    if (!e->calib_set) {
        slr_set_error("slr_bucket_triangulate: call slr_set_calib first");
        return SLR_ERR_STATE;
    }
    SLR_REQUIRE(e->calib.row0 == 0, "slr_bucket_triangulate: the Gray-only path does not partition by row bands "
                                    "(projector-cell buckets span the image); row offset must be 0");
    return slr_launch_bucket_triangulate(e, d_col, d_row, d_mask, batch, scan_w, scan_h, d_sum, d_cnt, d_n_cells);
}

extern "C" slr_status slr_run_mf(slr_engine *e, const uint8_t *d_stack, int batch, int F, int S, int black_thr,
                                 int mode, float *d_xyz, uint8_t *d_valid, int32_t *d_match_k,
                                 unsigned long long *d_n_points)
{
    SLR_ENTER(e);
    SLR_REQUIRE(d_stack && batch > 0 && (e->n_targets > 0 || (d_xyz && d_valid)), "slr_run_mf: bad argument");
    if (!e->calib_set) {
        slr_set_error("slr_run_mf: call slr_set_calib first");
        return SLR_ERR_STATE;
    }
    if (e->n_targets > 0) slr_own_target_block(e, &d_xyz, &d_valid);
    e->targets_written = false;
    const slr_status st = slr_launch_fused_mf(e, d_stack, batch, F, S, black_thr, mode, d_xyz, d_valid, d_match_k, d_n_points);
    return st == SLR_OK ? slr_copy_to_targets(e, batch) : st;
}

extern "C" slr_status slr_run_mf_raw(slr_engine *e, const uint8_t *d_raw_stack, int batch, int F, int S, int black_thr,
                                     int mode, float *d_xyz, uint8_t *d_valid, int32_t *d_match_k,
                                     unsigned long long *d_n_points)
{
    SLR_ENTER(e);
    SLR_REQUIRE(d_raw_stack && batch > 0 && (e->n_targets > 0 || (d_xyz && d_valid)), "slr_run_mf_raw: bad argument");
    if (!e->calib_set || !e->maps_set) {
        slr_set_error("slr_run_mf_raw: call slr_set_calib and slr_set_rectify_maps first");
        return SLR_ERR_STATE;
    }
    if (mode == SLR_MODE_STRICT)
        SLR_REQUIRE(F == 3 && S == 4, "strict mode reproduces the reference's hard-coded 3 frequencies x 4 steps; got F=%d S=%d", F, S);
    else
        SLR_REQUIRE(mode == SLR_MODE_CORRECTED && F >= 1 && F <= 8 && S >= 3 && S <= 16,
                    "corrected mode supports 1<=F<=8, 3<=S<=16; got mode %d F=%d S=%d", mode, F, S);
    if (e->n_targets > 0) slr_own_target_block(e, &d_xyz, &d_valid);
    e->targets_written = false;
    const slr_status st = slr_launch_fused_mf_raw(e, d_raw_stack, batch, F, S, black_thr, mode, d_xyz, d_valid, d_match_k, d_n_points);
    return st == SLR_OK ? slr_copy_to_targets(e, batch) : st;
}

extern "C" slr_status slr_run_ge(slr_engine *e, const uint8_t *d_stack, int batch, int nbits_col, int black_thr,
                                 int white_thr, int scan_w, int have_color, float *d_xyz, uint8_t *d_valid,
                                 int32_t *d_match_k, uint8_t *d_color, unsigned long long *d_n_points)
{
    SLR_ENTER(e);
    SLR_REQUIRE(d_stack && batch > 0 && (e->n_targets > 0 || (d_xyz && d_valid)), "slr_run_ge: bad argument");
    SLR_REQUIRE(!have_color || d_color, "slr_run_ge: have_color set but d_color is NULL");
    if (!e->calib_set) {
        slr_set_error("slr_run_ge: call slr_set_calib first");
        return SLR_ERR_STATE;
    }
    if (e->n_targets > 0) slr_own_target_block(e, &d_xyz, &d_valid);
    e->targets_written = false;
    const slr_status st = slr_launch_fused_ge(e, d_stack, batch, nbits_col, black_thr, white_thr, scan_w, have_color, d_xyz,
                                              d_valid, d_match_k, d_color, d_n_points);
    return st == SLR_OK ? slr_copy_to_targets(e, batch) : st;
}

// ------------------------------------------------------------------------------------------------
// widths that are not a multiple of 16
// ------------------------------------------------------------------------------------------------
// The match kernels move rows with TMA bulk copies (16-byte rows).  Other widths run on a child engine of the padded
// width Wp over zero-padded device copies: padded columns have white == black == 0 (MF/GE stacks) or mask == 0
// (phase / code rows), so they carry no phase or code, can never be matched and never match; column indices of real
// pixels are unchanged.  Outputs are copied back row by row.  Costs one extra pass over the data, for odd widths only.
static slr_status pad_buffer(slr_engine *e, int slot, size_t bytes)
{
    if (e->pad_bytes[slot] >= bytes) return SLR_OK;
    cudaFree(e->d_pad[slot]);
    e->d_pad[slot] = nullptr;
    e->pad_bytes[slot] = 0;
    SLR_CHECK_CUDA(cudaMalloc(&e->d_pad[slot], bytes));
    e->pad_bytes[slot] = bytes;
    return SLR_OK;
}

slr_status slr_padded_run(slr_engine *e, int kind, const void *d_in0, const uint8_t *d_in1, int batch, int planes,
                          int a0, int a1, int a2, int a3, int a4, float *d_xyz, uint8_t *d_valid, int32_t *d_match_k,
                          uint8_t *d_color, unsigned long long *d_n_points)
{
    const int W = e->W, H = e->H, Wp = (W + 15) & ~15;
    if (!e->child) {
        slr_status st = slr_create(&e->child, e->device, Wp, H, e->max_batch);
        if (st != SLR_OK) return st;
    }
    slr_engine *c = e->child;
    c->stream = e->stream;
    if (c->calib_version == 0 || e->child_calib_version != e->calib_version) {
        c->calib.row0 = e->calib.row0;
        slr_status st = slr_set_calib(c, e->cams, e->calib.Q, e->calib.has_rigid ? e->calib.rigid : nullptr);
        if (st != SLR_OK) return st;
        e->child_calib_version = e->calib_version;
    }
    const size_t rows_in = (size_t)batch * 2 * planes * H, rows_out = (size_t)batch * H;
    const size_t es0 = (kind <= 1) ? 1 : 4;   // element size of input 0: image bytes, or f32 phase / i32 code
    slr_status st;
    if ((st = pad_buffer(e, 0, rows_in * Wp * es0)) != SLR_OK) return st;
    if (kind >= 2 && (st = pad_buffer(e, 1, rows_in * Wp)) != SLR_OK) return st;
    if ((st = pad_buffer(e, 2, rows_out * Wp * 12)) != SLR_OK) return st;
    if ((st = pad_buffer(e, 3, rows_out * Wp)) != SLR_OK) return st;
    if (d_match_k && (st = pad_buffer(e, 4, rows_out * Wp * 4)) != SLR_OK) return st;
    if (d_color && (st = pad_buffer(e, 5, rows_out * Wp)) != SLR_OK) return st;
    cudaStream_t cs = e->stream;
    // input 0 needs no defined padding for phase / code rows (their mask decides), image stacks need zeros
    SLR_CHECK_CUDA(cudaMemsetAsync(e->d_pad[kind >= 2 ? 1 : 0], 0, rows_in * Wp * (kind >= 2 ? 1 : es0), cs));
    if (kind >= 2) SLR_CHECK_CUDA(cudaMemsetAsync(e->d_pad[0], 0, rows_in * Wp * es0, cs));
    SLR_CHECK_CUDA(cudaMemcpy2DAsync(e->d_pad[0], (size_t)Wp * es0, d_in0, (size_t)W * es0, (size_t)W * es0, rows_in,
                                     cudaMemcpyDeviceToDevice, cs));
    if (kind >= 2)
        SLR_CHECK_CUDA(cudaMemcpy2DAsync(e->d_pad[1], Wp, d_in1, W, W, rows_in, cudaMemcpyDeviceToDevice, cs));
    float *p_xyz = (float *)e->d_pad[2];
    uint8_t *p_valid = (uint8_t *)e->d_pad[3];
    int32_t *p_k = d_match_k ? (int32_t *)e->d_pad[4] : nullptr;
    uint8_t *p_color = d_color ? (uint8_t *)e->d_pad[5] : nullptr;
    const unsigned long long l0 = c->launches;
    switch (kind) {
    case 0:  // a0 = F, a1 = S, a2 = black_thr, a3 = mode
        st = slr_launch_fused_mf(c, (const uint8_t *)e->d_pad[0], batch, a0, a1, a2, a3, p_xyz, p_valid, p_k, d_n_points);
        break;
    case 1:  // a0 = nbits_col, a1 = black_thr, a2 = white_thr, a3 = scan_w, a4 = have_color
        st = slr_launch_fused_ge(c, (const uint8_t *)e->d_pad[0], batch, a0, a1, a2, a3, a4, p_xyz, p_valid, p_k, p_color,
                                 d_n_points);
        break;
    case 2:
        st = slr_launch_match_phase(c, (const float *)e->d_pad[0], (const uint8_t *)e->d_pad[1], batch, p_xyz, p_valid, p_k,
                                    d_n_points);
        break;
    default:
        st = slr_launch_match_code(c, (const int32_t *)e->d_pad[0], (const uint8_t *)e->d_pad[1], batch, nullptr, 0, p_xyz,
                                   p_valid, p_k, nullptr, d_n_points);
        break;
    }
    e->launches += c->launches - l0;
    if (st != SLR_OK) return st;
    SLR_CHECK_CUDA(cudaMemcpy2DAsync(d_xyz, (size_t)W * 12, p_xyz, (size_t)Wp * 12, (size_t)W * 12, rows_out,
                                     cudaMemcpyDeviceToDevice, cs));
    SLR_CHECK_CUDA(cudaMemcpy2DAsync(d_valid, W, p_valid, Wp, W, rows_out, cudaMemcpyDeviceToDevice, cs));
    if (d_match_k)
        SLR_CHECK_CUDA(cudaMemcpy2DAsync(d_match_k, (size_t)W * 4, p_k, (size_t)Wp * 4, (size_t)W * 4, rows_out,
                                         cudaMemcpyDeviceToDevice, cs));
    if (d_color) SLR_CHECK_CUDA(cudaMemcpy2DAsync(d_color, W, p_color, Wp, W, rows_out, cudaMemcpyDeviceToDevice, cs));
    return SLR_OK;
}

// ------------------------------------------------------------------------------------------------
// K5: mesh indexing (MeshCreator's vertex numbering + faces)
// ------------------------------------------------------------------------------------------------
static slr_status ensure_mesh_scratch(slr_engine *e, size_t px, bool host_staging)
{
    if (e->mesh_px < px) {
        cudaFree(e->d_mesh_pn);
        cudaFree(e->d_mesh_tiles);
        e->d_mesh_pn = e->d_mesh_tiles = nullptr;
        e->mesh_px = 0;
        SLR_CHECK_CUDA(cudaMalloc(&e->d_mesh_pn, px * sizeof(int)));
        SLR_CHECK_CUDA(cudaMalloc(&e->d_mesh_tiles, (px / 1024 + 2) * sizeof(int)));
        e->mesh_px = px;
    }
    if (host_staging && e->mesh_host_px < px) {
        cudaFree(e->d_mesh_sum);
        cudaFree(e->d_mesh_count);
        cudaFree(e->d_mesh_vertices);
        cudaFree(e->d_mesh_src);
        cudaFree(e->d_mesh_faces);
        cudaFree(e->d_mesh_counts);
        e->d_mesh_sum = e->d_mesh_vertices = nullptr;
        e->d_mesh_count = nullptr;
        e->d_mesh_src = e->d_mesh_faces = nullptr;
        e->d_mesh_counts = nullptr;
        e->mesh_host_px = 0;
        SLR_CHECK_CUDA(cudaMalloc(&e->d_mesh_sum, px * 3 * sizeof(float)));
        SLR_CHECK_CUDA(cudaMalloc(&e->d_mesh_count, px));
        SLR_CHECK_CUDA(cudaMalloc(&e->d_mesh_vertices, px * 3 * sizeof(float)));
        SLR_CHECK_CUDA(cudaMalloc(&e->d_mesh_src, px * sizeof(int32_t)));
        SLR_CHECK_CUDA(cudaMalloc(&e->d_mesh_faces, px * 6 * sizeof(int32_t)));
        SLR_CHECK_CUDA(cudaMalloc(&e->d_mesh_counts, 2 * sizeof(unsigned long long)));
        e->mesh_host_px = px;
    }
    return SLR_OK;
}

extern "C" slr_status slr_mesh_index(slr_engine *e, const float *d_sum, const uint8_t *d_count, int w, int h,
                                     int first_vertex, float *d_vertices, int32_t *d_vertex_src, int32_t *d_faces,
                                     unsigned long long *d_counts)
{
    SLR_ENTER(e);
    SLR_REQUIRE(d_sum && d_count && d_vertices && d_faces && d_counts && w > 0 && h > 0 &&
                (long long)w * h < (1LL << 30) && (first_vertex == 0 || first_vertex == 1), "slr_mesh_index: bad argument");
    slr_status st = ensure_mesh_scratch(e, (size_t)w * h, false);
    if (st != SLR_OK) return st;
    return slr_launch_mesh_index(e, d_sum, d_count, w, h, first_vertex, e->d_mesh_pn, e->d_mesh_tiles, d_vertices,
                                 d_vertex_src, d_faces, d_counts);
}

extern "C" slr_status slr_mesh_index_host(slr_engine *e, const float *h_sum, const uint8_t *h_count, int w, int h,
                                          int first_vertex, float *h_vertices, int32_t *h_vertex_src, int32_t *h_faces,
                                          unsigned long long *h_counts)
{
    SLR_ENTER(e);
    SLR_REQUIRE(h_sum && h_count && h_vertices && h_faces && h_counts && w > 0 && h > 0 &&
                (long long)w * h < (1LL << 30) && (first_vertex == 0 || first_vertex == 1),
                "slr_mesh_index_host: bad argument");
    const size_t px = (size_t)w * h;
    slr_status st = ensure_mesh_scratch(e, px, true);
    if (st != SLR_OK) return st;
    SLR_CHECK_CUDA(cudaMemcpyAsync(e->d_mesh_sum, h_sum, px * 3 * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    SLR_CHECK_CUDA(cudaMemcpyAsync(e->d_mesh_count, h_count, px, cudaMemcpyHostToDevice, e->stream));
    st = slr_launch_mesh_index(e, e->d_mesh_sum, e->d_mesh_count, w, h, first_vertex, e->d_mesh_pn, e->d_mesh_tiles,
                               e->d_mesh_vertices, e->d_mesh_src, e->d_mesh_faces, e->d_mesh_counts);
    if (st != SLR_OK) return st;
    SLR_CHECK_CUDA(cudaMemcpyAsync(h_counts, e->d_mesh_counts, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                   e->stream));
    SLR_CHECK_CUDA(cudaStreamSynchronize(e->stream));
    const size_t nv = (size_t)h_counts[0], nf = (size_t)h_counts[1];
    if (nv) SLR_CHECK_CUDA(cudaMemcpyAsync(h_vertices, e->d_mesh_vertices, nv * 3 * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    if (nv && h_vertex_src)
        SLR_CHECK_CUDA(cudaMemcpyAsync(h_vertex_src, e->d_mesh_src, nv * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    if (nf) SLR_CHECK_CUDA(cudaMemcpyAsync(h_faces, e->d_mesh_faces, nf * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    SLR_CHECK_CUDA(cudaStreamSynchronize(e->stream));
    return SLR_OK;
}

extern "C" slr_status slr_synth_mf(slr_engine *e, uint8_t *d_stack, int batch, int proj_w, unsigned seed,
                                   int integer_disparity, float noise_dn)
{
    SLR_ENTER(e);
    SLR_REQUIRE(d_stack && batch > 0 && batch <= 32767 && proj_w > 0, "slr_synth_mf: bad argument");
    return slr_launch_synth_mf(e, d_stack, batch, 3, 4, proj_w, seed, integer_disparity, noise_dn);
}

extern "C" slr_status slr_synth_mf_fs(slr_engine *e, uint8_t *d_stack, int batch, int F, int S, int proj_w,
                                      unsigned seed, int integer_disparity, float noise_dn)
{
    SLR_ENTER(e);
    SLR_REQUIRE(d_stack && batch > 0 && batch <= 32767 && proj_w > 0 && F >= 1 && F <= 8 && S >= 3 && S <= 16,
                "slr_synth_mf_fs: bad argument");
    return slr_launch_synth_mf(e, d_stack, batch, F, S, proj_w, seed, integer_disparity, noise_dn);
}

extern "C" slr_status slr_synth_gray(slr_engine *e, uint8_t *d_stack, int batch, int scan_w, unsigned seed,
                                     int integer_disparity, float noise_dn)
{
    SLR_ENTER(e);
    SLR_REQUIRE(d_stack && batch > 0 && batch <= 32767 && scan_w > 1, "slr_synth_gray: bad argument");
    return slr_launch_synth_gray(e, d_stack, batch, scan_w, seed, integer_disparity, noise_dn);
}

// ------------------------------------------------------------------------------------------------
// K0: rectification on load
// ------------------------------------------------------------------------------------------------
extern "C" slr_status slr_set_rectify_maps(slr_engine *e, const int16_t *h_map1, const uint16_t *h_map2)
{
    SLR_ENTER(e);
    SLR_REQUIRE(h_map1 && h_map2, "slr_set_rectify_maps: NULL map");
    const size_t P = (size_t)e->W * e->H;
    if (!e->d_map1) SLR_CHECK_CUDA(cudaMalloc(&e->d_map1, 2 * P * 2 * sizeof(int16_t)));
    if (!e->d_map2) SLR_CHECK_CUDA(cudaMalloc(&e->d_map2, 2 * P * sizeof(uint16_t)));
    SLR_CHECK_CUDA(cudaMemcpyAsync(e->d_map1, h_map1, 2 * P * 2 * sizeof(int16_t), cudaMemcpyHostToDevice, e->stream));
    SLR_CHECK_CUDA(cudaMemcpyAsync(e->d_map2, h_map2, 2 * P * sizeof(uint16_t), cudaMemcpyHostToDevice, e->stream));
    SLR_CHECK_CUDA(cudaStreamSynchronize(e->stream));
    e->maps_set = true;
    return SLR_OK;
}

extern "C" slr_status slr_rectify_stack(slr_engine *e, const uint8_t *d_raw, int batch, int n_images, uint8_t *d_out)
{
    SLR_ENTER(e);
    SLR_REQUIRE(d_raw && d_out && d_raw != d_out && batch > 0 && n_images > 0 && batch <= 32767,
                "slr_rectify_stack: bad argument (in-place is not supported)");
    if (!e->maps_set) {
        slr_set_error("slr_rectify_stack: call slr_set_rectify_maps first");
        return SLR_ERR_STATE;
    }
    return slr_launch_rectify(e, d_raw, batch, n_images, d_out);
}

extern "C" slr_status slr_set_host_input_raw(slr_engine *e, int raw)
{
    SLR_ENTER(e);
    if (raw && !e->maps_set) {
        slr_set_error("slr_set_host_input_raw: call slr_set_rectify_maps first");
        return SLR_ERR_STATE;
    }
    e->host_input_raw = raw != 0;
    return SLR_OK;
}

extern "C" slr_status slr_auto_contrast(slr_engine *e, uint8_t *d_images, int n_images)
{
    SLR_ENTER(e);
    SLR_REQUIRE(d_images && n_images > 0, "slr_auto_contrast: bad argument");
    return slr_launch_auto_contrast(e, d_images, n_images);
}

extern "C" slr_status slr_set_auto_contrast(slr_engine *e, int on)
{
    SLR_ENTER(e);
    e->auto_contrast = on != 0;
    return SLR_OK;
}

// ------------------------------------------------------------------------------------------------
// host-buffer pipelines
// ------------------------------------------------------------------------------------------------
extern "C" slr_status slr_host_alloc(void **out, size_t bytes)
{
    SLR_REQUIRE(out != nullptr, "slr_host_alloc: out is NULL");
    SLR_CHECK_CUDA(cudaMallocHost(out, bytes));
    return SLR_OK;
}

extern "C" slr_status slr_host_free(void *p)
{
    if (p) SLR_CHECK_CUDA(cudaFreeHost(p));
    return SLR_OK;
}

static slr_status ensure_stage(slr_engine *e, size_t in_bytes_per_scan, bool want_color)
{
    const size_t P = (size_t)e->W * e->H;
    for (int i = 0; i < 2; i++) {
        if (e->stage_in_bytes < in_bytes_per_scan) {
            if (e->d_stage_in[i]) SLR_CHECK_CUDA(cudaFree(e->d_stage_in[i]));
            e->d_stage_in[i] = nullptr;
            SLR_CHECK_CUDA(cudaMalloc(&e->d_stage_in[i], in_bytes_per_scan));
        }
        if (!e->d_stage_xyz[i]) SLR_CHECK_CUDA(cudaMalloc(&e->d_stage_xyz[i], P * 3 * sizeof(float)));
        if (!e->d_stage_valid[i]) SLR_CHECK_CUDA(cudaMalloc(&e->d_stage_valid[i], P));
        if (!e->d_stage_k[i]) SLR_CHECK_CUDA(cudaMalloc(&e->d_stage_k[i], P * sizeof(int32_t)));
        if (want_color && !e->d_stage_color[i]) SLR_CHECK_CUDA(cudaMalloc(&e->d_stage_color[i], P));
    }
    if (e->stage_in_bytes < in_bytes_per_scan) e->stage_in_bytes = in_bytes_per_scan;
    if (e->host_input_raw && e->stage_rect_bytes < in_bytes_per_scan) {
        for (int i = 0; i < 2; i++) {
            if (e->d_stage_rect[i]) SLR_CHECK_CUDA(cudaFree(e->d_stage_rect[i]));
            e->d_stage_rect[i] = nullptr;
            SLR_CHECK_CUDA(cudaMalloc(&e->d_stage_rect[i], in_bytes_per_scan));
        }
        e->stage_rect_bytes = in_bytes_per_scan;
    }
    return SLR_OK;
}

// Scan-by-scan software pipeline over three streams: H2D of scan s+1 overlaps the kernels of scan s
// and the D2H of scan s-1 (PCIe is full duplex; the kernels take microseconds, the copies ~1 ms).
template <typename LaunchFn>
static slr_status host_pipeline_body(slr_engine *e, const uint8_t *h_stack, size_t in_bytes, int batch, bool color,
                                     float *h_xyz, uint8_t *h_valid, int32_t *h_match_k, uint8_t *h_color,
                                     unsigned long long *h_n_points, LaunchFn launch, bool rectify_in_launch)
{
    const size_t P = (size_t)e->W * e->H;
    slr_status st = ensure_stage(e, in_bytes, color);
    if (st != SLR_OK) return st;
    cudaStream_t cs = e->stream;
    SLR_CHECK_CUDA(cudaMemsetAsync(e->d_counter, 0, sizeof(unsigned long long), cs));
    for (int s = 0; s < batch; s++) {
        const int b = s & 1;
        // stage_in[b] is free once the kernels of scan s-2 are done
        SLR_CHECK_CUDA(cudaStreamWaitEvent(e->copy_in, e->ev_k[b], 0));
        SLR_CHECK_CUDA(cudaMemcpyAsync(e->d_stage_in[b], h_stack + (size_t)s * in_bytes, in_bytes,
                                       cudaMemcpyHostToDevice, e->copy_in));
        SLR_CHECK_CUDA(cudaEventRecord(e->ev_in[b], e->copy_in));
        // outputs of buffer b are free once the D2H of scan s-2 is done
        SLR_CHECK_CUDA(cudaStreamWaitEvent(cs, e->ev_in[b], 0));
        SLR_CHECK_CUDA(cudaStreamWaitEvent(cs, e->ev_out[b], 0));
        uint8_t *d_in = e->d_stage_in[b];
        if (e->host_input_raw && !rectify_in_launch) {  // K0: rectify the raw camera images on the GPU (stereoRect::doStereoRectify)
            st = slr_launch_rectify(e, d_in, 1, (int)(in_bytes / (2 * P)), e->d_stage_rect[b]);
            if (st != SLR_OK) return st;
            d_in = e->d_stage_rect[b];
        }
        if (e->auto_contrast) {   // Utilities::autoContrast on every loaded image (Duke/reconstruct.cpp:182-183), in place
            st = slr_launch_auto_contrast(e, d_in, (int)(in_bytes / P));
            if (st != SLR_OK) return st;
        }
        st = launch(d_in, e->d_stage_xyz[b], e->d_stage_valid[b], e->d_stage_k[b], e->d_stage_color[b]);
        if (st != SLR_OK) return st;
        SLR_CHECK_CUDA(cudaEventRecord(e->ev_k[b], cs));
        SLR_CHECK_CUDA(cudaStreamWaitEvent(e->copy_out, e->ev_k[b], 0));
        SLR_CHECK_CUDA(cudaMemcpyAsync(h_xyz + (size_t)s * P * 3, e->d_stage_xyz[b], P * 3 * sizeof(float),
                                       cudaMemcpyDeviceToHost, e->copy_out));
        SLR_CHECK_CUDA(cudaMemcpyAsync(h_valid + (size_t)s * P, e->d_stage_valid[b], P, cudaMemcpyDeviceToHost,
                                       e->copy_out));
        if (h_match_k)
            SLR_CHECK_CUDA(cudaMemcpyAsync(h_match_k + (size_t)s * P, e->d_stage_k[b], P * sizeof(int32_t),
                                           cudaMemcpyDeviceToHost, e->copy_out));
        if (color && h_color)
            SLR_CHECK_CUDA(cudaMemcpyAsync(h_color + (size_t)s * P, e->d_stage_color[b], P, cudaMemcpyDeviceToHost,
                                           e->copy_out));
        SLR_CHECK_CUDA(cudaEventRecord(e->ev_out[b], e->copy_out));
    }
    SLR_CHECK_CUDA(cudaMemcpyAsync(e->h_counter, e->d_counter, sizeof(unsigned long long), cudaMemcpyDeviceToHost, cs));
    SLR_CHECK_CUDA(cudaStreamSynchronize(cs));
    SLR_CHECK_CUDA(cudaStreamSynchronize(e->copy_out));
    SLR_CHECK_CUDA(cudaStreamSynchronize(e->copy_in));
    if (h_n_points) *h_n_points = *e->h_counter;
    return SLR_OK;
}

// On an error in the middle of a batch, asynchronous copies from h_stack / into the caller's output buffers may still
// be in flight: drain all three streams before the caller gets the (first) error back and frees or reuses them.
template <typename LaunchFn>
static slr_status host_pipeline(slr_engine *e, const uint8_t *h_stack, size_t in_bytes, int batch, bool color,
                                float *h_xyz, uint8_t *h_valid, int32_t *h_match_k, uint8_t *h_color,
                                unsigned long long *h_n_points, LaunchFn launch, bool rectify_in_launch = false)
{
    const slr_status st = host_pipeline_body(e, h_stack, in_bytes, batch, color, h_xyz, h_valid, h_match_k, h_color,
                                             h_n_points, launch, rectify_in_launch);
    if (st != SLR_OK) {
        char first[512];
        snprintf(first, sizeof(first), "%s", slr_last_error());
        cudaStreamSynchronize(e->copy_in);
        cudaStreamSynchronize(e->stream);
        cudaStreamSynchronize(e->copy_out);
        cudaGetLastError();
        slr_set_error("%s", first);
    }
    return st;
}

extern "C" slr_status slr_run_mf_host(slr_engine *e, const uint8_t *h_stack, int batch, int F, int S,
                                      int black_thr, int mode, float *h_xyz, uint8_t *h_valid,
                                      int32_t *h_match_k, unsigned long long *h_n_points)
{
    SLR_ENTER(e);
    SLR_REQUIRE(h_stack && h_xyz && h_valid && batch > 0, "slr_run_mf_host: bad argument");
    SLR_REQUIRE(F >= 1 && S >= 3 && F * S <= 128, "slr_run_mf_host: bad F/S");
    if (!e->calib_set) {
        slr_set_error("slr_run_mf_host: call slr_set_calib first");
        return SLR_ERR_STATE;
    }
    const size_t in_bytes = (size_t)2 * (2 + F * S) * e->W * e->H;
    // raw camera images (MFReconstruct::loadCamImgs rectifies, never stretches: Duke/mfreconstruct.cpp:111-146): the
    // rectification runs inside the fused kernel's stage fill, so a raw scan is ONE kernel like a rectified one
    const bool raw = e->host_input_raw && !e->auto_contrast;
    return host_pipeline(e, h_stack, in_bytes, batch, false, h_xyz, h_valid, h_match_k, nullptr, h_n_points,
                         [&](uint8_t *d_in, float *d_xyz, uint8_t *d_valid, int32_t *d_k, uint8_t *) {
                             if (raw)
                                 return slr_launch_fused_mf_raw(e, d_in, 1, F, S, black_thr, mode, d_xyz, d_valid,
                                                                h_match_k ? d_k : nullptr, e->d_counter);
                             return slr_launch_fused_mf(e, d_in, 1, F, S, black_thr, mode, d_xyz, d_valid,
                                                        h_match_k ? d_k : nullptr, e->d_counter);
                         }, raw);
}

extern "C" slr_status slr_merge_scans(slr_engine *e, const float *d_xyz_all, const uint8_t *d_valid_all, int n_scans,
                                      const float *h_rigid3x4, const uint8_t *h_has_rigid, float *d_points,
                                      long long *d_source, unsigned long long *d_count)
{
    SLR_ENTER(e);
    SLR_REQUIRE(d_xyz_all && d_valid_all && d_points && d_count && n_scans > 0, "slr_merge_scans: bad argument");
    return slr_launch_merge(e, d_xyz_all, d_valid_all, n_scans, h_rigid3x4, h_has_rigid, d_points, d_source, d_count);
}

// ------------------------------------------------------------------------------------------------
// PNG ingest: images arrive one by one, as their host threads finish inflating them
// ------------------------------------------------------------------------------------------------
extern "C" slr_status slr_ingest_begin(slr_engine *e, int n_images)
{
    SLR_ENTER(e);
    SLR_REQUIRE(n_images > 0 && n_images % 2 == 0 && n_images <= 2 * 130, "slr_ingest_begin: bad image count %d", n_images);
    SLR_REQUIRE(e->W % 4 == 0, "slr_ingest_begin: the GPU unfilter needs a width that is a multiple of 4");
    const size_t P = (size_t)e->W * e->H, FB = (size_t)(e->W + 1) * e->H;
    const slr_status st = ensure_stage(e, (size_t)n_images * P, false);
    if (st != SLR_OK) return st;
    if (e->ingest_bytes < (size_t)n_images * FB) {
        if (e->d_ingest) SLR_CHECK_CUDA(cudaFree(e->d_ingest));
        e->d_ingest = nullptr;
        e->ingest_bytes = 0;
        SLR_CHECK_CUDA(cudaMalloc(&e->d_ingest, (size_t)n_images * FB));
        e->ingest_bytes = (size_t)n_images * FB;
    }
    if (!e->ev_ingest) SLR_CHECK_CUDA(cudaEventCreateWithFlags(&e->ev_ingest, cudaEventDisableTiming));
    // the stage buffer may still be read by the kernels of an earlier call on the engine's stream
    SLR_CHECK_CUDA(cudaEventRecord(e->ev_ingest, e->stream));
    SLR_CHECK_CUDA(cudaStreamWaitEvent(e->copy_in, e->ev_ingest, 0));
    e->ingest_images = n_images;
    return SLR_OK;
}

extern "C" slr_status slr_ingest_image(slr_engine *e, int index, const uint8_t *h_data, int filtered, int has_up_rows)
{
    SLR_ENTER(e);
    SLR_REQUIRE(h_data && index >= 0 && index < e->ingest_images, "slr_ingest_image: bad argument (slr_ingest_begin first)");
    const size_t P = (size_t)e->W * e->H, FB = (size_t)(e->W + 1) * e->H;
    uint8_t *plane = e->d_stage_in[0] + (size_t)index * P;
    if (!filtered) {
        SLR_CHECK_CUDA(cudaMemcpyAsync(plane, h_data, P, cudaMemcpyHostToDevice, e->copy_in));
        return SLR_OK;
    }
    uint8_t *f = e->d_ingest + (size_t)index * FB;
    SLR_CHECK_CUDA(cudaMemcpyAsync(f, h_data, FB, cudaMemcpyHostToDevice, e->copy_in));
    return slr_launch_png_unfilter(e, e->copy_in, f, plane, has_up_rows != 0);
}

extern "C" slr_status slr_ingest_abort(slr_engine *e)
{
    SLR_ENTER(e);
    e->ingest_images = 0;
    SLR_CHECK_CUDA(cudaStreamSynchronize(e->copy_in));   // the caller's pinned images are no longer read
    return SLR_OK;
}

extern "C" slr_status slr_png_unfilter(slr_engine *e, const uint8_t *d_scanlines, uint8_t *d_pixels, int has_up_rows)
{
    SLR_ENTER(e);
    SLR_REQUIRE(d_scanlines && d_pixels && e->W % 4 == 0 && ((uintptr_t)d_pixels & 3) == 0,
                "slr_png_unfilter: bad argument (width and output must be multiples of 4)");
    return slr_launch_png_unfilter(e, e->stream, d_scanlines, d_pixels, has_up_rows != 0);
}

// the tail shared by slr_run_mf_ingested / slr_run_ge_ingested: the cloud of d_stage_xyz[0] / d_stage_valid[0] (and the
// colour plane) back to the host in the requested forms; synchronises, error or not (the caller's buffers must be quiet)
static slr_status ingested_outputs(slr_engine *e, slr_status st, bool color, int scan_w, int scan_h, float *h_sum, uint8_t *h_cnt,
                                   uint8_t *h_cell_gray, float *h_xyz, uint8_t *h_valid, uint8_t *h_color,
                                   unsigned long long *h_n_points)
{
    const size_t P = (size_t)e->W * e->H;
    cudaStream_t cs = e->stream;
    while (st == SLR_OK) {
        if (h_sum && h_cnt) {
            const size_t cells = (size_t)scan_w * scan_h;
            if (e->cloud_cells < cells) {
                cudaFree(e->d_cloud_sum);
                cudaFree(e->d_cloud_cnt);
                cudaFree(e->d_cloud_gray);
                e->d_cloud_sum = nullptr;
                e->d_cloud_cnt = e->d_cloud_gray = nullptr;
                e->cloud_cells = 0;
                if (cudaMalloc(&e->d_cloud_sum, cells * 3 * sizeof(float)) != cudaSuccess || cudaMalloc(&e->d_cloud_cnt, cells) != cudaSuccess ||
                    cudaMalloc(&e->d_cloud_gray, cells) != cudaSuccess) {
                    slr_set_error("slr_run_*_ingested: out of device memory");
                    st = SLR_ERR_NOMEM;
                    break;
                }
                e->cloud_cells = cells;
            }
            const bool cg = color && h_cell_gray;
            st = slr_launch_cloud_image(e, e->d_stage_xyz[0], e->d_stage_valid[0], cg ? e->d_stage_color[0] : nullptr, scan_w, scan_h,
                                        e->d_cloud_sum, e->d_cloud_cnt, cg ? e->d_cloud_gray : nullptr);
            if (st != SLR_OK) break;
            cudaMemcpyAsync(h_sum, e->d_cloud_sum, cells * 3 * sizeof(float), cudaMemcpyDeviceToHost, cs);
            cudaMemcpyAsync(h_cnt, e->d_cloud_cnt, cells, cudaMemcpyDeviceToHost, cs);
            if (cg) cudaMemcpyAsync(h_cell_gray, e->d_cloud_gray, cells, cudaMemcpyDeviceToHost, cs);
        }
        if (h_xyz && h_valid) {
            cudaMemcpyAsync(h_xyz, e->d_stage_xyz[0], P * 3 * sizeof(float), cudaMemcpyDeviceToHost, cs);
            cudaMemcpyAsync(h_valid, e->d_stage_valid[0], P, cudaMemcpyDeviceToHost, cs);
            if (color && h_color) cudaMemcpyAsync(h_color, e->d_stage_color[0], P, cudaMemcpyDeviceToHost, cs);
        }
        cudaMemcpyAsync(e->h_counter, e->d_counter, sizeof(unsigned long long), cudaMemcpyDeviceToHost, cs);
        break;
    }
    const cudaError_t ce = cudaStreamSynchronize(cs);
    cudaStreamSynchronize(e->copy_in);
    e->ingest_images = 0;
    if (st != SLR_OK) return st;
    SLR_CHECK_CUDA(ce);
    if (h_n_points) *h_n_points = *e->h_counter;
    return SLR_OK;
}

// ingested images are complete on the device once copy_in's work so far is: make the engine's stream wait for it
static slr_status ingested_ready(slr_engine *e)
{
    if (cudaEventRecord(e->ev_ingest, e->copy_in) != cudaSuccess || cudaStreamWaitEvent(e->stream, e->ev_ingest, 0) != cudaSuccess ||
        cudaMemsetAsync(e->d_counter, 0, sizeof(unsigned long long), e->stream) != cudaSuccess) {
        slr_set_error("slr_run_*_ingested: stream ordering failed");
        return SLR_ERR_CUDA;
    }
    return SLR_OK;
}

extern "C" slr_status slr_run_mf_ingested(slr_engine *e, int F, int S, int black_thr, int mode, int scan_w, int scan_h,
                                          float *h_sum, uint8_t *h_cnt, float *h_xyz, uint8_t *h_valid,
                                          unsigned long long *h_n_points)
{
    SLR_ENTER(e);
    SLR_REQUIRE(F >= 1 && S >= 3 && e->ingest_images == 2 * (2 + F * S), "slr_run_mf_ingested: %d images ingested, the stack needs %d",
                e->ingest_images, 2 * (2 + F * S));
    SLR_REQUIRE((h_sum && h_cnt && scan_w > 0 && scan_h > 0) || (h_xyz && h_valid), "slr_run_mf_ingested: no output buffer");
    if (!e->calib_set) {
        slr_set_error("slr_run_mf_ingested: call slr_set_calib first");
        return SLR_ERR_STATE;
    }
    slr_status st = ingested_ready(e);
    if (st == SLR_OK) {
        if (e->host_input_raw)
            st = slr_launch_fused_mf_raw(e, e->d_stage_in[0], 1, F, S, black_thr, mode, e->d_stage_xyz[0], e->d_stage_valid[0],
                                         nullptr, e->d_counter);
        else
            st = slr_launch_fused_mf(e, e->d_stage_in[0], 1, F, S, black_thr, mode, e->d_stage_xyz[0], e->d_stage_valid[0],
                                     nullptr, e->d_counter);
    }
    return ingested_outputs(e, st, false, scan_w, scan_h, h_sum, h_cnt, nullptr, h_xyz, h_valid, nullptr, h_n_points);
}

extern "C" slr_status slr_run_ge_ingested(slr_engine *e, int nbits_col, int black_thr, int white_thr, int code_w, int have_color,
                                          int scan_w, int scan_h, float *h_sum, uint8_t *h_cnt, uint8_t *h_cell_gray,
                                          float *h_xyz, uint8_t *h_valid, uint8_t *h_color, unsigned long long *h_n_points)
{
    SLR_ENTER(e);
    SLR_REQUIRE(nbits_col >= 1 && nbits_col <= 16 && e->ingest_images == 2 * (2 + 2 * nbits_col),
                "slr_run_ge_ingested: %d images ingested, the stack needs %d", e->ingest_images, 2 * (2 + 2 * nbits_col));
    SLR_REQUIRE((h_sum && h_cnt && scan_w > 0 && scan_h > 0) || (h_xyz && h_valid), "slr_run_ge_ingested: no output buffer");
    if (!e->calib_set) {
        slr_set_error("slr_run_ge_ingested: call slr_set_calib first");
        return SLR_ERR_STATE;
    }
    const size_t P = (size_t)e->W * e->H, in_bytes = (size_t)e->ingest_images * P;
    slr_status st = ensure_stage(e, in_bytes, have_color != 0);   // colour buffers, the rectified copy
    if (st == SLR_OK) st = ingested_ready(e);
    uint8_t *d_in = e->d_stage_in[0];
    if (st == SLR_OK && e->host_input_raw) {   // as Reconstruct::loadCamImgs: rectify, then (optionally) stretch
        st = slr_launch_rectify(e, d_in, 1, e->ingest_images / 2, e->d_stage_rect[0]);
        d_in = e->d_stage_rect[0];
    }
    if (st == SLR_OK && e->auto_contrast) st = slr_launch_auto_contrast(e, d_in, e->ingest_images);
    if (st == SLR_OK)
        st = slr_launch_fused_ge(e, d_in, 1, nbits_col, black_thr, white_thr, code_w, have_color, e->d_stage_xyz[0],
                                 e->d_stage_valid[0], nullptr, have_color ? e->d_stage_color[0] : nullptr, e->d_counter);
    return ingested_outputs(e, st, have_color != 0, scan_w, scan_h, h_sum, h_cnt, h_cell_gray, h_xyz, h_valid, h_color, h_n_points);
}

extern "C" slr_status slr_run_ge_host(slr_engine *e, const uint8_t *h_stack, int batch, int nbits_col,
                                      int black_thr, int white_thr, int scan_w, int have_color, float *h_xyz,
                                      uint8_t *h_valid, int32_t *h_match_k, uint8_t *h_color,
                                      unsigned long long *h_n_points)
{
    SLR_ENTER(e);
    SLR_REQUIRE(h_stack && h_xyz && h_valid && batch > 0, "slr_run_ge_host: bad argument");
    SLR_REQUIRE(nbits_col >= 1 && nbits_col <= 16, "slr_run_ge_host: bad nbits_col %d", nbits_col);
    SLR_REQUIRE(!have_color || h_color, "slr_run_ge_host: have_color set but h_color is NULL");
    if (!e->calib_set) {
        slr_set_error("slr_run_ge_host: call slr_set_calib first");
        return SLR_ERR_STATE;
    }
    const size_t in_bytes = (size_t)2 * (2 + 2 * nbits_col) * e->W * e->H;
    return host_pipeline(e, h_stack, in_bytes, batch, have_color != 0, h_xyz, h_valid, h_match_k, h_color,
                         h_n_points,
                         [&](uint8_t *d_in, float *d_xyz, uint8_t *d_valid, int32_t *d_k, uint8_t *d_col) {
                             return slr_launch_fused_ge(e, d_in, 1, nbits_col, black_thr, white_thr, scan_w,
                                                        have_color, d_xyz, d_valid, h_match_k ? d_k : nullptr,
                                                        have_color ? d_col : nullptr, e->d_counter);
                         });
}

// Reconstruct::runReconstruction minus image IO (Gray-only, un-rectified): K2 with row codes + K3c, one scan at a
// time, synchronous staging (this path is latency bound and not on the north-star bench).
extern "C" slr_status slr_run_gray_host(slr_engine *e, const uint8_t *h_stack, int batch, int nbits_col, int nbits_row,
                                        int black_thr, int white_thr, int scan_w, int scan_h, float *h_sum,
                                        uint8_t *h_cnt, unsigned long long *h_n_cells)
{
    SLR_ENTER(e);
    SLR_REQUIRE(h_stack && h_sum && h_cnt && batch > 0, "slr_run_gray_host: bad argument");
    SLR_REQUIRE(nbits_col >= 1 && nbits_col <= 16 && nbits_row >= 1 && nbits_row <= 16 && scan_w > 0 && scan_h > 0,
                "slr_run_gray_host: bad bit counts / scan size");
    if (!e->calib_set) {
        slr_set_error("slr_run_gray_host: call slr_set_calib first");
        return SLR_ERR_STATE;
    }
    const size_t P = (size_t)e->W * e->H, ncell = (size_t)scan_w * scan_h;
    const size_t N = (size_t)(2 + 2 * nbits_col + 2 * nbits_row);
    // engine-owned scratch, grown on demand and kept for the next call
    const size_t want[6] = {2 * N * P, 2 * P, 2 * P * sizeof(int32_t), 2 * P * sizeof(int32_t), ncell * 3 * sizeof(float), ncell};
    for (int k = 0; k < 6; k++) {
        if (e->gray_bytes[k] >= want[k]) continue;
        cudaFree(e->d_gray[k]);
        e->d_gray[k] = nullptr;
        e->gray_bytes[k] = 0;
        SLR_CHECK_CUDA(cudaMalloc(&e->d_gray[k], want[k]));
        e->gray_bytes[k] = want[k];
    }
    uint8_t *d_stack = (uint8_t *)e->d_gray[0], *d_mask = (uint8_t *)e->d_gray[1], *d_cnt = (uint8_t *)e->d_gray[5];
    int32_t *d_col = (int32_t *)e->d_gray[2], *d_row = (int32_t *)e->d_gray[3];
    float *d_sum = (float *)e->d_gray[4];
    slr_status st = SLR_OK;
    cudaError_t ce = cudaMemsetAsync(e->d_counter, 0, sizeof(unsigned long long), e->stream);
    for (int b = 0; b < batch && ce == cudaSuccess && st == SLR_OK; b++) {
        // one scan at a time: upload, decode, triangulate, download on the engine's stream (stream order makes the
        // scratch reusable without a host synchronisation per scan)
        ce = cudaMemcpyAsync(d_stack, h_stack + (size_t)b * 2 * N * P, 2 * N * P, cudaMemcpyHostToDevice, e->stream);
        if (ce != cudaSuccess) break;
        if (e->auto_contrast) st = slr_launch_auto_contrast(e, d_stack, (int)(2 * N));
        if (st == SLR_OK)
            st = slr_launch_gray_decode(e, d_stack, 2, nbits_col, nbits_row, black_thr, white_thr, scan_w, scan_h, d_col, d_row, d_mask);
        if (st == SLR_OK) st = slr_launch_bucket_triangulate(e, d_col, d_row, d_mask, 1, scan_w, scan_h, d_sum, d_cnt, e->d_counter);
        if (st != SLR_OK) break;
        ce = cudaMemcpyAsync(h_sum + (size_t)b * ncell * 3, d_sum, ncell * 3 * sizeof(float), cudaMemcpyDeviceToHost, e->stream);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(h_cnt + (size_t)b * ncell, d_cnt, ncell, cudaMemcpyDeviceToHost, e->stream);
    }
    if (ce == cudaSuccess && st == SLR_OK)
        ce = cudaMemcpyAsync(e->h_counter, e->d_counter, sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream);
    const cudaError_t cs = cudaStreamSynchronize(e->stream);   // also on errors: no copy may outlive the call
    if (ce == cudaSuccess) ce = cs;
    if (ce == cudaSuccess && st == SLR_OK && h_n_cells) *h_n_cells = *e->h_counter;
    if (ce != cudaSuccess) {
        slr_set_error("slr_run_gray_host: %s", cudaGetErrorString(ce));
        return SLR_ERR_CUDA;
    }
    return st;
}

// ------------------------------------------------------------------------------------------------
// un-fused pipelines (first implementation of the fused entry points; kept as the fallback for
// shapes the fused kernels do not cover).  Chunked by max_batch through the engine's scratch.
// ------------------------------------------------------------------------------------------------
slr_status slr_unfused_mf(slr_engine *e, const uint8_t *d_stack, int batch, int F, int S, int black_thr, int mode,
                          float *d_xyz, uint8_t *d_valid, int32_t *d_match_k, unsigned long long *d_n_points)
{
    slr_status st = ensure_scratch(e, true, false);
    if (st != SLR_OK) return st;
    const size_t P = (size_t)e->W * e->H;
    const size_t N = (size_t)(2 + F * S);
    for (int b0 = 0; b0 < batch; b0 += e->max_batch) {
        const int nb = (batch - b0 < e->max_batch) ? batch - b0 : e->max_batch;
        st = slr_launch_mf_decode(e, d_stack + (size_t)b0 * 2 * N * P, nb * 2, F, S, black_thr, mode, e->d_phase,
                                  e->d_mask);
        if (st != SLR_OK) return st;
        st = slr_launch_match_phase(e, e->d_phase, e->d_mask, nb, d_xyz + (size_t)b0 * P * 3, d_valid + (size_t)b0 * P,
                                    d_match_k ? d_match_k + (size_t)b0 * P : nullptr, d_n_points);
        if (st != SLR_OK) return st;
    }
    return SLR_OK;
}

slr_status slr_unfused_ge(slr_engine *e, const uint8_t *d_stack, int batch, int nbits_col, int black_thr,
                          int white_thr, int scan_w, int have_color, float *d_xyz, uint8_t *d_valid,
                          int32_t *d_match_k, uint8_t *d_color, unsigned long long *d_n_points)
{
    slr_status st = ensure_scratch(e, false, true);
    if (st != SLR_OK) return st;
    const size_t P = (size_t)e->W * e->H;
    const size_t N = (size_t)(2 + 2 * nbits_col);
    for (int b0 = 0; b0 < batch; b0 += e->max_batch) {
        const int nb = (batch - b0 < e->max_batch) ? batch - b0 : e->max_batch;
        const uint8_t *stk = d_stack + (size_t)b0 * 2 * N * P;
        st = slr_launch_gray_decode(e, stk, nb * 2, nbits_col, 0, black_thr, white_thr, scan_w, 0, e->d_code, nullptr,
                                    e->d_mask);
        if (st != SLR_OK) return st;
        // white image of view v is plane 0 of that view: stride between views = N*P bytes
        st = slr_launch_match_code(e, e->d_code, e->d_mask, nb, have_color ? stk : nullptr, N * P,
                                   d_xyz + (size_t)b0 * P * 3, d_valid + (size_t)b0 * P,
                                   d_match_k ? d_match_k + (size_t)b0 * P : nullptr,
                                   (have_color && d_color) ? d_color + (size_t)b0 * P : nullptr, d_n_points);
        if (st != SLR_OK) return st;
    }
    return SLR_OK;
}
