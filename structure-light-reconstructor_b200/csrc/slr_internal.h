// slr_internal.h — engine state and host-side helpers shared by the .cu files of libslr_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "slr_b200.h"

#define SLR_QNAN_BITS 0x7FC00000u
#define SLR_MAX_TARGETS 8

// Reference constants.  Duke/mfreconstruct.cpp:5 `float PI = 3.1416;`
#define SLR_PI_DEC 3.1416f

struct slr_calib_dev {
    // Q (stereoRect::Q, 4x4 CV_64F) and the optional 3x4 rigid transform, passed to kernels by value.
    double Q[16];
    float rigid[12];
    int has_rigid;
    int q_std;  // Q has the cv::stereoRectify sparsity pattern (see reproject_q)
    int row0;   // image row of the engine's row 0 (row-band partition of a scan over several GPUs; 0 otherwise)
};

struct slr_engine {
    int device = 0;
    int W = 0, H = 0, max_batch = 0;
    int num_sms = 0;
    cudaStream_t own_stream = nullptr;   // created by the engine
    cudaStream_t stream = nullptr;       // the stream work is issued on (own or caller's)
    cudaStream_t copy_in = nullptr, copy_out = nullptr;  // host pipeline streams
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};

    bool calib_set = false;
    slr_camera cams[2];
    slr_calib_dev calib;
    // Utilities::undistortPoints maps, float [H][W]: left x, left y, right x.
    float *d_undist_lx = nullptr, *d_undist_ly = nullptr, *d_undist_rx = nullptr;
    // atan(float(q)) for q in [-255, 255] (index q+255), built on the host with the host libm.
    float *d_atan_lut = nullptr;
    // strict-mode tables of the fused kernel (k_fused.cu: slr_build_strict_tables)
    int32_t *d_ptab = nullptr;      // [SLR_PTAB_SIZE] wrapped phases in units of 2^-24
    uint32_t *d_btab = nullptr;     // [SLR_BTAB_SIZE] reciprocal multiplier + row base per b = G1-G3

    // scratch for the un-fused pipelines / host entry points
    float *d_phase = nullptr;      // [max_batch][2][H][W]
    int32_t *d_code = nullptr;     // [max_batch][2][H][W]
    uint8_t *d_mask = nullptr;     // [max_batch][2][H][W]
    // host pipeline staging (double buffered, one scan each)
    uint8_t *d_stage_in[2] = {nullptr, nullptr};
    size_t stage_in_bytes = 0;
    float *d_stage_xyz[2] = {nullptr, nullptr};
    uint8_t *d_stage_valid[2] = {nullptr, nullptr};
    int32_t *d_stage_k[2] = {nullptr, nullptr};
    uint8_t *d_stage_color[2] = {nullptr, nullptr};
    unsigned long long *d_counter = nullptr;  // device point counter
    unsigned long long *h_counter = nullptr;  // pinned

    // K0 rectification maps (cv::initUndistortRectifyMap, CV_16SC2): [2][H][W] short2 and [2][H][W] u16
    int16_t *d_map1 = nullptr;
    uint16_t *d_map2 = nullptr;
    bool maps_set = false;
    bool host_input_raw = false;     // host entry points rectify the uploaded stacks first
    bool auto_contrast = false;      // host entry points stretch every image (Utilities::autoContrast) before decoding
    void *d_minmax = nullptr;        // K_contrast: int2 {min, max} per image
    int minmax_images = 0;
    // slr_run_gray_host scratch (one scan): stack, mask, col, row, cell sums, cell counts
    void *d_gray[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t gray_bytes[6] = {0, 0, 0, 0, 0, 0};
    uint8_t *d_stage_rect[2] = {nullptr, nullptr};
    size_t stage_rect_bytes = 0;

    // K5 mesh index scratch: vertex numbers + tile sums, and the staging of the host entry point
    int *d_mesh_pn = nullptr, *d_mesh_tiles = nullptr;
    size_t mesh_px = 0;
    float *d_mesh_sum = nullptr, *d_mesh_vertices = nullptr;
    uint8_t *d_mesh_count = nullptr;
    int32_t *d_mesh_src = nullptr, *d_mesh_faces = nullptr;
    unsigned long long *d_mesh_counts = nullptr;
    size_t mesh_host_px = 0;

    void *d_bucket_scratch = nullptr;  // K3c counting-sort scratch (one scan)
    float *d_rays = nullptr;           // K3c: unit ray of every pixel of both cameras [2][H*W][3], per calibration
    unsigned rays_version = 0;
    size_t bucket_scratch_bytes = 0;

    // widths that are not a multiple of 16 (TMA bulk rows need 16-byte rows): the match stages run on a child engine
    // of the padded width over zero-padded copies (slr_engine.cu: padded_*)
    slr_engine *child = nullptr;
    unsigned calib_version = 0, child_calib_version = 0;
    void *d_pad[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t pad_bytes[6] = {0, 0, 0, 0, 0, 0};

    // gather targets (slr_set_gather_targets): the assembled clouds of all ranks, [0] = this GPU's own, the others
    // peer mappings; the MF / GE pipelines write every output row into each of them
    int n_targets = 0;
    float *xyz_t[SLR_MAX_TARGETS] = {};
    uint8_t *valid_t[SLR_MAX_TARGETS] = {};
    long long target_first_scan = 0;
    bool targets_written = false;   // set by a launcher whose kernel stored to all targets itself

    // PNG ingest (k_png.cu): filtered scanlines of the images of one scan as uploaded, the image count, and the
    // PointCloudImage-layout outputs of slr_run_mf_ingested
    uint8_t *d_ingest = nullptr;
    size_t ingest_bytes = 0;
    int ingest_images = 0;
    cudaEvent_t ev_ingest = nullptr;
    float *d_cloud_sum = nullptr;
    uint8_t *d_cloud_cnt = nullptr, *d_cloud_gray = nullptr;
    size_t cloud_cells = 0;

    void *d_merge = nullptr;         // slr_merge_scans: tile offsets + per-scan transforms
    size_t merge_bytes = 0;

    unsigned long long launches = 0;
};

void slr_set_error(const char *fmt, ...);

#define SLR_CHECK_CUDA(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            slr_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return SLR_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define SLR_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            slr_set_error(__VA_ARGS__);   \
            return SLR_ERR_INVALID;       \
        }                                 \
    } while (0)

#define SLR_CHECK_LAUNCH(e)                                  \
    do {                                                     \
        (e)->launches++;                                     \
        SLR_CHECK_CUDA(cudaGetLastError());                  \
    } while (0)

// true when any of the (possibly NULL) pointers is not 16-byte aligned: such calls are staged through aligned copies
template <typename... P>
static inline bool slr_misaligned16(const P *...ptrs)
{
    return ((... | (uintptr_t)ptrs) & 15u) != 0;
}

// kernel launchers (defined in the k*.cu files)
slr_status slr_launch_mf_decode(slr_engine *e, const uint8_t *d_stack, int views, int F, int S,
                                int black_thr, int mode, float *d_phase, uint8_t *d_mask);
slr_status slr_launch_match_phase(slr_engine *e, const float *d_phase, const uint8_t *d_mask, int batch,
                                  float *d_xyz, uint8_t *d_valid, int32_t *d_match_k,
                                  unsigned long long *d_n_points);
slr_status slr_launch_gray_decode(slr_engine *e, const uint8_t *d_stack, int views, int nbits_col,
                                  int nbits_row, int black_thr, int white_thr, int scan_w, int scan_h,
                                  int32_t *d_col, int32_t *d_row, uint8_t *d_mask);
slr_status slr_launch_match_code(slr_engine *e, const int32_t *d_col, const uint8_t *d_mask, int batch,
                                 const uint8_t *d_white, size_t white_view_stride,
                                 float *d_xyz, uint8_t *d_valid, int32_t *d_match_k, uint8_t *d_color,
                                 unsigned long long *d_n_points);
slr_status slr_launch_bucket_triangulate(slr_engine *e, const int32_t *d_col, const int32_t *d_row,
                                         const uint8_t *d_mask, int batch, int scan_w, int scan_h,
                                         float *d_sum, uint8_t *d_cnt, unsigned long long *d_n_cells);
slr_status slr_launch_fused_mf(slr_engine *e, const uint8_t *d_stack, int batch, int F, int S,
                               int black_thr, int mode, float *d_xyz, uint8_t *d_valid,
                               int32_t *d_match_k, unsigned long long *d_n_points);
slr_status slr_launch_fused_mf_raw(slr_engine *e, const uint8_t *d_raw, int batch, int F, int S, int black_thr, int mode,
                                   float *d_xyz, uint8_t *d_valid, int32_t *d_match_k, unsigned long long *d_n_points);
slr_status slr_launch_fused_ge(slr_engine *e, const uint8_t *d_stack, int batch, int nbits_col,
                               int black_thr, int white_thr, int scan_w, int have_color,
                               float *d_xyz, uint8_t *d_valid, int32_t *d_match_k, uint8_t *d_color,
                               unsigned long long *d_n_points);
slr_status slr_launch_undistort_maps(slr_engine *e);
slr_status slr_launch_rectify(slr_engine *e, const uint8_t *d_raw, int batch, int N, uint8_t *d_out);
slr_status slr_build_strict_tables(slr_engine *e);
slr_status slr_launch_merge(slr_engine *e, const float *d_xyz, const uint8_t *d_valid, int n_scans, const float *h_rigid,
                            const uint8_t *h_has_rigid, float *d_points, long long *d_source, unsigned long long *d_count);
slr_status slr_launch_png_unfilter(slr_engine *e, cudaStream_t stream, const uint8_t *d_filtered, uint8_t *d_plane,
                                   bool has_up_rows);
slr_status slr_launch_cloud_image(slr_engine *e, const float *d_xyz, const uint8_t *d_valid, const uint8_t *d_gray, int scan_w,
                                  int scan_h, float *d_sum, uint8_t *d_cnt, uint8_t *d_cell_gray);
slr_status slr_launch_auto_contrast(slr_engine *e, uint8_t *d_images, int n_images);
// padded route for e->W % 16 != 0 (slr_engine.cu); kind: 0 = image stacks (MF), 1 = image stacks (GE),
// 2 = phase + mask rows, 3 = code + mask rows
slr_status slr_padded_run(slr_engine *e, int kind, const void *d_in0, const uint8_t *d_in1, int batch, int planes,
                          int a0, int a1, int a2, int a3, int a4, float *d_xyz, uint8_t *d_valid, int32_t *d_match_k,
                          uint8_t *d_color, unsigned long long *d_n_points);
slr_status slr_launch_mesh_index(slr_engine *e, const float *d_sum, const uint8_t *d_count, int w, int h,
                                 int first_vertex, int *d_pn, int *d_tiles, float *d_vertices, int32_t *d_vertex_src,
                                 int32_t *d_faces, unsigned long long *d_counts);
int slr_mesh_tiles(int w, int h);
slr_status slr_launch_synth_mf(slr_engine *e, uint8_t *d_stack, int batch, int F, int S, int proj_w, unsigned seed,
                               int integer_disparity, float noise_dn);
slr_status slr_launch_synth_gray(slr_engine *e, uint8_t *d_stack, int batch, int scan_w, unsigned seed,
                                 int integer_disparity, float noise_dn);
