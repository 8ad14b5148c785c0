/*
 * slr_b200.h — C ABI of libslr_b200.so, the B200 (sm_100a) structured-light
 * decode + stereo-match + triangulate engine.
 *
 * This is the drop-in boundary for the reference's reconstruction hot path
 * (DrawZeroPoint/Structure-Light-Reconstructor @ 4a79a80).  The reference has no FFI layer:
 * its boundary is the C++ class surface `Reconstruct` / `MFReconstruct` (Duke/reconstruct.h:14-42,
 * Duke/mfreconstruct.h:12-20) that MainWindow::startreconstruct calls (Duke/mainwindow.cpp:562-652).
 * The Qt-free facade classes in structure-light-reconstructor_b200/facade/ keep those names and
 * call the entry points below; INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions
 *  - plain C: opaque handle, pointers and sizes only; no torch / OpenCV / Qt types.
 *  - every entry point returns slr_status; slr_last_error() gives the text for the calling thread.
 *  - one engine per GPU.  Calls are stream-ordered on the engine's stream (slr_set_stream) and
 *    asynchronous unless stated; an engine is thread-compatible, not thread-safe.
 *  - "d_" pointers are device memory on the engine's GPU, "h_" pointers are host memory.
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Data layout (HBM):
 *   image stacks   uint8  [batch][cam=2][N][H][W]   (cam 0 = left, 1 = right; rectified for the
 *                                                   MF / Gray-EPI paths, raw for Gray-only)
 *     MF   stack:  N = 2 + F*S; [0]=white, [1]=black, [2 + S*f + s] = frequency f, shift s
 *                  (Duke/multifrequency.cpp:16-17,30; Duke/mfreconstruct.cpp:239-242)
 *     Gray stack:  N = 2 + 2*nbits_col (+ 2*nbits_row); [0]=white, [1]=black,
 *                  column bit c (MSB first) at [2+2c] (pattern) and [3+2c] (inverse), row bits after
 *                  (Duke/graycodes.cpp:63-110; Duke/reconstruct.cpp:387-400, 349-360)
 *   phase          float  [batch][2][H][W]   NaN where the pixel carries no phase
 *   code           int32  [batch][2][H][W]   -1 where masked
 *   mask / valid   uint8  same shape, 0/1
 *   xyz            float  [batch][H][W][3]   NaN where no point (indexed by the LEFT pixel)
 *   match_k        int32  [batch][H][W]      matched right column, -1 if none
 */
#ifndef SLR_B200_H
#define SLR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SLR_API __attribute__((visibility("default")))
#else
#define SLR_API
#endif

typedef struct slr_engine slr_engine;

typedef enum {
    SLR_OK = 0,
    SLR_ERR_INVALID = 1,     /* bad argument / unsupported shape */
    SLR_ERR_CUDA = 2,        /* CUDA runtime error (or no device: there is no CPU fallback) */
    SLR_ERR_STATE = 3,       /* call order (e.g. calibration not set) */
    SLR_ERR_NOMEM = 4
} slr_status;

/* Decode arithmetic.  STRICT reproduces Duke/mfreconstruct.cpp:231-269 bit for bit (integer-division
 * atan, PI = 3.1416f) and requires F=3, S=4.  CORRECTED is the physically
 * meaningful atan2 + heterodyne cascade (no reference counterpart). */
typedef enum { SLR_MODE_STRICT = 0, SLR_MODE_CORRECTED = 1 } slr_mode;

/* The fields of the reference's VirtualCamera that the path reads (Duke/virtualcamera.h:27-37;
 * all CV_32F there). */
typedef struct {
    float fc[2];   /* VirtualCamera::fc   */
    float cc[2];   /* VirtualCamera::cc   */
    float dist[5]; /* VirtualCamera::distortion (k1 k2 p1 p2 k3; k3 unused, Duke/utilities.cpp:66) */
    float R[9];    /* VirtualCamera::rotationMatrix, row-major */
    float t[3];    /* VirtualCamera::translationVector */
} slr_camera;

/* ---- engine --------------------------------------------------------------------------------- */

/* One engine per GPU.  width/height = camera image size (Reconstruct::getParameters camw/camh,
 * Duke/reconstruct.cpp:615-621); max_batch = scans per call the host-buffer entry points stage.
 * Any width is accepted, as in the reference: the match kernels move rows in 16-byte multiples (TMA), so widths that
 * are not a multiple of 16 run the same kernels over zero-padded device copies (one extra pass over the data);
 * multiples of 16 with 16-byte aligned device pointers take the fused single-kernel paths. */
SLR_API slr_status slr_create(slr_engine **out, int device, int width, int height, int max_batch);
SLR_API slr_status slr_destroy(slr_engine *e);
/* Run on a caller-owned cudaStream_t (e.g. torch's current stream).  As in the CUDA API, NULL is the
 * legacy default stream; SLR_STREAM_OWN selects the engine's own non-blocking stream, which is also
 * the state after slr_create. */
#define SLR_STREAM_OWN ((void *)(intptr_t)-1)
SLR_API slr_status slr_set_stream(slr_engine *e, void *cuda_stream);
SLR_API slr_status slr_synchronize(slr_engine *e);
/* Replaces stereoRect::Q (Duke/stereorect.h:21), MFReconstruct/Reconstruct::cameras[2]
 * (Duke/reconstruct.h:24) and the optional scan/transfer_mat<sn>.txt 3x4 matrix
 * (Duke/mfreconstruct.cpp:276-282; NULL when scanSN == 0).  Also precomputes the per-pixel
 * Utilities::undistortPoints maps (Duke/utilities.cpp:58-94) used by the MF emitter. */
SLR_API slr_status slr_set_calib(slr_engine *e, const slr_camera cams[2], const double Q[16],
                                 const float *rigid3x4);
SLR_API const char *slr_last_error(void);
SLR_API const char *slr_version(void);
/* Host-only self-check of the strict-mode lookup tables the MF kernels decode with (they replace the branches of
 * Duke/mfreconstruct.cpp:246-261: C++ int division inside atan, PI = 3.1416f).  Builds the tables as an engine does and
 * compares the lookup with the branch form for every a = G4-G2, b = G1-G3 in [-255, 255]; h_fx (NULL, or 511 x 511
 * int32) receives the wrapped phases [b + 255][a + 255] in units of 2^-24 (INT32_MIN: the pixel is dropped, :254). */
SLR_API slr_status slr_strict_tables_check(int32_t *h_fx);

/* ---- pattern synthesis (host; rows a1, a2 of SURVEY.md §8) ------------------------------------- */
/* GrayCodes::calNumOfImgs, Duke/graycodes.cpp:22-30 */
SLR_API int slr_gray_num_bits(int n);
SLR_API int slr_gray_num_imgs(int scan_w, int scan_h, int use_epi);
/* GrayCodes::generateGrays, Duke/graycodes.cpp:55-114.  h_out = [nimgs][H][W] */
SLR_API slr_status slr_generate_gray_patterns(uint8_t *h_out, int W, int H, int use_epi);
/* MultiFrequency::generateMutiFreq, Duke/multifrequency.cpp:14-33.  h_out = [14][projH][projW] */
SLR_API slr_status slr_generate_mf_patterns(uint8_t *h_out, int projW, int projH);

/* ---- K0: rectification on load ------------------------------------------------ */
/* stereoRect::calParameters' maps: cv::initUndistortRectifyMap(..., CV_16SC2, map1, map2), Duke/stereorect.cpp:42-43.
 * h_map1 = int16 [2][H][W][2] (integer source x, y), h_map2 = uint16 [2][H][W] (5+5 fractional bits). */
SLR_API slr_status slr_set_rectify_maps(slr_engine *e, const int16_t *h_map1, const uint16_t *h_map2);
/* stereoRect::doStereoRectify == cv::remap(INTER_LINEAR) on every image of both cameras' stacks,
 * Duke/stereorect.cpp:26-34.  d_raw, d_out = [batch][2][n_images][H][W]; not in place. */
SLR_API slr_status slr_rectify_stack(slr_engine *e, const uint8_t *d_raw, int batch, int n_images, uint8_t *d_out);
/* raw != 0: the host-buffer entry points (slr_run_*_host) take RAW camera stacks and rectify them on the GPU
 * first, as MFReconstruct::loadCamImgs / Reconstruct::loadCamImgs do on the CPU. */
SLR_API slr_status slr_set_host_input_raw(slr_engine *e, int raw);

/* ---- K1: multi-frequency phase decode -------------------------------------------------------- */
/* MFReconstruct::computeShadows + decodePatterns + getPhase, Duke/mfreconstruct.cpp:190-269. */
SLR_API slr_status slr_mf_decode(slr_engine *e, const uint8_t *d_stack, int batch, int F, int S,
                                 int black_thr, int mode, float *d_phase, uint8_t *d_mask);

/* ---- K2: Gray-code decode --------------------------------------------------------------------- */
/* Reconstruct::computeShadows + decodePatterns_GE/getProjPixel_GE (nbits_row == 0) or
 * decodePaterns/getProjPixel (nbits_row > 0) + GrayCodes::grayToDec,
 * Duke/reconstruct.cpp:210-227, 79-97, 381-407, 56-74, 325-370; Duke/graycodes.cpp:116-128.
 * d_row may be NULL when nbits_row == 0. */
SLR_API slr_status slr_gray_decode(slr_engine *e, const uint8_t *d_stack, int batch,
                                   int nbits_col, int nbits_row, int black_thr, int white_thr,
                                   int scan_w, int scan_h,
                                   int32_t *d_col, int32_t *d_row, uint8_t *d_mask);

/* ---- K3a: phase match + Q-matrix triangulation ------------------------------------------------ */
/* MFReconstruct::triangulation, Duke/mfreconstruct.cpp:272-334 (+ Utilities::undistortPoints).
 * d_match_k and d_n_points may be NULL.  *d_n_points (device, one uint64) is ADDED to. */
SLR_API slr_status slr_match_triangulate_phase(slr_engine *e, const float *d_phase, const uint8_t *d_mask,
                                               int batch, float *d_xyz, uint8_t *d_valid,
                                               int32_t *d_match_k, unsigned long long *d_n_points);

/* ---- K3b: Gray-EPI code match + Q-matrix triangulation ---------------------------------------- */
/* Reconstruct::triangulation_ge, Duke/reconstruct.cpp:555-611.  d_white = the [batch][2][H][W]
 * white images when haveColor is on (else NULL, and d_color NULL). */
SLR_API slr_status slr_match_triangulate_code(slr_engine *e, const int32_t *d_col, const uint8_t *d_mask,
                                              int batch, const uint8_t *d_white,
                                              float *d_xyz, uint8_t *d_valid, int32_t *d_match_k,
                                              uint8_t *d_color, unsigned long long *d_n_points);

/* ---- K3c: Gray-only projector-cell bucket triangulation ---------------------------------------- */
/* Reconstruct::decodePaterns bucketing + Reconstruct::triangulation + cam2WorldSpace +
 * Utilities::line_lineIntersection + PointCloudImage::addPoint,
 * Duke/reconstruct.cpp:56-74, 417-481, 310-322; Duke/utilities.cpp:399-425; Duke/pointcloudimage.cpp:86-97.
 * d_sum = float [batch][scan_w*scan_h][3] indexed by the reference's ac(x,y) = x*scan_h + y,
 * d_cnt = uint8 [batch][scan_w*scan_h] (the reference's wrapping u8 point count). */
SLR_API slr_status slr_bucket_triangulate(slr_engine *e, const int32_t *d_col, const int32_t *d_row,
                                          const uint8_t *d_mask, int batch, int scan_w, int scan_h,
                                          float *d_sum, uint8_t *d_cnt, unsigned long long *d_n_cells);

/* ---- fused pipelines --------------------------------------------------------------------------- */
/* MFReconstruct::runReconstruction minus image IO, Duke/mfreconstruct.cpp:160-187: one kernel reads
 * each stack byte once and writes each output byte once (no phase round trip through HBM). */
SLR_API slr_status slr_run_mf(slr_engine *e, const uint8_t *d_stack, int batch, int F, int S,
                              int black_thr, int mode, float *d_xyz, uint8_t *d_valid,
                              int32_t *d_match_k, unsigned long long *d_n_points);
/* The same on RAW (un-rectified) camera stacks: stereoRect::doStereoRectify of every image (Duke/stereorect.cpp:26-34,
 * called per image by MFReconstruct::loadCamImgs, Duke/mfreconstruct.cpp:127-134) happens while the fused kernel fills
 * its shared-memory stage — raw bytes are read once, the rectified images never exist in HBM, one kernel per call
 * (SURVEY.md 8f row N1).  Needs slr_set_rectify_maps.  d_raw_stack = [batch][2][2+F*S][H][W]. */
SLR_API slr_status slr_run_mf_raw(slr_engine *e, const uint8_t *d_raw_stack, int batch, int F, int S,
                                  int black_thr, int mode, float *d_xyz, uint8_t *d_valid,
                                  int32_t *d_match_k, unsigned long long *d_n_points);
/* Reconstruct::runReconstruction_GE minus image IO, Duke/reconstruct.cpp:271-307 (K2 then K3b through engine
 * scratch; integer codes make the intermediate cheap: 5 bytes per pixel). */
SLR_API slr_status slr_run_ge(slr_engine *e, const uint8_t *d_stack, int batch, int nbits_col,
                              int black_thr, int white_thr, int scan_w, int have_color,
                              float *d_xyz, uint8_t *d_valid, int32_t *d_match_k, uint8_t *d_color,
                              unsigned long long *d_n_points);

/* ---- host-buffer entry points (what the facade classes call) ------------------------------------ */
/* Same as slr_run_mf with HOST buffers: stages host->device copies, the kernels and the
 * device->host copy of the cloud on the engine's streams, scan by scan, and returns when h_xyz /
 * h_valid are complete.  batch may exceed max_batch (processed in chunks).  h_match_k may be NULL.
 * Pinned host memory (slr_host_alloc) makes the copies asynchronous. */
SLR_API slr_status slr_run_mf_host(slr_engine *e, const uint8_t *h_stack, int batch, int F, int S,
                                   int black_thr, int mode, float *h_xyz, uint8_t *h_valid,
                                   int32_t *h_match_k, unsigned long long *h_n_points);
SLR_API slr_status slr_run_ge_host(slr_engine *e, const uint8_t *h_stack, int batch, int nbits_col,
                                   int black_thr, int white_thr, int scan_w, int have_color,
                                   float *h_xyz, uint8_t *h_valid, int32_t *h_match_k, uint8_t *h_color,
                                   unsigned long long *h_n_points);
/* ---- PNG ingest ---------------------------------------------------------------
 * MFReconstruct::loadCamImgs' cv::imread(path, 0) of every scan image (Duke/mfreconstruct.cpp:119-134) split where the
 * work stops being sequential: the caller inflates each image's zlib stream on a host thread (one stream per thread) and
 * hands over the bytes exactly as the stream holds them; copy, PNG unfiltering, rectification, decode, match and
 * triangulation run on the GPU while the remaining images are still being inflated.
 *   slr_ingest_begin   start a scan of n_images = 2 * (2 + F*S) images (index = cam * N + i)
 *   slr_ingest_image   one image, from pinned host memory, asynchronously.  filtered != 0: h_data = H scanlines of
 *                      [filter type byte][W filtered bytes] (8-bit grey, filter types 0 / 1 / 2 only; has_up_rows != 0 when
 *                      a row uses type 2); filtered == 0: h_data = H*W finished pixels (other formats, or images with
 *                      Average / Paeth rows, decoded by the caller).  h_data must stay valid until slr_run_mf_ingested
 *                      returns.  Calls may come from any host thread, one at a time.
 *   slr_run_mf_ingested  run the MF pipeline on the ingested scan (raw input + rectification in the fused kernel when
 *                      slr_set_host_input_raw is on) and return the cloud as h_xyz / h_valid (as slr_run_mf_host) and /
 *                      or in the storage layout of the reference's PointCloudImage(scan_w, scan_h) after
 *                      MFReconstruct::triangulation's addPoint(i, j, p) calls (Duke/mfreconstruct.cpp:326,
 *                      Duke/pointcloudimage.cpp:86-97): h_sum = float [scan_h][scan_w][3], h_cnt = uint8 [scan_h][scan_w],
 *                      cell (i_w = image row, j_h = image column), points with i_w >= scan_w or j_h >= scan_h dropped.
 *                      Either output pair may be NULL.  Synchronous. */
/* The GPU unfilter on its own, device pointers, stream-ordered: d_scanlines = H x [type][W bytes] (types 0 / 1 / 2),
 * d_pixels = H x W. */
SLR_API slr_status slr_png_unfilter(slr_engine *e, const uint8_t *d_scanlines, uint8_t *d_pixels, int has_up_rows);
SLR_API slr_status slr_ingest_begin(slr_engine *e, int n_images);
SLR_API slr_status slr_ingest_image(slr_engine *e, int index, const uint8_t *h_data, int filtered, int has_up_rows);
/* Gives up a scan begun with slr_ingest_begin (an image turned out to be missing): waits until no copy reads the
 * caller's buffers any more. */
SLR_API slr_status slr_ingest_abort(slr_engine *e);
SLR_API slr_status slr_run_mf_ingested(slr_engine *e, int F, int S, int black_thr, int mode, int scan_w, int scan_h,
                                       float *h_sum, uint8_t *h_cnt, float *h_xyz, uint8_t *h_valid,
                                       unsigned long long *h_n_points);
/* The Gray-EPI pipeline on an ingested scan of 2 * (2 + 2*nbits_col) images: Reconstruct::loadCamImgs' rectification
 * (slr_set_host_input_raw) and autoContrast (slr_set_auto_contrast) + runReconstruction_GE (Duke/reconstruct.cpp:150-196,
 * 271-307, 555-611).  code_w = the scan width the codes are checked against (slr_run_ge's scan_w); scan_w / scan_h = the
 * PointCloudImage size of the h_sum / h_cnt outputs; h_cell_gray = uint8 [scan_h][scan_w], the colour every cell received
 * (have_color; the reference stores it as a grey triple); h_xyz / h_valid / h_color as slr_run_ge_host.  Synchronous. */
SLR_API slr_status slr_run_ge_ingested(slr_engine *e, int nbits_col, int black_thr, int white_thr, int code_w, int have_color,
                                       int scan_w, int scan_h, float *h_sum, uint8_t *h_cnt, uint8_t *h_cell_gray,
                                       float *h_xyz, uint8_t *h_valid, uint8_t *h_color, unsigned long long *h_n_points);
/* Reconstruct::runReconstruction minus image IO (Gray-only: column + row codes on UN-rectified images, ray-ray
 * triangulation), Duke/reconstruct.cpp:230-265.  h_stack = [batch][2][2+2*nbits_col+2*nbits_row][H][W];
 * h_sum = float [batch][scan_w*scan_h][3], h_cnt = uint8 [batch][scan_w*scan_h], indexed ac(x,y) = x*scan_h + y. */
SLR_API slr_status slr_run_gray_host(slr_engine *e, const uint8_t *h_stack, int batch, int nbits_col, int nbits_row,
                                     int black_thr, int white_thr, int scan_w, int scan_h, float *h_sum,
                                     uint8_t *h_cnt, unsigned long long *h_n_cells);
/* ---- multi-GPU assembly ------------------------------------------------------------- */
/* Scans are independent, so the data path needs no collective (one engine per GPU, each with its own scans).  Where a
 * caller wants every rank's cloud on every GPU (the north star's "single NCCL all-gather of the output point cloud"),
 * slr_allgather assembles them IN PLACE over NVLink: d_xyz_all = float [world*scans_per_rank][H][W][3] and
 * d_valid_all = uint8 [world*scans_per_rank][H][W], of which this rank has already filled block `rank` (pass
 * d_xyz_all + rank*scans_per_rank*H*W*3 as the d_xyz of slr_run_mf / slr_run_ge).  Stream-ordered on the engine's
 * stream.  nccl_comm is an ncclComm_t of the NCCL library in the process (libnccl.so.2 is bound at run time with
 * dlopen; libslr_b200.so has no link-time NCCL dependency) — the caller's own, or one made by the helpers below,
 * which wrap ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy (id128 = 128 bytes, created on one rank and
 * handed to the others by any means). */
SLR_API slr_status slr_allgather(slr_engine *e, void *nccl_comm, int world, int rank, int scans_per_rank,
                                 float *d_xyz_all, uint8_t *d_valid_all);
SLR_API slr_status slr_nccl_unique_id(void *id128);
SLR_API slr_status slr_nccl_comm_create(slr_engine *e, void **comm, int world, int rank, const void *id128);
SLR_API slr_status slr_nccl_comm_destroy(void *comm);

/* The same assembly without a collective call (preferred): every rank maps the assembled cloud buffers of all ranks
 * (CUDA IPC over NVLink / NVSwitch peer access) and registers them as gather targets; slr_run_mf / slr_run_ge then
 * store every output row into this rank's block of EVERY target straight from the kernel's registers, so the
 * transfer overlaps the kernel that produces the data (k_fused_flow's epilogue; other routes copy the finished block
 * to the peers on the engine's stream).  One process per GPU:
 *   slr_peer_alloc   cudaMalloc + IPC handle (64 bytes) for the buffers this rank owns
 *   slr_peer_open    map another rank's buffer from its handle (handles travel by any means, e.g. torch.distributed)
 *   slr_set_gather_targets  d_xyz_all[t] = float [scans][H][W][3], d_valid_all[t] = uint8 [scans][H][W] of target t;
 *                    target 0 MUST be this rank's own buffer; first_scan = index of this rank's first scan in the
 *                    assembled cloud.  With targets set, slr_run_mf / slr_run_ge ignore their d_xyz / d_valid arguments.
 *                    n_targets = 0 switches back.  After the call on every rank completes and the ranks have
 *                    synchronised (any barrier), every target holds the whole cloud.
 * Row bands: an engine of height H_band with slr_set_row_offset(first_row) treats its rows as rows first_row .. of
 * the full image (undistortPoints maps, the Q reprojection of the Gray-EPI path), so ONE scan can be split over N
 * GPUs by horizontal bands (each rank gets the same rows of both cameras; no halo) and assembled with the targets. */
SLR_API slr_status slr_peer_alloc(slr_engine *e, size_t bytes, void **d_ptr, void *ipc_handle64);
SLR_API slr_status slr_peer_open(slr_engine *e, const void *ipc_handle64, void **d_ptr);
SLR_API slr_status slr_peer_close(slr_engine *e, void *d_ptr);
SLR_API slr_status slr_peer_free(slr_engine *e, void *d_ptr);
SLR_API slr_status slr_set_gather_targets(slr_engine *e, int n_targets, float *const *d_xyz_all,
                                          uint8_t *const *d_valid_all, long long first_scan);
SLR_API slr_status slr_set_row_offset(slr_engine *e, int first_row);

/* ---- rigid alignment + multi-scan merge ----------------------------------------
 * DotMatch::calMatrix (Duke/dotmatch.cpp:1180-1324): matched marker positions of two consecutive scans -> transfer matrix.
 *   slr_horn_method    mrpt::scanmatching::HornMethod as the reference calls it (:1238; MRPT 1.2.2 is a third-party
 *                      dependency that is not part of the reference tree — restated from Horn 1987, see k_merge.cu):
 *                      pairs = n x {base xyz, moving xyz} (the layout of :1230-1237), out7 = {tx ty tz qr qx qy qz} of
 *                      the transform that moves `moving` onto `base`; force_unit_scale = 0 is MRPT's default (the
 *                      translation then carries the estimated scale, which *scale_out also returns).  Host function.
 *   slr_register_scan  the whole of calMatrix's arithmetic: Horn -> quaternion -> 3x4 matrix (:1241-1250) chained onto
 *                      the previous scan's accumulated matrix (:1312-1317; prev3x4 = NULL for scanSN == 1) =
 *                      the contents of scan/transfer_mat<sn>.txt, row-major doubles.  Host function.
 *   slr_merge_scans    d_xyz_all = float [n_scans][H][W][3] + d_valid_all (e.g. the assembled cloud of
 *                      slr_set_gather_targets / slr_allgather), h_rigid3x4 = float [n_scans][12] transfer matrices
 *                      (NULL: none; h_has_rigid[s] == 0: scan s is already in the common frame, as scanSN == 0 is) ->
 *                      d_points = float [count][3], every valid point transformed as MFReconstruct::triangulation
 *                      does (Duke/mfreconstruct.cpp:315-323), scans in order, pixels in row-major order; d_source
 *                      (may be NULL) = int64 [count], index scan*H*W + pixel of each point; *d_count (device) = count.
 *                      d_points / d_source must hold n_scans*H*W entries.  Stream-ordered. */
SLR_API slr_status slr_horn_method(const double *pairs, int n, int force_unit_scale, double out7[7], double *scale_out);
SLR_API slr_status slr_register_scan(const double *pairs, int n, const double *prev3x4, double out3x4[12]);
SLR_API slr_status slr_merge_scans(slr_engine *e, const float *d_xyz_all, const uint8_t *d_valid_all, int n_scans,
                                   const float *h_rigid3x4, const uint8_t *h_has_rigid, float *d_points,
                                   long long *d_source, unsigned long long *d_count);

/* ---- mesh indexing ---------------------------------------------------------- */
/* The index passes of MeshCreator::exportPlyMesh / exportObjMesh, Duke/meshcreator.cpp:16-65, 67-166, on a
 * PointCloudImage stored as the reference stores it (Duke/pointcloudimage.cpp:3-13): d_sum = float [h][w][3]
 * coordinate sums, d_count = uint8 [h][w] points per pixel; w, h = PointCloudImage width / height.
 * Traversal order i in [0,w) outer, j in [0,h) inner.  first_vertex = 0 reproduces the PLY numbering (whose first
 * vertex reads as "absent" when faces are formed, meshcreator.cpp:79,100), 1 the OBJ numbering.
 *   d_vertices   float [w*h][3]   getPoint(i,j) of every pixel that has one, in traversal order
 *   d_vertex_src int32 [w*h]      storage element j*w+i each vertex came from (for per-vertex colour); may be NULL
 *   d_faces      int32 [2*w*h][3] vertex numbers of every face, in the reference's output order and orientation
 *   d_counts     uint64 [2]       number of vertices, number of faces */
SLR_API slr_status slr_mesh_index(slr_engine *e, const float *d_sum, const uint8_t *d_count, int w, int h,
                                  int first_vertex, float *d_vertices, int32_t *d_vertex_src, int32_t *d_faces,
                                  unsigned long long *d_counts);
/* Same with HOST buffers (what the facade's MeshCreator calls): h_vertices / h_vertex_src / h_faces must hold the
 * worst case (w*h vertices, 2*w*h faces); h_counts[2] receives the actual numbers. */
SLR_API slr_status slr_mesh_index_host(slr_engine *e, const float *h_sum, const uint8_t *h_count, int w, int h,
                                       int first_vertex, float *h_vertices, int32_t *h_vertex_src, int32_t *h_faces,
                                       unsigned long long *h_counts);
/* Utilities::autoContrast (Duke/utilities.cpp:340-355) as Reconstruct::loadCamImgs applies it to every loaded image
 * when the "auto contrast" setting is on (Duke/reconstruct.cpp:182-183): per image, min/max stretch with OpenCV's
 * saturating 8-bit arithmetic, in place.  d_images = n_images images of width*height bytes.  (The reference indexes
 * channels 1 and 2 of a one-channel image — undefined behaviour; channel 0's arithmetic is what is restated.) */
SLR_API slr_status slr_auto_contrast(slr_engine *e, uint8_t *d_images, int n_images);
/* The host-buffer pipelines (slr_run_ge_host, slr_run_gray_host, slr_run_mf_host) stretch every image after upload
 * (and rectification) when this is on: Reconstruct::getParameters' autocontrast flag. */
SLR_API slr_status slr_set_auto_contrast(slr_engine *e, int on);
SLR_API slr_status slr_host_alloc(void **out, size_t bytes); /* pinned */
SLR_API slr_status slr_host_free(void *p);

/* ---- synthetic scans (bench / smoke inputs) -------------------------------------- */
/* Renders `batch` rectified stereo MF stacks of a plane+bump scene straight into device memory:
 * d_stack = [batch][2][14][H][W].  integer_disparity != 0 makes left/right samples coincide exactly. */
SLR_API slr_status slr_synth_mf(slr_engine *e, uint8_t *d_stack, int batch, int proj_w,
                                unsigned seed, int integer_disparity, float noise_dn);
/* The same scene under F frequencies x S shifts (BASELINE config 5: F = 4, S = 8; frequencies 70, 64, 59, 56, ...):
 * d_stack = [batch][2][2+F*S][H][W], plane [2 + S*f + s].  F = 3, S = 4 gives slr_synth_mf's bytes. */
SLR_API slr_status slr_synth_mf_fs(slr_engine *e, uint8_t *d_stack, int batch, int F, int S, int proj_w,
                                   unsigned seed, int integer_disparity, float noise_dn);
/* d_stack = [batch][2][2+2*nbits_col][H][W] Gray-EPI stacks of the same scene family. */
SLR_API slr_status slr_synth_gray(slr_engine *e, uint8_t *d_stack, int batch, int scan_w,
                                  unsigned seed, int integer_disparity, float noise_dn);

/* Number of kernels this library has launched on this engine since creation (bench accounting). */
SLR_API unsigned long long slr_kernel_launches(const slr_engine *e);

#ifdef __cplusplus
}
#endif
#endif /* SLR_B200_H */
