/*
 * slr_oracle.h — CPU ORACLE for the structured-light decode + match + triangulate path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.  The product
 * (structure-light-reconstructor_b200/, libslr_b200.so) never links, imports or calls it.
 *
 * What it is: a plain-C restatement of the reference's per-pixel CPU loops
 * (DrawZeroPoint/Structure-Light-Reconstructor @ 4a79a80, Duke sources).  Every function
 * cites the reference file:line it follows.
 *
 * Pin status: the reference ships NO tests, golden vectors, fixtures or sample data
 * (SURVEY.md §4, §8c), so nothing of the reference's own pins this path:
 *   ** parity unpinned by reference fixtures **
 * The pins we do have: (1) known-answer vectors derived by hand from the cited lines
 * (tests/golden/), (2) oracle/_ref — the reference's own hot-path sources compiled
 * unmodified from /root/reference against a small header shim (see oracle/Makefile,
 * oracle/ref_shim/), whose outputs are compared with this restatement in tests/ and
 * committed as fixtures under tests/golden/.
 *
 * Floating point contract (strict mode): IEEE-754 binary32/binary64, round-to-nearest,
 * no FMA contraction, no x87 excess precision (compile with -ffp-contract=off on
 * x86-64/SSE2).  `atan(float)` is the C++ float overload (reference: <math.h> via
 * opencv/cv.h), i.e. atanf.
 */
#ifndef SLR_ORACLE_H
#define SLR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Duke/virtualcamera.h:27-37 — the fields the hot path reads (all CV_32F in the reference). */
typedef struct {
    float fc[2];   /* focal length  (cam_matrix[0][0], [1][1])      virtualcamera.cpp:41-44 */
    float cc[2];   /* principal pt  (cam_matrix[0][2], [1][2])                              */
    float dist[5]; /* k1 k2 p1 p2 k3; k3 is forced to 0 by utilities.cpp:66                */
    float R[9];    /* rotationMatrix, row-major 3x3                                         */
    float t[3];    /* translationVector 3x1                                                 */
} orc_camera;

enum { ORC_MODE_STRICT = 0, ORC_MODE_CORRECTED = 1 };

/* ---- pattern synthesis (a1, a2) ------------------------------------------------------- */
/* Duke/graycodes.cpp:24-25: ceil(log(n)/log(2)) */
int  orc_gray_num_bits(int n);
/* Duke/graycodes.cpp:22-30: 2+2nC (EPI) or 2+2nC+2nR */
int  orc_gray_num_imgs(int scan_w, int scan_h, int use_epi);
/* Duke/graycodes.cpp:55-114. out = [nimgs][H][W] u8 */
void orc_generate_gray(uint8_t *out, int W, int H, int use_epi);
/* Duke/graycodes.cpp:116-128. bits[0] is the MSB. */
int  orc_gray_to_dec(const uint8_t *bits, int n);
/* Duke/multifrequency.cpp:14-33. out = [14][projH][projW] u8 */
void orc_generate_mf(uint8_t *out, int projW, int projH);

/* ---- decode (a3, a4, a5, a7, a8) ------------------------------------------------------ */
/* Duke/mfreconstruct.cpp:190-207 / Duke/reconstruct.cpp:210-227 */
void orc_shadow_mask(const uint8_t *white, const uint8_t *black, int npix, int black_thr, uint8_t *mask);
/* Duke/mfreconstruct.cpp:231-269 for one pixel.  G = 12 gray values, G[4*f+s].
 * Returns 1 and writes *phase, or 0 if the pixel hit the degenerate branch (:254-255). */
int  orc_get_phase_strict(const int *G, float *phase);
/* The per-frequency wrapped phase table of :246-261 as a function of the two differences
 * a = G4-G2, b = G1-G3 (so tests can sweep all 511x511 pairs).  Returns 0 if degenerate. */
int  orc_wrapped_phase_strict(int a, int b, float *P);
/* Duke/mfreconstruct.cpp:190-228 for one camera.  stack = [2+F*S][H][W] u8.
 * phase: float[H*W] (NaN where no phase), mask: u8[H*W] (1 = pixel carries a phase). */
int  orc_mf_decode(const uint8_t *stack, int W, int H, int F, int S, int black_thr, int mode,
                   float *phase, uint8_t *mask);
/* same, pixels split over `nthreads` OpenMP threads (pixels are independent) */
int  orc_mf_decode_mt(const uint8_t *stack, int W, int H, int F, int S, int black_thr, int mode,
                      float *phase, uint8_t *mask, int nthreads);
/* Duke/reconstruct.cpp:210-227 + 79-97 + 381-407 (nbits_row==0, GRAY_EPI) or
 * 56-74 + 325-370 (nbits_row>0, GRAY_ONLY) for one camera.
 * stack = [2+2*nbits_col+2*nbits_row][H][W]. col,row: int32[H*W] (-1 where masked). */
void orc_gray_decode(const uint8_t *stack, int W, int H, int nbits_col, int nbits_row,
                     int black_thr, int white_thr, int scan_w, int scan_h,
                     int32_t *col, int32_t *row, uint8_t *mask);

/* ---- geometry helpers (a11, a12, a13) -------------------------------------------------- */
/* Duke/utilities.cpp:58-94 */
void orc_undistort_point(float px, float py, const orc_camera *cam, float *ox, float *oy);
/* Duke/reconstruct.cpp:310-322 */
void orc_cam2world(const orc_camera *cam, float p[3]);
/* Duke/utilities.cpp:19-28 */
void orc_normalize(float v[3]);
/* Duke/utilities.cpp:399-425; returns 1 if ok */
int  orc_line_line_intersection(const float p1[3], const float v1[3], const float p2[3],
                                const float v2[3], float p[3]);

/* ---- match + triangulate (a6, a9, a10) ------------------------------------------------- */
/* Duke/mfreconstruct.cpp:272-334.  xyz = float[H*W*3] (NaN where no point), valid = u8[H*W],
 * match_k = int32[H*W] (-1 where unmatched; may be NULL).  rigid = 3x4 row-major or NULL
 * (scanSN == 0).  nthreads <= 1: the reference's single-threaded loop order; > 1: same
 * per-row loop, rows split over OpenMP threads (rows are independent).
 * Returns the number of points. */
int64_t orc_mf_triangulate(const float *phL, const uint8_t *mkL, const float *phR, const uint8_t *mkR,
                           int W, int H, const orc_camera *camL, const orc_camera *camR,
                           const double *Q, const float *rigid,
                           float *xyz, uint8_t *valid, int32_t *match_k, int nthreads);
/* Duke/reconstruct.cpp:555-611.  whiteL/whiteR = image 0 of each camera or NULL (haveColor off);
 * color = u8[H*W] gray value or NULL. */
int64_t orc_ge_triangulate(const int32_t *colL, const uint8_t *mkL, const int32_t *colR, const uint8_t *mkR,
                           int W, int H, const double *Q, const float *rigid,
                           const uint8_t *whiteL, const uint8_t *whiteR,
                           float *xyz, uint8_t *valid, int32_t *match_k, uint8_t *color, int nthreads);
/* Duke/reconstruct.cpp:56-74 (bucketing, camera col-major push order) + 417-481 (all-pairs ray-ray
 * midpoints) + Duke/pointcloudimage.cpp:28-37,86-97 (sum / u8 count accumulation).
 * sum = float[scan_w*scan_h*3] indexed [x*scan_h + y] (the reference's ac(x,y)), cnt = u8 same index.
 * Returns the number of projector cells holding at least one point. */
int64_t orc_gray_triangulate(const int32_t *colL, const int32_t *rowL, const uint8_t *mkL,
                             const int32_t *colR, const int32_t *rowR, const uint8_t *mkR,
                             int W, int H, int scan_w, int scan_h,
                             const orc_camera *camL, const orc_camera *camR, const float *rigid,
                             float *sum, uint8_t *cnt);

/* ---- output container (a15) ------------------------------------------------------------- */
/* Duke/pointcloudimage.cpp:86-97 applied to a dense (row, col) cloud as the MF / GE callers do
 * (mfreconstruct.cpp:326, reconstruct.cpp:603: addPoint(row, col, p)):
 * points = float[scan_h*scan_w*3] laid out as the reference's h x w CV_32FC3 Mat, element
 * (j_h, i_w) at [(j_h*scan_w + i_w)*3]; count = u8[scan_h*scan_w].  F7 drop rule included. */
void orc_pointcloud_from_dense(const float *xyz, const uint8_t *valid, int W, int H,
                               int scan_w, int scan_h, float *points, uint8_t *count);

/* Duke/pointcloudimage.cpp:28-37, 86-97 for an arbitrary addPoint(i_w, j_h, p) sequence (u8 count wraps). */
void orc_pointcloud_add(int w, int h, const int32_t *iw, const int32_t *jh, const float *pts, int n,
                        float *points, uint8_t *count);

/* Duke/stereorect.cpp:26-34: cv::remap(INTER_LINEAR) with CV_16SC2 maps (map1 = [H][W][2] int16, map2 = [H][W] u16) */
void orc_remap_linear(const uint8_t *src, int W, int H, const int16_t *map1, const uint16_t *map2, uint8_t *dst);

/* Duke/utilities.cpp:340-355 (Utilities::autoContrast) on one CV_8U image, in place, as Reconstruct::loadCamImgs
 * calls it (Duke/reconstruct.cpp:182-183).  Channel 0 only: the reference indexes bgr[1], bgr[2] of a one-channel
 * split, which is undefined behaviour, so this function has NO reference pin ("parity unpinned", oracle-only). */
void orc_auto_contrast(uint8_t *img, int W, int H);

/* ---- whole-pipeline conveniences used by bench.py's CPU legs ------------------------------ */
/* MF pipeline on one scan: stacks = [2][14][H][W].  Returns points; *n_pixels unused. */
int64_t orc_run_mf(const uint8_t *stacks, int W, int H, int F, int S, int black_thr, int mode,
                   const orc_camera *cams, const double *Q, const float *rigid,
                   float *xyz, uint8_t *valid, int32_t *match_k, int nthreads);

int orc_max_threads(void);

/* N3: MeshCreator index passes + text export (Duke/meshcreator.cpp:16-166) on a PointCloudImage stored as
 * points float [h][w][3] / count u8 [h][w]; see slr_oracle.c. */
void orc_mesh_index(const float *points, const uint8_t *count, int w, int h, int first_vertex, int *pixel_num,
                    float *vertices, int32_t *vertex_src, int32_t *faces, int64_t *nv, int64_t *nf);
int orc_export_mesh(const float *points, const uint8_t *count, const int32_t *color, int w, int h, int obj,
                    const char *path);

#ifdef __cplusplus
}
#endif
#endif
