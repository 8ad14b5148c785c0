/*
 * slr_oracle.c — CPU ORACLE (test infrastructure, never shipped / never on the product path).
 * Plain-C restatement of the reference's hot-path loops; see slr_oracle.h for the contract
 * and the pin status ("parity unpinned by reference fixtures"; pinned against oracle/_ref).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -fPIC -shared (oracle/Makefile).
 */
#include "slr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Duke/mfreconstruct.cpp:5  `float PI = 3.1416;`  (decoder constant; not pi) */
static const float PI_DEC = 3.1416f;
/* Duke/multifrequency.h:5   `#define PI 3.1416`   (generator constant; double) */
#define PI_GEN 3.1416
/* Duke/multifrequency.cpp:3 */
static const int FREQ_GEN[3] = {70, 64, 59};

static float qnanf(void)
{
    union { uint32_t u; float f; } v;
    v.u = 0x7FC00000u;
    return v.f;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* a2: Gray patterns                                                                            */
/* ------------------------------------------------------------------------------------------ */

/* Duke/graycodes.cpp:24-25 */
int orc_gray_num_bits(int n)
{
    return (int)ceil(log((double)n) / log(2.0));
}

/* Duke/graycodes.cpp:22-30 */
int orc_gray_num_imgs(int scan_w, int scan_h, int use_epi)
{
    int nc = orc_gray_num_bits(scan_w);
    int nr = orc_gray_num_bits(scan_h);
    return use_epi ? 2 + 2 * nc : 2 + 2 * nc + 2 * nr;
}

/* Duke/graycodes.cpp:55-114 */
void orc_generate_gray(uint8_t *out, int W, int H, int use_epi)
{
    const int nc = orc_gray_num_bits(W);
    const int nr = orc_gray_num_bits(H);
    const size_t P = (size_t)W * H;
    memset(out, 255, P);     /* grayCodes[0] = 255  (:60) */
    memset(out + P, 0, P);   /* grayCodes[1] = 0    (:61) */
    for (int j = 0; j < W; j++) {                       /* :63 */
        int num = j, prev_rem = j % 2;
        for (int k = 0; k < nc; k++) {
            num = num / 2;
            int rem = num % 2;
            int flag = (rem != prev_rem);               /* :68-73 */
            uint8_t on = (uint8_t)(flag * 255), off = (uint8_t)(on ? 0 : 255);
            uint8_t *img = out + (size_t)(2 * nc - 2 * k) * P;     /* :76 */
            uint8_t *inv = out + (size_t)(2 * nc - 2 * k + 1) * P; /* :81 */
            for (int i = 0; i < H; i++) {
                img[(size_t)i * W + j] = on;
                inv[(size_t)i * W + j] = off;
            }
            prev_rem = rem;
        }
    }
    if (!use_epi) {                                     /* :87 */
        for (int i = 0; i < H; i++) {
            int num = i, prev_rem = i % 2;
            for (int k = 0; k < nr; k++) {
                num = num / 2;
                int rem = num % 2;
                int flag = (rem != prev_rem);
                uint8_t on = (uint8_t)(flag * 255), off = (uint8_t)(on ? 0 : 255);
                uint8_t *img = out + (size_t)(2 * nr - 2 * k + 2 * nc) * P;     /* :100 */
                uint8_t *inv = out + (size_t)(2 * nr - 2 * k + 2 * nc + 1) * P; /* :105 */
                for (int j = 0; j < W; j++) {
                    img[(size_t)i * W + j] = on;
                    inv[(size_t)i * W + j] = off;
                }
                prev_rem = rem;
            }
        }
    }
}

/* Duke/graycodes.cpp:116-128 */
int orc_gray_to_dec(const uint8_t *bits, int n)
{
    int dec = 0;
    int tmp = bits[0] ? 1 : 0;
    if (tmp)
        dec += (int)powf(2.0f, (float)(n - 1));
    for (int i = 1; i < n; i++) {
        tmp = (tmp != (bits[i] ? 1 : 0)); /* Utilities::XOR, utilities.cpp:11-17 */
        if (tmp)
            dec += (int)powf(2.0f, (float)(n - i - 1));
    }
    return dec;
}

/* ------------------------------------------------------------------------------------------ */
/* a1: multi-frequency fringe patterns                                                          */
/* ------------------------------------------------------------------------------------------ */

/* Duke/multifrequency.cpp:14-33 */
void orc_generate_mf(uint8_t *out, int projW, int projH)
{
    const size_t P = (size_t)projW * projH;
    memset(out, 255, P);   /* :16 */
    memset(out + P, 0, P); /* :17 */
    for (int f = 0; f < 3; f++) {
        for (int phi = 0; phi < 4; phi++) {
            uint8_t *img = out + (size_t)(4 * f + phi + 2) * P; /* :30 */
            for (int w = 0; w < projW; w++) {
                /* :27  PI*2*w*frequency[f]/projW + PI*phi/2, evaluated left to right in double */
                double arg = PI_GEN * 2 * (double)w * (double)FREQ_GEN[f] / (double)projW
                             + PI_GEN * (double)phi / 2;
                float c = cosf((float)arg);
                float v = 135.0f + 79.0f * c;
                uint8_t px = (uint8_t)v; /* float -> uchar truncation */
                for (int h = 0; h < projH; h++)
                    img[(size_t)h * projW + w] = px;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* a3: shadow mask                                                                              */
/* ------------------------------------------------------------------------------------------ */

/* Duke/mfreconstruct.cpp:196-204, Duke/reconstruct.cpp:216-224 */
void orc_shadow_mask(const uint8_t *white, const uint8_t *black, int npix, int black_thr, uint8_t *mask)
{
    for (int p = 0; p < npix; p++) {
        float blackVal = (float)black[p];
        float whiteVal = (float)white[p];
        mask[p] = (whiteVal - blackVal > (float)black_thr) ? 1 : 0;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* a4: strict-mode phase                                                                        */
/* ------------------------------------------------------------------------------------------ */

/* Duke/mfreconstruct.cpp:246-261 with a = G4-G2, b = G1-G3; first matching branch wins. */
int orc_wrapped_phase_strict(int a, int b, float *P)
{
    const float PI = PI_DEC;
    if (a == 0 && b > 0)
        *P = 0.0f;                                   /* :246 */
    else if (a == 0 && b < 0)
        *P = PI;                                     /* :248 */
    else if (b == 0 && a > 0)
        *P = 3.0f * PI / 2.0f;                       /* :250  3*PI/2 (int*float, /int) */
    else if (b == 0 && a < 0)
        *P = PI / 2.0f;                              /* :252 */
    else if (b == 0 && a == 0)
        return 0;                                    /* :254-255 mask cleared, P undefined (F4) */
    else if (b < 0)
        *P = atanf((float)(a / b)) + PI;             /* :257  C++ int division, atan(float) */
    else if (b > 0 && a > 0)
        *P = atanf((float)(a / b)) + 2.0f * PI;      /* :259 */
    else
        *P = atanf((float)(a / b));                  /* :261 */
    return 1;
}

/* Duke/mfreconstruct.cpp:231-269 */
int orc_get_phase_strict(const int *G, float *phase)
{
    const float PI = PI_DEC;
    double P[3];
    for (int count = 0; count < 3; count++) {
        int G1 = G[4 * count + 0], G2 = G[4 * count + 1];
        int G3 = G[4 * count + 2], G4 = G[4 * count + 3];
        float p;
        if (!orc_wrapped_phase_strict(G4 - G2, G1 - G3, &p))
            return 0; /* oracle definition of F4: the pixel is dropped */
        P[count] = (double)p;
    }
    /* :265-268 — P[] double, 2*PI float promoted to double, result narrowed to float */
    float P12 = (float)((P[0] > P[1]) ? (P[0] - P[1]) : (P[0] - P[1] + (double)(2.0f * PI)));
    float P23 = (float)((P[1] > P[2]) ? (P[1] - P[2]) : (P[1] - P[2] + (double)(2.0f * PI)));
    float d = P12 - P23;
    float P123 = (P12 > P23) ? d : (d + 2.0f * PI);
    float q = P123 / (2.0f * PI);
    *phase = q * 255.0f;
    return 1;
}

/* corrected mode (no reference counterpart): S-step atan2 wrapped phase per
 * frequency, cascade of wrapped differences down to one beat phase, scaled to 0..255. */
static int corrected_phase(const uint8_t *const *imgs, size_t p, int F, int S,
                           const float *cs, const float *sn, float *phase)
{
    const float TWO_PI = 6.28318530717958647692f;
    float lvl[16];
    for (int f = 0; f < F; f++) {
        float num, den;
        if (S == 4) {
            /* exact integer form: num = G4-G2, den = G1-G3 */
            int G1 = imgs[2 + 4 * f + 0][p], G2 = imgs[2 + 4 * f + 1][p];
            int G3 = imgs[2 + 4 * f + 2][p], G4 = imgs[2 + 4 * f + 3][p];
            num = (float)(G4 - G2);
            den = (float)(G1 - G3);
            if (num == 0.0f && den == 0.0f)
                return 0;
        } else {
            num = 0.0f;
            den = 0.0f;
            for (int s = 0; s < S; s++) {
                float I = (float)imgs[2 + S * f + s][p];
                num = num - I * sn[s];
                den = den + I * cs[s];
            }
            if (num * num + den * den < 0.25f)
                return 0;
        }
        float ph = atan2f(num, den);
        if (ph < 0.0f)
            ph = ph + TWO_PI;
        lvl[f] = ph;
    }
    for (int n = F; n > 1; n--)
        for (int i = 0; i + 1 < n; i++) {
            float d = lvl[i] - lvl[i + 1];
            lvl[i] = (d < 0.0f) ? d + TWO_PI : d;
        }
    *phase = lvl[0] / TWO_PI * 255.0f;
    return 1;
}

/* Duke/mfreconstruct.cpp:190-228 */
int orc_mf_decode_mt(const uint8_t *stack, int W, int H, int F, int S, int black_thr, int mode,
                     float *phase, uint8_t *mask, int nthreads)
{
    const size_t P = (size_t)W * H;
    if (F < 1 || F > 16 || S < 3 || S > 16)
        return -1;
    if (mode == ORC_MODE_STRICT && (F != 3 || S != 4))
        return -1; /* the reference hard-codes 3x4 (F8) */
    const uint8_t *imgs[2 + 16 * 16];
    for (int n = 0; n < 2 + F * S; n++)
        imgs[n] = stack + (size_t)n * P;
    float cs[16], sn[16];
    for (int s = 0; s < S; s++) {
        cs[s] = (float)cos(2.0 * 3.14159265358979323846 * s / S);
        sn[s] = (float)sin(2.0 * 3.14159265358979323846 * s / S);
    }
    orc_shadow_mask(imgs[0], imgs[1], (int)P, black_thr, mask); /* computeShadows :190-207 */
    /* decodePatterns :215-225.  Pixels are independent; nthreads > 1 splits them over OpenMP threads. */
#pragma omp parallel for schedule(static) num_threads(nthreads > 1 ? nthreads : 1)
    for (size_t p = 0; p < P; p++) {
        float ph = qnanf();
        if (mask[p]) {
            int ok;
            if (mode == ORC_MODE_STRICT) {
                int G[12];
                for (int n = 0; n < 12; n++)
                    G[n] = imgs[2 + n][p];
                ok = orc_get_phase_strict(G, &ph);
            } else {
                ok = corrected_phase(imgs, p, F, S, cs, sn, &ph);
            }
            if (!ok) {
                mask[p] = 0;
                ph = qnanf();
            }
        }
        phase[p] = ph;
    }
    return 0;
}

int orc_mf_decode(const uint8_t *stack, int W, int H, int F, int S, int black_thr, int mode,
                  float *phase, uint8_t *mask)
{
    return orc_mf_decode_mt(stack, W, H, F, S, black_thr, mode, phase, mask, 1);
}

/* ------------------------------------------------------------------------------------------ */
/* a7, a8: Gray decode                                                                          */
/* ------------------------------------------------------------------------------------------ */

/* Duke/reconstruct.cpp:381-407 (EPI) / 325-370 (col+row) for one pixel. Returns error flag. */
static int gray_pixel(const uint8_t *stack, size_t P, size_t p, int first_img, int nbits,
                      int white_thr, int *dec)
{
    uint8_t bits[32];
    int error = 0;
    for (int count = 0; count < nbits; count++) {
        double val1 = stack[(size_t)(first_img + count * 2) * P + p];     /* :390 */
        double val2 = stack[(size_t)(first_img + count * 2 + 1) * P + p]; /* :391 */
        if (fabs(val1 - val2) < (double)white_thr)                        /* :393 */
            error = 1;
        bits[count] = (val1 > val2) ? 1 : 0;                              /* :396-399 */
    }
    *dec = orc_gray_to_dec(bits, nbits);
    return error;
}

void orc_gray_decode(const uint8_t *stack, int W, int H, int nbits_col, int nbits_row,
                     int black_thr, int white_thr, int scan_w, int scan_h,
                     int32_t *col, int32_t *row, uint8_t *mask)
{
    const size_t P = (size_t)W * H;
    orc_shadow_mask(stack, stack + P, (int)P, black_thr, mask); /* reconstruct.cpp:210-227 */
    for (size_t p = 0; p < P; p++) {
        int x = -1, y = -1;
        if (mask[p]) {
            int error = gray_pixel(stack, P, p, 2, nbits_col, white_thr, &x);
            if (nbits_row > 0) {
                error |= gray_pixel(stack, P, p, 2 + 2 * nbits_col, nbits_row, white_thr, &y); /* :349-363 */
                if (y > scan_h || x > scan_w)   /* :365 */
                    error = 1;
            } else {
                if (x > scan_w)                 /* :403 */
                    error = 1;
            }
            if (error) {                        /* :88-92 / :65-68 */
                mask[p] = 0;
                x = -1;
                y = -1;
            }
        }
        col[p] = x;
        if (row)
            row[p] = y;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* a11-a13: geometry helpers                                                                    */
/* ------------------------------------------------------------------------------------------ */

/* Duke/utilities.cpp:58-94 */
void orc_undistort_point(float px, float py, const orc_camera *cam, float *ox, float *oy)
{
    double k[5] = {0, 0, 0, 0, 0}, fx, fy, ifx, ify, cx, cy;
    k[0] = cam->dist[0];
    k[1] = cam->dist[1];
    k[2] = cam->dist[2];
    k[3] = cam->dist[3];
    k[4] = 0;                /* :66 */
    fx = cam->fc[0];
    fy = cam->fc[1];
    ifx = 1. / fx;
    ify = 1. / fy;
    cx = cam->cc[0];
    cy = cam->cc[1];
    double x, y, x0, y0;
    x = px;
    y = py;
    x0 = x = (x - cx) * ifx;
    y0 = y = (y - cy) * ify;
    for (int jj = 0; jj < 5; jj++) {
        double r2 = x * x + y * y;
        double icdist = 1. / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
        double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x);
        double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
    }
    /* :93  cv::Point2f((float)(x*fx)+cx, (float)(y*fy)+cy) — float + double cx, narrowed */
    *ox = (float)((double)(float)(x * fx) + cx);
    *oy = (float)((double)(float)(y * fy) + cy);
}

/* Duke/reconstruct.cpp:310-322:  tmp = -R^T * t ; tmpPoint = R^T * p ; p = tmp + tmpPoint  (CV_32F Mats).
 * The two 3x3 * 3x1 products are cv::Mat arithmetic (third party): each element is accumulated in double,
 * k ascending, and narrowed to float (OpenCV's CV_32F gemm uses a double accumulator); the unary minus is
 * applied to the transposed matrix first, as the expression parses; the final sum is a float addition. */
void orc_cam2world(const orc_camera *cam, float p[3])
{
    const float *R = cam->R, *t = cam->t;
    float o[3];
    for (int i = 0; i < 3; i++) {
        double st = 0.0, sp = 0.0;
        for (int k = 0; k < 3; k++) {
            st += (double)(-R[k * 3 + i]) * (double)t[k];
            sp += (double)R[k * 3 + i] * (double)p[k];
        }
        o[i] = (float)st + (float)sp;
    }
    p[0] = o[0];
    p[1] = o[1];
    p[2] = o[2];
}

/* Duke/utilities.cpp:19-28 */
void orc_normalize(float v[3])
{
    float ss = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    double mag = (double)sqrtf(ss);
    double m = (0.000001 > mag) ? 0.000001 : mag;
    v[0] /= (float)m;
    v[1] /= (float)m;
    v[2] /= (float)m;
}

static float dot3(const float a[3], const float b[3])
{
    float s = 0.0f;
    for (int i = 0; i < 3; i++)
        s += a[i] * b[i];
    return s;
}

/* Duke/utilities.cpp:399-425 */
int orc_line_line_intersection(const float p1[3], const float v1[3], const float p2[3],
                               const float v2[3], float p[3])
{
    float v12[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
    float v1_dot_v1 = dot3(v1, v1);
    float v2_dot_v2 = dot3(v2, v2);
    float v1_dot_v2 = dot3(v1, v2);
    float v12_dot_v1 = dot3(v12, v1);
    float v12_dot_v2 = dot3(v12, v2);
    float denom = v1_dot_v1 * v2_dot_v2 - v1_dot_v2 * v1_dot_v2;
    if ((double)fabsf(denom) < 0.1)
        return 0;
    float s = (v1_dot_v2 / denom) * v12_dot_v2 - (v2_dot_v2 / denom) * v12_dot_v1;
    float t = -(v1_dot_v2 / denom) * v12_dot_v1 + (v1_dot_v1 / denom) * v12_dot_v2;
    for (int i = 0; i < 3; i++) {
        float a = p1[i] + s * v1[i];
        float b = p2[i] + t * v2[i];
        float sum = a + b;
        p[i] = (float)(0.5 * (double)sum);
    }
    return 1;
}

/* Q * [x y d 1]^T in double, sums left to right (cv::Mat 4x4 * 4x1, CV_64F), then /w, narrowed.
 * Duke/mfreconstruct.cpp:299-311, Duke/reconstruct.cpp:570-582 */
static void reproject_q(const double *Q, const double p2d[4], float out[3])
{
    double r[4];
    for (int i = 0; i < 4; i++)
        r[i] = Q[4 * i + 0] * p2d[0] + Q[4 * i + 1] * p2d[1] + Q[4 * i + 2] * p2d[2] + Q[4 * i + 3] * p2d[3];
    out[0] = (float)(r[0] / r[3]);
    out[1] = (float)(r[1] / r[3]);
    out[2] = (float)(r[2] / r[3]);
}

/* 3x4 CV_32F * 4x1 CV_32F (Duke/mfreconstruct.cpp:315-323); products accumulated in double
 * (OpenCV's float gemm uses a double accumulator), narrowed to float. */
static void apply_rigid(const float *M, float p[3])
{
    float o[3];
    for (int i = 0; i < 3; i++) {
        double s = (double)M[4 * i + 0] * (double)p[0] + (double)M[4 * i + 1] * (double)p[1]
                   + (double)M[4 * i + 2] * (double)p[2] + (double)M[4 * i + 3] * 1.0;
        o[i] = (float)s;
    }
    p[0] = o[0];
    p[1] = o[1];
    p[2] = o[2];
}

/* ------------------------------------------------------------------------------------------ */
/* a6: MF match + triangulate                                                                   */
/* ------------------------------------------------------------------------------------------ */

/* one rectified row of Duke/mfreconstruct.cpp:284-331 */
static int64_t mf_row(int i, const float *phL, const uint8_t *mkL, const float *phR, const uint8_t *mkR,
                      int W, const orc_camera *camL, const orc_camera *camR, const double *Q,
                      const float *rigid, float *xyz, uint8_t *valid, int32_t *match_k)
{
    int64_t n = 0;
    const size_t base = (size_t)i * W;
    for (int j = 0; j < W; j++) {
        float *o = xyz + (base + j) * 3;
        o[0] = o[1] = o[2] = qnanf();
        valid[base + j] = 0;
        if (match_k)
            match_k[base + j] = -1;
        if (!mkL[base + j])              /* :287 cam1Pix.size() == 0 */
            continue;
        float pl = phL[base + j];
        for (int k = 0; k < W; k++) {    /* :289 */
            if (!mkR[base + k])          /* :292 */
                continue;
            float d = pl - phR[base + k];
            if ((double)fabsf(d) < 0.1) { /* :295 fabs(float) < 0.1 (double) */
                float ulx, uly, urx, ury;
                orc_undistort_point((float)j, (float)i, camL, &ulx, &uly); /* :297 */
                orc_undistort_point((float)k, (float)i, camR, &urx, &ury); /* :298 */
                float disp = ulx - urx;
                double p2d[4] = {ulx, uly, disp, 1};                       /* :299 */
                float pt[3];
                reproject_q(Q, p2d, pt);
                if (rigid)                                                 /* :315 scanSN > 0 */
                    apply_rigid(rigid, pt);
                o[0] = pt[0];
                o[1] = pt[1];
                o[2] = pt[2];
                valid[base + j] = 1;
                if (match_k)
                    match_k[base + j] = k;
                n++;
                break;                                                     /* :327 */
            }
        }
    }
    return n;
}

int64_t orc_mf_triangulate(const float *phL, const uint8_t *mkL, const float *phR, const uint8_t *mkR,
                           int W, int H, const orc_camera *camL, const orc_camera *camR,
                           const double *Q, const float *rigid,
                           float *xyz, uint8_t *valid, int32_t *match_k, int nthreads)
{
    int64_t n = 0;
    if (nthreads <= 1) {
        for (int i = 0; i < H; i++)
            n += mf_row(i, phL, mkL, phR, mkR, W, camL, camR, Q, rigid, xyz, valid, match_k);
    } else {
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads) reduction(+ : n)
        for (int i = 0; i < H; i++)
            n += mf_row(i, phL, mkL, phR, mkR, W, camL, camR, Q, rigid, xyz, valid, match_k);
    }
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* a9: GE match + triangulate                                                                   */
/* ------------------------------------------------------------------------------------------ */

/* one row of Duke/reconstruct.cpp:555-611 */
static int64_t ge_row(int i, const int32_t *colL, const uint8_t *mkL, const int32_t *colR, const uint8_t *mkR,
                      int W, const double *Q, const float *rigid, const uint8_t *whiteL, const uint8_t *whiteR,
                      float *xyz, uint8_t *valid, int32_t *match_k, uint8_t *color)
{
    int64_t n = 0;
    const size_t base = (size_t)i * W;
    int kstart = 0;                                 /* :556 */
    for (int j = 0; j < W; j++) {
        float *o = xyz + (base + j) * 3;
        o[0] = o[1] = o[2] = qnanf();
        valid[base + j] = 0;
        if (match_k)
            match_k[base + j] = -1;
        if (color)
            color[base + j] = 0;
        if (!mkL[base + j])                         /* :559 */
            continue;
        for (int k = kstart; k < W; k++) {          /* :561 */
            if (!mkR[base + k])
                continue;
            if (colL[base + j] == colR[base + k]) { /* :565 */
                double p2d[4] = {(double)j, (double)i, (double)(j - k), 1}; /* :570 */
                float pt[3];
                reproject_q(Q, p2d, pt);
                if (rigid)
                    apply_rigid(rigid, pt);
                o[0] = pt[0];
                o[1] = pt[1];
                o[2] = pt[2];
                valid[base + j] = 1;
                if (match_k)
                    match_k[base + j] = k;
                if (color && whiteL && whiteR)      /* :597-600 */
                    color[base + j] = (uint8_t)(((int)whiteL[base + j] + (int)whiteR[base + k]) / 2);
                kstart = k;                         /* :604 */
                n++;
                break;
            }
        }
    }
    return n;
}

int64_t orc_ge_triangulate(const int32_t *colL, const uint8_t *mkL, const int32_t *colR, const uint8_t *mkR,
                           int W, int H, const double *Q, const float *rigid,
                           const uint8_t *whiteL, const uint8_t *whiteR,
                           float *xyz, uint8_t *valid, int32_t *match_k, uint8_t *color, int nthreads)
{
    int64_t n = 0;
    if (nthreads <= 1) {
        for (int i = 0; i < H; i++)
            n += ge_row(i, colL, mkL, colR, mkR, W, Q, rigid, whiteL, whiteR, xyz, valid, match_k, color);
    } else {
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads) reduction(+ : n)
        for (int i = 0; i < H; i++)
            n += ge_row(i, colL, mkL, colR, mkR, W, Q, rigid, whiteL, whiteR, xyz, valid, match_k, color);
    }
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* a10: GRAY_ONLY bucket triangulation                                                          */
/* ------------------------------------------------------------------------------------------ */

typedef struct { int32_t x, y; } pix_t;

/* Duke/reconstruct.cpp:56-74: camera pixels pushed into projector cell ac(x,y)=x*scan_h+y in camera
 * column-major order.  Cells whose index falls outside the scan_w*scan_h array (xDec == scan_w,
 * an out-of-bounds write in the reference) are dropped here. */
static void build_buckets(const int32_t *col, const int32_t *row, const uint8_t *mk, int W, int H,
                          int scan_w, int scan_h, int32_t *start, pix_t *items)
{
    const size_t ncell = (size_t)scan_w * scan_h;
    memset(start, 0, (ncell + 1) * sizeof(int32_t));
    for (int c = 0; c < W; c++)
        for (int r = 0; r < H; r++) {
            size_t p = (size_t)r * W + c;
            if (!mk[p])
                continue;
            size_t cell = (size_t)col[p] * scan_h + (size_t)row[p];
            if (cell < ncell)
                start[cell + 1]++;
        }
    for (size_t c = 0; c < ncell; c++)
        start[c + 1] += start[c];
    int32_t *fill = (int32_t *)malloc(ncell * sizeof(int32_t));
    memcpy(fill, start, ncell * sizeof(int32_t));
    for (int c = 0; c < W; c++)
        for (int r = 0; r < H; r++) {
            size_t p = (size_t)r * W + c;
            if (!mk[p])
                continue;
            size_t cell = (size_t)col[p] * scan_h + (size_t)row[p];
            if (cell < ncell) {
                items[fill[cell]].x = c;
                items[fill[cell]].y = r;
                fill[cell]++;
            }
        }
    free(fill);
}

/* Duke/reconstruct.cpp:417-481 */
int64_t orc_gray_triangulate(const int32_t *colL, const int32_t *rowL, const uint8_t *mkL,
                             const int32_t *colR, const int32_t *rowR, const uint8_t *mkR,
                             int W, int H, int scan_w, int scan_h,
                             const orc_camera *camL, const orc_camera *camR, const float *rigid,
                             float *sum, uint8_t *cnt)
{
    const size_t ncell = (size_t)scan_w * scan_h;
    const size_t P = (size_t)W * H;
    int32_t *s1 = (int32_t *)malloc((ncell + 1) * sizeof(int32_t));
    int32_t *s2 = (int32_t *)malloc((ncell + 1) * sizeof(int32_t));
    pix_t *it1 = (pix_t *)malloc(P * sizeof(pix_t));
    pix_t *it2 = (pix_t *)malloc(P * sizeof(pix_t));
    build_buckets(colL, rowL, mkL, W, H, scan_w, scan_h, s1, it1);
    build_buckets(colR, rowR, mkR, W, H, scan_w, scan_h, s2, it2);

    /* reconstruct.cpp:239-240: position = cam2WorldSpace((0,0,0)) */
    float pos1[3] = {0, 0, 0}, pos2[3] = {0, 0, 0};
    orc_cam2world(camL, pos1);
    orc_cam2world(camR, pos2);

    memset(sum, 0, ncell * 3 * sizeof(float));
    memset(cnt, 0, ncell);
    int64_t ncells_out = 0;
    for (size_t cell = 0; cell < ncell; cell++) {       /* :428-429, ac(i,j) = i*scan_h + j */
        int n1 = s1[cell + 1] - s1[cell], n2 = s2[cell + 1] - s2[cell];
        if (n1 == 0 || n2 == 0)                         /* :436 */
            continue;
        for (int c1 = 0; c1 < n1; c1++) {
            pix_t a = it1[s1[cell] + c1];
            float ux, uy;
            orc_undistort_point((float)a.x, (float)a.y, camL, &ux, &uy);      /* :441 */
            float pt1[3] = {(ux - camL->cc[0]) / camL->fc[0], (uy - camL->cc[1]) / camL->fc[1], 1.0f}; /* utilities.cpp:47-56 */
            orc_cam2world(camL, pt1);
            float ray1[3] = {pos1[0] - pt1[0], pos1[1] - pt1[1], pos1[2] - pt1[2]}; /* :445 */
            orc_normalize(ray1);
            for (int c2 = 0; c2 < n2; c2++) {
                pix_t b = it2[s2[cell] + c2];
                orc_undistort_point((float)b.x, (float)b.y, camR, &ux, &uy);
                float pt2[3] = {(ux - camR->cc[0]) / camR->fc[0], (uy - camR->cc[1]) / camR->fc[1], 1.0f};
                orc_cam2world(camR, pt2);
                float ray2[3] = {pos2[0] - pt2[0], pos2[1] - pt2[1], pos2[2] - pt2[2]};
                orc_normalize(ray2);
                float ip[3];
                if (!orc_line_line_intersection(pos1, ray1, pos2, ray2, ip))  /* :460-463 */
                    continue;
                if (rigid)
                    apply_rigid(rigid, ip);
                /* PointCloudImage::addPoint, pointcloudimage.cpp:86-97 (u8 count wraps) */
                float *acc = sum + cell * 3;
                if (cnt[cell] == 0) {
                    acc[0] = ip[0];
                    acc[1] = ip[1];
                    acc[2] = ip[2];
                    cnt[cell] = 1;
                } else {
                    acc[0] = ip[0] + acc[0];
                    acc[1] = ip[1] + acc[1];
                    acc[2] = ip[2] + acc[2];
                    cnt[cell] = (uint8_t)(cnt[cell] + 1);
                }
            }
        }
        if (cnt[cell])
            ncells_out++;
    }
    free(s1);
    free(s2);
    free(it1);
    free(it2);
    return ncells_out;
}

/* ------------------------------------------------------------------------------------------ */
/* a15: PointCloudImage accumulation as called by the MF / GE paths                             */
/* ------------------------------------------------------------------------------------------ */

void orc_pointcloud_from_dense(const float *xyz, const uint8_t *valid, int W, int H,
                               int scan_w, int scan_h, float *points, uint8_t *count)
{
    /* PointCloudImage(scan_w, scan_h): points = Mat(h, w, CV_32FC3) (uninitialised in the reference;
     * zero here), numOfPointsForPixel = Mat(h, w, CV_8U, 0).  pointcloudimage.cpp:3-13 */
    memset(points, 0, (size_t)scan_w * scan_h * 3 * sizeof(float));
    memset(count, 0, (size_t)scan_w * scan_h);
    for (int i = 0; i < H; i++)
        for (int j = 0; j < W; j++) {
            size_t p = (size_t)i * W + j;
            if (!valid[p])
                continue;
            int i_w = i, j_h = j;                       /* addPoint(i, j, p): mfreconstruct.cpp:326 */
            if (i_w >= scan_w || j_h >= scan_h)         /* pointcloudimage.cpp:88 (F7) */
                continue;
            size_t q = (size_t)j_h * scan_w + i_w;      /* matSet3D(points, i_w, j_h): at(row=j_h, col=i_w) */
            if (count[q] == 0) {
                points[q * 3 + 0] = xyz[p * 3 + 0];
                points[q * 3 + 1] = xyz[p * 3 + 1];
                points[q * 3 + 2] = xyz[p * 3 + 2];
                count[q] = 1;
            } else {
                points[q * 3 + 0] += xyz[p * 3 + 0];
                points[q * 3 + 1] += xyz[p * 3 + 1];
                points[q * 3 + 2] += xyz[p * 3 + 2];
                count[q] = (uint8_t)(count[q] + 1);
            }
        }
}

/* ------------------------------------------------------------------------------------------ */
/* N3: MeshCreator — vertex numbering, grid-neighbour faces, PLY / OBJ text                      */
/* ------------------------------------------------------------------------------------------ */

/* PointCloudImage::getPoint, pointcloudimage.cpp:56-67: Vec3d(points) / (float)num == points * (1.f/num)
 * evaluated in double (OpenCV 2.4 Vec operator/), narrowed to Point3f. */
static int orc_get_point(const float *points, const uint8_t *count, int w, int h, int i_w, int j_h, float out[3])
{
    if (i_w >= w || j_h >= h)
        return 0;
    size_t q = (size_t)j_h * w + i_w;
    uint8_t num = count[q];
    if (num == 0)
        return 0;
    float inv = 1.f / (float)num;
    for (int c = 0; c < 3; c++)
        out[c] = (float)((double)points[q * 3 + c] * inv);
    return 1;
}

/* The index passes of exportPlyMesh (meshcreator.cpp:75-110, first_vertex = 0) and exportObjMesh (:25-63,
 * first_vertex = 1).  vertices [w*h][3], vertex_src [w*h] (storage element j*w+i), faces [2*w*h][3];
 * returns the counts through nv / nf.  pixel_num is caller scratch [w*h]. */
void orc_mesh_index(const float *points, const uint8_t *count, int w, int h, int first_vertex, int *pixel_num,
                    float *vertices, int32_t *vertex_src, int32_t *faces, int64_t *nv, int64_t *nf)
{
    int vertex = first_vertex;
    int64_t n = 0;
    for (int i = 0; i < w; i++)
        for (int j = 0; j < h; j++) {
            float pt[3];
            if (orc_get_point(points, count, w, h, i, j, pt)) {
                pixel_num[i * h + j] = vertex++;          /* access(i,j) = i*h + j, meshcreator.cpp:168-171 */
                vertices[n * 3 + 0] = pt[0];
                vertices[n * 3 + 1] = pt[1];
                vertices[n * 3 + 2] = pt[2];
                if (vertex_src)
                    vertex_src[n] = (int32_t)((size_t)j * w + i);
                n++;
            } else {
                pixel_num[i * h + j] = 0;
            }
        }
    *nv = n;
    int64_t f = 0;
    for (int i = 0; i < w; i++)
        for (int j = 0; j < h; j++) {
            int v1 = pixel_num[i * h + j], v2, v3;
            v2 = (i < w - 1) ? pixel_num[(i + 1) * h + j] : 0;
            v3 = (j < h - 1) ? pixel_num[i * h + j + 1] : 0;
            if (v1 != 0 && v2 != 0 && v3 != 0) {          /* "3 v1 v2 v3", :151-152 */
                faces[f * 3 + 0] = v1, faces[f * 3 + 1] = v2, faces[f * 3 + 2] = v3;
                f++;
            }
            v3 = (j > 0 && i < w - 1) ? pixel_num[(i + 1) * h + j - 1] : 0;
            if (v1 != 0 && v2 != 0 && v3 != 0) {          /* "3 v1 v3 v2", :159-160 */
                faces[f * 3 + 0] = v1, faces[f * 3 + 1] = v3, faces[f * 3 + 2] = v2;
                f++;
            }
        }
    *nf = f;
}

/* exportPlyMesh (:67-166) / exportObjMesh (:16-65) text.  `ostream << float` is printf("%g") (precision 6).
 * colour: sums int [h][w][3] or NULL; without a colour plane getPoint yields (100, 0, 0) — the reference's
 * `(cv::Point3i)(100,100,100)` is a comma expression (pointcloudimage.cpp:50) — printed as "0 0 100". */
int orc_export_mesh(const float *points, const uint8_t *count, const int32_t *color, int w, int h, int obj,
                    const char *path)
{
    size_t px = (size_t)w * h;
    int *pn = (int *)malloc(px * sizeof(int));
    float *vert = (float *)malloc(px * 3 * sizeof(float));
    int32_t *src = (int32_t *)malloc(px * sizeof(int32_t));
    int32_t *faces = (int32_t *)malloc(px * 6 * sizeof(int32_t));
    int64_t nv = 0, nf = 0;
    FILE *fp = fopen(path, "w");
    if (!pn || !vert || !src || !faces || !fp) {
        free(pn), free(vert), free(src), free(faces);
        if (fp)
            fclose(fp);
        return -1;
    }
    orc_mesh_index(points, count, w, h, obj ? 1 : 0, pn, vert, src, faces, &nv, &nf);
    if (!obj) {
        fprintf(fp, "ply\nformat ascii 1.0\nelement vertex %lld\n", (long long)nv);
        fprintf(fp, "property float x\nproperty float y\nproperty float z\n");
        fprintf(fp, "property uchar red\nproperty uchar green\nproperty uchar blue\n");
        fprintf(fp, "element face %lld\nproperty list uchar int vertex_indices\nend_header\n", (long long)nf);
    }
    for (int64_t v = 0; v < nv; v++) {
        if (obj) {
            fprintf(fp, "v %g %g %g\n", vert[v * 3], vert[v * 3 + 1], vert[v * 3 + 2]);
        } else {
            int c[3] = {100, 0, 0};
            if (color) {
                float inv = 1.f / (float)count[src[v]];   /* Vec3i / float: saturate_cast<int>(v * (1.f/num)) = cvRound */
                for (int k = 0; k < 3; k++)
                    c[k] = (int)lrint((double)color[(size_t)src[v] * 3 + k] * inv);
            }
            fprintf(fp, "%g %g %g %d %d %d\n", vert[v * 3], vert[v * 3 + 1], vert[v * 3 + 2], c[2], c[1], c[0]);
        }
    }
    for (int64_t f = 0; f < nf; f++) {
        const int32_t *t = faces + f * 3;
        if (obj)
            fprintf(fp, "f %d/%d %d/%d %d/%d\n", t[0], t[0], t[1], t[1], t[2], t[2]);
        else
            fprintf(fp, "3 %d %d %d\n", t[0], t[1], t[2]);
    }
    fclose(fp);
    free(pn), free(vert), free(src), free(faces);
    return 0;
}

/* PointCloudImage::addPoint applied to a sequence of (i_w, j_h, point) calls: pointcloudimage.cpp:86-97 with
 * setPoint :28-37.  The u8 count wraps at 256, after which the next add RESETS the sum (num == 0 -> setPoint). */
void orc_pointcloud_add(int w, int h, const int32_t *iw, const int32_t *jh, const float *pts, int n,
                        float *points, uint8_t *count)
{
    memset(points, 0, (size_t)w * h * 3 * sizeof(float));
    memset(count, 0, (size_t)w * h);
    for (int k = 0; k < n; k++) {
        if (iw[k] >= w || jh[k] >= h)               /* :88 */
            continue;
        size_t q = (size_t)jh[k] * w + iw[k];
        if (count[q] == 0) {                        /* :91-92 */
            points[q * 3 + 0] = pts[3 * k + 0];
            points[q * 3 + 1] = pts[3 * k + 1];
            points[q * 3 + 2] = pts[3 * k + 2];
            count[q] = 1;
        } else {                                    /* :93-95  point + p */
            points[q * 3 + 0] = pts[3 * k + 0] + points[q * 3 + 0];
            points[q * 3 + 1] = pts[3 * k + 1] + points[q * 3 + 1];
            points[q * 3 + 2] = pts[3 * k + 2] + points[q * 3 + 2];
            count[q] = (uint8_t)(count[q] + 1);
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* N1: stereoRect::doStereoRectify == cv::remap(img, out, map1, map2, INTER_LINEAR)               */
/* ------------------------------------------------------------------------------------------ */

/* Duke/stereorect.cpp:26-34 with the CV_16SC2 fixed-point maps of :42-43.  OpenCV 2.4 imgproc (third party,
 * restated from the published algorithm): map1 = integer source coordinates, map2 = 5+5 fractional bits;
 * weights = bilinear products scaled to 2^15 and stored as shorts (the (0,0) entry saturates to 32767 and the
 * table normalisation gives the missing 1 to the diagonal tap); result = (sum + 2^14) >> 15;
 * BORDER_CONSTANT, value 0. */
void orc_remap_linear(const uint8_t *src, int W, int H, const int16_t *map1, const uint16_t *map2, uint8_t *dst)
{
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const size_t o = (size_t)y * W + x;
            const int sx = map1[2 * o], sy = map1[2 * o + 1];
            const int a = map2[o] & 1023, fx = a & 31, fy = a >> 5;
            int w[4] = {(32 - fx) * (32 - fy) * 32, fx * (32 - fy) * 32, (32 - fx) * fy * 32, fx * fy * 32};
            if (a == 0) {
                w[0] = 32767;
                w[3] = 1;
            }
            int sum = 0;
            for (int t = 0; t < 4; t++) {
                const int xx = sx + (t & 1), yy = sy + (t >> 1);
                const int v = (xx >= 0 && xx < W && yy >= 0 && yy < H) ? src[(size_t)yy * W + xx] : 0;
                sum += v * w[t];
            }
            dst[o] = (uint8_t)((sum + (1 << 14)) >> 15);
        }
}

/* ------------------------------------------------------------------------------------------ */
/* Utilities::autoContrast, Duke/utilities.cpp:340-355, channel 0                                */
/* ------------------------------------------------------------------------------------------ */
/* cv::saturate_cast<uchar>(cvRound(x)): lrint (round half to even), then clamp to 0..255 */
static uint8_t sat_round_u8(float x)
{
    long r = lrintf(x);
    return (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
}

void orc_auto_contrast(uint8_t *img, int W, int H)
{
    const size_t P = (size_t)W * H;
    int mn = 255, mx = 0;                                  /* cv::minMaxIdx, :346 */
    for (size_t k = 0; k < P; k++) {
        if (img[k] < mn) mn = img[k];
        if (img[k] > mx) mx = img[k];
    }
    if (P == 0) return;
    const double min = (double)mn + 255 * 0.05;            /* :347 */
    const double a = 255 / ((double)mx - min);             /* :349 */
    /* bgr[i] -= min (:350): cv::subtract with a non-integer scalar computes in CV_32F and saturates to 8 bits;
     * bgr[i] *= a (:351): Mat::operator*= -> convertTo(alpha = a) -> cvtScale_<uchar, uchar, float> */
    const float minf = (float)min, af = (float)a;
    for (size_t k = 0; k < P; k++) {
        const uint8_t t = sat_round_u8((float)img[k] - minf);
        img[k] = sat_round_u8((float)t * af);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* whole MF pipeline for one scan (Duke/mfreconstruct.cpp:160-187 minus image IO)               */
/* ------------------------------------------------------------------------------------------ */

int64_t orc_run_mf(const uint8_t *stacks, int W, int H, int F, int S, int black_thr, int mode,
                   const orc_camera *cams, const double *Q, const float *rigid,
                   float *xyz, uint8_t *valid, int32_t *match_k, int nthreads)
{
    const size_t P = (size_t)W * H;
    const size_t N = (size_t)(2 + F * S);
    float *ph = (float *)malloc(2 * P * sizeof(float));
    uint8_t *mk = (uint8_t *)malloc(2 * P);
    int64_t n = -1;
    if (orc_mf_decode_mt(stacks, W, H, F, S, black_thr, mode, ph, mk, nthreads) == 0 &&
        orc_mf_decode_mt(stacks + N * P, W, H, F, S, black_thr, mode, ph + P, mk + P, nthreads) == 0)
        n = orc_mf_triangulate(ph, mk, ph + P, mk + P, W, H, &cams[0], &cams[1], Q, rigid,
                               xyz, valid, match_k, nthreads);
    free(ph);
    free(mk);
    return n;
}
