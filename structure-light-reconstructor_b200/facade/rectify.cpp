#include "rectify.h"

#include <float.h>
#include <math.h>
#include <string.h>

namespace duke {

namespace {

typedef double M3[3][3];

void mul33(const M3 a, const M3 b, M3 out)
{
    M3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r[i][j] = a[i][0] * b[0][j] + a[i][1] * b[1][j] + a[i][2] * b[2][j];
    memcpy(out, r, sizeof(M3));
}

void transpose33(const M3 a, M3 out)
{
    M3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r[i][j] = a[j][i];
    memcpy(out, r, sizeof(M3));
}

// nearest rotation U*V^T of a 3x3 matrix by one-sided Jacobi SVD (cvRodrigues2 orthonormalises its input first)
void orthonormalize(const M3 in, M3 out)
{
    double A[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    memcpy(A, in, sizeof(A));
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < 3; i++) {
                    alpha += A[i][p] * A[i][p];
                    beta += A[i][q] * A[i][q];
                    gamma += A[i][p] * A[i][q];
                }
                off += fabs(gamma);
                if (fabs(gamma) < 1e-300) continue;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int i = 0; i < 3; i++) {
                    const double ap = A[i][p], aq = A[i][q];
                    A[i][p] = c * ap - s * aq;
                    A[i][q] = s * ap + c * aq;
                    const double vp = V[i][p], vq = V[i][q];
                    V[i][p] = c * vp - s * vq;
                    V[i][q] = s * vp + c * vq;
                }
            }
        if (off < 1e-18) break;
    }
    // A = U * S (columns), so U = A * S^-1 and U*V^T is the polar factor
    double U[3][3];
    for (int j = 0; j < 3; j++) {
        double n = sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
        if (n < 1e-300) n = 1;
        for (int i = 0; i < 3; i++) U[i][j] = A[i][j] / n;
    }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) out[i][j] = U[i][0] * V[j][0] + U[i][1] * V[j][1] + U[i][2] * V[j][2];
}

// cvRodrigues2, matrix -> vector
void rodrigues_to_vec(const M3 Rin, double r[3])
{
    M3 R;
    orthonormalize(Rin, R);
    double rx = R[2][1] - R[1][2], ry = R[0][2] - R[2][0], rz = R[1][0] - R[0][1];
    const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0][0] + R[1][1] + R[2][2] - 1) * 0.5;
    c = c > 1. ? 1. : c < -1. ? -1. : c;
    double theta = acos(c);
    if (s < 1e-5) {
        if (c > 0) {
            rx = ry = rz = 0;
        } else {
            double t;
            t = (R[0][0] + 1) * 0.5;
            rx = sqrt(t > 0. ? t : 0.);
            t = (R[1][1] + 1) * 0.5;
            ry = sqrt(t > 0. ? t : 0.) * (R[0][1] < 0 ? -1. : 1.);
            t = (R[2][2] + 1) * 0.5;
            rz = sqrt(t > 0. ? t : 0.) * (R[0][2] < 0 ? -1. : 1.);
            if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[1][2] > 0) != (ry * rz > 0)) rz = -rz;
            theta /= sqrt(rx * rx + ry * ry + rz * rz);
            rx *= theta;
            ry *= theta;
            rz *= theta;
        }
    } else {
        const double vth = 1 / (2 * s) * theta;
        rx *= vth;
        ry *= vth;
        rz *= vth;
    }
    r[0] = rx;
    r[1] = ry;
    r[2] = rz;
}

// cvRodrigues2, vector -> matrix
void rodrigues_to_mat(const double r[3], M3 R)
{
    const double theta = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (theta < DBL_EPSILON) {
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) R[i][j] = (i == j);
        return;
    }
    const double c = cos(theta), s = sin(theta), c1 = 1. - c, it = 1. / theta;
    const double x = r[0] * it, y = r[1] * it, z = r[2] * it;
    const double rrt[9] = {x * x, x * y, x * z, x * y, y * y, y * z, x * z, y * z, z * z};
    const double rx[9] = {0, -z, y, z, 0, -x, -y, x, 0};
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int k = 0; k < 9; k++) R[k / 3][k % 3] = c * I[k] + c1 * rrt[k] + s * rx[k];
}

// cvUndistortPoints without R / P (normalised output), 5 iterations, float in / float out
void undistort_normalized(float &px, float &py, const Matrix &A, const Matrix &D)
{
    const double fx = A.at(0, 0), fy = A.at(1, 1), ifx = 1. / fx, ify = 1. / fy, cx = A.at(0, 2), cy = A.at(1, 2);
    double k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < (int)D.v.size() && i < 8; i++) k[i] = D.v[i];
    double x = px, y = py;
    const double x0 = x = (x - cx) * ifx, y0 = y = (y - cy) * ify;
    for (int j = 0; j < 5; j++) {
        const double r2 = x * x + y * y;
        const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
        const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x);
        const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
    }
    px = (float)x;
    py = (float)y;
}

Matrix to_matrix(const M3 m)
{
    Matrix r(3, 3);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.at(i, j) = m[i][j];
    return r;
}

}  // namespace

RectifyResult stereo_rectify(const Matrix &M1, const Matrix &D1, const Matrix &M2, const Matrix &D2, Size size,
                             const Matrix &Rm, const Matrix &Tm, RectifyVariant variant)
{
    const int nx = size.width, ny = size.height;
    M3 R, r_r, wR, Ri1, Ri2, r_rT;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R[i][j] = Rm.at(i, j);
    const double T[3] = {Tm.v[0], Tm.v[1], Tm.v[2]};
    double om[3], t[3], uu[3] = {0, 0, 0}, ww[3];
    rodrigues_to_vec(R, om);
    for (int i = 0; i < 3; i++) om[i] *= -0.5;  // average rotation
    rodrigues_to_mat(om, r_r);
    for (int i = 0; i < 3; i++) t[i] = r_r[i][0] * T[0] + r_r[i][1] * T[1] + r_r[i][2] * T[2];
    const int idx = fabs(t[0]) > fabs(t[1]) ? 0 : 1;
    const double c = t[idx], nt = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
    uu[idx] = c > 0 ? 1 : -1;
    ww[0] = t[1] * uu[2] - t[2] * uu[1];  // global Z rotation: t x uu
    ww[1] = t[2] * uu[0] - t[0] * uu[2];
    ww[2] = t[0] * uu[1] - t[1] * uu[0];
    const double nw = sqrt(ww[0] * ww[0] + ww[1] * ww[1] + ww[2] * ww[2]);
    if (nw > 0.0) {
        const double sc = acos(fabs(c) / nt) / nw;
        for (int i = 0; i < 3; i++) ww[i] *= sc;
    }
    rodrigues_to_mat(ww, wR);
    transpose33(r_r, r_rT);
    mul33(wR, r_rT, Ri1);  // R1 = wR * r_r^T
    mul33(wR, r_r, Ri2);   // R2 = wR * r_r
    for (int i = 0; i < 3; i++) t[i] = Ri2[i][0] * T[0] + Ri2[i][1] * T[1] + Ri2[i][2] * T[2];

    // new focal length and principal points
    double fc_new = DBL_MAX;
    const Matrix *A[2] = {&M1, &M2}, *Dk[2] = {&D1, &D2};
    for (int k = 0; k < 2; k++) {
        const double dk1 = Dk[k]->v.empty() ? 0 : Dk[k]->v[0];
        double fc = A[k]->at(idx ^ 1, idx ^ 1);
        if (dk1 < 0) fc *= 1 + dk1 * (nx * nx + ny * ny) / (4 * fc * fc);
        fc_new = fc_new < fc ? fc_new : fc;
    }
    if (variant == RECTIFY_MODERN) fc_new = (A[0]->at(idx ^ 1, idx ^ 1) + A[1]->at(idx ^ 1, idx ^ 1)) * 0.5;
    double cc_new[2][2];
    for (int k = 0; k < 2; k++) {
        const M3 &Rk = (k == 0) ? Ri1 : Ri2;
        // cvProjectPoints2 converts the rotation matrix to a vector and back
        double rv[3];
        M3 Rp;
        rodrigues_to_vec(Rk, rv);
        rodrigues_to_mat(rv, Rp);
        double ax = 0, ay = 0;
        for (int i = 0; i < 4; i++) {
            const int j = (i < 2) ? 0 : 1;
            float px = (float)((i % 2) * (nx - 1)), py = (float)(j * (ny - 1));
            undistort_normalized(px, py, *A[k], *Dk[k]);
            const double X = Rp[0][0] * px + Rp[0][1] * py + Rp[0][2], Y = Rp[1][0] * px + Rp[1][1] * py + Rp[1][2],
                         Z = Rp[2][0] * px + Rp[2][1] * py + Rp[2][2];
            const double z = Z ? 1. / Z : 1;
            ax += (double)(float)(fc_new * (X * z));  // projected points are stored as float
            ay += (double)(float)(fc_new * (Y * z));
        }
        if (variant == RECTIFY_MODERN) {
            cc_new[k][0] = (nx - 1) * 0.5 - ax / 4;
            cc_new[k][1] = (ny - 1) * 0.5 - ay / 4;
        } else {
            cc_new[k][0] = (nx - 1) / 2 - ax / 4;  // (nx-1)/2 is integer division in the 2.4 source
            cc_new[k][1] = (ny - 1) / 2 - ay / 4;
        }
    }
    if (idx == 0)  // horizontal stereo, flags == 0
        cc_new[0][1] = cc_new[1][1] = (cc_new[0][1] + cc_new[1][1]) * 0.5;
    else
        cc_new[0][0] = cc_new[1][0] = (cc_new[0][0] + cc_new[1][0]) * 0.5;

    RectifyResult out;
    out.R1 = to_matrix(Ri1);
    out.R2 = to_matrix(Ri2);
    double pp2_t = t[idx] * fc_new;  // baseline * focal length
    // alpha < 0: no rescaling (s = 1), but the principal points pass through newSize*c/oldSize
    const double cx1 = nx * cc_new[0][0] / nx, cy1 = ny * cc_new[0][1] / ny;
    const double cx2 = nx * cc_new[1][0] / nx, cy2 = ny * cc_new[1][1] / ny;
    const double s = 1.;
    fc_new *= s;
    out.P1 = Matrix(3, 4);
    out.P2 = Matrix(3, 4);
    out.P1.at(0, 0) = out.P1.at(1, 1) = fc_new;
    out.P1.at(0, 2) = cx1;
    out.P1.at(1, 2) = cy1;
    out.P1.at(2, 2) = 1;
    out.P2.at(0, 0) = out.P2.at(1, 1) = fc_new;
    out.P2.at(0, 2) = cx2;
    out.P2.at(1, 2) = cy2;
    out.P2.at(2, 2) = 1;
    out.P2.at(idx, 3) = s * pp2_t;
    out.Q = Matrix(4, 4);
    const double q[16] = {1, 0, 0, -cx1, 0, 1, 0, -cy1, 0, 0, 0, fc_new, 0, 0, -1. / t[idx],
                          (idx == 0 ? cx1 - cx2 : cy1 - cy2) / t[idx]};
    out.Q.v.assign(q, q + 16);
    return out;
}

void init_undistort_rectify_map(const Matrix &M, const Matrix &D, const Matrix &R, const Matrix &P, Size size,
                                std::vector<int16_t> &map1, std::vector<uint16_t> &map2)
{
    // iR = (P[:, :3] * R)^-1, closed-form 3x3 inverse (cv::invert's 3x3 branch)
    double S[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) S[i][j] = P.at(i, 0) * R.at(0, j) + P.at(i, 1) * R.at(1, j) + P.at(i, 2) * R.at(2, j);
    double d = S[0][0] * (S[1][1] * S[2][2] - S[1][2] * S[2][1]) - S[0][1] * (S[1][0] * S[2][2] - S[1][2] * S[2][0]) +
               S[0][2] * (S[1][0] * S[2][1] - S[1][1] * S[2][0]);
    double ir[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (d != 0.) {
        d = 1. / d;
        ir[0] = (S[1][1] * S[2][2] - S[1][2] * S[2][1]) * d;
        ir[1] = (S[0][2] * S[2][1] - S[0][1] * S[2][2]) * d;
        ir[2] = (S[0][1] * S[1][2] - S[0][2] * S[1][1]) * d;
        ir[3] = (S[1][2] * S[2][0] - S[1][0] * S[2][2]) * d;
        ir[4] = (S[0][0] * S[2][2] - S[0][2] * S[2][0]) * d;
        ir[5] = (S[0][2] * S[1][0] - S[0][0] * S[1][2]) * d;
        ir[6] = (S[1][0] * S[2][1] - S[1][1] * S[2][0]) * d;
        ir[7] = (S[0][1] * S[2][0] - S[0][0] * S[2][1]) * d;
        ir[8] = (S[0][0] * S[1][1] - S[0][1] * S[1][0]) * d;
    }
    const double u0 = M.at(0, 2), v0 = M.at(1, 2), fx = M.at(0, 0), fy = M.at(1, 1);
    double k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < (int)D.v.size() && i < 8; i++) k[i] = D.v[i];
    const double k1 = k[0], k2 = k[1], p1 = k[2], p2 = k[3], k3 = k[4], k4 = k[5], k5 = k[6], k6 = k[7];
    const int W = size.width, H = size.height;
    map1.assign((size_t)W * H * 2, 0);
    map2.assign((size_t)W * H, 0);
    for (int i = 0; i < H; i++) {
        double _x = i * ir[1] + ir[2], _y = i * ir[4] + ir[5], _w = i * ir[7] + ir[8];
        for (int j = 0; j < W; j++, _x += ir[0], _y += ir[3], _w += ir[6]) {
            const double w = 1. / _w, x = _x * w, y = _y * w;
            const double x2 = x * x, y2 = y * y, r2 = x2 + y2, _2xy = 2 * x * y;
            const double kr = (1 + ((k3 * r2 + k2) * r2 + k1) * r2) / (1 + ((k6 * r2 + k5) * r2 + k4) * r2);
            const double u = fx * (x * kr + p1 * _2xy + p2 * (r2 + 2 * x2)) + u0;
            const double v = fy * (y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy) + v0;
            const int iu = (int)lrint(u * 32), iv = (int)lrint(v * 32);  // saturate_cast<int> == cvRound
            map1[((size_t)i * W + j) * 2 + 0] = (int16_t)(iu >> 5);
            map1[((size_t)i * W + j) * 2 + 1] = (int16_t)(iv >> 5);
            map2[(size_t)i * W + j] = (uint16_t)((iv & 31) * 32 + (iu & 31));
        }
    }
}

}  // namespace duke
