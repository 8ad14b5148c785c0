// rectify.h — the calibration-side arithmetic behind stereoRect::calParameters (Duke/stereorect.cpp:36-44):
// cv::stereoRectify(M1, D1, M2, D2, size, R, T, R1, R2, P1, P2, Q, flags = 0, alpha = -1) and
// cv::initUndistortRectifyMap(M, D, R, P, size, CV_16SC2, map1, map2), restated from the published OpenCV 2.4
// algorithm (third-party code the reference links: opencv_calib3d249 / opencv_imgproc249, Duke/Duke.pro:51-57).
// Host code: it runs once per calibration on 3x3 matrices; the per-pixel remap itself is a CUDA kernel
// (slr_rectify_stack).  Parity for this third-party arithmetic is checked against cv2 4.13 in the tests
// (tolerance, not bit-exactness: different OpenCV versions).
#pragma once
#include <stdint.h>

#include <vector>

#include "duke_types.h"

namespace duke {

struct RectifyResult {
    Matrix R1, R2, P1, P2, Q;  // 3x3, 3x3, 3x4, 3x4, 4x4
};

// Two details of cvStereoRectify changed after the 2.4 series the reference links; the product follows 2.4.9,
// the MODERN variant exists so the rest of the restatement can be validated against the cv2 4.x of this image:
//   new focal length   2.4.9: min over cameras of fy, shrunk by 1 + k1 (nx^2+ny^2)/(4 fy^2) when k1 < 0
//                      4.x  : mean of the two fy
//   principal point    2.4.9: (nx-1)/2 in INTEGER division        4.x: (nx-1)*0.5
enum RectifyVariant { RECTIFY_CV249 = 0, RECTIFY_MODERN = 1 };

// M: 3x3 camera matrix, D: 5x1 (k1 k2 p1 p2 k3), R: 3x3, T: 3x1
RectifyResult stereo_rectify(const Matrix &M1, const Matrix &D1, const Matrix &M2, const Matrix &D2, Size size,
                             const Matrix &R, const Matrix &T, RectifyVariant variant = RECTIFY_CV249);

// CV_16SC2 fixed-point maps: map1 = [H][W][2] int16 (x, y integer parts), map2 = [H][W] uint16 (5+5 fractional bits)
void init_undistort_rectify_map(const Matrix &M, const Matrix &D, const Matrix &R, const Matrix &P, Size size,
                                std::vector<int16_t> &map1, std::vector<uint16_t> &map2);

}  // namespace duke
