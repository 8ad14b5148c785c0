// pointcloudimage.h — PointCloudImage with the reference's interface (Duke/pointcloudimage.h:8-35).
// Accumulator semantics of Duke/pointcloudimage.cpp: float sums, a u8 count that wraps, (i_w, j_h) addressing
// with the `i_w >= w || j_h >= h -> drop` rule, getPoint = sum * (1.f / count) (OpenCV's Vec / float).
#pragma once
#include <stdint.h>

#include <vector>

#include "duke_types.h"

class PointCloudImage {
public:
    PointCloudImage(int imageW, int imageH, bool color);
    ~PointCloudImage();

    bool setPoint(int i_w, int j_h, duke::Point3f point, duke::Vec3i colorgray);
    bool setPoint(int i_w, int j_h, duke::Point3f point);
    bool getPoint(int i_w, int j_h, duke::Point3f &pointOut);
    bool getPoint(int i_w, int j_h, duke::Point3f &pointOut, duke::Vec3i &colorgray);
    bool addPoint(int i_w, int j_h, duke::Point3f point, duke::Vec3i colorgray);
    bool addPoint(int i_w, int j_h, duke::Point3f point);
    void exportXYZ(const char *path, bool exportOffPixels = true, bool colorFlag = true);
    int getWidth();
    int getHeight();

    // bulk fill from the engine's dense cloud, applying addPoint(row, col, p) for every valid pixel in the
    // reference's row-major order (mfreconstruct.cpp:284-326, reconstruct.cpp:555-603)
    void addDense(const float *xyz, const uint8_t *valid, const uint8_t *gray, int W, int H);
    // raw views (h x w, element (j_h, i_w))
    const std::vector<float> &sums() const { return points_; }
    const std::vector<uint8_t> &counts() const { return num_; }

private:
    int w, h;
    bool has_color_;
    std::vector<float> points_;   // h x w x 3
    std::vector<uint8_t> num_;    // h x w
    std::vector<int> color_;      // h x w x 3 (sums; the reference keeps CV_8UC3, which saturates per add)
};
