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
    // The cloud as the GPU pipeline delivers it (slr_run_mf_ingested): sums = float [imageH][imageW][3], counts = uint8
    // [imageH][imageW] already in this class's storage layout, in (pinned) memory the object keeps until it is
    // destroyed, when release(sums, counts) hands it back.  No host pass over the 17 MB of a 1280x1024 cloud.
    // cell_gray (may be NULL: no colour) = uint8 [imageH][imageW], the grey value every cell with a point received as
    // its (g, g, g) colour (Reconstruct::triangulation_ge with haveColor, Duke/reconstruct.cpp:596-603); copied.
    PointCloudImage(int imageW, int imageH, float *sums, uint8_t *counts, void (*release)(float *, uint8_t *),
                    const uint8_t *cell_gray = nullptr);
    ~PointCloudImage();
    PointCloudImage(const PointCloudImage &) = delete;
    PointCloudImage &operator=(const PointCloudImage &) = delete;

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
    template <typename T>
    struct View {
        const T *p;
        size_t n;
        const T *data() const { return p; }
        size_t size() const { return n; }
        const T &operator[](size_t i) const { return p[i]; }
    };
    View<float> sums() const { return {points_, (size_t)w * h * 3}; }
    View<uint8_t> counts() const { return {num_, (size_t)w * h}; }

private:
    int w, h;
    bool has_color_;
    float *points_;               // h x w x 3
    uint8_t *num_;                // h x w
    std::vector<float> own_points_;   // backing store unless adopted
    std::vector<uint8_t> own_num_;
    void (*release_)(float *, uint8_t *) = nullptr;
    std::vector<int> color_;      // h x w x 3 (sums; the reference keeps CV_8UC3, which saturates per add)
};
