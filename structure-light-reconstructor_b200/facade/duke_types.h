// duke_types.h — the few value types the reference's public surface carries (cv::Point3f, cv::Vec3i,
// cv::Size, a small dense matrix for the text-file matrices, QString -> std::string), so that the facade
// classes keep the reference's member names without Qt or OpenCV.  See INTEGRATION.md for the two-line
// adapter a Qt/OpenCV build adds on top.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace duke {

struct Point2f { float x = 0, y = 0; Point2f() {} Point2f(float a, float b) : x(a), y(b) {} };
struct Point3f { float x = 0, y = 0, z = 0; Point3f() {} Point3f(float a, float b, float c) : x(a), y(b), z(c) {} };
struct Vec3i { int v[3] = {0, 0, 0}; Vec3i() {} Vec3i(int a, int b, int c) { v[0] = a; v[1] = b; v[2] = c; } int &operator[](int i) { return v[i]; } int operator[](int i) const { return v[i]; } };
struct Size { int width = 0, height = 0; Size() {} Size(int w, int h) : width(w), height(h) {} };

// row-major dense matrix of doubles; holds what VirtualCamera::loadMatrix / stereoRect::loadMatrix read
struct Matrix {
    int rows = 0, cols = 0;
    std::vector<double> v;
    Matrix() {}
    Matrix(int r, int c) : rows(r), cols(c), v((size_t)r * c, 0.0) {}
    bool empty() const { return v.empty(); }
    double &at(int r, int c) { return v[(size_t)r * cols + c]; }
    double at(int r, int c) const { return v[(size_t)r * cols + c]; }
};

// 8-bit single-channel image
struct Image {
    int width = 0, height = 0;
    std::vector<uint8_t> pix;
    bool empty() const { return pix.empty(); }
};

}  // namespace duke
