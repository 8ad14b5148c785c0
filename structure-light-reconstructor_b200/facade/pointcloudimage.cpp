#include "pointcloudimage.h"

#include <math.h>

#include <fstream>
#include <thread>

PointCloudImage::PointCloudImage(int imageW, int imageH, bool colorFlag)
    : w(imageW), h(imageH), has_color_(colorFlag), own_points_((size_t)imageW * imageH * 3, 0.0f), own_num_((size_t)imageW * imageH, 0)
{
    points_ = own_points_.data();
    num_ = own_num_.data();
    if (colorFlag) color_.assign((size_t)imageW * imageH * 3, 0);
}
PointCloudImage::PointCloudImage(int imageW, int imageH, float *sums, uint8_t *counts, void (*release)(float *, uint8_t *),
                                 const uint8_t *cell_gray)
    : w(imageW), h(imageH), has_color_(cell_gray != nullptr), points_(sums), num_(counts), release_(release)
{
    if (!cell_gray) return;
    const size_t cells = (size_t)w * h;
    color_.resize(cells * 3);
    auto band = [&](size_t c0, size_t c1) {
        for (size_t c = c0; c < c1; c++) color_[3 * c] = color_[3 * c + 1] = color_[3 * c + 2] = cell_gray[c];
    };
    unsigned nt = std::thread::hardware_concurrency();
    nt = nt == 0 ? 1 : nt > 8 ? 8 : nt;
    if (cells < 65536) nt = 1;
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nt; t++) pool.emplace_back(band, cells * t / nt, cells * (t + 1) / nt);
    band(0, cells / nt);
    for (auto &t : pool) t.join();
}
PointCloudImage::~PointCloudImage()
{
    if (release_) release_(points_, num_);
}

bool PointCloudImage::setPoint(int i_w, int j_h, duke::Point3f point, duke::Vec3i colorgray)
{
    if (i_w >= w || j_h >= h) return false;
    setPoint(i_w, j_h, point);
    if (has_color_)
        for (int k = 0; k < 3; k++) {
            const int v = colorgray[k];
            color_[((size_t)j_h * w + i_w) * 3 + k] = v < 0 ? 0 : v > 255 ? 255 : v;  // CV_8UC3 store saturates
        }
    return true;
}

bool PointCloudImage::setPoint(int i_w, int j_h, duke::Point3f point)
{
    if (i_w >= w || j_h >= h) return false;
    const size_t q = (size_t)j_h * w + i_w;  // matSet3D(points, i_w, j_h): Mat(row = j_h, col = i_w)
    points_[q * 3 + 0] = point.x;
    points_[q * 3 + 1] = point.y;
    points_[q * 3 + 2] = point.z;
    num_[q] = 1;
    return true;
}

bool PointCloudImage::getPoint(int i_w, int j_h, duke::Point3f &pointOut, duke::Vec3i &colorOut)
{
    if (i_w >= w || j_h >= h) return false;
    const size_t q = (size_t)j_h * w + i_w;
    const uint8_t num = num_[q];
    if (num == 0) return false;
    getPoint(i_w, j_h, pointOut);
    if (has_color_) {
        const float inv = 1.f / (float)num;
        for (int k = 0; k < 3; k++) colorOut[k] = (int)lrintf((float)((double)color_[q * 3 + k] * inv));
    } else {
        colorOut = duke::Vec3i(100, 0, 0);  // `(cv::Point3i)(100,100,100)` is a comma expression: (100, 0, 0)
    }
    return true;
}

bool PointCloudImage::getPoint(int i_w, int j_h, duke::Point3f &pointOut)
{
    if (i_w >= w || j_h >= h) return false;
    const size_t q = (size_t)j_h * w + i_w;
    const uint8_t num = num_[q];
    if (num == 0) return false;
    const float inv = 1.f / (float)num;  // Vec3d / float == multiply by 1.f/alpha (OpenCV 2.4 operator/)
    pointOut.x = (float)((double)points_[q * 3 + 0] * inv);
    pointOut.y = (float)((double)points_[q * 3 + 1] * inv);
    pointOut.z = (float)((double)points_[q * 3 + 2] * inv);
    return true;
}

bool PointCloudImage::addPoint(int i_w, int j_h, duke::Point3f point, duke::Vec3i colorgray)
{
    if (i_w >= w || j_h >= h) return false;
    const size_t q = (size_t)j_h * w + i_w;
    if (num_[q] == 0) return setPoint(i_w, j_h, point, colorgray);
    addPoint(i_w, j_h, point);
    if (!has_color_) return false;
    for (int k = 0; k < 3; k++) {
        const int v = colorgray[k] + color_[q * 3 + k];
        color_[q * 3 + k] = v < 0 ? 0 : v > 255 ? 255 : v;
    }
    return true;
}

bool PointCloudImage::addPoint(int i_w, int j_h, duke::Point3f point)
{
    if (i_w >= w || j_h >= h) return false;
    const size_t q = (size_t)j_h * w + i_w;
    if (num_[q] == 0) return setPoint(i_w, j_h, point);
    points_[q * 3 + 0] = point.x + points_[q * 3 + 0];
    points_[q * 3 + 1] = point.y + points_[q * 3 + 1];
    points_[q * 3 + 2] = point.z + points_[q * 3 + 2];
    num_[q] = (uint8_t)(num_[q] + 1);
    return true;
}

void PointCloudImage::addDense(const float *xyz, const uint8_t *valid, const uint8_t *gray, int W, int H)
{
    // addPoint(i, j, p) lands in cell (i_w = i, j_h = j): image rows map to distinct cells, so bands of rows can be
    // filled by different host threads and every cell still sees its additions in the reference's order
    auto band = [&](int i0, int i1) {
        for (int i = i0; i < i1; i++)
            for (int j = 0; j < W; j++) {
                const size_t p = (size_t)i * W + j;
                if (!valid[p]) continue;
                const duke::Point3f pt(xyz[p * 3], xyz[p * 3 + 1], xyz[p * 3 + 2]);
                if (gray && has_color_)
                    addPoint(i, j, pt, duke::Vec3i(gray[p], gray[p], gray[p]));
                else
                    addPoint(i, j, pt);
            }
    };
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (nt > 16) nt = 16;
    if ((size_t)W * H < 65536) nt = 1;
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nt; t++) pool.emplace_back(band, (int)((long long)H * t / nt), (int)((long long)H * (t + 1) / nt));
    band(0, (int)((long long)H / nt));
    for (auto &t : pool) t.join();
}

void PointCloudImage::exportXYZ(const char *path, bool exportOffPixels, bool colorFlag)
{
    std::ofstream out(path);
    duke::Point3f p;
    duke::Vec3i c;
    for (int i = 0; i < w; i++)
        for (int j = 0; j < h; j++) {
            const uint8_t num = num_[(size_t)j * w + i];
            if (!exportOffPixels && num == 0) continue;
            getPoint(i, j, p, c);
            if (exportOffPixels && num == 0) {
                p = duke::Point3f(0, 0, 0);
                c = duke::Vec3i(0, 0, 0);
            }
            out << p.x << " " << p.y << " " << p.z;
            if (colorFlag && has_color_)
                out << " " << c[2] << " " << c[1] << " " << c[0] << "\n";
            else
                out << "\n";
        }
}

int PointCloudImage::getWidth() { return w; }
int PointCloudImage::getHeight() { return h; }
