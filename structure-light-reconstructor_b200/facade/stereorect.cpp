#include "stereorect.h"

#include <mutex>
#include <thread>

#include <fstream>

#include "rectify.h"

stereoRect::stereoRect(const std::string &projectPath, duke::Size size) : ppath(projectPath), img_size(size) {}

void stereoRect::getParameters()
{
    bool ok = true;
    ok &= loadMatrix(M1, 3, 3, ppath + "/calib/left/cam_stereo.txt");
    ok &= loadMatrix(D1, 5, 1, ppath + "/calib/left/distortion_stereo.txt");
    ok &= loadMatrix(M2, 3, 3, ppath + "/calib/right/cam_stereo.txt");
    ok &= loadMatrix(D2, 5, 1, ppath + "/calib/right/distortion_stereo.txt");
    ok &= loadMatrix(R, 3, 3, ppath + "/calib/R_stereo.txt");
    ok &= loadMatrix(T, 3, 1, ppath + "/calib/T_stereo.txt");
    loaded_ = ok;
}

namespace {
// The reference recomputes stereoRectify + both map pairs on every loadCamImgs (mfreconstruct.cpp:117): a pure
// function of the six calibration matrices and the image size, so the last result is kept for the process.
struct CalCache {
    std::vector<double> key;
    duke::RectifyResult r;
    std::shared_ptr<const std::vector<int16_t>> map1;
    std::shared_ptr<const std::vector<uint16_t>> map2;
    unsigned long long id = 0;
};
std::mutex g_cal_mu;
CalCache g_cal;
}  // namespace

void stereoRect::calParameters()
{
    if (!loaded_) return;
    std::vector<double> key = {(double)img_size.width, (double)img_size.height};
    for (const duke::Matrix *m : {&M1, &D1, &M2, &D2, &R, &T}) key.insert(key.end(), m->v.begin(), m->v.end());
    std::lock_guard<std::mutex> lk(g_cal_mu);
    if (g_cal.key != key) {
        g_cal.r = duke::stereo_rectify(M1, D1, M2, D2, img_size, R, T);
        std::vector<int16_t> a1, b1;
        std::vector<uint16_t> a2, b2;
        std::thread right([&] { duke::init_undistort_rectify_map(M2, D2, g_cal.r.R2, g_cal.r.P2, img_size, b1, b2); });
        duke::init_undistort_rectify_map(M1, D1, g_cal.r.R1, g_cal.r.P1, img_size, a1, a2);
        right.join();
        a1.insert(a1.end(), b1.begin(), b1.end());
        a2.insert(a2.end(), b2.begin(), b2.end());
        g_cal.map1 = std::make_shared<const std::vector<int16_t>>(std::move(a1));
        g_cal.map2 = std::make_shared<const std::vector<uint16_t>>(std::move(a2));
        g_cal.key = key;
        g_cal.id++;
    }
    R1 = g_cal.r.R1;
    R2 = g_cal.r.R2;
    P1 = g_cal.r.P1;
    P2 = g_cal.r.P2;
    Q = g_cal.r.Q;
    map1_ = g_cal.map1;
    map2_ = g_cal.map2;
    cal_id_ = g_cal.id;
}

void stereoRect::doStereoRectify(duke::Image &img, bool isleft)
{
    // cv::remap(INTER_LINEAR), CV_16SC2 maps, BORDER_CONSTANT 0 — same fixed point as k0_rectify.cu
    const int W = img_size.width, H = img_size.height;
    if (img.empty() || img.width != W || img.height != H || map2_->empty()) return;
    const std::vector<int16_t> &map1_ = *this->map1_;
    const std::vector<uint16_t> &map2_ = *this->map2_;
    const size_t P = (size_t)W * H, off = isleft ? 0 : P;
    std::vector<uint8_t> out(P);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const size_t o = (size_t)y * W + x;
            const int sx = map1_[(off + o) * 2], sy = map1_[(off + o) * 2 + 1];
            const int a = map2_[off + o] & 1023, fx = a & 31, fy = a >> 5;
            int w[4] = {(32 - fx) * (32 - fy) * 32, fx * (32 - fy) * 32, (32 - fx) * fy * 32, fx * fy * 32};
            if (a == 0) {
                w[0] = 32767;
                w[3] = 1;
            }
            int sum = 0;
            for (int t = 0; t < 4; t++) {
                const int xx = sx + (t & 1), yy = sy + (t >> 1);
                if (xx >= 0 && xx < W && yy >= 0 && yy < H) sum += img.pix[(size_t)yy * W + xx] * w[t];
            }
            out[o] = (uint8_t)((sum + (1 << 14)) >> 15);
        }
    img.pix.swap(out);
}

bool stereoRect::loadMatrix(duke::Matrix &matrix, int rows, int cols, const std::string &file)
{
    std::ifstream in1(file.c_str());
    if (!in1) return false;
    matrix = duke::Matrix(rows, cols);
    for (int i = 0; i < rows; i++)
        for (int j = 0; j < cols; j++) {
            float val = 0;  // parsed as float, stored in a CV_64F Mat (stereorect.cpp:54-59)
            in1 >> val;
            matrix.at(i, j) = val;
        }
    return true;
}
