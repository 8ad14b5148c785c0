// stereorect.h — stereoRect with the reference's interface (Duke/stereorect.h:13-29): loads the stereo
// calibration text files, computes R1/R2/P1/P2/Q and the CV_16SC2 rectification maps (cv::stereoRectify +
// cv::initUndistortRectifyMap restated in rectify.cpp), and rectifies images.  In the drop-in the per-image
// remap runs on the GPU (slr_rectify_stack via slr_set_host_input_raw); doStereoRectify() is kept for callers
// that rectify a single image (e.g. DotMatch) and uses the same fixed-point arithmetic on the host.
#pragma once
#include <stdint.h>

#include <memory>
#include <string>
#include <vector>

#include "duke_types.h"

class stereoRect {
public:
    stereoRect(const std::string &projectPath, duke::Size size);
    void doStereoRectify(duke::Image &img, bool isleft);
    void getParameters();   // reads calib/{left,right}/cam_stereo.txt, distortion_stereo.txt, calib/R_stereo.txt, T_stereo.txt
    void calParameters();   // stereoRectify(flags = 0, alpha = -1) + 2 x initUndistortRectifyMap(CV_16SC2)
    duke::Matrix R1, P1, R2, P2, Q;

    // maps in the layout slr_set_rectify_maps takes: [2][H][W][2] int16 and [2][H][W] uint16
    const std::vector<int16_t> &map1() const { return *map1_; }
    const std::vector<uint16_t> &map2() const { return *map2_; }
    bool loaded() const { return loaded_; }
    // identifies the result of calParameters(): equal ids = the same maps and Q (lets callers skip re-uploading 16 MB of
    // maps for every scan of a session); 0 before calParameters()
    unsigned long long calibrationId() const { return cal_id_; }

private:
    std::string ppath;
    duke::Size img_size;
    duke::Matrix M1, D1, M2, D2, R, T;
    // shared with the process-wide cache of the last calParameters() result (15.7 MB at 1280x1024: not copied per scan)
    std::shared_ptr<const std::vector<int16_t>> map1_ = std::make_shared<std::vector<int16_t>>();
    std::shared_ptr<const std::vector<uint16_t>> map2_ = std::make_shared<std::vector<uint16_t>>();
    unsigned long long cal_id_ = 0;
    bool loaded_ = false;
    bool loadMatrix(duke::Matrix &matrix, int rows, int cols, const std::string &file);
};
