// reconstruct.h — Reconstruct with the reference's public surface (Duke/reconstruct.h:14-42): Gray-code decode,
// match and triangulation on the B200 through the C ABI of libslr_b200.so.
#pragma once
#include <string>

#include "pointcloudimage.h"
#include "slr_b200.h"
#include "stereorect.h"
#include "virtualcamera.h"

class Reconstruct {
public:
    Reconstruct(bool useEpi);
    ~Reconstruct();

    bool loadCameras();
    bool runReconstruction();      // GRAY_ONLY: column + row codes, ray-ray triangulation (reconstruct.cpp:230-265)
    bool runReconstruction_GE();   // GRAY_EPI: column codes on rectified images, Q reprojection (:271-307)

    VirtualCamera *cameras;
    std::string *calibFolder;
    PointCloudImage *points3DProjView;
    void setBlackThreshold(int val);
    void setWhiteThreshold(int val);
    void setCalibPath(const std::string &path1st, int cam_no);
    void enableRaySampling();
    void disableRaySampling();
    void cam2WorldSpace(VirtualCamera cam, duke::Point3f &p);
    void getParameters(int scanw, int scanh, int camw, int camh, bool autocontrast, bool havecolor,
                       const std::string &savePath);
    std::string savePath_;
    int scanSN = 0;   // assign BEFORE getParameters (the image prefix is built there, reconstruct.cpp:645-648)

    int device = 0;   // addition: CUDA device of the engine
    stereoRect *rectifier() { return sr; }

private:
    bool EPI;
    stereoRect *sr = nullptr;
    std::string scanFolder[2], imgPrefix[2], imgSuffix;
    int blackThreshold = 40, whiteThreshold = 0;
    bool pathSet = false, autoContrast_ = false, raySampling_ = false, haveColor = false;
    int cameraWidth = 0, cameraHeight = 0, scan_w = 0, scan_h = 0;
};
