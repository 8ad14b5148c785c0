// mfreconstruct.h — MFReconstruct with the reference's public surface (Duke/mfreconstruct.h:12-20); the decode,
// match and triangulation run on the B200 through the C ABI of libslr_b200.so.
#pragma once
#include <string>

#include "pointcloudimage.h"
#include "slr_b200.h"
#include "stereorect.h"
#include "virtualcamera.h"

class MFReconstruct {
public:
    explicit MFReconstruct(void *parent = 0);
    ~MFReconstruct();
    // scansn also names the output model; blackt / whitet as in the Set dialog (Duke/mainwindow.cpp:594)
    void getParameters(int scansn, int scanw, int scanh, int camw, int camh, int blackt, int whitet,
                       const std::string &savePath);
    bool runReconstruction();

    PointCloudImage *points3DProjView;

    // ---- additions (not in the reference) ----
    int mode = SLR_MODE_STRICT;   // SLR_MODE_CORRECTED selects the atan2 + heterodyne decode
    int device = 0;               // CUDA device of the engine
    stereoRect *rectifier() { return sr; }
    unsigned long long pointCount() const { return n_points_; }

private:
    int scanSN = 0;
    std::string savePath_;
    std::string calibFolder[2], scanFolder[2], imgPrefix[2], imgSuffix;
    int numberOfImgs;
    int blackThreshold = 40, whiteThreshold = 0;
    bool pathSet = false, camerasLoaded = false;
    int cameraWidth = 0, cameraHeight = 0, scan_w = 0, scan_h = 0;
    VirtualCamera *cameras;
    stereoRect *sr = nullptr;
    unsigned long long n_points_ = 0;
    bool loadCameras();
};
