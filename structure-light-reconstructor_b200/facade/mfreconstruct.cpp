#include "mfreconstruct.h"

#include <chrono>

#include <stdio.h>

#include "ingest.h"
#include "reconstruct_common.h"

MFReconstruct::MFReconstruct(void *) : points3DProjView(nullptr), imgSuffix(".png"), numberOfImgs(14)
{
    cameras = new VirtualCamera[2];
}

MFReconstruct::~MFReconstruct()
{
    delete[] cameras;
    delete sr;
    delete points3DProjView;
}

void MFReconstruct::getParameters(int scansn, int scanw, int scanh, int camw, int camh, int blackt, int whitet,
                                  const std::string &savePath)
{
    scanSN = scansn;
    scan_w = scanw;
    scan_h = scanh;
    cameraWidth = camw;
    cameraHeight = camh;
    blackThreshold = blackt;
    whiteThreshold = whitet;
    savePath_ = savePath;
    delete sr;
    sr = new stereoRect(savePath, duke::Size(camw, camh));
    sr->getParameters();
    for (int i = 0; i < 2; i++) {
        const char *side = i == 0 ? "left" : "right";
        scanFolder[i] = savePath + "/scan/" + side + "/";
        imgPrefix[i] = std::to_string(scanSN) + (i == 0 ? "/L" : "/R");
        calibFolder[i] = savePath + "/calib/" + side + "/";
    }
    pathSet = true;
    camerasLoaded = loadCameras();
    if (!camerasLoaded) fprintf(stderr, "Get Param: Load Calibration files failed.\n");
}

bool MFReconstruct::loadCameras()
{
    bool loaded = false;
    for (int i = 0; i < 2; i++) {
        loaded = cameras[i].loadCameraMatrix(calibFolder[i] + "cam_matrix.txt");
        if (!loaded) break;
        cameras[i].loadDistortion(calibFolder[i] + "cam_distortion.txt");
        cameras[i].loadRotationMatrix(calibFolder[i] + "cam_rotation_matrix.txt");
        cameras[i].loadTranslationVector(calibFolder[i] + "cam_trans_vectror.txt");
        cameras[i].loadFundamentalMatrix(savePath_ + "/calib/fundamental_stereo.txt");
        cameras[i].height = cameraHeight;
        cameras[i].width = cameraWidth;
    }
    return loaded;
}

bool MFReconstruct::runReconstruction()
{
    if (!pathSet || !camerasLoaded || !sr || !sr->loaded()) {
        fprintf(stderr, "MFReconstruct: calibration is not loaded\n");
        return false;
    }
    const int W = cameraWidth, H = cameraHeight;
    const size_t P = (size_t)W * H;
    // DUKE_TIMING=1: where the wall time of a drop-in call goes (stderr)
    const bool timing = getenv("DUKE_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_prev = now();
    auto lap = [&](const char *what) {
        const double t = now();
        if (timing) fprintf(stderr, "[MFReconstruct] %-34s %8.1f ms\n", what, t - t_prev);
        t_prev = t;
    };
    sr->calParameters();
    lap("stereoRect::calParameters (host)");

    slr_engine *eng = duke::shared_engine(device, W, H);
    if (!eng) {
        fprintf(stderr, "MFReconstruct: %s\n", slr_last_error());
        return false;
    }
    bool ok = false;
    void *h_stack = nullptr, *h_xyz = nullptr, *h_valid = nullptr;
    do {
        slr_camera cams[2] = {duke::to_slr_camera(cameras[0]), duke::to_slr_camera(cameras[1])};
        float rigid[12];
        const float *rg = nullptr;
        if (scanSN > 0) {  // mfreconstruct.cpp:276-282
            if (duke::load_rigid(savePath_ + "/scan/transfer_mat" + std::to_string(scanSN) + ".txt", rigid)) rg = rigid;
        }
        // calibration and the 16 MB of rectification maps go to the GPU once per session, not once per scan
        if (!duke::upload_calibration(eng, cams, sr->Q.v.data(), rg, sr->calibrationId(), sr->map1().data(), sr->map2().data())) break;
        if (slr_set_host_input_raw(eng, 1) != SLR_OK) break;
        if (slr_set_auto_contrast(eng, 0) != SLR_OK) break;   // MFReconstruct has no such setting; the engine is shared
        lap("CUDA context + engine + calib");
        if (W % 4 == 0 && !getenv("DUKE_HOST_DECODE")) {
            // files -> GPU: inflate on all host threads, upload + PNG unfiltering + rectification + decode + match +
            // triangulation on the GPU, the cloud back in PointCloudImage's own storage layout (ingest.h)
            std::string err;
            if (!duke::ingest_scan(eng, scanFolder, imgPrefix, imgSuffix, numberOfImgs, W, H, &err)) {
                fprintf(stderr, "%s\n", err.c_str());
                slr_ingest_abort(eng);
                break;
            }
            lap("image files -> inflate -> GPU");
            float *sums = nullptr;
            uint8_t *counts = nullptr;
            if (!duke::cloud_storage_acquire((size_t)scan_w * scan_h, &sums, &counts)) break;
            n_points_ = 0;
            if (slr_run_mf_ingested(eng, 3, 4, blackThreshold, mode, scan_w, scan_h, sums, counts, nullptr, nullptr, &n_points_) != SLR_OK) {
                duke::cloud_storage_release(sums, counts);
                break;
            }
            lap("slr_run_mf_ingested (fused kernel, D2H)");
            delete points3DProjView;
            points3DProjView = new PointCloudImage(scan_w, scan_h, sums, counts, duke::cloud_storage_release);
            lap("PointCloudImage (adopts the buffers)");
            ok = true;
            break;
        }
        if (!(h_stack = duke::pinned_scratch(0, 2 * (size_t)numberOfImgs * P))) break;
        if (!(h_xyz = duke::pinned_scratch(1, P * 3 * sizeof(float)))) break;
        if (!(h_valid = duke::pinned_scratch(2, P))) break;
        const bool loaded = duke::load_stacks(scanFolder, imgPrefix, imgSuffix, 2, numberOfImgs, W, H, (uint8_t *)h_stack);
        if (!loaded) break;
        lap("image files -> pinned stack");
        n_points_ = 0;
        if (slr_run_mf_host(eng, (const uint8_t *)h_stack, 1, 3, 4, blackThreshold, mode, (float *)h_xyz, (uint8_t *)h_valid,
                            nullptr, &n_points_) != SLR_OK)
            break;
        lap("slr_run_mf_host (H2D, fused, D2H)");
        delete points3DProjView;
        points3DProjView = new PointCloudImage(scan_w, scan_h, false);
        points3DProjView->addDense((const float *)h_xyz, (const uint8_t *)h_valid, nullptr, W, H);
        lap("PointCloudImage::addDense");
        ok = true;
    } while (false);
    if (!ok && slr_last_error()[0]) fprintf(stderr, "MFReconstruct: %s\n", slr_last_error());
    return ok;   // engine and pinned buffers stay with the process (reconstruct_common.h)
}
