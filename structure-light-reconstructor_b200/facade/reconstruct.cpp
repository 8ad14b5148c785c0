#include "reconstruct.h"

#include <stdio.h>

#include <vector>

#include "ingest.h"
#include "reconstruct_common.h"

Reconstruct::Reconstruct(bool useEpi) : points3DProjView(nullptr), EPI(useEpi), imgSuffix(".png")
{
    cameras = new VirtualCamera[2];
    calibFolder = new std::string[2];
}

Reconstruct::~Reconstruct()
{
    delete points3DProjView;
    delete sr;
    delete[] cameras;
    delete[] calibFolder;
}

void Reconstruct::enableRaySampling() { raySampling_ = true; }    // stored, never read (reconstruct.cpp:33-41)
void Reconstruct::disableRaySampling() { raySampling_ = false; }
void Reconstruct::setBlackThreshold(int val) { blackThreshold = val; }
void Reconstruct::setWhiteThreshold(int val) { whiteThreshold = val; }

void Reconstruct::setCalibPath(const std::string &folder, int cam_no)
{
    calibFolder[cam_no] = folder;
    pathSet = true;
}

void Reconstruct::getParameters(int scanw, int scanh, int camw, int camh, bool autocontrast, bool havecolor,
                                const std::string &savePath)
{
    scan_w = scanw;
    scan_h = scanh;
    cameraWidth = camw;
    cameraHeight = camh;
    autoContrast_ = autocontrast;
    savePath_ = savePath;
    haveColor = havecolor;
    if (EPI) {
        delete sr;
        sr = new stereoRect(savePath_, duke::Size(camw, camh));
        sr->getParameters();
    }
    for (int i = 0; i < 2; i++) {
        scanFolder[i] = savePath + (i == 0 ? "/scan/left/" : "/scan/right/");
        imgPrefix[i] = std::to_string(scanSN) + (i == 0 ? "/L" : "/R");
    }
}

bool Reconstruct::loadCameras()
{
    bool loaded = false;
    for (int i = 0; i < 2; i++) {
        loaded = cameras[i].loadCameraMatrix(calibFolder[i] + "cam_matrix.txt");
        if (!loaded) break;
        cameras[i].loadDistortion(calibFolder[i] + "cam_distortion.txt");
        cameras[i].loadRotationMatrix(calibFolder[i] + "cam_rotation_matrix.txt");
        cameras[i].loadTranslationVector(calibFolder[i] + "cam_trans_vectror.txt");
        cameras[i].loadFundamentalMatrix(savePath_ + "/calib/fundamental_stereo.txt");
        cameras[i].loadHomoMatrix(savePath_ + "/calib/H1_mat.txt", 1);
        cameras[i].loadHomoMatrix(savePath_ + "/calib/H2_mat.txt", 2);
        cameras[i].height = 0;
        cameras[i].width = 0;
    }
    return loaded;
}

// p <- R^T p + (-R^T t): the two 3x3 * 3x1 cv::Mat products accumulate in double and narrow to float
// (reconstruct.cpp:310-322)
void Reconstruct::cam2WorldSpace(VirtualCamera cam, duke::Point3f &p)
{
    if (cam.rotationMatrix.v.size() < 9 || cam.translationVector.v.size() < 3) return;
    const float in[3] = {p.x, p.y, p.z};
    float o[3];
    for (int i = 0; i < 3; i++) {
        double st = 0.0, sp = 0.0;
        for (int k = 0; k < 3; k++) {
            st += (double)(-(float)cam.rotationMatrix.at(k, i)) * (double)(float)cam.translationVector.v[k];
            sp += (double)(float)cam.rotationMatrix.at(k, i) * (double)in[k];
        }
        o[i] = (float)st + (float)sp;
    }
    p = duke::Point3f(o[0], o[1], o[2]);
}

bool Reconstruct::runReconstruction_GE()
{
    if (!sr || !sr->loaded()) {
        fprintf(stderr, "Reconstruct: stereo calibration is not loaded\n");
        return false;
    }
    const int W = cameraWidth, H = cameraHeight;
    const size_t P = (size_t)W * H;
    const int nbits = slr_gray_num_bits(scan_w);       // GrayCodes(scan_w, scan_h, true)
    const int nimg = 2 + 2 * nbits;
    for (int i = 0; i < 2; i++) {                       // :279-280
        cameras[i].position = duke::Point3f(0, 0, 0);
        cam2WorldSpace(cameras[i], cameras[i].position);
        cameras[i].width = W;
        cameras[i].height = H;
    }
    sr->calParameters();
    slr_engine *eng = duke::shared_engine(device, W, H);
    if (!eng) {
        fprintf(stderr, "Reconstruct: %s\n", slr_last_error());
        return false;
    }
    bool ok = false;
    void *h_stack = nullptr, *h_xyz = nullptr, *h_valid = nullptr, *h_color = nullptr;
    do {
        slr_camera cams[2] = {duke::to_slr_camera(cameras[0]), duke::to_slr_camera(cameras[1])};
        float rigid[12];
        const float *rg = nullptr;
        if (scanSN > 0 && duke::load_rigid(savePath_ + "/scan/transfer_mat" + std::to_string(scanSN) + ".txt", rigid)) rg = rigid;
        if (!duke::upload_calibration(eng, cams, sr->Q.v.data(), rg, sr->calibrationId(), sr->map1().data(), sr->map2().data())) break;
        if (slr_set_host_input_raw(eng, 1) != SLR_OK) break;
        // Utilities::autoContrast on every rectified image (Duke/reconstruct.cpp:182-183), on the GPU after K0
        if (slr_set_auto_contrast(eng, autoContrast_ ? 1 : 0) != SLR_OK) break;
        unsigned long long n = 0;
        if (W % 4 == 0 && !getenv("DUKE_HOST_DECODE")) {
            // files -> GPU (facade/ingest.h): inflate on all host threads, everything after it on the GPU, the cloud back
            // in PointCloudImage's own storage layout
            std::string err;
            if (!duke::ingest_scan(eng, scanFolder, imgPrefix, imgSuffix, nimg, W, H, &err)) {
                fprintf(stderr, "%s\n", err.c_str());
                slr_ingest_abort(eng);
                break;
            }
            const size_t cells = (size_t)scan_w * scan_h;
            float *sums = nullptr;
            uint8_t *counts = nullptr;
            if (haveColor && !(h_color = duke::pinned_scratch(3, cells))) break;
            if (!duke::cloud_storage_acquire(cells, &sums, &counts)) break;
            if (slr_run_ge_ingested(eng, nbits, blackThreshold, whiteThreshold, scan_w, haveColor ? 1 : 0, scan_w, scan_h, sums, counts,
                                    (uint8_t *)h_color, nullptr, nullptr, nullptr, &n) != SLR_OK) {
                duke::cloud_storage_release(sums, counts);
                break;
            }
            delete points3DProjView;
            points3DProjView = new PointCloudImage(scan_w, scan_h, sums, counts, duke::cloud_storage_release,
                                                   haveColor ? (const uint8_t *)h_color : nullptr);
            ok = true;
            break;
        }
        if (!(h_stack = duke::pinned_scratch(0, 2 * (size_t)nimg * P))) break;
        if (!(h_xyz = duke::pinned_scratch(1, P * 3 * sizeof(float)))) break;
        if (!(h_valid = duke::pinned_scratch(2, P))) break;
        if (haveColor && !(h_color = duke::pinned_scratch(3, P))) break;
        const bool loaded = duke::load_stacks(scanFolder, imgPrefix, imgSuffix, 2, nimg, W, H, (uint8_t *)h_stack);
        if (!loaded) break;
        if (slr_run_ge_host(eng, (const uint8_t *)h_stack, 1, nbits, blackThreshold, whiteThreshold, scan_w, haveColor ? 1 : 0,
                            (float *)h_xyz, (uint8_t *)h_valid, nullptr, (uint8_t *)h_color, &n) != SLR_OK)
            break;
        delete points3DProjView;
        points3DProjView = new PointCloudImage(scan_w, scan_h, haveColor);
        points3DProjView->addDense((const float *)h_xyz, (const uint8_t *)h_valid, (const uint8_t *)h_color, W, H);
        ok = true;
    } while (false);
    if (!ok && slr_last_error()[0]) fprintf(stderr, "Reconstruct: %s\n", slr_last_error());
    return ok;   // engine and pinned buffers stay with the process (reconstruct_common.h)
}

bool Reconstruct::runReconstruction()
{
    const int W = cameraWidth, H = cameraHeight;
    const size_t P = (size_t)W * H;
    const int nc = slr_gray_num_bits(scan_w), nr = slr_gray_num_bits(scan_h);   // GrayCodes(scan_w, scan_h, false)
    const int nimg = 2 + 2 * nc + 2 * nr;
    for (int i = 0; i < 2; i++) {                       // :239-240
        cameras[i].position = duke::Point3f(0, 0, 0);
        cam2WorldSpace(cameras[i], cameras[i].position);
        cameras[i].width = W;
        cameras[i].height = H;
    }
    slr_engine *eng = duke::shared_engine(device, W, H);
    if (!eng) {
        fprintf(stderr, "Reconstruct: %s\n", slr_last_error());
        return false;
    }
    bool ok = false;
    const size_t ncell = (size_t)scan_w * scan_h;
    std::vector<uint8_t> stack(2 * (size_t)nimg * P), cnt(ncell);
    std::vector<float> sum(ncell * 3);
    do {
        slr_camera cams[2] = {duke::to_slr_camera(cameras[0]), duke::to_slr_camera(cameras[1])};
        const double Qid[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};   // unused by the ray-ray path
        float rigid[12];
        const float *rg = nullptr;
        if (scanSN > 0 && duke::load_rigid(savePath_ + "/scan/transfer_mat" + std::to_string(scanSN) + ".txt", rigid)) rg = rigid;
        if (!duke::upload_calibration(eng, cams, Qid, rg, 0, nullptr, nullptr)) break;
        if (slr_set_host_input_raw(eng, 0) != SLR_OK) break;   // the shared engine may have rectified for a GE / MF scan
        if (slr_set_auto_contrast(eng, autoContrast_ ? 1 : 0) != SLR_OK) break;
        const bool loaded = duke::load_stacks(scanFolder, imgPrefix, imgSuffix, 2, nimg, W, H, stack.data());
        if (!loaded) break;
        unsigned long long n = 0;
        if (slr_run_gray_host(eng, stack.data(), 1, nc, nr, blackThreshold, whiteThreshold, scan_w, scan_h, sum.data(),
                              cnt.data(), &n) != SLR_OK)
            break;
        delete points3DProjView;
        // colour storage as in Duke/reconstruct.cpp:262; the Gray-only triangulation never sets a colour (:417-481),
        // so with haveColor the export shows "0 0 0" columns, exactly as the reference's
        points3DProjView = new PointCloudImage(scan_w, scan_h, haveColor);
        // cells arrive in the reference's ac(i, j) = i*scan_h + j order; the accumulated state (sum, wrapped count)
        // is reproduced by one setPoint plus count-1 zero additions
        for (int i = 0; i < scan_w; i++)
            for (int j = 0; j < scan_h; j++) {
                const size_t cell = (size_t)i * scan_h + j;
                if (!cnt[cell]) continue;
                points3DProjView->setPoint(i, j, duke::Point3f(sum[cell * 3], sum[cell * 3 + 1], sum[cell * 3 + 2]));
                for (int k = 1; k < cnt[cell]; k++) points3DProjView->addPoint(i, j, duke::Point3f(0, 0, 0));
            }
        ok = true;
    } while (false);
    if (!ok && slr_last_error()[0]) fprintf(stderr, "Reconstruct: %s\n", slr_last_error());
    return ok;
}
