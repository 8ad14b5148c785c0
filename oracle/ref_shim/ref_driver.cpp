// ref_driver.cpp -- extern "C" entry points that drive the REFERENCE's own code (Duke/*.cpp, compiled
// unmodified from /root/reference against the shim headers in this directory) on in-memory inputs, so that
// tests can compare the oracle restatement with what the reference's sources actually compute.
// Test infrastructure only.  Built by `make -C oracle ref` into oracle/_ref/libref.so with
// -fno-access-control (the hot-path methods are private in the reference) and -fpermissive (MSVC-isms).
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <string>

#include "graycodes.h"
#include "mfreconstruct.h"
#include "multifrequency.h"
#include "pointcloudimage.h"
#include "meshcreator.h"
#include "reconstruct.h"
#include "utilities.h"

struct ref_camera {  // same layout as orc_camera / slr_camera
    float fc[2], cc[2], dist[5], R[9], t[3];
};

static cv::Mat mat_from_u8(const uint8_t *src, int W, int H)
{
    cv::Mat m(H, W, CV_8U);
    for (int r = 0; r < H; r++) memcpy(m.data + (size_t)r * m.step, src + (size_t)r * W, (size_t)W);
    return m;
}

static void fill_camera(VirtualCamera &vc, const ref_camera &c, int W, int H)
{
    vc.fc.x = c.fc[0];
    vc.fc.y = c.fc[1];
    vc.cc.x = c.cc[0];
    vc.cc.y = c.cc[1];
    vc.distortion = cv::Mat(5, 1, CV_32F);
    for (int i = 0; i < 5; i++) vc.distortion.at<float>(i) = c.dist[i];
    vc.rotationMatrix = cv::Mat(3, 3, CV_32F);
    for (int i = 0; i < 9; i++) vc.rotationMatrix.at<float>(i / 3, i % 3) = c.R[i];
    vc.translationVector = cv::Mat(3, 1, CV_32F);
    for (int i = 0; i < 3; i++) vc.translationVector.at<float>(i) = c.t[i];
    vc.width = W;
    vc.height = H;
}

static std::string write_rigid(const float *rigid, int sn)
{
    // Reconstruct / MFReconstruct::triangulation load <savePath>/scan/transfer_mat<sn>.txt (mfreconstruct.cpp:278-282)
    char dir[] = "/tmp/slr_refXXXXXX";
    if (!mkdtemp(dir)) return std::string();
    std::string base(dir);
    std::string cmd = "mkdir -p " + base + "/scan";
    if (system(cmd.c_str()) != 0) return std::string();
    std::string path = base + "/scan/transfer_mat" + std::to_string(sn) + ".txt";
    FILE *f = fopen(path.c_str(), "w");
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 4; j++) fprintf(f, "%.9g\t", rigid[4 * i + j]);
        fprintf(f, "\n");
    }
    fclose(f);
    return base;
}

static void dump_cloud(PointCloudImage *pc, int scan_w, int scan_h, float *points, uint8_t *count)
{
    // PointCloudImage::points is h x w CV_32FC3, numOfPointsForPixel h x w CV_8U (pointcloudimage.cpp:3-13)
    for (int r = 0; r < scan_h; r++)
        for (int c = 0; c < scan_w; c++) {
            const uint8_t n = pc->numOfPointsForPixel.at<uchar>(r, c);
            count[(size_t)r * scan_w + c] = n;
            cv::Vec3f v = pc->points.at<cv::Vec3f>(r, c);
            for (int k = 0; k < 3; k++) points[((size_t)r * scan_w + c) * 3 + k] = n ? v[k] : 0.0f;
        }
}

extern "C" {

// ---- patterns ----------------------------------------------------------------------------------------
int ref_gray_layout(int W, int H, int epi, int *ncol, int *nrow)
{
    GrayCodes g(W, H, epi != 0);
    *ncol = g.getNumOfColBits();
    *nrow = g.getNumOfRowBits();
    return g.getNumOfImgs();
}

void ref_generate_gray(int W, int H, int epi, uint8_t *out)
{
    GrayCodes g(W, H, epi != 0);
    g.generateGrays();
    for (int n = 0; n < g.getNumOfImgs(); n++)
        for (int r = 0; r < H; r++) memcpy(out + ((size_t)n * H + r) * W, g.grayCodes[n].data + (size_t)r * g.grayCodes[n].step, (size_t)W);
}

int ref_gray_to_dec(const uint8_t *bits, int n)
{
    cv::vector<bool> v;
    for (int i = 0; i < n; i++) v.push_back(bits[i] != 0);
    return GrayCodes::grayToDec(v);
}

void ref_generate_mf(int W, int H, uint8_t *out)
{
    MultiFrequency mf(0, W, H);
    mf.generateMutiFreq();
    for (int n = 0; n < mf.getNumOfImgs(); n++)
        for (int r = 0; r < H; r++)
            memcpy(out + ((size_t)n * H + r) * W, mf.MultiFreqImages[n].data + (size_t)r * mf.MultiFreqImages[n].step, (size_t)W);
}

// ---- MF decode: computeShadows + decodePatterns (+ getPhase) ----------------------------------------------
// phase/has: what decodePatterns pushed (has = the pixel's vector is non-empty); mask: the final mask Mat.
void ref_mf_decode(const uint8_t *stack, int W, int H, int black_thr, float *phase, uint8_t *has, uint8_t *mask)
{
    MFReconstruct m;
    m.cameraWidth = W;
    m.cameraHeight = H;
    m.blackThreshold = black_thr;
    m.savePath_ = "/tmp";
    for (int n = 0; n < 14; n++) m.camImgs.push_back(mat_from_u8(stack + (size_t)n * W * H, W, H));
    m.camPixels = new cv::vector<float>[(size_t)W * H];
    m.computeShadows();
    m.decodePatterns();
    for (size_t p = 0; p < (size_t)W * H; p++) {
        has[p] = m.camPixels[p].size() ? 1 : 0;
        phase[p] = has[p] ? m.camPixels[p][0] : 0.0f;
        mask[p] = m.mask.at<uchar>((int)(p / W), (int)(p % W));
    }
    delete[] m.camPixels;
}

// ---- MF triangulation ------------------------------------------------------------------------------------
void ref_mf_triangulate(const float *phL, const uint8_t *hasL, const float *phR, const uint8_t *hasR, int W, int H,
                        const ref_camera *camL, const ref_camera *camR, const double *Q, const float *rigid,
                        int scan_w, int scan_h, float *points, uint8_t *count)
{
    MFReconstruct m;
    m.cameraWidth = W;
    m.cameraHeight = H;
    m.scan_w = scan_w;
    m.scan_h = scan_h;
    m.scanSN = 0;
    std::string base;
    if (rigid) {
        m.scanSN = 1;
        base = write_rigid(rigid, 1);
        m.savePath_ = QString(base);
    }
    fill_camera(m.cameras[0], *camL, W, H);
    fill_camera(m.cameras[1], *camR, W, H);
    m.sr = new stereoRect(QString("/tmp"), cv::Size(W, H));
    m.sr->Q = cv::Mat(4, 4, CV_64F);
    for (int i = 0; i < 16; i++) m.sr->Q.at<double>(i / 4, i % 4) = Q[i];
    cv::vector<float> *L = new cv::vector<float>[(size_t)W * H], *R = new cv::vector<float>[(size_t)W * H];
    for (size_t p = 0; p < (size_t)W * H; p++) {
        if (hasL[p]) L[p].push_back(phL[p]);
        if (hasR[p]) R[p].push_back(phR[p]);
    }
    m.points3DProjView = new PointCloudImage(scan_w, scan_h, false);
    m.triangulation(L, m.cameras[0], R, m.cameras[1]);
    dump_cloud(m.points3DProjView, scan_w, scan_h, points, count);
    delete[] L;
    delete[] R;
    delete m.points3DProjView;
    if (!base.empty()) { std::string cmd = "rm -rf " + base; if (system(cmd.c_str())) {} }
}

// ---- Gray decode: computeShadows + decodePatterns_GE / decodePaterns ------------------------------------------
// EPI (nbits_row == 0): col[p] = pushed xDec or -1.  Gray-only: col/row = the projector cell of camera pixel p.
void ref_gray_decode(const uint8_t *stack, int W, int H, int nbits_col, int nbits_row, int black_thr, int white_thr,
                     int scan_w, int scan_h, int32_t *col, int32_t *row, uint8_t *mask)
{
    const bool epi = (nbits_row == 0);
    Reconstruct r(false);  // EPI flag only steers loadCamImgs / the destructor's `delete sr`
    r.cameraWidth = W;
    r.cameraHeight = H;
    r.scan_w = scan_w;
    r.scan_h = scan_h;
    r.numOfColBits = nbits_col;
    r.numOfRowBits = nbits_row;
    r.numberOfImgs = 2 + 2 * nbits_col + 2 * nbits_row;
    r.blackThreshold = black_thr;
    r.whiteThreshold = white_thr;
    r.camera = &r.cameras[0];
    r.camera->width = W;
    r.camera->height = H;
    for (int n = 0; n < r.numberOfImgs; n++) r.camImgs.push_back(mat_from_u8(stack + (size_t)n * W * H, W, H));
    r.computeShadows();
    for (size_t p = 0; p < (size_t)W * H; p++) col[p] = row[p] = -1;
    if (epi) {
        r.camPixels_GE = new cv::vector<int>[(size_t)W * H];
        r.decodePatterns_GE();
        for (size_t p = 0; p < (size_t)W * H; p++)
            if (r.camPixels_GE[p].size()) col[p] = r.camPixels_GE[p][0];
        delete[] r.camPixels_GE;
    } else {
        // one spare column of cells: xDec == scan_w is accepted by getProjPixel (:365 uses >) and indexes past the
        // reference's scan_h*scan_w array (out-of-bounds write there); give it room so the run is defined
        const size_t ncell = (size_t)scan_h * (scan_w + 2);
        r.camPixels = new cv::vector<cv::Point>[ncell];
        r.decodePaterns();
        for (size_t cell = 0; cell < ncell; cell++)
            for (size_t k = 0; k < r.camPixels[cell].size(); k++) {
                const cv::Point px = r.camPixels[cell][k];
                const size_t p = (size_t)px.y * W + px.x;
                col[p] = (int32_t)(cell / scan_h);
                row[p] = (int32_t)(cell % scan_h);
            }
        delete[] r.camPixels;
    }
    for (size_t p = 0; p < (size_t)W * H; p++) mask[p] = r.mask.at<uchar>((int)(p / W), (int)(p % W));
}

// ---- GE triangulation ------------------------------------------------------------------------------------
void ref_ge_triangulate(const int32_t *colL, const uint8_t *hasL, const int32_t *colR, const uint8_t *hasR, int W, int H,
                        const double *Q, const float *rigid, const uint8_t *whiteL, const uint8_t *whiteR,
                        int scan_w, int scan_h, float *points, uint8_t *count, uint8_t *color)
{
    Reconstruct r(false);
    r.cameraWidth = W;
    r.cameraHeight = H;
    r.scan_w = scan_w;
    r.scan_h = scan_h;
    r.scanSN = 0;
    r.haveColor = (whiteL != 0);
    std::string base;
    if (rigid) {
        r.scanSN = 1;
        base = write_rigid(rigid, 1);
        r.savePath_ = QString(base);
    }
    r.sr = new stereoRect(QString("/tmp"), cv::Size(W, H));
    r.sr->Q = cv::Mat(4, 4, CV_64F);
    for (int i = 0; i < 16; i++) r.sr->Q.at<double>(i / 4, i % 4) = Q[i];
    if (whiteL) {
        r.colorImgs.push_back(mat_from_u8(whiteL, W, H));
        r.colorImgs.push_back(mat_from_u8(whiteR, W, H));
    }
    cv::vector<int> *L = new cv::vector<int>[(size_t)W * H], *R = new cv::vector<int>[(size_t)W * H];
    for (size_t p = 0; p < (size_t)W * H; p++) {
        if (hasL[p]) L[p].push_back(colL[p]);
        if (hasR[p]) R[p].push_back(colR[p]);
    }
    r.points3DProjView = new PointCloudImage(scan_w, scan_h, r.haveColor);
    r.triangulation_ge(L, r.cameras[0], R, r.cameras[1]);
    dump_cloud(r.points3DProjView, scan_w, scan_h, points, count);
    if (color && whiteL)
        for (int y = 0; y < scan_h; y++)
            for (int x = 0; x < scan_w; x++) color[(size_t)y * scan_w + x] = r.points3DProjView->color.at<cv::Vec3b>(y, x)[0];
    delete[] L;
    delete[] R;
    delete r.sr;
    r.sr = 0;
    if (!base.empty()) { std::string cmd = "rm -rf " + base; if (system(cmd.c_str())) {} }
}

// ---- Gray-only bucket triangulation ---------------------------------------------------------------------------
// col/row/has per camera pixel (as ref_gray_decode returns them); buckets are rebuilt in the reference's push
// order (camera column-major, reconstruct.cpp:60-61) and handed to Reconstruct::triangulation.
void ref_gray_triangulate(const int32_t *colL, const int32_t *rowL, const uint8_t *hasL, const int32_t *colR,
                          const int32_t *rowR, const uint8_t *hasR, int W, int H, int scan_w, int scan_h,
                          const ref_camera *camL, const ref_camera *camR, const float *rigid, float *sum, uint8_t *cnt)
{
    Reconstruct r(false);
    r.cameraWidth = W;
    r.cameraHeight = H;
    r.scan_w = scan_w;
    r.scan_h = scan_h;
    r.scanSN = 0;
    r.haveColor = false;
    std::string base;
    if (rigid) {
        r.scanSN = 1;
        base = write_rigid(rigid, 1);
        r.savePath_ = QString(base);
    }
    fill_camera(r.cameras[0], *camL, W, H);
    fill_camera(r.cameras[1], *camR, W, H);
    for (int i = 0; i < 2; i++) {  // runReconstruction :239-240
        r.cameras[i].position = cv::Point3f(0, 0, 0);
        r.cam2WorldSpace(r.cameras[i], r.cameras[i].position);
    }
    const size_t ncell = (size_t)scan_w * scan_h;
    cv::vector<cv::Point> *B[2] = {new cv::vector<cv::Point>[ncell], new cv::vector<cv::Point>[ncell]};
    const int32_t *cols[2] = {colL, colR}, *rows[2] = {rowL, rowR};
    const uint8_t *has[2] = {hasL, hasR};
    for (int cam = 0; cam < 2; cam++)
        for (int c = 0; c < W; c++)
            for (int y = 0; y < H; y++) {
                const size_t p = (size_t)y * W + c;
                if (!has[cam][p]) continue;
                const size_t cell = (size_t)cols[cam][p] * scan_h + rows[cam][p];
                if (cell < ncell) B[cam][cell].push_back(cv::Point(c, y));
            }
    r.points3DProjView = new PointCloudImage(scan_w, scan_h, false);
    r.triangulation(B[0], r.cameras[0], B[1], r.cameras[1]);
    // addPoint(i, j) with i in [0,w), j in [0,h): stored at Mat(row = j, col = i); report in ac(i,j) = i*scan_h + j order
    for (int i = 0; i < scan_w; i++)
        for (int j = 0; j < scan_h; j++) {
            const size_t cell = (size_t)i * scan_h + j;
            const uint8_t n = r.points3DProjView->numOfPointsForPixel.at<uchar>(j, i);
            cnt[cell] = n;
            cv::Vec3f v = r.points3DProjView->points.at<cv::Vec3f>(j, i);
            for (int k = 0; k < 3; k++) sum[cell * 3 + k] = n ? v[k] : 0.0f;
        }
    delete[] B[0];
    delete[] B[1];
    if (!base.empty()) { std::string cmd = "rm -rf " + base; if (system(cmd.c_str())) {} }
}

// ---- helpers ---------------------------------------------------------------------------------------------
void ref_undistort(float x, float y, const ref_camera *cam, float *ox, float *oy)
{
    VirtualCamera vc;
    fill_camera(vc, *cam, 0, 0);
    cv::Point2f p = Utilities::undistortPoints(cv::Point2f(x, y), vc);
    *ox = p.x;
    *oy = p.y;
}

int ref_line_line(const float *p1, const float *v1, const float *p2, const float *v2, float *out)
{
    cv::Point3f p;
    bool ok = Utilities::line_lineIntersection(cv::Point3f(p1[0], p1[1], p1[2]), cv::Vec3f(v1[0], v1[1], v1[2]),
                                               cv::Point3f(p2[0], p2[1], p2[2]), cv::Vec3f(v2[0], v2[1], v2[2]), p);
    out[0] = p.x;
    out[1] = p.y;
    out[2] = p.z;
    return ok ? 1 : 0;
}

void ref_cam2world(const ref_camera *cam, float *p)
{
    Reconstruct r(false);
    VirtualCamera vc;
    fill_camera(vc, *cam, 0, 0);
    cv::Point3f q(p[0], p[1], p[2]);
    r.cam2WorldSpace(vc, q);
    p[0] = q.x;
    p[1] = q.y;
    p[2] = q.z;
}

void ref_normalize(float *v)
{
    cv::Vec3f x(v[0], v[1], v[2]);
    Utilities::normalize(x);
    v[0] = x[0];
    v[1] = x[1];
    v[2] = x[2];
}

// PointCloudImage::addPoint sequence -> sums and u8 counts (pointcloudimage.cpp:86-97)
void ref_pointcloud_add(int w, int h, const int32_t *iw, const int32_t *jh, const float *pts, int n, float *points,
                        uint8_t *count)
{
    PointCloudImage pc(w, h, false);
    for (int k = 0; k < n; k++) pc.addPoint(iw[k], jh[k], cv::Point3f(pts[3 * k], pts[3 * k + 1], pts[3 * k + 2]));
    dump_cloud(&pc, w, h, points, count);
}

// MeshCreator::exportPlyMesh / exportObjMesh (meshcreator.cpp:16-166) on a PointCloudImage filled from
// sums [h][w][3] + counts [h][w] (+ optional colour sums u8 [h][w][3]).
int ref_export_mesh(const float *points, const uint8_t *count, const uint8_t *color, int w, int h, int obj, const char *path)
{
    PointCloudImage pc(w, h, color != nullptr);
    for (int r = 0; r < h; r++)
        for (int c = 0; c < w; c++) {
            const size_t q = (size_t)r * w + c;
            pc.numOfPointsForPixel.at<uchar>(r, c) = count[q];
            pc.points.at<cv::Vec3f>(r, c) = cv::Vec3f(points[q * 3], points[q * 3 + 1], points[q * 3 + 2]);
            if (color) pc.color.at<cv::Vec3b>(r, c) = cv::Vec3b(color[q * 3], color[q * 3 + 1], color[q * 3 + 2]);
        }
    MeshCreator mc(&pc);
    if (obj)
        mc.exportObjMesh(QString(path));
    else
        mc.exportPlyMesh(QString(path));
    return 0;
}

// PointCloudImage::exportXYZ (pointcloudimage.cpp:99-122) on a cloud given as sums + counts (+ colour u8)
int ref_export_xyz(const float *points, const uint8_t *count, const uint8_t *color, int w, int h, int export_off, int color_flag,
                   const char *path)
{
    PointCloudImage pc(w, h, color != nullptr);
    for (int r = 0; r < h; r++)
        for (int c = 0; c < w; c++) {
            const size_t q = (size_t)r * w + c;
            pc.numOfPointsForPixel.at<uchar>(r, c) = count[q];
            pc.points.at<cv::Vec3f>(r, c) = cv::Vec3f(points[q * 3], points[q * 3 + 1], points[q * 3 + 2]);
            if (color) pc.color.at<cv::Vec3b>(r, c) = cv::Vec3b(color[q * 3], color[q * 3 + 1], color[q * 3 + 2]);
        }
    pc.exportXYZ((char *)path, export_off != 0, color_flag != 0);
    return 0;
}

}  // extern "C"
