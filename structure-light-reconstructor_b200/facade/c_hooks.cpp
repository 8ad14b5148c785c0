// c_hooks.cpp — extern "C" probes into the facade's host logic so the CPU test-suite can exercise it through ctypes
// (image IO, the stereoRectify / initUndistortRectifyMap restatement, stereoRect's text loaders) without a GPU.
#include <string.h>

#include "imageio.h"
#include "inflate.h"
#include "ingest.h"
#include "meshcreator.h"
#include "rectify.h"
#include "stereorect.h"
#include "virtualcamera.h"

extern "C" {

// in: M1[9] D1[5] M2[9] D2[5] R[9] T[3]; out: R1[9] R2[9] P1[12] P2[12] Q[16]
void duke_stereo_rectify(const double *in, int w, int h, int variant, double *out)
{
    duke::Matrix M1(3, 3), D1(5, 1), M2(3, 3), D2(5, 1), R(3, 3), T(3, 1);
    duke::Matrix *ms[6] = {&M1, &D1, &M2, &D2, &R, &T};
    for (auto *m : ms) {
        memcpy(m->v.data(), in, m->v.size() * sizeof(double));
        in += m->v.size();
    }
    duke::RectifyResult r = duke::stereo_rectify(M1, D1, M2, D2, duke::Size(w, h), R, T, (duke::RectifyVariant)variant);
    const duke::Matrix *os[5] = {&r.R1, &r.R2, &r.P1, &r.P2, &r.Q};
    for (auto *m : os) {
        memcpy(out, m->v.data(), m->v.size() * sizeof(double));
        out += m->v.size();
    }
}

// M[9] D[5] R[9] P[12] -> map1 int16 [h][w][2], map2 uint16 [h][w]
void duke_init_undistort_rectify_map(const double *M, const double *D, const double *R, const double *P, int w, int h,
                                     int16_t *map1, uint16_t *map2)
{
    duke::Matrix m(3, 3), d(5, 1), r(3, 3), p(3, 4);
    memcpy(m.v.data(), M, 9 * sizeof(double));
    memcpy(d.v.data(), D, 5 * sizeof(double));
    memcpy(r.v.data(), R, 9 * sizeof(double));
    memcpy(p.v.data(), P, 12 * sizeof(double));
    std::vector<int16_t> a;
    std::vector<uint16_t> b;
    duke::init_undistort_rectify_map(m, d, r, p, duke::Size(w, h), a, b);
    memcpy(map1, a.data(), a.size() * sizeof(int16_t));
    memcpy(map2, b.data(), b.size() * sizeof(uint16_t));
}

int duke_read_gray_image(const char *path, int *w, int *h, uint8_t *pix, int cap)
{
    duke::Image img;
    if (!duke::read_gray_image(path, img)) return -1;
    *w = img.width;
    *h = img.height;
    if ((int)img.pix.size() > cap) return -2;
    memcpy(pix, img.pix.data(), img.pix.size());
    return 0;
}

// the host half of the PNG ingest path for one image: returns the route (ingest.h) and fills out[H * (1 + W)]
int duke_decode_scan_image(const char *base, const char *suffix, int w, int h, uint8_t *out)
{
    return duke::decode_scan_image(base, suffix, w, h, out, nullptr);
}

// zlib stream -> bytes with the ingest path's own decoder (inflate.cpp); 0 = ok, -1 = rejected
int duke_zlib_inflate(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len)
{
    return duke::zlib_inflate(in, in_len, out, out_len) ? 0 : -1;
}

int duke_write_png_gray(const char *path, const uint8_t *pix, int w, int h) { return duke::write_png_gray(path, pix, w, h) ? 0 : -1; }
int duke_write_pgm(const char *path, const uint8_t *pix, int w, int h) { return duke::write_pgm(path, pix, w, h) ? 0 : -1; }

// stereoRect on a project directory: Q[16] + doStereoRectify of one image (host fixed-point path)
int duke_stereorect_probe(const char *project, int w, int h, double *Q, uint8_t *img_inout, int isleft)
{
    stereoRect sr(project, duke::Size(w, h));
    sr.getParameters();
    if (!sr.loaded()) return -1;
    sr.calParameters();
    memcpy(Q, sr.Q.v.data(), 16 * sizeof(double));
    duke::Image im;
    im.width = w;
    im.height = h;
    im.pix.assign(img_inout, img_inout + (size_t)w * h);
    sr.doStereoRectify(im, isleft != 0);
    memcpy(img_inout, im.pix.data(), (size_t)w * h);
    return 0;
}

// VirtualCamera::loadCameraMatrix -> fc, cc ; returns 0 when the file is missing (the reference's `false`)
int duke_load_camera_matrix(const char *path, float *fc_cc)
{
    VirtualCamera vc;
    if (!vc.loadCameraMatrix(path)) return 0;
    fc_cc[0] = vc.fc.x;
    fc_cc[1] = vc.fc.y;
    fc_cc[2] = vc.cc.x;
    fc_cc[3] = vc.cc.y;
    return 1;
}


// MeshCreator on a cloud given as sums [h][w][3] + counts [h][w] (+ colour values int [h][w][3] or NULL):
// fills a PointCloudImage the way the reconstructors do and exports <path> (needs a GPU).
int duke_export_mesh(const float *sums, const uint8_t *counts, const int *color, int w, int h, int obj, const char *path,
                     unsigned long long *nv, unsigned long long *nf)
{
    PointCloudImage pc(w, h, color != nullptr);
    for (int j = 0; j < h; j++)
        for (int i = 0; i < w; i++) {
            const size_t q = (size_t)j * w + i;
            for (int k = 0; k < counts[q]; k++) {   // first call sets, later calls add (pointcloudimage.cpp:86-97)
                const duke::Point3f p(k == 0 ? sums[q * 3] : -0.f, k == 0 ? sums[q * 3 + 1] : -0.f, k == 0 ? sums[q * 3 + 2] : -0.f);  // x + (-0) == x, signed zeros included
                if (color)
                    pc.addPoint(i, j, p, k == 0 ? duke::Vec3i(color[q * 3], color[q * 3 + 1], color[q * 3 + 2]) : duke::Vec3i(0, 0, 0));
                else
                    pc.addPoint(i, j, p);
            }
        }
    MeshCreator mc(&pc);
    if (obj)
        mc.exportObjMesh(path);
    else
        mc.exportPlyMesh(path);
    if (nv) *nv = mc.vertexCount();
    if (nf) *nf = mc.faceCount();
    return mc.ok() ? 0 : -1;
}

// The text stage of MeshCreator alone (no GPU): index arrays supplied by the caller (the CPU tests pass the oracle's)
int duke_write_mesh_text(const float *sums, const uint8_t *counts, const int *color, int w, int h, int obj, const float *vert,
                         const int32_t *src, const int32_t *faces, unsigned long long nv, unsigned long long nf, const char *path)
{
    PointCloudImage pc(w, h, color != nullptr);
    for (int j = 0; j < h; j++)
        for (int i = 0; i < w; i++) {
            const size_t q = (size_t)j * w + i;
            for (int k = 0; k < counts[q]; k++) {
                const duke::Point3f p(k == 0 ? sums[q * 3] : -0.f, k == 0 ? sums[q * 3 + 1] : -0.f, k == 0 ? sums[q * 3 + 2] : -0.f);
                if (color)
                    pc.addPoint(i, j, p, k == 0 ? duke::Vec3i(color[q * 3], color[q * 3 + 1], color[q * 3 + 2]) : duke::Vec3i(0, 0, 0));
                else
                    pc.addPoint(i, j, p);
            }
        }
    return duke::write_mesh_text(path, obj != 0, &pc, vert, src, faces, (size_t)nv, (size_t)nf) ? 0 : -1;
}

// PointCloudImage::exportXYZ of a cloud given as sums + counts (+ colour values), host only
int duke_export_xyz(const float *sums, const uint8_t *counts, const int *color, int w, int h, int export_off, int color_flag,
                    const char *path)
{
    PointCloudImage pc(w, h, color != nullptr);
    for (int j = 0; j < h; j++)
        for (int i = 0; i < w; i++) {
            const size_t q = (size_t)j * w + i;
            for (int k = 0; k < counts[q]; k++) {
                const duke::Point3f p(k == 0 ? sums[q * 3] : -0.f, k == 0 ? sums[q * 3 + 1] : -0.f, k == 0 ? sums[q * 3 + 2] : -0.f);
                if (color)
                    pc.addPoint(i, j, p, k == 0 ? duke::Vec3i(color[q * 3], color[q * 3 + 1], color[q * 3 + 2]) : duke::Vec3i(0, 0, 0));
                else
                    pc.addPoint(i, j, p);
            }
        }
    pc.exportXYZ(path, export_off != 0, color_flag != 0);
    return 0;
}
}  // extern "C"
