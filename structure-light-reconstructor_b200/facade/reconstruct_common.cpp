#include "reconstruct_common.h"

#include <stdio.h>
#include <string.h>

#include <fstream>

#include "imageio.h"

namespace duke {

slr_camera to_slr_camera(const VirtualCamera &vc)
{
    slr_camera c;
    memset(&c, 0, sizeof(c));
    c.fc[0] = vc.fc.x;
    c.fc[1] = vc.fc.y;
    c.cc[0] = vc.cc.x;
    c.cc[1] = vc.cc.y;
    for (int i = 0; i < 5 && i < (int)vc.distortion.v.size(); i++) c.dist[i] = (float)vc.distortion.v[i];
    for (int i = 0; i < 9; i++) c.R[i] = (i < (int)vc.rotationMatrix.v.size()) ? (float)vc.rotationMatrix.v[i] : (i % 4 == 0 ? 1.f : 0.f);
    for (int i = 0; i < 3 && i < (int)vc.translationVector.v.size(); i++) c.t[i] = (float)vc.translationVector.v[i];
    return c;
}

bool load_stack(const std::string &folder, const std::string &prefix, const std::string &suffix, int n, int W, int H,
                uint8_t *dst)
{
    for (int i = 0; i < n; i++) {
        const std::string base = folder + prefix + std::to_string(i);
        Image img;
        std::string err;
        if (!read_gray_image(base + suffix, img, &err) && !read_gray_image(base + ".pgm", img, &err)) {
            fprintf(stderr, "Load Images: Scan Images not found! (%s: %s)\n", (base + suffix).c_str(), err.c_str());
            return false;
        }
        if (img.width != W || img.height != H) {
            fprintf(stderr, "Load Images: %s is %dx%d, expected %dx%d\n", (base + suffix).c_str(), img.width, img.height, W, H);
            return false;
        }
        memcpy(dst + (size_t)i * W * H, img.pix.data(), (size_t)W * H);
    }
    return true;
}

bool load_rigid(const std::string &path, float out[12])
{
    std::ifstream in1(path.c_str());
    if (!in1) return false;
    for (int i = 0; i < 12; i++) {
        float v = 0;
        in1 >> v;
        out[i] = v;
    }
    return true;
}

}  // namespace duke
