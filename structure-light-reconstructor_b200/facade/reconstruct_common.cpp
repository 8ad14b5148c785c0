#include "reconstruct_common.h"

#include <stdio.h>
#include <string.h>

#include <atomic>
#include <fstream>
#include <mutex>
#include <thread>
#include <vector>

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
    const std::string folders[2] = {folder, folder}, prefixes[2] = {prefix, prefix};
    return load_stacks(folders, prefixes, suffix, 1, n, W, H, dst);
}

bool load_stacks(const std::string folders[2], const std::string prefixes[2], const std::string &suffix, int cams, int n_per_cam,
                 int W, int H, uint8_t *dst)
{
    const int n = cams * n_per_cam;
    // The images of a stack are independent files: decode them on all host threads (PNG inflate of a 1280x1024
    // frame costs ~10 ms, the GPU pipeline for the whole scan ~0.05 ms).  Errors are reported for the lowest failing
    // index, as the reference's sequential loop would (mfreconstruct.cpp:119-139).
    std::vector<std::string> errors((size_t)n);
    std::vector<char> failed((size_t)n, 0);
    std::atomic<int> next(0);
    auto worker = [&]() {
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= n) return;
            const std::string base = folders[i / n_per_cam] + prefixes[i / n_per_cam] + std::to_string(i % n_per_cam);
            Image img;
            std::string err;
            char msg[1024];
            bool got = false;
            try {
                got = read_gray_image(base + suffix, img, &err) || read_gray_image(base + ".pgm", img, &err);
            } catch (const std::exception &ex) {   // e.g. bad_alloc on a hostile header: report, do not terminate the thread
                err = ex.what();
            }
            if (!got) {
                snprintf(msg, sizeof(msg), "Load Images: Scan Images not found! (%s: %s)", (base + suffix).c_str(), err.c_str());
                errors[i] = msg;
                failed[i] = 1;
                continue;
            }
            if (img.width != W || img.height != H) {
                snprintf(msg, sizeof(msg), "Load Images: %s is %dx%d, expected %dx%d", (base + suffix).c_str(), img.width,
                         img.height, W, H);
                errors[i] = msg;
                failed[i] = 1;
                continue;
            }
            memcpy(dst + (size_t)i * W * H, img.pix.data(), (size_t)W * H);
        }
    };
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (nt > (unsigned)n) nt = (unsigned)n;
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nt; t++) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
    for (int i = 0; i < n; i++)
        if (failed[i]) {
            fprintf(stderr, "%s\n", errors[i].c_str());
            return false;
        }
    return true;
}

namespace {
std::mutex g_mu;
struct EngineSlot {
    int device, W, H;
    slr_engine *e;
    bool calib_valid = false, has_rigid = false;
    slr_camera cams[2];
    double Q[16];
    float rigid[12];
    unsigned long long maps_id = 0;   // stereoRect::calibrationId() of the maps on the GPU (0: none)
};
std::vector<EngineSlot> g_engines;      // never destroyed: the driver tears the context down at process exit
void *g_pinned[4] = {nullptr, nullptr, nullptr, nullptr};
size_t g_pinned_bytes[4] = {0, 0, 0, 0};
}  // namespace

slr_engine *shared_engine(int device, int W, int H)
{
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto &s : g_engines)
        if (s.device == device && s.W == W && s.H == H) return s.e;
    slr_engine *e = nullptr;
    if (slr_create(&e, device, W, H, 1) != SLR_OK) return nullptr;
    EngineSlot slot;
    slot.device = device;
    slot.W = W;
    slot.H = H;
    slot.e = e;
    g_engines.push_back(slot);
    return e;
}

bool upload_calibration(slr_engine *eng, const slr_camera cams[2], const double Q[16], const float *rigid3x4,
                        unsigned long long cal_id, const int16_t *h_map1, const uint16_t *h_map2)
{
    std::lock_guard<std::mutex> lk(g_mu);
    EngineSlot *s = nullptr;
    for (auto &g : g_engines)
        if (g.e == eng) s = &g;
    const bool same = s && s->calib_valid && memcmp(s->cams, cams, sizeof(s->cams)) == 0 && memcmp(s->Q, Q, sizeof(s->Q)) == 0 &&
                      s->has_rigid == (rigid3x4 != nullptr) && (!rigid3x4 || memcmp(s->rigid, rigid3x4, sizeof(s->rigid)) == 0);
    if (!same) {
        if (s) s->calib_valid = false;
        if (slr_set_calib(eng, cams, Q, rigid3x4) != SLR_OK) return false;
        if (s) {
            memcpy(s->cams, cams, sizeof(s->cams));
            memcpy(s->Q, Q, sizeof(s->Q));
            s->has_rigid = rigid3x4 != nullptr;
            if (rigid3x4) memcpy(s->rigid, rigid3x4, sizeof(s->rigid));
            s->calib_valid = true;
        }
    }
    if (h_map1 && h_map2 && !(s && cal_id != 0 && s->maps_id == cal_id)) {
        if (s) s->maps_id = 0;
        if (slr_set_rectify_maps(eng, h_map1, h_map2) != SLR_OK) return false;
        if (s) s->maps_id = cal_id;
    }
    return true;
}

void *pinned_scratch(int slot, size_t bytes)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (slot < 0 || slot >= 4) return nullptr;
    if (g_pinned_bytes[slot] >= bytes && g_pinned[slot]) return g_pinned[slot];
    if (g_pinned[slot]) slr_host_free(g_pinned[slot]);
    g_pinned[slot] = nullptr;
    g_pinned_bytes[slot] = 0;
    if (slr_host_alloc(&g_pinned[slot], bytes) != SLR_OK) return nullptr;
    g_pinned_bytes[slot] = bytes;
    return g_pinned[slot];
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
