#include "ingest.h"

#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <thread>
#include <vector>

#include "imageio.h"
#include "inflate.h"
#include "reconstruct_common.h"

namespace duke {
namespace {

uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

bool read_file(const std::string &path, std::vector<uint8_t> &buf)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize(n > 0 ? (size_t)n : 0);
    const size_t got = n > 0 ? fread(buf.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    return got == buf.size();
}

int paeth(int a, int b, int c)
{
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// PNG scanlines [H][1 + W] -> pixels [H][W], in place (row y is written below where it is read; the row above is final)
void unfilter_in_place(uint8_t *buf, int W, int H)
{
    for (int y = 0; y < H; y++) {
        const uint8_t *src = buf + (size_t)y * (W + 1);
        const int ft = src[0];
        src += 1;
        uint8_t *dst = buf + (size_t)y * W;
        const uint8_t *up = y ? dst - W : nullptr;
        for (int x = 0; x < W; x++) {
            const int a = x ? dst[x - 1] : 0, b = up ? up[x] : 0, c = (up && x) ? up[x - 1] : 0;
            int v = src[x];
            switch (ft) {
            case 1: v += a; break;
            case 2: v += b; break;
            case 3: v += (a + b) >> 1; break;
            case 4: v += paeth(a, b, c); break;
            default: break;
            }
            dst[x] = (uint8_t)v;
        }
    }
}

enum Route { ROUTE_FAILED, ROUTE_FILTERED, ROUTE_FILTERED_UP, ROUTE_PIXELS };

// One image file into its pinned slot (room for H * (1 + W) bytes).
Route decode_into(const std::string &base, const std::string &suffix, int W, int H, uint8_t *slot, std::string *err)
{
    std::vector<uint8_t> buf;
    std::string path = base + suffix;
    if (!read_file(path, buf)) {
        path = base + ".pgm";
        if (!read_file(path, buf)) {
            *err = "cannot read " + base + suffix;
            return ROUTE_FAILED;
        }
    }
    static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    if (buf.size() >= 8 + 25 && memcmp(buf.data(), sig, 8) == 0 && be32(&buf[8]) >= 13 && memcmp(&buf[12], "IHDR", 4) == 0 &&
        buf[24] == 8 && buf[25] == 0 && buf[28] == 0) {
        // 8-bit grey, not interlaced: the zlib stream holds H x (1 + W) bytes.  Collect the IDAT payloads in place.
        if ((int)be32(&buf[16]) != W || (int)be32(&buf[20]) != H) {
            char msg[256];
            snprintf(msg, sizeof(msg), "%s is %ux%u, expected %dx%d", path.c_str(), be32(&buf[16]), be32(&buf[20]), W, H);
            *err = msg;
            return ROUTE_FAILED;
        }
        size_t pos = 8, zlen = 0;
        uint8_t *z = buf.data();
        while (pos + 12 <= buf.size()) {
            const size_t len = be32(&buf[pos]);
            if (len > buf.size() - pos - 12) break;
            if (!memcmp(&buf[pos + 4], "IDAT", 4)) {
                memmove(z + zlen, &buf[pos + 8], len);      // moves down: payloads only ever land on bytes already read
                zlen += len;
            } else if (!memcmp(&buf[pos + 4], "IEND", 4)) {
                break;
            }
            pos += 12 + len;
        }
        std::string zerr;
        if (!zlib_inflate(z, zlen, slot, (size_t)(W + 1) * H, &zerr)) {
            *err = path + ": " + zerr;
            return ROUTE_FAILED;
        }
        int max_type = 0;
        bool up = false;
        for (int y = 0; y < H; y++) {
            const int t = slot[(size_t)y * (W + 1)];
            if (t > max_type) max_type = t;
            up |= t == 2;
        }
        if (max_type > 4) {
            *err = path + ": bad PNG filter type";
            return ROUTE_FAILED;
        }
        if (max_type <= 2) return up ? ROUTE_FILTERED_UP : ROUTE_FILTERED;
        unfilter_in_place(slot, W, H);                         // Average / Paeth rows: a serial recurrence, done here
        return ROUTE_PIXELS;
    }
    // everything else cv::imread(path, 0) takes: the general reader
    Image img;
    if (!read_gray_image(path, img, err)) return ROUTE_FAILED;
    if (img.width != W || img.height != H) {
        char msg[256];
        snprintf(msg, sizeof(msg), "%s is %dx%d, expected %dx%d", path.c_str(), img.width, img.height, W, H);
        *err = msg;
        return ROUTE_FAILED;
    }
    memcpy(slot, img.pix.data(), (size_t)W * H);
    return ROUTE_PIXELS;
}

}  // namespace

int decode_scan_image(const std::string &base, const std::string &suffix, int W, int H, uint8_t *slot, std::string *err)
{
    std::string e;
    const Route r = decode_into(base, suffix, W, H, slot, &e);
    if (err) *err = e;
    return r == ROUTE_FILTERED ? 1 : r == ROUTE_FILTERED_UP ? 2 : r == ROUTE_PIXELS ? 3 : 0;
}

bool ingest_scan(slr_engine *eng, const std::string folder[2], const std::string prefix[2], const std::string &suffix,
                 int n, int W, int H, std::string *err)
{
    const int total = 2 * n;
    const size_t slot_bytes = (size_t)(W + 1) * H;
    uint8_t *pinned = (uint8_t *)pinned_scratch(0, (size_t)total * slot_bytes);
    if (!pinned) {
        if (err) *err = slr_last_error();
        return false;
    }
    if (slr_ingest_begin(eng, total) != SLR_OK) {
        if (err) *err = slr_last_error();
        return false;
    }
    std::vector<std::string> errors((size_t)total);
    std::vector<char> failed((size_t)total, 0);
    std::atomic<int> next(0);
    std::mutex engine_mu;   // the engine is thread-compatible: one slr_ call at a time
    auto worker = [&]() {
        for (;;) {
            const int k = next.fetch_add(1);
            if (k >= total) return;
            const int cam = k / n, i = k % n;
            uint8_t *slot = pinned + (size_t)k * slot_bytes;
            std::string e;
            Route r = ROUTE_FAILED;
            try {
                r = decode_into(folder[cam] + prefix[cam] + std::to_string(i), suffix, W, H, slot, &e);
            } catch (const std::exception &ex) {
                e = ex.what();
            }
            if (r != ROUTE_FAILED) {
                std::lock_guard<std::mutex> lk(engine_mu);
                if (slr_ingest_image(eng, k, slot, r != ROUTE_PIXELS, r == ROUTE_FILTERED_UP) != SLR_OK) {
                    e = slr_last_error();
                    r = ROUTE_FAILED;
                }
            }
            if (r == ROUTE_FAILED) {
                errors[k] = "Load Images: Scan Images not found! (" + e + ")";
                failed[k] = 1;
            }
        }
    };
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (nt > (unsigned)total) nt = (unsigned)total;
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nt; t++) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
    for (int k = 0; k < total; k++)
        if (failed[k]) {
            if (err) *err = errors[k];
            return false;
        }
    return true;
}

namespace {
struct CloudBlock {
    size_t cells;
    float *sums;
    uint8_t *counts;
    bool busy;
};
std::mutex g_cloud_mu;
std::vector<CloudBlock> g_cloud;
}  // namespace

bool cloud_storage_acquire(size_t cells, float **sums, uint8_t **counts)
{
    std::lock_guard<std::mutex> lk(g_cloud_mu);
    for (auto &b : g_cloud)
        if (!b.busy && b.cells == cells) {
            b.busy = true;
            *sums = b.sums;
            *counts = b.counts;
            return true;
        }
    void *s = nullptr, *c = nullptr;
    if (slr_host_alloc(&s, cells * 3 * sizeof(float)) != SLR_OK) return false;
    if (slr_host_alloc(&c, cells) != SLR_OK) {
        slr_host_free(s);
        return false;
    }
    g_cloud.push_back({cells, (float *)s, (uint8_t *)c, true});
    *sums = (float *)s;
    *counts = (uint8_t *)c;
    return true;
}

void cloud_storage_release(float *sums, uint8_t *counts)
{
    std::lock_guard<std::mutex> lk(g_cloud_mu);
    int idle = 0;
    for (auto &b : g_cloud) idle += !b.busy;
    for (size_t k = 0; k < g_cloud.size(); k++)
        if (g_cloud[k].sums == sums && g_cloud[k].counts == counts) {
            if (idle >= 2) {   // keep at most two idle blocks: a session alternates between one or two cloud sizes
                slr_host_free(sums);
                slr_host_free(counts);
                g_cloud.erase(g_cloud.begin() + (long)k);
            } else {
                g_cloud[k].busy = false;
            }
            return;
        }
}

}  // namespace duke
