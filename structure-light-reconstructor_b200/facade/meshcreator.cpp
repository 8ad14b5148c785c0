#include "meshcreator.h"

#include <stdio.h>
#include <string.h>

#include <charconv>
#include <thread>
#include <vector>

#include "reconstruct_common.h"
#include "slr_b200.h"

MeshCreator::MeshCreator(PointCloudImage *in) : cloud(in), w(in->getWidth()), h(in->getHeight()) {}
MeshCreator::~MeshCreator() {}

namespace {

// `ostream << float` with the default format == printf("%g"): std::to_chars(general, precision 6) produces the
// same characters without locale or stream overhead.
inline char *put_float(char *p, float v)
{
    auto r = std::to_chars(p, p + 32, (double)v, std::chars_format::general, 6);
    return r.ptr;
}
inline char *put_int(char *p, int v)
{
    auto r = std::to_chars(p, p + 16, v);
    return r.ptr;
}

// format items [lo, hi) with fn(char*, index) -> char* into one buffer per worker, all host threads
template <typename Fn>
std::vector<std::string> format_parallel(size_t n, size_t max_bytes_per_item, Fn fn)
{
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (nt > 64) nt = 64;
    if (n < 4096) nt = 1;
    std::vector<std::string> out(nt);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++) {
        const size_t lo = n * t / nt, hi = n * (t + 1) / nt;
        th.emplace_back([&, t, lo, hi] {
            std::string &s = out[t];
            s.resize((hi - lo) * max_bytes_per_item);
            char *p = &s[0];
            for (size_t k = lo; k < hi; k++) p = fn(p, k);
            s.resize((size_t)(p - &s[0]));
        });
    }
    for (auto &x : th) x.join();
    return out;
}

}  // namespace

namespace duke {

// The text of exportPlyMesh / exportObjMesh (meshcreator.cpp:112-163 / :30-62) from the index arrays of
// slr_mesh_index: nv vertices (x y z, source element j*w+i for the colour lookup), nf faces.
bool write_mesh_text(const std::string &path, bool obj, PointCloudImage *pc, const float *vert, const int32_t *src,
                     const int32_t *faces, size_t nv, size_t nf)
{
    const int W = pc->getWidth();
    FILE *fp = fopen(path.c_str(), "wb");
    if (!fp) return false;
    if (!obj) {   // meshcreator.cpp:112-123
        fprintf(fp, "ply\nformat ascii 1.0\nelement vertex %zu\n", nv);
        fprintf(fp, "property float x\nproperty float y\nproperty float z\n");
        fprintf(fp, "property uchar red\nproperty uchar green\nproperty uchar blue\n");
        fprintf(fp, "element face %zu\nproperty list uchar int vertex_indices\nend_header\n", nf);
    }
    auto vtext = format_parallel(nv, 3 * 32 + 3 * 12 + 8, [&](char *p, size_t v) {
        if (obj) *p++ = 'v', *p++ = ' ';                       // "v x y z"            (:30)
        p = put_float(p, vert[3 * v + 0]), *p++ = ' ';
        p = put_float(p, vert[3 * v + 1]), *p++ = ' ';
        p = put_float(p, vert[3 * v + 2]);
        if (!obj) {                                             // "x y z c2 c1 c0"     (:129)
            duke::Point3f pt;
            duke::Vec3i c;
            pc->getPoint(src[v] % W, src[v] / W, pt, c);
            *p++ = ' ', p = put_int(p, c[2]), *p++ = ' ', p = put_int(p, c[1]), *p++ = ' ', p = put_int(p, c[0]);
        }
        *p++ = '\n';
        return p;
    });
    for (auto &s : vtext) fwrite(s.data(), 1, s.size(), fp);
    auto ftext = format_parallel(nf, 6 * 12 + 16, [&](char *p, size_t f) {
        const int32_t *t = &faces[3 * f];
        if (obj) {                                              // "f a/a b/b c/c"      (:47, :60)
            *p++ = 'f';
            for (int k = 0; k < 3; k++) *p++ = ' ', p = put_int(p, t[k]), *p++ = '/', p = put_int(p, t[k]);
        } else {                                                // "3 a b c"            (:151, :159)
            *p++ = '3';
            for (int k = 0; k < 3; k++) *p++ = ' ', p = put_int(p, t[k]);
        }
        *p++ = '\n';
        return p;
    });
    for (auto &s : ftext) fwrite(s.data(), 1, s.size(), fp);
    return fclose(fp) == 0;
}

}  // namespace duke

bool MeshCreator::exportMesh(const std::string &path, bool obj)
{
    ok_ = false;
    nv_ = nf_ = 0;
    const size_t px = (size_t)w * h;
    std::vector<float> vert(px * 3);
    std::vector<int32_t> src(px), faces(px * 6);
    unsigned long long counts[2] = {0, 0};

    slr_engine *eng = duke::shared_engine(0, w, h);
    if (!eng) {
        fprintf(stderr, "MeshCreator: %s\n", slr_last_error());
        return false;
    }
    const slr_status st = slr_mesh_index_host(eng, cloud->sums().data(), cloud->counts().data(), w, h, obj ? 1 : 0,
                                              vert.data(), src.data(), faces.data(), counts);
    if (st != SLR_OK) fprintf(stderr, "MeshCreator: %s\n", slr_last_error());
    if (st != SLR_OK) return false;
    const size_t nv = (size_t)counts[0], nf = (size_t)counts[1];

    const bool good = duke::write_mesh_text(path, obj, cloud, vert.data(), src.data(), faces.data(), nv, nf);
    nv_ = nv;
    nf_ = nf;
    ok_ = good;
    return good;
}

void MeshCreator::exportObjMesh(const std::string &path) { exportMesh(path, true); }
void MeshCreator::exportPlyMesh(const std::string &path) { exportMesh(path, false); }
