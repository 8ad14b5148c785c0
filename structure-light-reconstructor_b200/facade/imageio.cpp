#include "imageio.h"

#include "inflate.h"

#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <vector>

namespace duke {

static bool fail(std::string *err, const std::string &m)
{
    if (err) *err = m;
    return false;
}

static bool read_file(const std::string &path, std::vector<uint8_t> &buf)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize(n > 0 ? (size_t)n : 0);
    const size_t got = n > 0 ? fread(buf.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    return got == buf.size();
}

static uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

static int paeth(int a, int b, int c)
{
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

static bool read_png(const std::vector<uint8_t> &buf, Image &out, std::string *err)
{
    static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    if (buf.size() < 33 || memcmp(buf.data(), sig, 8) != 0) return fail(err, "not a PNG");
    size_t pos = 8;
    uint32_t W = 0, H = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat;
    while (pos + 12 <= buf.size()) {
        const uint32_t len = be32(&buf[pos]);
        const uint8_t *type = &buf[pos + 4], *data = &buf[pos + 8];
        if (pos + 12 + len > buf.size()) return fail(err, "truncated PNG chunk");
        if (!memcmp(type, "IHDR", 4)) {
            if (len < 13) return fail(err, "PNG IHDR chunk too short");
            W = be32(data);
            H = be32(data + 4);
            depth = data[8];
            ctype = data[9];
            interlace = data[12];
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + len;
    }
    // the file is not trusted: sizes are checked before anything is allocated from them
    if (W == 0 || H == 0 || W > 65535 || H > 65535) return fail(err, "PNG without a usable IHDR (size 0 or beyond 65535)");
    if (depth != 8 || interlace != 0) return fail(err, "only 8-bit non-interlaced PNG is supported");
    int ch;
    switch (ctype) {
    case 0: ch = 1; break;
    case 2: ch = 3; break;
    case 4: ch = 2; break;
    case 6: ch = 4; break;
    default: return fail(err, "palette PNG is not supported");
    }
    const size_t stride = (size_t)W * ch;
    std::vector<uint8_t> raw((stride + 1) * H);
    std::string zerr;
    if (!zlib_inflate(idat.data(), idat.size(), raw.data(), raw.size(), &zerr)) return fail(err, "PNG " + zerr);
    std::vector<uint8_t> img(stride * H);
    for (uint32_t y = 0; y < H; y++) {
        const uint8_t ft = raw[y * (stride + 1)];
        const uint8_t *src = &raw[y * (stride + 1) + 1];
        uint8_t *dst = &img[y * stride];
        const uint8_t *up = y ? &img[(y - 1) * stride] : nullptr;
        for (size_t x = 0; x < stride; x++) {
            const int a = x >= (size_t)ch ? dst[x - ch] : 0, b = up ? up[x] : 0, c = (up && x >= (size_t)ch) ? up[x - ch] : 0;
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
    out.width = (int)W;
    out.height = (int)H;
    out.pix.resize((size_t)W * H);
    for (size_t p = 0; p < (size_t)W * H; p++) {
        const uint8_t *px = &img[p * ch];
        if (ch <= 2)
            out.pix[p] = px[0];
        else  // OpenCV's BGR2GRAY fixed point: (R*4899 + G*9617 + B*1868 + 8192) >> 14
            out.pix[p] = (uint8_t)((px[0] * 4899 + px[1] * 9617 + px[2] * 1868 + 8192) >> 14);
    }
    return true;
}

static bool read_pgm(const std::vector<uint8_t> &buf, Image &out, std::string *err)
{
    size_t pos = 2;
    int vals[3], n = 0;
    while (n < 3 && pos < buf.size()) {
        while (pos < buf.size() && (buf[pos] == ' ' || buf[pos] == '\n' || buf[pos] == '\r' || buf[pos] == '\t')) pos++;
        if (pos < buf.size() && buf[pos] == '#') {
            while (pos < buf.size() && buf[pos] != '\n') pos++;
            continue;
        }
        int v = 0;
        bool any = false;
        while (pos < buf.size() && buf[pos] >= '0' && buf[pos] <= '9') v = v * 10 + (buf[pos++] - '0'), any = true;
        if (!any) return fail(err, "bad PGM header");
        vals[n++] = v;
    }
    pos++;  // single whitespace after maxval
    if (n != 3 || vals[2] != 255 || pos + (size_t)vals[0] * vals[1] > buf.size()) return fail(err, "unsupported PGM");
    out.width = vals[0];
    out.height = vals[1];
    out.pix.assign(buf.begin() + pos, buf.begin() + pos + (size_t)vals[0] * vals[1]);
    return true;
}

bool read_gray_image(const std::string &path, Image &out, std::string *err)
{
    std::vector<uint8_t> buf;
    if (!read_file(path, buf)) return fail(err, "cannot read " + path);
    if (buf.size() > 2 && buf[0] == 'P' && buf[1] == '5') return read_pgm(buf, out, err);
    return read_png(buf, out, err);
}

static void put_be32(std::vector<uint8_t> &v, uint32_t x)
{
    v.push_back(x >> 24);
    v.push_back(x >> 16);
    v.push_back(x >> 8);
    v.push_back(x);
}
static void put_chunk(std::vector<uint8_t> &out, const char *type, const std::vector<uint8_t> &data)
{
    put_be32(out, (uint32_t)data.size());
    std::vector<uint8_t> td(type, type + 4);
    td.insert(td.end(), data.begin(), data.end());
    out.insert(out.end(), td.begin(), td.end());
    put_be32(out, (uint32_t)crc32(0L, td.data(), (uInt)td.size()));
}

bool write_png_gray(const std::string &path, const uint8_t *pix, int w, int h)
{
    std::vector<uint8_t> out = {137, 80, 78, 71, 13, 10, 26, 10}, ihdr;
    put_be32(ihdr, (uint32_t)w);
    put_be32(ihdr, (uint32_t)h);
    const uint8_t tail[5] = {8, 0, 0, 0, 0};
    ihdr.insert(ihdr.end(), tail, tail + 5);
    put_chunk(out, "IHDR", ihdr);
    std::vector<uint8_t> raw((size_t)(w + 1) * h);
    for (int y = 0; y < h; y++) {
        raw[(size_t)y * (w + 1)] = 0;
        memcpy(&raw[(size_t)y * (w + 1) + 1], pix + (size_t)y * w, (size_t)w);
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 1) != Z_OK) return false;
    comp.resize(clen);
    put_chunk(out, "IDAT", comp);
    put_chunk(out, "IEND", {});
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    return ok;
}

bool write_pgm(const std::string &path, const uint8_t *pix, int w, int h)
{
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return false;
    fprintf(f, "P5\n%d %d\n255\n", w, h);
    const bool ok = fwrite(pix, 1, (size_t)w * h, f) == (size_t)w * h;
    fclose(f);
    return ok;
}

}  // namespace duke
