// imageio.h — reads the scan images the reference loads with cv::imread(path, 0)
// (Duke/mfreconstruct.cpp:125, Duke/reconstruct.cpp:164): 8-bit PNG (gray, gray+alpha, RGB, RGBA, palette-free,
// non-interlaced) decoded with zlib, and binary PGM (P5).  Colour inputs are reduced to gray with OpenCV's
// fixed-point BT.601 weights.  Also writes PNG/PGM (used by tests to lay out a project directory).
#pragma once
#include <string>

#include "duke_types.h"

namespace duke {
bool read_gray_image(const std::string &path, Image &out, std::string *err = nullptr);
bool write_png_gray(const std::string &path, const uint8_t *pix, int w, int h);
bool write_pgm(const std::string &path, const uint8_t *pix, int w, int h);
}  // namespace duke
