// ingest.h — scan images from disk to the GPU: the drop-in's replacement for the
// cv::imread(path, 0) loop of MFReconstruct::loadCamImgs (Duke/mfreconstruct.cpp:119-134).
#pragma once
#include <stdint.h>

#include <string>

#include "slr_b200.h"

namespace duke {

// Streams the 2 x n images <folder[c]><prefix[c]><i><suffix> (".pgm" where the ".png" is absent), i in [0, n), of an
// MF scan into the engine between slr_ingest_begin and slr_run_mf_ingested: every host thread takes images off a
// shared counter, parses the file, inflates its zlib stream with the decoder of inflate.cpp straight into pinned
// memory and hands the scanlines to the GPU (copy + PNG unfiltering there) while the other threads are still
// decoding.  8-bit grey PNGs whose rows use filters None / Sub / Up (what OpenCV's encoder writes) go up as they are;
// anything else (Average / Paeth rows, colour, alpha, PGM) is finished on the host thread and goes up as pixels.
// Returns false, with the message the reference would show, for the lowest-numbered image that is missing or has the
// wrong size.
bool ingest_scan(slr_engine *eng, const std::string folder[2], const std::string prefix[2], const std::string &suffix,
                 int n, int W, int H, std::string *err);

// The host half of one image, as ingest_scan's workers run it: <base><suffix> (or <base>.pgm) into `slot` (room for
// H * (1 + W) bytes).  Returns 0 on failure (*err says why), 1 = PNG scanlines with filter types 0 / 1 only, 2 = PNG
// scanlines that also have Up rows (both: H x [type][W bytes], to be unfiltered on the GPU), 3 = H x W finished pixels.
int decode_scan_image(const std::string &base, const std::string &suffix, int W, int H, uint8_t *slot, std::string *err);

// Pinned sums / counts of a PointCloudImage(w, h) from a small process-wide pool, and their way back
// (PointCloudImage's release hook).  nullptr when pinned memory cannot be had.
bool cloud_storage_acquire(size_t cells, float **sums, uint8_t **counts);
void cloud_storage_release(float *sums, uint8_t *counts);

}  // namespace duke
