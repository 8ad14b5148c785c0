// reconstruct_common.h — plumbing shared by the Reconstruct / MFReconstruct facades: engine ownership, camera
// upload, image stack loading into pinned host memory.
#pragma once
#include <stdint.h>

#include <string>

#include "slr_b200.h"
#include "stereorect.h"
#include "virtualcamera.h"

namespace duke {

slr_camera to_slr_camera(const VirtualCamera &vc);

// Loads <folder><prefix><i><suffix>, i in [0, n), into dst[cam][i][H][W] (reference naming:
// scan/left/<sn>/L<i>.png).  Falls back to ".pgm" when the ".png" file is absent.  Returns false on the first
// missing / mismatching image, after printing what the reference shows in a message box.
bool load_stack(const std::string &folder, const std::string &prefix, const std::string &suffix, int n, int W, int H,
                uint8_t *dst);
// Both cameras' stacks in one pass over all host threads (2 x n images share one work queue: three rounds of decoding
// instead of four for 2 x 24 Gray-code images on 16 threads): dst[cam][i][H][W].
bool load_stacks(const std::string folders[2], const std::string prefixes[2], const std::string &suffix, int cams, int n_per_cam,
                 int W, int H, uint8_t *dst);

// Process-wide engine for (device, W, H): created on first use and kept, so that only the first reconstruction of a
// session pays for CUDA context creation (~0.5 s) and the device / pinned allocations.  The reference news a
// reconstructor per scan (mainwindow.cpp:577-649) on one GUI thread; so do the facades, which is why a plain
// mutex-protected cache is enough.  Returns nullptr (slr_last_error() says why) when no engine can be created.
slr_engine *shared_engine(int device, int W, int H);
// Grow-only pinned host buffers shared by the facades (slot 0 image stack, 1 xyz / sums, 2 valid / counts, 3 colour)
void *pinned_scratch(int slot, size_t bytes);

// slr_set_calib + slr_set_rectify_maps, skipped when this engine already holds exactly these values (cal_id =
// stereoRect::calibrationId() of the maps; h_map1 == nullptr: no maps).  The reference builds a new reconstructor per
// scan with the same project calibration; re-uploading 16 MB of maps and rebuilding the undistortPoints tables
// each time would cost more than the scan.
bool upload_calibration(slr_engine *eng, const slr_camera cams[2], const double Q[16], const float *rigid3x4,
                        unsigned long long cal_id, const int16_t *h_map1, const uint16_t *h_map2);

// 3x4 matrix of scan/transfer_mat<sn>.txt (mfreconstruct.cpp:276-282); false if unreadable
bool load_rigid(const std::string &path, float out[12]);

}  // namespace duke
