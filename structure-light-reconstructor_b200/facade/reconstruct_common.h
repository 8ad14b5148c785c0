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

// 3x4 matrix of scan/transfer_mat<sn>.txt (mfreconstruct.cpp:276-282); false if unreadable
bool load_rigid(const std::string &path, float out[12]);

}  // namespace duke
