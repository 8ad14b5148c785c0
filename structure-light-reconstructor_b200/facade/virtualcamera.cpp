#include "virtualcamera.h"

#include <fstream>

VirtualCamera::VirtualCamera() {}
VirtualCamera::~VirtualCamera() {}

void VirtualCamera::loadDistortion(const std::string &path) { loadMatrix(distortion, 5, 1, path); }

bool VirtualCamera::loadCameraMatrix(const std::string &path)
{
    std::ifstream probe(path.c_str());
    if (!probe) {
        fprintf(stderr, "Matrix not found: File '%s' need to be added.\n", path.c_str());  // QMessageBox in the reference
        return false;
    }
    duke::Matrix m;
    loadMatrix(m, 3, 3, path);
    cc.x = (float)m.at(0, 2);
    cc.y = (float)m.at(1, 2);
    fc.x = (float)m.at(0, 0);
    fc.y = (float)m.at(1, 1);
    return true;
}

void VirtualCamera::loadRotationMatrix(const std::string &path) { loadMatrix(rotationMatrix, 3, 3, path); }
void VirtualCamera::loadTranslationVector(const std::string &path) { loadMatrix(translationVector, 3, 1, path); }
void VirtualCamera::loadFundamentalMatrix(const std::string &path) { loadMatrix(fundamentalMatrix, 3, 3, path); }
void VirtualCamera::loadHomoMatrix(const std::string &path, int i) { loadMatrix(i == 1 ? homoMat1 : homoMat2, 3, 3, path); }

int VirtualCamera::loadMatrix(duke::Matrix &matrix, int rows, int cols, const std::string &file)
{
    std::ifstream in1(file.c_str());
    if (!in1) return -1;
    matrix = duke::Matrix(rows, cols);
    for (int i = 0; i < rows; i++)
        for (int j = 0; j < cols; j++) {
            float val = 0;  // the reference reads `float val; in1 >> val;` into a CV_32F Mat
            in1 >> val;
            matrix.at(i, j) = val;
        }
    return 1;
}
