// virtualcamera.h — VirtualCamera with the reference's member names (Duke/virtualcamera.h:13-41).
// Qt-free: QString -> std::string, cv::Mat -> duke::Matrix, cv::Point*f -> duke::Point*f.
#pragma once
#include <string>

#include "duke_types.h"

class VirtualCamera {
public:
    VirtualCamera();
    ~VirtualCamera();
    void loadDistortion(const std::string &path);          // 5x1      (virtualcamera.cpp:25-28)
    bool loadCameraMatrix(const std::string &path);        // 3x3 -> fc, cc; false if the file is missing (:30-45)
    void loadRotationMatrix(const std::string &path);      // 3x3
    void loadTranslationVector(const std::string &path);   // 3x1
    void loadFundamentalMatrix(const std::string &path);   // 3x3
    void loadHomoMatrix(const std::string &path, int i);   // 3x3, i = 1 | 2
    // whitespace separated decimals, row major, each parsed as float (virtualcamera.cpp:70-88); -1 if unreadable
    int loadMatrix(duke::Matrix &matrix, int rows, int cols, const std::string &file);

    duke::Matrix distortion, rotationMatrix, translationVector, fundamentalMatrix, homoMat1, homoMat2;
    duke::Point3f position;
    duke::Point2f fc, cc;
    int width = 0, height = 0;
};
