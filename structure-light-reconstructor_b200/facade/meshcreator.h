// meshcreator.h — MeshCreator with the reference's interface (Duke/meshcreator.h:7-23): PLY / OBJ export of a
// PointCloudImage with grid-neighbour faces.  The vertex numbering and face enumeration of
// Duke/meshcreator.cpp:16-166 run on the GPU (slr_mesh_index_host, kernels in csrc/k5_mesh.cu); the text is
// formatted by all host threads and is byte-identical to what the reference's `ostream <<` loops write.
#pragma once
#include <string>

#include "pointcloudimage.h"

namespace duke {
bool write_mesh_text(const std::string &path, bool obj, PointCloudImage *pc, const float *vert, const int32_t *src,
                     const int32_t *faces, size_t nv, size_t nf);
}

class MeshCreator {
public:
    MeshCreator(PointCloudImage *in);
    ~MeshCreator();
    void exportObjMesh(const std::string &path);   // reference: QString path
    void exportPlyMesh(const std::string &path);

    // what the last export wrote (0 on failure)
    unsigned long long vertexCount() const { return nv_; }
    unsigned long long faceCount() const { return nf_; }
    bool ok() const { return ok_; }

private:
    bool exportMesh(const std::string &path, bool obj);
    PointCloudImage *cloud;
    int w, h;
    unsigned long long nv_ = 0, nf_ = 0;
    bool ok_ = false;
};
