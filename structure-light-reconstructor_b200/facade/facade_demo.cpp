// facade_demo.cpp — drives the drop-in classes exactly as MainWindow::startreconstruct does
// (Duke/mainwindow.cpp:577-637) on a project directory and dumps the resulting PointCloudImage (+ the Q matrix
// and rectification maps) for the tests:  facade_demo <mf|ge|gray> <project> <sn> <scanw> <scanh> <camw> <camh>
//                                                      <black> <white> <havecolor> <out.bin>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include <chrono>

#include "meshcreator.h"
#include "mfreconstruct.h"
#include "reconstruct.h"

static void dump(const char *path, PointCloudImage *pc, stereoRect *sr)
{
    FILE *f = fopen(path, "wb");
    const int w = pc->getWidth(), h = pc->getHeight();
    fwrite(&w, 4, 1, f);
    fwrite(&h, 4, 1, f);
    fwrite(pc->sums().data(), sizeof(float), pc->sums().size(), f);
    fwrite(pc->counts().data(), 1, pc->counts().size(), f);
    int has_sr = sr ? 1 : 0;
    fwrite(&has_sr, 4, 1, f);
    if (sr) {
        fwrite(sr->Q.v.data(), sizeof(double), 16, f);
        fwrite(sr->map1().data(), sizeof(int16_t), sr->map1().size(), f);
        fwrite(sr->map2().data(), sizeof(uint16_t), sr->map2().size(), f);
    }
    // a few getPoint() probes (mean = sum * (1.f / count))
    for (int k = 0; k < 16; k++) {
        duke::Point3f p;
        const int i = (k * 7919) % w, j = (k * 104729) % h;
        float rec[4] = {0, 0, 0, 0};
        if (pc->getPoint(i, j, p)) rec[0] = 1, rec[1] = p.x, rec[2] = p.y, rec[3] = p.z;
        fwrite(rec, sizeof(float), 4, f);
    }
    fclose(f);
}

static double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// MainWindow::startreconstruct's export step (mainwindow.cpp:637-646): <project>/reconstruction/<sn>.ply when the
// environment variable DUKE_EXPORT_PLY names a file
static void maybe_export(PointCloudImage *pc)
{
    const char *path = getenv("DUKE_EXPORT_PLY");
    if (!path) return;
    const double t0 = now_ms();
    MeshCreator mc(pc);
    mc.exportPlyMesh(path);
    fprintf(stderr, "[facade_demo] exportPlyMesh: %llu vertices, %llu faces, %.1f ms\n", mc.vertexCount(), mc.faceCount(),
            now_ms() - t0);
}

int main(int argc, char **argv)
{
    if (argc != 12) {
        fprintf(stderr, "usage: %s <mf|ge|gray> <project> <sn> <scanw> <scanh> <camw> <camh> <black> <white> <havecolor> <out.bin>\n", argv[0]);
        return 2;
    }
    const std::string kind = argv[1], project = argv[2];
    const int sn = atoi(argv[3]), scanw = atoi(argv[4]), scanh = atoi(argv[5]), camw = atoi(argv[6]), camh = atoi(argv[7]);
    const int black = atoi(argv[8]), white = atoi(argv[9]), havecolor = atoi(argv[10]);
    if (kind == "mf") {
        MFReconstruct *mfr = new MFReconstruct();
        mfr->getParameters(sn, scanw, scanh, camw, camh, black, white, project);
        const int repeat = getenv("DUKE_REPEAT") ? atoi(getenv("DUKE_REPEAT")) : 1;   // later calls reuse context + engine
        for (int rep = 0; rep < repeat; rep++) {
            const double t0 = now_ms();
            if (!mfr->runReconstruction()) return 1;
            fprintf(stderr, "[facade_demo] MFReconstruct::runReconstruction #%d (image files -> PointCloudImage): %.1f ms\n", rep,
                    now_ms() - t0);
        }
        maybe_export(mfr->points3DProjView);
        dump(argv[11], mfr->points3DProjView, mfr->rectifier());
        printf("mf: %llu points\n", mfr->pointCount());
        delete mfr;
    } else {
        Reconstruct *r = new Reconstruct(kind == "ge");
        r->scanSN = sn;
        const bool autocontrast = getenv("DUKE_AUTOCONTRAST") && atoi(getenv("DUKE_AUTOCONTRAST")) != 0;   // Set dialog flag
        r->getParameters(scanw, scanh, camw, camh, autocontrast, havecolor != 0, project);
        r->setCalibPath(project + "/calib/left/", 0);
        r->setCalibPath(project + "/calib/right/", 1);
        if (!r->loadCameras()) return 1;
        r->setBlackThreshold(black);
        r->setWhiteThreshold(white);
        r->disableRaySampling();
        const int repeat = getenv("DUKE_REPEAT") ? atoi(getenv("DUKE_REPEAT")) : 1;   // later calls reuse context + engine
        for (int rep = 0; rep < repeat; rep++) {
            const double t0 = now_ms();
            const bool ok = (kind == "ge") ? r->runReconstruction_GE() : r->runReconstruction();
            if (!ok) return 1;
            fprintf(stderr, "[facade_demo] Reconstruct::runReconstruction%s #%d (image files -> PointCloudImage): %.1f ms\n",
                    kind == "ge" ? "_GE" : "", rep, now_ms() - t0);
        }
        maybe_export(r->points3DProjView);
        dump(argv[11], r->points3DProjView, kind == "ge" ? r->rectifier() : nullptr);
        delete r;
    }
    return 0;
}
