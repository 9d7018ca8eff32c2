// Headless driver for the UNMODIFIED reference renderer (test infrastructure).
// Follows /root/reference/app/main.cpp:43-126 with the nanogui window (lines 84-109)
// removed: Embree device + scene, Job, parseScene, Job::integrator(), Integrator::run.
// Adds: --root DIR (the reference chdir("..")s into its repo root, main.cpp:60),
// wall-clock timing around Integrator::run, optional fp32 dump of the final image.
//
//   pathed_ref_headless --root DIR job.json [--raw out.f32] [--warmup warmup_job.json]
//
// --warmup runs Integrator::run once with another job file (same scene and resolution, fewer spp) before the
// timed run, for bench.py's warm-up steps.
//
// The raw dump is Image::m_raw as the reference stores it: 3*W*H floats, RGB,
// scanline 0 = TOP of the image (Image::set flips, src/image.cpp:21-26).
#define private public
#include "image.h"
#undef private

#include "globals.h"
#include "integrator.h"
#include "job.h"
#include "render_status.h"
#include "scene.h"
#include "scene_parser.h"

#include <embree3/rtcore.h>
#define STB_IMAGE_WRITE_IMPLEMENTATION
#include "stb_image_write.h"
#define STB_IMAGE_IMPLEMENTATION
#include "stb_image.h"
#define TINYEXR_IMPLEMENTATION
#include "tinyexr.h"

#include <omp.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>

Job *g_job;
RTCDevice g_rtcDevice;
RTCScene g_rtcScene;

int main(int argc, char *argv[])
{
    std::string root = ".", jobPath = "job.json", rawPath, warmupPath;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--root") && i + 1 < argc) { root = argv[++i]; }
        else if (!strcmp(argv[i], "--raw") && i + 1 < argc) { rawPath = argv[++i]; }
        else if (!strcmp(argv[i], "--warmup") && i + 1 < argc) { warmupPath = argv[++i]; }
        else { jobPath = argv[i]; }
    }

    g_rtcDevice = rtcNewDevice(NULL);
    g_rtcScene = rtcNewScene(g_rtcDevice);
    if (!g_rtcDevice || !g_rtcScene) { fprintf(stderr, "embree init failed\n"); return 1; }

    // job path is interpreted before the chdir if absolute, else relative to root
    if (chdir(root.c_str()) != 0) { perror("chdir"); return 1; }

    std::ifstream jsonJob(jobPath);
    if (!jsonJob) { fprintf(stderr, "cannot open job %s\n", jobPath.c_str()); return 1; }
    g_job = new Job(jsonJob);
    g_job->init();

    const int width = g_job->width();
    const int height = g_job->height();
    Image image(width, height);

    const auto t0 = std::chrono::steady_clock::now();
    std::ifstream jsonScene(g_job->scene());
    if (!jsonScene) { fprintf(stderr, "cannot open scene %s\n", g_job->scene().c_str()); return 1; }
    Scene scene = parseScene(jsonScene);
    const auto t1 = std::chrono::steady_clock::now();

    std::shared_ptr<Integrator> integrator = g_job->integrator();

    bool quit = false;
    if (!warmupPath.empty()) {
        std::ifstream warmupFile(warmupPath);
        if (!warmupFile) { fprintf(stderr, "cannot open warm-up job %s\n", warmupPath.c_str()); return 1; }
        Job *timedJob = g_job;
        g_job = new Job(warmupFile);
        g_job->init();
        Image scratch(width, height);
        g_job->integrator()->run(scratch, scene, [](RenderStatus) {}, &quit);
        g_job = timedJob;
    }
    const auto tr = std::chrono::steady_clock::now();
    integrator->run(image, scene, [](RenderStatus) {}, &quit);
    const auto t2 = std::chrono::steady_clock::now();

    const double buildS = std::chrono::duration<double>(t1 - t0).count();
    const double renderS = std::chrono::duration<double>(t2 - tr).count();
    const double samples = double(width) * height * g_job->spp();

    if (!rawPath.empty()) {
        FILE *f = fopen(rawPath.c_str(), "wb");
        if (f) { fwrite(image.m_raw.data(), sizeof(float), image.m_raw.size(), f); fclose(f); }
    }

    printf("REF_RESULT {\"width\": %d, \"height\": %d, \"spp\": %d, \"threads\": %d, "
           "\"scene_build_s\": %.4f, \"render_wall_s\": %.4f, \"msamples_per_s\": %.6f}\n",
           width, height, g_job->spp(), omp_get_max_threads(), buildS, renderS,
           samples / renderS * 1e-6);

    rtcReleaseScene(g_rtcScene);
    rtcReleaseDevice(g_rtcDevice);
    return 0;
}
