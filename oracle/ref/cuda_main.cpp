// pathed_ref_cuda -- test infrastructure: the UNMODIFIED reference (its Job, scene / OBJ / PLY parsers, Image, Integrator::run)
// rendering through libpathed_cuda.  The reference's sources are compiled exactly as for pathed_ref_headless (oracle/ref/build_ref.sh);
// instead of Embree they are linked against oracle/ref/embree_shim.cpp, which records the geometry the parsers hand to the rtc* API.
// After the reference's own parseScene this harness replays what it built -- geometry from the shim, materials / camera / lights /
// media from the reference's objects -- into the C ABI of include/pathed_cuda.h and renders with a CudaPathTracer that is a subclass of
// the reference's Integrator: the binding INTEGRATION.md describes, made at the library seams instead of by patching sources.
//
//   pathed_ref_cuda --root DIR job.json [--raw out.f32]
//
// Output: the reference's own files (Image::save / saveCheckpoint through Integrator::run), REF_CUDA_RESULT {...} on stdout, and with
// --raw Image::m_raw (3*W*H floats, scanline 0 = top).  A scene through this binary and through pathed_b200/pathed (the repository's own
// host layer) must give the same image bit for bit: same geometry, materials and Philox streams (tests/test_gpu_host.py).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>
#include <unistd.h>

#include <embree3/rtcore.h>

// the reference keeps material / camera / light parameters private without getters (include/lambertian.h:34-36, glass.h:26-27, ...):
// this translation unit reads them where they lie (class layout does not depend on access specifiers)
#define private public
#define protected public
#include "beckmann.h"
#include "camera.h"
#include "checkerboard.h"
#include "environment_light.h"
#include "ggx.h"
#include "glass.h"
#include "globals.h"
#include "homogeneous_medium.h"
#include "image.h"
#include "integrator.h"
#include "job.h"
#include "lambertian.h"
#include "microfacet.h"
#include "mirror.h"
#include "oren_nayar.h"
#include "passthrough.h"
#include "plastic.h"
#include "render_status.h"
#include "rtc_manager.h"
#include "scene.h"
#include "scene_parser.h"
#include "texture.h"
#include "transform.h"
#undef private
#undef protected

#define STB_IMAGE_WRITE_IMPLEMENTATION
#include "stb_image_write.h"
#define STB_IMAGE_IMPLEMENTATION
#include "stb_image.h"
#define TINYEXR_IMPLEMENTATION
#include "tinyexr.h"

#include "embree_shim.h"
#include "pathed_cuda.h"

Job *g_job;
RTCDevice g_rtcDevice;
RTCScene g_rtcScene;

namespace {

ptc_ctx *g_ptc = nullptr;

void check(int status, const char *what)
{
    if (status != PTC_OK) { fprintf(stderr, "%s: %s\n", what, g_ptc ? ptc_last_error(g_ptc) : "no context"); exit(1); }
}

// OrenNayar keeps A and B, not sigma (src/oren_nayar.cpp:11-19); the ABI takes sigma and derives A and B with the same float
// arithmetic, so a sigma is searched that reproduces both exactly
float sigmaFor(float A, float B)
{
    auto derive = [](float sigma, float &a, float &b) {
        const float sigma2 = sigma * sigma;
        a = 1.f - (sigma2 / (2.f * (sigma2 + 0.33f)));
        b = (0.45f * sigma2) / (sigma2 + 0.09f);
    };
    const double s2 = 0.66 * (1.0 - (double)A) / (1.0 - 2.0 * (1.0 - (double)A)); // A = 1 - s2 / (2 (s2 + 0.33))
    float guess = (float)std::sqrt(std::max(0.0, s2));
    float lo = guess, hi = guess;
    for (int step = 0; step < 4096; step++) {
        float a, b;
        derive(lo, a, b); if (a == A && b == B) { return lo; }
        derive(hi, a, b); if (a == A && b == B) { return hi; }
        lo = std::nextafterf(lo, 0.f); hi = std::nextafterf(hi, 1e30f);
    }
    fprintf(stderr, "no sigma reproduces OrenNayar A = %.9g, B = %.9g\n", A, B);
    exit(1);
}

struct Binding {
    std::map<const Material *, uint32_t> materials;
    std::map<const Texture *, uint32_t> textures;
    std::map<const Medium *, uint32_t> media;
    std::map<const ShimScene *, uint32_t> scenes; // instance scenes already described (ptc_begin_instance ids)
};

void albedoInto(const std::shared_ptr<Albedo> &albedo, ptc_material_desc &d, Binding &binding)
{
    if (!albedo) { return; }
    if (const Checkerboard *c = dynamic_cast<const Checkerboard *>(albedo.get())) {
        d.albedo_kind = PTC_ALBEDO_CHECKERBOARD;
        d.checker_on[0] = c->m_onColor.r(); d.checker_on[1] = c->m_onColor.g(); d.checker_on[2] = c->m_onColor.b();
        d.checker_off[0] = c->m_offColor.r(); d.checker_off[1] = c->m_offColor.g(); d.checker_off[2] = c->m_offColor.b();
        d.checker_resolution[0] = c->m_resolution.u; d.checker_resolution[1] = c->m_resolution.v;
        return;
    }
    if (Texture *t = dynamic_cast<Texture *>(albedo.get())) {
        if (!binding.textures.count(t)) {
            if (!t->m_data) { t->load(); }
            uint32_t id = 0;
            check(ptc_add_texture(g_ptc, t->m_data, t->m_width, t->m_height, &id), "ptc_add_texture"); // stbi_load(..., 3): RGB
            binding.textures[t] = id;
        }
        d.albedo_kind = PTC_ALBEDO_TEXTURE; d.texture = binding.textures[t];
        return;
    }
    fprintf(stderr, "an albedo class the ABI does not know\n"); exit(1);
}

void distributionInto(const MicrofacetDistribution *distribution, ptc_material_desc &d)
{
    if (const Beckmann *b = dynamic_cast<const Beckmann *>(distribution)) { d.distribution = PTC_BECKMANN; d.alpha = b->m_alpha; return; }
    if (const GGX *g = dynamic_cast<const GGX *>(distribution)) { d.distribution = PTC_GGX; d.alpha = g->m_alpha; return; }
    fprintf(stderr, "a microfacet distribution the ABI does not know\n"); exit(1);
}

uint32_t materialId(const Material *material, Binding &binding)
{
    auto found = binding.materials.find(material);
    if (found != binding.materials.end()) { return found->second; }
    ptc_material_desc d;
    memset(&d, 0, sizeof(d));
    d.ior = 1.4f; d.albedo_kind = PTC_ALBEDO_CONSTANT;
    d.emit[0] = material->m_emit.r(); d.emit[1] = material->m_emit.g(); d.emit[2] = material->m_emit.b();
    auto diffuse = [&](const Color &c) { d.diffuse[0] = c.r(); d.diffuse[1] = c.g(); d.diffuse[2] = c.b(); };
    if (const Lambertian *l = dynamic_cast<const Lambertian *>(material)) { d.type = PTC_LAMBERTIAN; diffuse(l->m_diffuse); albedoInto(l->m_albedo, d, binding); }
    else if (const OrenNayar *o = dynamic_cast<const OrenNayar *>(material)) { d.type = PTC_OREN_NAYAR; diffuse(o->m_diffuse); d.sigma = sigmaFor(o->m_A, o->m_B); }
    else if (dynamic_cast<const Mirror *>(material)) { d.type = PTC_MIRROR; }
    else if (const Glass *g = dynamic_cast<const Glass *>(material)) { d.type = PTC_GLASS; d.ior = g->m_ior; }
    else if (const Microfacet *m = dynamic_cast<const Microfacet *>(material)) { d.type = PTC_MICROFACET; distributionInto(m->m_distributionPtr.get(), d); }
    else if (const Plastic *p = dynamic_cast<const Plastic *>(material)) {
        d.type = PTC_PLASTIC;
        diffuse(p->m_lambertianPtr->m_diffuse); albedoInto(p->m_lambertianPtr->m_albedo, d, binding);
        distributionInto(p->m_microfacet.m_distributionPtr.get(), d);
    }
    else if (dynamic_cast<const Passthrough *>(material)) { d.type = PTC_PASSTHROUGH; }
    else { fprintf(stderr, "a material class the ABI does not know\n"); exit(1); }
    uint32_t id = 0;
    check(ptc_add_material(g_ptc, &d, &id), "ptc_add_material");
    binding.materials[material] = id;
    return id;
}

// one recorded scene into the context, geometry by geometry in attach order (= geometry ids on both sides); an instance scene is
// described (ptc_begin_instance .. ptc_end_instance) right before its first placement
void describe(const ShimScene *scene, const RTCManager &manager, Binding &binding, bool root)
{
    const NestedSurfaceVector &surfaces = manager.m_rtcSceneToSurfaces.at((RTCScene)scene);
    for (size_t geomID = 0; geomID < scene->geometries.size(); geomID++) {
        const ShimGeometry *g = scene->geometries[geomID];
        uint32_t id = 0;
        if (g->type == RTC_GEOMETRY_TYPE_INSTANCE) {
            if (!binding.scenes.count(g->instanced)) {
                uint32_t sceneId = 0;
                check(ptc_begin_instance(g_ptc, &sceneId), "ptc_begin_instance");
                binding.scenes[g->instanced] = sceneId;
                describe(g->instanced, manager, binding, false);
                check(ptc_end_instance(g_ptc), "ptc_end_instance");
            }
            check(ptc_add_instance(g_ptc, binding.scenes[g->instanced], g->transform, &id), "ptc_add_instance");
        } else if (g->type == RTC_GEOMETRY_TYPE_TRIANGLE) {
            const std::vector<std::shared_ptr<Surface>> &faces = surfaces.at(geomID);
            if (faces.size() != g->indexCount) { fprintf(stderr, "geometry %zu: %zu surfaces for %zu triangles\n", geomID, faces.size(), g->indexCount); exit(1); }
            std::vector<float> P(3 * g->vertexCount), N(3 * g->vertexCount, 0.f), UV(2 * g->vertexCount, 0.f);
            for (size_t v = 0; v < g->vertexCount; v++) {
                memcpy(&P[3 * v], &g->vertices[v * g->vertexStride], 12);
                if (g->attributeCount[0] == g->vertexCount) { memcpy(&UV[2 * v], &g->attributes[0][v * g->attributeStride[0]], 8); } // slot 0: uvs
                if (g->attributeCount[1] == g->vertexCount) { memcpy(&N[3 * v], &g->attributes[1][v * g->attributeStride[1]], 12); } // slot 1: normals
            }
            std::vector<uint32_t> I(3 * g->indexCount), M(g->indexCount);
            for (size_t t = 0; t < g->indexCount; t++) {
                memcpy(&I[3 * t], &g->indices[t * g->indexStride], 12);
                M[t] = materialId(faces[t]->getMaterial().get(), binding);
            }
            check(ptc_add_triangle_mesh(g_ptc, P.data(), N.data(), UV.data(), (uint32_t)g->vertexCount, I.data(), M.data(), (uint32_t)g->indexCount, &id), "ptc_add_triangle_mesh");
        } else if (g->type == RTC_GEOMETRY_TYPE_SPHERE_POINT) {
            const std::vector<std::shared_ptr<Surface>> &faces = surfaces.at(geomID);
            float centerRadius[4];
            memcpy(centerRadius, g->vertices.data(), 16);
            check(ptc_add_sphere(g_ptc, centerRadius, materialId(faces.at(0)->getMaterial().get(), binding), &id), "ptc_add_sphere");
        } else { fprintf(stderr, "geometry %zu has a type the ABI does not cover (%d)\n", geomID, (int)g->type); exit(1); }
        if (id != geomID) { fprintf(stderr, "geometry ids diverged: Embree-side %zu, ptc %u\n", geomID, id); exit(1); }
        // Surface::m_internalMedium: the same medium for every surface of a model (src/scene_parser.cpp:324-343, :370-381, :503-514)
        if (root && g->type != RTC_GEOMETRY_TYPE_INSTANCE && !surfaces.at(geomID).empty()) {
            const std::shared_ptr<Medium> medium = surfaces.at(geomID)[0]->getInternalMedium();
            if (medium) {
                if (!binding.media.count(medium.get())) {
                    const HomogeneousMedium *h = dynamic_cast<const HomogeneousMedium *>(medium.get());
                    if (!h) { fprintf(stderr, "a medium class the ABI does not know\n"); exit(1); }
                    const float st[3] = {h->m_sigmaT.r(), h->m_sigmaT.g(), h->m_sigmaT.b()}, ss[3] = {h->m_sigmaS.r(), h->m_sigmaS.g(), h->m_sigmaS.b()};
                    uint32_t mediumId = 0;
                    check(ptc_add_medium(g_ptc, st, ss, &mediumId), "ptc_add_medium");
                    binding.media[medium.get()] = mediumId;
                }
                check(ptc_set_internal_medium(g_ptc, id, binding.media[medium.get()]), "ptc_set_internal_medium");
            }
        }
    }
}

void bindScene(Scene &scene)
{
    check(ptc_create(0, &g_ptc), "ptc_create");
    Binding binding;
    describe((const ShimScene *)g_rtcScene, *scene.m_rtcManagerPtr, binding, true);
    if (scene.m_environmentLight) { // EnvironmentLight::EnvironmentLight, src/environment_light.cpp:14-54
        const EnvironmentLight &e = *scene.m_environmentLight;
        check(ptc_set_environment(g_ptc, e.m_data, e.m_width, e.m_height, e.m_scale, &e.m_mapToWorld.m_matrix[0][0], &e.m_worldToMap.m_matrix[0][0]), "ptc_set_environment");
    }
    const Camera &camera = *scene.m_camera;
    // Camera does not keep flipHandedness: it is the one of the two look-at matrices the camera holds (src/camera.cpp:13-30)
    int flip = -1;
    for (int candidate = 0; candidate < 2 && flip < 0; candidate++) {
        const Transform t = lookAt(camera.m_origin, camera.m_target, camera.m_up, candidate != 0);
        if (!memcmp(t.m_matrix, camera.m_cameraToWorld.m_matrix, sizeof(t.m_matrix))) { flip = candidate; }
    }
    if (flip < 0) { fprintf(stderr, "the camera's matrix is neither look-at variant\n"); exit(1); }
    const float origin[3] = {camera.m_origin.x(), camera.m_origin.y(), camera.m_origin.z()};
    const float target[3] = {camera.m_target.x(), camera.m_target.y(), camera.m_target.z()};
    const float up[3] = {camera.m_up.x(), camera.m_up.y(), camera.m_up.z()};
    check(ptc_set_camera(g_ptc, origin, target, up, camera.m_verticalFOV, camera.m_resolution.x, camera.m_resolution.y, flip), "ptc_set_camera");
    check(ptc_commit(g_ptc), "ptc_commit");
}

// one spp per sampleImage call, accumulated into the HOST radianceLookup like src/sample_integrator.cpp:61-63;
// Integrator::run (src/integrator.cpp:19-106: per-sample callback, image.set, power-of-two checkpoints) stays the reference's
class CudaPathTracer : public Integrator {
public:
    CudaPathTracer(int integrator, int startBounce, int lastBounce) : m_start(startBounce), m_last(lastBounce) { check(ptc_set_integrator(g_ptc, integrator), "ptc_set_integrator"); }

protected:
    void sampleImage(std::vector<float> &radianceLookup, std::vector<Sample> &, Scene &, RandomGenerator &) override
    {
        check(ptc_render(g_ptc, 0x5EED, m_sample++, 1, m_start, m_last, radianceLookup.data()), "ptc_render");
    }

private:
    int m_start, m_last;
    uint32_t m_sample = 0;
};

} // namespace

int main(int argc, char *argv[])
{
    std::string root = ".", jobPath = "job.json", rawPath;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--root") && i + 1 < argc) { root = argv[++i]; }
        else if (!strcmp(argv[i], "--raw") && i + 1 < argc) { rawPath = argv[++i]; }
        else { jobPath = argv[i]; }
    }
    g_rtcDevice = rtcNewDevice(NULL);
    g_rtcScene = rtcNewScene(g_rtcDevice);
    if (chdir(root.c_str()) != 0) { perror("chdir"); return 1; }
    std::ifstream jsonJob(jobPath);
    if (!jsonJob) { fprintf(stderr, "cannot open job %s\n", jobPath.c_str()); return 1; }
    nlohmann::json jobJson = nlohmann::json::parse(jsonJob);
    jsonJob.clear(); jsonJob.seekg(0);
    g_job = new Job(jsonJob);
    g_job->init();
    const int width = g_job->width(), height = g_job->height();
    Image image(width, height);

    const auto t0 = std::chrono::steady_clock::now();
    std::ifstream jsonScene(g_job->scene());
    if (!jsonScene) { fprintf(stderr, "cannot open scene %s\n", g_job->scene().c_str()); return 1; }
    Scene scene = parseScene(jsonScene); // the reference's own parser; its rtc* calls land in the shim
    bindScene(scene);
    const auto t1 = std::chrono::steady_clock::now();

    const std::string name = jobJson["integrator"].get<std::string>();
    if (name != "PathTracer" && name != "VolumePathTracer") { fprintf(stderr, "pathed_ref_cuda binds PathTracer and VolumePathTracer (src/job.cpp:69-72)\n"); return 1; }
    CudaPathTracer integrator(name == "VolumePathTracer" ? PTC_INTEGRATOR_VOLUME_PATH_TRACER : PTC_INTEGRATOR_PATH_TRACER,
                              jobJson.value("startBounce", 0), jobJson.value("lastBounce", 10));
    bool quit = false;
    integrator.run(image, scene, [](RenderStatus) {}, &quit);
    image.save(g_job->outputName());
    const auto t2 = std::chrono::steady_clock::now();
    if (!rawPath.empty()) {
        FILE *f = fopen(rawPath.c_str(), "wb");
        if (f) { fwrite(image.m_raw.data(), sizeof(float), image.m_raw.size(), f); fclose(f); }
    }
    uint32_t lights = 0;
    ptc_num_lights(g_ptc, &lights);
    printf("REF_CUDA_RESULT {\"width\": %d, \"height\": %d, \"spp\": %d, \"lights\": %u, \"scene_build_s\": %.4f, \"render_wall_s\": %.4f}\n", width, height,
           g_job->spp(), lights, std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t2 - t1).count());
    ptc_destroy(g_ptc);
    return 0;
}
