// Probe library over the UNMODIFIED reference (test infrastructure; built into
// oracle/_ref/libpathed_ref_probe.so by build_ref.sh, driven from Python by ctypes).
//
// Every entry point calls the reference's own classes and functions on flat arrays:
//   ref_intersect / ref_occluded      -> Scene::testIntersect / testOcclusion (src/scene.cpp:91-223, 355-381)
//                                        plus a raw rtcIntersect1 for geomID/primID/u/v/Ng
//   ref_bsdf_eval / ref_bsdf_sample   -> Material::f / Material::sample (include/material.h:18-46)
//   ref_triangle_* / ref_sphere_*     -> Triangle / Sphere ::sample, ::pdf (src/triangle.cpp, src/sphere.cpp)
//   ref_env_*                         -> EnvironmentLight::sample / emit / emitPDF (src/environment_light.cpp)
//   ref_scene_*                       -> Scene::sampleDirectLights / lightsPDF / environmentL / environmentPDF
//   ref_camera_rays                   -> Camera::generateRay(float,float) (src/camera.cpp:32-47)
//   ref_radiance                      -> SampleIntegrator::samplePixel's body + PathTracer::L
//                                        (src/sample_integrator.cpp:18-59, src/path_tracer.cpp:19-77)
// Random numbers are replayed from caller-supplied arrays (see random_generator.h here).
#include "area_light.h"
#include "beckmann.h"
#include "camera.h"
#include "checkerboard.h"
#include "environment_light.h"
#include "ggx.h"
#include "glass.h"
#include "globals.h"
#include "job.h"
#include "lambertian.h"
#include "microfacet.h"
#include "mirror.h"
#include "oren_nayar.h"
#include "path_tracer.h"
#include "plastic.h"
#include "texture.h"
#include "ray.h"
#include "scene.h"
#include "scene_parser.h"
#include "sphere.h"
#include "triangle.h"
#include "volume_helper.h"
#include "volume_path_tracer.h"

#include <embree3/rtcore.h>
#define STB_IMAGE_WRITE_IMPLEMENTATION
#include "stb_image_write.h"
#define STB_IMAGE_IMPLEMENTATION
#include "stb_image.h"
#define TINYEXR_IMPLEMENTATION
#include "tinyexr.h"

#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <string>

Job *g_job;
RTCDevice g_rtcDevice;
RTCScene g_rtcScene;

static Scene *s_scene = nullptr;
static std::shared_ptr<PathTracer> s_pathTracer;
static std::shared_ptr<VolumePathTracer> s_volumeTracer;
static int s_integrator = 0; // 0 PathTracer, 1 VolumePathTracer (ref_set_integrator)

static inline Vector3 vec(const float *p) { return Vector3(p[0], p[1], p[2]); }
static inline Point3 pnt(const float *p) { return Point3(p[0], p[1], p[2]); }
static inline void put(float *o, const Vector3 &v) { o[0] = v.x(); o[1] = v.y(); o[2] = v.z(); }
static inline void put(float *o, const Point3 &v) { o[0] = v.x(); o[1] = v.y(); o[2] = v.z(); }
static inline void put(float *o, const Color &c) { o[0] = c.r(); o[1] = c.g(); o[2] = c.b(); }

// exposes PathTracer's protected samplePixel-level logic without touching the reference
struct ProbeTracer : public PathTracer {
    using PathTracer::PathTracer;
};

extern "C" {

int ref_init(const char *root, const char *sceneJson, int width, int height, int startBounce, int lastBounce)
{
    if (s_scene) { return -1; } // one scene per process (the reference uses process globals)
    if (chdir(root) != 0) { return -2; }

    g_rtcDevice = rtcNewDevice(NULL);
    g_rtcScene = rtcNewScene(g_rtcDevice);

    char jobPath[] = "/tmp/pathed_probe_job_XXXXXX";
    int fd = mkstemp(jobPath);
    if (fd < 0) { return -3; }
    std::ostringstream job;
    job << "{\"spp\": 1, \"integrator\": \"PathTracer\", \"scene\": \"" << sceneJson << "\", "
        << "\"startBounce\": " << startBounce << ", \"lastBounce\": " << lastBounce << ", "
        << "\"output_directory\": \"/tmp/pathed_probe_out\", \"showUI\": false, \"force\": true, "
        << "\"width\": " << width << ", \"height\": " << height << ", \"output_name\": \"probe\"}";
    const std::string text = job.str();
    if (write(fd, text.c_str(), text.size()) < 0) { return -3; }
    close(fd);

    std::ifstream jobFile(jobPath);
    g_job = new Job(jobFile);
    unlink(jobPath);

    std::ifstream sceneFile(sceneJson);
    if (!sceneFile) { return -4; }
    s_scene = new Scene(parseScene(sceneFile));
    s_pathTracer = std::make_shared<PathTracer>(g_job->bounceController());
    s_volumeTracer = std::make_shared<VolumePathTracer>(g_job->bounceController());
    return 0;
}

// which L() ref_radiance calls: Job::integrator (src/job.cpp:66-75)
void ref_set_integrator(int integrator) { s_integrator = integrator; }

int ref_num_lights() { return s_scene ? (int)s_scene->lights().size() : -1; }

int ref_num_geometries() { return s_scene ? (int)s_scene->getSurfaces().size() : -1; }

int ref_geometry_size(int geomID) { return (int)s_scene->getSurfaces().at(geomID).size(); }

// emission of the surface registered at (geomID, primID): pins the light table order
void ref_surface_emit(int geomID, int primID, float *rgb)
{
    put(rgb, s_scene->getSurfaces().at(geomID).at(primID)->getMaterial()->emit());
}

void ref_camera_rays(int n, const float *rowCol, float *rays)
{
    for (int i = 0; i < n; i++) {
        const Ray ray = s_scene->getCamera()->generateRay(rowCol[2 * i], rowCol[2 * i + 1]);
        put(rays + 6 * i, ray.origin());
        put(rays + 6 * i + 3, ray.direction());
    }
}

// raw Embree record, traced exactly as Scene::testIntersect sets the ray up
void ref_intersect_raw(
    int n, const float *rays,
    float *t, unsigned *geomID, unsigned *primID, float *uv, float *Ng
) {
    #pragma omp parallel for
    for (int i = 0; i < n; i++) {
        RTCRayHit rh;
        rh.ray.org_x = rays[6 * i + 0]; rh.ray.org_y = rays[6 * i + 1]; rh.ray.org_z = rays[6 * i + 2];
        rh.ray.dir_x = rays[6 * i + 3]; rh.ray.dir_y = rays[6 * i + 4]; rh.ray.dir_z = rays[6 * i + 5];
        rh.ray.tnear = 1e-3f;
        rh.ray.tfar = 1e5f;
        rh.ray.flags = 0;
        rh.ray.time = 0.f;
        rh.ray.mask = -1;
        rh.hit.geomID = RTC_INVALID_GEOMETRY_ID;
        rh.hit.instID[0] = RTC_INVALID_GEOMETRY_ID;
        // same context type the reference's filter callback expects (src/scene.cpp:42-49): pass-throughs are intersected
        CustomRTCIntersectContext context;
        rtcInitIntersectContext(&context.context);
        context.rtcManagerPtr = nullptr;
        context.shouldIntersectPassthroughs = true;
        rtcIntersect1(g_rtcScene, &context.context, &rh);
        t[i] = rh.ray.tfar;
        geomID[i] = rh.hit.geomID;
        primID[i] = rh.hit.primID;
        uv[2 * i] = rh.hit.u; uv[2 * i + 1] = rh.hit.v;
        Ng[3 * i] = rh.hit.Ng_x; Ng[3 * i + 1] = rh.hit.Ng_y; Ng[3 * i + 2] = rh.hit.Ng_z;
    }
}

// RTCHit::instID of the same query (hierarchical instancing, RTC_MAX_INSTANCE_LEVEL_COUNT = 2)
void ref_intersect_inst(int n, const float *rays, unsigned *instID /*2n*/)
{
    #pragma omp parallel for
    for (int i = 0; i < n; i++) {
        RTCRayHit rh;
        rh.ray.org_x = rays[6 * i + 0]; rh.ray.org_y = rays[6 * i + 1]; rh.ray.org_z = rays[6 * i + 2];
        rh.ray.dir_x = rays[6 * i + 3]; rh.ray.dir_y = rays[6 * i + 4]; rh.ray.dir_z = rays[6 * i + 5];
        rh.ray.tnear = 1e-3f; rh.ray.tfar = 1e5f; rh.ray.flags = 0; rh.ray.time = 0.f; rh.ray.mask = -1;
        rh.hit.geomID = RTC_INVALID_GEOMETRY_ID;
        for (int l = 0; l < RTC_MAX_INSTANCE_LEVEL_COUNT; l++) { rh.hit.instID[l] = RTC_INVALID_GEOMETRY_ID; }
        CustomRTCIntersectContext context;
        rtcInitIntersectContext(&context.context);
        context.rtcManagerPtr = nullptr;
        context.shouldIntersectPassthroughs = true;
        rtcIntersect1(g_rtcScene, &context.context, &rh);
        const bool hit = rh.hit.geomID != RTC_INVALID_GEOMETRY_ID;
        instID[2 * i] = hit ? rh.hit.instID[0] : RTC_INVALID_GEOMETRY_ID;
        instID[2 * i + 1] = hit && RTC_MAX_INSTANCE_LEVEL_COUNT > 1 ? rh.hit.instID[1] : RTC_INVALID_GEOMETRY_ID;
    }
}

// processed Intersection as the integrator sees it
void ref_intersect(
    int n, const float *rays,
    int *hit, float *t, float *point, float *normal, float *shadingNormal, float *texUV,
    float *emit, int *isDelta
) {
    #pragma omp parallel for
    for (int i = 0; i < n; i++) {
        const Ray ray(pnt(rays + 6 * i), vec(rays + 6 * i + 3));
        const Intersection isect = s_scene->testIntersect(ray);
        hit[i] = isect.hit ? 1 : 0;
        t[i] = isect.t;
        put(point + 3 * i, isect.point);
        put(normal + 3 * i, isect.normal);
        put(shadingNormal + 3 * i, isect.shadingNormal);
        texUV[2 * i] = isect.uv.u; texUV[2 * i + 1] = isect.uv.v;
        if (isect.hit) {
            put(emit + 3 * i, isect.material->emit());
            isDelta[i] = isect.material->isDelta() ? 1 : 0;
        } else {
            emit[3 * i] = emit[3 * i + 1] = emit[3 * i + 2] = 0.f;
            isDelta[i] = 0;
        }
    }
}

void ref_occluded(int n, const float *rays, const float *maxT, unsigned char *occluded)
{
    #pragma omp parallel for
    for (int i = 0; i < n; i++) {
        const Ray ray(pnt(rays + 6 * i), vec(rays + 6 * i + 3));
        occluded[i] = s_scene->testOcclusion(ray, maxT[i]) ? 1 : 0;
    }
}

// Scene::testVolumetricOcclusion (src/scene.cpp:383-424): events of unoccluded rays, up to `maxEvents` stored per ray
void ref_volumetric_occluded(int n, const float *rays, const float *maxT, unsigned char *occluded, int *nEvents, float *eventT, int maxEvents)
{
    #pragma omp parallel for
    for (int i = 0; i < n; i++) {
        const Ray ray(pnt(rays + 6 * i), vec(rays + 6 * i + 3));
        const OcclusionResult result = s_scene->testVolumetricOcclusion(ray, maxT[i]);
        occluded[i] = result.isOccluded ? 1 : 0;
        nEvents[i] = result.isOccluded ? 0 : (int)result.volumeEvents.size();
        for (int e = 0; e < maxEvents; e++) {
            eventT[(size_t)i * maxEvents + e] = (!result.isOccluded && e < (int)result.volumeEvents.size()) ? result.volumeEvents[e].t : 0.f;
        }
    }
}

// Scene::testVolumetricIntersect (src/scene.cpp:225-353)
void ref_volumetric_intersect(int n, const float *rays, int *hit, float *t, float *point, float *emit, int *nEvents, float *eventT, int maxEvents)
{
    #pragma omp parallel for
    for (int i = 0; i < n; i++) {
        const Ray ray(pnt(rays + 6 * i), vec(rays + 6 * i + 3));
        const IntersectionResult result = s_scene->testVolumetricIntersect(ray);
        const Intersection &isect = result.intersection;
        hit[i] = isect.hit ? 1 : 0;
        t[i] = isect.t;
        put(point + 3 * i, isect.point);
        if (isect.hit) { put(emit + 3 * i, isect.material->emit()); } else { emit[3 * i] = emit[3 * i + 1] = emit[3 * i + 2] = 0.f; }
        nEvents[i] = (int)result.volumeEvents.size();
        for (int e = 0; e < maxEvents; e++) { eventT[(size_t)i * maxEvents + e] = e < (int)result.volumeEvents.size() ? result.volumeEvents[e].t : 0.f; }
    }
}

// ------------------------------------------------------------------ materials
// type: 0 lambertian, 1 oren-nayar, 2 mirror, 3 glass, 4 microfacet, 5 plastic
// p[0..2] diffuse, p[3..5] emit, p[6] sigma | ior, p[7] distribution (0 beckmann, 1 ggx), p[8] alpha,
// p[9] albedo kind (0 constant, 1 checkerboard), p[10..12] on, p[13..15] off, p[16..17] resolution u,v
void *ref_material_new(int type, const float *p)
{
    auto dist = [&]() -> std::unique_ptr<MicrofacetDistribution> {
        if (p[7] == 0.f) { return std::make_unique<Beckmann>(p[8]); }
        return std::make_unique<GGX>(p[8]);
    };
    const Color diffuse(p[0], p[1], p[2]);
    const Color emit(p[3], p[4], p[5]);
    Material *m = nullptr;
    switch (type) {
    case 0:
        if (p[9] == 1.f) {
            auto checker = std::make_shared<Checkerboard>(
                Color(p[10], p[11], p[12]), Color(p[13], p[14], p[15]), UV{p[16], p[17]});
            m = new Lambertian(checker, emit);
        } else {
            m = new Lambertian(diffuse, emit);
        }
        break;
    case 1: m = new OrenNayar(diffuse, p[6]); break;
    case 2: m = new Mirror(); break;
    case 3: m = new Glass(p[6]); break;
    case 4: m = new Microfacet(dist()); break;
    case 5: m = new Plastic(diffuse, dist()); break;
    }
    return m;
}

// Texture::load's decode: stbi_load(path, &w, &h, &channels, 3) of the reference's vendored stb_image (src/texture.cpp:16-21).
// Returns 0 and the size; fills rgb when it has room.
int ref_load_image(const char *path, unsigned char *rgb, int capacity, int *width, int *height)
{
    int channels = 0;
    unsigned char *data = stbi_load(path, width, height, &channels, 3);
    if (!data) { return -1; }
    if (rgb && capacity >= *width * *height * 3) { memcpy(rgb, data, (size_t)*width * *height * 3); }
    stbi_image_free(data);
    return 0;
}

// N1: Lambertian / Plastic whose albedo is an image texture, built the way parseMaterial does (src/scene_parser.cpp:625-647)
void *ref_material_new_textured(int type, const float *p, const char *texturePath)
{
    auto texture = std::make_shared<Texture>(texturePath);
    texture->load();
    if (type == 0) { return new Lambertian(texture, Color(p[3], p[4], p[5])); }
    std::unique_ptr<MicrofacetDistribution> dist;
    if (p[7] == 0.f) { dist = std::make_unique<Beckmann>(p[8]); } else { dist = std::make_unique<GGX>(p[8]); }
    return new Plastic(std::make_unique<Lambertian>(texture, Color(0.f)), std::move(dist));
}

static Intersection makeIsect(const float *wo, const float *ng, const float *ns, const float *uv, Material *m)
{
    return Intersection(
        true, 1.f, Point3(0.f, 0.f, 0.f), vec(wo), vec(ng), vec(ns), UV{uv[0], uv[1]}, m, nullptr);
}

void ref_bsdf_eval(
    void *material, int n,
    const float *wo, const float *ng, const float *ns, const float *uv, const float *wi,
    float *f, float *pdf
) {
    Material *m = (Material *)material;
    for (int i = 0; i < n; i++) {
        const Intersection isect = makeIsect(wo + 3 * i, ng + 3 * i, ns + 3 * i, uv + 2 * i, m);
        float p = 0.f;
        const Color value = m->f(isect, vec(wi + 3 * i), &p);
        put(f + 3 * i, value);
        pdf[i] = p;
    }
}

void ref_bsdf_sample(
    void *material, int n,
    const float *wo, const float *ng, const float *ns, const float *uv, const float *xi /*3n*/,
    float *wi, float *pdf, float *throughput, int *consumed
) {
    Material *m = (Material *)material;
    RandomGenerator random;
    for (int i = 0; i < n; i++) {
        const Intersection isect = makeIsect(wo + 3 * i, ng + 3 * i, ns + 3 * i, uv + 2 * i, m);
        RandomGenerator::beginReplay(xi + 3 * i, 3);
        const BSDFSample s = m->sample(isect, random);
        consumed[i] = RandomGenerator::endReplay();
        put(wi + 3 * i, s.wiWorld);
        pdf[i] = s.pdf;
        put(throughput + 3 * i, s.throughput);
    }
}

// tangent frame: rows of worldToTangent (x, n, z axes)
void ref_tangent_frame(int n, const float *ns, const float *wo, float *frame /*9n*/)
{
    for (int i = 0; i < n; i++) {
        const Transform toWorld = normalToWorldSpace(vec(ns + 3 * i), vec(wo + 3 * i));
        put(frame + 9 * i + 0, toWorld.apply(Vector3(1.f, 0.f, 0.f)));
        put(frame + 9 * i + 3, toWorld.apply(Vector3(0.f, 1.f, 0.f)));
        put(frame + 9 * i + 6, toWorld.apply(Vector3(0.f, 0.f, 1.f)));
    }
}

// ------------------------------------------------------------------ shapes
void ref_triangle_sample(
    const float *p9, int n, const float *ref, const float *xi /*2n*/,
    float *point, float *normal, float *invPDF, int *measure
) {
    const Triangle tri(pnt(p9), pnt(p9 + 3), pnt(p9 + 6));
    RandomGenerator random;
    for (int i = 0; i < n; i++) {
        RandomGenerator::beginReplay(xi + 2 * i, 2);
        const SurfaceSample s = ((const Shape &)tri).sample(pnt(ref + 3 * i), random);
        RandomGenerator::endReplay();
        put(point + 3 * i, s.point); put(normal + 3 * i, s.normal);
        invPDF[i] = s.invPDF; measure[i] = s.measure == Measure::SolidAngle ? 0 : 1;
    }
}

void ref_triangle_pdf(const float *p9, int n, const float *point, const float *ref, float *pdf)
{
    const Triangle tri(pnt(p9), pnt(p9 + 3), pnt(p9 + 6));
    for (int i = 0; i < n; i++) {
        pdf[i] = tri.pdf(pnt(point + 3 * i), pnt(ref + 3 * i), Measure::SolidAngle);
    }
}

void ref_sphere_sample(
    const float *centerRadius, int n, const float *ref, const float *xi /*2n*/,
    float *point, float *normal, float *invPDF, int *measure
) {
    const Sphere sphere(pnt(centerRadius), centerRadius[3]);
    RandomGenerator random;
    for (int i = 0; i < n; i++) {
        RandomGenerator::beginReplay(xi + 2 * i, 2);
        const SurfaceSample s = sphere.sample(pnt(ref + 3 * i), random);
        RandomGenerator::endReplay();
        put(point + 3 * i, s.point); put(normal + 3 * i, s.normal);
        invPDF[i] = s.invPDF; measure[i] = s.measure == Measure::SolidAngle ? 0 : 1;
    }
}

void ref_sphere_pdf(const float *centerRadius, int n, const float *point, const float *ref, float *pdf)
{
    const Sphere sphere(pnt(centerRadius), centerRadius[3]);
    for (int i = 0; i < n; i++) {
        const float d = (pnt(centerRadius) - pnt(ref + 3 * i)).toVector().length();
        // the inside branch prints and throws in the reference (src/sphere.cpp:137-140); skip it
        pdf[i] = d <= centerRadius[3] ? -1.f : sphere.pdf(pnt(point + 3 * i), pnt(ref + 3 * i), Measure::SolidAngle);
    }
}

// ------------------------------------------------------------------ environment light
void *ref_env_new(const char *exrPath, float scale, const float *mapToWorld16, const float *worldToMap16)
{
    float m[4][4], inv[4][4];
    memcpy(m, mapToWorld16, sizeof(m));
    memcpy(inv, worldToMap16, sizeof(inv));
    return new EnvironmentLight(exrPath, scale, Transform(m, inv));
}

void ref_env_sample(
    void *env, int n, const float *ref, const float *xi /*2n*/,
    float *point, float *normal, float *invPDF
) {
    const EnvironmentLight *light = (const EnvironmentLight *)env;
    RandomGenerator random;
    for (int i = 0; i < n; i++) {
        RandomGenerator::beginReplay(xi + 2 * i, 2);
        const SurfaceSample s = light->sample(pnt(ref + 3 * i), random);
        RandomGenerator::endReplay();
        put(point + 3 * i, s.point); put(normal + 3 * i, s.normal);
        invPDF[i] = s.invPDF;
    }
}

// radiance arriving from direction `dir` (= Scene::environmentL(dir), src/scene.cpp:486-492)
void ref_env_emit(void *env, int n, const float *dir, float *rgb)
{
    const EnvironmentLight *light = (const EnvironmentLight *)env;
    for (int i = 0; i < n; i++) {
        put(rgb + 3 * i, light->emit(-vec(dir + 3 * i)));
    }
}

void ref_env_pdf(void *env, int n, const float *dir, float *pdf)
{
    const EnvironmentLight *light = (const EnvironmentLight *)env;
    for (int i = 0; i < n; i++) {
        pdf[i] = light->emitPDF(vec(dir + 3 * i), Measure::SolidAngle);
    }
}

// ------------------------------------------------------------------ scene-level light queries
void ref_scene_sample_direct_lights(
    int n, const float *ref, const float *xi /*3n*/,
    float *point, float *normal, float *invPDF, int *measure, float *solidAnglePDF, float *emit
) {
    RandomGenerator random;
    for (int i = 0; i < n; i++) {
        RandomGenerator::beginReplay(xi + 3 * i, 3);
        const LightSample s = s_scene->sampleDirectLights(pnt(ref + 3 * i), random);
        RandomGenerator::endReplay();
        put(point + 3 * i, s.point); put(normal + 3 * i, s.normal);
        invPDF[i] = s.invPDF; measure[i] = s.measure == Measure::SolidAngle ? 0 : 1;
        solidAnglePDF[i] = s.solidAnglePDF(pnt(ref + 3 * i));
        const Vector3 lightWo = -((s.point - pnt(ref + 3 * i)).toVector().normalized());
        put(emit + 3 * i, s.light->emit(lightWo));
    }
}

// traces origin+dir; if it lands on an emitter returns Scene::lightsPDF, else -1;
// on a miss returns -2 - environmentPDF when an environment light exists
void ref_scene_lights_pdf(int n, const float *rays, float *pdf)
{
    for (int i = 0; i < n; i++) {
        const Ray ray(pnt(rays + 6 * i), vec(rays + 6 * i + 3));
        const Intersection isect = s_scene->testIntersect(ray);
        if (isect.hit && isect.isEmitter()) {
            pdf[i] = s_scene->lightsPDF(ray.origin(), isect, Measure::SolidAngle);
        } else if (!isect.hit && !s_scene->environmentL(ray.direction()).isBlack()) {
            pdf[i] = -2.f - s_scene->environmentPDF(ray.direction(), Measure::SolidAngle);
        } else {
            pdf[i] = -1.f;
        }
    }
}

void ref_scene_environment(int n, const float *dir, float *rgb)
{
    for (int i = 0; i < n; i++) { put(rgb + 3 * i, s_scene->environmentL(vec(dir + 3 * i))); }
}

// ------------------------------------------------------------------ whole paths
// One radiance sample per primary ray, with the random stream replayed from xi[i*stride ...]:
// the body of SampleIntegrator::samplePixel (src/sample_integrator.cpp:18-59, container branch included)
// followed by PathTracer::L or VolumePathTracer::L (ref_set_integrator).
void ref_radiance(int n, const float *rays, const float *xi, int stride, float *rgb, int *consumed)
{
    #pragma omp parallel for
    for (int i = 0; i < n; i++) {
        RandomGenerator random;
        const Ray ray(pnt(rays + 6 * i), vec(rays + 6 * i + 3));
        Color color(0.f);
        RandomGenerator::beginReplay(xi + (size_t)i * stride, stride);
        const Intersection isect = s_scene->testIntersect(ray);
        if (isect.hit) {
            Sample sample;
            if (g_job->bounceController().checkCounts(0)) {
                const Color emit = isect.material->emit();
                if (!emit.isBlack() && !IntersectionHelper::checkBacksideIntersection(isect)) {
                    color += emit;
                }
                if (isect.material->isContainer()) { // src/sample_integrator.cpp:35-51
                    const IntersectionResult volumetricResult = s_scene->testVolumetricIntersect(ray);
                    const Intersection &volumetricIntersection = volumetricResult.intersection;
                    const Color transmittance = VolumeHelper::rayTransmission(ray, volumetricResult.volumeEvents, nullptr);
                    if (volumetricIntersection.hit) { color += volumetricIntersection.material->emit() * transmittance; }
                    else { color += s_scene->environmentL(ray.direction()) * transmittance; }
                }
            }
            if (s_integrator == 1) { color += s_volumeTracer->L(isect, *s_scene, random, 0, sample); }
            else { color += s_pathTracer->L(isect, *s_scene, random, 0, sample); }
        } else {
            color += s_scene->environmentL(ray.direction());
        }
        consumed[i] = RandomGenerator::endReplay();
        put(rgb + 3 * i, color);
    }
}

} // extern "C"
