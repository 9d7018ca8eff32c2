// Flat scene description: what Pathed's parsers produce, captured at the rtc* sink
// (/root/reference/src/geometry_parser.cpp:5-97, src/quad.cpp:34-150, src/sphere.cpp:20-47) and at the
// Material / Camera / EnvironmentLight constructors, in registration order.  It is fed to the CUDA
// library through the C ABI (include/pathed_cuda.h) by feedScene().
#pragma once

#include "../../include/pathed_cuda.h"

#include <string>
#include <vector>

namespace pathed {

struct GeometryDesc {
    bool isSphere = false;
    // triangle mesh (one Embree geometry): per-vertex position/normal/uv, per-face indices + material
    std::vector<float> positions, normals, uvs;
    std::vector<uint32_t> indices, materialOfTri;
    // sphere
    float centerRadius[4] = {0.f, 0.f, 0.f, 0.f};
    uint32_t sphereMaterial = 0;
    // "internal_medium" of the model (src/scene_parser.cpp:324-343, :370-381, :503-514): index into SceneDescription::media, -1 none
    int internalMedium = -1;
    // "instanced" model (src/scene_parser.cpp:449-492): a placement of SceneDescription::instanceScenes[instanceScene] under a
    // column-major 4x4 local-to-world matrix; it takes a geometry id like any other geometry (rtcAttachGeometry)
    bool isInstance = false;
    uint32_t instanceScene = 0;
    float instanceTransform[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
};

// "instance" model (parseInstance, src/scene_parser.cpp:231-249): a named scene of its own (rtcNewScene) that "instanced" models place
struct InstanceSceneDesc {
    std::string name;
    std::vector<GeometryDesc> geometries; // index == geometry id inside the instance scene
};

struct MediumDesc { // HomogeneousMedium (include/homogeneous_medium.h), parseMedia src/scene_parser.cpp:202-229
    std::string name;
    float sigmaT[3], sigmaS[3];
};

struct CameraDesc {
    float origin[3], target[3], up[3];
    float verticalFov; // radians
    int width, height;
    bool flipHandedness;
};

struct EnvironmentDesc {
    bool present = false;
    std::string filename;
    std::vector<float> rgba;
    int width = 0, height = 0;
    float scale = 1.f;
    float mapToWorld[16], worldToMap[16];
};

struct TextureDesc { // Texture (include/texture.h): decoded at parse time like Texture::load
    std::string filename;
    std::vector<uint8_t> rgb;
    int width = 0, height = 0;
};

struct SceneDescription {
    std::vector<TextureDesc> textures; // ptc_material_desc::texture indexes this list
    std::vector<ptc_material_desc> materials;
    std::vector<MediumDesc> media;
    std::vector<GeometryDesc> geometries; // index == Embree geomID
    std::vector<InstanceSceneDesc> instanceScenes; // in order of completion: a scene only places scenes that come before it
    CameraDesc camera;
    EnvironmentDesc environment;
};

// Table of the C-ABI scene-description entry points; lets the same feeder drive any library that
// exports them (the product's libpathed_cuda.so; in tests also the CPU checker).
struct SceneSink {
    void *ctx;
    int (*add_material)(void *, const ptc_material_desc *, uint32_t *);
    int (*add_triangle_mesh)(void *, const float *, const float *, const float *, uint32_t, const uint32_t *,
                             const uint32_t *, uint32_t, uint32_t *);
    int (*add_sphere)(void *, const float *, uint32_t, uint32_t *);
    int (*set_environment)(void *, const float *, int, int, float, const float *, const float *);
    int (*set_camera)(void *, const float *, const float *, const float *, float, int, int, int);
    int (*commit)(void *);
    int (*add_texture)(void *, const uint8_t *, int, int, uint32_t *);
    int (*add_medium)(void *, const float *, const float *, uint32_t *);
    int (*set_internal_medium)(void *, uint32_t, uint32_t);
    // SURVEY 8(f) N4 (may be null for a sink that does not take instances: feeding a scene that has some then fails)
    int (*begin_instance)(void *, uint32_t *);
    int (*end_instance)(void *);
    int (*add_instance)(void *, uint32_t, const float *, uint32_t *);
};

// returns the first non-zero status of the sink, 0 on success
int feedScene(const SceneDescription &scene, const SceneSink &sink);

} // namespace pathed
