// C entry points over the host layer, for Python (ctypes) tests and bench.py.
// The C++ classes (Job, Image, Scene, CudaPathTracer) are the drop-in surface for C++ callers;
// these wrappers only expose the parsers so scripts can load a scene file into any library that
// exports the ptc_* scene-description calls.
#include "exr_io.hpp"
#include "scene_description.hpp"
#include "scene_parser.hpp"

#include <cstdio>
#include <cstring>
#include <exception>

using namespace pathed;

extern "C" {

int pth_scene_load(const char *sceneJson, const char *rootDirectory, int width, int height, void **out, char *err, int errLen)
{
    try {
        *out = new SceneDescription(parseScene(sceneJson, rootDirectory ? rootDirectory : "", width, height));
        return 0;
    } catch (const std::exception &e) {
        if (err && errLen > 0) { snprintf(err, (size_t)errLen, "%s", e.what()); }
        return -1;
    }
}

void pth_scene_free(void *scene) { delete (SceneDescription *)scene; }

int pth_scene_feed(void *scene, const SceneSink *sink) { return feedScene(*(SceneDescription *)scene, *sink); }

void pth_scene_counts(void *scene, uint32_t *geometries, uint32_t *triangles, uint32_t *spheres, uint32_t *materials)
{
    const SceneDescription &s = *(SceneDescription *)scene;
    uint32_t tris = 0, sph = 0;
    for (const GeometryDesc &g : s.geometries) { if (g.isSphere) { sph++; } else { tris += (uint32_t)g.materialOfTri.size(); } }
    *geometries = (uint32_t)s.geometries.size(); *triangles = tris; *spheres = sph; *materials = (uint32_t)s.materials.size();
}

// geometry access for tests (vertex-level fixtures)
int pth_scene_geometry(void *scene, uint32_t geom, const float **positions, uint32_t *nVertices, const uint32_t **indices,
                       const uint32_t **materialOfTri, uint32_t *nTriangles)
{
    const SceneDescription &s = *(SceneDescription *)scene;
    if (geom >= s.geometries.size() || s.geometries[geom].isSphere) { return -1; }
    const GeometryDesc &g = s.geometries[geom];
    *positions = g.positions.data(); *nVertices = (uint32_t)(g.positions.size() / 3);
    *indices = g.indices.data(); *materialOfTri = g.materialOfTri.data(); *nTriangles = (uint32_t)g.materialOfTri.size();
    return 0;
}

int pth_scene_material(void *scene, uint32_t id, ptc_material_desc *out)
{
    const SceneDescription &s = *(SceneDescription *)scene;
    if (id >= s.materials.size()) { return -1; }
    *out = s.materials[id];
    return 0;
}

int pth_exr_write_rgb_f32(const char *path, int width, int height, const float *rgb /*interleaved, top row first*/)
{
    try {
        std::vector<float> planes[3];
        for (int c = 0; c < 3; c++) { planes[c].resize((size_t)width * height); }
        for (size_t i = 0; i < (size_t)width * height; i++) { for (int c = 0; c < 3; c++) { planes[c][i] = rgb[3 * i + c]; } }
        saveEXR(path, width, height, {"B", "G", "R"}, {planes[2].data(), planes[1].data(), planes[0].data()}, false);
        return 0;
    } catch (const std::exception &) { return -1; }
}

int pth_exr_read_rgba(const char *path, float *rgba, int capacityPixels, int *width, int *height)
{
    try {
        std::vector<float> data;
        loadEXR(path, data, *width, *height);
        if (rgba && (size_t)capacityPixels >= (size_t)*width * *height) { memcpy(rgba, data.data(), data.size() * sizeof(float)); }
        return 0;
    } catch (const std::exception &) { return -1; }
}

} // extern "C"
