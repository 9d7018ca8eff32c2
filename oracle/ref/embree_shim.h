// Test infrastructure: what the Embree-API shim (embree_shim.cpp) recorded while the UNMODIFIED reference parsed a scene.
// The shim implements the part of the Embree 3 C API the reference calls (ext/embree/include/embree3/rtcore_*.h) and keeps the
// scene graph instead of building a BVH; cuda_main.cpp replays it into libpathed_cuda through include/pathed_cuda.h.
#pragma once

#include <embree3/rtcore.h>

#include <cstddef>
#include <vector>

struct ShimScene;

struct ShimGeometry {
    int references = 1;
    RTCGeometryType type;
    std::vector<unsigned char> vertices, indices, attributes[2]; // as the caller filled them (rtcSetNewGeometryBuffer)
    size_t vertexStride = 0, vertexCount = 0, indexStride = 0, indexCount = 0, attributeStride[2] = {0, 0}, attributeCount[2] = {0, 0};
    float transform[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; // column-major 4x4 of an instance (rtcSetGeometryTransform)
    ShimScene *instanced = nullptr;                                         // rtcSetGeometryInstancedScene
    bool hasFilter = false;
};

struct ShimScene {
    std::vector<ShimGeometry *> geometries; // attach order: index = geometry id
    bool committed = false;
};
