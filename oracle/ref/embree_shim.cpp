// Test infrastructure: the Embree 3 C API as the reference calls it (17 entry points, `nm -u` of the reference's objects), recording the
// scene graph instead of building Embree's BVH.  Linked in place of Embree into oracle/_ref/pathed_ref_cuda, it is the seam at which
// the UNMODIFIED reference (its parsers, Job, Image, Integrator::run) hands its geometry to libpathed_cuda: SURVEY 8(b), INTEGRATION.md.
//   scene construction  rtcNewScene, rtcNewGeometry, rtcSetNewGeometryBuffer, rtcSetGeometryVertexAttributeCount, rtcSetGeometryTransform,
//                       rtcSetGeometryInstancedScene, rtcSetGeometryTimeStepCount, rtcCommitGeometry, rtcAttachGeometry,
//                       rtcReleaseGeometry, rtcCommitScene, rtcGetGeometry, filter registration  -> recorded (embree_shim.h)
//   ray queries         rtcIntersect1, rtcOccluded1, rtcInterpolate: the CUDA integrator never calls them (the whole of
//                       PathTracer::L runs on the device); a CPU integrator that did would stop here with a message
#include "embree_shim.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {
size_t formatSize(RTCFormat format)
{
    switch (format) {
    case RTC_FORMAT_FLOAT2: return 8;
    case RTC_FORMAT_FLOAT3: return 12;
    case RTC_FORMAT_FLOAT4: return 16;
    case RTC_FORMAT_UINT3: return 12;
    default: return 0;
    }
}
[[noreturn]] void unsupported(const char *what)
{
    fprintf(stderr, "embree shim: %s is not served here -- pathed_ref_cuda traces rays on the GPU through libpathed_cuda\n", what);
    abort();
}
} // namespace

extern "C" {

RTCDevice rtcNewDevice(const char *) { return (RTCDevice) new int(0); }
void rtcReleaseDevice(RTCDevice device) { delete (int *)device; }
RTCScene rtcNewScene(RTCDevice) { return (RTCScene) new ShimScene(); }
void rtcReleaseScene(RTCScene) {} // the recorded graph lives until the process ends
void rtcSetSceneBuildQuality(RTCScene, enum RTCBuildQuality) {}
void rtcCommitScene(RTCScene scene) { ((ShimScene *)scene)->committed = true; }

RTCGeometry rtcNewGeometry(RTCDevice, enum RTCGeometryType type)
{
    ShimGeometry *g = new ShimGeometry();
    g->type = type;
    return (RTCGeometry)g;
}

void *rtcSetNewGeometryBuffer(RTCGeometry geometry, enum RTCBufferType type, unsigned int slot, enum RTCFormat format, size_t byteStride, size_t itemCount)
{
    ShimGeometry *g = (ShimGeometry *)geometry;
    if (formatSize(format) == 0 || byteStride < formatSize(format)) { unsupported("this buffer format"); }
    std::vector<unsigned char> *store = nullptr;
    if (type == RTC_BUFFER_TYPE_VERTEX && slot == 0) { store = &g->vertices; g->vertexStride = byteStride; g->vertexCount = itemCount; }
    else if (type == RTC_BUFFER_TYPE_INDEX && slot == 0) { store = &g->indices; g->indexStride = byteStride; g->indexCount = itemCount; }
    else if (type == RTC_BUFFER_TYPE_VERTEX_ATTRIBUTE && slot < 2) { store = &g->attributes[slot]; g->attributeStride[slot] = byteStride; g->attributeCount[slot] = itemCount; }
    else { unsupported("this buffer type / slot"); }
    store->assign(byteStride * itemCount + 16, 0); // Embree pads buffers as well; the pointer stays valid until the geometry dies
    return store->data();
}

void rtcSetGeometryVertexAttributeCount(RTCGeometry, unsigned int count) { if (count > 2) { unsupported("more than two vertex attributes"); } }
void rtcSetGeometryTimeStepCount(RTCGeometry, unsigned int count) { if (count != 1) { unsupported("motion blur"); } }
void rtcSetGeometryInstancedScene(RTCGeometry geometry, RTCScene scene) { ((ShimGeometry *)geometry)->instanced = (ShimScene *)scene; }
void rtcSetGeometryTransform(RTCGeometry geometry, unsigned int timeStep, enum RTCFormat format, const void *xfm)
{
    if (timeStep != 0 || format != RTC_FORMAT_FLOAT4X4_COLUMN_MAJOR) { unsupported("this transform format"); }
    memcpy(((ShimGeometry *)geometry)->transform, xfm, 16 * sizeof(float));
}
void rtcCommitGeometry(RTCGeometry) {}
unsigned int rtcAttachGeometry(RTCScene scene, RTCGeometry geometry)
{
    ShimScene *s = (ShimScene *)scene;
    ShimGeometry *g = (ShimGeometry *)geometry;
    g->references++;
    s->geometries.push_back(g);
    return (unsigned int)s->geometries.size() - 1; // ids in attach order, as Embree hands them out for a scene nothing was detached from
}
void rtcReleaseGeometry(RTCGeometry geometry)
{
    ShimGeometry *g = (ShimGeometry *)geometry;
    if (--g->references == 0) { delete g; }
}
RTCGeometry rtcGetGeometry(RTCScene scene, unsigned int geomID)
{
    ShimScene *s = (ShimScene *)scene;
    return geomID < s->geometries.size() ? (RTCGeometry)s->geometries[geomID] : nullptr;
}
// Scene::registerOcclusionFilters (src/scene.cpp:42-84): libpathed_cuda derives the same filter from materials and media at ptc_commit
void rtcSetGeometryIntersectFilterFunction(RTCGeometry geometry, RTCFilterFunctionN) { ((ShimGeometry *)geometry)->hasFilter = true; }
void rtcSetGeometryOccludedFilterFunction(RTCGeometry geometry, RTCFilterFunctionN) { ((ShimGeometry *)geometry)->hasFilter = true; }

void rtcIntersect1(RTCScene, struct RTCIntersectContext *, struct RTCRayHit *) { unsupported("rtcIntersect1"); }
void rtcOccluded1(RTCScene, struct RTCIntersectContext *, struct RTCRay *) { unsupported("rtcOccluded1"); }
void rtcInterpolate(const struct RTCInterpolateArguments *) { unsupported("rtcInterpolate"); }

} // extern "C"
