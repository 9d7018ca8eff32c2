// Participating media on the device: the occlusion filter, the volumetric scene queries, HomogeneousMedium,
// VolumeHelper, DirectLightingHelper::Ld and VolumePathTracer::L.  Each function names the reference code it reproduces.
//
// Containers: a surface whose material is Passthrough (src/passthrough.cpp) AND that has an internal medium is never a hit for
// the "volumetric" queries and for Scene::testOcclusion; the Embree filter callback (src/scene.cpp:42-84) rejects it and records
// a volume event (t, medium) unless an event with the same t exists already.  Scene::testIntersect does hit containers.
#pragma once

#include "shading.cuh"

namespace ptc {

struct VolumeEvents {
    uint32_t count; // distinct events met; more than PTC_MAX_EVENTS are counted but not stored
    float t[PTC_MAX_EVENTS];
    int32_t medium[PTC_MAX_EVENTS];
};

__device__ __forceinline__ void eventAdd(VolumeEvents &ev, float t, int32_t medium) // src/scene.cpp:66-81
{
    const uint32_t stored = ev.count < PTC_MAX_EVENTS ? ev.count : PTC_MAX_EVENTS;
    for (uint32_t i = 0; i < stored; i++) { if (ev.t[i] == t) { return; } }
    if (ev.count < PTC_MAX_EVENTS) { ev.t[ev.count] = t; ev.medium[ev.count] = medium; }
    ev.count++;
}
__device__ __forceinline__ void eventsSort(VolumeEvents &ev) // std::sort by t, src/scene.cpp:337-343, :412-418
{
    const uint32_t stored = ev.count < PTC_MAX_EVENTS ? ev.count : PTC_MAX_EVENTS;
    for (uint32_t i = 1; i < stored; i++) {
        const float t = ev.t[i]; const int32_t m = ev.medium[i];
        uint32_t j = i;
        while (j > 0 && ev.t[j - 1] > t) { ev.t[j] = ev.t[j - 1]; ev.medium[j] = ev.medium[j - 1]; j--; }
        ev.t[j] = t; ev.medium[j] = m;
    }
}

// One triangle of the pending group with the filter applied before a hit is accepted (traversalTriangle + occlusionFilter)
template <bool COUNT>
__device__ __forceinline__ bool filteredTriangle(const DScene &s, TraversalState &st, VolumeEvents &ev, TraverseCounters &tc)
{
    const uint32_t bit = highestBit(st.tgroup.y);
    st.tgroup.y &= ~(1u << bit);
    const float4 *tri = s.bvh.triangles + (size_t)(st.tgroup.x + bit) * 3;
    const float4 a = loadNodeWord(tri), b = loadNodeWord(tri + 1), c = loadNodeWord(tri + 2);
    if (COUNT) { tc.tris++; }
    float T, U, V, absDen;
    if (!triangleTestRaw(a, b, c, st.ox, st.oy, st.oz, st.dx, st.dy, st.dz, st.tnear, st.hit.t, T, U, V, absDen)) { return false; }
    const float t = divIeee(T, absDen);
    const uint32_t prim = f2u(a.w);
    const int32_t medium = s.bvh.primEvent ? __ldg(s.bvh.primEvent + prim) : -1;
    if (medium >= 0) { eventAdd(ev, t, medium); return false; } // args->valid[0] = 0: the ray goes on, tfar unchanged
    if (!(st.found && t == st.hit.t && prim < st.hit.prim)) { st.hit.t = t; st.hit.u = U; st.hit.v = V; st.hitDen = absDen; st.hit.prim = prim; }
    st.found = true;
    return true;
}

// rtcIntersect1 / rtcOccluded1 with the filter registered on every geometry (src/rtc_manager.cpp:85-92) and
// shouldIntersectPassthroughs = false.  Closest hit: Embree calls the filter for every candidate closer than the closest
// accepted hit SO FAR, so which container surfaces behind the final hit leave an event depends on its traversal order; here
// the events kept are exactly the container surfaces in front of the final triangle hit (triangle meshes are traversed before
// the sphere points, as in Embree's per-type acceleration structures).  Any hit: every container surface in the interval when
// the ray is unoccluded -- order-independent.
// counters (optional): inner-node visits and triangle tests, the same quantities the wavefront traversal counts
// Kept out of line on purpose (like the other building blocks marked PTC_VOLUME_CALL below): the one-thread-per-path kernel
// reaches them from several places, and one copy of each keeps its code inside the instruction cache -- with everything inlined
// the kernel stalled on instruction fetch (no_instruction: 68 warps per issue) and spilled 470 GB per launch.
#define PTC_VOLUME_CALL __device__ __noinline__
template <bool ANY, bool COUNT = false>
PTC_VOLUME_CALL bool traverseFiltered(const DScene &s, V3 O, V3 D, float tnear, float tfar, RayHit &hit, VolumeEvents &ev, TraverseCounters *counters = nullptr)
{
    TraverseCounters tc = {0, 0};
    TraversalState st;
    traversalInit(st, O.x, O.y, O.z, D.x, D.y, D.z, tnear, tfar);
    ev.count = 0;
    bool stop = false;
    if (s.bvh.nNodes) {
        for (;;) {
            traversalNode<COUNT>(s.bvh, st, &tc);
            while (st.tgroup.y) { if (filteredTriangle<COUNT>(s, st, ev, tc) && ANY) { stop = true; break; } }
            if (stop || traversalPop(st)) { break; }
        }
    }
    if (!ANY && st.found) {
        st.hit.u = divIeee(st.hit.u, st.hitDen); st.hit.v = divIeee(st.hit.v, st.hitDen);
        uint32_t kept = 0; // events behind the closest triangle hit were candidates only in some traversal orders: dropped
        const uint32_t stored = ev.count < PTC_MAX_EVENTS ? ev.count : PTC_MAX_EVENTS;
        for (uint32_t i = 0; i < stored; i++) {
            if (ev.t[i] <= st.hit.t) { ev.t[kept] = ev.t[i]; ev.medium[kept] = ev.medium[i]; kept++; }
        }
        if (ev.count <= PTC_MAX_EVENTS) { ev.count = kept; }
    }
    if (!(ANY && st.found)) {
        for (uint32_t i = 0; i < s.bvh.nSpheres; i++) {
            float t, nx, ny, nz;
            if (sphereTest(loadNodeWord(s.bvh.spheres + i), st.ox, st.oy, st.oz, st.dx, st.dy, st.dz, st.tnear, st.hit.t, t, nx, ny, nz)) {
                const int32_t medium = s.bvh.sphereEvent ? __ldg(s.bvh.sphereEvent + i) : -1;
                if (medium >= 0) { eventAdd(ev, t, medium); continue; } // the far side of a rejected sphere is not tried (sphere_intersector.h:91-92)
                st.hit.t = t; st.hit.u = 0.f; st.hit.v = 0.f; st.hit.prim = PTC_SPHERE_FLAG | i;
                st.found = true;
                if (ANY) { break; }
            }
        }
    }
    eventsSort(ev);
    hit = st.hit;
    if (COUNT && counters) { counters->inner += tc.inner; counters->tris += tc.tris; }
    return st.found;
}

// Closest hit / any hit of the plain queries on a scene that may hold containers: Scene::testIntersect passes
// shouldIntersectPassthroughs = true (containers are ordinary hits), Scene::testOcclusion passes false (src/scene.cpp:369-370)
template <bool COUNT = false>
PTC_VOLUME_CALL bool sceneIntersect(const DScene &s, V3 O, V3 D, RayHit &h, TraverseCounters *counters = nullptr)
{
    TraverseCounters tc = {0, 0};
    const bool found = traverseBVH<false, COUNT>(s.bvh, O.x, O.y, O.z, D.x, D.y, D.z, PTC_TNEAR, PTC_TFAR, h, &tc);
    if (COUNT && counters) { counters->inner += tc.inner; counters->tris += tc.tris; }
    return found;
}
__device__ __forceinline__ bool sceneOccluded(const DScene &s, V3 O, V3 D, float maxT)
{
    RayHit h;
    if (s.hasFilter) { VolumeEvents ev; return traverseFiltered<true>(s, O, D, PTC_TNEAR, maxT - 1e-3f, h, ev); }
    return traverseBVH<true, false>(s.bvh, O.x, O.y, O.z, D.x, D.y, D.z, PTC_TNEAR, maxT - 1e-3f, h, nullptr);
}

// Surface::getInternalMedium of the surface an intersection lies on (-1: none)
__device__ __forceinline__ int32_t internalMedium(const DScene &s, uint32_t prim)
{
    if (!s.nMedia) { return -1; }
    const uint32_t geom = (prim & PTC_SPHERE_FLAG) ? __ldg(s.sphereIds + (prim & ~PTC_SPHERE_FLAG)).x : __ldg(s.primIds + prim).x;
    return __ldg(s.geomMedium + geom);
}

// out-of-line copies of the shading library's building blocks for the volume integrator
PTC_VOLUME_CALL void volIsect(const DScene &s, V3 O, V3 D, const RayHit &h, Isect &out) { makeIsect(s, O, D, h, out); }
PTC_VOLUME_CALL void volBsdfSample(const DScene &s, const Isect &i, Rng &r, BsdfSample &out) { bsdfSample(s.materials[i.material], i, r, out); }
PTC_VOLUME_CALL V3 volBsdfEval(const DScene &s, const Isect &i, V3 wi, float &pdf) { return bsdfEval(s.materials[i.material], i, wi, pdf); }
PTC_VOLUME_CALL const DLight *volSampleLights(const DScene &s, V3 ref, Rng &r, SurfSample &out) { return sampleDirectLights(s, ref, r, out); }
PTC_VOLUME_CALL V3 volLightRadiance(const DScene &s, const DLight *light, V3 towardsLight)
{
    return __ldg(&light->kind) == 2 ? envRadiance(s, towardsLight) : mk(__ldg(&light->emit[0]), __ldg(&light->emit[1]), __ldg(&light->emit[2]));
}
PTC_VOLUME_CALL V3 volDirectBsdf(const DScene &s, V3 point, float cosTheta, V3 wi, float pdf, V3 thr, bool delta, bool hit, const Isect *bi)
{
    return directBsdf(s, point, cosTheta, wi, pdf, thr, delta, hit, bi, false);
}

// HomogeneousMedium::transmittance, src/homogeneous_medium.cpp:13-17: util::exp(-sigmaT * |b - a|)
__device__ __forceinline__ V3 mediumTransmittance(const DScene &s, int32_t medium, V3 a, V3 b)
{
    const float4 st = __ldg(s.media + 2 * medium);
    const float d = length(b - a);
    return mk(expHost(-st.x * d), expHost(-st.y * d), expHost(-st.z * d));
}

// VolumeHelper::rayTransmission, src/volume_helper.cpp:72-123 (current = the medium the path is in, -1 = none)
__device__ V3 rayTransmission(const DScene &s, V3 O, V3 D, const VolumeEvents &ev, int32_t current)
{
    V3 tr = mk(1.f, 1.f, 1.f);
    if (ev.count == 0) { return tr; }
    if (current >= 0) {
        if (ev.count == 1) { tr = tr * mediumTransmittance(s, current, O, O + D * ev.t[0]); }
        else if (ev.count == 2) { tr = tr * mediumTransmittance(s, current, O + D * ev.t[0], O + D * ev.t[1]); }
    } else {
        const int32_t m = ev.medium[0];
        if (ev.count == 2) { tr = tr * mediumTransmittance(s, m, O + D * ev.t[0], O + D * ev.t[1]); }
        else if (ev.count == 1) { tr = tr * mediumTransmittance(s, m, O, O + D * ev.t[0]); }
    }
    return tr; // more than two events: the reference's asserts are compiled out and nothing is applied
}

// rays traced by one path and their traversal work (closest-hit and any-hit kept apart, like the wavefront counters)
struct VolumeWork { uint32_t closestRays, shadowRays; TraverseCounters closest, shadow; };

// VolumeHelper::directSampleLights, src/volume_helper.cpp:12-70: single scattering from a point inside `medium`
// (isotropic phase function 1 / 4 pi, no sigma_s factor)
template <bool COUNT>
__device__ V3 volumeDirectLights(const DScene &s, int32_t medium, V3 point, Rng &r, VolumeWork *work)
{
    if (s.nLights == 0u) { return mk(0.f, 0.f, 0.f); } // no light to sample (undefined in the reference, Q18): no in-scattered light
    SurfSample ls;
    const DLight *light = volSampleLights(s, point, r, ls);
    const V3 sd = ls.point - point;
    const V3 wi = normalize(sd);
    if (dot(ls.normal, wi) >= 0.f) { return mk(0.f, 0.f, 0.f); }
    const float dist = length(sd);
    RayHit h; VolumeEvents ev;
    work->shadowRays++;
    if (traverseFiltered<true, COUNT>(s, point, wi, PTC_TNEAR, dist - 1e-3f, h, ev, &work->shadow)) { return mk(0.f, 0.f, 0.f); }
    const float pdf = solidAnglePdf(ls, point);
    const V3 lwo = -normalize(sd);
    V3 tr = mk(0.f, 0.f, 0.f);
    if (ev.count == 1) { tr = mediumTransmittance(s, medium, point, point + wi * ev.t[0]); }
    else if (ev.count == 2) { tr = mediumTransmittance(s, medium, point + wi * ev.t[0], point + wi * ev.t[1]); }
    const V3 Le = volLightRadiance(s, light, -lwo);
    return (((Le * tr) * 1.f) / (float)(4.f * PTC_PI_D)) / pdf;
}

// VolumePathTracer::scatter -> HomogeneousMedium::integrate, src/volume_path_tracer.cpp:112-131, src/homogeneous_medium.cpp:36-66
template <bool COUNT>
__device__ V3 mediumScatter(const DScene &s, int32_t medium, V3 entry, V3 exit, Rng &r, VolumeWork *work)
{
    if (medium < 0) { return mk(0.f, 0.f, 0.f); }
    const float sigmaT = __ldg(s.media + 2 * medium).x;
    const V3 travel = exit - entry;
    const float distance = length(travel);
    const float xi = r.next();
    const float sampleT = -logHost(1 - xi) / sigmaT;
    if (sampleT >= distance) { return mk(0.f, 0.f, 0.f); }
    const V3 samplePoint = entry + normalize(travel) * sampleT;
    return volumeDirectLights<COUNT>(s, medium, samplePoint, r, work);
}

// DirectLightingHelper::Ld, src/direct_lighting_helper.cpp:37-187
template <bool COUNT>
__device__ V3 volumeLd(const DScene &s, const Isect &i, int32_t medium, const BsdfSample &bs, Rng &r, VolumeWork *work)
{
    const DMaterial &m = s.materials[i.material];
    if (__ldg(&m.type) == PTC_PASSTHROUGH) { return mk(0.f, 0.f, 0.f); } // isContainer
    if (__ldg(&m.emitter)) { return mk(0.f, 0.f, 0.f); }
    V3 result = mk(0.f, 0.f, 0.f);
    if (!bs.delta && s.nLights != 0u) { // directSampleLights, :75-137 (no light at all: undefined in the reference, skipped here)
        SurfSample ls;
        const DLight *light = volSampleLights(s, i.point, r, ls);
        const V3 ld = ls.point - i.point;
        const V3 wi = normalize(ld);
        if (!(dot(ls.normal, wi) >= 0.f)) {
            const float dist = length(ld);
            RayHit h; VolumeEvents ev;
            work->shadowRays++;
            if (!traverseFiltered<true, COUNT>(s, i.point, wi, PTC_TNEAR, dist - 1e-3f, h, ev, &work->shadow)) {
                const V3 tr = rayTransmission(s, i.point, wi, ev, medium);
                const float pdf = solidAnglePdf(ls, i.point);
                float brdfPDF;
                const V3 f = volBsdfEval(s, i, wi, brdfPDF);
                const float w = (1 * pdf) / (1 * pdf + 1 * brdfPDF);
                const V3 lwo = -normalize(ld);
                const V3 Le = volLightRadiance(s, light, -lwo);
                result = result + ((((Le * tr) * w) * f) * fabsf(dot(i.ns, wi))) / pdf;
            }
        }
    }
    { // directSampleBSDF, :139-187: the probe ray skips containers; an emitter counts from either side, no transmittance
        RayHit h; VolumeEvents ev; Isect bi;
        work->closestRays++;
        const bool isHit = traverseFiltered<false, COUNT>(s, i.point, bs.wi, PTC_TNEAR, PTC_TFAR, h, ev, &work->closest);
        if (isHit) { volIsect(s, i.point, bs.wi, h, bi); }
        result = result + volDirectBsdf(s, i.point, fabsf(dot(i.ns, bs.wi)), bs.wi, bs.pdf, bs.thr, bs.delta, isHit, &bi);
    }
    return result;
}

// SampleIntegrator::samplePixel's body (src/sample_integrator.cpp:18-59, container branch included) + VolumePathTracer::L
// (src/volume_path_tracer.cpp:14-95) for one primary ray; the caller has positioned `r` (Philox: vertex 0 consumed the jitter)
template <bool COUNT>
__device__ V3 volumeRadiance(const DScene &s, V3 O, V3 D, Rng &r, int start, int last, VolumeWork *work)
{
    V3 color = mk(0.f, 0.f, 0.f);
    RayHit h;
    work->closestRays++;
    if (!sceneIntersect<COUNT>(s, O, D, h, &work->closest)) { return envRadiance(s, D); }
    Isect lastI;
    volIsect(s, O, D, h, lastI);
    if (checkCounts(start, last, 0)) {
        const DMaterial &m = s.materials[lastI.material];
        if (m.emitter && !(dot(lastI.n, lastI.wo) < 0.f)) { color = mk(m.emit[0], m.emit[1], m.emit[2]); }
        if (m.type == PTC_PASSTHROUGH) { // what lies behind the container, attenuated (src/sample_integrator.cpp:35-51)
            VolumeEvents ev; RayHit vh;
            work->closestRays++;
            const bool vHit = traverseFiltered<false, COUNT>(s, O, D, PTC_TNEAR, PTC_TFAR, vh, ev, &work->closest);
            const V3 tr = rayTransmission(s, O, D, ev, -1);
            if (vHit) {
                Isect vi; volIsect(s, O, D, vh, vi);
                const DMaterial &vm = s.materials[vi.material];
                color = color + mk(vm.emit[0], vm.emit[1], vm.emit[2]) * tr;
            } else { color = color + envRadiance(s, D) * tr; }
        }
    }
    // VolumePathTracer::L as ONE loop (the reference peels the first vertex, src/volume_path_tracer.cpp:21-31; with modulation = 1
    // there `result += Ld * modulation` is the same value), so that every building block has a single call site
    int32_t medium = -1;
    V3 result = mk(0.f, 0.f, 0.f), modulation = mk(1.f, 1.f, 1.f);
    r.beginVertex(1);
    for (int bounce = 1;;) {
        BsdfSample bs;
        volBsdfSample(s, lastI, r, bs);
        if (checkCounts(start, last, bounce)) {
            const V3 Ld = volumeLd<COUNT>(s, lastI, medium, bs, r, work);
            result = result + Ld * modulation;
        }
        bounce++;
        if (checkDone(last, bounce)) { break; }
        if (dot(lastI.wo, bs.wi) < 0.f) { // refraction: the medium changes (:42-50)
            if (dot(lastI.n, bs.wi) < 0.f) { medium = internalMedium(s, lastI.prim); }
            else { medium = -1; }
        }
        work->closestRays++;
        if (!sceneIntersect<COUNT>(s, lastI.point, bs.wi, h, &work->closest)) { break; }
        Isect bi;
        volIsect(s, lastI.point, bs.wi, h, bi);
        const float invPDF = 1.f / bs.pdf;
        const float cosT = fabsf(dot(lastI.ns, bs.wi));
        modulation = modulation * ((bs.thr * cosT) * invPDF);
        r.beginVertex((uint32_t)bounce);
        const V3 Ls = mediumScatter<COUNT>(s, medium, lastI.point, bi.point, r, work);
        result = result + Ls * modulation;
        if (medium >= 0) { modulation = modulation * mediumTransmittance(s, medium, lastI.point, bi.point); }
        if (isBlack(modulation)) { break; }
        lastI = bi;
    }
    return color + result;
}

} // namespace ptc
