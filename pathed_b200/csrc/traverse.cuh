// Single-ray traversal of the compressed 8-wide BVH: closest hit (rtcIntersect1) and any hit (rtcOccluded1).
// Shared by the sm_100a kernels and by the host-side counting traversal (identical arithmetic, so the host
// counts of inner-node visits / triangle tests are exactly what the kernels execute).
//
// Triangle test = Embree's Moeller-Trumbore in its exact operation order, including the fused multiply-adds
// of its AVX2 build (ext/embree/kernels/geometry/triangle_intersector_moeller.h:75-113, common/math/vec3.h:216,221):
// inclusive edge tests, |den|*tnear < T <= |den|*tfar, no back-face culling.
// Sphere test = ext/embree/kernels/geometry/sphere_intersector.h:67-106.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PTC_HD __host__ __device__ __forceinline__
#else
#define PTC_HD inline
#endif

namespace ptc {

PTC_HD uint32_t f2u(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}
PTC_HD float u2f(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
// IEEE-rounded division / square root regardless of compiler flags: hit records must match Embree's, and unit vectors feed
// 1 - cos^2 terms whose cancellation multiplies every ulp (a fast-math build was measured: Beckmann alpha = 0.005 eval and
// sample fixtures miss the 1e-5 gate), so the library is built with the precise forms throughout.
// On the device an exactly zero numerator is answered without dividing: the division sequence (MUFU.RCP + FFMA refinement) guards itself
// with FCHK, which sends every operand with a zero exponent field -- 0.0 included -- to a ~150-instruction slow path that the whole warp
// waits for.  Zeros are common here (axis-aligned normals, black colour channels, a Beckmann D that underflowed): measured 9 slow-path calls
// per Lambertian vertex on the dragon workload's floor, 15 % of the Plastic kernel's instructions.  0 / b = +-0 (NaN for b = 0 or NaN).
PTC_HD float divIeee(float a, float b)
{
#if defined(__CUDA_ARCH__)
    const bool zero = a == 0.f;
    const float q = __fdiv_rn(zero ? 1.f : a, b);
    const float z = (b == 0.f || b != b) ? __int_as_float(0x7FFFFFFF) : __uint_as_float((__float_as_uint(a) ^ __float_as_uint(b)) & 0x80000000u);
    return zero ? z : q;
#else
    return a / b;
#endif
}
PTC_HD float sqrtIeee(float a)
{
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(a);
#else
    return sqrtf(a);
#endif
}
PTC_HD uint32_t highestBit(uint32_t x) // x != 0
{
#if defined(__CUDA_ARCH__)
    return 31u - (uint32_t)__clz((int)x);
#else
    return 31u - (uint32_t)__builtin_clz(x);
#endif
}
PTC_HD uint32_t popCount(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}
PTC_HD uint32_t lowestBit(uint32_t x) // x != 0
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ffs((int)x) - 1u;
#else
    return (uint32_t)__builtin_ctz(x);
#endif
}
// byte j of x as the float 32768 + byte (one PRMT on the device, no int -> float conversion).  The 2^15 pattern comes from constant
// memory: PRMT takes ONE immediate, and with the pattern as a literal ptxas spends it on the pattern and materialises the selector of
// every one of the 48 PRMTs of a node test in a register first (measured: ~45 extra MOV / IMAD.U32 per node visit, 13 % of the node phase);
// a constant-bank operand leaves the immediate slot to the selector.
#if defined(__CUDACC__)
static __constant__ uint32_t c_byteMagic = 0x47000000u;
#endif
PTC_HD float byteToMagic(uint32_t x, int j)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(__byte_perm(x, c_byteMagic, 0x7504u | ((uint32_t)j << 4)));
#else
    return u2f(0x47000000u | (((x >> (8 * j)) & 0xFFu) << 8));
#endif
}
// byte j of x, zero-extended (one PRMT on the device)
PTC_HD uint32_t byteOf(uint32_t x, int j)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, 0u, 0x4440u | (uint32_t)j);
#else
    return (x >> (8 * j)) & 0xFFu;
#endif
}
// a * b + c rounded towards -inf / +inf
PTC_HD float fmaDown(float a, float b, float c)
{
#if defined(__CUDA_ARCH__)
    return __fmaf_rd(a, b, c);
#else
    const double exact = (double)a * (double)b + (double)c;
    float r = (float)exact;
    if ((double)r > exact) { r = nextafterf(r, -INFINITY); }
    return r;
#endif
}
PTC_HD float fmaUp(float a, float b, float c)
{
#if defined(__CUDA_ARCH__)
    return __fmaf_ru(a, b, c);
#else
    const double exact = (double)a * (double)b + (double)c;
    float r = (float)exact;
    if ((double)r < exact) { r = nextafterf(r, INFINITY); }
    return r;
#endif
}
PTC_HD float4 loadNodeWord(const float4 *p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

struct BvhView {
    const float4 *nodes;     // 5 float4 per node
    const float4 *triangles; // 3 float4 per triangle
    const float4 *spheres;   // center.xyz, radius
    uint32_t nSpheres;
    uint32_t nNodes;
    // occlusion filter (src/scene.cpp:42-84): per triangle / per sphere the medium of a Passthrough surface that encloses one
    // (such a hit is rejected and leaves a volume event), -1 for every other surface; null when the scene has no such surface
    const int32_t *primEvent, *sphereEvent;
    // world-to-local maps of the flattened instance placements, 6 float4 each (rows of the outer placement's 3x4
    // map, then of the inner one's for a two-level placement); null without instances.  A leaf triangle of a placement keeps its
    // LOCAL-space corners and carries (placement + 1) | two-level flag << 31 in the spare word of its second float4
    const float4 *placements;
};

#define PTC_SPHERE_FLAG 0x80000000u
#define PTC_MISS 0xFFFFFFFFu

struct RayHit {
    float t, u, v;
    uint32_t prim; // global triangle primitive, or PTC_SPHERE_FLAG | sphere slot, or PTC_MISS
};

// Embree's AVX2 Vec3 helpers: dot = madd chain, cross = msub
PTC_HD float edot(float ax, float ay, float az, float bx, float by, float bz) { return fmaf(ax, bx, fmaf(ay, by, az * bz)); }

PTC_HD bool triangleTest(const float4 a, const float4 b, const float4 c, float ox, float oy, float oz, float dx, float dy,
                         float dz, float tnear, float tfar, float &t, float &u, float &v)
{
    // a = (v0, prim), b = (e1 = v0 - v1), c = (e2 = v2 - v0); Ng = e2 x e1
    const float ngx = fmaf(c.y, b.z, -(c.z * b.y)), ngy = fmaf(c.z, b.x, -(c.x * b.z)), ngz = fmaf(c.x, b.y, -(c.y * b.x));
    const float cx = a.x - ox, cy = a.y - oy, cz = a.z - oz;
    const float rx = fmaf(cy, dz, -(cz * dy)), ry = fmaf(cz, dx, -(cx * dz)), rz = fmaf(cx, dy, -(cy * dx));
    const float den = edot(ngx, ngy, ngz, dx, dy, dz);
    const float absDen = fabsf(den);
    const uint32_t sgn = f2u(den) & 0x80000000u;
    const float U = u2f(f2u(edot(rx, ry, rz, c.x, c.y, c.z)) ^ sgn);
    const float V = u2f(f2u(edot(rx, ry, rz, b.x, b.y, b.z)) ^ sgn);
    if (!(den != 0.f && U >= 0.f && V >= 0.f && U + V <= absDen)) { return false; }
    const float T = u2f(f2u(edot(ngx, ngy, ngz, cx, cy, cz)) ^ sgn);
    if (!(absDen * tnear < T && T <= absDen * tfar)) { return false; }
    t = divIeee(T, absDen); u = divIeee(U, absDen); v = divIeee(V, absDen);
    return true;
}

// The same test with the divisions left to the caller: T, U, V are the sign-corrected numerators over absDen.  Used by the
// cooperative triangle phase, where the ray's owner re-applies the depth interval with its current hit distance.
PTC_HD bool triangleTestRaw(const float4 a, const float4 b, const float4 c, float ox, float oy, float oz, float dx, float dy,
                            float dz, float tnear, float tfar, float &T, float &U, float &V, float &absDen)
{
    const float ngx = fmaf(c.y, b.z, -(c.z * b.y)), ngy = fmaf(c.z, b.x, -(c.x * b.z)), ngz = fmaf(c.x, b.y, -(c.y * b.x));
    const float cx = a.x - ox, cy = a.y - oy, cz = a.z - oz;
    const float rx = fmaf(cy, dz, -(cz * dy)), ry = fmaf(cz, dx, -(cx * dz)), rz = fmaf(cx, dy, -(cy * dx));
    const float den = edot(ngx, ngy, ngz, dx, dy, dz);
    absDen = fabsf(den);
    const uint32_t sgn = f2u(den) & 0x80000000u;
    U = u2f(f2u(edot(rx, ry, rz, c.x, c.y, c.z)) ^ sgn);
    V = u2f(f2u(edot(rx, ry, rz, b.x, b.y, b.z)) ^ sgn);
    if (!(den != 0.f && U >= 0.f && V >= 0.f && U + V <= absDen)) { return false; }
    T = u2f(f2u(edot(ngx, ngy, ngz, cx, cy, cz)) ^ sgn);
    return absDen * tnear < T && T <= absDen * tfar;
}

PTC_HD bool sphereTest(const float4 s, float ox, float oy, float oz, float dx, float dy, float dz, float tnear, float tfar,
                       float &t, float &ngx, float &ngy, float &ngz)
{
    const float rd2 = divIeee(1.f, edot(dx, dy, dz, dx, dy, dz));
    const float c0x = s.x - ox, c0y = s.y - oy, c0z = s.z - oz;
    const float projC0 = edot(c0x, c0y, c0z, dx, dy, dz) * rd2;
    const float px = c0x - dx * projC0, py = c0y - dy * projC0, pz = c0z - dz * projC0;
    const float l2 = edot(px, py, pz, px, py, pz);
    const float r2 = s.w * s.w;
    if (!(l2 <= r2)) { return false; }
    float td = sqrtIeee((r2 - l2) * rd2);
    const float tIn = projC0 - td, tOut = projC0 + td;
    const bool validIn = (tIn > tnear) && (tIn < tfar);
    const bool validOut = !validIn && (tOut > tnear) && (tOut < tfar);
    if (!validIn && !validOut) { return false; }
    if (validIn) { td = -1.0f * td; }
    t = validIn ? tIn : tOut;
    ngx = dx * td - px; ngy = dy * td - py; ngz = dz * td - pz;
    return true;
}

// Embree's InstanceIntersector1 (ext/embree/kernels/geometry/instance_intersector.cpp:52-109): the ray enters an instance by
// xfmPoint(world2local, org) / xfmVector(world2local, dir) (AVX2 madd chains, common/math/affinespace.h), tnear / tfar unchanged, and
// the triangle is tested there -- so t, u, v carry the rounding of the instance's space, not of world space.  The placements are
// flattened into the one BVH (world-space boxes); the triangle test repeats Embree's transforms, level by level.
PTC_HD void rayToPlacement(const float4 *placements, uint32_t tag, float &ox, float &oy, float &oz, float &dx, float &dy, float &dz)
{
    const float4 *m = placements + (size_t)((tag & 0x7FFFFFFFu) - 1u) * 6;
    for (uint32_t level = 0; level < 1u + (tag >> 31); level++, m += 3) {
        const float4 r0 = loadNodeWord(m), r1 = loadNodeWord(m + 1), r2 = loadNodeWord(m + 2);
        const float px = fmaf(ox, r0.x, fmaf(oy, r0.y, fmaf(oz, r0.z, r0.w))), py = fmaf(ox, r1.x, fmaf(oy, r1.y, fmaf(oz, r1.z, r1.w))),
                    pz = fmaf(ox, r2.x, fmaf(oy, r2.y, fmaf(oz, r2.z, r2.w)));
        const float vx = fmaf(dx, r0.x, fmaf(dy, r0.y, dz * r0.z)), vy = fmaf(dx, r1.x, fmaf(dy, r1.y, dz * r1.z)), vz = fmaf(dx, r2.x, fmaf(dy, r2.y, dz * r2.z));
        ox = px; oy = py; oz = pz; dx = vx; dy = vy; dz = vz;
    }
}

struct TraverseCounters { uint32_t inner, tris; };

// 1: assemble the traversal mask in a loop over the hit children (measured slower: the loop runs at 8 of 32 lanes)
#ifndef PTC_HITMASK_LOOP
#define PTC_HITMASK_LOOP 0
#endif

#define PTC_STACK_SIZE 40

// Per-ray traversal state: lets a kernel interleave the phases of many rays (persistent warps that refill idle lanes).
struct TraversalState {
    float ox, oy, oz, dx, dy, dz, idx, idy, idz, tnear;
    uint32_t octInv;
    uint2 ngroup, tgroup;
    int sp;
    bool found;
    RayHit hit;     // while the BVH is being traversed hit.u / hit.v hold the numerators U, V of the closest triangle so far
    float hitDen;   // and hitDen its |den|: the two divisions are done once per ray (traversalSpheres), not once per accepted hit
    uint2 stack[PTC_STACK_SIZE];
};

// The first PTC_FAST_STACK entries live in shared memory when a kernel provides a slice (fast != nullptr; entry e of this
// ray at fast[e * PTC_FAST_STRIDE]): pushes and pops then cost no L1/L2 traffic, which the node fetches need.  Deeper entries,
// and all entries of the scalar callers, go to the local-memory array.
#ifndef PTC_FAST_STACK
#define PTC_FAST_STACK 8
#endif
#define PTC_FAST_STRIDE 128 /* = threads per CTA of the traversal kernels */

PTC_HD void stackPush(TraversalState &st, uint2 v, uint2 *fast)
{
    if (fast && st.sp < PTC_FAST_STACK) { fast[st.sp * PTC_FAST_STRIDE] = v; }
    else { st.stack[st.sp] = v; }
    st.sp++;
}
PTC_HD uint2 stackPop(TraversalState &st, uint2 *fast)
{
    st.sp--;
    if (fast && st.sp < PTC_FAST_STACK) { return fast[st.sp * PTC_FAST_STRIDE]; }
    return st.stack[st.sp];
}

PTC_HD void traversalInit(TraversalState &st, float ox, float oy, float oz, float dx, float dy, float dz, float tnear, float tfar)
{
    st.ox = ox; st.oy = oy; st.oz = oz; st.dx = dx; st.dy = dy; st.dz = dz; st.tnear = tnear;
    // reciprocal direction; zero components are nudged so that 0 * inf never appears in the slab test
    const float eps = 1e-30f;
    st.idx = 1.f / (fabsf(dx) > eps ? dx : (f2u(dx) & 0x80000000u ? -eps : eps));
    st.idy = 1.f / (fabsf(dy) > eps ? dy : (f2u(dy) & 0x80000000u ? -eps : eps));
    st.idz = 1.f / (fabsf(dz) > eps ? dz : (f2u(dz) & 0x80000000u ? -eps : eps));
    st.octInv = 7u - ((dx < 0.f ? 4u : 0u) | (dy < 0.f ? 2u : 0u) | (dz < 0.f ? 1u : 0u));
    st.ngroup = make_uint2(0u, 0x80000000u); // root = slot (7 ^ octInv) of a virtual parent with no siblings
    st.tgroup = make_uint2(0u, 0u);
    st.sp = 0;
    st.found = false;
    st.hit.t = tfar; st.hit.u = 0.f; st.hit.v = 0.f; st.hit.prim = PTC_MISS;
    st.hitDen = 1.f;
}

// The traversal is split into three per-ray phases so that a kernel can run each phase for all the rays of a warp that
// need it (CWBVH-style triangle postponing, Ylitie et al. 2017): the node phase visits one inner node, the triangle phase
// tests ONE triangle of the pending triangle group, the pop phase fetches the next group from the stack.  A stack entry is a
// node group (child base, hit bits << 24 | inner mask) or a postponed triangle group (triangle base, triangle bits < 2^24).
template <bool COUNT>
PTC_HD void traversalNode(const BvhView &bvh, TraversalState &st, TraverseCounters *counters, uint2 *fast = nullptr)
{
    uint2 ngroup = st.ngroup;
    if (!(ngroup.y & 0xFF000000u)) { // a postponed triangle group came off the stack
        st.tgroup = ngroup;
        st.ngroup = make_uint2(0u, 0u);
        return;
    }
    uint2 tgroup;
    {
        const uint32_t hitsImask = ngroup.y;
        const uint32_t childBit = highestBit(hitsImask);
        ngroup.y &= ~(1u << childBit);
        if ((ngroup.y & 0xFF000000u) && st.sp < PTC_STACK_SIZE) { stackPush(st, ngroup, fast); }
        const uint32_t slot = (childBit - 24u) ^ st.octInv;
        const uint32_t relative = popCount(hitsImask & ~(0xFFFFFFFFu << slot) & 0xFFu);
        const float4 *node = bvh.nodes + (size_t)(ngroup.x + relative) * 5;
        const float4 n0 = loadNodeWord(node + 0), n1 = loadNodeWord(node + 1), n2 = loadNodeWord(node + 2),
                     n3 = loadNodeWord(node + 3), n4 = loadNodeWord(node + 4);
        if (COUNT) { counters->inner++; }
        const uint32_t e = f2u(n0.w);
        // plane distance t = (origin + q * 2^e - o) / d = q * a + b.  q is turned into a float without a conversion
        // instruction: the byte is dropped into mantissa bits 8..15 of 2^15, which reads 32768 + q exactly, and the 32768 * a
        // is taken out of b once per node (rounded down for entry planes, up for exit planes, so boxes stay conservative).
        const float ax = u2f((e & 0xFFu) << 23) * st.idx, ay = u2f(((e >> 8) & 0xFFu) << 23) * st.idy,
                    az = u2f(((e >> 16) & 0xFFu) << 23) * st.idz;
        const float bx = (n0.x - st.ox) * st.idx, by = (n0.y - st.oy) * st.idy, bz = (n0.z - st.oz) * st.idz;
        const float bx0 = fmaDown(-32768.f, ax, bx), by0 = fmaDown(-32768.f, ay, by), bz0 = fmaDown(-32768.f, az, bz);
        const float bx1 = fmaUp(-32768.f, ax, bx), by1 = fmaUp(-32768.f, ay, by), bz1 = fmaUp(-32768.f, az, bz);
        ngroup.x = f2u(n1.x);
        tgroup.x = f2u(n1.y);
#if PTC_HITMASK_LOOP
        uint32_t hits8 = 0;
#else
        // octant-ordered traversal mask: the per-child meta byte is 0b cccxxxxx (ccc = child bits to set, xxxxx = first bit;
        // inner children, xxxxx >= 24, have their slot xor-ed with the ray octant).  The four children of a meta word are decoded
        // at once with byte-parallel integer operations, so that the slab loop below only extracts two bytes per hit child.
        uint32_t hitmask = 0;
        const uint32_t octInv4 = st.octInv * 0x01010101u;
#endif
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const uint32_t qlox = f2u(half ? n2.y : n2.x), qloy = f2u(half ? n2.w : n2.z), qloz = f2u(half ? n3.y : n3.x);
            const uint32_t qhix = f2u(half ? n3.w : n3.z), qhiy = f2u(half ? n4.y : n4.x), qhiz = f2u(half ? n4.w : n4.z);
            const uint32_t xmin = st.dx < 0.f ? qhix : qlox, xmax = st.dx < 0.f ? qlox : qhix;
            const uint32_t ymin = st.dy < 0.f ? qhiy : qloy, ymax = st.dy < 0.f ? qloy : qhiy;
            const uint32_t zmin = st.dz < 0.f ? qhiz : qloz, zmax = st.dz < 0.f ? qloz : qhiz;
#if !PTC_HITMASK_LOOP
            const uint32_t meta4 = f2u(half ? n1.w : n1.z);
            const uint32_t inner4 = ((meta4 & (meta4 << 1)) & 0x10101010u) >> 4; // 1 in every byte of an inner child
            const uint32_t index4 = (meta4 ^ (octInv4 & (inner4 * 0xFFu))) & 0x1F1F1F1Fu;
            const uint32_t bits4 = (meta4 >> 5) & 0x07070707u;
#endif
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float t0x = fmaf(byteToMagic(xmin, j), ax, bx0), t1x = fmaf(byteToMagic(xmax, j), ax, bx1);
                const float t0y = fmaf(byteToMagic(ymin, j), ay, by0), t1y = fmaf(byteToMagic(ymax, j), ay, by1);
                const float t0z = fmaf(byteToMagic(zmin, j), az, bz0), t1z = fmaf(byteToMagic(zmax, j), az, bz1);
                const float tmin = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, st.tnear));
                const float tmax = fminf(fminf(t1x, t1y), fminf(t1z, st.hit.t)) * 1.0000004f;
#if PTC_HITMASK_LOOP
                if (tmin <= tmax) { hits8 |= 1u << (4 * half + j); }
#else
                if (tmin <= tmax) { hitmask |= byteOf(bits4, j) << byteOf(index4, j); }
#endif
            }
        }
#if PTC_HITMASK_LOOP
        // octant-ordered traversal mask, assembled for the children that were hit only
        uint32_t hitmask = 0;
        const uint32_t metaLo = f2u(n1.z), metaHi = f2u(n1.w);
        while (hits8) {
            const uint32_t i = lowestBit(hits8);
            hits8 &= hits8 - 1u;
            const uint32_t meta = ((i & 4u ? metaHi : metaLo) >> (8u * (i & 3u))) & 0xFFu;
            const uint32_t isInner = (meta & (meta << 1)) & 0x10u; // inner: 0b001xxxxx with xxxxx >= 24
            const uint32_t bitIndex = (meta ^ (isInner ? st.octInv : 0u)) & 0x1Fu;
            hitmask |= (meta >> 5) << bitIndex;
        }
#endif
        ngroup.y = (hitmask & 0xFF000000u) | (e >> 24);
        tgroup.y = hitmask & 0x00FFFFFFu;
    }
    st.ngroup = ngroup;
    st.tgroup = tgroup;
}

// Tests one triangle of the pending group (precondition: st.tgroup.y != 0).  Returns true when a hit was accepted.
// FILTER: Scene::testOcclusion's shouldIntersectPassthroughs = false -- container surfaces are not hits.
// INSTANCES = false: the caller knows the scene has no instance placements (the wavefront kernels of such a scene).
template <bool COUNT, bool FILTER = false, bool INSTANCES = true>
PTC_HD bool traversalTriangle(const BvhView &bvh, TraversalState &st, TraverseCounters *counters)
{
    const uint32_t bit = highestBit(st.tgroup.y);
    st.tgroup.y &= ~(1u << bit);
    const float4 *tri = bvh.triangles + (size_t)(st.tgroup.x + bit) * 3;
    const float4 a = loadNodeWord(tri), b = loadNodeWord(tri + 1), c = loadNodeWord(tri + 2);
    if (COUNT) { counters->tris++; }
    float T, U, V, absDen;
    if (INSTANCES && f2u(b.w)) {
        float ox = st.ox, oy = st.oy, oz = st.oz, dx = st.dx, dy = st.dy, dz = st.dz;
        rayToPlacement(bvh.placements, f2u(b.w), ox, oy, oz, dx, dy, dz);
        if (!triangleTestRaw(a, b, c, ox, oy, oz, dx, dy, dz, st.tnear, st.hit.t, T, U, V, absDen)) { return false; }
    } else if (!triangleTestRaw(a, b, c, st.ox, st.oy, st.oz, st.dx, st.dy, st.dz, st.tnear, st.hit.t, T, U, V, absDen)) { return false; }
    const float t = divIeee(T, absDen);
    const uint32_t prim = f2u(a.w);
    if (FILTER && bvh.primEvent[prim] >= 0) { return false; }
    // equal depth (shared edges, coincident faces): keep the larger primitive index, as a linear scan would
    if (!(st.found && t == st.hit.t && prim < st.hit.prim)) { st.hit.t = t; st.hit.u = U; st.hit.v = V; st.hitDen = absDen; st.hit.prim = prim; }
    st.found = true;
    return true;
}

// Puts the pending triangle group back on the stack (tested later, together with other rays' triangles); false = no room.
PTC_HD bool traversalPostpone(TraversalState &st, uint2 *fast = nullptr)
{
    if (st.sp >= PTC_STACK_SIZE) { return false; }
    stackPush(st, st.tgroup, fast);
    st.tgroup.y = 0u;
    return true;
}

// Next group once the current node group has no inner hits left.  Returns true when the BVH part of the traversal is finished.
PTC_HD bool traversalPop(TraversalState &st, uint2 *fast = nullptr)
{
    if (st.ngroup.y & 0xFF000000u) { return false; }
    if (st.sp == 0) { return true; }
    st.ngroup = stackPop(st, fast);
    return false;
}

// spheres: a handful per scene (mis-pbrt: 5), tested linearly after the mesh BVH; strict depth test
template <bool ANY, bool FILTER = false>
PTC_HD bool traversalSpheres(const BvhView &bvh, TraversalState &st)
{
    if (ANY && st.found) { return true; }
    if (st.found) { st.hit.u = divIeee(st.hit.u, st.hitDen); st.hit.v = divIeee(st.hit.v, st.hitDen); } // u = U / |den|, v = V / |den|
    for (uint32_t s = 0; s < bvh.nSpheres; s++) {
        float t, nx, ny, nz;
        if (sphereTest(loadNodeWord(bvh.spheres + s), st.ox, st.oy, st.oz, st.dx, st.dy, st.dz, st.tnear, st.hit.t, t, nx, ny, nz)) {
            if (FILTER && bvh.sphereEvent[s] >= 0) { continue; }
            st.hit.t = t; st.hit.u = 0.f; st.hit.v = 0.f; st.hit.prim = PTC_SPHERE_FLAG | s;
            st.found = true;
            if (ANY) { return true; }
        }
    }
    return st.found;
}

// ANY: return at the first accepted hit.  On return hit.t holds the closest t (or the input tfar on a miss).
template <bool ANY, bool COUNT>
PTC_HD bool traverseBVH(const BvhView &bvh, float ox, float oy, float oz, float dx, float dy, float dz, float tnear,
                        float tfar, RayHit &hit, TraverseCounters *counters)
{
    TraversalState st;
    traversalInit(st, ox, oy, oz, dx, dy, dz, tnear, tfar);
    if (bvh.nNodes) { // one ray on its own: node, all of its triangles, pop (nothing is postponed)
        for (;;) {
            traversalNode<COUNT>(bvh, st, counters);
            bool stop = false;
            while (st.tgroup.y) { if (traversalTriangle<COUNT>(bvh, st, counters) && ANY) { stop = true; break; } }
            if (stop || traversalPop(st)) { break; }
        }
    }
    const bool found = traversalSpheres<ANY>(bvh, st);
    hit = st.hit;
    return found;
}

} // namespace ptc
