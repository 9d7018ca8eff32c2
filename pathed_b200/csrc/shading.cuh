// Device-side shading library: Intersection construction, tangent frames, every BSDF's f / pdf / sample,
// area / sphere / environment light sampling and pdfs, camera rays, Philox streams.
// Each function names the reference code whose arithmetic it reproduces (paths relative to the reference root).
// Built with -fmad=false: the reference is compiled for baseline x86-64 (no FMA), so products and sums are kept
// separate; the only fused operations are the explicit fmaf() calls that mirror Embree's AVX2 interpolation.
#pragma once

#include "traverse.cuh"

#include "../../include/pathed_cuda.h"

namespace ptc {

#define PTC_INV_PI 0.3183098861837907f    /* include/util.h:10 */
#define PTC_TWO_PI_F 6.283185307179586f   /* include/util.h:11 */
#define PTC_PI_D 3.14159265358979323846   /* M_PI; the reference promotes a few expressions to double through it */
#define PTC_TNEAR 1e-3f                   /* src/scene.cpp:102 */
#define PTC_TFAR 1e5f                     /* src/scene.cpp:103 */

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float t) { return mk(a.x * t, a.y * t, a.z * t); }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ V3 operator/(V3 a, float t) { return mk(divIeee(a.x, t), divIeee(a.y, t), divIeee(a.z, t)); } // IEEE division, zero numerators answered directly
__device__ __forceinline__ V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }            // src/vector.cpp:18-21
__device__ __forceinline__ float length(V3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }          // :28-35
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return mk((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x)); } // :37-44
// :46-62
__device__ __forceinline__ V3 normalize(V3 a) { const float n = sqrtIeee(a.x * a.x + a.y * a.y + a.z * a.z); return mk(divIeee(a.x, n), divIeee(a.y, n), divIeee(a.z, n)); }
__device__ __forceinline__ V3 reflect(V3 w, V3 n) { return ((n * dot(w, n)) * 2.f) - w; }                    // :64-67
__device__ __forceinline__ bool isBlack(V3 c) { return c.x == 0.f && c.y == 0.f && c.z == 0.f; }              // src/color.cpp:14-17
__device__ __forceinline__ bool same(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(hi, fmaxf(v, lo)); }      // include/util.h:39-41

// ------------------------------------------------------------------------------------------------ host-matched libm
// Directions go through the host's libm in the reference (sinf / cosf of an azimuth, logf / atanf of a random number), and several
// quantities downstream are ill-conditioned in the direction: D(wh) of a narrow lobe (tan^2 = (1 - y^2) / y^2 moves by 1e-7 / theta^2 per
// ulp of y), the triangle a bounce ray lands on.  glibc's float sinf / cosf / logf return the correctly rounded value for 98.7 - 99.3 %
// of their inputs (measured against the rounded double result; they are evaluated in double), expf for 99.94 %; libdevice's float
// versions are only within 1 - 2 ulp.  So the sampling code calls the double-precision functions and rounds once (PTC_HOST_TRIG 1):
// sampled directions are then bit-identical to the reference's for ~97 % of the samples instead of about half.  atanf, acosf, atan2f
// and tanf of glibc are not correctly rounded often enough (84 - 96 %) for this to pin them; atanf still gains, the rest stay float.
#ifndef PTC_HOST_TRIG
#define PTC_HOST_TRIG 1
#endif
#if PTC_HOST_TRIG
__device__ __forceinline__ float sinHost(float x) { return (float)sin((double)x); }
__device__ __forceinline__ float cosHost(float x) { return (float)cos((double)x); }
__device__ __forceinline__ float logHost(float x) { return (float)log((double)x); }
__device__ __forceinline__ float atanHost(float x) { return (float)atan((double)x); }
__device__ __forceinline__ void sincosHost(float x, float &sn, float &cs) { double s, c; sincos((double)x, &s, &c); sn = (float)s; cs = (float)c; }
#else
__device__ __forceinline__ float sinHost(float x) { return sinf(x); }
__device__ __forceinline__ float cosHost(float x) { return cosf(x); }
__device__ __forceinline__ float logHost(float x) { return logf(x); }
__device__ __forceinline__ float atanHost(float x) { return atanf(x); }
__device__ __forceinline__ void sincosHost(float x, float &sn, float &cs) { sn = sinf(x); cs = cosf(x); }
#endif
// expf stays libdevice's: glibc's is correctly rounded for 99.94 % of inputs and nothing downstream amplifies its last bit (BSDF eval
// agrees with the reference within 2e-7 on all 2^16 tuples of every configuration either way)
__device__ __forceinline__ float expHost(float x) { return expf(x); }

// ------------------------------------------------------------------------------------------------ scene
struct DMaterial {
    int32_t type, distribution, albedoKind, emitter;
    float diffuse[3], sigmaA;  // sigmaA / sigmaB: OrenNayar's A and B (src/oren_nayar.cpp:11-19)
    float emit[3], sigmaB;
    float ior, alpha, resU, resV;
    float on[3]; int32_t texW;   // texW, texH, texels: albedoKind == PTC_ALBEDO_TEXTURE
    float off[3]; int32_t texH;
    const uint32_t *texels;      // r | g << 8 | b << 16 per texel, row 0 = top of the image (stbi_load order)
    uint64_t pad0;
};

struct DLight {
    int32_t kind, pad[3]; // 0 triangle, 1 sphere, 2 environment
    float p0[4], p1[4], p2[4];
    float centerRadius[4];
    float emit[4];
};

struct DScene {
    BvhView bvh;
    const float4 *positions; // per vertex
    const float4 *normals;
    const float2 *uvs;
    const uint4 *prims;      // per triangle: i0, i1, i2, material
    const uint2 *primIds;    // per triangle: geomID, primID (inside the scene the mesh is attached to)
    const uint2 *instIds;    // per triangle: RTCHit::instID[0..1] of the placement it was flattened from; null without instances (N4)
    const uint2 *sphereIds;  // per sphere: geomID, material
    const uint8_t *primClass, *sphereClass; // per triangle / sphere: material type | emitter << 3 (all the logic stage needs of a surface)
    // per triangle, everything Scene::testIntersect's post-processing needs in one 80-byte record (5 float4): unnormalised Ng as the
    // traversal computes it + material | n0.xyz n1.x | n1.yz n2.xy | n2.z uv0.xy uv1.x | uv1.y uv2.xy.  Ten scattered sectors
    // (index record, 3 positions, 3 normals, 3 uvs) become three contiguous ones; built at ptc_commit with the same arithmetic.
    const float4 *triShade;
    const DMaterial *materials;
    const DLight *lights;
    uint32_t nLights;
    int32_t hasEnv, envW, envH, envThetaEmpty;
    float envScale;
    const float4 *envRgba;
    const float *envThetaCdf, *envPhiCdf;
    const uint8_t *envPhiEmpty;
    // guide tables: guide[g] = first index whose cdf value is >= g / G (G a power of two), so a sample in [g/G, (g+1)/G) only
    // has to search cdf[guide[g] .. guide[g+1]]; same index as the full search, a handful of loads instead of log2(n)
    const uint16_t *envThetaGuide, *envPhiGuide;
    int32_t envThetaG, envPhiG;
    float envM2W[12], envW2M[12];
    float camToWorld[12];
    float vfov;
    int32_t width, height;
    // participating media (volume.cuh): sigma_t rgb / sigma_s rgb per medium; internal medium per geometry (-1 none);
    // hasFilter: some Passthrough surface encloses a medium, i.e. bvh.primEvent / bvh.sphereEvent are in use
    const float4 *media;
    const int32_t *geomMedium;
    int32_t nMedia, hasFilter;
};

__device__ __forceinline__ V3 xfVec(const float *m, V3 v) // src/transform.cpp:89-100
{
    return mk(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z, m[8] * v.x + m[9] * v.y + m[10] * v.z);
}

// ------------------------------------------------------------------------------------------------ RNG
// Replaces RandomGenerator (src/random_generator.cpp:4-11) and std::rand (src/camera.cpp:51-52) with
// counter-based Philox4x32-10: key = seed, counter = (pixel, sample, bounce, block); draw d = lane d&3 of block d>>2.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int round = 0; round < 10; round++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct Rng {
    const float *replay; // test hook: sequential draws from an explicit array
    uint32_t replayCount, replayUsed;
    uint32_t k0, k1, pixel, sample, bounce, draw;
    uint32_t block[4];
    uint32_t cached; // block index held in `block`, 0xFFFFFFFF = none

    __device__ __forceinline__ void initPhilox(uint64_t seed, uint32_t pixel_, uint32_t sample_)
    {
        replay = nullptr; replayCount = replayUsed = 0;
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32); pixel = pixel_; sample = sample_; bounce = 0; draw = 0; cached = 0xFFFFFFFFu;
    }
    __device__ __forceinline__ void initReplay(const float *xi, uint32_t count)
    {
        replay = xi; replayCount = count; replayUsed = 0; cached = 0xFFFFFFFFu; draw = 0; bounce = 0;
    }
    __device__ __forceinline__ void beginVertex(uint32_t b)
    {
        if (!replay) { bounce = b; draw = 0; cached = 0xFFFFFFFFu; }
    }
    __device__ __forceinline__ float next()
    {
        if (replay) { const float xi = replay[replayUsed % replayCount]; replayUsed++; return xi; }
        const uint32_t b = draw >> 2;
        if (b != cached) { philox4x32_10(pixel, sample, bounce, b, k0, k1, block); cached = b; }
        const uint32_t lane = draw & 3u;
        draw++;
        const uint32_t bits = lane == 0 ? block[0] : (lane == 1 ? block[1] : (lane == 2 ? block[2] : block[3]));
        return (float)(bits >> 8) * (1.0f / 16777216.0f); // [0, 1)
    }
};

// ------------------------------------------------------------------------------------------------ intersection
// The reference's Intersection (include/intersection.h:13-56); the two 4x4 frame Transforms reduce to 2 axes + ns.
struct Isect {
    V3 point, wo, n, ns, tx, tz;
    float u, v;
    uint32_t material;
    uint32_t prim; // global triangle index or PTC_SPHERE_FLAG | slot
};

__device__ __forceinline__ void makeFrame(V3 normal, V3 dir, V3 &xAxis, V3 &zAxis) // src/transform.cpp:182-218
{
    if (same(normal, dir)) {
        V3 xa;
        if (fabsf(normal.x) > fabsf(normal.y)) { xa = normalize(mk(-normal.z, 0.f, normal.x)); }
        else { xa = normalize(mk(0.f, -normal.z, normal.y)); }
        xAxis = xa; zAxis = cross(normal, xa);
        return;
    }
    xAxis = normalize(cross(normal, dir));
    zAxis = normalize(cross(normal, xAxis));
}
__device__ __forceinline__ V3 toWorld(const Isect &i, V3 l)
{
    return mk(i.tx.x * l.x + i.ns.x * l.y + i.tz.x * l.z, i.tx.y * l.x + i.ns.y * l.y + i.tz.y * l.z, i.tx.z * l.x + i.ns.z * l.y + i.tz.z * l.z);
}
__device__ __forceinline__ V3 toLocal(const Isect &i, V3 w)
{
    return mk(i.tx.x * w.x + i.tx.y * w.y + i.tx.z * w.z, i.ns.x * w.x + i.ns.y * w.y + i.ns.z * w.z, i.tz.x * w.x + i.tz.y * w.y + i.tz.z * w.z);
}

#ifndef PTC_FAT_TRIANGLES
#define PTC_FAT_TRIANGLES 1
#endif
// geometric normal exactly as the traversal computed it (Embree returns Ng = e2 x e1 with fused msub)
__device__ __forceinline__ V3 triangleNg(const DScene &s, uint32_t prim, uint4 &ix)
{
    ix = __ldg(s.prims + prim);
    const float4 v0 = __ldg(s.positions + ix.x), v1 = __ldg(s.positions + ix.y), v2 = __ldg(s.positions + ix.z);
    const float e1x = v0.x - v1.x, e1y = v0.y - v1.y, e1z = v0.z - v1.z;
    const float e2x = v2.x - v0.x, e2y = v2.y - v0.y, e2z = v2.z - v0.z;
    return mk(fmaf(e2y, e1z, -(e2z * e1y)), fmaf(e2z, e1x, -(e2x * e1z)), fmaf(e2x, e1y, -(e2y * e1x)));
}

// Scene::testIntersect post-processing, src/scene.cpp:122-219 (hit must be a hit)
__device__ __forceinline__ void makeIsect(const DScene &s, V3 O, V3 D, const RayHit &h, Isect &r, V3 *ngRaw = nullptr)
{
    V3 ngU, ns = mk(0.f, 0.f, 0.f);
    r.u = 0.f; r.v = 0.f;
    if (h.prim & PTC_SPHERE_FLAG) {
        const uint32_t slot = h.prim & ~PTC_SPHERE_FLAG;
        float t, nx, ny, nz;
        // re-derive Ng with the arithmetic of the hit test (deterministic): widen the interval around the stored t
        sphereTest(__ldg(s.bvh.spheres + slot), O.x, O.y, O.z, D.x, D.y, D.z, h.t * 0.999999f - 1e-30f, h.t * 1.000001f + 1e-30f, t, nx, ny, nz);
        ngU = mk(nx, ny, nz);
        r.material = __ldg(s.sphereIds + slot).y;
    } else {
#if PTC_FAT_TRIANGLES
        const float4 *rec = s.triShade + (size_t)h.prim * 5;
        const float4 r0 = __ldg(rec), r1 = __ldg(rec + 1), r2 = __ldg(rec + 2), r3 = __ldg(rec + 3), r4 = __ldg(rec + 4);
        ngU = mk(r0.x, r0.y, r0.z);
        uint4 ix; ix.w = __float_as_uint(r0.w);
        const float4 n0 = make_float4(r1.x, r1.y, r1.z, 0.f), n1 = make_float4(r1.w, r2.x, r2.y, 0.f), n2 = make_float4(r2.z, r2.w, r3.x, 0.f);
        const float2 t0 = make_float2(r3.y, r3.z), t1 = make_float2(r3.w, r4.x), t2 = make_float2(r4.y, r4.z);
#else
        uint4 ix;
        ngU = triangleNg(s, h.prim, ix);
        const float4 n0 = __ldg(s.normals + ix.x), n1 = __ldg(s.normals + ix.y), n2 = __ldg(s.normals + ix.z);
        const float2 t0 = __ldg(s.uvs + ix.x), t1 = __ldg(s.uvs + ix.y), t2 = __ldg(s.uvs + ix.z);
#endif
        // rtcInterpolate0, ext/embree/kernels/common/scene_triangle_mesh.cpp:248-253: madd(w, p0, madd(u, p1, v * p2))
        const float w = 1.0f - h.u - h.v;
        ns = mk(fmaf(w, n0.x, fmaf(h.u, n1.x, h.v * n2.x)), fmaf(w, n0.y, fmaf(h.u, n1.y, h.v * n2.y)), fmaf(w, n0.z, fmaf(h.u, n1.z, h.v * n2.z)));
        r.u = fmaf(w, t0.x, fmaf(h.u, t1.x, h.v * t2.x));
        r.v = fmaf(w, t0.y, fmaf(h.u, t1.y, h.v * t2.y));
        r.material = ix.w;
    }
    if (ngRaw) { *ngRaw = ngU; }
    const V3 ng = normalize(ngU);
    if (length(ns) == 0.f) { ns = ng; }
    r.point = O + D * h.t;
    r.wo = -D;
    r.n = ng;
    r.ns = normalize(ns);
    r.prim = h.prim;
    makeFrame(r.ns, r.wo, r.tx, r.tz);
}

// ------------------------------------------------------------------------------------------------ BSDFs
// include/tangent_frame.h (local frame is y-up)
__device__ __forceinline__ float tfCos2(V3 v) { return v.y * v.y; }
__device__ __forceinline__ float tfSin(V3 v) { return sqrtIeee(fmaxf(0.f, 1.f - tfCos2(v))); }
__device__ __forceinline__ float tfSin2(V3 v) { return 1.f - tfCos2(v); }
__device__ __forceinline__ float tfTan(V3 v) { return divIeee(tfSin(v), v.y); }
__device__ __forceinline__ float tfTan2(V3 v) { return divIeee(tfSin2(v), tfCos2(v)); }
__device__ __forceinline__ float tfCosPhi(V3 v) // :69-76
{
    const float s = tfSin(v);
    if (s == 0.f) { return 1.f; }
    return clampf(divIeee(v.x, s), -1.f, 1.f);
}
__device__ __forceinline__ float tfSinPhi(V3 v) // :11-37, :83-92 (axis snap at +-0.9999)
{
    const float max = 0.9999f;
    V3 c = v;
    if (v.x >= max) { c = mk(1.f, 0.f, 0.f); }
    else if (v.y >= max) { c = mk(0.f, 1.f, 0.f); }
    else if (v.z >= max) { c = mk(0.f, 0.f, 1.f); }
    else if (v.x <= -max) { c = mk(-1.f, 0.f, 0.f); }
    else if (v.y <= -max) { c = mk(0.f, -1.f, 0.f); }
    else if (v.z <= -max) { c = mk(0.f, 0.f, -1.f); }
    const float s = tfSin(c);
    if (s == 0.f) { return 0.f; }
    return clampf(divIeee(c.z, s), -1.f, 1.f);
}

__device__ __forceinline__ void cartToSph(V3 c, float &phi, float &theta) // src/coordinate.cpp:7-19
{
    phi = atan2f(c.z, c.x);
    if (phi < 0.f) { phi = (float)((double)phi + 2 * PTC_PI_D); }
    if (phi == PTC_TWO_PI_F) { phi = 0.f; }
    theta = acosf(clampf(c.y, -1.f, 1.f));
}
__device__ __forceinline__ V3 sphToCart(float phi, float cosTheta, float sinTheta) // :26-32
{
    float sn, cs;
    sincosHost(phi, sn, cs); // one double-precision sincos serves both
    return mk(sinTheta * cs, cosTheta, sinTheta * sn);
}

__device__ __forceinline__ V3 cosineSample(Rng &r) // src/monte_carlo.cpp:24-41
{
    const float xi1 = r.next();
    const float rad = sqrtf(xi1);
    const float phi = (float)(2 * PTC_PI_D * (double)r.next());
    float sn, cs;
    sincosHost(phi, sn, cs);
    return mk(rad * cs, sqrtf(1.f - xi1), rad * sn);
}

// pow(c / 255.f, 2.2f) for the 256 byte values, tabulated by the host's powf at ptc_commit (bit-identical to the reference's
// per-lookup powf, src/texture.cpp:44-48)
__constant__ float c_gammaTable[256];

__device__ __forceinline__ V3 lambertAlbedo(const DMaterial &m, const Isect &i) // src/checkerboard.cpp:9-20, src/texture.cpp:34-49
{
    if (m.albedoKind == PTC_ALBEDO_TEXTURE) {
        // wrap, flip v, nearest texel by roundf on (size - 1)
        const float u = i.u - (int)floorf(i.u);
        const float v = 1.f - (i.v - (int)floorf(i.v));
        const int x = (int)roundf(u * (m.texW - 1)), y = (int)roundf(v * (m.texH - 1));
        const uint32_t texel = __ldg(m.texels + (size_t)y * m.texW + x);
        return mk(c_gammaTable[texel & 0xFFu], c_gammaTable[(texel >> 8) & 0xFFu], c_gammaTable[(texel >> 16) & 0xFFu]);
    }
    if (m.albedoKind == PTC_ALBEDO_CHECKERBOARD) {
        const int ui = (int)floorf(i.u * m.resU), vi = (int)floorf(i.v * m.resV);
        if (ui % 2 == vi % 2) { return mk(m.on[0], m.on[1], m.on[2]); }
        return mk(m.off[0], m.off[1], m.off[2]);
    }
    return mk(m.diffuse[0], m.diffuse[1], m.diffuse[2]);
}

__device__ __forceinline__ float fresnelDielectric(float cosI, float etaI, float etaT) // src/fresnel.cpp:30-64, src/snell.cpp:51-57
{
    const float sinT = (etaI / etaT) * sqrtf(fmaxf(0.f, 1.f - cosI * cosI));
    if (sinT > 1.f) { return 1.f; }
    const float cosT = sqrtf(fmaxf(0.f, 1.f - sinT * sinT));
    const float rPar = (etaT * cosI - etaI * cosT) / (etaT * cosI + etaI * cosT);
    const float rPerp = (etaI * cosI - etaT * cosT) / (etaI * cosI + etaT * cosT);
    return 0.5f * (rPar * rPar + rPerp * rPerp);
}

__device__ __forceinline__ float mfD(const DMaterial &m, V3 wh) // src/beckmann.cpp:45-65, src/ggx.cpp:31-46
{
    const float alpha2 = m.alpha * m.alpha;
    const float tan2 = tfTan2(wh);
    if (isinf(tan2)) { return 0.f; }
    const float cos2 = tfCos2(wh);
    const float cos4 = cos2 * cos2;
    if (m.distribution == PTC_BECKMANN) {
        const float cp = tfCosPhi(wh), sp = tfSinPhi(wh);
        const float num = expHost(-tan2 * (divIeee(cp * cp, alpha2) + divIeee(sp * sp, alpha2)));
        const float den = (float)(PTC_PI_D * (double)alpha2 * (double)cos4);
        return divIeee(num, den); // num underflows to 0 for directions far off a narrow lobe
    }
    const float sum = alpha2 + tan2;
    const float den = (float)(PTC_PI_D * (double)cos4 * (double)sum * (double)sum);
    return alpha2 / den;
}
__device__ __forceinline__ float mfPdf(const DMaterial &m, V3 wh) { return mfD(m, wh) * fabsf(wh.y); }
__device__ __forceinline__ float beckmannLambda(float alpha, V3 w) // src/beckmann.cpp:67-81
{
    const float absTan = fabsf(tfTan(w));
    if (isinf(absTan)) { return 0.f; }
    const float cp = tfCosPhi(w), sp = tfSinPhi(w);
    const float a_ = sqrtf((cp * cp) * alpha * alpha + (sp * sp) * alpha * alpha);
    const float a = 1.f / (a_ * absTan);
    if (a >= 1.6f) { return 0.f; }
    return (1 - 1.259f * a + 0.396f * a * a) / (3.535f * a + 2.181f * a * a);
}
__device__ __forceinline__ float ggxG1(float alpha, V3 v) // src/ggx.cpp:48-58
{
    const float tan2 = tfTan2(v);
    if (isinf(tan2)) { return 0.f; }
    const float s = (1 + alpha * alpha * tan2);
    return 2.f / (1 + sqrtf(s));
}
__device__ __forceinline__ float mfG(const DMaterial &m, V3 wo, V3 wi)
{
    if (m.distribution == PTC_BECKMANN) { return 1.f / (1.f + beckmannLambda(m.alpha, wo) + beckmannLambda(m.alpha, wi)); }
    return ggxG1(m.alpha, wo) * ggxG1(m.alpha, wi);
}
__device__ __forceinline__ V3 mfSampleWh(const DMaterial &m, Rng &r) // src/beckmann.cpp:13-40, src/ggx.cpp:13-24
{
    if (m.distribution == PTC_BECKMANN) {
        const float phi = (float)((double)r.next() * PTC_PI_D * (double)2.f); // phi is drawn first
        const float xi = r.next();
        float logXi = logHost(xi);
        if (isinf(logXi)) { logXi = 0.f; }
        const float tan2 = -m.alpha * m.alpha * logXi;
        const float cosT = 1.f / sqrtf(1.f + tan2);
        const float sinT = sqrtf(fmaxf(0.f, 1.f - (cosT * cosT)));
        return sphToCart(phi, cosT, sinT);
    }
    const float xi1 = r.next(), xi2 = r.next();
    const float theta = atanHost((m.alpha * sqrtf(xi1)) / sqrtf(1.f - xi1));
    const float phi = PTC_TWO_PI_F * xi2;
    float sn, cs;
    sincosHost(theta, sn, cs);
    return sphToCart(phi, cs, sn);
}

__device__ __forceinline__ V3 lambertF(const DMaterial &m, const Isect &i, V3 wiW, float &pdf) // src/lambertian.cpp:16-41
{
    if (dot(i.wo, i.ns) < 0.f || dot(wiW, i.ns) < 0.f) { pdf = 0.f; return mk(0.f, 0.f, 0.f); }
    const V3 wi = normalize(toLocal(i, wiW));
    pdf = wi.y * PTC_INV_PI;
    return lambertAlbedo(m, i) / (float)PTC_PI_D;
}

__device__ __forceinline__ V3 microfacetF(const DMaterial &m, const Isect &i, V3 wiW, float &pdf) // src/microfacet.cpp:12-58
{
    if (dot(i.wo, i.ns) < 0.f || dot(wiW, i.ns) < 0.f) { pdf = 0.f; return mk(0.f, 0.f, 0.f); }
    const V3 wo = normalize(toLocal(i, i.wo)), wi = normalize(toLocal(i, wiW));
    const float cosO = fabsf(wo.y), cosI = fabsf(wi.y);
    const V3 wh = normalize(wo + wi);
    pdf = divIeee(mfPdf(m, wh), 4.f * dot(wo, wh));
    if (cosO == 0.f || cosI == 0.f) { return mk(0.f, 0.f, 0.f); }
    if (wh.x == 0.f && wh.y == 0.f && wh.z == 0.f) { return mk(0.f, 0.f, 0.f); }
    const float F = fresnelDielectric(clampf(dot(wi, wh), 0.f, 1.f), 1.f, 1.5f); // Fresnel fixed at 1 -> 1.5 (Q10)
    const float D = mfD(m, wh);
    const float G = mfG(m, wo, wi);
    const float val = divIeee((1.f * D) * G * F, 4 * cosI * cosO);
    return mk(val, val, val);
}

// Material::f(isect, wi, &pdf).  TYPE >= 0 fixes the material class at compile time (material-sorted shading queues);
// TYPE < 0 dispatches on m.type.
template <int TYPE = -1>
__device__ __forceinline__ V3 bsdfEval(const DMaterial &m, const Isect &i, V3 wiW, float &pdf)
{
    switch (TYPE >= 0 ? TYPE : m.type) {
    case PTC_LAMBERTIAN: return lambertF(m, i, wiW, pdf);
    case PTC_OREN_NAYAR: { // src/oren_nayar.cpp:21-69; back-side cases report pdf = 1 (Q12)
        pdf = 1.f;
        if (dot(i.n, i.wo) < 0.f || dot(i.ns, i.wo) < 0.f) { return mk(0.f, 0.f, 0.f); }
        const V3 lwo = normalize(toLocal(i, i.wo)), lwi = normalize(toLocal(i, wiW));
        if (lwo.y < 0.f || lwi.y < 0.f) { return mk(0.f, 0.f, 0.f); }
        float phiI, thetaI, phiO, thetaO;
        cartToSph(lwi, phiI, thetaI); cartToSph(lwo, phiO, thetaO);
        const float alpha = fmaxf(thetaI, thetaO), beta = fminf(thetaI, thetaO);
        pdf = lwi.y * PTC_INV_PI;
        const float thr = PTC_INV_PI * (m.sigmaA + m.sigmaB * fmaxf(0.f, cosHost(phiI - phiO)) * sinHost(alpha) * tanf(beta));
        return mk(m.diffuse[0] * thr, m.diffuse[1] * thr, m.diffuse[2] * thr);
    }
    case PTC_MICROFACET: return microfacetF(m, i, wiW, pdf);
    case PTC_PLASTIC: { // src/plastic.cpp:19-35: lobes summed, pdfs averaged (Q13)
        float pl, pm;
        const V3 fl = lambertF(m, i, wiW, pl);
        const V3 fm = microfacetF(m, i, wiW, pm);
        pdf = (pl + pm) / 2.f;
        return fl + fm;
    }
    default: pdf = 0.f; return mk(0.f, 0.f, 0.f); // Mirror / Glass: src/mirror.cpp:11-19, src/glass.cpp:20-28
    }
}

struct BsdfSample { V3 wi; float pdf; V3 thr; bool delta; };

__device__ __forceinline__ void lambertSample(const DMaterial &m, const Isect &i, Rng &r, BsdfSample &s) // src/lambertian.cpp:43-58
{
    float unused;
    const V3 l = cosineSample(r);
    s.wi = toWorld(i, l); s.pdf = l.y * PTC_INV_PI; s.thr = lambertF(m, i, s.wi, unused); s.delta = false;
}
__device__ __forceinline__ void microfacetSample(const DMaterial &m, const Isect &i, Rng &r, BsdfSample &s) // src/microfacet.cpp:60-78
{
    float unused;
    const V3 wo = toLocal(i, i.wo);
    const V3 wh = mfSampleWh(m, r);
    s.wi = toWorld(i, reflect(wo, wh));
    s.pdf = divIeee(mfPdf(m, wh), 4.f * dot(wo, wh));
    s.thr = microfacetF(m, i, s.wi, unused); s.delta = false;
}

// Material::sample(isect, random)
template <int TYPE = -1>
__device__ __forceinline__ void bsdfSample(const DMaterial &m, const Isect &i, Rng &r, BsdfSample &s)
{
    switch (TYPE >= 0 ? TYPE : m.type) {
    case PTC_LAMBERTIAN: lambertSample(m, i, r, s); return;
    case PTC_OREN_NAYAR: { // src/oren_nayar.cpp:71-85
        float unused;
        const V3 l = cosineSample(r);
        s.wi = toWorld(i, l); s.pdf = l.y * PTC_INV_PI; s.thr = bsdfEval<PTC_OREN_NAYAR>(m, i, s.wi, unused); s.delta = false;
        return;
    }
    case PTC_MIRROR: { // src/mirror.cpp:21-37
        const V3 lwi = reflect(toLocal(i, i.wo), mk(0.f, 1.f, 0.f));
        const float t = fmaxf(0.f, 1.f / lwi.y);
        s.wi = toWorld(i, lwi); s.pdf = 1.f; s.thr = mk(t, t, t); s.delta = true;
        return;
    }
    case PTC_GLASS: { // src/glass.cpp:30-85 with Snell::refract (src/snell.cpp:9-37)
        const V3 lwo = toLocal(i, i.wo);
        float etaI = 1.f, etaT = m.ior;
        if (lwo.y < 0.f) { etaI = m.ior; etaT = 1.f; }
        V3 normal = mk(0.f, 1.f, 0.f);
        if (lwo.y < 0.f) { normal = normal * -1.f; }
        const V3 incPerp = lwo - (normal * dot(lwo, normal));
        const V3 transPerp = (-incPerp) * (etaI / etaT);
        const float perpLen2 = length(transPerp) * length(transPerp);
        const V3 transPar = normal * -sqrtf(fmaxf(0.f, 1.f - perpLen2));
        const V3 refracted = normalize(transPar + transPerp);
        const float R = fresnelDielectric(fabsf(lwo.y), etaI, etaT);
        s.delta = true;
        if (r.next() < R) {
            const V3 lwi = reflect(lwo, mk(0.f, 1.f, 0.f));
            const float t = R / fabsf(lwi.y);
            s.wi = toWorld(i, lwi); s.pdf = R; s.thr = mk(t, t, t);
        } else { // no eta^2 radiance scaling (Q11); the reference exit(1)s on TIR here, which R == 1 makes unreachable
            const float T = 1.f - R;
            const float t = T / fabsf(refracted.y);
            s.wi = toWorld(i, refracted); s.pdf = T; s.thr = mk(t, t, t);
        }
        return;
    }
    case PTC_MICROFACET: microfacetSample(m, i, r, s); return;
    case PTC_PASSTHROUGH: { // src/passthrough.cpp:26-40: straight on, throughput 1 / |n_s . wo| (cancels the path's cosine)
        const float cosTheta = fabsf(dot(-i.ns, -i.wo));
        const float t = 1.f * (1.f / cosTheta); // Color(1.f) / cosTheta multiplies by the reciprocal (src/color.cpp:127-134)
        s.wi = -i.wo; s.pdf = 1.f; s.thr = mk(t, t, t); s.delta = true;
        return;
    }
    default: { // PTC_PLASTIC, src/plastic.cpp:37-66: xi > 0.5 -> diffuse lobe
        // The two branches of the reference differ only in how wi and the sampled lobe's pdf come about; both then evaluate BOTH lobes for
        // wi (the sampled lobe inside its own sample(), the other one for the sum).  Only the first part is divergent here: the
        // evaluations run once, for the whole warp (same functions, same arguments, same floats).
        const float xi = r.next();
        const bool diffuse = xi > 0.5f;
        float sampledPdf;
        if (diffuse) { // Lambertian::sample, src/lambertian.cpp:43-58
            const V3 l = cosineSample(r);
            s.wi = toWorld(i, l); sampledPdf = l.y * PTC_INV_PI;
        } else {       // Microfacet::sample, src/microfacet.cpp:60-78
            const V3 wo = toLocal(i, i.wo);
            const V3 wh = mfSampleWh(m, r);
            s.wi = toWorld(i, reflect(wo, wh));
            sampledPdf = divIeee(mfPdf(m, wh), 4.f * dot(wo, wh));
        }
        float pl, pm;
        const V3 fl = lambertF(m, i, s.wi, pl);
        const V3 fm = microfacetF(m, i, s.wi, pm);
        s.pdf = (sampledPdf + (diffuse ? pm : pl)) / 2.f;
        s.thr = fl + fm; // diffuse: lambertF + microfacetF, specular: microfacetF + lambertF -- the same sums
        s.delta = false;
        return;
    }
    }
}

// ------------------------------------------------------------------------------------------------ lights
struct SurfSample { V3 point, normal; float invPDF; int measure; }; // include/shape.h:15-20; measure 0 = solid angle, 1 = area

__device__ __forceinline__ float triArea(V3 p0, V3 p1, V3 p2) { return fabsf(length(cross(p1 - p0, p2 - p0)) / 2.f); } // src/triangle.cpp:62-69
__device__ __forceinline__ float uniformConePdf(float cosMax) { return (float)(1.f / (2.f * PTC_PI_D * (double)(1.f - cosMax))); } // src/sphere.cpp:72-75

__device__ __forceinline__ void triSample(const DLight &l, Rng &r, SurfSample &s) // src/triangle.cpp:16-37
{
    const V3 p0 = mk(l.p0[0], l.p0[1], l.p0[2]), p1 = mk(l.p1[0], l.p1[1], l.p1[2]), p2 = mk(l.p2[0], l.p2[1], l.p2[2]);
    const float r1 = r.next(), r2 = r.next();
    const float a = 1 - sqrtf(r1);
    const float b = sqrtf(r1) * (1 - r2);
    const float c = 1 - a - b;
    s.point = ((p0 * a) + (p1 * b)) + (p2 * c);
    s.normal = normalize(cross(p1 - p0, p2 - p0));
    s.invPDF = triArea(p0, p1, p2); s.measure = 1;
}

__device__ void sphereSample(const DLight &l, V3 ref, Rng &r, SurfSample &s) // src/sphere.cpp:54-128
{
    const V3 center = mk(l.centerRadius[0], l.centerRadius[1], l.centerRadius[2]);
    const float radius = l.centerRadius[3];
    const float cd = length(center - ref);
    const float cd2 = cd * cd;
    if (cd <= radius) { // inside: uniform over the area
        const float z = 1 - 2 * r.next();
        const float rr = sqrtf(fmaxf(0.f, 1 - z * z));
        const float phi = (float)(2 * PTC_PI_D * (double)r.next());
        float sn, cs;
        sincosHost(phi, sn, cs);
        const V3 v = mk(rr * cs, rr * sn, z);
        s.point = center + v * radius; s.normal = normalize(v);
        s.invPDF = (float)(4 * PTC_PI_D * (double)radius * (double)radius); s.measure = 1;
        return;
    }
    const float radius2 = radius * radius;
    const float sin2Max = radius * radius / cd2;
    const float cosMax = sqrtf(fmaxf(0.f, 1.f - sin2Max));
    const float xi1 = r.next();
    const float cosTheta = (1.f - xi1) + xi1 * cosMax;
    const float phi = (float)((double)(r.next() * 2.f) * PTC_PI_D);
    const float sinTheta = sqrtf(fmaxf(0.f, 1.f - (cosTheta * cosTheta)));
    const float opp = cd * sinTheta;
    const float helper = sqrtf(fmaxf(0.f, radius * radius - opp * opp));
    const float sd = cd * cosTheta - helper;
    const float sd2 = sd * sd;
    const float cosAlpha = clampf((cd2 + radius2 - sd2) / (2.f * radius * cd), 0.f, 1.f);
    const float sinAlpha = sqrtf(fmaxf(0.f, 1.f - (cosAlpha * cosAlpha)));
    const V3 local = sphToCart(phi, cosAlpha, sinAlpha);
    const V3 nrm = normalize(ref - center);
    V3 xa, za;
    makeFrame(nrm, nrm, xa, za); // single-argument normalToWorldSpace
    V3 world = mk(xa.x * local.x + nrm.x * local.y + za.x * local.z, xa.y * local.x + nrm.y * local.y + za.y * local.z, xa.z * local.x + nrm.z * local.y + za.z * local.z);
    world = normalize(world);
    s.point = center + world * radius; s.normal = normalize(world);
    s.invPDF = 1.f / uniformConePdf(cosMax); s.measure = 0;
}

// radiance arriving from direction `dir`: Scene::environmentL -> EnvironmentLight::emit (src/environment_light.cpp:61-80); nearest texel (Q9)
__device__ V3 envRadiance(const DScene &s, V3 dir)
{
    if (!s.hasEnv) { return mk(0.f, 0.f, 0.f); }
    float phi, theta;
    cartToSph(normalize(xfVec(s.envW2M, dir)), phi, theta);
    const float phiC = clampf(phi / PTC_TWO_PI_F, 0.f, 1.f);
    const float thetaC = clampf((float)((double)theta / PTC_PI_D), 0.f, 1.f);
    const int ps = min((int)floorf(s.envW * phiC), s.envW - 1);
    const int ts = min((int)floorf(s.envH * thetaC), s.envH - 1);
    const float4 px = __ldg(s.envRgba + (size_t)ts * s.envW + ps);
    return mk(px.x * s.envScale, px.y * s.envScale, px.z * s.envScale);
}
__device__ __forceinline__ float cdfPdf(const float *cdf, bool empty, int i) // src/distribution.cpp:56-65
{
    if (empty) { return 0.f; }
    return i == 0 ? __ldg(cdf) : __ldg(cdf + i) - __ldg(cdf + i - 1);
}
// src/distribution.cpp:35-53 scans for the first i with xi <= cdf[i]; the cdf is non-decreasing, so a binary search finds the same i
__device__ __forceinline__ int cdfSample(const float *cdf, int n, float xi, float &pdf)
{
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (xi <= __ldg(cdf + mid)) { hi = mid; } else { lo = mid + 1; }
    }
    pdf = lo > 0 ? __ldg(cdf + lo) - __ldg(cdf + lo - 1) : __ldg(cdf);
    return lo;
}
__device__ __forceinline__ int cdfSampleGuided(const float *cdf, const uint16_t *guide, int G, float xi, float &pdf)
{
    const int g = (int)(xi * (float)G); // exact: G is a power of two, xi < 1
    int lo = __ldg(guide + g), hi = __ldg(guide + g + 1);
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (xi <= __ldg(cdf + mid)) { hi = mid; } else { lo = mid + 1; }
    }
    pdf = lo > 0 ? __ldg(cdf + lo) - __ldg(cdf + lo - 1) : __ldg(cdf);
    return lo;
}
__device__ float envPdf(const DScene &s, V3 dir) // EnvironmentLight::emitPDF, src/environment_light.cpp:117-138
{
    float phi, theta;
    cartToSph(xfVec(s.envW2M, dir), phi, theta);
    const float phiC = phi / PTC_TWO_PI_F;
    const float thetaC = (float)((double)theta / PTC_PI_D);
    const int ps = min((int)floorf(phiC * s.envW), s.envW - 1);
    const int ts = min((int)floorf(thetaC * s.envH), s.envH - 1);
    const float tp = cdfPdf(s.envThetaCdf, s.envThetaEmpty != 0, ts);
    const float pp = cdfPdf(s.envPhiCdf + (size_t)ts * s.envW, __ldg(s.envPhiEmpty + ts) != 0, ps);
    return (float)((double)(tp * pp * s.envW * s.envH) / ((double)(sinHost(theta) * PTC_TWO_PI_F) * PTC_PI_D));
}
__device__ void envSample(const DScene &s, V3 ref, Rng &r, SurfSample &out) // src/environment_light.cpp:82-105
{
    float tp, pp;
    const int ts = cdfSampleGuided(s.envThetaCdf, s.envThetaGuide, s.envThetaG, r.next(), tp);
    const int ps = cdfSampleGuided(s.envPhiCdf + (size_t)ts * s.envW, s.envPhiGuide + (size_t)ts * (s.envPhiG + 1), s.envPhiG, r.next(), pp);
    const float phiC = (ps + 0.5f) / s.envW;
    const float thetaC = (ts + 0.5f) / s.envH;
    const float phi = phiC * PTC_TWO_PI_F;
    const float theta = (float)((double)thetaC * PTC_PI_D);
    float sinT, cosT;
    sincosHost(theta, sinT, cosT);
    const float pdf = (float)((double)(tp * pp * s.envW * s.envH) / ((double)(sinT * PTC_TWO_PI_F) * PTC_PI_D));
    const V3 dir = xfVec(s.envM2W, sphToCart(phi, cosT, sinT));
    out.point = ref + dir * 10000.f; out.normal = dir * -1.f; out.invPDF = 1.f / pdf; out.measure = 0;
}

// Scene::sampleDirectLights, src/scene.cpp:446-467: uniform light choice (Q8), environment light is the last entry
__device__ __forceinline__ const DLight *sampleDirectLights(const DScene &s, V3 ref, Rng &r, SurfSample &out)
{
    const int count = (int)s.nLights;
    const int index = (int)floorf(r.next() * count);
    const DLight *l = s.lights + index;
    const int kind = __ldg(&l->kind);
    if (kind == 0) { triSample(*l, r, out); }
    else if (kind == 1) { sphereSample(*l, ref, r, out); }
    else { envSample(s, ref, r, out); }
    const float choicePDF = 1.f / count;
    out.invPDF = out.invPDF * (1.f / choicePDF);
    return l;
}
__device__ __forceinline__ float solidAnglePdf(const SurfSample &s, V3 ref) // LightSample::solidAnglePDF, include/scene.h:66-80
{
    if (s.measure == 0) { return 1.f / s.invPDF; }
    const V3 ld = s.point - ref;
    const V3 lwo = -normalize(ld);
    const float d = length(ld);
    return (1.f / s.invPDF) * (d * d) / fmaxf(0.f, dot(s.normal, lwo));
}
// Scene::lightsPDF (src/scene.cpp:469-484) for an emitter intersection: Triangle::pdf / Sphere::pdf in solid-angle measure
__device__ float lightsPdf(const DScene &s, V3 ref, const Isect &li)
{
    float m;
    if (li.prim & PTC_SPHERE_FLAG) {
        const float4 cr = __ldg(s.bvh.spheres + (li.prim & ~PTC_SPHERE_FLAG));
        const float cd = length(mk(cr.x, cr.y, cr.z) - ref);
        if (cd <= cr.w) { m = 1.f / (float)(4 * PTC_PI_D * (double)cr.w * (double)cr.w); } // the reference throws here (src/sphere.cpp:137-140)
        else { m = uniformConePdf(sqrtf(fmaxf(0.f, 1.f - cr.w * cr.w / (cd * cd)))); }
    } else {
        const uint4 ix = __ldg(s.prims + li.prim);
        const float4 a = __ldg(s.positions + ix.x), b = __ldg(s.positions + ix.y), c = __ldg(s.positions + ix.z);
        const V3 p0 = mk(a.x, a.y, a.z), p1 = mk(b.x, b.y, b.z), p2 = mk(c.x, c.y, c.z);
        const float areaPDF = 1.f / triArea(p0, p1, p2);
        const V3 normal = normalize(cross(p1 - p0, p2 - p0));
        const V3 sd = ref - li.point; // include/measure.h:13-28
        const float d = length(sd);
        m = areaPDF * (d * d) / fmaxf(0.f, dot(normal, normalize(sd)));
    }
    return m / (int)s.nLights;
}

// ------------------------------------------------------------------------------------------------ camera
__device__ __forceinline__ void cameraRay(const DScene &s, float row, float col, V3 &origin, V3 &direction) // src/camera.cpp:32-47
{
    const float zNear = 0.01f;
    const float height = 2 * tanf(s.vfov / 2) * zNear;
    const float width = height * s.width / s.height;
    const V3 d = normalize(mk(width * (col + 0.5f) / s.width - width / 2.f, height * (row + 0.5f) / s.height - height / 2.f, -zNear));
    const float *m = s.camToWorld;
    origin = mk(m[0] * 0.f + m[1] * 0.f + m[2] * 0.f + m[3], m[4] * 0.f + m[5] * 0.f + m[6] * 0.f + m[7], m[8] * 0.f + m[9] * 0.f + m[10] * 0.f + m[11]);
    direction = xfVec(m, d);
}

// ------------------------------------------------------------------------------------------------ direct lighting
// PathTracer::directSampleLights up to the shadow test (src/path_tracer.cpp:113-165): returns the contribution that
// applies when the shadow ray is unoccluded, plus the shadow ray; false = no shadow ray needed (contribution 0)
template <int TYPE = -1>
__device__ __forceinline__ bool directLightsSetup(const DScene &s, const DMaterial &m, const Isect &i, const BsdfSample &bs, Rng &r,
                                                  V3 &contribution, V3 &shadowDir, float &shadowMaxT)
{
    // a scene without any light: Scene::sampleDirectLights indexes m_lights[0] of an empty vector in the reference (undefined, Q18);
    // defined here as "no direct lighting"
    if (bs.delta || s.nLights == 0u) { return false; }
    SurfSample ls;
    const DLight *light = sampleDirectLights(s, i.point, r, ls);
    const V3 ld = ls.point - i.point;
    const V3 wi = normalize(ld);
    if (dot(ls.normal, wi) >= 0.f) { return false; } // back of the light
    const float pdf = solidAnglePdf(ls, i.point);
    float brdfPDF;
    const V3 f = bsdfEval<TYPE>(m, i, wi, brdfPDF);
    const float w = (1 * pdf) / (1 * pdf + 1 * brdfPDF); // MIS::balanceWeight, include/mis.h:4-7
    const V3 lwo = -normalize(ld);
    const V3 Le = __ldg(&light->kind) == 2 ? envRadiance(s, -lwo) : mk(__ldg(&light->emit[0]), __ldg(&light->emit[1]), __ldg(&light->emit[2]));
    contribution = (((Le * w) * f) * fabsf(dot(i.ns, wi))) / pdf;
    shadowDir = wi; shadowMaxT = length(ld);
    return true;
}

// PathTracer::directSampleBSDF (src/path_tracer.cpp:167-216) given the already traced bounce hit
// frontOnly = false: DirectLightingHelper's copy (src/direct_lighting_helper.cpp:139-187), which has no front-side test
__device__ V3 directBsdf(const DScene &s, V3 point, float cosTheta, V3 wi, float pdf, V3 thr, bool delta, bool hit, const Isect *bi, bool frontOnly = true)
{
    V3 Le; float lightPDF;
    if (hit) {
        const DMaterial &bm = s.materials[bi->material];
        if (!__ldg(&bm.emitter) || (frontOnly && !(dot(bi->wo, bi->ns) >= 0.f))) { return mk(0.f, 0.f, 0.f); }
        Le = mk(__ldg(&bm.emit[0]), __ldg(&bm.emit[1]), __ldg(&bm.emit[2]));
        lightPDF = lightsPdf(s, point, *bi);
    } else {
        Le = envRadiance(s, wi);
        if (isBlack(Le)) { return mk(0.f, 0.f, 0.f); }
        lightPDF = envPdf(s, wi) / (float)s.nLights; // Scene::environmentPDF, src/scene.cpp:494-502
    }
    const float w = delta ? 1.f : (1 * pdf) / (1 * pdf + 1 * lightPDF);
    return (((Le * w) * thr) * cosTheta) / pdf; // cosTheta = |n_s . wi| at the vertex the ray left
}

__device__ __forceinline__ bool checkDone(int last, int b) { return last == -1 ? false : b > last; }                       // src/bounce_controller.cpp:20-25
__device__ __forceinline__ bool checkCounts(int start, int last, int b) { return start > b ? false : !checkDone(last, b); } // :14-18

} // namespace ptc
