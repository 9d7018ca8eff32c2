/*
 * pathed_oracle.c — plain-C restatement of the reference's surface path tracer (see pathed_oracle.h).
 * TEST INFRASTRUCTURE ONLY; parity PINNED against the compiled reference (oracle/_ref).
 *
 * Every function cites the reference file:line it follows (paths relative to the reference root;
 * Embree paths under ext/embree).  All arithmetic is fp32 except where the reference itself
 * promotes to double through M_PI (those spots are marked "double as in the reference").
 * Compile with -ffp-contract=off: the reference is built for baseline x86-64 (no FMA); the only
 * fused operations are the explicit fmaf() calls that mirror Embree's AVX2 madd/msub.
 */
#include "pathed_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define INV_PI 0.3183098861837907f        /* include/util.h:10 */
#define M_TWO_PI_F 6.283185307179586f     /* include/util.h:11 */
#define TNEAR 1e-3f                       /* src/scene.cpp:102 */
#define TFAR 1e5f                         /* src/scene.cpp:103 */

typedef struct { float x, y, z; } v3;

static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, float t) { return V(a.x * t, a.y * t, a.z * t); }
static inline v3 vmulv(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
/* src/vector.cpp:18-21 */
static inline float vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* src/vector.cpp:28-35 */
static inline float vlen(v3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
/* src/vector.cpp:37-44 */
static inline v3 vcross(v3 a, v3 b)
{
    return V((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x));
}
/* src/vector.cpp:46-62 */
static inline v3 vnorm(v3 a)
{
    const float n = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
    return V(a.x / n, a.y / n, a.z / n);
}
/* src/vector.cpp:64-67: (normal * dot(normal) * 2) - this */
static inline v3 vreflect(v3 w, v3 n) { return vsub(vmul(vmul(n, vdot(w, n)), 2.f), w); }
static inline int veq(v3 a, v3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
static inline int black(v3 c) { return c.x == 0.f && c.y == 0.f && c.z == 0.f; } /* src/color.cpp:14-17 */
static inline float clampf(float v, float lo, float hi) { return fminf(hi, fmaxf(v, lo)); } /* include/util.h:39-41 */

/* ------------------------------------------------------------------------------------------ RNG */
/* Replaces RandomGenerator (src/random_generator.cpp:4-11) and std::rand (src/camera.cpp:51-52):
 * Philox4x32-10 (Salmon et al., Random123), key = seed, counter = (pixel, sample, bounce, block). */
void orc_philox4x32_10(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = counter[0], c1 = counter[1], c2 = counter[2], c3 = counter[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; round++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

float orc_uniform(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t d)
{
    const uint32_t counter[4] = {pixel, sample, bounce, d >> 2};
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t out[4];
    orc_philox4x32_10(counter, key, out);
    return (float)(out[d & 3] >> 8) * (1.0f / 16777216.0f); /* [0, 1): strictly below 1 like the reference */
}

typedef struct {
    /* replay mode: sequential draws from xi[] (the reference's consumption order) */
    const float *replay;
    uint32_t replay_count, replay_used;
    /* philox mode */
    uint64_t seed;
    uint32_t pixel, sample, bounce, draw;
} rng_t;

static inline void rng_begin_vertex(rng_t *r, uint32_t bounce)
{
    if (!r->replay) { r->bounce = bounce; r->draw = 0; }
}

static inline float rng_next(rng_t *r)
{
    if (r->replay) {
        const float xi = r->replay[r->replay_used % r->replay_count];
        r->replay_used++;
        return xi;
    }
    return orc_uniform(r->seed, r->pixel, r->sample, r->bounce, r->draw++);
}

/* ------------------------------------------------------------------------------------------ scene */
typedef struct {
    ptc_material_desc d;
    float A, B; /* OrenNayar, src/oren_nayar.cpp:11-19 */
    const unsigned char *tex; int tex_w, tex_h; /* Texture::m_data / m_width / m_height, src/texture.cpp:23-27 */
} material_t;

typedef struct { unsigned char *rgb; int w, h; } texture_t;

typedef struct {
    uint32_t first_vertex, first_prim, n_prims;
    int is_sphere;
    float center_radius[4];
    int medium; /* Surface::m_internalMedium of the geometry's surfaces, stored + 1 (0 = none, so memset-initialised geometries have none) */
    uint32_t scene, local_id; /* the scene the geometry is attached to (0 = root) and its geometry id there (= rtcAttachGeometry's) */
} geom_t;

/* a placement of an instance scene (RTC_GEOMETRY_TYPE_INSTANCE, src/scene_parser.cpp:449-492) */
typedef struct {
    uint32_t scene;      /* the instanced scene */
    uint32_t geom_id;    /* geometry id of the placement inside its owner scene (what ends up in RTCHit::instID) */
    float l2w[12], w2l[12]; /* rows of the 3x4 affine maps; w2l = rcp(l2w) as Instance::setTransform keeps it */
} instance_t;
typedef struct {
    uint32_t n_geoms;                 /* geometry ids handed out in this scene */
    instance_t *insts; uint32_t n_insts;
    uint32_t order_first, order_count; /* this scene's triangle prims in c->order */
    uint32_t root_node; int has_nodes;
} iscene_t;
#define ORC_MAX_SCENE_DEPTH 8

typedef struct { v3 sigma_t, sigma_s; } medium_t; /* HomogeneousMedium, include/homogeneous_medium.h */

typedef struct {
    int kind; /* 0 triangle, 1 sphere, 2 environment */
    v3 p0, p1, p2;
    float center_radius[4];
    v3 emit;
} light_t;

typedef struct { v3 lo, hi; uint32_t left, right, first, count; } bnode_t;

struct orc_ctx {
    char err[256];
    /* flattened geometry */
    float *pos, *nrm, *uv;        /* per vertex: 3, 3, 2 */
    uint32_t n_vertices, cap_vertices;
    uint32_t *idx;                /* per triangle prim: 3 global vertex ids */
    uint32_t *prim_material, *prim_geom, *prim_local;
    uint32_t n_prims, cap_prims;  /* triangle prims only */
    geom_t *geoms; uint32_t n_geoms;
    uint32_t *prim_scene;         /* per triangle prim: the scene it belongs to (0 = root) */
    iscene_t *scenes; uint32_t n_scenes; /* [0] = the root scene (created on first use) */
    uint32_t scene_stack[ORC_MAX_SCENE_DEPTH]; int scene_depth; /* scenes being described: parseInstance recurses */
    uint32_t *root_geoms; uint32_t n_root_geoms;               /* root geometry id -> index into geoms */
    material_t *materials; uint32_t n_materials;
    texture_t *textures; uint32_t n_textures;
    uint32_t *sphere_geoms; uint32_t n_spheres; uint32_t *sphere_material;
    /* lights (src/scene_parser.cpp:173-190) */
    light_t *lights; uint32_t n_lights;
    /* surface -> light lookup for lightsPDF */
    /* environment (src/environment_light.cpp:14-54) */
    int has_env; float *env_rgba; int env_w, env_h; float env_scale;
    float env_m2w[16], env_w2m[16];
    float *env_theta_cdf; float *env_phi_cdf; uint8_t *env_phi_empty; int env_theta_empty;
    /* camera (src/camera.cpp:13-30) */
    int has_camera; float cam_to_world[16]; float vfov; int width, height;
    /* acceleration */
    int committed, brute_force, threads;
    bnode_t *nodes; uint32_t n_nodes; uint32_t *order;
    /* participating media: media, per-prim / per-sphere filter table (medium of a Passthrough surface that encloses
     * one, else -1; src/scene.cpp:42-84), integrator choice (src/job.cpp:66-75) */
    medium_t *media; uint32_t n_media;
    int *prim_event, *sphere_event; int has_filter;
    int integrator;
    /* stats */
    uint64_t closest_rays, shadow_rays, samples;
};

#define FAIL(ctx, code, ...) do { snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__); return (code); } while (0)

int orc_create(int unused, orc_ctx **out)
{
    (void)unused;
    orc_ctx *c = (orc_ctx *)calloc(1, sizeof(orc_ctx));
    if (!c) { return PTC_ERR_NOMEM; }
#ifdef _OPENMP
    c->threads = omp_get_max_threads();
#else
    c->threads = 1;
#endif
    *out = c;
    return PTC_OK;
}

void orc_destroy(orc_ctx *c)
{
    if (!c) { return; }
    free(c->pos); free(c->nrm); free(c->uv); free(c->idx); free(c->prim_material); free(c->prim_geom);
    free(c->prim_local); free(c->prim_scene); free(c->root_geoms);
    for (uint32_t i = 0; i < c->n_scenes; i++) { free(c->scenes[i].insts); }
    free(c->scenes); free(c->geoms); free(c->materials); free(c->sphere_geoms); free(c->sphere_material);
    free(c->lights); free(c->env_rgba); free(c->env_theta_cdf); free(c->env_phi_cdf); free(c->env_phi_empty);
    free(c->nodes); free(c->order);
    for (uint32_t t = 0; t < c->n_textures; t++) { free(c->textures[t].rgb); }
    free(c->textures);
    free(c);
}

const char *orc_last_error(orc_ctx *c) { return c ? c->err : "null context"; }

/* Texture::load, src/texture.cpp:12-32: keeps the 8-bit RGB texels as stbi_load(..., 3) returns them */
int orc_add_texture(orc_ctx *c, const uint8_t *rgb, int width, int height, uint32_t *id)
{
    if (!rgb || width <= 0 || height <= 0) { FAIL(c, PTC_ERR_INVALID, "Error loading texture"); }
    c->textures = (texture_t *)realloc(c->textures, (c->n_textures + 1) * sizeof(texture_t));
    texture_t *t = &c->textures[c->n_textures];
    t->w = width; t->h = height;
    t->rgb = (unsigned char *)malloc((size_t)width * height * 3);
    memcpy(t->rgb, rgb, (size_t)width * height * 3);
    if (id) { *id = c->n_textures; }
    c->n_textures++;
    return PTC_OK;
}

int orc_add_material(orc_ctx *c, const ptc_material_desc *d, uint32_t *id)
{
    if (!d || d->type < 0 || d->type > PTC_PASSTHROUGH) { FAIL(c, PTC_ERR_INVALID, "Unimplemented material"); }
    if (d->albedo_kind == PTC_ALBEDO_TEXTURE && ((d->type != PTC_LAMBERTIAN && d->type != PTC_PLASTIC) || d->texture >= c->n_textures)) {
        FAIL(c, PTC_ERR_INVALID, "bad texture reference");
    }
    c->materials = (material_t *)realloc(c->materials, (c->n_materials + 1) * sizeof(material_t));
    material_t *m = &c->materials[c->n_materials];
    m->d = *d;
    m->tex = NULL; m->tex_w = m->tex_h = 0;
    if (d->albedo_kind == PTC_ALBEDO_TEXTURE) { /* the texel storage itself never moves (malloc'd per texture) */
        m->tex = c->textures[d->texture].rgb; m->tex_w = c->textures[d->texture].w; m->tex_h = c->textures[d->texture].h;
    }
    const float sigma2 = d->sigma * d->sigma;
    m->A = 1.f - (sigma2 / (2.f * (sigma2 + 0.33f)));
    m->B = (0.45f * sigma2) / (sigma2 + 0.09f);
    if (id) { *id = c->n_materials; }
    c->n_materials++;
    return PTC_OK;
}

int orc_add_medium(orc_ctx *c, const float sigma_t[3], const float sigma_s[3], uint32_t *id)
{
    if (c->committed) { FAIL(c, PTC_ERR_STATE, "scene already committed"); }
    c->media = (medium_t *)realloc(c->media, (c->n_media + 1) * sizeof(medium_t));
    c->media[c->n_media].sigma_t = V(sigma_t[0], sigma_t[1], sigma_t[2]);
    c->media[c->n_media].sigma_s = sigma_s ? V(sigma_s[0], sigma_s[1], sigma_s[2]) : V(0, 0, 0);
    if (id) { *id = c->n_media; }
    c->n_media++;
    return PTC_OK;
}

int orc_set_internal_medium(orc_ctx *c, uint32_t geom, uint32_t medium)
{
    if (c->committed) { FAIL(c, PTC_ERR_STATE, "scene already committed"); }
    if (geom >= c->n_root_geoms || c->root_geoms[geom] == PTC_INVALID_ID) { FAIL(c, PTC_ERR_INVALID, "geometry id out of range"); }
    if (medium != PTC_NO_MEDIUM && medium >= c->n_media) { FAIL(c, PTC_ERR_INVALID, "medium id out of range"); }
    c->geoms[c->root_geoms[geom]].medium = medium == PTC_NO_MEDIUM ? 0 : (int)medium + 1;
    return PTC_OK;
}

int orc_set_integrator(orc_ctx *c, int integrator)
{
    if (integrator != PTC_INTEGRATOR_PATH_TRACER && integrator != PTC_INTEGRATOR_VOLUME_PATH_TRACER) { FAIL(c, PTC_ERR_INVALID, "Unimplemented integrator"); }
    c->integrator = integrator;
    return PTC_OK;
}

/* ---- scenes: [0] is the root; ptc_begin_instance opens another one (parseInstance, src/scene_parser.cpp:231-249) */
static void ensure_root_scene(orc_ctx *c)
{
    if (c->n_scenes) { return; }
    c->scenes = (iscene_t *)calloc(1, sizeof(iscene_t));
    c->n_scenes = 1; c->scene_depth = 0;
}
static uint32_t current_scene(orc_ctx *c) { ensure_root_scene(c); return c->scene_depth ? c->scene_stack[c->scene_depth - 1] : 0u; }
/* root geometry id -> index into geoms (PTC_INVALID_ID for an instance placement, which takes an id as well) */
static void root_geom_slot(orc_ctx *c, uint32_t index)
{
    c->root_geoms = (uint32_t *)realloc(c->root_geoms, (c->n_root_geoms + 1) * sizeof(uint32_t));
    c->root_geoms[c->n_root_geoms++] = index;
}
/* hands out the next geometry id of the current scene and records the geometry's place */
static void attach_geometry(orc_ctx *c, geom_t *g, uint32_t index, uint32_t *geom_id)
{
    const uint32_t s = current_scene(c);
    g->scene = s; g->local_id = c->scenes[s].n_geoms++;
    if (s == 0) { root_geom_slot(c, index); }
    if (geom_id) { *geom_id = g->local_id; }
}

int orc_begin_instance(orc_ctx *c, uint32_t *scene_out)
{
    if (c->committed) { FAIL(c, PTC_ERR_STATE, "scene already committed"); }
    ensure_root_scene(c);
    if (c->scene_depth >= ORC_MAX_SCENE_DEPTH) { FAIL(c, PTC_ERR_INVALID, "instance definitions nested too deeply"); }
    c->scenes = (iscene_t *)realloc(c->scenes, (c->n_scenes + 1) * sizeof(iscene_t));
    memset(&c->scenes[c->n_scenes], 0, sizeof(iscene_t));
    c->scene_stack[c->scene_depth++] = c->n_scenes;
    if (scene_out) { *scene_out = c->n_scenes; }
    c->n_scenes++;
    return PTC_OK;
}

int orc_end_instance(orc_ctx *c)
{
    if (c->committed) { FAIL(c, PTC_ERR_STATE, "scene already committed"); }
    if (!c->scene_depth) { FAIL(c, PTC_ERR_STATE, "ptc_end_instance without ptc_begin_instance"); }
    c->scene_depth--;
    return PTC_OK;
}

/* Embree's AffineSpace3fa from a column-major 4x4 (RTC_FORMAT_FLOAT4X4_COLUMN_MAJOR) and its rcp(), which Instance::setTransform
 * (kernels/common/scene_instance.cpp:77-85, part of the lowest-ISA build: SSE2, no fused multiply-add, no dpps) stores as world2local0:
 * common/math/affinespace.h:91 il = rcp(l), p' = -(il * p); linearspace3.h:57-63 il = adjoint(l) / det(l) -- a true division per
 * element (vec3fa.h:187) --, det = dot(vx, cross(vy, vz)) summed as (x + y) + z (vec3fa.h:251-256), il * p = p.x*il.vx + (p.y*il.vy +
 * p.z*il.vz) (linearspace3.h:159 with the non-FMA madd, vec3fa.h:225) */
static void affine_inverse(const float l2w[12], float w2l[12])
{
    /* rows of l2w: r0 = (m00 m01 m02 tx) ...; columns vx = (m00 m10 m20) ... */
    const float vx[3] = {l2w[0], l2w[4], l2w[8]}, vy[3] = {l2w[1], l2w[5], l2w[9]}, vz[3] = {l2w[2], l2w[6], l2w[10]};
    const float px = l2w[3], py = l2w[7], pz = l2w[11];
    const float cyz[3] = {vy[1] * vz[2] - vy[2] * vz[1], vy[2] * vz[0] - vy[0] * vz[2], vy[0] * vz[1] - vy[1] * vz[0]};
    const float czx[3] = {vz[1] * vx[2] - vz[2] * vx[1], vz[2] * vx[0] - vz[0] * vx[2], vz[0] * vx[1] - vz[1] * vx[0]};
    const float cxy[3] = {vx[1] * vy[2] - vx[2] * vy[1], vx[2] * vy[0] - vx[0] * vy[2], vx[0] * vy[1] - vx[1] * vy[0]};
    const float det = (vx[0] * cyz[0] + vx[1] * cyz[1]) + vx[2] * cyz[2];
    /* adjoint = rows (cyz, czx, cxy): inverse rows */
    const float inv[9] = {cyz[0] / det, cyz[1] / det, cyz[2] / det, czx[0] / det, czx[1] / det, czx[2] / det, cxy[0] / det, cxy[1] / det, cxy[2] / det};
    for (int row = 0; row < 3; row++) {
        w2l[4 * row] = inv[3 * row]; w2l[4 * row + 1] = inv[3 * row + 1]; w2l[4 * row + 2] = inv[3 * row + 2];
        w2l[4 * row + 3] = -(px * inv[3 * row] + (py * inv[3 * row + 1] + pz * inv[3 * row + 2]));
    }
}

int orc_add_instance(orc_ctx *c, uint32_t scene, const float m[16], uint32_t *geom_id)
{
    if (c->committed) { FAIL(c, PTC_ERR_STATE, "scene already committed"); }
    ensure_root_scene(c);
    if (!m || scene == 0 || scene >= c->n_scenes) { FAIL(c, PTC_ERR_INVALID, "unknown instance scene %u", scene); }
    for (int d = 0; d < c->scene_depth; d++) { if (c->scene_stack[d] == scene) { FAIL(c, PTC_ERR_INVALID, "an instance scene cannot contain itself"); } }
    const uint32_t owner = current_scene(c);
    iscene_t *sc = &c->scenes[owner];
    sc->insts = (instance_t *)realloc(sc->insts, (sc->n_insts + 1) * sizeof(instance_t));
    instance_t *in = &sc->insts[sc->n_insts++];
    in->scene = scene; in->geom_id = sc->n_geoms++;
    if (owner == 0) { root_geom_slot(c, PTC_INVALID_ID); }
    for (int row = 0; row < 3; row++) { for (int col = 0; col < 4; col++) { in->l2w[4 * row + col] = m[4 * col + row]; } } /* column-major input */
    affine_inverse(in->l2w, in->w2l);
    if (geom_id) { *geom_id = in->geom_id; }
    return PTC_OK;
}

int orc_add_triangle_mesh(orc_ctx *c, const float *P, const float *N, const float *UV, uint32_t nv,
                          const uint32_t *I, const uint32_t *mat, uint32_t nt, uint32_t *geom_id)
{
    if (c->committed) { FAIL(c, PTC_ERR_STATE, "scene already committed"); }
    for (uint32_t t = 0; t < nt; t++) {
        if (mat[t] >= c->n_materials) { FAIL(c, PTC_ERR_INVALID, "material id out of range"); }
        for (int k = 0; k < 3; k++) { if (I[3 * t + k] >= nv) { FAIL(c, PTC_ERR_INVALID, "vertex index out of range"); } }
    }
    c->pos = (float *)realloc(c->pos, (size_t)(c->n_vertices + nv) * 3 * sizeof(float));
    c->nrm = (float *)realloc(c->nrm, (size_t)(c->n_vertices + nv) * 3 * sizeof(float));
    c->uv = (float *)realloc(c->uv, (size_t)(c->n_vertices + nv) * 2 * sizeof(float));
    memcpy(c->pos + 3 * (size_t)c->n_vertices, P, (size_t)nv * 3 * sizeof(float));
    if (N) { memcpy(c->nrm + 3 * (size_t)c->n_vertices, N, (size_t)nv * 3 * sizeof(float)); }
    else { memset(c->nrm + 3 * (size_t)c->n_vertices, 0, (size_t)nv * 3 * sizeof(float)); }
    if (UV) { memcpy(c->uv + 2 * (size_t)c->n_vertices, UV, (size_t)nv * 2 * sizeof(float)); }
    else { memset(c->uv + 2 * (size_t)c->n_vertices, 0, (size_t)nv * 2 * sizeof(float)); }
    c->idx = (uint32_t *)realloc(c->idx, (size_t)(c->n_prims + nt) * 3 * sizeof(uint32_t));
    c->prim_material = (uint32_t *)realloc(c->prim_material, (size_t)(c->n_prims + nt) * sizeof(uint32_t));
    c->prim_geom = (uint32_t *)realloc(c->prim_geom, (size_t)(c->n_prims + nt) * sizeof(uint32_t));
    c->prim_local = (uint32_t *)realloc(c->prim_local, (size_t)(c->n_prims + nt) * sizeof(uint32_t));
    c->prim_scene = (uint32_t *)realloc(c->prim_scene, (size_t)(c->n_prims + nt) * sizeof(uint32_t));
    for (uint32_t t = 0; t < nt; t++) {
        for (int k = 0; k < 3; k++) { c->idx[3 * (size_t)(c->n_prims + t) + k] = I[3 * t + k] + c->n_vertices; }
        c->prim_material[c->n_prims + t] = mat[t];
        c->prim_geom[c->n_prims + t] = c->n_geoms;
        c->prim_local[c->n_prims + t] = t;
        c->prim_scene[c->n_prims + t] = current_scene(c);
    }
    c->geoms = (geom_t *)realloc(c->geoms, (c->n_geoms + 1) * sizeof(geom_t));
    geom_t *g = &c->geoms[c->n_geoms];
    memset(g, 0, sizeof(*g));
    g->first_vertex = c->n_vertices; g->first_prim = c->n_prims; g->n_prims = nt;
    attach_geometry(c, g, c->n_geoms, geom_id);
    c->n_geoms++; c->n_vertices += nv; c->n_prims += nt;
    return PTC_OK;
}

int orc_add_sphere(orc_ctx *c, const float cr[4], uint32_t material, uint32_t *geom_id)
{
    if (c->committed) { FAIL(c, PTC_ERR_STATE, "scene already committed"); }
    if (material >= c->n_materials) { FAIL(c, PTC_ERR_INVALID, "material id out of range"); }
    if (current_scene(c) != 0) { FAIL(c, PTC_ERR_INVALID, "only triangle meshes can be instanced (src/sphere.cpp:46 attaches spheres to the global scene)"); }
    c->geoms = (geom_t *)realloc(c->geoms, (c->n_geoms + 1) * sizeof(geom_t));
    geom_t *g = &c->geoms[c->n_geoms];
    memset(g, 0, sizeof(*g));
    g->is_sphere = 1; g->n_prims = 1; memcpy(g->center_radius, cr, 4 * sizeof(float));
    attach_geometry(c, g, c->n_geoms, NULL);
    c->sphere_geoms = (uint32_t *)realloc(c->sphere_geoms, (c->n_spheres + 1) * sizeof(uint32_t));
    c->sphere_material = (uint32_t *)realloc(c->sphere_material, (c->n_spheres + 1) * sizeof(uint32_t));
    c->sphere_geoms[c->n_spheres] = c->n_geoms; c->sphere_material[c->n_spheres] = material;
    if (geom_id) { *geom_id = g->local_id; }
    c->n_spheres++; c->n_geoms++;
    return PTC_OK;
}

/* src/distribution.cpp:6-33: cdf[i] = values[i]/sum + cdf[i-1]; last forced to 1; empty if sum == 0 */
static int build_cdf(const float *values, int n, float *cdf)
{
    float sum = 0.f;
    for (int i = 0; i < n; i++) { sum += values[i]; }
    if (sum == 0.f) { for (int i = 0; i < n; i++) { cdf[i] = 0.f; } return 1; }
    for (int i = 0; i < n; i++) {
        cdf[i] = values[i] / sum;
        if (i > 0) { cdf[i] += cdf[i - 1]; }
    }
    cdf[n - 1] = 1.f;
    return 0;
}

int orc_set_environment(orc_ctx *c, const float *rgba, int w, int h, float scale, const float m2w[16], const float w2m[16])
{
    if (!rgba || w <= 0 || h <= 0) { FAIL(c, PTC_ERR_INVALID, "bad environment map"); }
    free(c->env_rgba); free(c->env_theta_cdf); free(c->env_phi_cdf); free(c->env_phi_empty);
    c->env_rgba = (float *)malloc((size_t)w * h * 4 * sizeof(float));
    memcpy(c->env_rgba, rgba, (size_t)w * h * 4 * sizeof(float));
    c->env_w = w; c->env_h = h; c->env_scale = scale;
    memcpy(c->env_m2w, m2w, sizeof(c->env_m2w)); memcpy(c->env_w2m, w2m, sizeof(c->env_w2m));
    /* src/environment_light.cpp:29-53: weight = R+G+B (no sin theta), row sums -> theta marginal */
    float *row = (float *)malloc((size_t)w * sizeof(float));
    float *theta = (float *)malloc((size_t)h * sizeof(float));
    c->env_phi_cdf = (float *)malloc((size_t)w * h * sizeof(float));
    c->env_theta_cdf = (float *)malloc((size_t)h * sizeof(float));
    c->env_phi_empty = (uint8_t *)malloc((size_t)h);
    for (int t = 0; t < h; t++) {
        float thetaSum = 0.f;
        for (int p = 0; p < w; p++) {
            const float *px = rgba + 4 * ((size_t)t * w + p);
            float value = 0.f;
            value += px[0]; value += px[1]; value += px[2];
            thetaSum += value;
            row[p] = value;
        }
        c->env_phi_empty[t] = (uint8_t)build_cdf(row, w, c->env_phi_cdf + (size_t)t * w);
        theta[t] = thetaSum;
    }
    c->env_theta_empty = build_cdf(theta, h, c->env_theta_cdf);
    free(row); free(theta);
    c->has_env = 1;
    return PTC_OK;
}

/* src/transform.cpp:138-164 (lookAt) */
int orc_set_camera(orc_ctx *c, const float o[3], const float t[3], const float up[3], float vfov, int w, int h, int flip)
{
    if (w <= 0 || h <= 0) { FAIL(c, PTC_ERR_INVALID, "bad resolution"); }
    const v3 source = V(o[0], o[1], o[2]);
    const v3 dir = vnorm(vsub(source, V(t[0], t[1], t[2])));
    const v3 upv = V(up[0], up[1], up[2]);
    if (veq(dir, upv)) { FAIL(c, PTC_ERR_INVALID, "Look direction cannot equal up vector"); }
    const v3 xa = vnorm(vcross(vnorm(upv), dir));
    const v3 ya = vcross(dir, xa);
    const float sign = flip ? -1.f : 1.f;
    const float m[16] = {
        sign * xa.x, ya.x, dir.x, source.x,
        sign * xa.y, ya.y, dir.y, source.y,
        sign * xa.z, ya.z, dir.z, source.z,
        0.f, 0.f, 0.f, 1.f};
    memcpy(c->cam_to_world, m, sizeof(m));
    c->vfov = vfov; c->width = w; c->height = h; c->has_camera = 1;
    return PTC_OK;
}

/* src/transform.cpp:89-100 (vector) / :61-72 (point) */
static inline v3 xf_vec(const float *m, v3 v)
{
    return V(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z, m[8] * v.x + m[9] * v.y + m[10] * v.z);
}
static inline v3 xf_pnt(const float *m, v3 v)
{
    return V(m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3], m[4] * v.x + m[5] * v.y + m[6] * v.z + m[7],
             m[8] * v.x + m[9] * v.y + m[10] * v.z + m[11]);
}

/* src/camera.cpp:32-47 */
static void camera_ray(const orc_ctx *c, float row, float col, v3 *origin, v3 *direction)
{
    const float zNear = 0.01f;
    const float height = 2 * tanf(c->vfov / 2) * zNear;
    const float width = height * c->width / c->height;
    const v3 d = vnorm(V(width * (col + 0.5f) / c->width - width / 2.f, height * (row + 0.5f) / c->height - height / 2.f, -zNear));
    *origin = xf_pnt(c->cam_to_world, V(0.f, 0.f, 0.f));
    *direction = xf_vec(c->cam_to_world, d);
}

/* ------------------------------------------------------------------------------------------ intersection */
static inline v3 vert(const orc_ctx *c, uint32_t v) { return V(c->pos[3 * (size_t)v], c->pos[3 * (size_t)v + 1], c->pos[3 * (size_t)v + 2]); }

/* Embree's AVX2 Vec3 helpers: common/math/vec3.h:216 (dot = madd chain), :221 (cross = msub) */
static inline float edot(v3 a, v3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
static inline v3 ecross(v3 a, v3 b)
{
    return V(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}

typedef struct { float t, u, v; uint32_t prim; int sphere; v3 ng; uint32_t inst[2]; } rawhit_t; /* prim = global prim or sphere slot; inst = RTCHit::instID */

/* kernels/geometry/triangle_intersector_moeller.h:75-113 (+ :119-127, triangle.h:53-54):
 * stored v0, e1 = v0-v1, e2 = v2-v0, Ng = e2 x e1; accept iff den != 0, U >= 0, V >= 0, U+V <= |den|,
 * |den|*tnear < T <= |den|*tfar; t = T/|den|, u = U/|den|, v = V/|den| */
static inline int tri_test(const orc_ctx *c, uint32_t prim, v3 O, v3 D, float tnear, float tfar, float *t, float *u, float *v, v3 *Ng)
{
    const v3 v0 = vert(c, c->idx[3 * (size_t)prim]);
    const v3 v1 = vert(c, c->idx[3 * (size_t)prim + 1]);
    const v3 v2 = vert(c, c->idx[3 * (size_t)prim + 2]);
    const v3 e1 = vsub(v0, v1), e2 = vsub(v2, v0);
    const v3 ng = ecross(e2, e1);
    const v3 C = vsub(v0, O);
    const v3 R = ecross(C, D);
    const float den = edot(ng, D);
    const float absDen = fabsf(den);
    const float sgn = den < 0.f || (den == 0.f && signbit(den)) ? -1.f : 1.f;
    const float U = edot(R, e2) * sgn;
    const float Vv = edot(R, e1) * sgn;
    if (!(den != 0.f && U >= 0.f && Vv >= 0.f && U + Vv <= absDen)) { return 0; }
    const float T = edot(ng, C) * sgn;
    if (!(absDen * tnear < T && T <= absDen * tfar)) { return 0; }
    *t = T / absDen; *u = U / absDen; *v = Vv / absDen; *Ng = ng;
    return 1;
}

/* kernels/geometry/sphere_intersector.h:67-106 */
static inline int sphere_test(const float cr[4], v3 O, v3 D, float tnear, float tfar, float *t, v3 *Ng)
{
    const float rd2 = 1.f / edot(D, D);
    const v3 c0 = vsub(V(cr[0], cr[1], cr[2]), O);
    const float projC0 = edot(c0, D) * rd2;
    const v3 perp = vsub(c0, vmul(D, projC0));
    const float l2 = edot(perp, perp);
    const float r2 = cr[3] * cr[3];
    if (!(l2 <= r2)) { return 0; }
    float td = sqrtf((r2 - l2) * rd2);
    const float t_in = projC0 - td, t_out = projC0 + td;
    const int valid_in = (t_in > tnear) && (t_in < tfar);
    const int valid_out = !valid_in && (t_out > tnear) && (t_out < tfar);
    if (!valid_in && !valid_out) { return 0; }
    if (valid_in) { td = -1.0f * td; }
    *t = valid_in ? t_in : t_out;
    *Ng = vsub(vmul(D, td), perp);
    return 1;
}

static inline int box_hit(const bnode_t *n, v3 O, v3 inv, float tnear, float tfar)
{
    float t0 = tnear, t1 = tfar;
    const float lo[3] = {n->lo.x, n->lo.y, n->lo.z}, hi[3] = {n->hi.x, n->hi.y, n->hi.z};
    const float o[3] = {O.x, O.y, O.z}, iv[3] = {inv.x, inv.y, inv.z};
    for (int a = 0; a < 3; a++) {
        float ta = (lo[a] - o[a]) * iv[a], tb = (hi[a] - o[a]) * iv[a];
        if (ta > tb) { const float s = ta; ta = tb; tb = s; }
        if (ta != ta || tb != tb) { continue; } /* 0 * inf: ray parallel and on the slab plane -> do not cull */
        tb *= 1.0000005f; ta *= 0.9999995f;
        if (ta > t0) { t0 = ta; }
        if (tb < t1) { t1 = tb; }
    }
    return t0 <= t1;
}

/* Embree's xfmPoint / xfmVector on an AffineSpace3fa (common/math/affinespace.h, AVX2: madd chains), rows of a 3x4 map */
static inline v3 affine_point(const float *m, v3 p)
{
    return V(fmaf(p.x, m[0], fmaf(p.y, m[1], fmaf(p.z, m[2], m[3]))), fmaf(p.x, m[4], fmaf(p.y, m[5], fmaf(p.z, m[6], m[7]))),
             fmaf(p.x, m[8], fmaf(p.y, m[9], fmaf(p.z, m[10], m[11]))));
}
static inline v3 affine_vector(const float *m, v3 d)
{
    return V(fmaf(d.x, m[0], fmaf(d.y, m[1], d.z * m[2])), fmaf(d.x, m[4], fmaf(d.y, m[5], d.z * m[6])), fmaf(d.x, m[8], fmaf(d.y, m[9], d.z * m[10])));
}

typedef struct { int found; float best; rawhit_t h; } trace_state_t;

/* one scene's triangles, then its instance placements: InstanceIntersector1 (kernels/geometry/instance_intersector.cpp:52-109)
 * transforms origin and direction with world2local, keeps tnear / tfar, and traverses the instanced scene */
static int trace_scene(const orc_ctx *c, uint32_t scene, v3 O, v3 D, float tnear, int any, trace_state_t *st, uint32_t inst0, uint32_t inst1, int level)
{
    const iscene_t *sc = c->n_scenes ? &c->scenes[scene] : NULL;
    const uint32_t first = sc ? sc->order_first : 0, count = sc ? sc->order_count : c->n_prims;
    if (c->brute_force || !sc || !sc->has_nodes) {
        for (uint32_t i = 0; i < count; i++) {
            const uint32_t p = c->order ? c->order[first + i] : i;
            float t, u, v; v3 ng;
            if (tri_test(c, p, O, D, tnear, st->best, &t, &u, &v, &ng)) {
                st->best = t; st->h.t = t; st->h.u = u; st->h.v = v; st->h.prim = p; st->h.sphere = 0; st->h.ng = ng; st->found = 1;
                st->h.inst[0] = inst0; st->h.inst[1] = inst1;
                if (any) { return 1; }
            }
        }
    } else {
        const v3 inv = V(1.f / D.x, 1.f / D.y, 1.f / D.z);
        uint32_t stack[128]; int sp = 0; stack[sp++] = sc->root_node;
        while (sp) {
            const bnode_t *n = &c->nodes[stack[--sp]];
            if (!box_hit(n, O, inv, tnear, st->best)) { continue; }
            if (n->count) {
                for (uint32_t i = 0; i < n->count; i++) {
                    const uint32_t p = c->order[n->first + i];
                    float t, u, v; v3 ng;
                    if (tri_test(c, p, O, D, tnear, st->best, &t, &u, &v, &ng)) {
                        /* keep the brute-force tie rule: among equal t the larger prim index wins */
                        if (st->found && t == st->best && !st->h.sphere && p < st->h.prim) { continue; }
                        st->best = t; st->h.t = t; st->h.u = u; st->h.v = v; st->h.prim = p; st->h.sphere = 0; st->h.ng = ng; st->found = 1;
                        st->h.inst[0] = inst0; st->h.inst[1] = inst1;
                        if (any) { return 1; }
                    }
                }
            } else if (sp + 2 <= 128) { stack[sp++] = n->left; stack[sp++] = n->right; }
        }
    }
    if (sc && level < 2) { /* RTC_MAX_INSTANCE_LEVEL_COUNT = 2 */
        for (uint32_t i = 0; i < sc->n_insts; i++) {
            const instance_t *in = &sc->insts[i];
            const v3 Ol = affine_point(in->w2l, O), Dl = affine_vector(in->w2l, D);
            if (trace_scene(c, in->scene, Ol, Dl, tnear, any, st, level == 0 ? in->geom_id : inst0, level == 0 ? PTC_INVALID_ID : in->geom_id, level + 1) && any) { return 1; }
        }
    }
    return st->found;
}

/* closest hit; any != 0 -> first accepted hit (rtcOccluded1 semantics) */
static int trace(const orc_ctx *c, v3 O, v3 D, float tnear, float tfar, int any, rawhit_t *out)
{
    trace_state_t st; memset(&st, 0, sizeof(st));
    st.best = tfar; st.h.inst[0] = st.h.inst[1] = PTC_INVALID_ID;
    if (trace_scene(c, 0, O, D, tnear, any, &st, PTC_INVALID_ID, PTC_INVALID_ID, 0) && any) { *out = st.h; return 1; }
    for (uint32_t s = 0; s < c->n_spheres; s++) {
        float t; v3 ng;
        if (sphere_test(c->geoms[c->sphere_geoms[s]].center_radius, O, D, tnear, st.best, &t, &ng)) {
            st.best = t; st.h.t = t; st.h.u = 0.f; st.h.v = 0.f; st.h.prim = s; st.h.sphere = 1; st.h.ng = ng; st.found = 1;
            st.h.inst[0] = st.h.inst[1] = PTC_INVALID_ID;
            if (any) { *out = st.h; return 1; }
        }
    }
    if (st.found) { *out = st.h; }
    return st.found;
}

/* the reference's Intersection (include/intersection.h:13-56) with the two frame matrices reduced to 3 axes */
typedef struct {
    int hit; float t; v3 point, wo, n, ns; float u, v; uint32_t material; int sphere; uint32_t prim;
    v3 tx, tz; /* tangentToWorld columns x and z (column y = ns) */
} isect_t;

/* src/transform.cpp:182-218 */
static void make_frame(v3 normal, v3 dir, v3 *xAxis, v3 *zAxis)
{
    if (veq(normal, dir)) {
        v3 xa;
        if (fabsf(normal.x) > fabsf(normal.y)) { xa = vnorm(V(-normal.z, 0.f, normal.x)); }
        else { xa = vnorm(V(0.f, -normal.z, normal.y)); }
        *xAxis = xa; *zAxis = vcross(normal, xa);
        return;
    }
    *xAxis = vnorm(vcross(normal, dir));
    *zAxis = vnorm(vcross(normal, *xAxis));
}
static inline v3 to_world(const isect_t *i, v3 l) { return V(i->tx.x * l.x + i->ns.x * l.y + i->tz.x * l.z, i->tx.y * l.x + i->ns.y * l.y + i->tz.y * l.z, i->tx.z * l.x + i->ns.z * l.y + i->tz.z * l.z); }
static inline v3 to_local(const isect_t *i, v3 w) { return V(i->tx.x * w.x + i->tx.y * w.y + i->tx.z * w.z, i->ns.x * w.x + i->ns.y * w.y + i->ns.z * w.z, i->tz.x * w.x + i->tz.y * w.y + i->tz.z * w.z); }

/* Scene::testIntersect, src/scene.cpp:91-223 */
static isect_t make_isect(const orc_ctx *c, v3 O, v3 D, int found, rawhit_t h)
{
    isect_t r; memset(&r, 0, sizeof(r));
    if (!found) { r.t = 3.402823466e+38f; return r; }
    v3 ns = V(0.f, 0.f, 0.f);
    const v3 ng = vnorm(h.ng);
    if (!h.sphere) {
        /* rtcInterpolate0, kernels/common/scene_triangle_mesh.cpp:248-253: madd(w,p0,madd(u,p1,v*p2)) */
        const uint32_t *ix = &c->idx[3 * (size_t)h.prim];
        const float w = 1.0f - h.u - h.v;
        const float *n0 = &c->nrm[3 * (size_t)ix[0]], *n1 = &c->nrm[3 * (size_t)ix[1]], *n2 = &c->nrm[3 * (size_t)ix[2]];
        const float *t0 = &c->uv[2 * (size_t)ix[0]], *t1 = &c->uv[2 * (size_t)ix[1]], *t2 = &c->uv[2 * (size_t)ix[2]];
        ns = V(fmaf(w, n0[0], fmaf(h.u, n1[0], h.v * n2[0])), fmaf(w, n0[1], fmaf(h.u, n1[1], h.v * n2[1])), fmaf(w, n0[2], fmaf(h.u, n1[2], h.v * n2[2])));
        r.u = fmaf(w, t0[0], fmaf(h.u, t1[0], h.v * t2[0]));
        r.v = fmaf(w, t0[1], fmaf(h.u, t1[1], h.v * t2[1]));
        r.material = c->prim_material[h.prim];
    } else {
        r.material = c->sphere_material[h.prim];
    }
    if (vlen(ns) == 0.f) { ns = ng; }
    r.hit = 1; r.t = h.t; r.point = vadd(O, vmul(D, h.t)); r.wo = vneg(D); r.n = ng; r.ns = vnorm(ns);
    r.sphere = h.sphere; r.prim = h.prim;
    make_frame(r.ns, r.wo, &r.tx, &r.tz);
    return r;
}
static isect_t test_intersect(orc_ctx *c, v3 O, v3 D)
{
    rawhit_t h; memset(&h, 0, sizeof(h));
    const int found = trace(c, O, D, TNEAR, TFAR, 0, &h);
    return make_isect(c, O, D, found, h);
}

/* ---- occlusion filter (src/scene.cpp:42-84, registered for both query kinds by src/rtc_manager.cpp:85-92) ---- */
typedef struct { uint32_t count; float t[PTC_MAX_EVENTS]; int medium[PTC_MAX_EVENTS]; } events_t; /* std::vector<VolumeEvent> */

static void event_add(events_t *ev, float t, int medium) /* src/scene.cpp:66-81: an event with the same t is not added twice */
{
    const uint32_t stored = ev->count < PTC_MAX_EVENTS ? ev->count : PTC_MAX_EVENTS;
    for (uint32_t i = 0; i < stored; i++) { if (ev->t[i] == t) { return; } }
    if (ev->count < PTC_MAX_EVENTS) { ev->t[ev->count] = t; ev->medium[ev->count] = medium; }
    ev->count++;
}
static void events_sort(events_t *ev) /* std::sort by t, src/scene.cpp:337-343, :412-418 */
{
    const uint32_t stored = ev->count < PTC_MAX_EVENTS ? ev->count : PTC_MAX_EVENTS;
    for (uint32_t i = 1; i < stored; i++) {
        const float t = ev->t[i]; const int m = ev->medium[i];
        uint32_t j = i;
        while (j > 0 && ev->t[j - 1] > t) { ev->t[j] = ev->t[j - 1]; ev->medium[j] = ev->medium[j - 1]; j--; }
        ev->t[j] = t; ev->medium[j] = m;
    }
}

/* rtcIntersect1 / rtcOccluded1 with shouldIntersectPassthroughs = false: a Passthrough surface that encloses a medium is
 * rejected by the filter and leaves an event.  Brute force over every primitive (the scenes with media are small), triangles
 * first, then spheres -- Embree keeps one acceleration structure per geometry type and visits them in that order.
 * Closest hit: the filter sees every candidate nearer than the closest accepted hit so far, so which containers BEHIND the
 * final hit leave an event depends on Embree's traversal order; this restatement keeps the order-independent part, the
 * containers in front of the final triangle hit.  Any hit: every container in the interval when unoccluded. */
static int trace_filtered(const orc_ctx *c, v3 O, v3 D, float tnear, float tfar, int any, rawhit_t *out, events_t *ev)
{
    int found = 0;
    float best = tfar;
    rawhit_t h; memset(&h, 0, sizeof(h));
    ev->count = 0;
    events_t all; all.count = 0;
    for (uint32_t p = 0; p < c->n_prims && !(any && found); p++) {
        float t, u, v; v3 ng;
        if (!tri_test(c, p, O, D, tnear, any ? tfar : best, &t, &u, &v, &ng)) { continue; }
        if (c->prim_event && c->prim_event[p] >= 0) { event_add(any ? ev : &all, t, c->prim_event[p]); continue; }
        best = t; h.t = t; h.u = u; h.v = v; h.prim = p; h.sphere = 0; h.ng = ng; found = 1;
    }
    if (!any) { /* a brute-force scan meets containers before it knows the final hit: keep those in front of it */
        const uint32_t stored = all.count < PTC_MAX_EVENTS ? all.count : PTC_MAX_EVENTS;
        for (uint32_t i = 0; i < stored; i++) { if (all.t[i] <= best) { event_add(ev, all.t[i], all.medium[i]); } }
        if (all.count > PTC_MAX_EVENTS) { ev->count = all.count; }
    }
    for (uint32_t s = 0; s < c->n_spheres && !(any && found); s++) {
        float t; v3 ng;
        if (!sphere_test(c->geoms[c->sphere_geoms[s]].center_radius, O, D, tnear, best, &t, &ng)) { continue; }
        /* a rejected near side does not make Embree try the far side (sphere_intersector.h:91-92) */
        if (c->sphere_event && c->sphere_event[s] >= 0) { event_add(ev, t, c->sphere_event[s]); continue; }
        best = t; h.t = t; h.u = 0.f; h.v = 0.f; h.prim = s; h.sphere = 1; h.ng = ng; found = 1;
    }
    events_sort(ev);
    if (found) { *out = h; }
    return found;
}

/* Scene::testOcclusion, src/scene.cpp:355-381 (shouldIntersectPassthroughs = false, :369-370) */
static int test_occlusion(orc_ctx *c, v3 O, v3 D, float maxT)
{
    rawhit_t h;
    if (c->has_filter) { events_t ev; return trace_filtered(c, O, D, TNEAR, maxT - 1e-3f, 1, &h, &ev); }
    return trace(c, O, D, TNEAR, maxT - 1e-3f, 1, &h);
}

/* ------------------------------------------------------------------------------------------ BSDFs */
/* include/tangent_frame.h */
static inline float tf_cos2(v3 v) { return v.y * v.y; }
static inline float tf_sin(v3 v) { return sqrtf(fmaxf(0.f, 1.f - tf_cos2(v))); }
static inline float tf_sin2(v3 v) { return 1.f - tf_cos2(v); }
static inline float tf_tan(v3 v) { return tf_sin(v) / v.y; }
static inline float tf_tan2(v3 v) { return tf_sin2(v) / tf_cos2(v); }
static inline float tf_cosphi(v3 v) /* :69-76 */
{
    const float s = tf_sin(v);
    if (s == 0.f) { return 1.f; }
    return clampf(v.x / s, -1.f, 1.f);
}
static inline v3 tf_clamp(v3 v) /* :11-37 */
{
    const float max = 0.9999f;
    if (v.x >= max) { return V(1.f, 0.f, 0.f); }
    if (v.y >= max) { return V(0.f, 1.f, 0.f); }
    if (v.z >= max) { return V(0.f, 0.f, 1.f); }
    if (v.x <= -max) { return V(-1.f, 0.f, 0.f); }
    if (v.y <= -max) { return V(0.f, -1.f, 0.f); }
    if (v.z <= -max) { return V(0.f, 0.f, -1.f); }
    return v;
}
static inline float tf_sinphi(v3 v) /* :83-92 */
{
    const v3 cl = tf_clamp(v);
    const float s = tf_sin(cl);
    if (s == 0.f) { return 0.f; }
    return clampf(cl.z / s, -1.f, 1.f);
}
static inline float tf_cos2phi(v3 v) { return tf_cosphi(v) * tf_cosphi(v); }
static inline float tf_sin2phi(v3 v) { return tf_sinphi(v) * tf_sinphi(v); }

/* src/coordinate.cpp:7-19 */
static void cart_to_sph(v3 c, float *phi, float *theta)
{
    *phi = atan2f(c.z, c.x);
    if (*phi < 0.f) { *phi = (float)((double)*phi + 2 * M_PI); } /* double as in the reference */
    if (*phi == M_TWO_PI_F) { *phi = 0; }
    *theta = acosf(clampf(c.y, -1.f, 1.f));
}
/* src/coordinate.cpp:26-32 */
static v3 sph_to_cart(float phi, float cosTheta, float sinTheta) { return V(sinTheta * cosf(phi), cosTheta, sinTheta * sinf(phi)); }

/* src/monte_carlo.cpp:24-41 */
static v3 cosine_sample(rng_t *r)
{
    const float xi1 = rng_next(r);
    const float rad = sqrtf(xi1);
    const float phi = (float)(2 * M_PI * (double)rng_next(r)); /* double as in the reference */
    return V(rad * cosf(phi), sqrtf(1.f - xi1), rad * sinf(phi));
}

/* src/checkerboard.cpp:9-20 */
static v3 lambert_albedo(const material_t *m, const isect_t *i)
{
    if (m->d.albedo_kind == PTC_ALBEDO_TEXTURE) { /* Texture::lookup, src/texture.cpp:34-49 */
        const float u = i->u - (int)floorf(i->u);
        const float v = 1.f - (i->v - (int)floorf(i->v));
        const int x = (int)roundf(u * (m->tex_w - 1));
        const int y = (int)roundf(v * (m->tex_h - 1));
        const unsigned char *t = m->tex + 3 * ((size_t)y * m->tex_w + x);
        return V(powf(t[0] / 255.f, 2.2f), powf(t[1] / 255.f, 2.2f), powf(t[2] / 255.f, 2.2f));
    }
    if (m->d.albedo_kind == PTC_ALBEDO_CHECKERBOARD) {
        const int ui = (int)floorf(i->u * m->d.checker_resolution[0]);
        const int vi = (int)floorf(i->v * m->d.checker_resolution[1]);
        if (ui % 2 == vi % 2) { return V(m->d.checker_on[0], m->d.checker_on[1], m->d.checker_on[2]); }
        return V(m->d.checker_off[0], m->d.checker_off[1], m->d.checker_off[2]);
    }
    return V(m->d.diffuse[0], m->d.diffuse[1], m->d.diffuse[2]);
}

/* src/fresnel.cpp:30-64 with src/snell.cpp:51-57 */
static float fresnel_dielectric(float cosI, float etaI, float etaT)
{
    const float sinT = (etaI / etaT) * sqrtf(fmaxf(0.f, 1.f - cosI * cosI));
    if (sinT > 1.f) { return 1.f; }
    const float cosT = sqrtf(fmaxf(0.f, 1.f - sinT * sinT));
    const float rPar = (etaT * cosI - etaI * cosT) / (etaT * cosI + etaI * cosT);
    const float rPerp = (etaI * cosI - etaT * cosT) / (etaI * cosI + etaT * cosT);
    return 0.5f * (rPar * rPar + rPerp * rPerp);
}

/* src/beckmann.cpp:45-89, src/ggx.cpp:26-63 */
static float mf_D(const material_t *m, v3 wh)
{
    const float alpha2 = m->d.alpha * m->d.alpha;
    if (m->d.distribution == PTC_BECKMANN) {
        const float tan2 = tf_tan2(wh);
        if (isinf(tan2)) { return 0.f; }
        const float cos2 = tf_cos2(wh);
        const float cos4 = cos2 * cos2;
        const float num = expf(-tan2 * ((tf_cos2phi(wh) / alpha2) + (tf_sin2phi(wh) / alpha2)));
        const float den = (float)(M_PI * (double)alpha2 * (double)cos4); /* double as in the reference */
        return num / den;
    } else {
        const float cos2 = tf_cos2(wh);
        const float cos4 = cos2 * cos2;
        const float tan2 = tf_tan2(wh);
        if (isinf(tan2)) { return 0.f; }
        const float sum = alpha2 + tan2;
        const float den = (float)(M_PI * (double)cos4 * (double)sum * (double)sum); /* double as in the reference */
        return alpha2 / den;
    }
}
static float mf_pdf(const material_t *m, v3 wh) { return mf_D(m, wh) * fabsf(wh.y); }
static float beckmann_lambda(float alpha, v3 w)
{
    const float absTan = fabsf(tf_tan(w));
    if (isinf(absTan)) { return 0.f; }
    const float a_ = sqrtf(tf_cos2phi(w) * alpha * alpha + tf_sin2phi(w) * alpha * alpha);
    const float a = 1.f / (a_ * absTan);
    if (a >= 1.6f) { return 0.f; }
    return (1 - 1.259f * a + 0.396f * a * a) / (3.535f * a + 2.181f * a * a);
}
static float ggx_G1(float alpha, v3 v)
{
    const float tan2 = tf_tan2(v);
    if (isinf(tan2)) { return 0.f; }
    const float s = (1 + alpha * alpha * tan2);
    return 2.f / (1 + sqrtf(s));
}
static float mf_G(const material_t *m, v3 wo, v3 wi)
{
    if (m->d.distribution == PTC_BECKMANN) { return 1.f / (1.f + beckmann_lambda(m->d.alpha, wo) + beckmann_lambda(m->d.alpha, wi)); }
    return ggx_G1(m->d.alpha, wo) * ggx_G1(m->d.alpha, wi);
}
/* src/beckmann.cpp:13-40, src/ggx.cpp:13-24 */
static v3 mf_sample_wh(const material_t *m, rng_t *r)
{
    if (m->d.distribution == PTC_BECKMANN) {
        const float phi = (float)((double)rng_next(r) * M_PI * (double)2.f); /* double as in the reference */
        const float xi = rng_next(r);
        float logXi = logf(xi);
        if (isinf(logXi)) { logXi = 0.f; }
        const float tan2 = -m->d.alpha * m->d.alpha * logXi;
        const float cosT = 1.f / sqrtf(1.f + tan2);
        const float sinT = sqrtf(fmaxf(0.f, 1.f - (cosT * cosT)));
        return sph_to_cart(phi, cosT, sinT);
    } else {
        const float xi1 = rng_next(r), xi2 = rng_next(r);
        const float theta = atanf((m->d.alpha * sqrtf(xi1)) / sqrtf(1.f - xi1));
        const float phi = M_TWO_PI_F * xi2;
        return sph_to_cart(phi, cosf(theta), sinf(theta));
    }
}

static v3 lambert_f(const material_t *m, const isect_t *i, v3 wiW, float *pdf) /* src/lambertian.cpp:16-41 */
{
    if (vdot(i->wo, i->ns) < 0.f) { *pdf = 0.f; return V(0, 0, 0); }
    if (vdot(wiW, i->ns) < 0.f) { *pdf = 0.f; return V(0, 0, 0); }
    const v3 wi = vnorm(to_local(i, wiW));
    *pdf = wi.y * INV_PI;
    const v3 a = lambert_albedo(m, i);
    const float pi = (float)M_PI;
    return V(a.x / pi, a.y / pi, a.z / pi);
}

static v3 microfacet_f(const material_t *m, const isect_t *i, v3 wiW, float *pdf) /* src/microfacet.cpp:12-58 */
{
    const v3 wo = vnorm(to_local(i, i->wo));
    const v3 wi = vnorm(to_local(i, wiW));
    if (vdot(i->wo, i->ns) < 0.f) { *pdf = 0.f; return V(0, 0, 0); }
    if (vdot(wiW, i->ns) < 0.f) { *pdf = 0.f; return V(0, 0, 0); }
    const float cosO = fabsf(wo.y), cosI = fabsf(wi.y);
    const v3 wh = vnorm(vadd(wo, wi));
    *pdf = mf_pdf(m, wh) / (4.f * vdot(wo, wh));
    if (cosO == 0.f || cosI == 0.f) { return V(0, 0, 0); }
    if (wh.x == 0.f && wh.y == 0.f && wh.z == 0.f) { return V(0, 0, 0); }
    const float cosInc = clampf(vdot(wi, wh), 0.f, 1.f);
    const float F = fresnel_dielectric(cosInc, 1.f, 1.5f);
    const float D = mf_D(m, wh);
    const float G = mf_G(m, wo, wi);
    /* albedo(1) * D * G * F / (4 * cosI * cosO), Color ops left to right */
    const float den = 4 * cosI * cosO;
    const float val = ((1.f * D) * G * F) / den;
    return V(val, val, val);
}

/* Material::f dispatch: returns f, writes pdf */
static v3 bsdf_f(const material_t *m, const isect_t *i, v3 wiW, float *pdf)
{
    switch (m->d.type) {
    case PTC_LAMBERTIAN: return lambert_f(m, i, wiW, pdf);
    case PTC_OREN_NAYAR: { /* src/oren_nayar.cpp:21-69 */
        if (vdot(i->n, i->wo) < 0.f) { *pdf = 1.f; return V(0, 0, 0); }
        if (vdot(i->ns, i->wo) < 0.f) { *pdf = 1.f; return V(0, 0, 0); }
        const v3 lwo = vnorm(to_local(i, i->wo)), lwi = vnorm(to_local(i, wiW));
        if (lwo.y < 0.f) { *pdf = 1.f; return V(0, 0, 0); }
        if (lwi.y < 0.f) { *pdf = 1.f; return V(0, 0, 0); }
        float phiI, thetaI, phiO, thetaO;
        cart_to_sph(lwi, &phiI, &thetaI); cart_to_sph(lwo, &phiO, &thetaO);
        const float alpha = fmaxf(thetaI, thetaO), beta = fminf(thetaI, thetaO);
        *pdf = lwi.y * INV_PI;
        const float thr = INV_PI * (m->A + m->B * fmaxf(0.f, cosf(phiI - phiO)) * sinf(alpha) * tanf(beta));
        return V(m->d.diffuse[0] * thr, m->d.diffuse[1] * thr, m->d.diffuse[2] * thr);
    }
    case PTC_MIRROR: case PTC_GLASS: *pdf = 0.f; return V(0, 0, 0); /* src/mirror.cpp:11-19, src/glass.cpp:20-28 */
    case PTC_MICROFACET: return microfacet_f(m, i, wiW, pdf);
    case PTC_PLASTIC: { /* src/plastic.cpp:19-35 */
        float pl, pm;
        const v3 fl = lambert_f(m, i, wiW, &pl);
        const v3 fm = microfacet_f(m, i, wiW, &pm);
        *pdf = (pl + pm) / 2.f;
        return vadd(fl, fm);
    }
    }
    *pdf = 0.f; return V(0, 0, 0);
}

typedef struct { v3 wi; float pdf; v3 thr; int delta; } bsdf_sample_t;

static int is_delta(const material_t *m) { return m->d.type == PTC_MIRROR || m->d.type == PTC_GLASS || m->d.type == PTC_PASSTHROUGH; }

static bsdf_sample_t lambert_sample(const material_t *m, const isect_t *i, rng_t *r) /* src/lambertian.cpp:43-58 */
{
    bsdf_sample_t s; float unused;
    const v3 l = cosine_sample(r);
    s.wi = to_world(i, l); s.pdf = l.y * INV_PI; s.thr = lambert_f(m, i, s.wi, &unused); s.delta = 0;
    return s;
}
static bsdf_sample_t microfacet_sample(const material_t *m, const isect_t *i, rng_t *r) /* src/microfacet.cpp:60-78 */
{
    bsdf_sample_t s; float unused;
    const v3 wo = to_local(i, i->wo);
    const v3 wh = mf_sample_wh(m, r);
    const v3 wi = vreflect(wo, wh);
    s.wi = to_world(i, wi);
    s.pdf = mf_pdf(m, wh) / (4.f * vdot(wo, wh));
    s.thr = microfacet_f(m, i, s.wi, &unused); s.delta = 0;
    return s;
}

static bsdf_sample_t bsdf_sample(const material_t *m, const isect_t *i, rng_t *r)
{
    bsdf_sample_t s; memset(&s, 0, sizeof(s));
    float unused;
    switch (m->d.type) {
    case PTC_LAMBERTIAN: return lambert_sample(m, i, r);
    case PTC_OREN_NAYAR: { /* src/oren_nayar.cpp:71-85 */
        const v3 l = cosine_sample(r);
        s.wi = to_world(i, l); s.pdf = l.y * INV_PI; s.thr = bsdf_f(m, i, s.wi, &unused);
        return s;
    }
    case PTC_MIRROR: { /* src/mirror.cpp:21-37 */
        const v3 lwo = to_local(i, i->wo);
        const v3 lwi = vreflect(lwo, V(0.f, 1.f, 0.f));
        const float t = fmaxf(0.f, 1.f / lwi.y);
        s.wi = to_world(i, lwi); s.pdf = 1.f; s.thr = V(t, t, t); s.delta = 1;
        return s;
    }
    case PTC_GLASS: { /* src/glass.cpp:30-85, src/snell.cpp:9-37 */
        const v3 lwo = to_local(i, i->wo);
        float etaI = 1.f, etaT = m->d.ior;
        if (lwo.y < 0.f) { const float sw = etaI; etaI = etaT; etaT = sw; }
        v3 normal = V(0.f, 1.f, 0.f);
        if (lwo.y < 0.f) { normal = vmul(normal, -1.f); }
        const v3 wIncPerp = vsub(lwo, vmul(normal, vdot(lwo, normal)));
        const v3 wTransPerp = vmul(vneg(wIncPerp), etaI / etaT);
        const float perpLen2 = vlen(wTransPerp) * vlen(wTransPerp);
        const float parLen = sqrtf(fmaxf(0.f, 1.f - perpLen2));
        const v3 wTransPar = vmul(normal, -parLen);
        const v3 refracted = vnorm(vadd(wTransPar, wTransPerp));
        const float R = fresnel_dielectric(fabsf(lwo.y), etaI, etaT);
        s.delta = 1;
        if (rng_next(r) < R) {
            const v3 lwi = vreflect(lwo, V(0.f, 1.f, 0.f));
            const float t = R / fabsf(lwi.y);
            s.wi = to_world(i, lwi); s.pdf = R; s.thr = V(t, t, t);
        } else {
            const float T = 1.f - R;
            const float t = T / fabsf(refracted.y);
            s.wi = to_world(i, refracted); s.pdf = T; s.thr = V(t, t, t);
        }
        return s;
    }
    case PTC_MICROFACET: return microfacet_sample(m, i, r);
    case PTC_PLASTIC: { /* src/plastic.cpp:37-66 */
        const float xi = rng_next(r);
        if (xi > 0.5f) {
            s = lambert_sample(m, i, r);
            float pm; const v3 fm = microfacet_f(m, i, s.wi, &pm);
            s.pdf = (s.pdf + pm) / 2.f; s.thr = vadd(s.thr, fm);
        } else {
            s = microfacet_sample(m, i, r);
            float pl; const v3 fl = lambert_f(m, i, s.wi, &pl);
            s.pdf = (s.pdf + pl) / 2.f; s.thr = vadd(s.thr, fl);
        }
        return s;
    }
    case PTC_PASSTHROUGH: { /* src/passthrough.cpp:26-40; Color(1.f) / cosTheta multiplies by 1 / cosTheta (src/color.cpp:127-134) */
        const float cosTheta = fabsf(vdot(vneg(i->ns), vneg(i->wo)));
        const float t = 1.f * (1.f / cosTheta);
        s.wi = vneg(i->wo); s.pdf = 1.f; s.thr = V(t, t, t); s.delta = 1;
        return s;
    }
    }
    return s;
}

/* ------------------------------------------------------------------------------------------ lights */
typedef struct { v3 point, normal; float invPDF; int measure; /* 0 solid angle, 1 area */ } surf_sample_t;

static float tri_area(const light_t *l) /* src/triangle.cpp:62-69 */
{
    const v3 cr = vcross(vsub(l->p1, l->p0), vsub(l->p2, l->p0));
    return fabsf(vlen(cr) / 2.f);
}
static surf_sample_t tri_sample(const light_t *l, rng_t *r) /* src/triangle.cpp:16-37 */
{
    surf_sample_t s;
    const float r1 = rng_next(r), r2 = rng_next(r);
    const float a = 1 - sqrtf(r1);
    const float b = sqrtf(r1) * (1 - r2);
    const float c = 1 - a - b;
    s.point = vadd(vadd(vmul(l->p0, a), vmul(l->p1, b)), vmul(l->p2, c));
    s.normal = vnorm(vcross(vsub(l->p1, l->p0), vsub(l->p2, l->p0)));
    s.invPDF = tri_area(l); s.measure = 1;
    return s;
}
static float uniform_cone_pdf(float cosThetaMax) { return (float)(1.f / (2.f * M_PI * (double)(1.f - cosThetaMax))); } /* src/sphere.cpp:72-75 */
static surf_sample_t sphere_sample(const light_t *l, v3 ref, rng_t *r) /* src/sphere.cpp:54-128 */
{
    surf_sample_t s;
    const v3 center = V(l->center_radius[0], l->center_radius[1], l->center_radius[2]);
    const float radius = l->center_radius[3];
    const float cd = vlen(vsub(center, ref));
    const float cd2 = cd * cd;
    if (cd <= radius) {
        const float z = 1 - 2 * rng_next(r);
        const float rr = sqrtf(fmaxf(0, 1 - z * z));
        const float phi = (float)(2 * M_PI * (double)rng_next(r));
        const v3 v = V(rr * cosf(phi), rr * sinf(phi), z);
        s.point = vadd(center, vmul(v, radius)); s.normal = vnorm(v);
        s.invPDF = (float)(4 * M_PI * (double)radius * (double)radius); s.measure = 1;
        return s;
    }
    const float radius2 = radius * radius;
    const float sin2Max = radius * radius / cd2;
    const float cosMax = sqrtf(fmaxf(0.f, 1.f - sin2Max));
    const float xi1 = rng_next(r);
    const float cosTheta = (1.f - xi1) + xi1 * cosMax;
    const float phi = (float)((double)(rng_next(r) * 2.f) * M_PI);
    const float sinTheta = sqrtf(fmaxf(0.f, 1.f - (cosTheta * cosTheta)));
    const float opp = cd * sinTheta;
    const float helper = sqrtf(fmaxf(0.f, radius * radius - opp * opp));
    const float sd = cd * cosTheta - helper;
    const float sd2 = sd * sd;
    const float cosAlpha = clampf((cd2 + radius2 - sd2) / (2.f * radius * cd), 0.f, 1.f);
    const float sinAlpha = sqrtf(fmaxf(0.f, 1.f - (cosAlpha * cosAlpha)));
    const v3 local = sph_to_cart(phi, cosAlpha, sinAlpha);
    const v3 nrm = vnorm(vsub(ref, center));
    v3 xa, za; make_frame(nrm, nrm, &xa, &za); /* single-argument normalToWorldSpace */
    v3 world = V(xa.x * local.x + nrm.x * local.y + za.x * local.z, xa.y * local.x + nrm.y * local.y + za.y * local.z, xa.z * local.x + nrm.z * local.y + za.z * local.z);
    world = vnorm(world);
    s.point = vadd(center, vmul(world, radius)); s.normal = vnorm(world);
    s.invPDF = 1.f / uniform_cone_pdf(cosMax); s.measure = 0;
    return s;
}

/* src/environment_light.cpp:61-80 (emit through Scene::environmentL: radiance arriving from `dir`) */
static v3 env_radiance(const orc_ctx *c, v3 dir)
{
    if (!c->has_env) { return V(0, 0, 0); }
    float phi, theta;
    cart_to_sph(vnorm(xf_vec(c->env_w2m, dir)), &phi, &theta);
    const float phiC = clampf(phi / M_TWO_PI_F, 0.f, 1.f);
    const float thetaC = clampf((float)((double)theta / M_PI), 0.f, 1.f);
    int ps = (int)floorf(c->env_w * phiC); if (ps > c->env_w - 1) { ps = c->env_w - 1; }
    int ts = (int)floorf(c->env_h * thetaC); if (ts > c->env_h - 1) { ts = c->env_h - 1; }
    const float *px = c->env_rgba + 4 * ((size_t)ts * c->env_w + ps);
    return V(px[0] * c->env_scale, px[1] * c->env_scale, px[2] * c->env_scale);
}
static float cdf_pdf(const float *cdf, int empty, int i) /* src/distribution.cpp:56-65 */
{
    if (empty) { return 0.f; }
    return i == 0 ? cdf[0] : cdf[i] - cdf[i - 1];
}
static int cdf_sample(const float *cdf, int n, float xi, float *pdf) /* src/distribution.cpp:35-53: first i with xi <= cdf[i] */
{
    for (int i = 0; i < n; i++) {
        if (xi <= cdf[i]) { *pdf = i > 0 ? cdf[i] - cdf[i - 1] : cdf[i]; return i; }
    }
    *pdf = 0.f; return n - 1;
}
static float env_pdf(const orc_ctx *c, v3 dir) /* src/environment_light.cpp:117-138 */
{
    float phi, theta;
    cart_to_sph(xf_vec(c->env_w2m, dir), &phi, &theta);
    const float phiC = phi / M_TWO_PI_F;
    const float thetaC = (float)((double)theta / M_PI);
    int ps = (int)floorf(phiC * c->env_w); if (ps > c->env_w - 1) { ps = c->env_w - 1; }
    int ts = (int)floorf(thetaC * c->env_h); if (ts > c->env_h - 1) { ts = c->env_h - 1; }
    const float tp = cdf_pdf(c->env_theta_cdf, c->env_theta_empty, ts);
    const float pp = cdf_pdf(c->env_phi_cdf + (size_t)ts * c->env_w, c->env_phi_empty[ts], ps);
    return (float)((double)(tp * pp * c->env_w * c->env_h) / ((double)(sinf(theta) * M_TWO_PI_F) * M_PI));
}
static surf_sample_t env_sample(const orc_ctx *c, v3 ref, rng_t *r) /* src/environment_light.cpp:82-105 */
{
    surf_sample_t s;
    float tp, pp;
    const int ts = cdf_sample(c->env_theta_cdf, c->env_h, rng_next(r), &tp);
    const int ps = cdf_sample(c->env_phi_cdf + (size_t)ts * c->env_w, c->env_w, rng_next(r), &pp);
    const float phiC = (ps + 0.5f) / c->env_w;
    const float thetaC = (ts + 0.5f) / c->env_h;
    const float phi = phiC * M_TWO_PI_F;
    const float theta = (float)((double)thetaC * M_PI);
    const float pdf = (float)((double)(tp * pp * c->env_w * c->env_h) / ((double)(sinf(theta) * M_TWO_PI_F) * M_PI));
    const v3 dir = xf_vec(c->env_m2w, sph_to_cart(phi, cosf(theta), sinf(theta)));
    s.point = vadd(ref, vmul(dir, 10000.f)); s.normal = vmul(dir, -1.f); s.invPDF = 1.f / pdf; s.measure = 0;
    return s;
}

typedef struct { surf_sample_t s; const light_t *light; } light_sample_t;

/* Scene::sampleDirectLights, src/scene.cpp:446-467 */
static light_sample_t sample_direct_lights(const orc_ctx *c, v3 ref, rng_t *r)
{
    light_sample_t ls;
    const int count = (int)c->n_lights;
    const int index = (int)floorf(rng_next(r) * count);
    const light_t *l = &c->lights[index];
    if (l->kind == 0) { ls.s = tri_sample(l, r); }
    else if (l->kind == 1) { ls.s = sphere_sample(l, ref, r); }
    else { ls.s = env_sample(c, ref, r); }
    const float choicePDF = 1.f / count;
    ls.s.invPDF = ls.s.invPDF * (1.f / choicePDF);
    ls.light = l;
    return ls;
}
/* LightSample::solidAnglePDF, include/scene.h:66-80 */
static float solid_angle_pdf(const surf_sample_t *s, v3 ref)
{
    if (s->measure == 0) { return 1.f / s->invPDF; }
    const v3 ld = vsub(s->point, ref);
    const v3 lwo = vneg(vnorm(ld));
    const float d = vlen(ld);
    const float d2 = d * d;
    const float proj = fmaxf(0.f, vdot(s->normal, lwo));
    return (1.f / s->invPDF) * d2 / proj;
}
/* Scene::lightsPDF, src/scene.cpp:469-484 -> Triangle::pdf src/triangle.cpp:48-60 / Sphere::pdf src/sphere.cpp:130-149 */
static float lights_pdf(const orc_ctx *c, v3 ref, const isect_t *li)
{
    float m;
    if (li->sphere) {
        const float *cr = c->geoms[c->sphere_geoms[li->prim]].center_radius;
        const float cd = vlen(vsub(V(cr[0], cr[1], cr[2]), ref));
        const float cd2 = cd * cd;
        if (cd <= cr[3]) { m = 1.f / (float)(4 * M_PI * (double)cr[3] * (double)cr[3]); } /* the reference throws here */
        else {
            const float sin2Max = cr[3] * cr[3] / cd2;
            m = uniform_cone_pdf(sqrtf(fmaxf(0.f, 1.f - sin2Max)));
        }
    } else {
        light_t l;
        l.p0 = vert(c, c->idx[3 * (size_t)li->prim]); l.p1 = vert(c, c->idx[3 * (size_t)li->prim + 1]); l.p2 = vert(c, c->idx[3 * (size_t)li->prim + 2]);
        const float areaPDF = 1.f / tri_area(&l);
        const v3 normal = vnorm(vcross(vsub(l.p1, l.p0), vsub(l.p2, l.p0)));
        /* include/measure.h:13-28 */
        const v3 sd = vsub(ref, li->point);
        const v3 swo = vnorm(sd);
        const float d = vlen(sd);
        m = areaPDF * (d * d) / fmaxf(0.f, vdot(normal, swo));
    }
    return m / (int)c->n_lights;
}

/* ------------------------------------------------------------------------------------------ path tracer */
static inline int check_done(int last, int b) { return last == -1 ? 0 : b > last; }              /* src/bounce_controller.cpp:20-25 */
static inline int check_counts(int start, int last, int b) { return start > b ? 0 : !check_done(last, b); } /* :14-18 */

typedef struct { uint64_t closest, shadow; } counts_t;

/* PathTracer::directSampleLights, src/path_tracer.cpp:113-165 */
static v3 direct_lights(orc_ctx *c, const isect_t *i, const bsdf_sample_t *bs, rng_t *r, counts_t *n)
{
    /* no light at all: m_lights[0] of an empty vector in the reference (undefined, Q18); defined as "no direct lighting", as on the device */
    if (bs->delta || c->n_lights == 0) { return V(0, 0, 0); }
    const material_t *m = &c->materials[i->material];
    const light_sample_t ls = sample_direct_lights(c, i->point, r);
    const v3 ld = vsub(ls.s.point, i->point);
    const v3 wi = vnorm(ld);
    if (vdot(ls.s.normal, wi) >= 0.f) { return V(0, 0, 0); }
    const float dist = vlen(ld);
    n->shadow++;
    if (test_occlusion(c, i->point, wi, dist)) { return V(0, 0, 0); }
    const float pdf = solid_angle_pdf(&ls.s, i->point);
    float brdfPDF, unused;
    bsdf_f(m, i, wi, &brdfPDF);
    const float w = (1 * pdf) / (1 * pdf + 1 * brdfPDF); /* include/mis.h:4-7 */
    const v3 lwo = vneg(vnorm(ld));
    const v3 Le = ls.light->kind == 2 ? env_radiance(c, vneg(lwo)) : ls.light->emit;
    const v3 f = bsdf_f(m, i, wi, &unused);
    const float cosT = fabsf(vdot(i->ns, wi));
    v3 out = vmul(Le, w);
    out = vmulv(out, f);
    out = vmul(out, cosT);
    return V(out.x / pdf, out.y / pdf, out.z / pdf);
}

/* PathTracer::directSampleBSDF, src/path_tracer.cpp:167-216, given the already traced bounce intersection */
static v3 direct_bsdf(orc_ctx *c, const isect_t *i, const bsdf_sample_t *bs, const isect_t *bi)
{
    v3 Le; float lightPDF;
    if (bi->hit) {
        const material_t *bm = &c->materials[bi->material];
        Le = V(bm->d.emit[0], bm->d.emit[1], bm->d.emit[2]);
        if (black(Le) || !(vdot(bi->wo, bi->ns) >= 0.f)) { return V(0, 0, 0); }
        lightPDF = lights_pdf(c, i->point, bi);
    } else {
        Le = env_radiance(c, bs->wi);
        if (black(Le)) { return V(0, 0, 0); }
        lightPDF = env_pdf(c, bs->wi) / (float)c->n_lights; /* src/scene.cpp:494-502 */
    }
    const float w = bs->delta ? 1.f : (1 * bs->pdf) / (1 * bs->pdf + 1 * lightPDF);
    v3 out = vmul(Le, w);
    out = vmulv(out, bs->thr);
    out = vmul(out, fabsf(vdot(i->ns, bs->wi)));
    return V(out.x / bs->pdf, out.y / bs->pdf, out.z / bs->pdf);
}

/* ------------------------------------------------------------------------------------------ participating media (N3) */
/* Scene::testVolumetricIntersect, src/scene.cpp:225-353 */
static isect_t test_volumetric_intersect(orc_ctx *c, v3 O, v3 D, events_t *ev)
{
    rawhit_t h; memset(&h, 0, sizeof(h));
    const int found = trace_filtered(c, O, D, TNEAR, TFAR, 0, &h, ev);
    return make_isect(c, O, D, found, h);
}
/* Scene::testVolumetricOcclusion, src/scene.cpp:383-424 */
static int test_volumetric_occlusion(orc_ctx *c, v3 O, v3 D, float maxT, events_t *ev)
{
    rawhit_t h;
    return trace_filtered(c, O, D, TNEAR, maxT - 1e-3f, 1, &h, ev);
}
/* Surface::getInternalMedium of the surface an intersection lies on (-1 none) */
static int internal_medium(const orc_ctx *c, const isect_t *i)
{
    const uint32_t geom = i->sphere ? c->sphere_geoms[i->prim] : c->prim_geom[i->prim];
    return c->geoms[geom].medium - 1;
}
/* HomogeneousMedium::transmittance, src/homogeneous_medium.cpp:13-17 */
static v3 medium_transmittance(const orc_ctx *c, int medium, v3 a, v3 b)
{
    const v3 st = c->media[medium].sigma_t;
    const float d = vlen(vsub(b, a));
    return V(expf(-st.x * d), expf(-st.y * d), expf(-st.z * d));
}
static inline v3 ray_at(v3 O, v3 D, float t) { return vadd(O, vmul(D, t)); }
/* VolumeHelper::rayTransmission, src/volume_helper.cpp:72-123 */
static v3 ray_transmission(const orc_ctx *c, v3 O, v3 D, const events_t *ev, int current)
{
    v3 tr = V(1.f, 1.f, 1.f);
    if (ev->count == 0) { return tr; }
    if (current >= 0) {
        if (ev->count == 1) { tr = vmulv(tr, medium_transmittance(c, current, O, ray_at(O, D, ev->t[0]))); }
        else if (ev->count == 2) { tr = vmulv(tr, medium_transmittance(c, current, ray_at(O, D, ev->t[0]), ray_at(O, D, ev->t[1]))); }
    } else {
        const int m = ev->medium[0];
        if (ev->count == 2) { tr = vmulv(tr, medium_transmittance(c, m, ray_at(O, D, ev->t[0]), ray_at(O, D, ev->t[1]))); }
        else if (ev->count == 1) { tr = vmulv(tr, medium_transmittance(c, m, O, ray_at(O, D, ev->t[0]))); }
    }
    return tr; /* more than two events: asserts are compiled out (F8), nothing is applied */
}
/* VolumeHelper::directSampleLights, src/volume_helper.cpp:12-70 */
static v3 volume_direct_lights(orc_ctx *c, int medium, v3 point, rng_t *r, counts_t *n)
{
    if (c->n_lights == 0) { return V(0, 0, 0); }
    const light_sample_t ls = sample_direct_lights(c, point, r);
    const v3 sd = vsub(ls.s.point, point);
    const v3 wi = vnorm(sd);
    if (vdot(ls.s.normal, wi) >= 0.f) { return V(0, 0, 0); }
    const float dist = vlen(sd);
    events_t ev;
    n->shadow++;
    if (test_volumetric_occlusion(c, point, wi, dist, &ev)) { return V(0, 0, 0); }
    const float pdf = solid_angle_pdf(&ls.s, point);
    const v3 lwo = vneg(vnorm(sd));
    v3 tr = V(0, 0, 0);
    if (ev.count == 1) { tr = medium_transmittance(c, medium, point, ray_at(point, wi, ev.t[0])); }
    else if (ev.count == 2) { tr = medium_transmittance(c, medium, ray_at(point, wi, ev.t[0]), ray_at(point, wi, ev.t[1])); }
    const v3 Le = ls.light->kind == 2 ? env_radiance(c, vneg(lwo)) : ls.light->emit;
    v3 out = vmul(vmulv(Le, tr), 1.f);
    const float fourPi = (float)(4.f * M_PI);
    out = V(out.x / fourPi, out.y / fourPi, out.z / fourPi);
    return V(out.x / pdf, out.y / pdf, out.z / pdf);
}
/* VolumePathTracer::scatter -> HomogeneousMedium::integrate, src/volume_path_tracer.cpp:112-131, src/homogeneous_medium.cpp:36-66 */
static v3 medium_scatter(orc_ctx *c, int medium, v3 entry, v3 exit_, rng_t *r, counts_t *n)
{
    if (medium < 0) { return V(0, 0, 0); }
    const float sigmaT = c->media[medium].sigma_t.x;
    const v3 travel = vsub(exit_, entry);
    const float distance = vlen(travel);
    const float xi = rng_next(r);
    const float sampleT = -logf(1 - xi) / sigmaT;
    if (sampleT >= distance) { return V(0, 0, 0); }
    const v3 samplePoint = ray_at(entry, vnorm(travel), sampleT);
    return volume_direct_lights(c, medium, samplePoint, r, n);
}
/* DirectLightingHelper::Ld, src/direct_lighting_helper.cpp:37-187 */
static v3 volume_ld(orc_ctx *c, const isect_t *i, int medium, const bsdf_sample_t *bs, rng_t *r, counts_t *n)
{
    const material_t *m = &c->materials[i->material];
    if (m->d.type == PTC_PASSTHROUGH) { return V(0, 0, 0); }
    if (!black(V(m->d.emit[0], m->d.emit[1], m->d.emit[2]))) { return V(0, 0, 0); }
    v3 result = V(0, 0, 0);
    if (!bs->delta && c->n_lights != 0) { /* directSampleLights, :75-137 */
        const light_sample_t ls = sample_direct_lights(c, i->point, r);
        const v3 ld = vsub(ls.s.point, i->point);
        const v3 wi = vnorm(ld);
        if (!(vdot(ls.s.normal, wi) >= 0.f)) {
            const float dist = vlen(ld);
            events_t ev;
            n->shadow++;
            if (!test_volumetric_occlusion(c, i->point, wi, dist, &ev)) {
                const v3 tr = ray_transmission(c, i->point, wi, &ev, medium);
                const float pdf = solid_angle_pdf(&ls.s, i->point);
                float brdfPDF;
                const v3 f = bsdf_f(m, i, wi, &brdfPDF);
                const float w = (1 * pdf) / (1 * pdf + 1 * brdfPDF);
                const v3 lwo = vneg(vnorm(ld));
                const v3 Le = ls.light->kind == 2 ? env_radiance(c, vneg(lwo)) : ls.light->emit;
                v3 out = vmulv(Le, tr);
                out = vmul(out, w);
                out = vmulv(out, f);
                out = vmul(out, fabsf(vdot(i->ns, wi)));
                result = vadd(result, V(out.x / pdf, out.y / pdf, out.z / pdf));
            }
        }
    }
    { /* directSampleBSDF, :139-187: no front-side test, no transmittance */
        events_t ev;
        n->closest++;
        const isect_t bi = test_volumetric_intersect(c, i->point, bs->wi, &ev);
        v3 Le; float lightPDF; int contributes = 1;
        if (bi.hit) {
            const material_t *bm = &c->materials[bi.material];
            Le = V(bm->d.emit[0], bm->d.emit[1], bm->d.emit[2]);
            if (black(Le)) { contributes = 0; } else { lightPDF = lights_pdf(c, i->point, &bi); }
        } else {
            Le = env_radiance(c, bs->wi);
            if (black(Le)) { contributes = 0; } else { lightPDF = env_pdf(c, bs->wi) / (float)c->n_lights; }
        }
        if (contributes) {
            const float w = bs->delta ? 1.f : (1 * bs->pdf) / (1 * bs->pdf + 1 * lightPDF);
            v3 out = vmul(Le, w);
            out = vmulv(out, bs->thr);
            out = vmul(out, fabsf(vdot(i->ns, bs->wi)));
            result = vadd(result, V(out.x / bs->pdf, out.y / bs->pdf, out.z / bs->pdf));
        }
    }
    return result;
}
/* the container branch of SampleIntegrator::samplePixel, src/sample_integrator.cpp:35-51 */
static v3 camera_container_term(orc_ctx *c, v3 O, v3 D, counts_t *n)
{
    events_t ev;
    n->closest++;
    const isect_t vi = test_volumetric_intersect(c, O, D, &ev);
    const v3 tr = ray_transmission(c, O, D, &ev, -1);
    if (vi.hit) { const material_t *vm = &c->materials[vi.material]; return vmulv(V(vm->d.emit[0], vm->d.emit[1], vm->d.emit[2]), tr); }
    return vmulv(env_radiance(c, D), tr);
}
/* SampleIntegrator::samplePixel body + VolumePathTracer::L, src/volume_path_tracer.cpp:14-95 */
static v3 volume_radiance(orc_ctx *c, v3 O, v3 D, rng_t *r, int start, int last, counts_t *n)
{
    v3 color = V(0, 0, 0);
    n->closest++;
    isect_t lastI = test_intersect(c, O, D);
    if (!lastI.hit) { return env_radiance(c, D); }
    if (check_counts(start, last, 0)) {
        const material_t *m = &c->materials[lastI.material];
        const v3 emit = V(m->d.emit[0], m->d.emit[1], m->d.emit[2]);
        if (!black(emit) && !(vdot(lastI.n, lastI.wo) < 0.f)) { color = vadd(color, emit); }
        if (m->d.type == PTC_PASSTHROUGH) { color = vadd(color, camera_container_term(c, O, D, n)); }
    }
    int medium = -1;
    rng_begin_vertex(r, 1);
    bsdf_sample_t bs = bsdf_sample(&c->materials[lastI.material], &lastI, r);
    v3 result = V(0, 0, 0);
    if (check_counts(start, last, 1)) { result = volume_ld(c, &lastI, medium, &bs, r, n); }
    v3 modulation = V(1.f, 1.f, 1.f);
    for (int bounce = 2; !check_done(last, bounce); bounce++) {
        if (vdot(lastI.wo, bs.wi) < 0.f) { /* refraction: the medium changes, :42-50 */
            if (vdot(lastI.n, bs.wi) < 0.f) { medium = internal_medium(c, &lastI); }
            else { medium = -1; }
        }
        n->closest++;
        const isect_t bi = test_intersect(c, lastI.point, bs.wi);
        if (!bi.hit) { break; }
        const float invPDF = 1.f / bs.pdf;
        const float cosT = fabsf(vdot(lastI.ns, bs.wi));
        modulation = vmulv(modulation, vmul(vmul(bs.thr, cosT), invPDF));
        rng_begin_vertex(r, (uint32_t)bounce);
        const v3 Ls = medium_scatter(c, medium, lastI.point, bi.point, r, n);
        result = vadd(result, vmulv(Ls, modulation));
        if (medium >= 0) { modulation = vmulv(modulation, medium_transmittance(c, medium, lastI.point, bi.point)); }
        if (black(modulation)) { break; }
        bs = bsdf_sample(&c->materials[bi.material], &bi, r);
        lastI = bi;
        if (check_counts(start, last, bounce)) {
            const v3 Ld = volume_ld(c, &bi, medium, &bs, r, n);
            result = vadd(result, vmulv(Ld, modulation));
        }
    }
    return vadd(color, result);
}

/* SampleIntegrator::samplePixel body (src/sample_integrator.cpp:18-59) + PathTracer::L (src/path_tracer.cpp:19-77).
 * The MIS probe ray (path_tracer.cpp:175) and the continuation ray (:44) are the same ray; traced once (Q6). */
static v3 radiance(orc_ctx *c, v3 O, v3 D, rng_t *r, int start, int last, counts_t *n)
{
    if (c->integrator == PTC_INTEGRATOR_VOLUME_PATH_TRACER) { return volume_radiance(c, O, D, r, start, last, n); }
    v3 color = V(0, 0, 0);
    n->closest++;
    isect_t isect = test_intersect(c, O, D);
    if (!isect.hit) { return vadd(color, env_radiance(c, D)); }
    if (check_counts(start, last, 0)) {
        const material_t *m = &c->materials[isect.material];
        const v3 emit = V(m->d.emit[0], m->d.emit[1], m->d.emit[2]);
        if (!black(emit) && !(vdot(isect.n, isect.wo) < 0.f)) { color = vadd(color, emit); }
        if (c->has_filter && m->d.type == PTC_PASSTHROUGH) { color = vadd(color, camera_container_term(c, O, D, n)); }
    }
    rng_begin_vertex(r, 1);
    bsdf_sample_t bs = bsdf_sample(&c->materials[isect.material], &isect, r);
    v3 result = V(0, 0, 0);
    v3 modulation = V(1.f, 1.f, 1.f);
    int bounce = 1;
    for (;;) {
        /* direct() at the current vertex (bounce), src/path_tracer.cpp:79-111 */
        const int wantDirect = check_counts(start, last, bounce) && black(V(c->materials[isect.material].d.emit[0], c->materials[isect.material].d.emit[1], c->materials[isect.material].d.emit[2]));
        v3 Ld = V(0, 0, 0);
        if (wantDirect) { Ld = vadd(Ld, direct_lights(c, &isect, &bs, r, n)); }
        const int wantNext = !check_done(last, bounce + 1);
        if (!wantDirect && !wantNext) { break; }
        n->closest++;
        isect_t bi = test_intersect(c, isect.point, bs.wi);
        if (wantDirect) {
            Ld = vadd(Ld, direct_bsdf(c, &isect, &bs, &bi));
            result = vadd(result, vmulv(Ld, modulation));
        }
        if (!wantNext) { break; }
        bounce++;
        if (!bi.hit) { break; }
        const float invPDF = 1.f / bs.pdf;
        const float cosT = fabsf(vdot(isect.ns, bs.wi));
        modulation = vmulv(modulation, vmul(vmul(bs.thr, cosT), invPDF));
        if (black(modulation)) { break; }
        rng_begin_vertex(r, (uint32_t)bounce);
        bs = bsdf_sample(&c->materials[bi.material], &bi, r);
        isect = bi;
    }
    return vadd(color, result);
}

/* ------------------------------------------------------------------------------------------ commit */
static void node_bounds(const orc_ctx *c, const uint32_t *order, uint32_t first, uint32_t count, v3 *lo, v3 *hi)
{
    *lo = V(1e30f, 1e30f, 1e30f); *hi = V(-1e30f, -1e30f, -1e30f);
    for (uint32_t i = 0; i < count; i++) {
        for (int k = 0; k < 3; k++) {
            const v3 p = vert(c, c->idx[3 * (size_t)order[first + i] + k]);
            lo->x = fminf(lo->x, p.x); lo->y = fminf(lo->y, p.y); lo->z = fminf(lo->z, p.z);
            hi->x = fmaxf(hi->x, p.x); hi->y = fmaxf(hi->y, p.y); hi->z = fmaxf(hi->z, p.z);
        }
    }
}
static const orc_ctx *g_sort_ctx; static int g_sort_axis;
static int cmp_centroid(const void *a, const void *b)
{
    const orc_ctx *c = g_sort_ctx;
    float ca = 0.f, cb = 0.f;
    for (int k = 0; k < 3; k++) {
        ca += c->pos[3 * (size_t)c->idx[3 * (size_t)(*(const uint32_t *)a) + k] + g_sort_axis];
        cb += c->pos[3 * (size_t)c->idx[3 * (size_t)(*(const uint32_t *)b) + k] + g_sort_axis];
    }
    return (ca > cb) - (ca < cb);
}
/* the oracle's own accelerator: median-split BVH2, padded boxes; only a speed-up for rendering */
static uint32_t build_node(orc_ctx *c, uint32_t first, uint32_t count)
{
    const uint32_t id = c->n_nodes++;
    bnode_t *n = &c->nodes[id];
    v3 lo, hi; node_bounds(c, c->order, first, count, &lo, &hi);
    const float pad = 1e-5f * fmaxf(fmaxf(hi.x - lo.x, hi.y - lo.y), hi.z - lo.z) + 1e-6f;
    n->lo = V(lo.x - pad, lo.y - pad, lo.z - pad); n->hi = V(hi.x + pad, hi.y + pad, hi.z + pad);
    n->first = first; n->count = 0; n->left = n->right = 0;
    if (count <= 4) { n->count = count; return id; }
    const float ex = hi.x - lo.x, ey = hi.y - lo.y, ez = hi.z - lo.z;
    g_sort_ctx = c; g_sort_axis = ex > ey ? (ex > ez ? 0 : 2) : (ey > ez ? 1 : 2);
    qsort(c->order + first, count, sizeof(uint32_t), cmp_centroid);
    const uint32_t half = count / 2;
    const uint32_t l = build_node(c, first, half);
    const uint32_t r = build_node(c, first + half, count - half);
    c->nodes[id].left = l; c->nodes[id].right = r;
    return id;
}

int orc_commit(orc_ctx *c)
{
    if (c->committed) { FAIL(c, PTC_ERR_STATE, "scene already committed"); }
    /* light table: emissive surfaces in registration order, environment last (src/scene_parser.cpp:173-190) */
    uint32_t cap = 0;
    for (uint32_t g = 0; g < c->n_geoms; g++) { cap += c->geoms[g].n_prims; }
    c->lights = (light_t *)malloc((size_t)(cap + 1) * sizeof(light_t));
    c->n_lights = 0;
    uint32_t sphere_slot = 0;
    for (uint32_t g = 0; g < c->n_geoms; g++) {
        const geom_t *ge = &c->geoms[g];
        if (ge->is_sphere) {
            const material_t *m = &c->materials[c->sphere_material[sphere_slot++]];
            const v3 e = V(m->d.emit[0], m->d.emit[1], m->d.emit[2]);
            if (!black(e)) { light_t l; memset(&l, 0, sizeof(l)); l.kind = 1; memcpy(l.center_radius, ge->center_radius, 16); l.emit = e; c->lights[c->n_lights++] = l; }
            continue;
        }
        if (ge->scene != 0) { continue; } /* only root-scene surfaces become lights, src/scene_parser.cpp:173-182 */
        for (uint32_t p = ge->first_prim; p < ge->first_prim + ge->n_prims; p++) {
            const material_t *m = &c->materials[c->prim_material[p]];
            const v3 e = V(m->d.emit[0], m->d.emit[1], m->d.emit[2]);
            if (black(e)) { continue; }
            light_t l; memset(&l, 0, sizeof(l)); l.kind = 0; l.emit = e;
            l.p0 = vert(c, c->idx[3 * (size_t)p]); l.p1 = vert(c, c->idx[3 * (size_t)p + 1]); l.p2 = vert(c, c->idx[3 * (size_t)p + 2]);
            c->lights[c->n_lights++] = l;
        }
    }
    if (c->has_env) { light_t l; memset(&l, 0, sizeof(l)); l.kind = 2; c->lights[c->n_lights++] = l; }
    /* the occlusion filter's table: Passthrough surfaces that enclose a medium (src/scene.cpp:59-63) */
    c->has_filter = 0;
    if (c->n_media) {
        c->prim_event = (int *)malloc(((size_t)c->n_prims + 1) * sizeof(int));
        c->sphere_event = (int *)malloc(((size_t)c->n_spheres + 1) * sizeof(int));
        for (uint32_t p = 0; p < c->n_prims; p++) {
            const int medium = c->geoms[c->prim_geom[p]].medium - 1;
            c->prim_event[p] = (medium >= 0 && c->materials[c->prim_material[p]].d.type == PTC_PASSTHROUGH) ? medium : -1;
            if (c->prim_event[p] >= 0) { c->has_filter = 1; }
        }
        for (uint32_t s = 0; s < c->n_spheres; s++) {
            const int medium = c->geoms[c->sphere_geoms[s]].medium - 1;
            c->sphere_event[s] = (medium >= 0 && c->materials[c->sphere_material[s]].d.type == PTC_PASSTHROUGH) ? medium : -1;
            if (c->sphere_event[s] >= 0) { c->has_filter = 1; }
        }
    }
    ensure_root_scene(c);
    if (c->scene_depth) { FAIL(c, PTC_ERR_STATE, "ptc_begin_instance without ptc_end_instance"); }
    if (c->n_scenes > 1 && c->n_media) { FAIL(c, PTC_ERR_INVALID, "instancing together with participating media is not supported"); }
    if (c->n_prims) {
        /* one BVH per scene (the root and every instance scene) over a partition of c->order */
        c->order = (uint32_t *)malloc((size_t)c->n_prims * sizeof(uint32_t));
        c->nodes = (bnode_t *)malloc((size_t)(2 * c->n_prims + c->n_scenes + 1) * sizeof(bnode_t));
        c->n_nodes = 0;
        uint32_t at = 0;
        for (uint32_t sc = 0; sc < c->n_scenes; sc++) {
            c->scenes[sc].order_first = at;
            for (uint32_t p = 0; p < c->n_prims; p++) { if (c->prim_scene[p] == sc) { c->order[at++] = p; } }
            c->scenes[sc].order_count = at - c->scenes[sc].order_first;
            c->scenes[sc].has_nodes = 0;
            if (c->scenes[sc].order_count) { c->scenes[sc].root_node = build_node(c, c->scenes[sc].order_first, c->scenes[sc].order_count); c->scenes[sc].has_nodes = 1; }
        }
    }
    c->committed = 1;
    return PTC_OK;
}

#define NEED_COMMIT(c) do { if (!(c)->committed) { FAIL(c, PTC_ERR_STATE, "scene not committed"); } } while (0)

/* ------------------------------------------------------------------------------------------ API: queries */
int orc_intersect_instanced(orc_ctx *c, const ptc_ray *rays, uint32_t n, ptc_hit *hits, uint32_t *inst_ids)
{
    NEED_COMMIT(c);
    #pragma omp parallel for schedule(dynamic, 256) num_threads(c->threads)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        rawhit_t h; ptc_hit *o = &hits[i];
        const v3 O = V(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]), D = V(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]);
        if (trace(c, O, D, TNEAR, TFAR, 0, &h)) {
            o->t = h.t; o->u = h.u; o->v = h.v; o->ng[0] = h.ng.x; o->ng[1] = h.ng.y; o->ng[2] = h.ng.z;
            if (h.sphere) { o->geom_id = c->geoms[c->sphere_geoms[h.prim]].local_id; o->prim_id = 0; }
            else { o->geom_id = c->geoms[c->prim_geom[h.prim]].local_id; o->prim_id = c->prim_local[h.prim]; }
            if (inst_ids) { inst_ids[2 * i] = h.inst[0]; inst_ids[2 * i + 1] = h.inst[1]; }
        } else {
            memset(o, 0, sizeof(*o)); o->t = TFAR; o->geom_id = PTC_INVALID_ID; o->prim_id = PTC_INVALID_ID;
            if (inst_ids) { inst_ids[2 * i] = inst_ids[2 * i + 1] = PTC_INVALID_ID; }
        }
    }
    return PTC_OK;
}
int orc_intersect(orc_ctx *c, const ptc_ray *rays, uint32_t n, ptc_hit *hits) { return orc_intersect_instanced(c, rays, n, hits, NULL); }

static void export_isect(const isect_t *s, ptc_isect *o)
{
    memset(o, 0, sizeof(*o));
    o->hit = s->hit; o->t = s->t;
    o->point[0] = s->point.x; o->point[1] = s->point.y; o->point[2] = s->point.z;
    o->wo[0] = s->wo.x; o->wo[1] = s->wo.y; o->wo[2] = s->wo.z;
    o->normal[0] = s->n.x; o->normal[1] = s->n.y; o->normal[2] = s->n.z;
    o->shading_normal[0] = s->ns.x; o->shading_normal[1] = s->ns.y; o->shading_normal[2] = s->ns.z;
    o->uv[0] = s->u; o->uv[1] = s->v; o->material = s->hit ? s->material : PTC_INVALID_ID;
}
static isect_t import_isect(const ptc_isect *p)
{
    isect_t s; memset(&s, 0, sizeof(s));
    s.hit = 1; s.t = p->t; s.point = V(p->point[0], p->point[1], p->point[2]); s.wo = V(p->wo[0], p->wo[1], p->wo[2]);
    s.n = V(p->normal[0], p->normal[1], p->normal[2]); s.ns = V(p->shading_normal[0], p->shading_normal[1], p->shading_normal[2]);
    s.u = p->uv[0]; s.v = p->uv[1]; s.material = p->material;
    make_frame(s.ns, s.wo, &s.tx, &s.tz);
    return s;
}

int orc_intersect_full(orc_ctx *c, const ptc_ray *rays, uint32_t n, ptc_isect *out)
{
    NEED_COMMIT(c);
    #pragma omp parallel for schedule(dynamic, 256) num_threads(c->threads)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const isect_t s = test_intersect(c, V(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]), V(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]));
        export_isect(&s, &out[i]);
    }
    return PTC_OK;
}

static void export_events(const events_t *ev, uint32_t i, uint32_t *n_events, float *event_t, uint32_t *event_medium)
{
    if (n_events) { n_events[i] = ev->count; }
    for (uint32_t e = 0; e < PTC_MAX_EVENTS; e++) {
        const int have = e < ev->count;
        if (event_t) { event_t[(size_t)i * PTC_MAX_EVENTS + e] = have ? ev->t[e] : 0.f; }
        if (event_medium) { event_medium[(size_t)i * PTC_MAX_EVENTS + e] = have ? (uint32_t)ev->medium[e] : PTC_NO_MEDIUM; }
    }
}

int orc_intersect_volumetric(orc_ctx *c, const ptc_ray *rays, uint32_t n, ptc_isect *out, uint32_t *n_events, float *event_t, uint32_t *event_medium)
{
    NEED_COMMIT(c);
    #pragma omp parallel for schedule(dynamic, 256) num_threads(c->threads)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        events_t ev;
        const isect_t s = test_volumetric_intersect(c, V(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]), V(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]), &ev);
        export_isect(&s, &out[i]);
        export_events(&ev, (uint32_t)i, n_events, event_t, event_medium);
    }
    return PTC_OK;
}

int orc_occluded(orc_ctx *c, const ptc_ray *rays, const float *max_t, uint32_t n, uint8_t *occluded)
{
    NEED_COMMIT(c);
    #pragma omp parallel for schedule(dynamic, 256) num_threads(c->threads)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        occluded[i] = (uint8_t)test_occlusion(c, V(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]), V(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]), max_t[i]);
    }
    return PTC_OK;
}

int orc_occluded_volumetric(orc_ctx *c, const ptc_ray *rays, const float *max_t, uint32_t n, uint8_t *occluded, uint32_t *n_events,
                            float *event_t, uint32_t *event_medium)
{
    NEED_COMMIT(c);
    #pragma omp parallel for schedule(dynamic, 256) num_threads(c->threads)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        events_t ev;
        const int occ = test_volumetric_occlusion(c, V(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]), V(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]), max_t[i], &ev);
        if (occ) { ev.count = 0; }
        occluded[i] = (uint8_t)occ;
        export_events(&ev, (uint32_t)i, n_events, event_t, event_medium);
    }
    return PTC_OK;
}

int orc_camera_rays(orc_ctx *c, const float *row_col, uint32_t n, ptc_ray *rays)
{
    if (!c->has_camera) { FAIL(c, PTC_ERR_STATE, "no camera"); }
    for (uint32_t i = 0; i < n; i++) {
        v3 o, d; camera_ray(c, row_col[2 * i], row_col[2 * i + 1], &o, &d);
        rays[i].origin[0] = o.x; rays[i].origin[1] = o.y; rays[i].origin[2] = o.z;
        rays[i].direction[0] = d.x; rays[i].direction[1] = d.y; rays[i].direction[2] = d.z;
    }
    return PTC_OK;
}

int orc_bsdf_eval(orc_ctx *c, uint32_t material, const ptc_isect *isects, const float *wi, uint32_t n, float *f_rgb, float *pdf)
{
    if (material >= c->n_materials) { FAIL(c, PTC_ERR_INVALID, "material id out of range"); }
    for (uint32_t i = 0; i < n; i++) {
        const isect_t s = import_isect(&isects[i]);
        const v3 f = bsdf_f(&c->materials[material], &s, V(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), &pdf[i]);
        f_rgb[3 * i] = f.x; f_rgb[3 * i + 1] = f.y; f_rgb[3 * i + 2] = f.z;
    }
    return PTC_OK;
}

int orc_bsdf_sample(orc_ctx *c, uint32_t material, const ptc_isect *isects, const float *xi, uint32_t n, float *wi, float *pdf, float *thr)
{
    if (material >= c->n_materials) { FAIL(c, PTC_ERR_INVALID, "material id out of range"); }
    for (uint32_t i = 0; i < n; i++) {
        const isect_t s = import_isect(&isects[i]);
        rng_t r; memset(&r, 0, sizeof(r)); r.replay = xi + 3 * i; r.replay_count = 3;
        const bsdf_sample_t b = bsdf_sample(&c->materials[material], &s, &r);
        wi[3 * i] = b.wi.x; wi[3 * i + 1] = b.wi.y; wi[3 * i + 2] = b.wi.z; pdf[i] = b.pdf;
        thr[3 * i] = b.thr.x; thr[3 * i + 1] = b.thr.y; thr[3 * i + 2] = b.thr.z;
    }
    return PTC_OK;
}

int orc_light_sample(orc_ctx *c, const float *ref, const float *xi, uint32_t n, ptc_light_sample_t *out)
{
    NEED_COMMIT(c);
    if (!c->n_lights) { FAIL(c, PTC_ERR_STATE, "scene has no lights"); }
    for (uint32_t i = 0; i < n; i++) {
        const v3 p = V(ref[3 * i], ref[3 * i + 1], ref[3 * i + 2]);
        rng_t r; memset(&r, 0, sizeof(r)); r.replay = xi + 3 * i; r.replay_count = 3;
        const light_sample_t ls = sample_direct_lights(c, p, &r);
        ptc_light_sample_t *o = &out[i];
        o->point[0] = ls.s.point.x; o->point[1] = ls.s.point.y; o->point[2] = ls.s.point.z;
        o->normal[0] = ls.s.normal.x; o->normal[1] = ls.s.normal.y; o->normal[2] = ls.s.normal.z;
        o->inv_pdf = ls.s.invPDF; o->measure = ls.s.measure; o->solid_angle_pdf = solid_angle_pdf(&ls.s, p);
        const v3 lwo = vneg(vnorm(vsub(ls.s.point, p)));
        const v3 e = ls.light->kind == 2 ? env_radiance(c, vneg(lwo)) : ls.light->emit;
        o->emit[0] = e.x; o->emit[1] = e.y; o->emit[2] = e.z;
    }
    return PTC_OK;
}

int orc_light_pdf(orc_ctx *c, const ptc_ray *rays, uint32_t n, float *pdf)
{
    NEED_COMMIT(c);
    for (uint32_t i = 0; i < n; i++) {
        const v3 O = V(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]), D = V(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]);
        const isect_t s = test_intersect(c, O, D);
        if (s.hit) {
            const material_t *m = &c->materials[s.material];
            pdf[i] = black(V(m->d.emit[0], m->d.emit[1], m->d.emit[2])) ? -1.f : lights_pdf(c, O, &s);
        } else if (c->has_env && !black(env_radiance(c, D))) {
            pdf[i] = -2.f - env_pdf(c, D) / (float)c->n_lights;
        } else { pdf[i] = -1.f; }
    }
    return PTC_OK;
}

int orc_environment_radiance(orc_ctx *c, const float *dirs, uint32_t n, float *rgb)
{
    for (uint32_t i = 0; i < n; i++) {
        const v3 e = env_radiance(c, V(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]));
        rgb[3 * i] = e.x; rgb[3 * i + 1] = e.y; rgb[3 * i + 2] = e.z;
    }
    return PTC_OK;
}

int orc_radiance_replay(orc_ctx *c, const ptc_ray *rays, const float *xi, uint32_t stride, uint32_t n, int start, int last, float *rgb)
{
    NEED_COMMIT(c);
    #pragma omp parallel for schedule(dynamic, 64) num_threads(c->threads)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        rng_t r; memset(&r, 0, sizeof(r)); r.replay = xi + (size_t)i * stride; r.replay_count = stride;
        counts_t cn = {0, 0};
        const v3 L = radiance(c, V(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]), V(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]), &r, start, last, &cn);
        rgb[3 * i] = L.x; rgb[3 * i + 1] = L.y; rgb[3 * i + 2] = L.z;
    }
    return PTC_OK;
}

/* ------------------------------------------------------------------------------------------ API: render */
/* n_spp iterations of SampleIntegrator::sampleImage (src/sample_integrator.cpp:80-113): one jittered sample per
 * pixel per iteration, accumulated (+=) in sample order into radianceLookup[3*(row*W+col)+c] */
int orc_render(orc_ctx *c, uint64_t seed, uint32_t first_sample, uint32_t n_spp, int start, int last, float *accum)
{
    NEED_COMMIT(c);
    if (!c->has_camera) { FAIL(c, PTC_ERR_STATE, "no camera"); }
    if (last == -1 || last > PTC_MAX_BOUNCES) { last = PTC_MAX_BOUNCES; }
    uint64_t closest = 0, shadow = 0;
    #pragma omp parallel for schedule(dynamic, 1) num_threads(c->threads) reduction(+ : closest, shadow)
    for (int row = 0; row < c->height; row++) {
        for (int col = 0; col < c->width; col++) {
            const uint32_t pixel = (uint32_t)(row * c->width + col);
            for (uint32_t s = first_sample; s < first_sample + n_spp; s++) {
                rng_t r; memset(&r, 0, sizeof(r));
                r.seed = seed; r.pixel = pixel; r.sample = s;
                rng_begin_vertex(&r, 0);
                /* src/camera.cpp:49-55: box-filter jitter in [-0.5, 0.5) */
                const float jitterX = rng_next(&r) - 0.5f;
                const float jitterY = rng_next(&r) - 0.5f;
                v3 O, D; camera_ray(c, row + jitterY, col + jitterX, &O, &D);
                counts_t cn = {0, 0};
                const v3 L = radiance(c, O, D, &r, start, last, &cn);
                accum[3 * (size_t)pixel + 0] += L.x; accum[3 * (size_t)pixel + 1] += L.y; accum[3 * (size_t)pixel + 2] += L.z;
                closest += cn.closest; shadow += cn.shadow;
            }
        }
    }
    c->closest_rays += closest; c->shadow_rays += shadow; c->samples += (uint64_t)c->width * c->height * n_spp;
    return PTC_OK;
}

int orc_num_lights(orc_ctx *c, uint32_t *out) { NEED_COMMIT(c); *out = c->n_lights; return PTC_OK; }

int orc_get_stats(orc_ctx *c, ptc_stats *out)
{
    memset(out, 0, sizeof(*out));
    out->closest_rays = c->closest_rays; out->shadow_rays = c->shadow_rays; out->samples = c->samples;
    out->bvh_nodes = c->n_nodes; out->bvh_triangles = c->n_prims;
    return PTC_OK;
}
int orc_reset_stats(orc_ctx *c) { c->closest_rays = c->shadow_rays = c->samples = 0; return PTC_OK; }

int orc_set_option(orc_ctx *c, const char *name, int64_t value)
{
    if (!strcmp(name, "brute_force")) { c->brute_force = value != 0; return PTC_OK; }
    if (!strcmp(name, "threads")) { c->threads = value > 0 ? (int)value : 1; return PTC_OK; }
    FAIL(c, PTC_ERR_INVALID, "unknown option %s", name);
}
