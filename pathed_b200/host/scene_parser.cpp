#include "scene_parser.hpp"

#include "exr_io.hpp"
#include "image_loader.hpp"
#include "json.hpp"
#include "obj_parser.hpp"
#include "transform.hpp"

#include <cmath>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <sstream>
#include <stdexcept>

namespace pathed {

namespace {

// All scene scalars are JSON *strings* parsed with stof (src/scene_parser.cpp:814-835)
float parseFloat(const Json &j) { return std::stof(j.asString()); }
float parseFloat(const Json &j, float fallback)
{
    try { return parseFloat(j); } catch (const std::runtime_error &) { return fallback; }
}
bool checkFloat(const Json &j, float *value)
{
    try { *value = parseFloat(j); return true; } catch (const std::runtime_error &) { return false; }
}
bool parseBool(const Json &j, bool fallback) { return j.isBool() ? j.asBool() : fallback; }
std::string parseString(const Json &j, const std::string &fallback) { return j.isString() ? j.asString() : fallback; }
void parseTriple(const Json &j, float out[3])
{
    for (int i = 0; i < 3; i++) { out[i] = std::stof(j[(size_t)i].asString()); }
}
bool parseColor(const Json &j, float out[3], float fallback)
{
    if (j.isArray()) { parseTriple(j, out); return true; }
    out[0] = out[1] = out[2] = fallback;
    return false;
}

std::string resolve(const std::string &root, const std::string &path)
{
    if (root.empty() || (!path.empty() && path[0] == '/')) { return path; }
    return root + "/" + path;
}

// src/scene_parser.cpp:716-812: scale, then rotations (Z,X,Y or legacy X,Y,Z), then translate, with the
// analytic inverse built in the opposite order
Transform parseTransform(const Json &j)
{
    const bool legacy = parseBool(j["legacy"], false);
    float sx = 1.f, sy = 1.f, sz = 1.f, rx = 0.f, ry = 0.f, rz = 0.f, tx = 0.f, ty = 0.f, tz = 0.f;
    if (j["scale"].isArray()) { sx = parseFloat(j["scale"][0]); sy = parseFloat(j["scale"][1]); sz = parseFloat(j["scale"][2]); }
    if (j["rotate"].isArray()) {
        // float * double / float, as the reference writes it
        rx = (float)(parseFloat(j["rotate"][0]) * M_PI / 180.f);
        ry = (float)(parseFloat(j["rotate"][1]) * M_PI / 180.f);
        rz = (float)(parseFloat(j["rotate"][2]) * M_PI / 180.f);
        if (legacy) { rx *= -1; rz *= -1; } else { ry *= -1; }
    }
    if (j["translate"].isArray()) { tx = parseFloat(j["translate"][0]); ty = parseFloat(j["translate"][1]); tz = parseFloat(j["translate"][2]); }

    Transform t;
    Transform::scale(t.m, sx, sy, sz);
    if (legacy) { Transform::rotateX(t.m, rx); Transform::rotateY(t.m, ry); Transform::rotateZ(t.m, rz); }
    else { Transform::rotateZ(t.m, rz); Transform::rotateX(t.m, rx); Transform::rotateY(t.m, ry); }
    Transform::translate(t.m, tx, ty, tz);

    Transform::translate(t.inv, -tx, -ty, -tz);
    if (legacy) { Transform::rotateZ(t.inv, -rz); Transform::rotateY(t.inv, -ry); Transform::rotateX(t.inv, -rx); }
    else { Transform::rotateY(t.inv, -ry); Transform::rotateX(t.inv, -rx); Transform::rotateZ(t.inv, -rz); }
    Transform::scale(t.inv, 1.f / sx, 1.f / sy, 1.f / sz);
    return t;
}
Transform parseTransformOrIdentity(const Json &j) { return j.isObject() ? parseTransform(j) : Transform(); }

void parseDistribution(const Json &j, ptc_material_desc &d) // src/scene_parser.cpp:702-714
{
    d.alpha = parseFloat(j["alpha"]);
    const std::string type = parseString(j["type"], "<missing>");
    if (type == "beckmann") { d.distribution = PTC_BECKMANN; }
    else if (type == "ggx") { d.distribution = PTC_GGX; }
    else { throw std::runtime_error("Unimplemented distribution: " + type); }
}

thread_local std::string g_sceneRoot; // directory that repo-relative asset paths resolve against (the reference uses the cwd)

// std::make_shared<Texture>(path) + load(), src/scene_parser.cpp:630-632, 646-648: every mention decodes the file again in the
// reference; here a file is decoded once and shared.  Paths are relative to the working directory (the scene root).
uint32_t parseTexture(const std::string &path, SceneDescription &scene)
{
    for (size_t t = 0; t < scene.textures.size(); t++) { if (scene.textures[t].filename == path) { return (uint32_t)t; } }
    TextureDesc texture;
    texture.filename = path;
    loadImageRGB8(resolve(g_sceneRoot, path), texture.rgb, texture.width, texture.height);
    scene.textures.push_back(std::move(texture));
    return (uint32_t)scene.textures.size() - 1;
}

// parseMaterial, src/scene_parser.cpp:604-700.  Returns -1 for "no bsdf given" (nullptr in the reference).
int parseMaterial(const Json &j, const MaterialMap &lookup, SceneDescription &scene)
{
    if (!j.isObject()) { return -1; }
    const std::string type = parseString(j["type"], "<missing>");
    if (type == "reference") {
        auto it = lookup.find(j["name"].asString());
        if (it == lookup.end()) { throw std::runtime_error("unknown material reference: " + j["name"].asString()); }
        return (int)it->second;
    }
    ptc_material_desc d;
    memset(&d, 0, sizeof(d));
    d.ior = 1.4f; // Glass::Glass(), src/glass.cpp:16-18
    if (type == "mirror") { d.type = PTC_MIRROR; }
    else if (type == "glass") {
        d.type = PTC_GLASS;
        float ior;
        if (checkFloat(j["ior"], &ior)) { d.ior = ior; } // "alpha" / "diffuseReflectance" ignored (Q11)
    } else if (type == "oren-nayar") {
        d.type = PTC_OREN_NAYAR;
        parseColor(j["diffuseReflectance"], d.diffuse, 1.f);
        d.sigma = parseFloat(j["sigma"]);
    } else if (type == "microfacet") {
        d.type = PTC_MICROFACET;
        parseDistribution(j["distribution"], d);
    } else if (type == "plastic") {
        d.type = PTC_PLASTIC;
        parseColor(j["diffuseReflectance"], d.diffuse, 0.f);
        parseDistribution(j["distribution"], d);
        if (j["texture"].isString()) { // Plastic(Lambertian(texture, 0), distribution): diffuseReflectance is ignored, src/scene_parser.cpp:629-639
            d.albedo_kind = PTC_ALBEDO_TEXTURE;
            d.texture = parseTexture(j["texture"].asString(), scene);
        }
    } else if (type == "lambertian") {
        d.type = PTC_LAMBERTIAN;
        parseColor(j["diffuseReflectance"], d.diffuse, 0.f);
        parseColor(j["emit"], d.emit, 0.f);
        if (j["texture"].isString()) { // the texture wins over "albedo", src/scene_parser.cpp:645-651
            d.albedo_kind = PTC_ALBEDO_TEXTURE;
            d.texture = parseTexture(j["texture"].asString(), scene);
        } else if (j["albedo"].isObject() && parseString(j["albedo"]["type"], "") == "checkerboard") {
            const Json &albedo = j["albedo"];
            d.albedo_kind = PTC_ALBEDO_CHECKERBOARD;
            parseColor(albedo["onColor"], d.checker_on, 0.f);
            parseColor(albedo["offColor"], d.checker_off, 0.f);
            d.checker_resolution[0] = std::stof(albedo["resolution"]["u"].asString());
            d.checker_resolution[1] = std::stof(albedo["resolution"]["v"].asString());
        }
    } else if (type == "passthrough") { // src/scene_parser.cpp:593-594: the container material of a participating medium
        d.type = PTC_PASSTHROUGH;
    } else {
        // phong / disney / ptex / perfect-transmission exist in the reference but are outside
        // the surface path-tracing hot path (SURVEY §2.1 "BSDFs (other)")
        throw std::runtime_error("Unimplemented material: " + type);
    }
    scene.materials.push_back(d);
    return (int)scene.materials.size() - 1;
}

GeometryDesc makeQuad(const Transform &transform, uint32_t material, bool zUp) // Quad::parse, src/quad.cpp:27-151
{
    static const float yUp[6][3] = {{-1, 0, -1}, {-1, 0, 1}, {1, 0, -1}, {-1, 0, 1}, {1, 0, 1}, {1, 0, -1}};
    static const float zUpPoints[6][3] = {{-1, -1, 0}, {1, -1, 0}, {-1, 1, 0}, {-1, 1, 0}, {1, -1, 0}, {1, 1, 0}};
    static const float uv[6][2] = {{0, 0}, {1, 0}, {0, 1}, {0, 1}, {1, 0}, {1, 1}};
    GeometryDesc g;
    const Vec3 normal = transform.applyVector(zUp ? Vec3(0.f, 0.f, 1.f) : Vec3(0.f, 1.f, 0.f)).normalized();
    for (int i = 0; i < 6; i++) {
        const float *p = zUp ? zUpPoints[i] : yUp[i];
        const Vec3 q = transform.applyPoint(Vec3(p[0], p[1], p[2]));
        g.positions.insert(g.positions.end(), {q.x, q.y, q.z});
        g.normals.insert(g.normals.end(), {normal.x, normal.y, normal.z});
        g.uvs.insert(g.uvs.end(), {uv[i][0], uv[i][1]});
        g.indices.push_back((uint32_t)i);
    }
    g.materialOfTri = {material, material};
    return g;
}

} // namespace

SceneDescription parseScene(const std::string &sceneJsonPath, const std::string &root, int width, int height)
{
    g_sceneRoot = root;
    std::ifstream file(resolve(root, sceneJsonPath));
    if (!file) { throw std::runtime_error("cannot open scene file: " + sceneJsonPath); }
    std::stringstream buffer;
    buffer << file.rdbuf();
    const Json sceneJson = Json::parse(buffer.str());

    SceneDescription scene;

    // camera, src/scene_parser.cpp:146-158
    const Json &sensor = sceneJson["sensor"];
    const float fov = parseFloat(sensor["fov"]);
    parseTriple(sensor["lookAt"]["origin"], scene.camera.origin);
    parseTriple(sensor["lookAt"]["target"], scene.camera.target);
    parseTriple(sensor["lookAt"]["up"], scene.camera.up);
    scene.camera.verticalFov = (float)(fov / 180.f * M_PI); // float / float * double -> float parameter
    scene.camera.width = width; scene.camera.height = height;
    scene.camera.flipHandedness = parseBool(sensor["flipHandedness"], false);

    // named materials, src/scene_parser.cpp:589-602
    MaterialMap lookup;
    for (const Json &materialJson : sceneJson["materials"].items()) {
        const std::string name = materialJson["name"].asString();
        const int id = parseMaterial(materialJson, lookup, scene);
        if (id >= 0) { lookup[name] = (uint32_t)id; }
    }

    // parseMedia, src/scene_parser.cpp:202-229 (homogeneous media; the heterogeneous grid needs .vol assets and stays out)
    std::map<std::string, int> mediaLookup;
    if (sceneJson["media"].isArray()) {
        for (const Json &mediumJson : sceneJson["media"].items()) {
            const std::string type = parseString(mediumJson["type"], "");
            if (type == "heterogeneous") { throw std::runtime_error("heterogeneous media are outside the accelerated path (SURVEY N3: HomogeneousMedium)"); }
            if (type != "homogeneous") { continue; }
            MediumDesc medium;
            medium.name = mediumJson["name"].asString();
            parseColor(mediumJson["sigma_t"], medium.sigmaT, 0.f);
            parseColor(mediumJson["sigma_s"], medium.sigmaS, 0.f);
            auto known = mediaLookup.find(medium.name);
            if (known != mediaLookup.end()) { scene.media[(size_t)known->second] = medium; } // media[name] = ...: the last one wins
            else { mediaLookup[medium.name] = (int)scene.media.size(); scene.media.push_back(medium); }
        }
    }
    // checkString(json["internal_medium"]) + media[key]: an unknown key default-constructs a null medium in the reference's map
    auto internalMedium = [&](const Json &object) {
        if (!object["internal_medium"].isString()) { return -1; }
        auto it = mediaLookup.find(object["internal_medium"].asString());
        return it == mediaLookup.end() ? -1 : it->second;
    };

    // models, parseObjects src/scene_parser.cpp:251-291; geometry ids follow attach order inside the scene being filled
    // (`out`: the root scene's list, or an instance scene's while parseInstance recurses)
    std::map<std::string, uint32_t> instanceLookup; // InstanceMap: name -> index into scene.instanceScenes
    std::function<void(const Json &, std::vector<GeometryDesc> &, bool)> parseObjects = [&](const Json &models, std::vector<GeometryDesc> &out, bool isRoot) {
    for (const Json &object : models.items()) {
        if (parseBool(object["skip"], false)) { continue; }
        const std::string type = parseString(object["type"], "");
        if (!isRoot && (type == "ply" || type == "sphere" || type == "quad")) {
            // the reference attaches these to the GLOBAL Embree scene even inside an instance definition (src/ply_parser.cpp:143,
            // src/sphere.cpp:46, src/quad.cpp:149) while registering their surfaces with the instance scene: its tables go out of step
            throw std::runtime_error("model type '" + type + "' inside an instance: only obj meshes can be instanced");
        }
        if (type == "obj") {
            const Transform transform = parseTransformOrIdentity(object["transform"]);
            const int material = parseMaterial(object["bsdf"], lookup, scene);
            out.push_back(parseObj(resolve(root, object["filename"].asString()), root, transform, lookup,
                                                parseString(object["materialPrefix"], ""), material, scene));
            out.back().internalMedium = internalMedium(object);
        } else if (type == "ply") {
            const Transform transform = parseTransformOrIdentity(object["transform"]);
            int material = parseMaterial(object["bsdf"], lookup, scene);
            if (material < 0) { // the PLY parser's own green Lambertian, src/ply_parser.cpp:115-117
                ptc_material_desc d; memset(&d, 0, sizeof(d));
                d.type = PTC_LAMBERTIAN; d.diffuse[1] = 1.f; d.ior = 1.4f;
                scene.materials.push_back(d);
                material = (int)scene.materials.size() - 1;
            }
            out.push_back(parsePly(resolve(root, object["filename"].asString()), transform, (uint32_t)material));
            out.back().internalMedium = internalMedium(object);
        } else if (type == "sphere") {
            const int material = parseMaterial(object["bsdf"], lookup, scene);
            if (material < 0) { throw std::runtime_error("sphere without bsdf"); }
            GeometryDesc g;
            g.isSphere = true;
            float center[3];
            parseTriple(object["center"], center);
            const Transform transform = parseTransformOrIdentity(object["transform"]);
            const Vec3 c = transform.applyPoint(Vec3(center[0], center[1], center[2])); // Sphere::create, src/sphere.cpp:30-35
            g.centerRadius[0] = c.x; g.centerRadius[1] = c.y; g.centerRadius[2] = c.z;
            g.centerRadius[3] = parseFloat(object["radius"]);
            g.sphereMaterial = (uint32_t)material;
            g.internalMedium = internalMedium(object);
            out.push_back(g);
        } else if (type == "quad") {
            const int material = parseMaterial(object["bsdf"], lookup, scene);
            if (material < 0) { throw std::runtime_error("quad without bsdf"); }
            const Transform transform = parseTransformOrIdentity(object["transform"]);
            bool zUp = false;
            if (object["upAxis"].isString()) {
                const std::string axis = object["upAxis"].asString();
                if (axis == "z") { zUp = true; } else if (axis != "y") { throw std::runtime_error("Unsupported axis: " + axis); }
            }
            out.push_back(makeQuad(transform, (uint32_t)material, zUp));
        } else if (type == "instance") { // parseInstance, src/scene_parser.cpp:231-249
            InstanceSceneDesc definition;
            definition.name = parseString(object["name"], "");
            parseObjects(object["models"], definition.geometries, false);
            instanceLookup[definition.name] = (uint32_t)scene.instanceScenes.size(); // instanceLookup[name] = ...: the last one wins
            scene.instanceScenes.push_back(std::move(definition));
        } else if (type == "instanced") { // parseInstanced, src/scene_parser.cpp:449-492
            auto known = instanceLookup.find(parseString(object["instance_name"], ""));
            if (known == instanceLookup.end()) { throw std::runtime_error("instanced: unknown instance_name " + parseString(object["instance_name"], "<missing>")); }
            GeometryDesc g;
            g.isInstance = true;
            g.instanceScene = known->second;
            for (size_t k = 0; k < 16; k++) { g.instanceTransform[k] = parseFloat(object["transform"][k]); } // RTC_FORMAT_FLOAT4X4_COLUMN_MAJOR
            out.push_back(g);
        } else if (type == "pbrt-curve" || type == "b-spline") {
            throw std::runtime_error("model type '" + type + "' is outside the surface path-tracing hot path (SURVEY §8(f))");
        }
        // unknown types are silently ignored, like the reference's if-chain
    }
    };
    parseObjects(sceneJson["models"], scene.geometries, true);

    // environment light, src/scene_parser.cpp:544-557
    const Json &environment = sceneJson["environmentLight"];
    if (environment.isObject()) {
        EnvironmentDesc &env = scene.environment;
        env.present = true;
        env.filename = environment["filename"].asString();
        env.scale = parseFloat(environment["scale"], 1.f);
        const Transform t = parseTransformOrIdentity(environment["transform"]);
        memcpy(env.mapToWorld, t.m, sizeof(env.mapToWorld));
        memcpy(env.worldToMap, t.inv, sizeof(env.worldToMap));
        loadEXR(resolve(root, env.filename), env.rgba, env.width, env.height);
    }
    return scene;
}

int feedScene(const SceneDescription &scene, const SceneSink &sink)
{
    int status = 0;
    for (const TextureDesc &texture : scene.textures) {
        if (!sink.add_texture) { return PTC_ERR_INVALID; }
        if ((status = sink.add_texture(sink.ctx, texture.rgb.data(), texture.width, texture.height, nullptr))) { return status; }
    }
    for (const ptc_material_desc &material : scene.materials) {
        if ((status = sink.add_material(sink.ctx, &material, nullptr))) { return status; }
    }
    for (const MediumDesc &medium : scene.media) {
        if (!sink.add_medium) { return PTC_ERR_INVALID; }
        if ((status = sink.add_medium(sink.ctx, medium.sigmaT, medium.sigmaS, nullptr))) { return status; }
    }
    std::vector<uint32_t> sinkScene(scene.instanceScenes.size(), 0); // instance scene -> the id the sink gave it
    auto feedGeometry = [&](const GeometryDesc &g, bool isRoot) -> int {
        uint32_t geomId = 0;
        int rc;
        if (g.isInstance) {
            if (!sink.add_instance || g.instanceScene >= sinkScene.size()) { return PTC_ERR_INVALID; }
            return sink.add_instance(sink.ctx, sinkScene[g.instanceScene], g.instanceTransform, &geomId);
        }
        if (g.isSphere) { rc = sink.add_sphere(sink.ctx, g.centerRadius, g.sphereMaterial, &geomId); }
        else {
            rc = sink.add_triangle_mesh(sink.ctx, g.positions.data(), g.normals.data(), g.uvs.data(),
                                        (uint32_t)(g.positions.size() / 3), g.indices.data(), g.materialOfTri.data(),
                                        (uint32_t)g.materialOfTri.size(), &geomId);
        }
        if (rc) { return rc; }
        if (g.internalMedium >= 0 && isRoot) {
            if (!sink.set_internal_medium) { return PTC_ERR_INVALID; }
            if ((rc = sink.set_internal_medium(sink.ctx, geomId, (uint32_t)g.internalMedium))) { return rc; }
        }
        return 0;
    };
    // instance definitions first, in order of completion (a scene only places scenes defined before it), each in its own
    // begin / end pair; then the root scene's geometries
    for (size_t i = 0; i < scene.instanceScenes.size(); i++) {
        if (!sink.begin_instance || !sink.end_instance) { return PTC_ERR_INVALID; }
        if ((status = sink.begin_instance(sink.ctx, &sinkScene[i]))) { return status; }
        for (const GeometryDesc &g : scene.instanceScenes[i].geometries) { if ((status = feedGeometry(g, false))) { return status; } }
        if ((status = sink.end_instance(sink.ctx))) { return status; }
    }
    for (const GeometryDesc &g : scene.geometries) { if ((status = feedGeometry(g, true))) { return status; } }
    if (scene.environment.present) {
        const EnvironmentDesc &e = scene.environment;
        if ((status = sink.set_environment(sink.ctx, e.rgba.data(), e.width, e.height, e.scale, e.mapToWorld, e.worldToMap))) { return status; }
    }
    const CameraDesc &c = scene.camera;
    if ((status = sink.set_camera(sink.ctx, c.origin, c.target, c.up, c.verticalFov, c.width, c.height, c.flipHandedness ? 1 : 0))) { return status; }
    return sink.commit(sink.ctx);
}

} // namespace pathed
