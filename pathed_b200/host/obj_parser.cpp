#include "obj_parser.hpp"

#include <cstring>
#include <fstream>
#include <stdexcept>
#include <utility>

namespace pathed {

namespace {

std::string lTrim(const std::string &token) // src/string_util.cpp:27-37
{
    const std::string::size_type first = token.find_first_not_of(" \t");
    if (first == std::string::npos) { return ""; }
    return token.substr(first);
}

std::vector<std::string> tokenize(const std::string &line) // src/string_util.cpp:7-25
{
    std::vector<std::string> tokens;
    std::string remaining = lTrim(line);
    while (!remaining.empty()) {
        const std::string::size_type end = remaining.find_first_of(" \t");
        if (end == std::string::npos) { tokens.push_back(remaining); break; }
        tokens.push_back(remaining.substr(0, end));
        remaining = lTrim(remaining.substr(end));
    }
    return tokens;
}

ptc_material_desc lambertianDesc(float r, float g, float b, float er, float eg, float eb)
{
    ptc_material_desc d;
    memset(&d, 0, sizeof(d));
    d.type = PTC_LAMBERTIAN;
    d.diffuse[0] = r; d.diffuse[1] = g; d.diffuse[2] = b;
    d.emit[0] = er; d.emit[1] = eg; d.emit[2] = eb;
    d.ior = 1.4f;
    return d;
}

// src/mtl_parser.cpp: only newmtl / Kd / Ke are read; every entry becomes a Lambertian(diffuse, emit)
MaterialMap parseMtl(const std::string &path, SceneDescription &scene)
{
    struct Entry { float kd[3] = {0.f, 0.f, 0.f}; float ke[3] = {0.f, 0.f, 0.f}; };
    std::map<std::string, Entry> entries;
    std::string current;
    std::ifstream file(path);
    std::string line;
    while (std::getline(file, line)) {
        const std::vector<std::string> tokens = tokenize(line);
        if (tokens.empty()) { continue; }
        if (tokens[0] == "newmtl" && tokens.size() >= 2) { current = tokens[1]; entries[current] = Entry(); }
        else if ((tokens[0] == "Kd" || tokens[0] == "Ke") && tokens.size() >= 4) {
            float *dst = tokens[0] == "Kd" ? entries[current].kd : entries[current].ke;
            for (int i = 0; i < 3; i++) { dst[i] = std::stof(tokens[1 + i]); }
        }
    }
    MaterialMap lookup;
    for (const auto &item : entries) { // std::map order, like bakeLookup (src/mtl_parser.cpp:26-38)
        scene.materials.push_back(lambertianDesc(item.second.kd[0], item.second.kd[1], item.second.kd[2],
                                                 item.second.ke[0], item.second.ke[1], item.second.ke[2]));
        lookup[item.first] = (uint32_t)scene.materials.size() - 1;
    }
    return lookup;
}

struct FaceVertex { int vertex, normal, uv; };
struct Face { FaceVertex v[3]; };

// classifies one whitespace-separated face token: "7", "7/8/9" or "7//9"
bool parseFaceToken(const std::string &token, int &v, int &t, int &n, int &kind)
{
    const char *s = token.c_str();
    char *end = nullptr;
    v = (int)strtol(s, &end, 10);
    if (end == s) { return false; }
    if (*end == 0) { kind = 0; return true; }
    if (*end != '/') { return false; }
    s = end + 1;
    if (*s == '/') {
        s++;
        n = (int)strtol(s, &end, 10);
        if (end == s || *end != 0) { return false; }
        kind = 2;
        return true;
    }
    t = (int)strtol(s, &end, 10);
    if (end == s || *end != '/') { return false; }
    s = end + 1;
    n = (int)strtol(s, &end, 10);
    if (end == s || *end != 0) { return false; }
    kind = 1;
    return true;
}

} // namespace

GeometryDesc parseObj(const std::string &path, const std::string &rootDirectory, const Transform &transform,
                      const MaterialMap &sceneMaterials, const std::string &materialPrefix, int defaultMaterial,
                      SceneDescription &scene)
{
    std::ifstream file(path);
    if (!file) { throw std::runtime_error("cannot open OBJ file: " + path); }

    if (defaultMaterial < 0) { // src/obj_parser.cpp:40-45
        scene.materials.push_back(lambertianDesc(1.f, 0.f, 0.f, 0.f, 0.f, 0.f));
        defaultMaterial = (int)scene.materials.size() - 1;
    }

    std::vector<Vec3> vertices, normals;
    std::vector<std::pair<float, float>> uvs;
    std::vector<Vec3> vertexNormals;
    std::vector<std::pair<float, float>> vertexUVs;
    std::vector<Face> faces;
    std::vector<uint32_t> faceMaterials;
    MaterialMap mtlLookup;
    std::string currentGroup, currentMaterial;

    auto resolveIndex = [](int index, size_t count) { return index < 0 ? index + (int)count : index - 1; }; // :226-232
    auto resolveMaterial = [&]() -> uint32_t { // src/obj_parser.cpp:241-252
        const std::string groupKey = materialPrefix + currentGroup, mtlKey = materialPrefix + currentMaterial;
        auto it = sceneMaterials.find(groupKey);
        if (it != sceneMaterials.end()) { return it->second; }
        it = sceneMaterials.find(mtlKey);
        if (it != sceneMaterials.end()) { return it->second; }
        it = mtlLookup.find(currentMaterial);
        if (it != mtlLookup.end()) { return it->second; }
        return (uint32_t)defaultMaterial;
    };
    auto checkVertex = [&](int index) {
        if (index < 0 || index >= (int)vertices.size()) { throw std::runtime_error("OBJ vertex index out of range in " + path); }
    };
    // the three processTriangle overloads, src/obj_parser.cpp:278-372
    auto addTriangle = [&](const FaceVertex (&in)[3], int kind) {
        Face face;
        for (int i = 0; i < 3; i++) {
            face.v[i].vertex = resolveIndex(in[i].vertex, vertices.size());
            checkVertex(face.v[i].vertex);
            face.v[i].normal = kind >= 1 ? resolveIndex(in[i].normal, normals.size()) : -1;
            face.v[i].uv = kind == 1 ? resolveIndex(in[i].uv, uvs.size()) : -1;
        }
        faceMaterials.push_back(resolveMaterial());
        if (kind == 1) {
            vertexUVs.resize(vertices.size(), {0.f, 0.f});
            for (int i = 0; i < 3; i++) { vertexUVs[face.v[i].vertex] = uvs.at(face.v[i].uv); }
        }
        if (kind >= 1) {
            vertexNormals.resize(vertices.size(), Vec3());
            for (int i = 0; i < 3; i++) { vertexNormals[face.v[i].vertex] = normals.at(face.v[i].normal); }
        }
        faces.push_back(face);
    };

    std::string line;
    while (std::getline(file, line)) { // parseLine, src/obj_parser.cpp:130-167
        if (line.empty()) { continue; }
        const std::string::size_type space = line.find_first_of(" \t");
        if (space == std::string::npos) { continue; }
        const std::string command = line.substr(0, space);
        if (command[0] == '#') { continue; }
        const std::string rest = lTrim(line.substr(space + 1));

        if (command == "v" || command == "vn") {
            const char *s = rest.c_str(); char *end = nullptr;
            const float x = strtof(s, &end); s = end;
            const float y = strtof(s, &end); s = end;
            const float z = strtof(s, &end);
            if (command == "v") { vertices.push_back(transform.applyPoint(Vec3(x, y, z))); }
            else { normals.push_back(transform.applyVector(Vec3(x, y, z))); } // forward matrix, not inverse-transpose (Q4)
        } else if (command == "vt") {
            const char *s = rest.c_str(); char *end = nullptr;
            const float u = strtof(s, &end); s = end;
            const float v = strtof(s, &end);
            uvs.push_back({u, v});
        } else if (command == "g") {
            currentGroup = lTrim(rest);
        } else if (command == "usemtl") {
            currentMaterial = rest;
        } else if (command == "mtllib") {
            mtlLookup = parseMtl(rootDirectory.empty() || rest[0] == '/' ? rest : rootDirectory + "/" + rest, scene);
        } else if (command == "f") {
            if (currentMaterial == "hidden") { continue; } // src/obj_parser.cpp:151-153
            const std::vector<std::string> tokens = tokenize(rest);
            FaceVertex fv[4];
            int kinds[4] = {-1, -1, -1, -1};
            const size_t count = tokens.size() < 4 ? tokens.size() : 4;
            bool ok = tokens.size() >= 3;
            for (size_t i = 0; ok && i < count; i++) {
                fv[i] = {0, 0, 0};
                ok = parseFaceToken(tokens[i], fv[i].vertex, fv[i].uv, fv[i].normal, kinds[i]);
            }
            // the four regular expressions of src/obj_parser.cpp:374-476 plus the stoi fallback (:478-506)
            const bool same3 = ok && kinds[0] == kinds[1] && kinds[1] == kinds[2];
            const bool quad = same3 && tokens.size() == 4 && kinds[3] == kinds[0];
            if (quad && kinds[0] == 0) {
                const FaceVertex a[3] = {fv[0], fv[1], fv[2]}, b[3] = {fv[0], fv[2], fv[3]};
                addTriangle(a, 0); addTriangle(b, 0);
            } else if (same3 && tokens.size() == 3 && kinds[0] == 1) {
                const FaceVertex a[3] = {fv[0], fv[1], fv[2]};
                addTriangle(a, 1);
            } else if (same3 && tokens.size() == 3 && kinds[0] == 2) {
                const FaceVertex a[3] = {fv[0], fv[1], fv[2]};
                addTriangle(a, 2);
            } else if (quad && kinds[0] == 2) {
                const FaceVertex a[3] = {fv[0], fv[1], fv[2]}, b[3] = {fv[0], fv[2], fv[3]};
                addTriangle(a, 2); addTriangle(b, 2);
            } else if (same3 && kinds[0] == 0) {
                const FaceVertex a[3] = {fv[0], fv[1], fv[2]};
                addTriangle(a, 0);
            } else {
                throw std::runtime_error("unsupported OBJ face syntax: " + line); // std::stoi throws in the reference
            }
        }
    }

    // "cube-normal" correction, src/obj_parser.cpp:57-117: a vertex reused with a different normal index is
    // duplicated (appended), then per-vertex normals are rewritten face by face
    std::map<int, int> normalLookup;
    std::map<std::pair<int, int>, int> correctionLookup;
    for (Face &face : faces) {
        for (int j = 0; j < 3; j++) {
            FaceVertex &fv = face.v[j];
            auto seen = normalLookup.find(fv.vertex);
            if (seen == normalLookup.end()) {
                normalLookup[fv.vertex] = fv.normal;
            } else if (seen->second != fv.normal) {
                const std::pair<int, int> key(fv.vertex, fv.normal);
                auto fixed = correctionLookup.find(key);
                int corrected;
                if (fixed == correctionLookup.end()) {
                    vertices.push_back(vertices[fv.vertex]);
                    corrected = (int)vertices.size() - 1;
                    correctionLookup[key] = corrected;
                } else {
                    corrected = fixed->second;
                }
                fv.vertex = corrected;
            }
        }
    }
    vertexNormals.resize(vertices.size(), Vec3());
    for (const Face &face : faces) {
        for (int j = 0; j < 3; j++) {
            if (face.v[j].normal != -1) { vertexNormals[face.v[j].vertex] = normals.at(face.v[j].normal); }
        }
    }
    vertexUVs.resize(vertices.size(), {0.f, 0.f}); // src/geometry_parser.cpp:67-71

    GeometryDesc geometry;
    geometry.positions.reserve(vertices.size() * 3);
    for (size_t i = 0; i < vertices.size(); i++) {
        geometry.positions.insert(geometry.positions.end(), {vertices[i].x, vertices[i].y, vertices[i].z});
        geometry.normals.insert(geometry.normals.end(), {vertexNormals[i].x, vertexNormals[i].y, vertexNormals[i].z});
        geometry.uvs.insert(geometry.uvs.end(), {vertexUVs[i].first, vertexUVs[i].second});
    }
    for (const Face &face : faces) {
        for (int j = 0; j < 3; j++) { geometry.indices.push_back((uint32_t)face.v[j].vertex); }
    }
    geometry.materialOfTri = faceMaterials;
    return geometry;
}

GeometryDesc parsePly(const std::string &path, const Transform &transform, uint32_t material)
{
    std::ifstream file(path, std::ios::binary);
    if (!file) { throw std::runtime_error("cannot open PLY file: " + path); }
    auto expect = [&](const std::string &line, const std::string &a, const std::string &b = "") {
        if (line != a && (b.empty() || line != b)) { throw std::runtime_error("unsupported PLY header line '" + line + "' in " + path); }
    };
    auto count = [&](const std::string &line, const std::string &prefix) {
        if (line.compare(0, prefix.size(), prefix) != 0) { throw std::runtime_error("unsupported PLY header line '" + line + "' in " + path); }
        return std::stoi(line.substr(prefix.size()));
    };
    std::string line;
    std::getline(file, line); expect(line, "ply");
    std::getline(file, line); expect(line, "format binary_little_endian 1.0");
    std::getline(file, line); const int vertexCount = count(line, "element vertex ");
    std::getline(file, line); expect(line, "property float x");
    std::getline(file, line); expect(line, "property float y");
    std::getline(file, line); expect(line, "property float z");
    std::getline(file, line); const int faceCount = count(line, "element face ");
    std::getline(file, line); expect(line, "property list uint8 int vertex_indices", "property list uchar int vertex_indices");
    std::getline(file, line); expect(line, "end_header");

    GeometryDesc geometry;
    for (int i = 0; i < vertexCount; i++) {
        float p[3];
        file.read((char *)p, 12);
        const Vec3 v = transform.applyPoint(Vec3(p[0], p[1], p[2]));
        geometry.positions.insert(geometry.positions.end(), {v.x, v.y, v.z});
        geometry.normals.insert(geometry.normals.end(), {0.f, 0.f, 0.f});
        geometry.uvs.insert(geometry.uvs.end(), {0.f, 0.f});
    }
    for (int i = 0; i < faceCount; i++) {
        unsigned char faceSize = 0;
        file.read((char *)&faceSize, 1);
        if (faceSize != 3) { throw std::runtime_error("PLY faces must be triangles: " + path); }
        int index[3];
        file.read((char *)index, 12);
        for (int j = 0; j < 3; j++) {
            if (index[j] < 0 || index[j] >= vertexCount) { throw std::runtime_error("PLY vertex index out of range: " + path); }
            geometry.indices.push_back((uint32_t)index[j]);
        }
        geometry.materialOfTri.push_back(material);
    }
    if (!file) { throw std::runtime_error("truncated PLY file: " + path); }
    return geometry;
}

} // namespace pathed
