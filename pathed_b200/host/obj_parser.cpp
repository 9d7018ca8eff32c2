#include "obj_parser.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
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
bool parseFaceToken(const char *s, int &v, int &t, int &n, int &kind) // s: zero-terminated (no std::string: one per token was a heap allocation per token)
{
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

// Lines of an OBJ file as records.  Scanning the text (number parsing is what costs: 1.3 M lines for the dragon) is done by all host
// threads on line-aligned chunks of the file; the records are then replayed in file order by ONE thread, because the format is
// stateful (current group / material, indices relative to the vertices seen so far, "last face wins" per-vertex normals and uvs).
struct ObjRecord {
    enum Kind : uint8_t { Vertex, Normal, Uv, Group, UseMtl, MtlLib, Face, BadFace } kind;
    uint8_t tokens = 0;            // Face: number of whitespace-separated tokens (3, 4, or more)
    int8_t kinds[4] = {-1, -1, -1, -1}; // Face: per token 0 "7", 1 "7/8/9", 2 "7//9"
    float f[3] = {0.f, 0.f, 0.f};  // Vertex / Normal / Uv
    FaceVertex fv[4] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    uint32_t textBegin = 0, textEnd = 0; // Group / UseMtl / MtlLib: the argument; BadFace: the line (for the message)
};

bool isBlank(char c) { return c == ' ' || c == '\t'; }

// one token of a face line, [s, e): "7", "7/8/9" or "7//9" (parseFaceToken on a range)
bool scanFaceToken(const char *s, const char *e, FaceVertex &out, int8_t &kind)
{
    char buffer[64];
    const size_t n = (size_t)(e - s);
    if (n == 0 || n >= sizeof(buffer)) { return false; }
    memcpy(buffer, s, n); buffer[n] = 0;
    int v = 0, t = 0, nrm = 0, k = -1;
    if (!parseFaceToken(buffer, v, t, nrm, k)) { return false; }
    out.vertex = v; out.uv = t; out.normal = nrm; kind = (int8_t)k;
    return true;
}

void scanObjChunk(const char *text, size_t begin, size_t end, std::vector<ObjRecord> &out)
{
    size_t pos = begin;
    while (pos < end) {
        size_t eol = pos;
        while (eol < end && text[eol] != '\n') { eol++; }
        size_t lineEnd = eol;
        if (lineEnd > pos && text[lineEnd - 1] == '\r') { /* std::getline keeps the '\r': it stays part of the last token */ }
        const char *line = text + pos, *stop = text + lineEnd;
        pos = eol + 1;
        if (line == stop) { continue; }
        const char *space = line;
        while (space < stop && !isBlank(*space)) { space++; }
        if (space == stop || line[0] == '#') { continue; } // no argument / comment (parseLine, src/obj_parser.cpp:130-167)
        const size_t commandLength = (size_t)(space - line);
        const char *rest = space + 1;
        while (rest < stop && isBlank(*rest)) { rest++; }
        auto is = [&](const char *name) { return strlen(name) == commandLength && !memcmp(line, name, commandLength); };
        ObjRecord r;
        if (is("v") || is("vn") || is("vt")) {
            r.kind = is("v") ? ObjRecord::Vertex : (is("vn") ? ObjRecord::Normal : ObjRecord::Uv);
            // strtof on a copy of the line: it must not read past the end of the line into the next one
            char buffer[256];
            const size_t n = std::min((size_t)(stop - rest), sizeof(buffer) - 1);
            memcpy(buffer, rest, n); buffer[n] = 0;
            const char *p = buffer; char *next = nullptr;
            for (int c = 0; c < (r.kind == ObjRecord::Uv ? 2 : 3); c++) { r.f[c] = strtof(p, &next); p = next; }
        } else if (is("g") || is("usemtl") || is("mtllib")) {
            r.kind = is("g") ? ObjRecord::Group : (is("usemtl") ? ObjRecord::UseMtl : ObjRecord::MtlLib);
            r.textBegin = (uint32_t)(rest - text); r.textEnd = (uint32_t)(stop - text);
        } else if (is("f")) {
            r.kind = ObjRecord::Face;
            bool ok = true;
            const char *p = rest;
            unsigned count = 0;
            while (p < stop) {
                const char *tokenEnd = p;
                while (tokenEnd < stop && !isBlank(*tokenEnd)) { tokenEnd++; }
                if (count < 4 && ok) { ok = scanFaceToken(p, tokenEnd, r.fv[count], r.kinds[count]); }
                count++;
                p = tokenEnd;
                while (p < stop && isBlank(*p)) { p++; }
            }
            r.tokens = (uint8_t)std::min(count, 255u);
            if (!ok || count < 3) { r.kind = ObjRecord::BadFace; r.textBegin = (uint32_t)(line - text); r.textEnd = (uint32_t)(stop - text); }
        } else { continue; }
        out.push_back(r);
    }
}

} // namespace

GeometryDesc parseObj(const std::string &path, const std::string &rootDirectory, const Transform &transform,
                      const MaterialMap &sceneMaterials, const std::string &materialPrefix, int defaultMaterial,
                      SceneDescription &scene)
{
    const auto T0 = std::chrono::steady_clock::now(); auto lap = [&](const char *w){ if (getenv("PTH_OBJ_TIMING")) fprintf(stderr, "obj %s %.3f\n", w, std::chrono::duration<double>(std::chrono::steady_clock::now()-T0).count()); };
    std::string text;
    {
        std::ifstream file(path, std::ios::binary);
        if (!file) { throw std::runtime_error("cannot open OBJ file: " + path); }
        file.seekg(0, std::ios::end);
        text.resize((size_t)file.tellg());
        file.seekg(0);
        file.read(&text[0], (std::streamsize)text.size());
    }
    if (text.size() >= 0xFFFFFFFFull) { throw std::runtime_error("OBJ file larger than 4 GB: " + path); }

    if (defaultMaterial < 0) { // src/obj_parser.cpp:40-45
        scene.materials.push_back(lambertianDesc(1.f, 0.f, 0.f, 0.f, 0.f, 0.f));
        defaultMaterial = (int)scene.materials.size() - 1;
    }

    // scan: line-aligned chunks, one per host thread
    const size_t threads = std::max<size_t>(1, std::min<size_t>(std::thread::hardware_concurrency(), text.size() / (1 << 20) + 1));
    std::vector<size_t> cut(threads + 1, text.size());
    cut[0] = 0;
    for (size_t t = 1; t < threads; t++) {
        size_t at = std::max(cut[t - 1], text.size() * t / threads);
        while (at < text.size() && text[at] != '\n') { at++; }
        cut[t] = std::min(text.size(), at + 1);
    }
    std::vector<std::vector<ObjRecord>> chunks(threads);
    {
        std::vector<std::thread> pool;
        for (size_t t = 0; t < threads; t++) { pool.emplace_back([&, t]() { scanObjChunk(text.data(), cut[t], cut[t + 1], chunks[t]); }); }
        for (std::thread &t : pool) { t.join(); }
    }

    lap("scan");
    std::vector<Vec3> vertices, normals;
    std::vector<std::pair<float, float>> uvs;
    std::vector<Vec3> vertexNormals;
    std::vector<std::pair<float, float>> vertexUVs;
    std::vector<Face> faces;
    std::vector<uint32_t> faceMaterials;
    MaterialMap mtlLookup;
    std::string currentGroup, currentMaterial;
    bool materialKnown = false; // the material of the faces that follow, resolved once per group / usemtl / mtllib change
    uint32_t material = 0;

    auto resolveIndex = [](int index, size_t count) { return index < 0 ? index + (int)count : index - 1; }; // :226-232
    auto resolveMaterial = [&]() -> uint32_t { // src/obj_parser.cpp:241-252
        if (materialKnown) { return material; }
        const std::string groupKey = materialPrefix + currentGroup, mtlKey = materialPrefix + currentMaterial;
        materialKnown = true;
        auto it = sceneMaterials.find(groupKey);
        if (it != sceneMaterials.end()) { return material = it->second; }
        it = sceneMaterials.find(mtlKey);
        if (it != sceneMaterials.end()) { return material = it->second; }
        it = mtlLookup.find(currentMaterial);
        if (it != mtlLookup.end()) { return material = it->second; }
        return material = (uint32_t)defaultMaterial;
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

    size_t total = 0;
    for (const auto &chunk : chunks) { total += chunk.size(); }
    faces.reserve(total); faceMaterials.reserve(total); vertices.reserve(total / 2);
    for (const auto &chunk : chunks) { // replay, in file order
        for (const ObjRecord &r : chunk) {
            switch (r.kind) {
            case ObjRecord::Vertex: vertices.push_back(transform.applyPoint(Vec3(r.f[0], r.f[1], r.f[2]))); break;
            case ObjRecord::Normal: normals.push_back(transform.applyVector(Vec3(r.f[0], r.f[1], r.f[2]))); break; // forward matrix, not inverse-transpose (Q4)
            case ObjRecord::Uv: uvs.push_back({r.f[0], r.f[1]}); break;
            case ObjRecord::Group: currentGroup = lTrim(text.substr(r.textBegin, r.textEnd - r.textBegin)); materialKnown = false; break;
            case ObjRecord::UseMtl: currentMaterial = text.substr(r.textBegin, r.textEnd - r.textBegin); materialKnown = false; break;
            case ObjRecord::MtlLib: {
                const std::string rest = text.substr(r.textBegin, r.textEnd - r.textBegin);
                mtlLookup = parseMtl(rootDirectory.empty() || rest[0] == '/' ? rest : rootDirectory + "/" + rest, scene);
                materialKnown = false;
                break;
            }
            case ObjRecord::BadFace:
                if (currentMaterial == "hidden") { break; }
                throw std::runtime_error("unsupported OBJ face syntax: " + text.substr(r.textBegin, r.textEnd - r.textBegin)); // std::stoi throws in the reference
            case ObjRecord::Face: {
                if (currentMaterial == "hidden") { break; } // src/obj_parser.cpp:151-153
                const FaceVertex *fv = r.fv;
                const int8_t *kinds = r.kinds;
                // the four regular expressions of src/obj_parser.cpp:374-476 plus the stoi fallback (:478-506)
                const bool same3 = kinds[0] == kinds[1] && kinds[1] == kinds[2];
                const bool quad = same3 && r.tokens == 4 && kinds[3] == kinds[0];
                if (quad && kinds[0] == 0) {
                    const FaceVertex a[3] = {fv[0], fv[1], fv[2]}, b[3] = {fv[0], fv[2], fv[3]};
                    addTriangle(a, 0); addTriangle(b, 0);
                } else if (same3 && r.tokens == 3 && kinds[0] == 1) {
                    const FaceVertex a[3] = {fv[0], fv[1], fv[2]};
                    addTriangle(a, 1);
                } else if (same3 && r.tokens == 3 && kinds[0] == 2) {
                    const FaceVertex a[3] = {fv[0], fv[1], fv[2]};
                    addTriangle(a, 2);
                } else if (quad && kinds[0] == 2) {
                    const FaceVertex a[3] = {fv[0], fv[1], fv[2]}, b[3] = {fv[0], fv[2], fv[3]};
                    addTriangle(a, 2); addTriangle(b, 2);
                } else if (same3 && kinds[0] == 0) {
                    const FaceVertex a[3] = {fv[0], fv[1], fv[2]};
                    addTriangle(a, 0);
                } else {
                    throw std::runtime_error("unsupported OBJ face syntax in " + path); // std::stoi throws in the reference
                }
                break;
            }
            }
        }
    }

    lap("replay");
    // "cube-normal" correction, src/obj_parser.cpp:57-117: a vertex reused with a different normal index is
    // duplicated (appended), then per-vertex normals are rewritten face by face.  The reference keeps the first normal index seen per
    // vertex in a std::map; a flat table does the same (no iteration order is involved), the rare conflicts stay in a map.
    const int unseen = -2; // normal indices are >= -1
    std::vector<int> normalLookup(vertices.size(), unseen);
    std::map<std::pair<int, int>, int> correctionLookup;
    for (Face &face : faces) {
        for (int j = 0; j < 3; j++) {
            FaceVertex &fv = face.v[j];
            int &seen = normalLookup[(size_t)fv.vertex];
            if (seen == unseen) {
                seen = fv.normal;
            } else if (seen != fv.normal) {
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

    lap("correction");
    GeometryDesc geometry;
    geometry.positions.resize(vertices.size() * 3); geometry.normals.resize(vertices.size() * 3); geometry.uvs.resize(vertices.size() * 2);
    for (size_t i = 0; i < vertices.size(); i++) {
        geometry.positions[3 * i] = vertices[i].x; geometry.positions[3 * i + 1] = vertices[i].y; geometry.positions[3 * i + 2] = vertices[i].z;
        geometry.normals[3 * i] = vertexNormals[i].x; geometry.normals[3 * i + 1] = vertexNormals[i].y; geometry.normals[3 * i + 2] = vertexNormals[i].z;
        geometry.uvs[2 * i] = vertexUVs[i].first; geometry.uvs[2 * i + 1] = vertexUVs[i].second;
    }
    geometry.indices.resize(faces.size() * 3);
    for (size_t t = 0; t < faces.size(); t++) {
        for (int j = 0; j < 3; j++) { geometry.indices[3 * t + j] = (uint32_t)faces[t].v[j].vertex; }
    }
    geometry.materialOfTri = faceMaterials;
    lap("assemble");
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
