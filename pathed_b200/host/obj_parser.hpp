// Wavefront OBJ / MTL and binary PLY readers with the reference's exact semantics
// (/root/reference/src/obj_parser.cpp, src/mtl_parser.cpp, src/ply_parser.cpp).
#pragma once

#include "scene_description.hpp"
#include "transform.hpp"

#include <map>
#include <string>

namespace pathed {

using MaterialMap = std::map<std::string, uint32_t>; // name -> id in SceneDescription::materials

// One OBJ file = one geometry (src/obj_parser.cpp:50-128).  Material resolution per face (Q15):
// prefix+group -> prefix+usemtl -> MTL usemtl -> defaultMaterial (or a red Lambertian when < 0).
GeometryDesc parseObj(const std::string &path, const std::string &rootDirectory, const Transform &transform,
                      const MaterialMap &sceneMaterials, const std::string &materialPrefix, int defaultMaterial,
                      SceneDescription &scene);

// binary_little_endian PLY with float x,y,z and uchar/int triangle lists (src/ply_parser.cpp:29-151)
GeometryDesc parsePly(const std::string &path, const Transform &transform, uint32_t material);

} // namespace pathed
