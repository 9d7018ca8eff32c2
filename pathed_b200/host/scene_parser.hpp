// Scene JSON -> SceneDescription.  Mirrors parseScene (/root/reference/src/scene_parser.cpp:140-200):
// same keys, same string-encoded scalars (stof), same defaults, same errors (std::runtime_error).
#pragma once

#include "scene_description.hpp"

#include <string>

namespace pathed {

// `rootDirectory` plays the role of the reference's working directory after its chdir("..")
// (app/main.cpp:60): every filename inside scene files is relative to it.
SceneDescription parseScene(const std::string &sceneJsonPath, const std::string &rootDirectory, int width, int height);

} // namespace pathed
