// Stub for the reference's PtexLocal (src/ptex_local.cpp needs the Ptex library,
// which is out of scope). Same three entry points, no texture data.
#include "ptex_local.h"
PtexLocal::PtexLocal(const std::string &texturePath) : Texture(texturePath) {}
void PtexLocal::load() {}
Color PtexLocal::lookup(const Intersection &intersection) const { return Color(0.f); }
