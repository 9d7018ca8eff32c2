// OpenEXR scanline reader/writer covering what the reference does through tinyexr:
//   LoadEXR  -> RGBA fp32           /root/reference/src/environment_light.cpp:23
//   SaveEXRImageToFile, 3 x HALF, channel order B,G,R, no compression   src/image.cpp:80-154
// Reads uncompressed / RLE / ZIPS / ZIP / PIZ scanline files with HALF or FLOAT channels (every block header is range-checked);
// writes uncompressed files.
#pragma once

#include <string>
#include <vector>

namespace pathed {

// rgba: 4*width*height floats, scanline 0 = top (file order), alpha = 1 when absent
void loadEXR(const std::string &path, std::vector<float> &rgba, int &width, int &height);

// planar fp32 channels in the order given (names sorted by the caller as EXR requires)
void saveEXR(const std::string &path, int width, int height, const std::vector<std::string> &channelNames,
             const std::vector<const float *> &channels, bool asHalf);

// the same file as bytes (a checkpoint is written under two names), and a plain whole-file write
std::vector<unsigned char> encodeEXR(int width, int height, const std::vector<std::string> &channelNames, const std::vector<const float *> &channels, bool asHalf);
void writeFileBytes(const std::string &path, const std::vector<unsigned char> &bytes);

unsigned short floatToHalf(float value);
float halfToFloat(unsigned short bits);

} // namespace pathed
