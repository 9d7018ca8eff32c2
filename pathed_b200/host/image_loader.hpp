// Image files -> the 8-bit RGB texels Texture::load keeps (reference: src/texture.cpp:12-32, stbi_load(..., 3)).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace pathed {

// width * height * 3 bytes, row 0 = top of the image.  PNG and binary PNM are decoded here (lossless formats: the bytes
// equal stb_image's).  A JPEG is rejected with std::runtime_error: its texels depend on the decoder's IDCT and chroma
// upsampling, so bit parity with the reference needs the reference's own decoder -- a maintainer binding the C ABI keeps
// stbi_load in Texture::load and passes m_data to ptc_add_texture.
void loadImageRGB8(const std::string &path, std::vector<uint8_t> &rgb, int &width, int &height);

} // namespace pathed
