// 8-bit RGB image loading for Texture::load (reference: src/texture.cpp:12-32, which calls stbi_load(path, &w, &h, &n, 3)
// of the vendored stb_image v2.22).  The result is what that call returns: width * height * 3 bytes, row 0 = top.
//
// Decoders written here (no third-party code): PNG (non-interlaced and Adam7, bit depths 1-16, all colour types; zlib
// inflate from the system library) and binary PNM (P5 / P6, maxval <= 255).  Both formats are lossless, so the bytes are
// defined by the file format and equal stb_image's.  Conversion to three channels follows stb's rules: grey is replicated,
// alpha is dropped (never blended), 16-bit samples keep their high byte, low-bit-depth grey is scaled to 0..255.
// JPEG decoding is not exact across decoders (IDCT / upsampling are implementation-defined); see image_loader.hpp.
#include "image_loader.hpp"

#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace pathed {

namespace {

std::vector<uint8_t> readFile(const std::string &path)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { throw std::runtime_error("Error loading texture: cannot open " + path); }
    std::vector<uint8_t> data;
    uint8_t buffer[1 << 16];
    size_t n;
    while ((n = fread(buffer, 1, sizeof(buffer), f)) > 0) { data.insert(data.end(), buffer, buffer + n); }
    fclose(f);
    return data;
}

uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

int paeth(int a, int b, int c)
{
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    if (pa <= pb && pa <= pc) { return a; }
    return pb <= pc ? b : c;
}

// Undo the per-scanline filters of one (sub-)image: `raw` holds height * (1 + stride) bytes; returns height * stride bytes.
void unfilter(const uint8_t *raw, size_t stride, size_t bpp, uint32_t height, std::vector<uint8_t> &out)
{
    out.assign((size_t)height * stride, 0);
    for (uint32_t y = 0; y < height; y++) {
        const uint8_t *src = raw + (size_t)y * (stride + 1);
        const int filter = src[0];
        src++;
        uint8_t *cur = out.data() + (size_t)y * stride;
        const uint8_t *prev = y ? cur - stride : nullptr;
        for (size_t i = 0; i < stride; i++) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev ? prev[i] : 0, c = (prev && i >= bpp) ? prev[i - bpp] : 0;
            int v = src[i];
            switch (filter) {
            case 0: break;
            case 1: v += a; break;
            case 2: v += b; break;
            case 3: v += (a + b) >> 1; break;
            case 4: v += paeth(a, b, c); break;
            default: throw std::runtime_error("Error loading texture: bad PNG filter");
            }
            cur[i] = (uint8_t)v;
        }
    }
}

struct PngInfo {
    uint32_t width = 0, height = 0;
    int depth = 0, colorType = 0, interlace = 0;
    uint8_t palette[256][3];
    int paletteSize = 0;
};

int channelsOf(int colorType) { return colorType == 0 ? 1 : colorType == 2 ? 3 : colorType == 3 ? 1 : colorType == 4 ? 2 : 4; }

// sample s of a scanline as an 8-bit value (stb: 16-bit keeps the high byte, grey below 8 bits is scaled, palette indices are raw)
inline int sampleAt(const uint8_t *line, size_t s, int depth, bool scale)
{
    if (depth == 8) { return line[s]; }
    if (depth == 16) { return line[2 * s]; }
    const int perByte = 8 / depth, shift = (perByte - 1 - (int)(s % perByte)) * depth;
    const int v = (line[s / perByte] >> shift) & ((1 << depth) - 1);
    static const int factor[5] = {0, 0xFF, 0x55, 0, 0x11};
    return scale ? v * factor[depth] : v;
}

void storePixels(const PngInfo &info, const std::vector<uint8_t> &lines, size_t stride, uint32_t w, uint32_t h, uint32_t x0, uint32_t y0,
                 uint32_t dx, uint32_t dy, std::vector<uint8_t> &rgb)
{
    const int channels = channelsOf(info.colorType);
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t *line = lines.data() + (size_t)y * stride;
        for (uint32_t x = 0; x < w; x++) {
            uint8_t *dst = rgb.data() + 3 * ((size_t)(y0 + y * dy) * info.width + (x0 + x * dx));
            if (info.colorType == 3) {
                const int index = sampleAt(line, x, info.depth, false);
                if (index >= info.paletteSize) { throw std::runtime_error("Error loading texture: PNG palette index out of range"); }
                dst[0] = info.palette[index][0]; dst[1] = info.palette[index][1]; dst[2] = info.palette[index][2];
            } else if (channels <= 2) {
                const int g = sampleAt(line, (size_t)x * channels, info.depth, true);
                dst[0] = dst[1] = dst[2] = (uint8_t)g;
            } else {
                for (int c = 0; c < 3; c++) { dst[c] = (uint8_t)sampleAt(line, (size_t)x * channels + c, info.depth, false); }
            }
        }
    }
}

void decodePng(const std::vector<uint8_t> &file, std::vector<uint8_t> &rgb, int &width, int &height)
{
    PngInfo info;
    std::vector<uint8_t> idat;
    size_t pos = 8;
    bool sawHeader = false, sawEnd = false;
    while (!sawEnd && pos + 12 <= file.size()) {
        const uint32_t length = be32(&file[pos]);
        const uint8_t *type = &file[pos + 4], *data = &file[pos + 8];
        if (pos + 12 + (size_t)length > file.size()) { throw std::runtime_error("Error loading texture: truncated PNG"); }
        if (!memcmp(type, "IHDR", 4)) {
            if (length != 13) { throw std::runtime_error("Error loading texture: bad IHDR"); }
            info.width = be32(data); info.height = be32(data + 4);
            info.depth = data[8]; info.colorType = data[9]; info.interlace = data[12];
            if (data[10] != 0 || data[11] != 0 || info.interlace > 1) { throw std::runtime_error("Error loading texture: unsupported PNG method"); }
            sawHeader = true;
        } else if (!memcmp(type, "PLTE", 4)) {
            info.paletteSize = (int)(length / 3);
            if (info.paletteSize > 256) { throw std::runtime_error("Error loading texture: bad PLTE"); }
            memcpy(info.palette, data, (size_t)info.paletteSize * 3);
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + length);
        } else if (!memcmp(type, "IEND", 4)) {
            sawEnd = true;
        }
        pos += 12 + (size_t)length;
    }
    const bool depthOk = info.depth == 8 || info.depth == 16 || ((info.colorType == 0 || info.colorType == 3) && (info.depth == 1 || info.depth == 2 || info.depth == 4));
    if (!sawHeader || info.width == 0 || info.height == 0 || !depthOk || (info.colorType == 3 && info.depth == 16) ||
        !(info.colorType == 0 || info.colorType == 2 || info.colorType == 3 || info.colorType == 4 || info.colorType == 6)) {
        throw std::runtime_error("Error loading texture: unsupported PNG");
    }
    const int bitsPerPixel = channelsOf(info.colorType) * info.depth;
    const size_t bpp = (size_t)(bitsPerPixel + 7) / 8;
    auto strideOf = [&](uint32_t w) { return ((size_t)w * bitsPerPixel + 7) / 8; };

    // size of the filtered stream
    static const uint32_t xOrigin[7] = {0, 4, 0, 2, 0, 1, 0}, yOrigin[7] = {0, 0, 4, 0, 2, 0, 1}, xStep[7] = {8, 8, 4, 4, 2, 2, 1}, yStep[7] = {8, 8, 8, 4, 4, 2, 2};
    size_t expected = 0;
    if (!info.interlace) { expected = (size_t)info.height * (strideOf(info.width) + 1); }
    else {
        for (int p = 0; p < 7; p++) {
            const uint32_t w = (info.width - xOrigin[p] + xStep[p] - 1) / xStep[p], h = (info.height - yOrigin[p] + yStep[p] - 1) / yStep[p];
            if (w && h) { expected += (size_t)h * (strideOf(w) + 1); }
        }
    }
    std::vector<uint8_t> raw(expected);
    uLongf produced = (uLongf)expected;
    const int zr = uncompress(raw.data(), &produced, idat.data(), (uLong)idat.size());
    if ((zr != Z_OK && zr != Z_BUF_ERROR) || produced < expected) { throw std::runtime_error("Error loading texture: PNG inflate failed"); }

    rgb.assign((size_t)info.width * info.height * 3, 0);
    std::vector<uint8_t> lines;
    if (!info.interlace) {
        unfilter(raw.data(), strideOf(info.width), bpp, info.height, lines);
        storePixels(info, lines, strideOf(info.width), info.width, info.height, 0, 0, 1, 1, rgb);
    } else {
        size_t offset = 0;
        for (int p = 0; p < 7; p++) {
            const uint32_t w = (info.width - xOrigin[p] + xStep[p] - 1) / xStep[p], h = (info.height - yOrigin[p] + yStep[p] - 1) / yStep[p];
            if (!w || !h) { continue; }
            unfilter(raw.data() + offset, strideOf(w), bpp, h, lines);
            storePixels(info, lines, strideOf(w), w, h, xOrigin[p], yOrigin[p], xStep[p], yStep[p], rgb);
            offset += (size_t)h * (strideOf(w) + 1);
        }
    }
    width = (int)info.width; height = (int)info.height;
}

void decodePnm(const std::vector<uint8_t> &file, std::vector<uint8_t> &rgb, int &width, int &height)
{
    size_t pos = 2;
    auto nextInt = [&]() {
        for (;;) { // whitespace and comments
            while (pos < file.size() && (file[pos] == ' ' || file[pos] == '\t' || file[pos] == '\n' || file[pos] == '\r')) { pos++; }
            if (pos < file.size() && file[pos] == '#') { while (pos < file.size() && file[pos] != '\n' && file[pos] != '\r') { pos++; } }
            else { break; }
        }
        long v = 0; bool any = false;
        while (pos < file.size() && file[pos] >= '0' && file[pos] <= '9') { v = v * 10 + (file[pos] - '0'); pos++; any = true; }
        if (!any || v > (1 << 24)) { throw std::runtime_error("Error loading texture: bad PNM header"); }
        return (int)v;
    };
    const int channels = file[1] == '6' ? 3 : 1;
    width = nextInt(); height = nextInt();
    const int maxValue = nextInt();
    pos++; // the single whitespace byte after maxval
    if (width <= 0 || height <= 0 || maxValue <= 0 || maxValue > 255) { throw std::runtime_error("Error loading texture: unsupported PNM"); }
    const size_t count = (size_t)width * height;
    if (pos + count * channels > file.size()) { throw std::runtime_error("Error loading texture: truncated PNM"); }
    rgb.resize(count * 3);
    for (size_t i = 0; i < count; i++) {
        for (int c = 0; c < 3; c++) { rgb[3 * i + c] = file[pos + i * channels + (channels == 3 ? c : 0)]; }
    }
}

} // namespace

void loadImageRGB8(const std::string &path, std::vector<uint8_t> &rgb, int &width, int &height)
{
    const std::vector<uint8_t> file = readFile(path);
    static const uint8_t pngSignature[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1A, '\n'};
    if (file.size() >= 8 && !memcmp(file.data(), pngSignature, 8)) { decodePng(file, rgb, width, height); return; }
    if (file.size() >= 3 && file[0] == 'P' && (file[1] == '5' || file[1] == '6')) { decodePnm(file, rgb, width, height); return; }
    if (file.size() >= 2 && file[0] == 0xFF && file[1] == 0xD8) {
        throw std::runtime_error("Error loading texture: " + path + " is a JPEG; JPEG decoding is decoder-defined (the reference's texels come "
                                 "from stb_image's IDCT), convert it to PNG or hand the decoded texels to ptc_add_texture");
    }
    throw std::runtime_error("Error loading texture: unknown image format: " + path);
}

} // namespace pathed
