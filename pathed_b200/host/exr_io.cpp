#include "exr_io.hpp"

#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace pathed {

unsigned short floatToHalf(float value)
{
    uint32_t x; memcpy(&x, &value, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const int32_t exponent = (int32_t)((x >> 23) & 0xFF) - 127 + 15;
    uint32_t mantissa = x & 0x7FFFFFu;
    if (((x >> 23) & 0xFF) == 0xFF) { return (unsigned short)(sign | 0x7C00u | (mantissa ? 0x200u : 0u)); }
    if (exponent >= 31) { return (unsigned short)(sign | 0x7C00u); }
    if (exponent <= 0) {
        if (exponent < -10) { return (unsigned short)sign; }
        mantissa |= 0x800000u;
        const int shift = 14 - exponent;
        uint32_t half = mantissa >> shift;
        const uint32_t rem = mantissa & ((1u << shift) - 1), halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (half & 1))) { half++; }
        return (unsigned short)(sign | half);
    }
    uint32_t half = ((uint32_t)exponent << 10) | (mantissa >> 13);
    const uint32_t rem = mantissa & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (half & 1))) { half++; }
    return (unsigned short)(sign | half);
}

float halfToFloat(unsigned short h)
{
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exponent = (h >> 10) & 0x1F, mantissa = h & 0x3FFu, bits;
    if (exponent == 0) {
        if (mantissa == 0) { bits = sign; }
        else {
            int e = -1;
            do { e++; mantissa <<= 1; } while (!(mantissa & 0x400u));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((mantissa & 0x3FFu) << 13);
        }
    } else if (exponent == 31) { bits = sign | 0x7F800000u | (mantissa << 13); }
    else { bits = sign | ((exponent + 127 - 15) << 23) | (mantissa << 13); }
    float f; memcpy(&f, &bits, 4);
    return f;
}

namespace {
struct Channel { std::string name; int type; };

struct Reader {
    const std::vector<unsigned char> &d; size_t p = 0;
    explicit Reader(const std::vector<unsigned char> &data) : d(data) {}
    void need(size_t n) const { if (p + n > d.size()) { throw std::runtime_error("exr: truncated file"); } }
    int32_t i32() { need(4); int32_t v; memcpy(&v, &d[p], 4); p += 4; return v; }
    uint64_t u64() { need(8); uint64_t v; memcpy(&v, &d[p], 8); p += 8; return v; }
    std::string str() { std::string s; while (true) { need(1); const char c = (char)d[p++]; if (!c) { break; } s += c; } return s; }
};
} // namespace

void loadEXR(const std::string &path, std::vector<float> &rgba, int &width, int &height)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { throw std::runtime_error("exr: cannot open " + path); }
    std::vector<unsigned char> data;
    fseek(f, 0, SEEK_END); const long size = ftell(f); fseek(f, 0, SEEK_SET);
    data.resize((size_t)size);
    if (fread(data.data(), 1, data.size(), f) != data.size()) { fclose(f); throw std::runtime_error("exr: read failed"); }
    fclose(f);

    Reader r(data);
    if (r.i32() != 20000630) { throw std::runtime_error("exr: bad magic in " + path); }
    const int32_t version = r.i32();
    if (version & 0x1E00) { throw std::runtime_error("exr: tiled / multipart / deep files are not supported"); }
    std::vector<Channel> channels;
    int compression = 0, xMin = 0, yMin = 0, xMax = -1, yMax = -1;
    for (;;) {
        const std::string name = r.str();
        if (name.empty()) { break; }
        const std::string type = r.str();
        const int32_t attrSize = r.i32();
        const size_t start = r.p;
        if (name == "channels") {
            while (true) {
                const std::string channelName = r.str();
                if (channelName.empty()) { break; }
                Channel ch; ch.name = channelName; ch.type = r.i32();
                r.p += 4; /* pLinear + reserved */
                const int xs = r.i32(), ys = r.i32();
                if (xs != 1 || ys != 1) { throw std::runtime_error("exr: subsampled channels are not supported"); }
                channels.push_back(ch);
            }
        } else if (name == "compression") { r.need(1); compression = data[r.p]; }
        else if (name == "dataWindow") { xMin = r.i32(); yMin = r.i32(); xMax = r.i32(); yMax = r.i32(); }
        r.p = start + (size_t)attrSize;
    }
    width = xMax - xMin + 1; height = yMax - yMin + 1;
    if (width <= 0 || height <= 0 || channels.empty()) { throw std::runtime_error("exr: bad header in " + path); }
    int linesPerBlock;
    if (compression == 0 || compression == 2) { linesPerBlock = 1; }
    else if (compression == 3) { linesPerBlock = 16; }
    else { throw std::runtime_error("exr: unsupported compression (only NONE, ZIPS, ZIP) in " + path); }

    size_t bytesPerLine = 0;
    for (const Channel &ch : channels) {
        if (ch.type != 1 && ch.type != 2) { throw std::runtime_error("exr: only HALF and FLOAT channels are supported"); }
        bytesPerLine += (size_t)width * (ch.type == 1 ? 2 : 4);
    }
    const int blocks = (height + linesPerBlock - 1) / linesPerBlock;
    std::vector<uint64_t> offsets(blocks);
    for (int b = 0; b < blocks; b++) { offsets[b] = r.u64(); }

    rgba.assign((size_t)width * height * 4, 0.f);
    for (size_t i = 0; i < (size_t)width * height; i++) { rgba[4 * i + 3] = 1.f; }
    std::vector<unsigned char> raw, tmp;
    for (int b = 0; b < blocks; b++) {
        Reader c(data); c.p = (size_t)offsets[b];
        const int y = c.i32() - yMin;
        const int32_t packed = c.i32();
        c.need((size_t)packed);
        const int lines = std::min(linesPerBlock, height - y);
        const size_t expect = bytesPerLine * (size_t)lines;
        const unsigned char *src = &data[c.p];
        if (compression != 0 && (size_t)packed < expect) {
            tmp.resize(expect); raw.resize(expect);
            uLongf outLen = (uLongf)expect;
            if (uncompress(tmp.data(), &outLen, src, (uLong)packed) != Z_OK || outLen != expect) { throw std::runtime_error("exr: zlib failure"); }
            for (size_t i = 1; i < expect; i++) { tmp[i] = (unsigned char)(tmp[i - 1] + tmp[i] - 128); }
            const size_t half = (expect + 1) / 2;
            for (size_t i = 0; i < expect; i++) { raw[i] = (i & 1) ? tmp[half + i / 2] : tmp[i / 2]; }
            src = raw.data();
        }
        for (int line = 0; line < lines; line++) {
            const unsigned char *p = src + bytesPerLine * (size_t)line;
            for (const Channel &ch : channels) {
                int slot = -1;
                if (ch.name == "R") { slot = 0; } else if (ch.name == "G") { slot = 1; } else if (ch.name == "B") { slot = 2; } else if (ch.name == "A") { slot = 3; }
                for (int x = 0; x < width; x++) {
                    float v;
                    if (ch.type == 1) { unsigned short h; memcpy(&h, p + 2 * (size_t)x, 2); v = halfToFloat(h); }
                    else { memcpy(&v, p + 4 * (size_t)x, 4); }
                    if (slot >= 0) { rgba[4 * ((size_t)(y + line) * width + x) + slot] = v; }
                }
                p += (size_t)width * (ch.type == 1 ? 2 : 4);
            }
        }
    }
}

void saveEXR(const std::string &path, int width, int height, const std::vector<std::string> &names,
             const std::vector<const float *> &planes, bool asHalf)
{
    std::vector<unsigned char> out;
    auto i32 = [&](int32_t v) { unsigned char b[4]; memcpy(b, &v, 4); out.insert(out.end(), b, b + 4); };
    auto f32 = [&](float v) { unsigned char b[4]; memcpy(b, &v, 4); out.insert(out.end(), b, b + 4); };
    auto str = [&](const std::string &s) { out.insert(out.end(), s.begin(), s.end()); out.push_back(0); };
    auto attr = [&](const char *name, const char *type, int32_t size) { str(name); str(type); i32(size); };

    i32(20000630); i32(2);
    int32_t chSize = 1;
    for (const std::string &n : names) { chSize += (int32_t)n.size() + 1 + 16; }
    attr("channels", "chlist", chSize);
    for (const std::string &n : names) { str(n); i32(asHalf ? 1 : 2); i32(0); i32(1); i32(1); }
    out.push_back(0);
    attr("compression", "compression", 1); out.push_back(0);
    attr("dataWindow", "box2i", 16); i32(0); i32(0); i32(width - 1); i32(height - 1);
    attr("displayWindow", "box2i", 16); i32(0); i32(0); i32(width - 1); i32(height - 1);
    attr("lineOrder", "lineOrder", 1); out.push_back(0);
    attr("pixelAspectRatio", "float", 4); f32(1.f);
    attr("screenWindowCenter", "v2f", 8); f32(0.f); f32(0.f);
    attr("screenWindowWidth", "float", 4); f32(1.f);
    out.push_back(0);

    const size_t bytesPerLine = names.size() * (size_t)width * (asHalf ? 2 : 4);
    const size_t tableStart = out.size();
    out.resize(out.size() + 8 * (size_t)height);
    for (int y = 0; y < height; y++) {
        const uint64_t offset = out.size();
        memcpy(&out[tableStart + 8 * (size_t)y], &offset, 8);
        i32(y); i32((int32_t)bytesPerLine);
        for (size_t c = 0; c < names.size(); c++) {
            const float *row = planes[c] + (size_t)y * width;
            for (int x = 0; x < width; x++) {
                if (asHalf) { const unsigned short h = floatToHalf(row[x]); unsigned char b[2]; memcpy(b, &h, 2); out.insert(out.end(), b, b + 2); }
                else { f32(row[x]); }
            }
        }
    }
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) { throw std::runtime_error("exr: cannot write " + path); }
    fwrite(out.data(), 1, out.size(), f);
    fclose(f);
}

} // namespace pathed
