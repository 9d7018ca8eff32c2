#include "exr_io.hpp"

#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <thread>

namespace pathed {

// tinyexr's float_to_half_full (vendor/tinyexr.h:7164-7199), which SaveEXRImageToFile applies to every HALF channel the reference
// writes: the mantissa is truncated and then incremented when the highest dropped bit is set (ties go up, not to even), floats with a zero
// exponent field become signed zeros.  The files of this writer equal the reference's byte for byte (tests/test_gpu_host.py).
unsigned short floatToHalf(float value)
{
    uint32_t x; memcpy(&x, &value, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const uint32_t biased = (x >> 23) & 0xFFu;
    const uint32_t mantissa = x & 0x7FFFFFu;
    if (biased == 0) { return (unsigned short)sign; }
    if (biased == 0xFF) { return (unsigned short)(sign | 0x7C00u | (mantissa ? 0x200u : 0u)); }
    const int32_t exponent = (int32_t)biased - 127 + 15;
    if (exponent >= 31) { return (unsigned short)(sign | 0x7C00u); }
    if (exponent <= 0) {
        if (14 - exponent > 24) { return (unsigned short)sign; }
        const uint32_t full = mantissa | 0x800000u;
        uint32_t half = full >> (14 - exponent);
        if ((full >> (13 - exponent)) & 1u) { half++; }
        return (unsigned short)(sign | half);
    }
    uint32_t half = ((uint32_t)exponent << 10) | (mantissa >> 13);
    if (mantissa & 0x1000u) { half++; } // may carry into the exponent, up to infinity
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

// ---- PIZ (OpenEXR's wavelet + Huffman codec; tinyexr decodes it for the reference, e.g. test_scenes/1_pixel_test.exr).
// Restated from the published format (OpenEXR ImfPizCompressor / ImfHuf / ImfWav): per block a 16-bit value bitmap, a canonical
// Huffman stream of 16-bit symbols with a run-length symbol, and a two-dimensional Haar-like wavelet per channel component.
namespace piz {
const int kEncBits = 16, kEncSize = (1 << kEncBits) + 1;
const int kShortZeroRun = 59, kLongZeroRun = 63, kShortestLongRun = 2 + kLongZeroRun - kShortZeroRun;

struct BitReader {
    const unsigned char *p, *end; uint64_t c = 0; int lc = 0;
    BitReader(const unsigned char *b, const unsigned char *e) : p(b), end(e) {}
    uint32_t bits(int n)
    {
        while (lc < n) { if (p >= end) { throw std::runtime_error("exr: truncated PIZ block"); } c = (c << 8) | *p++; lc += 8; }
        lc -= n;
        return (uint32_t)((c >> lc) & ((1ull << n) - 1));
    }
};

// code lengths of the symbols im .. iM (6 bits each, runs of zeros packed), then the canonical codes
void unpackLengths(BitReader &br, int im, int iM, std::vector<uint64_t> &hcode)
{
    for (; im <= iM; im++) {
        const uint32_t l = br.bits(6);
        hcode[(size_t)im] = l;
        if (l == (uint32_t)kLongZeroRun) {
            int run = (int)br.bits(8) + kShortestLongRun;
            if (im + run > iM + 1) { throw std::runtime_error("exr: bad PIZ code table"); }
            while (run--) { hcode[(size_t)im++] = 0; }
            im--;
        } else if (l >= (uint32_t)kShortZeroRun) {
            int run = (int)l - kShortZeroRun + 2;
            if (im + run > iM + 1) { throw std::runtime_error("exr: bad PIZ code table"); }
            while (run--) { hcode[(size_t)im++] = 0; }
            im--;
        }
    }
}

void huffmanDecode(const unsigned char *in, size_t nIn, std::vector<uint16_t> &out)
{
    if (nIn < 20) { throw std::runtime_error("exr: truncated PIZ block"); }
    uint32_t h[5];
    memcpy(h, in, 20);
    const int im = (int)h[0], iM = (int)h[1];
    const uint64_t nBits = h[3];
    if (im < 0 || im >= kEncSize || iM < 0 || iM >= kEncSize || im > iM) { throw std::runtime_error("exr: bad PIZ header"); }
    BitReader table(in + 20, in + nIn);
    std::vector<uint64_t> hcode((size_t)kEncSize, 0);
    unpackLengths(table, im, iM, hcode);
    const unsigned char *data = table.p; // the table ends on a byte boundary of what was consumed
    if (nBits > 8 * (uint64_t)(in + nIn - data)) { throw std::runtime_error("exr: bad PIZ bit count"); }
    // canonical code: symbols of one length get consecutive codes in symbol order; first code per length from the length histogram
    uint64_t count[59] = {0}, first[59];
    for (int i = im; i <= iM; i++) { if (hcode[(size_t)i] > 58) { throw std::runtime_error("exr: bad PIZ code length"); } count[hcode[(size_t)i]]++; }
    uint64_t c = 0;
    for (int l = 58; l > 0; l--) { const uint64_t nc = (c + count[l]) >> 1; first[l] = c; c = nc; }
    std::vector<std::vector<uint32_t>> symbols(59);
    for (int i = im; i <= iM; i++) { const uint64_t l = hcode[(size_t)i]; if (l) { symbols[(size_t)l].push_back((uint32_t)i); } }
    BitReader br(data, in + nIn);
    size_t produced = 0;
    uint64_t consumed = 0;
    while (consumed < nBits && produced < out.size()) {
        uint64_t code = 0; int len = 0; uint32_t symbol = 0; bool found = false;
        while (len < 58) {
            code = (code << 1) | br.bits(1); len++; consumed++;
            if (count[len] && code >= first[len] && code - first[len] < count[len]) { symbol = symbols[(size_t)len][(size_t)(code - first[len])]; found = true; break; }
        }
        if (!found) { throw std::runtime_error("exr: bad PIZ code"); }
        if (symbol == (uint32_t)iM) { // run-length symbol: repeat the previous value
            const uint32_t run = br.bits(8); consumed += 8;
            if (produced == 0 || produced + run > out.size()) { throw std::runtime_error("exr: bad PIZ run"); }
            const uint16_t v = out[produced - 1];
            for (uint32_t k = 0; k < run; k++) { out[produced++] = v; }
        } else { out[produced++] = (uint16_t)symbol; }
    }
    if (produced != out.size()) { throw std::runtime_error("exr: short PIZ block"); }
}

inline void wdec14(uint16_t l, uint16_t h, uint16_t &a, uint16_t &b)
{
    const short ls = (short)l, hs = (short)h;
    const int hi = hs, ai = ls + (hi & 1) + (hi >> 1);
    a = (uint16_t)(short)ai; b = (uint16_t)(short)(ai - hi);
}
inline void wdec16(uint16_t l, uint16_t h, uint16_t &a, uint16_t &b)
{
    const int m = l, d = h;
    const int bb = (m - (d >> 1)) & 0xFFFF, aa = (d + bb - 0x8000) & 0xFFFF;
    b = (uint16_t)bb; a = (uint16_t)aa;
}
void waveletDecode(uint16_t *in, int nx, int ox, int ny, int oy, uint16_t mx)
{
    const bool w14 = mx < (1 << 14);
    const int n = nx > ny ? ny : nx;
    int p = 1;
    while (p <= n) { p <<= 1; }
    p >>= 1;
    int p2 = p;
    p >>= 1;
    while (p >= 1) {
        uint16_t *py = in, *ey = in + (ptrdiff_t)oy * (ny - p2);
        const ptrdiff_t oy1 = (ptrdiff_t)oy * p, oy2 = (ptrdiff_t)oy * p2, ox1 = (ptrdiff_t)ox * p, ox2 = (ptrdiff_t)ox * p2;
        uint16_t i00, i01, i10, i11;
        for (; py <= ey; py += oy2) {
            uint16_t *px = py, *ex = py + (ptrdiff_t)ox * (nx - p2);
            for (; px <= ex; px += ox2) {
                uint16_t *p01 = px + ox1, *p10 = px + oy1, *p11 = p10 + ox1;
                if (w14) { wdec14(*px, *p10, i00, i10); wdec14(*p01, *p11, i01, i11); wdec14(i00, i01, *px, *p01); wdec14(i10, i11, *p10, *p11); }
                else { wdec16(*px, *p10, i00, i10); wdec16(*p01, *p11, i01, i11); wdec16(i00, i01, *px, *p01); wdec16(i10, i11, *p10, *p11); }
            }
            if (nx & p) {
                uint16_t *p10 = px + oy1;
                if (w14) { wdec14(*px, *p10, i00, *p10); } else { wdec16(*px, *p10, i00, *p10); }
                *px = i00;
            }
        }
        if (ny & p) {
            uint16_t *px = py, *ex = py + (ptrdiff_t)ox * (nx - p2);
            for (; px <= ex; px += ox2) {
                uint16_t *p01 = px + ox1;
                if (w14) { wdec14(*px, *p01, i00, *p01); } else { wdec16(*px, *p01, i00, *p01); }
                *px = i00;
            }
        }
        p2 = p;
        p >>= 1;
    }
}

// one block: `lines` scanlines of `width` pixels, channel c has words[c] 16-bit words per pixel; out = the uncompressed block bytes
void decodeBlock(const unsigned char *src, size_t packed, int width, int lines, const std::vector<int> &words, std::vector<unsigned char> &out)
{
    size_t total = 0;
    for (int w : words) { total += (size_t)w * width * lines; }
    if (packed < 4) { throw std::runtime_error("exr: truncated PIZ block"); }
    uint16_t minNonZero, maxNonZero;
    memcpy(&minNonZero, src, 2); memcpy(&maxNonZero, src + 2, 2);
    size_t at = 4;
    std::vector<unsigned char> bitmap(8192, 0);
    if (minNonZero <= maxNonZero) {
        const size_t n = (size_t)maxNonZero - minNonZero + 1;
        if (maxNonZero >= 8192 || at + n > packed) { throw std::runtime_error("exr: bad PIZ bitmap"); }
        memcpy(&bitmap[minNonZero], src + at, n);
        at += n;
    }
    std::vector<uint16_t> lut(65536, 0);
    int k = 0;
    for (int i = 0; i < 65536; i++) { if (i == 0 || (bitmap[(size_t)i >> 3] & (1 << (i & 7)))) { lut[(size_t)k++] = (uint16_t)i; } }
    const uint16_t maxValue = (uint16_t)(k - 1);
    if (at + 4 > packed) { throw std::runtime_error("exr: truncated PIZ block"); }
    int32_t length; memcpy(&length, src + at, 4); at += 4;
    if (length < 0 || at + (size_t)length > packed) { throw std::runtime_error("exr: bad PIZ length"); }
    std::vector<uint16_t> tmp(total);
    huffmanDecode(src + at, (size_t)length, tmp);
    size_t base = 0;
    std::vector<size_t> start(words.size());
    for (size_t c = 0; c < words.size(); c++) {
        start[c] = base;
        for (int j = 0; j < words[c]; j++) { waveletDecode(&tmp[base + (size_t)j], width, words[c], lines, width * words[c], maxValue); }
        base += (size_t)words[c] * width * lines;
    }
    for (uint16_t &v : tmp) { v = lut[v]; }
    out.resize(total * 2);
    size_t o = 0;
    std::vector<size_t> cursor = start;
    for (int y = 0; y < lines; y++) {
        for (size_t c = 0; c < words.size(); c++) {
            const size_t n = (size_t)words[c] * width;
            memcpy(&out[o], &tmp[cursor[c]], n * 2);
            o += n * 2; cursor[c] += n;
        }
    }
}
} // namespace piz

namespace {
// OpenEXR's byte-run-length code (ImfRle): a negative count copies -count literal bytes, a count n >= 0 repeats the next byte n + 1 times
void rleDecode(const unsigned char *src, size_t packed, std::vector<unsigned char> &out, size_t expect)
{
    out.clear(); out.reserve(expect);
    size_t i = 0;
    while (i < packed) {
        const int count = (signed char)src[i++];
        if (count < 0) {
            const size_t n = (size_t)(-count);
            if (i + n > packed || out.size() + n > expect) { throw std::runtime_error("exr: bad RLE block"); }
            out.insert(out.end(), src + i, src + i + n); i += n;
        } else {
            const size_t n = (size_t)count + 1;
            if (i >= packed || out.size() + n > expect) { throw std::runtime_error("exr: bad RLE block"); }
            out.insert(out.end(), n, src[i++]);
        }
    }
    if (out.size() != expect) { throw std::runtime_error("exr: short RLE block"); }
}
// ZIP / RLE post-processing: undo the delta predictor, then interleave the two half-buffers
void unpredict(std::vector<unsigned char> &tmp, std::vector<unsigned char> &raw)
{
    const size_t expect = tmp.size();
    raw.resize(expect);
    for (size_t i = 1; i < expect; i++) { tmp[i] = (unsigned char)(tmp[i - 1] + tmp[i] - 128); }
    const size_t half = (expect + 1) / 2;
    for (size_t i = 0; i < expect; i++) { raw[i] = (i & 1) ? tmp[half + i / 2] : tmp[i / 2]; }
}
} // namespace

void loadEXR(const std::string &path, std::vector<float> &rgba, int &width, int &height)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { throw std::runtime_error("exr: cannot open " + path); }
    std::vector<unsigned char> data;
    fseek(f, 0, SEEK_END); const long size = ftell(f); fseek(f, 0, SEEK_SET);
    data.resize((size_t)std::max(size, 0L));
    if (fread(data.data(), 1, data.size(), f) != data.size()) { fclose(f); throw std::runtime_error("exr: read failed"); }
    fclose(f);

    Reader r(data);
    if (r.i32() != 20000630) { throw std::runtime_error("exr: bad magic in " + path); }
    const int32_t version = r.i32();
    if (version & 0x1E00) { throw std::runtime_error("exr: tiled / multipart / deep files are not supported"); }
    std::vector<Channel> channels;
    int compression = 0, xMin = 0, yMin = 0, xMax = -1, yMax = -1;
    bool decreasingY = false;
    for (;;) {
        const std::string name = r.str();
        if (name.empty()) { break; }
        const std::string type = r.str();
        const int32_t attrSize = r.i32();
        if (attrSize < 0) { throw std::runtime_error("exr: negative attribute size in " + path); }
        r.need((size_t)attrSize);
        const size_t start = r.p;
        if (name == "channels") {
            while (true) {
                const std::string channelName = r.str();
                if (channelName.empty()) { break; }
                Channel ch; ch.name = channelName; ch.type = r.i32();
                r.need(4); r.p += 4; /* pLinear + reserved */
                const int xs = r.i32(), ys = r.i32();
                if (xs != 1 || ys != 1) { throw std::runtime_error("exr: subsampled channels are not supported"); }
                channels.push_back(ch);
            }
        } else if (name == "compression") { r.need(1); compression = data[r.p]; }
        else if (name == "dataWindow") { xMin = r.i32(); yMin = r.i32(); xMax = r.i32(); yMax = r.i32(); }
        else if (name == "lineOrder") { r.need(1); decreasingY = data[r.p] == 1; }
        r.p = start + (size_t)attrSize;
    }
    (void)decreasingY; // blocks carry their own y, so either order of the offset table reads the same
    const int64_t w64 = (int64_t)xMax - xMin + 1, h64 = (int64_t)yMax - yMin + 1;
    if (w64 <= 0 || h64 <= 0 || w64 > (1 << 20) || h64 > (1 << 20) || channels.empty()) { throw std::runtime_error("exr: bad header in " + path); }
    width = (int)w64; height = (int)h64;
    int linesPerBlock;
    if (compression == 0 || compression == 1 || compression == 2) { linesPerBlock = 1; }   // NONE, RLE, ZIPS
    else if (compression == 3) { linesPerBlock = 16; }                                      // ZIP
    else if (compression == 4) { linesPerBlock = 32; }                                      // PIZ
    else {
        throw std::runtime_error("exr: unsupported compression " + std::to_string(compression) + " in " + path +
                                 " (reads NONE, RLE, ZIPS, ZIP and PIZ; re-save PXR24 / B44 / DWA files with one of those)");
    }

    size_t bytesPerLine = 0;
    std::vector<int> words;
    for (const Channel &ch : channels) {
        if (ch.type != 1 && ch.type != 2) { throw std::runtime_error("exr: only HALF and FLOAT channels are supported"); }
        bytesPerLine += (size_t)width * (ch.type == 1 ? 2 : 4);
        words.push_back(ch.type == 1 ? 1 : 2);
    }
    const int blocks = (height + linesPerBlock - 1) / linesPerBlock;
    std::vector<uint64_t> offsets((size_t)blocks);
    for (int b = 0; b < blocks; b++) { offsets[(size_t)b] = r.u64(); }

    rgba.assign((size_t)width * height * 4, 0.f);
    for (size_t i = 0; i < (size_t)width * height; i++) { rgba[4 * i + 3] = 1.f; }
    std::vector<unsigned char> raw, tmp;
    for (int b = 0; b < blocks; b++) {
        if (offsets[(size_t)b] > data.size()) { throw std::runtime_error("exr: block offset outside the file"); }
        Reader c(data); c.p = (size_t)offsets[(size_t)b];
        const int64_t y64 = (int64_t)c.i32() - yMin;
        if (y64 < 0 || y64 >= height) { throw std::runtime_error("exr: block scanline outside the data window"); }
        const int y = (int)y64;
        const int32_t packed = c.i32();
        if (packed < 0) { throw std::runtime_error("exr: negative block size"); }
        c.need((size_t)packed);
        const int lines = std::min(linesPerBlock, height - y);
        const size_t expect = bytesPerLine * (size_t)lines;
        const unsigned char *src = &data[c.p];
        if ((size_t)packed > expect) { throw std::runtime_error("exr: block larger than its scanlines"); }
        if ((size_t)packed < expect) {
            if (compression == 0) { throw std::runtime_error("exr: short uncompressed block"); }
            if (compression == 4) { piz::decodeBlock(src, (size_t)packed, width, lines, words, raw); }
            else {
                if (compression == 1) { rleDecode(src, (size_t)packed, tmp, expect); }
                else {
                    tmp.resize(expect);
                    uLongf outLen = (uLongf)expect;
                    if (uncompress(tmp.data(), &outLen, src, (uLong)packed) != Z_OK || outLen != expect) { throw std::runtime_error("exr: zlib failure"); }
                }
                unpredict(tmp, raw);
            }
            if (raw.size() != expect) { throw std::runtime_error("exr: decoded block has the wrong size"); }
            src = raw.data();
        } // packed == expect: every codec stores the block raw when compressing did not help
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

std::vector<unsigned char> encodeEXR(int width, int height, const std::vector<std::string> &names, const std::vector<const float *> &planes, bool asHalf)
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

    // offset table, then one block per scanline (y, byte count, the channels' rows one after the other): sizes are known up front,
    // so the scanlines are converted in place by all host threads
    const size_t bytesPerLine = names.size() * (size_t)width * (asHalf ? 2 : 4), blockBytes = 8 + bytesPerLine;
    const size_t tableStart = out.size(), dataStart = tableStart + 8 * (size_t)height;
    out.resize(dataStart + blockBytes * (size_t)height);
    unsigned char *base = out.data();
    const int threads = std::max(1, std::min(height, (int)std::thread::hardware_concurrency()));
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) {
        pool.emplace_back([&, t]() {
            for (int y = t; y < height; y += threads) {
                const uint64_t offset = dataStart + blockBytes * (size_t)y;
                memcpy(base + tableStart + 8 * (size_t)y, &offset, 8);
                unsigned char *p = base + offset;
                const int32_t line = y, count = (int32_t)bytesPerLine;
                memcpy(p, &line, 4); memcpy(p + 4, &count, 4); p += 8;
                for (size_t c = 0; c < names.size(); c++) {
                    const float *row = planes[c] + (size_t)y * width;
                    if (asHalf) { for (int x = 0; x < width; x++) { const unsigned short h = floatToHalf(row[x]); memcpy(p, &h, 2); p += 2; } }
                    else { memcpy(p, row, (size_t)width * 4); p += (size_t)width * 4; }
                }
            }
        });
    }
    for (std::thread &t : pool) { t.join(); }
    return out;
}

void writeFileBytes(const std::string &path, const std::vector<unsigned char> &bytes)
{
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) { throw std::runtime_error("exr: cannot write " + path); }
    fwrite(bytes.data(), 1, bytes.size(), f);
    fclose(f);
}

void saveEXR(const std::string &path, int width, int height, const std::vector<std::string> &names,
             const std::vector<const float *> &planes, bool asHalf)
{
    writeFileBytes(path, encodeEXR(width, height, names, planes, asHalf));
}

} // namespace pathed
