// Image: fp32 raw buffer (stored top scanline first), 8-bit gamma preview, EXR output.
// Behaviour follows /root/reference/src/image.cpp:13-160 (Q14: row 0 is the bottom scanline; EXR = 3 x HALF, B,G,R, uncompressed;
// auto.exr + auto-%05dspp.exr checkpoints).
#include "pathed.hpp"

#include "exr_io.hpp"

#include <cmath>
#include <cstdio>
#include <algorithm>
#include <fstream>
#include <memory>
#include <thread>
#include <vector>

namespace pathed {

Image::Image(int width, int height)
    : m_height(height), m_width(width), m_spp(0), m_data((size_t)3 * height * width), m_raw((size_t)3 * height * width)
{}

void Image::set(int row, int col, float r, float g, float b)
{
    const size_t flipped = 3 * ((size_t)(m_height - row - 1) * m_width + col);
    m_raw[flipped + 0] = r; m_raw[flipped + 1] = g; m_raw[flipped + 2] = b;
    const size_t index = 3 * ((size_t)row * m_width + col);
    // powf(x, 1/2.2) with a double exponent narrowed by the call, then truncation to a byte
    m_data[index + 0] = (unsigned char)(fminf(powf(r, 1 / 2.2), 1.f) * 255);
    m_data[index + 1] = (unsigned char)(fminf(powf(g, 1 / 2.2), 1.f) * 255);
    m_data[index + 2] = (unsigned char)(fminf(powf(b, 1 / 2.2), 1.f) * 255);
}

void Image::setAll(const float *rgb)
{
    const int threads = std::max(1, std::min(m_height, (int)std::thread::hardware_concurrency()));
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) {
        pool.emplace_back([this, rgb, t, threads]() {
            for (int row = t; row < m_height; row += threads) {
                const float *line = rgb + 3 * (size_t)row * m_width;
                for (int col = 0; col < m_width; col++) { set(row, col, line[3 * col], line[3 * col + 1], line[3 * col + 2]); }
            }
        });
    }
    for (std::thread &t : pool) { t.join(); }
}

void Image::save(const std::string &filestem) { save(filestem, false); }
void Image::saveCheckpoint(const std::string &filestem) { save(filestem, true); }

// The files are what the reference writes (src/image.cpp:80-154: <stem>.exr, and <stem>-%05dspp.exr for a checkpoint -- the same
// bytes twice); the image is encoded here and now, the bytes go to disk on a writer thread, in the order of the calls (auto.exr is
// rewritten by every checkpoint), so that the render loop does not wait for the file system.  flush() / the destructor wait.
void Image::save(const std::string &filestem, bool checkpoint)
{
    const size_t n = (size_t)m_width * m_height;
    std::vector<float> planes[3];
    for (int c = 0; c < 3; c++) { planes[c].resize(n); }
    for (size_t i = 0; i < n; i++) { for (int c = 0; c < 3; c++) { planes[c][i] = m_raw[3 * i + c]; } }
    const std::string directory = g_job ? g_job->outputDirectory() : std::string();
    std::vector<std::string> paths = {directory + filestem + ".exr"};
    if (checkpoint) {
        char suffix[32];
        snprintf(suffix, sizeof(suffix), "-%05dspp.exr", m_spp);
        paths.push_back(directory + filestem + suffix);
    }
    auto bytes = std::make_shared<std::vector<unsigned char>>(encodeEXR(m_width, m_height, {"B", "G", "R"}, {planes[2].data(), planes[1].data(), planes[0].data()}, true));
    std::shared_future<void> previous = m_lastWrite;
    m_lastWrite = std::async(std::launch::async, [previous, bytes, paths]() {
        if (previous.valid()) { previous.wait(); }
        for (const std::string &path : paths) {
            try {
                writeFileBytes(path, *bytes);
                printf("Saved exr file. [ %s ] \n", path.c_str());
            } catch (const std::exception &e) { fprintf(stderr, "Save EXR err: %s\n", e.what()); }
        }
    }).share();
}

void Image::flush()
{
    if (m_lastWrite.valid()) { m_lastWrite.wait(); }
}

Image::~Image() { flush(); }

void Image::write(const std::string &filename)
{
    // stbi_write_bmp(path, w, h, 3, data): 24-bit BMP, bottom-up rows of B,G,R padded to 4 bytes
    const std::string path = (g_job ? g_job->outputDirectory() : std::string()) + filename;
    const int pad = (4 - (m_width * 3) % 4) % 4;
    const uint32_t dataSize = (uint32_t)((m_width * 3 + pad) * m_height), fileSize = 54 + dataSize;
    std::ofstream out(path, std::ios::binary);
    auto u16 = [&](uint16_t v) { out.write((const char *)&v, 2); };
    auto u32 = [&](uint32_t v) { out.write((const char *)&v, 4); };
    out.write("BM", 2); u32(fileSize); u16(0); u16(0); u32(54);
    u32(40); u32((uint32_t)m_width); u32((uint32_t)m_height); u16(1); u16(24); u32(0); u32(0); u32(0); u32(0); u32(0); u32(0);
    const char zero[3] = {0, 0, 0};
    for (int row = m_height - 1; row >= 0; row--) {
        for (int col = 0; col < m_width; col++) {
            const unsigned char *px = &m_data[3 * ((size_t)row * m_width + col)];
            const char bgr[3] = {(char)px[2], (char)px[1], (char)px[0]};
            out.write(bgr, 3);
        }
        out.write(zero, pad);
    }
}

} // namespace pathed
