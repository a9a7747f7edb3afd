// writers.hpp -- dependency-free output writers for the CLI / tutorials:
//   * trajectory TXT in the reference CLI's format (CLI/main.cpp:239-258)
//   * legacy-VTK ASCII polylines (stand-in for VTKFileManager::SaveTrajectoryLinesAsVTP, which needs libvtk)
//   * PNG of one image channel through a viridis-like colour ramp, NaN = transparent (stand-in for the
//     reference's stb-based SaveToPNG, src/Common/ImageBuffer.hpp:94-137); stored-deflate, no zlib needed
#pragma once
#include "api/MOPS.h"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

namespace MOPS {
namespace writers {

inline bool SaveTrajectoryLinesAsTXT(const std::vector<TrajectoryLine>& lines, const std::string& path)
{
    std::ofstream out(path);
    if (!out.is_open()) return false;
    out << "Line_Index Point_Index Position_X Position_Y Position_Z Velocity_X Velocity_Y Velocity_Z\n";
    for (const auto& line : lines)
        for (size_t k = 0; k < line.points.size(); ++k) {
            const auto& p = line.points[k];
            const auto& v = line.velocity[k];
            out << line.lineID << " " << k << " " << p.x() << " " << p.y() << " " << p.z() << " " << v.x() << " " << v.y() << " " << v.z() << "\n";
        }
    return true;
}

inline bool SaveTrajectoryLinesAsVTK(const std::vector<TrajectoryLine>& lines, const std::string& path)
{
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) return false;
    size_t npts = 0;
    for (auto& l : lines) npts += l.points.size();
    std::fprintf(f, "# vtk DataFile Version 3.0\nMOPS trajectory lines\nASCII\nDATASET POLYDATA\nPOINTS %zu double\n", npts);
    for (auto& l : lines)
        for (auto& p : l.points) std::fprintf(f, "%.17g %.17g %.17g\n", p.x(), p.y(), p.z());
    std::fprintf(f, "LINES %zu %zu\n", lines.size(), npts + lines.size());
    size_t base = 0;
    for (auto& l : lines) {
        std::fprintf(f, "%zu", l.points.size());
        for (size_t k = 0; k < l.points.size(); ++k) std::fprintf(f, " %zu", base + k);
        std::fprintf(f, "\n");
        base += l.points.size();
    }
    std::fprintf(f, "POINT_DATA %zu\nVECTORS velocity double\n", npts);
    for (auto& l : lines)
        for (size_t k = 0; k < l.points.size(); ++k) {
            const vec3 v = k < l.velocity.size() ? l.velocity[k] : vec3();
            std::fprintf(f, "%.17g %.17g %.17g\n", v.x(), v.y(), v.z());
        }
    std::fprintf(f, "CELL_DATA %zu\nSCALARS lineID int 1\nLOOKUP_TABLE default\n", lines.size());
    for (auto& l : lines) std::fprintf(f, "%d\n", l.lineID);
    std::fclose(f);
    return true;
}

namespace detail {
inline uint32_t crc32(const unsigned char* d, size_t n, uint32_t crc = 0)
{
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ d[i]) & 0xFF] ^ (crc >> 8);
    return ~crc;
}
inline void be32(std::vector<unsigned char>& v, uint32_t x)
{
    v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x);
}
inline void chunk(std::vector<unsigned char>& png, const char* type, const std::vector<unsigned char>& data)
{
    be32(png, static_cast<uint32_t>(data.size()));
    std::vector<unsigned char> td(type, type + 4);
    td.insert(td.end(), data.begin(), data.end());
    png.insert(png.end(), td.begin(), td.end());
    be32(png, crc32(td.data(), td.size()));
}
// smooth dark-blue -> green -> yellow ramp (a polynomial fit in the spirit of viridis)
inline void ramp(float t, unsigned char* rgb)
{
    t = t < 0.f ? 0.f : (t > 1.f ? 1.f : t);
    const float r = 0.267f + t * (0.005f + t * (-1.9f + t * 2.62f));
    const float g = 0.005f + t * (1.40f + t * (-0.80f + t * 0.30f));
    const float b = 0.329f + t * (1.39f + t * (-3.70f + t * 2.12f));
    auto q = [](float x) { x = x < 0.f ? 0.f : (x > 1.f ? 1.f : x); return static_cast<unsigned char>(x * 255.0f); };
    rgb[0] = q(r); rgb[1] = q(g); rgb[2] = q(b);
}
} // namespace detail

// one channel (0..3) of an RGBA-double image -> colour-mapped RGBA PNG; NaN pixels transparent
inline bool SaveToPNG(const ImageBuffer<double>& img, const std::string& path, int channel = 2)
{
    const int w = img.getWidth(), h = img.getHeight();
    if (w <= 0 || h <= 0 || channel < 0 || channel > 3) return false;
    float lo = 3.4e38f, hi = -3.4e38f;
    for (size_t p = 0; p < static_cast<size_t>(w) * h; ++p) {
        const float v = static_cast<float>(img.mPixels[4 * p + channel]);
        if (!std::isnan(v)) { lo = v < lo ? v : lo; hi = v > hi ? v : hi; }
    }
    if (lo >= hi) hi = lo + 1e-5f;
    std::vector<unsigned char> raw;
    raw.reserve(static_cast<size_t>(h) * (1 + 4 * static_cast<size_t>(w)));
    for (int i = 0; i < h; ++i) {
        raw.push_back(0); // filter: none
        for (int j = 0; j < w; ++j) {
            const float v = static_cast<float>(img.mPixels[(static_cast<size_t>(i) * w + j) * 4 + channel]);
            unsigned char px[4] = {0, 0, 0, 0};
            if (!std::isnan(v)) { detail::ramp((v - lo) / (hi - lo), px); px[3] = 255; }
            raw.insert(raw.end(), px, px + 4);
        }
    }
    std::vector<unsigned char> z = {0x78, 0x01}; // zlib header, then stored (uncompressed) deflate blocks
    uint32_t a = 1, b = 0;
    for (unsigned char c : raw) { a = (a + c) % 65521u; b = (b + a) % 65521u; }
    for (size_t off = 0; off < raw.size(); off += 65535) {
        const size_t n = std::min<size_t>(65535, raw.size() - off);
        z.push_back(off + n >= raw.size() ? 1 : 0);
        z.push_back(n & 0xFF); z.push_back(n >> 8); z.push_back(~n & 0xFF); z.push_back((~n >> 8) & 0xFF);
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
    }
    detail::be32(z, (b << 16) | a);
    std::vector<unsigned char> png = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::vector<unsigned char> ihdr;
    detail::be32(ihdr, static_cast<uint32_t>(w)); detail::be32(ihdr, static_cast<uint32_t>(h));
    ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    detail::chunk(png, "IHDR", ihdr);
    detail::chunk(png, "IDAT", z);
    detail::chunk(png, "IEND", {});
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = std::fwrite(png.data(), 1, png.size(), f) == png.size();
    std::fclose(f);
    return ok;
}

} // namespace writers
} // namespace MOPS
