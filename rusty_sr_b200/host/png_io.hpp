// png_io.hpp — minimal PNG reader / writer (zlib only), enough for the CLI's host I/O.
// The reference uses the `image` crate for this (src/main.rs:164 image::open, :175 .save); image
// decode/encode is host-side plumbing, not part of the accelerated path.
// Reader: 8-bit non-interlaced grey / grey+alpha / RGB / RGBA / palette -> RGBA8.  Writer: RGBA8.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace rsr_host {

struct Image {
    int w = 0, h = 0;
    std::vector<uint8_t> rgba;  // [h][w][4]
};

inline uint32_t be32(const uint8_t* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

inline Image read_png(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("Error opening input image file.");  // main.rs:164
    std::vector<uint8_t> buf;
    uint8_t tmp[65536];
    size_t n;
    while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    fclose(f);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (buf.size() < 8 || memcmp(buf.data(), sig, 8) != 0) throw std::runtime_error("not a PNG file: " + path);
    int w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    for (size_t off = 8; off + 12 <= buf.size();) {
        const uint32_t len = be32(&buf[off]);
        const char* type = reinterpret_cast<const char*>(&buf[off + 4]);
        const uint8_t* data = &buf[off + 8];
        if (off + 12 + len > buf.size()) throw std::runtime_error("truncated PNG");
        if (crc32(crc32(0L, Z_NULL, 0), &buf[off + 4], 4 + len) != be32(&buf[off + 8 + len])) throw std::runtime_error("PNG chunk CRC mismatch");
        if (!memcmp(type, "IHDR", 4)) {
            if (len != 13) throw std::runtime_error("malformed PNG: IHDR length is not 13");
            w = (int)be32(data); h = (int)be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
        } else if (!memcmp(type, "PLTE", 4)) plte.assign(data, data + len);
        else if (!memcmp(type, "tRNS", 4)) trns.assign(data, data + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!memcmp(type, "IEND", 4)) break;
        off += 12 + len;
    }
    if (w <= 0 || h <= 0) throw std::runtime_error("PNG without IHDR");
    const int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!ch) throw std::runtime_error("unsupported PNG colour type");
    const bool sub_byte = (ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4);
    if ((depth != 8 && !sub_byte) || interlace != 0)
        throw std::runtime_error("unsupported PNG (8-bit, or 1/2/4-bit grey/palette; non-interlaced)");
    const size_t stride = sub_byte ? ((size_t)w * depth + 7) / 8 : (size_t)w * ch;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf rawlen = raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), idat.size()) != Z_OK || rawlen != raw.size())
        throw std::runtime_error("PNG inflate failed");
    std::vector<uint8_t> px(stride * h);
    const int fch = ch;  // filter distance in bytes (1 for sub-byte depths)
    (void)fch;
    for (int y = 0; y < h; ++y) {
        const uint8_t ft = raw[(stride + 1) * y];
        const uint8_t* in = &raw[(stride + 1) * y + 1];
        uint8_t* out = &px[stride * y];
        const uint8_t* up = y ? &px[stride * (y - 1)] : nullptr;
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= (size_t)ch ? out[i - ch] : 0, b = up ? up[i] : 0, c = (up && i >= (size_t)ch) ? up[i - ch] : 0;
            int v = in[i];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: {
                    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
                    v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    break;
                }
                default: throw std::runtime_error("bad PNG filter");
            }
            out[i] = (uint8_t)v;
        }
    }
    if (sub_byte) {  // expand packed samples to one byte each (grey is scaled to 0..255)
        std::vector<uint8_t> wide((size_t)w * h);
        const int maxv = (1 << depth) - 1;
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                const size_t bit = (size_t)x * depth;
                const int v = (px[stride * y + bit / 8] >> (8 - depth - bit % 8)) & maxv;
                wide[(size_t)y * w + x] = (uint8_t)(ctype == 0 ? v * 255 / maxv : v);
            }
        px.swap(wide);
    }
    Image img;
    img.w = w; img.h = h; img.rgba.resize((size_t)w * h * 4);
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        uint8_t r, g, b, a = 255;
        const uint8_t* s = &px[i * ch];
        if (ctype == 0) { r = g = b = s[0]; }
        else if (ctype == 4) { r = g = b = s[0]; a = s[1]; }
        else if (ctype == 2) { r = s[0]; g = s[1]; b = s[2]; }
        else if (ctype == 6) { r = s[0]; g = s[1]; b = s[2]; a = s[3]; }
        else {
            const size_t k = s[0];
            if (3 * k + 2 >= plte.size()) throw std::runtime_error("PNG palette index out of range");
            r = plte[3 * k]; g = plte[3 * k + 1]; b = plte[3 * k + 2];
            a = k < trns.size() ? trns[k] : 255;
        }
        img.rgba[4 * i] = r; img.rgba[4 * i + 1] = g; img.rgba[4 * i + 2] = b; img.rgba[4 * i + 3] = a;
    }
    return img;
}

inline void put_chunk(std::vector<uint8_t>& out, const char* type, const uint8_t* data, size_t len) {
    uint8_t hdr[8] = {(uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len,
                      (uint8_t)type[0], (uint8_t)type[1], (uint8_t)type[2], (uint8_t)type[3]};
    out.insert(out.end(), hdr, hdr + 8);
    if (len) out.insert(out.end(), data, data + len);
    uLong crc = crc32(0L, reinterpret_cast<const Bytef*>(type), 4);
    if (len) crc = crc32(crc, data, (uInt)len);
    const uint8_t c[4] = {(uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc};
    out.insert(out.end(), c, c + 4);
}

inline void write_png_rgba(const std::string& path, const uint8_t* rgba, int w, int h) {
    std::vector<uint8_t> raw(((size_t)w * 4 + 1) * h);
    for (int y = 0; y < h; ++y) {
        raw[((size_t)w * 4 + 1) * y] = 0;  // filter: none
        memcpy(&raw[((size_t)w * 4 + 1) * y + 1], rgba + (size_t)y * w * 4, (size_t)w * 4);
    }
    uLongf clen = compressBound(raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), raw.size(), 6) != Z_OK) throw std::runtime_error("PNG deflate failed");
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    uint8_t ihdr[13] = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w,
                        (uint8_t)(h >> 24), (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h, 8, 6, 0, 0, 0};
    put_chunk(out, "IHDR", ihdr, 13);
    put_chunk(out, "IDAT", comp.data(), clen);
    put_chunk(out, "IEND", nullptr, 0);
    FILE* f = fopen(path.c_str(), "wb");
    if (!f || fwrite(out.data(), 1, out.size(), f) != out.size()) {
        if (f) fclose(f);
        throw std::runtime_error("Could not write output file");  // main.rs:175
    }
    fclose(f);
}

}  // namespace rsr_host
