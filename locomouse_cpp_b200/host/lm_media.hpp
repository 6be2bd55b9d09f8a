// lm_media.hpp — the reference's own input media, read without OpenCV:
//   video       cv::VideoCapture + extractChannel(F, F, 0)   (LocoMouse_class.cpp:367-400, 1282-1293)
//   background  cv::imread(file, CV_LOAD_IMAGE_GRAYSCALE)    (LocoMouse_class.cpp:402-417)
// Video: AVI (RIFF, also OpenDML 'AVIX' extensions) whose stream 0 holds UNCOMPRESSED frames -- 8-bit grey ('Y800' /
// 'GREY' / 'Y8  ', what cv2.VideoWriter(fourcc = 0, isColor = False) writes), BI_RGB 8-bit with a palette, BI_RGB 24- / 32-bit
// (channel 0 = blue, as extractChannel takes it).  Compressed or YUV-planar streams need a decoder, which is outside the hot
// path: they are refused with a message that says so.  Background: PNG, 8 bits per channel, non-interlaced; colour images are
// reduced to grey as libpng does for OpenCV (png_set_rgb_to_gray with 0.299 / 0.587).  Both are checked against what the
// real OpenCV reads from the same files (tests/test_host_media.py).
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace lmmedia {

inline std::vector<uint8_t> slurp(const std::string &name, size_t limit = 0) {
    FILE *f = std::fopen(name.c_str(), "rb");
    if (!f) throw std::runtime_error("Cannot open file: " + name);
    std::vector<uint8_t> d;
    uint8_t buf[1 << 16];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) {
        d.insert(d.end(), buf, buf + n);
        if (limit && d.size() >= limit) break;
    }
    std::fclose(f);
    return d;
}
inline uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]; }
inline bool is_avi(const std::string &name) {
    const std::vector<uint8_t> h = slurp(name, 12);
    return h.size() >= 12 && !std::memcmp(h.data(), "RIFF", 4) && !std::memcmp(h.data() + 8, "AVI ", 4);
}
inline bool is_png(const std::string &name) {
    const std::vector<uint8_t> h = slurp(name, 8);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    return h.size() >= 8 && !std::memcmp(h.data(), sig, 8);
}

struct Video {
    int n = 0, rows = 0, cols = 0;
    std::vector<uint8_t> frames;  // channel 0 of every frame, [n][rows][cols]
};

// Streams the frame chunks of stream 0 ('00db' / '00dc') in file order.
inline Video read_avi(const std::string &name) {
    FILE *f = std::fopen(name.c_str(), "rb");
    if (!f) throw std::runtime_error("Cannot open file: " + name);
    struct Closer {
        FILE *f;
        ~Closer() { std::fclose(f); }
    } closer{f};
    auto fail = [&](const std::string &why) -> void { throw std::runtime_error("Video file " + name + ": " + why); };
    auto rd = [&](void *dst, size_t n) { return std::fread(dst, 1, n, f) == n; };
    auto skip = [&](uint64_t n) { return fseeko(f, (off_t)n, SEEK_CUR) == 0; };
    uint8_t h[12];
    if (!rd(h, 12) || std::memcmp(h, "RIFF", 4) || std::memcmp(h + 8, "AVI ", 4)) fail("not an AVI file");
    Video V;
    int bits = 0, height_signed = 0;
    uint32_t fourcc = 0;
    uint8_t palette_b[256];
    for (int i = 0; i < 256; ++i) palette_b[i] = (uint8_t)i;
    bool have_format = false, stream0_is_video = false;
    int stream_index = -1;
    std::vector<uint8_t> chunk;
    // A flat walk: LIST / RIFF headers are entered (only their 4-byte type is consumed), every other chunk is read or skipped.
    for (;;) {
        uint8_t ch[8];
        if (!rd(ch, 8)) break;
        const uint32_t size = le32(ch + 4);
        if (!std::memcmp(ch, "LIST", 4) || !std::memcmp(ch, "RIFF", 4)) {
            uint8_t type[4];
            if (!rd(type, 4)) break;
            if (!std::memcmp(type, "strl", 4)) ++stream_index;
            if (!std::memcmp(ch, "LIST", 4) && std::memcmp(type, "hdrl", 4) && std::memcmp(type, "strl", 4) && std::memcmp(type, "movi", 4)) {
                if (!skip((uint64_t)size - 4 + (size & 1))) break;  // INFO, odml ... : not needed
            }
            continue;
        }
        if (!std::memcmp(ch, "strh", 4) && stream_index == 0) {
            chunk.resize(size);
            if (!rd(chunk.data(), size)) fail("truncated stream header");
            stream0_is_video = size >= 4 && !std::memcmp(chunk.data(), "vids", 4);
            if (size & 1) skip(1);
            continue;
        }
        if (!std::memcmp(ch, "strf", 4) && stream_index == 0) {
            chunk.resize(size);
            if (!rd(chunk.data(), size) || size < 40) fail("truncated stream format");
            if (size & 1) skip(1);
            V.cols = (int)le32(chunk.data() + 4);
            height_signed = (int)le32(chunk.data() + 8);
            V.rows = height_signed < 0 ? -height_signed : height_signed;
            bits = chunk[14] | (chunk[15] << 8);
            fourcc = le32(chunk.data() + 16);
            const uint32_t used = le32(chunk.data() + 32);
            const size_t ncol = used ? used : (bits <= 8 ? (size_t)1 << bits : 0);
            for (size_t i = 0; i < ncol && 40 + 4 * i + 3 < size && i < 256; ++i) palette_b[i] = chunk[40 + 4 * i];  // RGBQUAD: blue first
            have_format = true;
            continue;
        }
        const bool frame_chunk = ch[0] == '0' && ch[1] == '0' && ch[2] == 'd' && (ch[3] == 'b' || ch[3] == 'c');
        if (!frame_chunk) {
            if (!skip((uint64_t)size + (size & 1))) break;
            continue;
        }
        if (!have_format || !stream0_is_video || V.rows <= 0 || V.cols <= 0) fail("frame data before a usable video stream format");
        auto cc = [](const char *s) { return le32(reinterpret_cast<const uint8_t *>(s)); };
        const bool grey = fourcc == cc("Y800") || fourcc == cc("GREY") || fourcc == cc("Y8  ") || fourcc == cc("Y8\0\0");
        const bool rgb = fourcc == 0 || fourcc == cc("DIB ") || fourcc == cc("RGB ") || fourcc == cc("RAW ");
        if (!(grey && bits == 8) && !(rgb && (bits == 8 || bits == 24 || bits == 32))) {
            char fc[5] = {(char)(fourcc & 255), (char)((fourcc >> 8) & 255), (char)((fourcc >> 16) & 255), (char)((fourcc >> 24) & 255), 0};
            fail(std::string("stream 0 is '") + fc + "' with " + std::to_string(bits) +
                 " bits per pixel; only uncompressed 8-bit grey (Y800) and BI_RGB 8 / 24 / 32-bit frames are read here. Decoding compressed or "
                 "YUV-planar video is outside this library (convert the video first).");
        }
        if (size == 0) continue;  // dropped frame: VideoCapture repeats nothing, it simply has one frame less
        chunk.resize(size);
        if (!rd(chunk.data(), size)) fail("truncated frame");
        if (size & 1) skip(1);
        const int bpp = bits / 8;
        const size_t tight = (size_t)V.cols * bpp, padded = (tight + 3) & ~(size_t)3;
        size_t stride;
        if (size >= padded * V.rows) stride = padded;
        else if (size >= tight * V.rows) stride = tight;
        else { fail("frame chunk smaller than one image"); return V; }
        const bool bottom_up = rgb && height_signed > 0;  // BI_RGB DIBs are stored bottom-up unless the height is negative
        const size_t base = V.frames.size();
        V.frames.resize(base + (size_t)V.rows * V.cols);
        for (int r = 0; r < V.rows; ++r) {
            const uint8_t *src = chunk.data() + (size_t)(bottom_up ? V.rows - 1 - r : r) * stride;
            uint8_t *dst = V.frames.data() + base + (size_t)r * V.cols;
            if (bpp == 1) {
                if (grey) std::memcpy(dst, src, (size_t)V.cols);
                else for (int c = 0; c < V.cols; ++c) dst[c] = palette_b[src[c]];
            } else {
                for (int c = 0; c < V.cols; ++c) dst[c] = src[(size_t)c * bpp];  // B of BGR(A): channel 0
            }
        }
        ++V.n;
    }
    if (V.n == 0) fail("no video frames found");
    return V;
}

struct Image {
    int rows = 0, cols = 0;
    std::vector<uint8_t> px;
};

inline Image read_png_gray(const std::string &name) {
    const std::vector<uint8_t> d = slurp(name);
    auto fail = [&](const std::string &why) -> void { throw std::runtime_error("Image file " + name + ": " + why); };
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (d.size() < 8 || std::memcmp(d.data(), sig, 8)) fail("not a PNG file");
    Image I;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte;
    for (size_t pos = 8; pos + 12 <= d.size();) {
        const uint32_t len = be32(&d[pos]);
        const uint8_t *type = &d[pos + 4], *data = &d[pos + 8];
        if (pos + 12 + (size_t)len > d.size()) fail("truncated chunk");
        if (!std::memcmp(type, "IHDR", 4) && len >= 13) {
            I.cols = (int)be32(data);
            I.rows = (int)be32(data + 4);
            depth = data[8];
            ctype = data[9];
            interlace = data[12];
        } else if (!std::memcmp(type, "PLTE", 4)) {
            plte.assign(data, data + len);
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (I.rows <= 0 || I.cols <= 0) fail("missing header");
    if (depth != 8 || interlace != 0) fail("only 8 bits per channel, non-interlaced PNG images are read here");
    int ch = 0;
    switch (ctype) {
        case 0: ch = 1; break;
        case 2: ch = 3; break;
        case 3: ch = 1; break;
        case 4: ch = 2; break;
        case 6: ch = 4; break;
        default: fail("unknown colour type");
    }
    const size_t stride = (size_t)I.cols * ch;
    std::vector<uint8_t> raw((stride + 1) * (size_t)I.rows);
    uLongf out_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size()) != Z_OK || out_len != raw.size()) fail("corrupt image data");
    std::vector<uint8_t> cur(stride), prev(stride, 0);
    I.px.resize((size_t)I.rows * I.cols);
    for (int r = 0; r < I.rows; ++r) {
        const uint8_t *src = &raw[(stride + 1) * (size_t)r];
        const int ft = src[0];
        ++src;
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= (size_t)ch ? cur[i - ch] : 0, b = prev[i], c = i >= (size_t)ch ? prev[i - ch] : 0;
            int pred = 0;
            switch (ft) {
                case 0: pred = 0; break;
                case 1: pred = a; break;
                case 2: pred = b; break;
                case 3: pred = (a + b) >> 1; break;
                case 4: {
                    const int p = a + b - c, pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
                    pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    break;
                }
                default: fail("unknown row filter");
            }
            cur[i] = (uint8_t)(src[i] + pred);
        }
        uint8_t *dst = &I.px[(size_t)r * I.cols];
        for (int x = 0; x < I.cols; ++x) {
            const uint8_t *p = &cur[(size_t)x * ch];
            int R, G, B;
            if (ctype == 0 || ctype == 4) {
                dst[x] = p[0];
                continue;
            }
            if (ctype == 3) {
                if ((size_t)p[0] * 3 + 2 >= plte.size()) fail("palette index out of range");
                R = plte[p[0] * 3];
                G = plte[p[0] * 3 + 1];
                B = plte[p[0] * 3 + 2];
            } else {
                R = p[0];
                G = p[1];
                B = p[2];
            }
            // libpng's png_set_rgb_to_gray(1, 0.299, 0.587) as OpenCV's PNG reader requests it: 15-bit fixed point (9797 / 19234 and
            // the remainder 3737 for blue), truncated -- verified against cv2.imread(IMREAD_GRAYSCALE) of OpenCV 4.13; equal
            // channels pass through unchanged
            dst[x] = (R == G && G == B) ? (uint8_t)R : (uint8_t)((R * 9797 + G * 19234 + B * 3737) >> 15);
        }
        std::swap(cur, prev);
    }
    return I;
}

}  // namespace lmmedia
