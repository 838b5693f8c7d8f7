// png_io.cpp — input staging: PNG file image -> pixels in the caller's (pinned) buffer, the
// way cv::imread(path, CV_LOAD_IMAGE_UNCHANGED) hands KITTI frames to Tracking::Track (main.cpp:160-162): 8-bit gray
// stays gray, 8-bit RGB(A) comes back as interleaved BGR(A) (svo_extract_bgr / svo_frame_in.channels = 3 take it from
// there and convert on the device), 16-bit gray (the depth / disparity PNGs, main.cpp:55) comes back as native u16.
// Host code: inflate is bit-serial per stream and a batch holds only tens of streams, so the decode stays on host cores
// (zlib) and writes straight into pinned memory; callers run one decode per core beside the GPU lanes.  Plain C ABI, no
// context needed.  Not built: interlaced and palette images (cv::imwrite and the KITTI tools never produce them).
#include <stdint.h>
#include <string.h>
#include <zlib.h>
#include <vector>
#include "../../include/svo_b200.h"

namespace {

inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

struct PngHeader { int w, h, depth, color, channels, interlace; };

// walks the chunks; returns 0 and fills hdr (and, if idat != NULL, the concatenated IDAT payload) or a negative status
int parse(const uint8_t *f, size_t n, PngHeader *hdr, std::vector<uint8_t> *idat)
{
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (!f || n < 8 + 25 || memcmp(f, sig, 8) != 0) return SVO_E_INVALID;
    size_t pos = 8;
    bool have_hdr = false;
    while (pos + 12 <= n) {
        const uint32_t len = be32(f + pos);
        const uint8_t *type = f + pos + 4, *data = f + pos + 8;
        if ((size_t)len > n - pos - 12) return SVO_E_INVALID;
        if (!memcmp(type, "IHDR", 4)) {
            if (len != 13) return SVO_E_INVALID;
            hdr->w = (int)be32(data); hdr->h = (int)be32(data + 4);
            hdr->depth = data[8]; hdr->color = data[9]; hdr->interlace = data[12];
            if (data[10] != 0 || data[11] != 0) return SVO_E_INVALID;           // compression / filter method
            switch (hdr->color) {
            case 0: hdr->channels = 1; break;
            case 2: hdr->channels = 3; break;
            case 4: hdr->channels = 2; break;
            case 6: hdr->channels = 4; break;
            default: return SVO_E_INVALID;                                      // palette images are not supported
            }
            if (hdr->w <= 0 || hdr->h <= 0 || hdr->w > (1 << 15) || hdr->h > (1 << 15)) return SVO_E_INVALID;
            if (!(hdr->depth == 8 || (hdr->depth == 16 && hdr->color == 0)) || hdr->interlace != 0) return SVO_E_INVALID;
            have_hdr = true;
            if (!idat) return SVO_OK;
        } else if (!memcmp(type, "IDAT", 4)) {
            if (!have_hdr) return SVO_E_INVALID;
            if (idat) idat->insert(idat->end(), data, data + len);
        } else if (!memcmp(type, "IEND", 4)) break;
        pos += 12 + (size_t)len;
    }
    return have_hdr ? SVO_OK : SVO_E_INVALID;
}

inline int paeth(int a, int b, int c)
{
    const int p = a + b - c, pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

}  // namespace

extern "C" {

int svo_png_info(const uint8_t *file, size_t n, int *w, int *h, int *channels, int *bit_depth)
{
    PngHeader hd;
    const int rc = parse(file, n, &hd, nullptr);
    if (rc != SVO_OK) return rc;
    if (w) *w = hd.w;
    if (h) *h = hd.h;
    if (channels) *channels = hd.channels;
    if (bit_depth) *bit_depth = hd.depth;
    return SVO_OK;
}

int svo_png_decode(const uint8_t *file, size_t n, uint8_t *dst, size_t dst_stride, size_t dst_cap)
{
    PngHeader hd;
    std::vector<uint8_t> idat;
    idat.reserve(n);
    int rc = parse(file, n, &hd, &idat);
    if (rc != SVO_OK) return rc;
    const size_t bpp = (size_t)hd.channels * (hd.depth / 8), row = (size_t)hd.w * bpp;
    if (!dst || dst_stride < row || (size_t)(hd.h - 1) * dst_stride + row > dst_cap) return SVO_E_CAPACITY;
    // inflate the whole image (filter byte + row bytes per scanline), then undo the filters row by row into dst
    std::vector<uint8_t> raw((row + 1) * (size_t)hd.h);
    uLongf out_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size()) != Z_OK || out_len != raw.size()) return SVO_E_INVALID;
    const uint8_t *prev = nullptr;
    for (int y = 0; y < hd.h; ++y) {
        const uint8_t *src = raw.data() + (size_t)y * (row + 1);
        const int ft = src[0];
        ++src;
        uint8_t *o = dst + (size_t)y * dst_stride;
        switch (ft) {
        case 0: memcpy(o, src, row); break;
        case 1:
            for (size_t i = 0; i < row; ++i) o[i] = (uint8_t)(src[i] + (i >= bpp ? o[i - bpp] : 0));
            break;
        case 2:
            for (size_t i = 0; i < row; ++i) o[i] = (uint8_t)(src[i] + (prev ? prev[i] : 0));
            break;
        case 3:
            for (size_t i = 0; i < row; ++i) o[i] = (uint8_t)(src[i] + (((i >= bpp ? o[i - bpp] : 0) + (prev ? prev[i] : 0)) >> 1));
            break;
        case 4:
            for (size_t i = 0; i < row; ++i)
                o[i] = (uint8_t)(src[i] + paeth(i >= bpp ? o[i - bpp] : 0, prev ? prev[i] : 0, (prev && i >= bpp) ? prev[i - bpp] : 0));
            break;
        default: return SVO_E_INVALID;
        }
        prev = o;
    }
    // cv::imread's memory order: BGR(A) for colour, native-endian for 16 bit (done after the filters: they need PNG's order)
    for (int y = 0; y < hd.h; ++y) {
        uint8_t *o = dst + (size_t)y * dst_stride;
        if (hd.depth == 16) {
            for (size_t i = 0; i + 1 < row; i += 2) { const uint8_t t = o[i]; o[i] = o[i + 1]; o[i + 1] = t; }
        } else if (hd.channels >= 3) {
            for (size_t i = 0; i + 2 < row; i += bpp) { const uint8_t t = o[i]; o[i] = o[i + 2]; o[i + 2] = t; }
        }
    }
    return SVO_OK;
}

}  // extern "C"
