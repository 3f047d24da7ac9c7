// ygl_imageio.cpp — the image files a scene names, decoded to the texels the reference's decoders give (stb_image for
// PNG / JPEG / Radiance HDR, tinyexr for OpenEXR): scene ingestion, SURVEY.md §8f rank 3. Host C++ only.
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ygl_hostio.h"

namespace ygl_io {

namespace {

// ---------------------------------------------------------------------------------------------------------------
// textures
// ---------------------------------------------------------------------------------------------------------------
// Radiance .hdr -> float rgba, the conversion rule of stb_image's stbi__hdr_load / stbi__hdr_convert (what the
// reference calls): value = mantissa * 2^(exponent - 136), alpha 1, a zero exponent gives black.
bool load_hdr(const std::vector<uint8_t>& data, HostTexture& tex) {
  const uint8_t* p   = data.data();
  const uint8_t* end = p + data.size();
  auto           token = [&](std::string& line) {
    line.clear();
    while (p < end && *p != '\n') line += (char)*p++;
    if (p < end) p++;
  };
  std::string line;
  token(line);
  if (line != "#?RADIANCE" && line != "#?RGBE") return false;
  bool valid = false;
  while (true) {
    token(line);
    if (line.empty()) break;
    if (line == "FORMAT=32-bit_rle_rgbe") valid = true;
    if (p >= end) return false;
  }
  if (!valid) return false;
  token(line);
  int w = 0, h = 0;
  if (sscanf(line.c_str(), "-Y %d +X %d", &h, &w) != 2 || w <= 0 || h <= 0) return false;
  if (w > (1 << 24) || h > (1 << 24) || (uint64_t)w * h > (1u << 27)) return false;  // (stb: "too large")
  tex.width = w, tex.height = h;
  tex.pixelsf.assign((size_t)w * h * 4, 0.0f);
  static const struct Scales {  // 2^(e - 136) for every exponent byte: what ldexp returns, computed once
    float of[256];
    Scales() {
      for (int e = 0; e < 256; e++) of[e] = (float)ldexp(1.0f, e - (int)(128 + 8));
    }
  } scales;
  auto convert = [](float* out, const uint8_t* rgbe) {
    if (rgbe[3] != 0) {
      float f1 = scales.of[rgbe[3]];
      out[0] = rgbe[0] * f1, out[1] = rgbe[1] * f1, out[2] = rgbe[2] * f1;
    } else {
      out[0] = out[1] = out[2] = 0;
    }
    out[3] = 1;
  };
  if (w < 8 || w >= 32768) {  // flat data
    if (end - p < (ptrdiff_t)w * h * 4) return false;
    for (size_t i = 0; i < (size_t)w * h; i++) convert(&tex.pixelsf[4 * i], p + 4 * i);
    return true;
  }
  std::vector<uint8_t> scanline((size_t)w * 4);
  for (int j = 0; j < h; j++) {
    if (end - p < 4) return false;
    int c1 = p[0], c2 = p[1], len = p[2];
    if (c1 != 2 || c2 != 2 || (len & 0x80)) {
      // not run-length encoded: the rest of the file is flat (stb switches mode for the whole image at row 0)
      if (j != 0) return false;
      if (end - p < (ptrdiff_t)w * h * 4) return false;
      for (size_t i = 0; i < (size_t)w * h; i++) convert(&tex.pixelsf[4 * i], p + 4 * i);
      return true;
    }
    len = (len << 8) | p[3];
    if (len != w) return false;
    p += 4;
    for (int k = 0; k < 4; k++) {
      int i = 0;
      while (i < w) {
        if (p >= end) return false;
        int count = *p++;
        if (count > 128) {
          count -= 128;
          if (p >= end || i + count > w) return false;
          const uint8_t value = *p++;
          for (int z = 0; z < count; z++) scanline[(size_t)(i++) * 4 + k] = value;
        } else {
          if (count == 0 || end - p < count || i + count > w) return false;
          for (int z = 0; z < count; z++) scanline[(size_t)(i++) * 4 + k] = *p++;
        }
      }
    }
    for (int i = 0; i < w; i++) convert(&tex.pixelsf[((size_t)j * w + i) * 4], &scanline[(size_t)i * 4]);
  }
  return true;
}

// PNG -> byte rgba as stbi_load(..., 4) returns it (stb_image.h, stbi__parse_png_file / stbi__create_png_image): 1 / 2 /
// 4 / 8 / 16-bit samples (sub-byte grey scaled to 0..255, 16 -> high byte), grey / grey+alpha / rgb / rgba / palette,
// tRNS transparency (a colour key, or per-entry alpha of a palette whose missing entries read as transparent black),
// Adam7 interlacing, and stb's acceptance rules for chunk sizes.
struct PngLayout {
  int w = 0, h = 0, depth = 0, ctype = 0, channels = 0;
};
// one (sub)image: `raw` holds h filtered scanlines of w pixels; the result has one byte per sample for depths <= 8
// (sub-byte samples unpacked, not yet scaled) and two for depth 16
bool png_unfilter(const uint8_t* raw, size_t raw_size, int w, int h, const PngLayout& png, std::vector<uint8_t>& out) {
  const size_t bpp    = std::max<size_t>(1, (size_t)png.channels * png.depth / 8);  // filter distance in bytes
  const size_t stride = ((size_t)png.channels * w * png.depth + 7) / 8;
  if (raw_size < (stride + 1) * h) return false;
  std::vector<uint8_t> img(stride * h);
  for (int j = 0; j < h; j++) {
    const uint8_t  filter = raw[(stride + 1) * j];
    const uint8_t* src    = &raw[(stride + 1) * j + 1];
    uint8_t*       dst    = &img[stride * j];
    const uint8_t* up     = j ? dst - stride : nullptr;
    if (filter > 4) return false;
    // PNG specification, section 9 (a: left, b: above, c: above left; bytes before the row or the image read as 0)
    switch (filter) {
      case 0: memcpy(dst, src, stride); break;
      case 1:
        for (size_t i = 0; i < stride; i++) dst[i] = (uint8_t)(src[i] + (i >= bpp ? dst[i - bpp] : 0));
        break;
      case 2:
        if (!up) memcpy(dst, src, stride);
        else
          for (size_t i = 0; i < stride; i++) dst[i] = (uint8_t)(src[i] + up[i]);
        break;
      case 3:
        for (size_t i = 0; i < stride; i++) dst[i] = (uint8_t)(src[i] + (((i >= bpp ? dst[i - bpp] : 0) + (up ? up[i] : 0)) >> 1));
        break;
      default:
        for (size_t i = 0; i < stride; i++) {
          const int a = i >= bpp ? dst[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0;
          const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
          dst[i] = (uint8_t)(src[i] + ((pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c)));
        }
        break;
    }
  }
  if (png.depth >= 8) return out = std::move(img), true;
  out.resize((size_t)w * h);  // one channel: unpack the most significant bits first
  const int mask = (1 << png.depth) - 1;
  for (int j = 0; j < h; j++)
    for (int i = 0; i < w; i++) {
      const size_t bit = (size_t)i * png.depth;
      out[(size_t)j * w + i] = (uint8_t)((img[stride * j + bit / 8] >> (8 - png.depth - bit % 8)) & mask);
    }
  return true;
}
bool load_png(const std::vector<uint8_t>& data, HostTexture& tex) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (data.size() < 8 || memcmp(data.data(), sig, 8)) return false;
  auto   be32 = [&](size_t o) { return (uint32_t)data[o] << 24 | (uint32_t)data[o + 1] << 16 | (uint32_t)data[o + 2] << 8 | data[o + 3]; };
  size_t pos = 8;
  PngLayout png;
  int       interlace = 0;
  bool      first = true, have_trns = false, ended = false;
  uint8_t   palette[256][4] = {};  // entries a PLTE chunk does not fill read as 0, 0, 0, 0 (stb: uninitialised memory)
  int       palette_len = 0;
  uint8_t   key8[3]  = {};         // colour key of an 8-bit image (already scaled like the samples)
  uint16_t  key16[3] = {};
  std::vector<uint8_t> idat;
  while (!ended) {
    if (pos + 8 > data.size()) return false;
    const uint32_t len = be32(pos);
    const char*    id  = (const char*)&data[pos + 4];
    if (pos + 12 + (size_t)len > data.size()) {
      if (memcmp(id, "IEND", 4) || pos + 8 + (size_t)len > data.size()) return false;  // stb never reads IEND's checksum
    }
    const uint8_t* body = &data[pos + 8];
    if (first && memcmp(id, "IHDR", 4)) return false;
    if (!memcmp(id, "IHDR", 4)) {
      if (!first || len != 13) return false;
      png.w = (int)be32(pos + 8), png.h = (int)be32(pos + 12);
      png.depth = body[8], png.ctype = body[9], interlace = body[12];
      if (be32(pos + 8) > (1u << 24) || be32(pos + 12) > (1u << 24) || png.w == 0 || png.h == 0) return false;
      if ((uint64_t)png.w * png.h > (1u << 28)) return false;  // (stb refuses what does not fit 2^30 bytes)
      if (png.depth != 1 && png.depth != 2 && png.depth != 4 && png.depth != 8 && png.depth != 16) return false;
      if (png.ctype > 6 || (png.ctype == 3 && png.depth == 16) || (png.ctype != 3 && (png.ctype & 1))) return false;
      if (body[10] != 0 || body[11] != 0 || interlace > 1) return false;
      png.channels = png.ctype == 3 ? 1 : ((png.ctype & 2) ? 3 : 1) + ((png.ctype & 4) ? 1 : 0);
    } else if (!memcmp(id, "PLTE", 4)) {
      if (len > 256 * 3 || len % 3 != 0) return false;
      palette_len = (int)len / 3;
      for (int k = 0; k < palette_len; k++)
        palette[k][0] = body[3 * k], palette[k][1] = body[3 * k + 1], palette[k][2] = body[3 * k + 2], palette[k][3] = 255;
    } else if (!memcmp(id, "tRNS", 4)) {
      if (!idat.empty()) return false;
      if (png.ctype == 3) {
        if (palette_len == 0 || (int)len > palette_len) return false;
        for (uint32_t k = 0; k < len; k++) palette[k][3] = body[k];
      } else {
        if (!(png.channels & 1) || len != (uint32_t)png.channels * 2) return false;
        static const uint8_t scale[9] = {0, 0xff, 0x55, 0, 0x11, 0, 0, 0, 0x01};
        for (int k = 0; k < png.channels; k++) {
          key16[k] = (uint16_t)(body[2 * k] << 8 | body[2 * k + 1]);
          if (png.depth < 16) key8[k] = (uint8_t)((key16[k] & 255) * scale[png.depth]);
        }
        have_trns = true;
      }
    } else if (!memcmp(id, "IDAT", 4)) {
      if (png.ctype == 3 && palette_len == 0) return false;
      idat.insert(idat.end(), body, body + len);
    } else if (!memcmp(id, "IEND", 4)) {
      ended = true;
    } else if (!(id[0] & 32)) {
      return false;  // an unknown critical chunk
    }
    first = false;
    pos += 12 + (size_t)len;
  }
  if (idat.empty()) return false;
  const int w = png.w, h = png.h;
  // inflate everything; interlaced images are seven sub-images one after the other
  const size_t bytes_per_sample = png.depth == 16 ? 2 : 1;
  const size_t bpp_out          = (size_t)png.channels * bytes_per_sample;
  size_t       expected         = 0;
  static const int xorig[7] = {0, 4, 0, 2, 0, 1, 0}, yorig[7] = {0, 0, 4, 0, 2, 0, 1};
  static const int xspc[7] = {8, 8, 4, 4, 2, 2, 1}, yspc[7] = {8, 8, 8, 4, 4, 2, 2};
  auto raw_size = [&](int pw, int ph) { return (((size_t)png.channels * pw * png.depth + 7) / 8 + 1) * ph; };
  if (!interlace) {
    expected = raw_size(w, h);
  } else {
    for (int p = 0; p < 7; p++) {
      const int pw = (w - xorig[p] + xspc[p] - 1) / xspc[p], ph = (h - yorig[p] + yspc[p] - 1) / yspc[p];
      if (pw > 0 && ph > 0) expected += raw_size(pw, ph);
    }
  }
  std::vector<uint8_t> raw(expected);
  uLongf rawlen = (uLongf)raw.size();
  if (uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size()) != Z_OK || rawlen != raw.size()) return false;
  std::vector<uint8_t> img((size_t)w * h * bpp_out);
  if (!interlace) {
    if (!png_unfilter(raw.data(), raw.size(), w, h, png, img)) return false;
  } else {
    size_t offset = 0;
    for (int p = 0; p < 7; p++) {
      const int pw = (w - xorig[p] + xspc[p] - 1) / xspc[p], ph = (h - yorig[p] + yspc[p] - 1) / yspc[p];
      if (pw <= 0 || ph <= 0) continue;
      std::vector<uint8_t> part;
      if (!png_unfilter(raw.data() + offset, raw.size() - offset, pw, ph, png, part)) return false;
      for (int j = 0; j < ph; j++)
        for (int i = 0; i < pw; i++)
          memcpy(&img[((size_t)(j * yspc[p] + yorig[p]) * w + (size_t)i * xspc[p] + xorig[p]) * bpp_out],
              &part[((size_t)j * pw + i) * bpp_out], bpp_out);
      offset += raw_size(pw, ph);
    }
  }
  if (png.depth < 8 && png.ctype == 0) {  // sub-byte grey spreads over 0..255
    static const uint8_t scale[9] = {0, 0xff, 0x55, 0, 0x11, 0, 0, 0, 0x01};
    for (auto& v : img) v = (uint8_t)(v * scale[png.depth]);
  }
  tex.width = w, tex.height = h;
  tex.pixelsb.resize((size_t)w * h * 4);
  const bool   wide = png.depth == 16;
  const size_t step = bytes_per_sample;  // 16-bit samples are big-endian: the first byte is the high byte stb keeps
  for (size_t i = 0; i < (size_t)w * h; i++) {
    const uint8_t* s = &img[i * bpp_out];
    uint8_t*       d = &tex.pixelsb[4 * i];
    auto sample16 = [&](int k) { return (uint16_t)(s[2 * k] << 8 | s[2 * k + 1]); };
    switch (png.ctype) {
      case 0: {
        d[0] = d[1] = d[2] = s[0], d[3] = 255;
        if (have_trns && (wide ? sample16(0) == key16[0] : s[0] == key8[0])) d[3] = 0;
      } break;
      case 2: {
        d[0] = s[0], d[1] = s[step], d[2] = s[2 * step], d[3] = 255;
        if (have_trns) {
          bool key = true;
          for (int k = 0; k < 3; k++) key = key && (wide ? sample16(k) == key16[k] : s[k] == key8[k]);
          if (key) d[3] = 0;
        }
      } break;
      case 3: memcpy(d, palette[s[0]], 4); break;
      case 4: d[0] = d[1] = d[2] = s[0], d[3] = s[step]; break;
      case 6: d[0] = s[0], d[1] = s[step], d[2] = s[2 * step], d[3] = s[3 * step]; break;
    }
  }
  return true;
}

// ---------------------------------------------------------------------------------------------------------------
// JPEG -> byte rgba as stbi_load(..., 4) returns it. The format leaves the inverse DCT, the chroma upsampling and the
// colour conversion to the decoder, so "the same bytes as the reference" means stb_image's choices (stb_image.h,
// stbi__idct_block / stbi__resample_row_* / stbi__YCbCr_to_RGB_row; its SSE2 variants are written to match them
// exactly): the 12-bit fixed-point "islow" IDCT with two extra bits between the passes, triangle-filter upsampling for
// 2x factors (3:1 weights, rounded) and replication for the others, and the reduced-precision YCbCr -> RGB.
// Baseline and progressive Huffman streams, 8-bit, 1 / 3 / 4 components, restart intervals, JFIF / Adobe markers.
// ---------------------------------------------------------------------------------------------------------------
class JpegDecoder {
 public:
  bool decode(const std::vector<uint8_t>& file, HostTexture& tex) {
    at_ = file.data(), end_ = file.data() + file.size();
    if (next_marker() != 0xD8) return false;  // SOI
    int m = next_marker();
    while (m != 0xC0 && m != 0xC1 && m != 0xC2) {  // tables and application segments up to the frame header
      if (!segment(m)) return false;
      m = next_marker();
      while (m == kNoMarker) {
        if (at_ >= end_) return false;
        m = next_marker();
      }
    }
    progressive_ = m == 0xC2;
    if (!frame_header()) return false;
    m = next_marker();
    while (m != 0xD9) {  // until EOI
      if (m == 0xDA) {
        if (!scan_header() || !scan()) return false;
        if (pending_ == kNoMarker) {  // zero padding after the entropy-coded data
          while (at_ < end_) {
            if (byte() == 255) {
              pending_ = byte();
              break;
            }
          }
        }
      } else if (m == 0xDC) {  // DNL
        const int len = u16(), lines = u16();
        if (len != 4 || lines != height_) return false;
      } else if (!segment(m)) {
        return false;
      }
      m = next_marker();
    }
    if (progressive_) finish_progressive();
    return assemble(tex);
  }

 private:
  static constexpr int kNoMarker = 0xff;
  struct Huffman {  // a table no DHT segment has filled decodes nothing: every lookup runs off its end
    uint8_t  size[257] = {}, values[256] = {};
    uint16_t code[256] = {};
    unsigned maxcode[18] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0xffffffffu};
    int      delta[17] = {};
    bool build(const int* count) {  // canonical codes in symbol order
      int k = 0;
      for (int i = 0; i < 16; i++)
        for (int j = 0; j < count[i]; j++) {
          if (k >= 256) return false;
          size[k++] = (uint8_t)(i + 1);
        }
      size[k] = 0;
      unsigned next = 0;
      k = 0;
      for (int len = 1; len <= 16; len++) {
        delta[len] = k - (int)next;
        if (size[k] == len) {
          while (size[k] == len) code[k++] = (uint16_t)(next++);
          if (next - 1 >= (1u << len)) return false;
        }
        maxcode[len] = next << (16 - len);
        next <<= 1;
      }
      maxcode[17] = 0xffffffff;
      return true;
    }
  };
  struct Component {
    int id = 0, h = 0, v = 0, tq = 0, hd = 0, ha = 0, dc_pred = 0;
    int x = 0, y = 0, w2 = 0, h2 = 0, coeff_w = 0;
    std::vector<uint8_t> data;    // w2 x h2 samples
    std::vector<int16_t> coeff;   // progressive: 64 per block
  };

  const uint8_t *at_ = nullptr, *end_ = nullptr;
  int  pending_ = kNoMarker;  // a marker met inside entropy-coded data
  bool progressive_ = false, jfif_ = false, no_more_ = false;
  int  adobe_transform_ = -1, rgb_ids_ = 0;
  int  width_ = 0, height_ = 0, ncomp_ = 0, h_max_ = 1, v_max_ = 1, mcu_x_ = 0, mcu_y_ = 0;
  int  restart_interval_ = 0, todo_ = 0, eob_run_ = 0;
  int  scan_n_ = 0, order_[4] = {}, spec_start_ = 0, spec_end_ = 0, succ_high_ = 0, succ_low_ = 0;
  uint32_t  bits_ = 0;
  int       nbits_ = 0;
  uint16_t  dequant_[4][64] = {};
  Huffman   dc_[4], ac_[4];
  Component comp_[4];

  int byte() { return at_ < end_ ? *at_++ : 0; }
  int u16() {
    const int hi = byte();
    return hi << 8 | byte();
  }
  void skip(int n) { at_ = n < 0 || end_ - at_ < n ? end_ : at_ + n; }
  int  next_marker() {
    if (pending_ != kNoMarker) {
      const int m = pending_;
      pending_    = kNoMarker;
      return m;
    }
    int x = byte();
    if (x != 0xff) return kNoMarker;
    while (x == 0xff) x = byte();
    return x;
  }
  static int zigzag(int k) {
    static const uint8_t order[64 + 15] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20,
        13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60,
        61, 54, 47, 55, 62, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};
    return order[k];
  }

  bool segment(int m) {
    switch (m) {
      case kNoMarker: return false;
      case 0xDD:
        if (u16() != 4) return false;
        restart_interval_ = u16();
        return true;
      case 0xDB: {
        int len = u16() - 2;
        while (len > 0) {
          const int q = byte(), wide = q >> 4, t = q & 15;
          if ((wide != 0 && wide != 1) || t > 3) return false;
          for (int i = 0; i < 64; i++) dequant_[t][zigzag(i)] = (uint16_t)(wide ? u16() : byte());
          len -= wide ? 129 : 65;
        }
        return len == 0;
      }
      case 0xC4: {
        int len = u16() - 2;
        while (len > 0) {
          const int q = byte(), cls = q >> 4, slot = q & 15;
          if (cls > 1 || slot > 3) return false;
          int sizes[16], n = 0;
          for (int& sz : sizes) n += sz = byte();
          len -= 17;
          Huffman& table = cls == 0 ? dc_[slot] : ac_[slot];
          if (!table.build(sizes)) return false;
          for (int i = 0; i < n; i++) table.values[i] = (uint8_t)byte();
          len -= n;
        }
        return len == 0;
      }
      default: break;
    }
    if ((m >= 0xE0 && m <= 0xEF) || m == 0xFE) {
      int len = u16();
      if (len < 2) return false;
      len -= 2;
      if (m == 0xE0 && len >= 5) {
        bool ok = true;
        for (char c : {'J', 'F', 'I', 'F', '\0'}) ok &= byte() == c;
        len -= 5;
        if (ok) jfif_ = true;
      } else if (m == 0xEE && len >= 12) {
        bool ok = true;
        for (char c : {'A', 'd', 'o', 'b', 'e', '\0'}) ok &= byte() == c;
        len -= 6;
        if (ok) {
          byte(), u16(), u16();
          adobe_transform_ = byte();
          len -= 6;
        }
      }
      skip(len);
      return true;
    }
    return false;
  }

  bool frame_header() {
    const int len = u16();
    if (len < 11 || byte() != 8) return false;
    height_ = u16(), width_ = u16();
    if (height_ == 0 || width_ == 0 || height_ > (1 << 24) || width_ > (1 << 24)) return false;
    ncomp_ = byte();
    if (ncomp_ != 1 && ncomp_ != 3 && ncomp_ != 4) return false;
    if (len != 8 + 3 * ncomp_) return false;
    for (int i = 0; i < ncomp_; i++) {
      Component& c = comp_[i];
      c.id         = byte();
      if (ncomp_ == 3 && c.id == "RGB"[i]) rgb_ids_++;
      const int q = byte();
      c.h = q >> 4, c.v = q & 15, c.tq = byte();
      if (!c.h || c.h > 4 || !c.v || c.v > 4 || c.tq > 3) return false;
    }
    if ((uint64_t)width_ * height_ * ncomp_ > (1u << 30)) return false;
    for (int i = 0; i < ncomp_; i++) h_max_ = std::max(h_max_, comp_[i].h), v_max_ = std::max(v_max_, comp_[i].v);
    for (int i = 0; i < ncomp_; i++)
      if (h_max_ % comp_[i].h != 0 || v_max_ % comp_[i].v != 0) return false;
    mcu_x_ = (width_ + h_max_ * 8 - 1) / (h_max_ * 8), mcu_y_ = (height_ + v_max_ * 8 - 1) / (v_max_ * 8);
    for (int i = 0; i < ncomp_; i++) {
      Component& c = comp_[i];
      c.x = (width_ * c.h + h_max_ - 1) / h_max_, c.y = (height_ * c.v + v_max_ - 1) / v_max_;
      c.w2 = mcu_x_ * c.h * 8, c.h2 = mcu_y_ * c.v * 8;  // whole MCUs: the surplus is cut at assembly
      c.data.assign((size_t)c.w2 * c.h2, 0);
      if (progressive_) c.coeff_w = c.w2 / 8, c.coeff.assign((size_t)c.w2 * c.h2, 0);
    }
    return true;
  }

  bool scan_header() {
    const int len = u16();
    scan_n_       = byte();
    if (scan_n_ < 1 || scan_n_ > 4 || scan_n_ > ncomp_ || len != 6 + 2 * scan_n_) return false;
    for (int i = 0; i < scan_n_; i++) {
      const int id = byte(), q = byte();
      int       which = 0;
      while (which < ncomp_ && comp_[which].id != id) which++;
      if (which == ncomp_) return false;
      comp_[which].hd = q >> 4, comp_[which].ha = q & 15;
      if (comp_[which].hd > 3 || comp_[which].ha > 3) return false;
      order_[i] = which;
    }
    spec_start_ = byte(), spec_end_ = byte();
    const int approx = byte();
    succ_high_ = approx >> 4, succ_low_ = approx & 15;
    if (progressive_) {
      if (spec_start_ > 63 || spec_end_ > 63 || spec_start_ > spec_end_ || succ_high_ > 13 || succ_low_ > 13) return false;
    } else {
      if (spec_start_ != 0 || succ_high_ != 0 || succ_low_ != 0) return false;
      spec_end_ = 63;
    }
    return true;
  }

  // ---- entropy-coded bits ----
  void fill() {
    do {
      const unsigned b = no_more_ ? 0 : (unsigned)byte();
      if (b == 0xff) {
        int c = byte();
        while (c == 0xff) c = byte();
        if (c != 0) {
          pending_ = c, no_more_ = true;
          return;
        }
      }
      bits_ |= b << (24 - nbits_);
      nbits_ += 8;
    } while (nbits_ <= 24);
  }
  int symbol(const Huffman& h) {
    if (nbits_ < 16) fill();
    const unsigned top = bits_ >> 16;
    int            len = 1;
    while (top >= h.maxcode[len]) len++;
    if (len == 17) {
      nbits_ -= 16;
      return -1;
    }
    if (len > nbits_) return -1;
    const int index = (int)((bits_ >> (32 - len)) & ((1u << len) - 1)) + h.delta[len];
    if (index < 0 || index > 255) return -1;
    nbits_ -= len, bits_ <<= len;
    return h.values[index];
  }
  int take(int n) {  // n unsigned bits
    if (n == 0) return 0;
    if (nbits_ < n) fill();
    const unsigned k = bits_ >> (32 - n);
    bits_ <<= n, nbits_ -= n;
    return (int)k;
  }
  int take_signed(int n) {  // "receive and extend"
    if (n == 0) return 0;
    if (nbits_ < n) fill();
    const bool     positive = bits_ >> 31;
    const unsigned k        = bits_ >> (32 - n);
    bits_ <<= n, nbits_ -= n;
    return (int)k + (positive ? 0 : (int)((~0u << n) + 1));
  }
  void reset() {
    nbits_ = 0, bits_ = 0, no_more_ = false, pending_ = kNoMarker, eob_run_ = 0;
    for (auto& c : comp_) c.dc_pred = 0;
    todo_ = restart_interval_ ? restart_interval_ : 0x7fffffff;
  }
  bool restart_due() {  // after an MCU: true = the scan ends here (no restart marker where one is due)
    if (--todo_ > 0) return false;
    if (nbits_ < 24) fill();
    if (!(pending_ >= 0xd0 && pending_ <= 0xd7)) return true;
    reset();
    return false;
  }

  bool block_baseline(int16_t* data, Component& c) {
    const int t = symbol(dc_[c.hd]);
    if (t < 0 || t > 15) return false;
    memset(data, 0, 64 * sizeof(int16_t));
    const uint16_t* dq = dequant_[c.tq];
    c.dc_pred += take_signed(t);
    data[0] = (int16_t)(c.dc_pred * dq[0]);
    for (int k = 1; k < 64;) {
      const int rs = symbol(ac_[c.ha]);
      if (rs < 0) return false;
      const int s = rs & 15, r = rs >> 4;
      if (s == 0) {
        if (rs != 0xf0) break;
        k += 16;
      } else {
        k += r;
        const int z = zigzag(k++);
        data[z]     = (int16_t)(take_signed(s) * dq[z]);
      }
    }
    return true;
  }
  bool block_dc_progressive(int16_t* data, Component& c) {
    if (spec_end_ != 0) return false;
    if (succ_high_ == 0) {
      memset(data, 0, 64 * sizeof(int16_t));
      const int t = symbol(dc_[c.hd]);
      if (t < 0 || t > 15) return false;
      c.dc_pred += take_signed(t);
      data[0] = (int16_t)(c.dc_pred * (1 << succ_low_));
    } else if (take(1)) {
      data[0] += (int16_t)(1 << succ_low_);
    }
    return true;
  }
  void refine(int16_t& v, int16_t bit) {
    if (take(1) && (v & bit) == 0) v = (int16_t)(v > 0 ? v + bit : v - bit);
  }
  bool block_ac_progressive(int16_t* data, Component& c) {
    if (spec_start_ == 0) return false;
    const Huffman& table = ac_[c.ha];
    if (succ_high_ == 0) {
      if (eob_run_) return --eob_run_, true;
      int k = spec_start_;
      do {
        const int rs = symbol(table);
        if (rs < 0) return false;
        const int s = rs & 15, r = rs >> 4;
        if (s == 0) {
          if (r < 15) {
            eob_run_ = (1 << r) + take(r) - 1;
            break;
          }
          k += 16;
        } else {
          k += r;
          data[zigzag(k++)] = (int16_t)(take_signed(s) * (1 << succ_low_));
        }
      } while (k <= spec_end_);
      return true;
    }
    const int16_t bit = (int16_t)(1 << succ_low_);
    if (eob_run_) {
      --eob_run_;
      for (int k = spec_start_; k <= spec_end_; k++)
        if (int16_t& v = data[zigzag(k)]; v != 0) refine(v, bit);
      return true;
    }
    int k = spec_start_;
    do {
      const int rs = symbol(table);
      if (rs < 0) return false;
      int s = rs & 15, r = rs >> 4;
      if (s == 0) {
        if (r < 15) {
          eob_run_ = (1 << r) - 1 + take(r);
          r        = 64;  // to the end of the band
        }
      } else {
        if (s != 1) return false;
        s = take(1) ? bit : -bit;
      }
      while (k <= spec_end_) {
        int16_t& v = data[zigzag(k++)];
        if (v != 0) {
          refine(v, bit);
        } else {
          if (r == 0) {
            v = (int16_t)s;
            break;
          }
          --r;
        }
      }
    } while (k <= spec_end_);
    return true;
  }

  // 8x8 inverse DCT on dequantised coefficients, writing clamped samples
  static void idct_1d(int s0, int s1, int s2, int s3, int s4, int s5, int s6, int s7, int (&x)[4], int (&t)[4]) {
    auto f = [](double v) { return (int)(v * 4096 + 0.5); };
    int p2 = s2, p3 = s6;
    int p1 = (p2 + p3) * f(0.5411961f);
    int t2 = p1 + p3 * f(-1.847759065f), t3 = p1 + p2 * f(0.765366865f);
    p2 = s0, p3 = s4;
    int t0 = (p2 + p3) * 4096, t1 = (p2 - p3) * 4096;
    x[0] = t0 + t3, x[3] = t0 - t3, x[1] = t1 + t2, x[2] = t1 - t2;
    t0 = s7, t1 = s5, t2 = s3, t3 = s1;
    p3 = t0 + t2;
    int p4 = t1 + t3;
    p1 = t0 + t3, p2 = t1 + t2;
    const int p5 = (p3 + p4) * f(1.175875602f);
    t0 = t0 * f(0.298631336f), t1 = t1 * f(2.053119869f), t2 = t2 * f(3.072711026f), t3 = t3 * f(1.501321110f);
    p1 = p5 + p1 * f(-0.899976223f), p2 = p5 + p2 * f(-2.562915447f);
    p3 = p3 * f(-1.961570560f), p4 = p4 * f(-0.390180644f);
    t[3] = t3 + p1 + p4, t[2] = t2 + p2 + p3, t[1] = t1 + p2 + p4, t[0] = t0 + p1 + p3;
  }
  static uint8_t clamp8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }
  static void    idct(uint8_t* out, int stride, const int16_t* d) {
    int val[64];
    for (int i = 0; i < 8; i++) {  // columns, keeping two extra bits
      const int16_t* c = d + i;
      int*           v = val + i;
      if (c[8] == 0 && c[16] == 0 && c[24] == 0 && c[32] == 0 && c[40] == 0 && c[48] == 0 && c[56] == 0) {
        const int dc = c[0] * 4;
        for (int r = 0; r < 8; r++) v[r * 8] = dc;
        continue;
      }
      int x[4], t[4];
      idct_1d(c[0], c[8], c[16], c[24], c[32], c[40], c[48], c[56], x, t);
      for (int& e : x) e += 512;
      v[0] = (x[0] + t[3]) >> 10, v[56] = (x[0] - t[3]) >> 10, v[8] = (x[1] + t[2]) >> 10, v[48] = (x[1] - t[2]) >> 10;
      v[16] = (x[2] + t[1]) >> 10, v[40] = (x[2] - t[1]) >> 10, v[24] = (x[3] + t[0]) >> 10, v[32] = (x[3] - t[0]) >> 10;
    }
    for (int i = 0; i < 8; i++, out += stride) {  // rows: remove 2^17, centre on 128
      const int* v = val + i * 8;
      int        x[4], t[4];
      idct_1d(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], x, t);
      for (int& e : x) e += 65536 + (128 << 17);
      out[0] = clamp8((x[0] + t[3]) >> 17), out[7] = clamp8((x[0] - t[3]) >> 17), out[1] = clamp8((x[1] + t[2]) >> 17);
      out[6] = clamp8((x[1] - t[2]) >> 17), out[2] = clamp8((x[2] + t[1]) >> 17), out[5] = clamp8((x[2] - t[1]) >> 17);
      out[3] = clamp8((x[3] + t[0]) >> 17), out[4] = clamp8((x[3] - t[0]) >> 17);
    }
  }

  bool scan() {
    reset();
    int16_t block[64];
    auto    one = [&](Component& c, int bx, int by) {  // the block at (bx, by) of component c
      if (!progressive_) {
        if (!block_baseline(block, c)) return false;
        idct(&c.data[(size_t)c.w2 * by * 8 + bx * 8], c.w2, block);
        return true;
      }
      int16_t* data = &c.coeff[64 * ((size_t)bx + (size_t)by * c.coeff_w)];
      return spec_start_ == 0 ? block_dc_progressive(data, c) : block_ac_progressive(data, c);
    };
    if (scan_n_ == 1) {  // one component: its own blocks in raster order
      Component& c = comp_[order_[0]];
      const int  w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
      for (int j = 0; j < h; j++)
        for (int i = 0; i < w; i++) {
          if (!one(c, i, j)) return false;
          if (restart_due()) return true;
        }
      return true;
    }
    for (int j = 0; j < mcu_y_; j++)  // interleaved MCUs
      for (int i = 0; i < mcu_x_; i++) {
        for (int k = 0; k < scan_n_; k++) {
          Component& c = comp_[order_[k]];
          for (int y = 0; y < c.v; y++)
            for (int x = 0; x < c.h; x++) {
              if (progressive_ && spec_start_ != 0) return false;  // AC scans carry one component
              if (!one(c, i * c.h + x, j * c.v + y)) return false;
            }
        }
        if (restart_due()) return true;
      }
    return true;
  }
  void finish_progressive() {
    for (int n = 0; n < ncomp_; n++) {
      Component& c = comp_[n];
      const int  w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
      for (int j = 0; j < h; j++)
        for (int i = 0; i < w; i++) {
          int16_t* data = &c.coeff[64 * ((size_t)i + (size_t)j * c.coeff_w)];
          for (int k = 0; k < 64; k++) data[k] = (int16_t)(data[k] * dequant_[c.tq][k]);
          idct(&c.data[(size_t)c.w2 * j * 8 + i * 8], c.w2, data);
        }
    }
  }

  // ---- full-resolution rows: 2x factors by the 3:1 triangle filter, others by replication ----
  static const uint8_t* upsample(uint8_t* out, const uint8_t* near, const uint8_t* far, int w, int hs, int vs) {
    if (hs == 1 && vs == 1) return near;
    if (hs == 1 && vs == 2) {
      for (int i = 0; i < w; i++) out[i] = (uint8_t)((3 * near[i] + far[i] + 2) >> 2);
    } else if (hs == 2 && vs == 1) {
      if (w == 1) return out[0] = out[1] = near[0], out;
      out[0] = near[0], out[1] = (uint8_t)((near[0] * 3 + near[1] + 2) >> 2);
      for (int i = 1; i < w - 1; i++) {
        const int n    = 3 * near[i] + 2;
        out[i * 2]     = (uint8_t)((n + near[i - 1]) >> 2);
        out[i * 2 + 1] = (uint8_t)((n + near[i + 1]) >> 2);
      }
      out[(w - 1) * 2] = (uint8_t)((near[w - 2] * 3 + near[w - 1] + 2) >> 2), out[(w - 1) * 2 + 1] = near[w - 1];
    } else if (hs == 2 && vs == 2) {
      if (w == 1) return out[0] = out[1] = (uint8_t)((3 * near[0] + far[0] + 2) >> 2), out;
      int t1 = 3 * near[0] + far[0];
      out[0] = (uint8_t)((t1 + 2) >> 2);
      for (int i = 1; i < w; i++) {
        const int t0   = t1;
        t1             = 3 * near[i] + far[i];
        out[i * 2 - 1] = (uint8_t)((3 * t0 + t1 + 8) >> 4);
        out[i * 2]     = (uint8_t)((3 * t1 + t0 + 8) >> 4);
      }
      out[w * 2 - 1] = (uint8_t)((t1 + 2) >> 2);
    } else {
      for (int i = 0; i < w; i++)
        for (int j = 0; j < hs; j++) out[i * hs + j] = near[i];
    }
    return out;
  }
  static void ycc_to_rgb(uint8_t* out, const uint8_t* y, const uint8_t* pcb, const uint8_t* pcr, int count) {
    auto fixed = [](float v) { return ((int)(v * 4096.0f + 0.5f)) << 8; };
    for (int i = 0; i < count; i++, out += 4) {
      const int luma = (y[i] << 20) + (1 << 19), cr = pcr[i] - 128, cb = pcb[i] - 128;
      int       r    = luma + cr * fixed(1.40200f);
      int       g    = luma + (cr * -fixed(0.71414f)) + (int)((unsigned)(cb * -fixed(0.34414f)) & 0xffff0000u);
      int       b    = luma + cb * fixed(1.77200f);
      out[0] = clamp8(r >> 20), out[1] = clamp8(g >> 20), out[2] = clamp8(b >> 20), out[3] = 255;
    }
  }
  static uint8_t scale8(uint8_t x, uint8_t y) {  // x * y / 255, rounded
    const unsigned t = x * y + 128;
    return (uint8_t)((t + (t >> 8)) >> 8);
  }
  bool assemble(HostTexture& tex) {
    const bool is_rgb = ncomp_ == 3 && (rgb_ids_ == 3 || (adobe_transform_ == 0 && !jfif_));
    struct Plane {
      int                  hs, vs, ystep, w_lores, ypos;
      const uint8_t *      line0, *line1;
      std::vector<uint8_t> row;
    } plane[4];
    for (int k = 0; k < ncomp_; k++) {
      Plane& p  = plane[k];
      p.hs = h_max_ / comp_[k].h, p.vs = v_max_ / comp_[k].v, p.ystep = p.vs >> 1;
      p.w_lores = (width_ + p.hs - 1) / p.hs, p.ypos = 0;
      p.line0 = p.line1 = comp_[k].data.data();
      p.row.resize((size_t)width_ + 3 + 8);
    }
    tex.width = width_, tex.height = height_;
    tex.pixelsb.resize((size_t)width_ * height_ * 4);
    for (int j = 0; j < height_; j++) {
      uint8_t*       out = &tex.pixelsb[(size_t)j * width_ * 4];
      const uint8_t* rows[4] = {};
      for (int k = 0; k < ncomp_; k++) {
        Plane&     p      = plane[k];
        const bool bottom = p.ystep >= (p.vs >> 1);
        rows[k]           = upsample(p.row.data(), bottom ? p.line1 : p.line0, bottom ? p.line0 : p.line1, p.w_lores, p.hs, p.vs);
        if (++p.ystep >= p.vs) {
          p.ystep = 0, p.line0 = p.line1;
          if (++p.ypos < comp_[k].y) p.line1 += comp_[k].w2;
        }
      }
      if (ncomp_ == 3 && is_rgb) {
        for (int i = 0; i < width_; i++) out[4 * i] = rows[0][i], out[4 * i + 1] = rows[1][i], out[4 * i + 2] = rows[2][i], out[4 * i + 3] = 255;
      } else if (ncomp_ == 3) {
        ycc_to_rgb(out, rows[0], rows[1], rows[2], width_);
      } else if (ncomp_ == 4 && adobe_transform_ == 0) {  // CMYK
        for (int i = 0; i < width_; i++) {
          const uint8_t m = rows[3][i];
          out[4 * i] = scale8(rows[0][i], m), out[4 * i + 1] = scale8(rows[1][i], m), out[4 * i + 2] = scale8(rows[2][i], m), out[4 * i + 3] = 255;
        }
      } else if (ncomp_ == 4) {
        ycc_to_rgb(out, rows[0], rows[1], rows[2], width_);
        if (adobe_transform_ == 2)  // YCCK
          for (int i = 0; i < width_; i++) {
            const uint8_t m = rows[3][i];
            for (int c = 0; c < 3; c++) out[4 * i + c] = scale8((uint8_t)(255 - out[4 * i + c]), m);
          }
      } else {
        for (int i = 0; i < width_; i++) out[4 * i] = out[4 * i + 1] = out[4 * i + 2] = rows[0][i], out[4 * i + 3] = 255;
      }
    }
    return true;
  }
};

// ---------------------------------------------------------------------------------------------------------------
// OpenEXR -> float rgba as tinyexr's LoadEXR returns it (exts/tinyexr/tinyexr.h:11607-11860, what load_texture calls
// for .exr): single-part scanline files, channels of type half (widened exactly), float, or uint (whose bits are then
// read as a float, like there), compression none / RLE / ZIPS / ZIP (zlib or run lengths, then the byte predictor and
// the two-halves interleave) / PIZ (below); a block whose stored size equals its raw size is taken as raw. R, G, B (+ A, else 1) are
// picked by name among the first four channels; a single channel fills all four components. Lines of a
// decreasing-Y file land mirrored, as in tinyexr. Tiled, multi-part, deep and PXR24 / B44 / DWA files are refused (the
// last three by tinyexr too).
// ---------------------------------------------------------------------------------------------------------------
float half_bits_to_float(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000) << 16, exponent = (h >> 10) & 31, mantissa = h & 0x3ff, bits;
  if (exponent == 31) {
    bits = sign | 0x7f800000u | (mantissa << 13);  // infinity, or a NaN that keeps its payload
  } else if (exponent != 0) {
    bits = sign | ((exponent + 112) << 23) | (mantissa << 13);
  } else if (mantissa == 0) {
    bits = sign;
  } else {  // a subnormal half is a normal float
    int shift = 0;
    while (!(mantissa & 0x400)) mantissa <<= 1, shift++;
    bits = sign | ((uint32_t)(113 - shift) << 23) | ((mantissa & 0x3ff) << 13);
  }
  float f;
  memcpy(&f, &bits, 4);
  return f;
}
// PIZ blocks (OpenEXR's lossless wavelet codec, 32 lines per block): [min, max of the 16-bit values in use][that slice
// of the 65536-bit "value in use" bitmap][length][Huffman data]. The 16-bit words of the block - every channel as a plane,
// a float as two interleaved words - are Huffman coded (canonical codes of up to 58 bits whose lengths arrive as a
// run-length packed 6-bit table; the last symbol in use means "repeat the previous word n times"), Haar-wavelet
// transformed level by level with the 14-bit (or, above 2^14 values in use, the 16-bit modulo) lifting step, and mapped
// through the bitmap's ranking. The codec is lossless: any correct decoder returns tinyexr's bytes.
struct PizHuffman {
  static constexpr int kSymbols = 65537, kFastBits = 14;
  std::vector<uint8_t>  length;                // per symbol
  std::vector<uint32_t> fast;                  // kFastBits-bit prefix -> symbol << 6 | length, 0 = longer code
  uint64_t              first[60] = {};        // lowest code of each length
  uint32_t              count[60] = {}, base[60] = {};
  std::vector<uint32_t> by_length;             // symbols ordered by (length, symbol)
  bool unpack(const uint8_t*& p, const uint8_t* end, int lo, int hi) {
    length.assign(kSymbols, 0);
    uint64_t acc = 0;
    int      have = 0;
    auto bits = [&](int n, int& v) {
      while (have < n) {
        if (p >= end) return false;
        acc = (acc << 8) | *p++, have += 8;
      }
      have -= n;
      v = (int)((acc >> have) & ((1u << n) - 1));
      return true;
    };
    for (int s = lo; s <= hi; s++) {
      int l;
      if (!bits(6, l)) return false;
      if (l == 63) {  // a long run of unused symbols
        int run;
        if (!bits(8, run)) return false;
        run += 6;
        if (s + run > hi + 1) return false;
        s += run - 1;
      } else if (l >= 59) {  // a short run
        const int run = l - 59 + 2;
        if (s + run > hi + 1) return false;
        s += run - 1;
      } else {
        length[s] = (uint8_t)l;
      }
    }
    // canonical codes: shorter codes are numerically higher, equal lengths ascend with the symbol
    uint64_t n[60] = {};
    for (int s = 0; s < kSymbols; s++) n[length[s]]++;
    uint64_t c = 0;
    for (int l = 58; l > 0; l--) {
      const uint64_t next = (c + n[l]) >> 1;
      count[l] = (uint32_t)n[l], first[l] = c;
      c = next;
    }
    uint32_t at = 0;
    for (int l = 1; l <= 58; l++) base[l] = at, at += count[l];
    by_length.assign(at, 0);
    uint32_t fill[60] = {};
    fast.assign((size_t)1 << kFastBits, 0);
    for (int s = 0; s < kSymbols; s++) {
      const int l = length[s];
      if (!l) continue;
      const uint32_t rank = fill[l]++;
      by_length[base[l] + rank] = (uint32_t)s;
      if (l <= kFastBits) {
        const uint64_t code = first[l] + rank;
        const size_t   from = (size_t)(code << (kFastBits - l)), span = (size_t)1 << (kFastBits - l);
        if (from + span > fast.size()) return false;
        for (size_t k = 0; k < span; k++) fast[from + k] = ((uint32_t)s << 6) | (uint32_t)l;
      }
    }
    return true;
  }
};
bool piz_huffman_decode(const uint8_t* data, size_t size, std::vector<uint16_t>& out) {
  if (size < 20) return false;
  auto u32 = [&](size_t at) { uint32_t v; memcpy(&v, data + at, 4); return v; };
  const uint32_t lo = u32(0), hi = u32(4), nbits = u32(12);
  if (lo >= 65537 || hi >= 65537) return false;
  const uint8_t *p = data + 20, *end = data + size;
  PizHuffman     table;
  if (!table.unpack(p, end, (int)lo, (int)hi)) return false;
  if ((uint64_t)nbits > (uint64_t)(end - p) * 8) return false;
  // MSB-first bit reader over exactly nbits bits
  uint64_t acc = 0, used = 0;
  int      have = 0;
  auto refill = [&]() {
    while (have <= 56 && p < end) acc |= (uint64_t)*p++ << (56 - have), have += 8;
  };
  size_t written = 0;
  while (used < nbits) {
    refill();
    uint32_t symbol;
    int      l;
    const uint32_t quick = table.fast[acc >> (64 - PizHuffman::kFastBits)];
    if (quick) {
      symbol = quick >> 6, l = (int)(quick & 63);
    } else {
      l = PizHuffman::kFastBits + 1;
      for (;; l++) {
        if (l > 58) return false;
        const uint64_t code = acc >> (64 - l);
        if (table.count[l] && code >= table.first[l] && code - table.first[l] < table.count[l]) {
          symbol = table.by_length[table.base[l] + (uint32_t)(code - table.first[l])];
          break;
        }
      }
    }
    if (used + (uint64_t)l > nbits) return false;
    acc <<= l, have -= l, used += (uint64_t)l;
    if (symbol == hi) {  // the run-length symbol: eight more bits give the count
      refill();
      if (used + 8 > nbits) return false;
      const unsigned run = (unsigned)(acc >> 56);
      acc <<= 8, have -= 8, used += 8;
      if (written == 0 || written + run > out.size()) return false;
      for (unsigned k = 0; k < run; k++) out[written + k] = out[written - 1];
      written += run;
    } else {
      if (written >= out.size()) return false;
      out[written++] = (uint16_t)symbol;
    }
  }
  return true;  // (tinyexr does not insist on the word count either)
}
// one inverse lifting step: (low, high) -> the two samples
inline void piz_unlift(bool narrow, uint16_t l, uint16_t h, uint16_t& a, uint16_t& b) {
  if (narrow) {
    const int hi = (int16_t)h, ai = (int16_t)l + (hi & 1) + (hi >> 1);
    a = (uint16_t)(int16_t)ai, b = (uint16_t)(int16_t)(ai - hi);
  } else {
    const int bb = ((int)l - ((int)h >> 1)) & 0xffff;
    a = (uint16_t)(((int)h + bb - 0x8000) & 0xffff), b = (uint16_t)bb;
  }
}
// Inverse 2-D Haar transform of an nx x ny grid of words (`sx`, `sy`: distance between neighbours in x and in y), in
// place, from the coarsest level down. At the level with step `half`, the word at a grid point (x, y) with x, y
// multiples of 2 * half is an average and its neighbours at +half are differences: first the two vertical pairs of a
// 2 x 2 cell are undone, then the two horizontal ones; a last odd column or row has only its one pair.
void piz_wavelet_decode(uint16_t* words, int nx, int sx, int ny, int sy, uint16_t max_value) {
  const bool narrow = max_value < (1 << 14);
  auto at = [&](int x, int y) -> uint16_t& { return words[(ptrdiff_t)x * sx + (ptrdiff_t)y * sy]; };
  int  step = 1;
  while (step <= std::min(nx, ny)) step <<= 1;
  step >>= 1;  // the largest power of two that still fits the smaller side: cell size of the coarsest level
  for (int half = step >> 1; half >= 1; step = half, half >>= 1) {
    int y = 0;
    for (; y + step <= ny; y += step) {
      int x = 0;
      for (; x + step <= nx; x += step) {
        uint16_t top_left, bottom_left, top_right, bottom_right;
        piz_unlift(narrow, at(x, y), at(x, y + half), top_left, bottom_left);
        piz_unlift(narrow, at(x + half, y), at(x + half, y + half), top_right, bottom_right);
        piz_unlift(narrow, top_left, top_right, at(x, y), at(x + half, y));
        piz_unlift(narrow, bottom_left, bottom_right, at(x, y + half), at(x + half, y + half));
      }
      if (nx & half) {  // a column without a right-hand partner
        uint16_t top, bottom;
        piz_unlift(narrow, at(x, y), at(x, y + half), top, bottom);
        at(x, y) = top, at(x, y + half) = bottom;
      }
    }
    if (ny & half) {  // a row without a partner below
      for (int x = 0; x + step <= nx; x += step) {
        uint16_t left, right;
        piz_unlift(narrow, at(x, y), at(x + half, y), left, right);
        at(x, y) = left, at(x + half, y) = right;
      }
    }
  }
}
bool piz_decode(const uint8_t* src, size_t stored, const std::vector<int>& channel_bytes, int width, int lines, uint8_t* out, size_t size) {
  if (stored == size) return memcpy(out, src, size), true;
  if (stored < 4) return false;
  uint16_t min_used, max_used;
  memcpy(&min_used, src, 2), memcpy(&max_used, src + 2, 2);
  std::vector<uint8_t> bitmap(8192, 0);
  size_t at = 4;
  if (max_used >= 8192) return false;
  if (min_used <= max_used) {
    const size_t n = (size_t)max_used - min_used + 1;
    if (stored < at + n) return false;
    memcpy(&bitmap[min_used], src + at, n);
    at += n;
  }
  std::vector<uint16_t> ranked(65536, 0);  // rank among the values in use -> value
  size_t                in_use = 0;
  for (int v = 0; v < 65536; v++)
    if (v == 0 || (bitmap[v >> 3] & (1 << (v & 7)))) ranked[in_use++] = (uint16_t)v;
  const uint16_t max_value = (uint16_t)(in_use - 1);
  if (stored < at + 4) return false;
  int32_t length;
  memcpy(&length, src + at, 4);
  at += 4;
  if (length < 0 || at + (size_t)length > stored) return false;
  std::vector<uint16_t> words(size / 2, 0);
  if (length > 0) piz_huffman_decode(src + at, (size_t)length, words);  // (a failed block decodes to what it got, like there)
  size_t start = 0;
  for (int bytes : channel_bytes) {
    const int per = bytes / 2;
    for (int j = 0; j < per; j++) piz_wavelet_decode(&words[start + j], width, per, lines, width * per, max_value);
    start += (size_t)width * lines * per;
  }
  for (auto& w : words) w = ranked[w];
  std::vector<size_t> cursor(channel_bytes.size());
  start = 0;
  for (size_t c = 0; c < channel_bytes.size(); c++) cursor[c] = start, start += (size_t)width * lines * (channel_bytes[c] / 2);
  for (int y = 0; y < lines; y++)
    for (size_t c = 0; c < channel_bytes.size(); c++) {
      const size_t n = (size_t)width * (channel_bytes[c] / 2);
      memcpy(out, &words[cursor[c]], n * 2);
      out += n * 2, cursor[c] += n;
    }
  return true;
}

bool load_exr(const std::vector<uint8_t>& file, HostTexture& tex) {
  const uint8_t *p = file.data(), *end = p + file.size();
  auto i32 = [&](const uint8_t* at) { int32_t v; memcpy(&v, at, 4); return v; };
  if (file.size() < 8 || i32(p) != 20000630 || p[4] != 2) return false;
  const uint32_t flags = (uint32_t)i32(p + 4) >> 8;  // 0x2 tiled, 0x4 long names, 0x8 deep, 0x10 multi-part
  if (flags & (0x2 | 0x8 | 0x10)) return false;
  p += 8;
  struct Channel {
    std::string name;
    int         type = 0, bytes = 0;
    size_t      offset = 0;
  };
  std::vector<Channel> channels;
  int  compression = -1, line_order = 0, window[4] = {0, 0, -1, -1};
  bool have_window = false;
  while (true) {  // attributes: name\0 type\0 size data
    if (p >= end) return false;
    if (*p == 0) {
      p++;
      break;
    }
    const uint8_t* z = (const uint8_t*)memchr(p, 0, end - p);
    if (!z) return false;
    const std::string name((const char*)p, (const char*)z);
    p = z + 1;
    z = (const uint8_t*)memchr(p, 0, end - p);
    if (!z) return false;
    const std::string type((const char*)p, (const char*)z);
    p = z + 1;
    if (end - p < 4) return false;
    const int64_t size = i32(p);
    p += 4;
    if (size < 0 || end - p < size) return false;
    if (name == "channels") {
      const uint8_t *c = p, *cend = p + size;
      while (c < cend && *c) {
        const uint8_t* cz = (const uint8_t*)memchr(c, 0, cend - c);
        if (!cz || cend - cz < 17) return false;
        Channel ch;
        ch.name = std::string((const char*)c, (const char*)cz);
        ch.type = i32(cz + 1);
        if (ch.type < 0 || ch.type > 2) return false;
        if (i32(cz + 9) != 1 || i32(cz + 13) != 1) return false;  // subsampled channels
        ch.bytes = ch.type == 1 ? 2 : 4;
        channels.push_back(ch);
        c = cz + 17;
      }
    } else if (name == "compression") {
      if (size < 1) return false;
      compression = p[0];
    } else if (name == "dataWindow") {
      if (size < 16) return false;
      for (int k = 0; k < 4; k++) window[k] = i32(p + 4 * k);
      have_window = true;
    } else if (name == "lineOrder") {
      if (size < 1) return false;
      line_order = p[0];
    }
    p += size;
  }
  if (channels.empty() || !have_window || compression < 0 || compression > 4) return false;
  if (window[2] < window[0] || window[3] < window[1]) return false;
  const int64_t width = (int64_t)window[2] - window[0] + 1, height = (int64_t)window[3] - window[1] + 1;
  if (width > (1 << 23) || height > (1 << 23) || width * height > (int64_t(1) << 28)) return false;
  size_t pixel_bytes = 0;
  for (auto& ch : channels) ch.offset = pixel_bytes, pixel_bytes += ch.bytes;
  const int    block_lines = compression == 3 ? 16 : compression == 4 ? 32 : 1;
  const size_t num_blocks  = (size_t)((height + block_lines - 1) / block_lines);
  if ((size_t)(end - p) < num_blocks * 8) return false;
  // every channel as 32-bit words (floats, or the uint's bits)
  std::vector<std::vector<uint32_t>> planes(channels.size(), std::vector<uint32_t>((size_t)(width * height), 0));
  std::vector<uint8_t> raw, scratch;
  for (size_t b = 0; b < num_blocks; b++) {
    uint64_t offset;
    memcpy(&offset, p + 8 * b, 8);
    if (offset >= file.size() || file.size() - offset < 8) return false;
    const uint8_t* chunk = file.data() + offset;
    int64_t        line  = i32(chunk);
    const int64_t  stored = i32(chunk + 4);
    if (stored <= 0 || (uint64_t)stored > file.size() - offset - 8) return false;
    if (line > (2 << 20) || line < -(2 << 20)) return false;
    const int64_t last = std::min<int64_t>(line + block_lines, (int64_t)window[3] + 1);
    const int64_t lines = last - line;
    line -= window[1];
    if (lines <= 0 || line < 0 || line + lines > height) return false;
    const size_t size = (size_t)(width * lines) * pixel_bytes;
    raw.resize(size);
    const uint8_t* src = chunk + 8;
    if (compression == 0 || (size_t)stored == size) {
      if ((size_t)stored < size) return false;
      memcpy(raw.data(), src, size);
    } else if (compression == 4) {
      std::vector<int> channel_bytes;
      for (auto& ch : channels) channel_bytes.push_back(ch.bytes);
      if (!piz_decode(src, (size_t)stored, channel_bytes, (int)width, (int)lines, raw.data(), size)) return false;
    } else {
      scratch.resize(size);
      if (compression == 1) {  // run lengths
        if (stored <= 2) return false;
        size_t  out = 0;
        int64_t in  = 0;
        while (in < stored) {
          const int8_t code = (int8_t)src[in++];
          if (code < 0) {
            const size_t count = (size_t)(-(int)code);
            if (out + count > size || in + (int64_t)count > stored) return false;
            memcpy(&scratch[out], src + in, count);
            out += count, in += (int64_t)count;
          } else {
            const size_t count = (size_t)code + 1;
            if (out + count > size || in >= stored) return false;
            memset(&scratch[out], src[in++], count);
            out += count;
          }
        }
        if (out != size) return false;
      } else {
        uLongf got = (uLongf)size;
        if (uncompress(scratch.data(), &got, src, (uLong)stored) != Z_OK) return false;
        // (a shorter result leaves the tail as it is: zeros)
        if (got < size) memset(&scratch[got], 0, size - got);
      }
      for (size_t i = 1; i < size; i++) scratch[i] = (uint8_t)(scratch[i - 1] + scratch[i] - 128);  // predictor
      const size_t half = (size + 1) / 2;                                                            // interleave
      for (size_t i = 0; i < size; i++) raw[i] = scratch[(i & 1) ? half + i / 2 : i / 2];
    }
    for (size_t c = 0; c < channels.size(); c++)
      for (int64_t v = 0; v < lines; v++) {
        const uint8_t* row = &raw[(size_t)v * pixel_bytes * (size_t)width + channels[c].offset * (size_t)width];
        const int64_t  y   = line_order == 0 ? line + v : height - 1 - (line + v);
        uint32_t*      dst = &planes[c][(size_t)(y * width)];
        if (channels[c].type == 1) {
          for (int64_t u = 0; u < width; u++) {
            uint16_t h;
            memcpy(&h, row + 2 * u, 2);
            const float f = half_bits_to_float(h);
            memcpy(&dst[u], &f, 4);
          }
        } else {
          memcpy(dst, row, (size_t)width * 4);
        }
      }
  }
  // LoadEXR's channel choice: names among the first four channels of the default layer (names without a '.')
  std::vector<size_t> layer;
  for (size_t c = 0; c < channels.size(); c++)
    if (channels[c].name.find('.') == std::string::npos) layer.push_back(c);
  if (layer.empty()) return false;
  int r = -1, g = -1, b = -1, a = -1;
  for (size_t k = 0; k < layer.size() && k < 4; k++) {
    const auto& name = channels[layer[k]].name;
    if (name == "R") r = (int)layer[k];
    else if (name == "G") g = (int)layer[k];
    else if (name == "B") b = (int)layer[k];
    else if (name == "A") a = (int)layer[k];
  }
  tex.width = (int)width, tex.height = (int)height;
  tex.pixelsf.resize((size_t)(width * height) * 4);
  const float one = 1.0f;
  uint32_t    one_bits;
  memcpy(&one_bits, &one, 4);
  uint32_t* out = (uint32_t*)tex.pixelsf.data();
  if (layer.size() == 1) {
    const auto& plane = planes[layer[0]];
    for (size_t i = 0; i < plane.size(); i++) out[4 * i] = out[4 * i + 1] = out[4 * i + 2] = out[4 * i + 3] = plane[i];
    return true;
  }
  if (r < 0 || g < 0 || b < 0) return false;
  for (size_t i = 0; i < (size_t)(width * height); i++)
    out[4 * i] = planes[r][i], out[4 * i + 1] = planes[g][i], out[4 * i + 2] = planes[b][i], out[4 * i + 3] = a >= 0 ? planes[a][i] : one_bits;
  return true;
}

}  // namespace

// load_texture, yocto_sceneio.cpp:1796-1837: the file type decides `linear`; nearest / clamp come from the JSON
bool load_texture(const std::string& filename, HostTexture& tex, std::string& error) {
  const auto ext = path_extension(filename);
  if (ext != ".hdr" && ext != ".png" && ext != ".jpg" && ext != ".jpeg" && ext != ".exr") return error = "unsupported format " + filename, false;
  std::vector<uint8_t> data;
  if (!read_file(filename, data, error)) return false;
  if (ext == ".jpg" || ext == ".jpeg") {
    if (!JpegDecoder().decode(data, tex)) return error = "cannot raed " + filename, false;
    tex.linear = 0;
  } else if (ext == ".exr") {
    if (!load_exr(data, tex)) return error = "cannot raed " + filename, false;
    tex.linear = 1;
  } else if (ext == ".hdr") {
    if (!load_hdr(data, tex)) return error = "cannot raed " + filename, false;  // (the reference's own spelling)
    tex.linear = 1;
  } else {
    if (!load_png(data, tex)) return error = "cannot raed " + filename, false;
    tex.linear = 0;
  }
  return true;
}

}  // namespace ygl_io
