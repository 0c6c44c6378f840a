// png_decode.cpp — PNG -> RGBA8 for glTF textures (host only; zlib does the inflate).
//
// The reference loads texture images through tinygltf, i.e. stb_image with 4 requested components (scene.cpp:144-159, 513-646).  What
// reaches the texture is therefore stb_image's reading of the PNG specification, restated here:
//   * all colour types (grey, RGB, palette, grey + alpha, RGBA), bit depths 1 / 2 / 4 / 8 / 16, Adam7 interlacing;
//   * 16-bit samples keep their HIGH byte; 1 / 2 / 4-bit grey samples are scaled by 255 / 85 / 17; palette indices are looked up;
//   * tRNS: per-index alpha for palette images, a colour key (exact sample match, alpha 0) for grey / RGB images;
//   * grey -> (g, g, g), missing alpha -> 255; gAMA / sRGB / iCCP and every other ancillary chunk are ignored; CRCs are not checked.
#include <zlib.h>
#include <cstring>
#include "common.h"
#include "gltf_import.h"

namespace eid {

static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

static int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// undo the scanline filters of one (sub-)image in place; `raw` holds h rows of (1 + rowBytes) bytes; returns rows of rowBytes
static void unfilter(uint8_t* raw, size_t rowBytes, uint32_t h, int bpp, std::vector<uint8_t>& out) {
  out.assign(rowBytes * h, 0);
  for (uint32_t y = 0; y < h; ++y) {
    const uint8_t* src = raw + (size_t)y * (rowBytes + 1);
    const int ft = src[0];
    ++src;
    uint8_t* cur = out.data() + (size_t)y * rowBytes;
    const uint8_t* up = y ? cur - rowBytes : nullptr;
    if (ft > 4) raise(EID_ERR_PARSE, "PNG: invalid filter type %d", ft);
    for (size_t i = 0; i < rowBytes; ++i) {
      const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= (size_t)bpp) ? up[i - bpp] : 0;
      int v = src[i];
      switch (ft) {
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: v += paeth(a, b, c); break;
        default: break;
      }
      cur[i] = (uint8_t)v;
    }
  }
}

bool isPng(const uint8_t* d, size_t n) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  return n >= 8 && !memcmp(d, sig, 8);
}

void decodePng(const uint8_t* d, size_t n, HostGltf::Image& img) {
  if (!isPng(d, n)) raise(EID_ERR_PARSE, "PNG: bad signature");
  uint32_t W = 0, H = 0;
  int depth = 0, ctype = 0, interlace = 0;
  std::vector<uint8_t> idat, plte, trns;
  bool haveHdr = false, done = false;
  for (size_t p = 8; !done;) {
    if (p + 12 > n) raise(EID_ERR_PARSE, "PNG: truncated chunk");
    const uint32_t len = be32(d + p);
    const uint8_t* tag = d + p + 4;
    if (len > n - p - 12) raise(EID_ERR_PARSE, "PNG: chunk overruns the file");
    const uint8_t* body = d + p + 8;
    if (!memcmp(tag, "IHDR", 4)) {
      if (len != 13) raise(EID_ERR_PARSE, "PNG: bad IHDR");
      W = be32(body); H = be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
      if (!W || !H || W > 32768 || H > 32768) raise(EID_ERR_PARSE, "PNG: bad size %ux%u", W, H);
      if (body[10] || body[11] || interlace > 1) raise(EID_ERR_PARSE, "PNG: unknown compression / filter / interlace method");
      const bool okDepth = (ctype == 0 && (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) ||
                           (ctype == 3 && (depth == 1 || depth == 2 || depth == 4 || depth == 8)) ||
                           ((ctype == 2 || ctype == 4 || ctype == 6) && (depth == 8 || depth == 16));
      if (!okDepth) raise(EID_ERR_PARSE, "PNG: colour type %d with bit depth %d", ctype, depth);
      haveHdr = true;
    } else if (!memcmp(tag, "PLTE", 4)) plte.assign(body, body + len);
    else if (!memcmp(tag, "tRNS", 4)) trns.assign(body, body + len);
    else if (!memcmp(tag, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
    else if (!memcmp(tag, "IEND", 4)) done = true;
    p += 12 + (size_t)len;
  }
  if (!haveHdr || idat.empty()) raise(EID_ERR_PARSE, "PNG: missing IHDR / IDAT");
  if (ctype == 3 && (plte.empty() || plte.size() % 3)) raise(EID_ERR_PARSE, "PNG: palette image without a valid PLTE");
  const int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : 4;
  const int bpp = std::max(1, channels * depth / 8);
  auto rowBytesOf = [&](uint32_t w) { return ((size_t)w * channels * depth + 7) / 8; };

  // sub-images: the whole image, or the 7 Adam7 passes
  struct Pass { uint32_t x0, y0, dx, dy, w, h; };
  std::vector<Pass> passes;
  if (!interlace) passes.push_back({0, 0, 1, 1, W, H});
  else {
    static const uint32_t X0[7] = {0, 4, 0, 2, 0, 1, 0}, Y0[7] = {0, 0, 4, 0, 2, 0, 1}, DX[7] = {8, 8, 4, 4, 2, 2, 1}, DY[7] = {8, 8, 8, 4, 4, 2, 2};
    for (int k = 0; k < 7; ++k) {
      const uint32_t w = (W > X0[k]) ? (W - X0[k] + DX[k] - 1) / DX[k] : 0, h = (H > Y0[k]) ? (H - Y0[k] + DY[k] - 1) / DY[k] : 0;
      if (w && h) passes.push_back({X0[k], Y0[k], DX[k], DY[k], w, h});
    }
  }
  size_t rawSize = 0;
  for (const auto& ps : passes) rawSize += (rowBytesOf(ps.w) + 1) * ps.h;
  std::vector<uint8_t> raw(rawSize);
  {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit(&zs) != Z_OK) raise(EID_ERR_PARSE, "PNG: inflateInit failed");
    zs.next_in = idat.data(); zs.avail_in = (uInt)idat.size();
    zs.next_out = raw.data(); zs.avail_out = (uInt)raw.size();
    const int rc = inflate(&zs, Z_FINISH);
    const size_t got = zs.total_out;
    inflateEnd(&zs);
    if ((rc != Z_STREAM_END && rc != Z_OK && rc != Z_BUF_ERROR) || got < raw.size()) raise(EID_ERR_PARSE, "PNG: corrupt or short image data (zlib %d, %zu of %zu bytes)", rc, got, raw.size());
  }

  img.width = W; img.height = H;
  img.rgba8.assign((size_t)W * H * 4, 255);
  const int scale = depth == 1 ? 255 : depth == 2 ? 85 : depth == 4 ? 17 : 1;
  uint32_t key[3] = {0, 0, 0};
  const bool hasKey = (ctype == 0 && trns.size() >= 2) || (ctype == 2 && trns.size() >= 6);
  if (hasKey) for (int k = 0; k < (ctype == 0 ? 1 : 3); ++k) key[k] = ((uint32_t)trns[2 * k] << 8) | trns[2 * k + 1];
  size_t off = 0;
  std::vector<uint8_t> rows;
  for (const auto& ps : passes) {
    const size_t rb = rowBytesOf(ps.w);
    unfilter(raw.data() + off, rb, ps.h, bpp, rows);
    off += (rb + 1) * ps.h;
    for (uint32_t y = 0; y < ps.h; ++y) {
      const uint8_t* r = rows.data() + (size_t)y * rb;
      for (uint32_t x = 0; x < ps.w; ++x) {
        uint32_t s[4] = {0, 0, 0, 0};                        // raw samples of the pixel (full 16 bits when depth == 16)
        for (int c = 0; c < channels; ++c) {
          if (depth == 8) s[c] = r[(size_t)x * channels + c];
          else if (depth == 16) s[c] = ((uint32_t)r[2 * ((size_t)x * channels + c)] << 8) | r[2 * ((size_t)x * channels + c) + 1];
          else { const size_t bit = (size_t)x * depth; s[c] = (r[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u); }
        }
        uint8_t* o = img.rgba8.data() + ((size_t)(ps.y0 + y * ps.dy) * W + (ps.x0 + x * ps.dx)) * 4;
        auto to8 = [&](uint32_t v) { return (uint8_t)(depth == 16 ? (v >> 8) : (depth < 8 ? v * scale : v)); };
        if (ctype == 3) {
          if (s[0] * 3 + 2 >= plte.size()) raise(EID_ERR_PARSE, "PNG: palette index %u out of range", s[0]);
          o[0] = plte[3 * s[0]]; o[1] = plte[3 * s[0] + 1]; o[2] = plte[3 * s[0] + 2];
          o[3] = s[0] < trns.size() ? trns[s[0]] : 255;
        } else if (ctype == 0 || ctype == 4) {
          o[0] = o[1] = o[2] = to8(s[0]);
          o[3] = ctype == 4 ? to8(s[1]) : (hasKey && s[0] == key[0] ? 0 : 255);
        } else {
          o[0] = to8(s[0]); o[1] = to8(s[1]); o[2] = to8(s[2]);
          o[3] = ctype == 6 ? to8(s[3]) : (hasKey && s[0] == key[0] && s[1] == key[1] && s[2] == key[2] ? 0 : 255);
        }
      }
    }
  }
}

}  // namespace eid
