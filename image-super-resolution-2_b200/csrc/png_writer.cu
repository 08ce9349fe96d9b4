// ff_png_encode_rgb8: host-side PNG writer of the plugin's save path (reference models/team29_FreqFusion/io.py:71-76 _save_image,
// Image.fromarray(arr).save(path, format="PNG")).  Pixel-identical files; the encoder is built for the throughput the GPU path
// needs (one 512 x 512 tile in ~1.5 ms on one host core, zlib's fastest setting takes 13-15 ms):
//   * every scanline Sub-filtered (PNG filter type 1),
//   * ONE deflate block with a dynamic Huffman code over literals only (no LZ77 matches: on photographic / super-resolved
//     content after the Sub filter a match search buys nothing -- zlib's Z_RLE and Z_HUFFMAN_ONLY give the same size),
//   * code lengths from a histogram pass (Huffman by two-queue merge, limited to 15 bits), emitted through a 64-bit bit buffer,
//   * slicing-by-8 CRC-32 and a blocked Adler-32.
// Plain host C++ (compiled by nvcc with the rest of the library); called through ctypes, which releases the GIL, so the
// plugin's encoder threads run in parallel.
#include "ff_common.cuh"
#include "../../include/ffb200.h"
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <dlfcn.h>
#include <vector>

namespace {

uint32_t g_crc_tab[8][256];
bool g_crc_ready = false;
void crc_init() {
  // (idempotent: concurrent first calls write identical values)
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
    g_crc_tab[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) g_crc_tab[t][i] = (g_crc_tab[t - 1][i] >> 8) ^ g_crc_tab[0][g_crc_tab[t - 1][i] & 0xFF];
  __atomic_store_n(&g_crc_ready, true, __ATOMIC_RELEASE);
}
uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
  if (!__atomic_load_n(&g_crc_ready, __ATOMIC_ACQUIRE)) crc_init();
  crc = ~crc;
  while (n && (reinterpret_cast<uintptr_t>(p) & 7)) { crc = g_crc_tab[0][(crc ^ *p++) & 0xFF] ^ (crc >> 8); --n; }
  while (n >= 8) {
    uint64_t v;
    memcpy(&v, p, 8);
    v ^= crc;
    crc = g_crc_tab[7][v & 0xFF] ^ g_crc_tab[6][(v >> 8) & 0xFF] ^ g_crc_tab[5][(v >> 16) & 0xFF] ^ g_crc_tab[4][(v >> 24) & 0xFF] ^
          g_crc_tab[3][(v >> 32) & 0xFF] ^ g_crc_tab[2][(v >> 40) & 0xFF] ^ g_crc_tab[1][(v >> 48) & 0xFF] ^ g_crc_tab[0][v >> 56];
    p += 8; n -= 8;
  }
  while (n--) crc = g_crc_tab[0][(crc ^ *p++) & 0xFF] ^ (crc >> 8);
  return ~crc;
}
uint32_t adler32_update(uint32_t adler, const uint8_t* p, size_t n) {
  uint32_t a = adler & 0xFFFF, b = adler >> 16;
  while (n) {
    size_t k = n < 5552 ? n : 5552;      // largest run before the 32-bit sums can overflow
    n -= k;
    for (; k >= 8; k -= 8, p += 8) {      // eight bytes per step: b gains 8 a + the weighted bytes
      const uint32_t s = p[0] + p[1] + p[2] + p[3] + p[4] + p[5] + p[6] + p[7];
      b += 8 * a + 8 * p[0] + 7 * p[1] + 6 * p[2] + 5 * p[3] + 4 * p[4] + 3 * p[5] + 2 * p[6] + p[7];
      a += s;
    }
    for (; k; --k) { a += *p++; b += a; }
    a %= 65521; b %= 65521;
  }
  return (b << 16) | a;
}

constexpr int NSYM = 257;      // literals 0..255 + end of block
constexpr int MAXBITS = 15;

// Huffman code lengths (<= 15 bits) of the symbols with freq > 0; canonical codes, bit-reversed for deflate's LSB-first packing.
void build_code(const uint32_t* freq, uint8_t* len, uint16_t* code) {
  int order[NSYM], n = 0;
  for (int s = 0; s < NSYM; ++s) { len[s] = 0; if (freq[s]) order[n++] = s; }
  if (n == 1) { len[order[0]] = 1; }      // (cannot happen with an end-of-block symbol next to data, kept for safety)
  else {
    std::sort(order, order + n, [&](int x, int y) { return freq[x] != freq[y] ? freq[x] < freq[y] : x < y; });
    // two-queue Huffman: leaves in increasing frequency, internal nodes are created in increasing weight
    uint64_t w[2 * NSYM];
    int parent[2 * NSYM];
    for (int i = 0; i < n; ++i) w[i] = freq[order[i]];
    int leaf = 0, inode = n, next = n;
    auto take = [&]() { return (leaf < n && (inode >= next || w[leaf] <= w[inode])) ? leaf++ : inode++; };
    while (next < 2 * n - 1) {
      const int x = take(), y = take();
      w[next] = w[x] + w[y];
      parent[x] = parent[y] = next;
      ++next;
    }
    int depth[2 * NSYM];
    depth[2 * n - 2] = 0;
    for (int i = 2 * n - 3; i >= 0; --i) depth[i] = depth[parent[i]] + 1;
    int count[64] = {0};
    for (int i = 0; i < n; ++i) ++count[depth[i] < 63 ? depth[i] : 63];
    // limit to 15 bits: fold the deeper levels into level 15, then repair the Kraft sum by splitting shallower leaves
    for (int d = MAXBITS + 1; d < 64; ++d) { count[MAXBITS] += count[d]; count[d] = 0; }
    uint64_t total = 0;
    for (int d = 1; d <= MAXBITS; ++d) total += (uint64_t)count[d] << (MAXBITS - d);
    while (total > (1ull << MAXBITS)) {
      --count[MAXBITS];
      for (int d = MAXBITS - 1; d > 0; --d)
        if (count[d]) { --count[d]; count[d + 1] += 2; break; }
      --total;
    }
    // the most frequent symbols get the shortest codes (order[] is increasing in frequency)
    int i = n - 1;
    for (int d = 1; d <= MAXBITS; ++d)
      for (int k = 0; k < count[d]; ++k) len[order[i--]] = (uint8_t)d;
  }
  int bl_count[MAXBITS + 1] = {0};
  for (int s = 0; s < NSYM; ++s) ++bl_count[len[s]];
  bl_count[0] = 0;
  uint32_t next_code[MAXBITS + 2] = {0};
  uint32_t c = 0;
  for (int b = 1; b <= MAXBITS; ++b) { c = (c + bl_count[b - 1]) << 1; next_code[b] = c; }
  for (int s = 0; s < NSYM; ++s) {
    if (!len[s]) { code[s] = 0; continue; }
    uint32_t v = next_code[len[s]]++, r = 0;
    for (int b = 0; b < len[s]; ++b) { r = (r << 1) | (v & 1); v >>= 1; }
    code[s] = (uint16_t)r;
  }
}

struct BitWriter {
  uint8_t* p;
  uint64_t buf = 0;
  int cnt = 0;
  inline void put(uint32_t v, int n) {
    buf |= (uint64_t)v << cnt;
    cnt += n;
    if (cnt >= 32) {
      memcpy(p, &buf, 4);
      p += 4;
      buf >>= 32;
      cnt -= 32;
    }
  }
  inline void finish() {
    while (cnt > 0) { *p++ = (uint8_t)buf; buf >>= 8; cnt -= 8; }
    cnt = 0;
  }
};

inline void be32(uint8_t* p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }

}  // namespace

extern "C" long long ff_png_bound_rgb8(int h, int w) {
  const long long raw = (long long)h * (1 + 3LL * w);
  return raw * 15 / 8 + 1024;      // worst case: every literal at 15 bits, plus headers
}

// rgb: uint8 [h][w][3] with `row_stride` bytes between rows; out: at least ff_png_bound_rgb8(h, w) bytes.
// Returns the file size, or a negative error code.
extern "C" long long ff_png_encode_rgb8(const unsigned char* rgb, int h, int w, long long row_stride, unsigned char* out, long long cap) {
  FF_CHECK_ARG(rgb && out && h > 0 && w > 0 && row_stride >= 3LL * w, "ff_png_encode_rgb8: bad arguments");
  FF_CHECK_ARG(cap >= ff_png_bound_rgb8(h, w), "ff_png_encode_rgb8: output buffer of %lld bytes, need %lld", cap, ff_png_bound_rgb8(h, w));
  const size_t rb = 1 + 3 * (size_t)w;
  uint8_t* filt = static_cast<uint8_t*>(malloc(rb * h));
  if (!filt) { ff_set_error("ff_png_encode_rgb8: out of host memory"); return FF_ERR_ARG; }
  uint32_t freq[NSYM] = {0};
  uint32_t f4[4][256];      // four histograms: consecutive bytes do not serialise on one counter
  memset(f4, 0, sizeof(f4));
  for (int y = 0; y < h; ++y) {
    const uint8_t* s = rgb + (size_t)y * row_stride;
    uint8_t* d = filt + (size_t)y * rb;
    d[0] = 1;      // filter type Sub
    d[1] = s[0]; d[2] = s[1]; d[3] = s[2];
    const size_t n = 3 * (size_t)w;
    for (size_t i = 3; i < n; ++i) d[1 + i] = (uint8_t)(s[i] - s[i - 3]);
    size_t i = 0;
    for (; i + 4 <= rb; i += 4) { ++f4[0][d[i]]; ++f4[1][d[i + 1]]; ++f4[2][d[i + 2]]; ++f4[3][d[i + 3]]; }
    for (; i < rb; ++i) ++f4[0][d[i]];
  }
  for (int s = 0; s < 256; ++s) freq[s] = f4[0][s] + f4[1][s] + f4[2][s] + f4[3][s];
  freq[256] = 1;
  uint8_t len[NSYM];
  uint16_t code[NSYM];
  build_code(freq, len, code);

  uint8_t* p = out;
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
  memcpy(p, sig, 8); p += 8;
  be32(p, 13); memcpy(p + 4, "IHDR", 4); be32(p + 8, (uint32_t)w); be32(p + 12, (uint32_t)h);
  p[16] = 8; p[17] = 2; p[18] = 0; p[19] = 0; p[20] = 0;      // 8-bit truecolour, deflate, adaptive filtering, no interlace
  be32(p + 21, crc32_update(0, p + 4, 17));
  p += 25;
  uint8_t* idat = p;      // length patched below
  memcpy(p + 4, "IDAT", 4);
  p += 8;
  *p++ = 0x78; *p++ = 0x01;      // zlib header: deflate, 32 KB window, fastest
  BitWriter bw;
  bw.p = p;
  bw.put(1, 1);      // BFINAL
  bw.put(2, 2);      // BTYPE = dynamic Huffman
  bw.put(NSYM - 257, 5);      // HLIT
  bw.put(1, 5);               // HDIST: two distance codes of one bit each, never used -- the form zlib itself writes for
                              // match-free data ("at least one distance code exists", trees.c), accepted by every inflater
  bw.put(19 - 4, 4);          // HCLEN: all 19 code-length-code lengths follow
  // code-length alphabet: symbols 0..15 with 4 bits each (a complete code), the run-length symbols 16 / 17 / 18 unused
  static const int cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  for (int i = 0; i < 19; ++i) bw.put(cl_order[i] < 16 ? 4 : 0, 3);
  auto put_cl = [&](int v) {      // canonical 4-bit code of symbol v is v itself; deflate packs Huffman codes MSB first
    const uint32_t r = ((v & 1) << 3) | ((v & 2) << 1) | ((v & 4) >> 1) | ((v & 8) >> 3);
    bw.put(r, 4);
  };
  for (int s = 0; s < NSYM; ++s) put_cl(len[s]);
  put_cl(1);
  put_cl(1);
  const size_t total = rb * (size_t)h;
  uint32_t cl[256];      // code | length << 16
  for (int s = 0; s < 256; ++s) cl[s] = code[s] | ((uint32_t)len[s] << 16);
  size_t i = 0;
  for (; i + 2 <= total; i += 2) {      // two literals per bit-buffer update (at most 30 bits on top of < 32 pending)
    const uint32_t e0 = cl[filt[i]], e1 = cl[filt[i + 1]];
    const int n0 = (int)(e0 >> 16);
    bw.buf |= ((uint64_t)(e0 & 0xFFFF) | ((uint64_t)(e1 & 0xFFFF) << n0)) << bw.cnt;
    bw.cnt += n0 + (int)(e1 >> 16);
    if (bw.cnt >= 32) {
      memcpy(bw.p, &bw.buf, 4);
      bw.p += 4;
      bw.buf >>= 32;
      bw.cnt -= 32;
    }
  }
  for (; i < total; ++i) bw.put(code[filt[i]], len[filt[i]]);
  bw.put(code[256], len[256]);
  bw.finish();
  p = bw.p;
  be32(p, adler32_update(1, filt, total));
  p += 4;
  free(filt);
  const uint32_t idat_len = (uint32_t)(p - (idat + 8));
  be32(idat, idat_len);
  be32(p, crc32_update(0, idat + 4, idat_len + 4));
  p += 4;
  be32(p, 0); memcpy(p + 4, "IEND", 4); be32(p + 8, crc32_update(0, p + 4, 4));
  p += 12;
  return (long long)(p - out);
}


// ------------------------------------------------------------------------------------------------
// ff_png_decode_rgb8: host-side PNG reader of the plugin's load path (models/team29_FreqFusion/io.py:64-68 _load_image,
// Image.open(path).convert("RGB")) for the files a test set holds: 8-bit grey / grey+alpha / RGB / RGBA, non-interlaced.  Chunk
// walk and scanline un-filtering (None / Sub / Up / Average / Paeth) are done here; the inflate is zlib's (`uncompress` from
// libz.so.1, resolved with dlopen the first time -- the library Python's own zlib module is linked against).  Alpha is dropped
// and grey is replicated, exactly what PIL's convert("RGB") does.  Anything else (palette, 16-bit, interlaced, a broken file)
// returns FF_PNG_UNSUPPORTED and the caller falls back to PIL.  Called through ctypes, i.e. without the GIL: PIL spends ~0.3 ms
// per small file holding it, which serialised the decode of the first batch of a folder.
// ------------------------------------------------------------------------------------------------
namespace {
typedef int (*UncompressFn)(unsigned char*, unsigned long*, const unsigned char*, unsigned long);
UncompressFn get_uncompress() {
  static UncompressFn fn = []() -> UncompressFn {
    void* h = dlopen("libz.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libz.so", RTLD_NOW | RTLD_GLOBAL);
    return h ? reinterpret_cast<UncompressFn>(dlsym(h, "uncompress")) : nullptr;
  }();
  return fn;
}
inline uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
}  // namespace

extern "C" int ff_png_decode_rgb8(const unsigned char* file, long long n, unsigned char* rgb, long long cap, int* out_h, int* out_w) {
  static const unsigned char sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
  if (!file || n < 8 + 25 || memcmp(file, sig, 8) != 0) return FF_PNG_UNSUPPORTED;
  long long pos = 8;
  if (be32(file + pos) != 13 || memcmp(file + pos + 4, "IHDR", 4) != 0) return FF_PNG_UNSUPPORTED;
  const uint32_t w = be32(file + pos + 8), h = be32(file + pos + 12);
  const int depth = file[pos + 16], ctype = file[pos + 17], comp = file[pos + 18], filt = file[pos + 19], lace = file[pos + 20];
  if (out_h) *out_h = (int)h;
  if (out_w) *out_w = (int)w;
  if (depth != 8 || comp != 0 || filt != 0 || lace != 0 || w == 0 || h == 0 || w > 65535 || h > 65535) return FF_PNG_UNSUPPORTED;
  const int ch = ctype == 0 ? 1 : ctype == 4 ? 2 : ctype == 2 ? 3 : ctype == 6 ? 4 : 0;
  if (!ch) return FF_PNG_UNSUPPORTED;
  if (!rgb) return FF_OK;      // header query
  if (cap < (long long)h * w * 3) return FF_ERR_ARG;
  UncompressFn unz = get_uncompress();
  if (!unz) return FF_PNG_UNSUPPORTED;
  pos += 12 + 13;
  // gather the IDAT payload (one chunk in most files: used in place)
  const unsigned char* z = nullptr;
  unsigned long zlen = 0;
  std::vector<unsigned char> joined;
  bool ended = false;
  while (pos + 12 <= n) {
    const uint32_t len = be32(file + pos);
    const unsigned char* type = file + pos + 4;
    if (pos + 12 + (long long)len > n) return FF_PNG_UNSUPPORTED;
    if (memcmp(type, "IDAT", 4) == 0) {
      if (!z) { z = file + pos + 8; zlen = len; }
      else {
        if (joined.empty()) joined.assign(z, z + zlen);
        joined.insert(joined.end(), file + pos + 8, file + pos + 8 + len);
      }
    } else if (memcmp(type, "IEND", 4) == 0) { ended = true; break; }
    pos += 12 + len;
  }
  if (!z || !ended) return FF_PNG_UNSUPPORTED;
  if (!joined.empty()) { z = joined.data(); zlen = joined.size(); }
  const size_t stride = (size_t)w * ch;
  std::vector<unsigned char> raw((stride + 1) * h);
  unsigned long rawlen = raw.size();
  if (unz(raw.data(), &rawlen, z, zlen) != 0 || rawlen != raw.size()) return FF_PNG_UNSUPPORTED;
  // un-filter in place (row r at raw[r * (stride + 1) + 1 ..]), then write RGB
  for (uint32_t r = 0; r < h; ++r) {
    unsigned char* cur = raw.data() + (size_t)r * (stride + 1) + 1;
    const unsigned char* up = r ? cur - (stride + 1) : nullptr;
    switch (cur[-1]) {
      case 0: break;
      case 1: for (size_t i = ch; i < stride; ++i) cur[i] = (unsigned char)(cur[i] + cur[i - ch]); break;
      case 2: if (up) for (size_t i = 0; i < stride; ++i) cur[i] = (unsigned char)(cur[i] + up[i]); break;
      case 3:
        for (size_t i = 0; i < stride; ++i) {
          const int a = i >= (size_t)ch ? cur[i - ch] : 0, b = up ? up[i] : 0;
          cur[i] = (unsigned char)(cur[i] + ((a + b) >> 1));
        }
        break;
      case 4:
        for (size_t i = 0; i < stride; ++i) {
          const int a = i >= (size_t)ch ? cur[i - ch] : 0, b = up ? up[i] : 0, c = (up && i >= (size_t)ch) ? up[i - ch] : 0;
          cur[i] = (unsigned char)(cur[i] + paeth(a, b, c));
        }
        break;
      default: return FF_PNG_UNSUPPORTED;
    }
    unsigned char* o = rgb + (size_t)r * w * 3;
    if (ch == 3) memcpy(o, cur, stride);
    else if (ch == 4) for (uint32_t x = 0; x < w; ++x) { o[3 * x] = cur[4 * x]; o[3 * x + 1] = cur[4 * x + 1]; o[3 * x + 2] = cur[4 * x + 2]; }
    else for (uint32_t x = 0; x < w; ++x) { const unsigned char g = cur[(size_t)x * ch]; o[3 * x] = o[3 * x + 1] = o[3 * x + 2] = g; }
  }
  return FF_OK;
}
