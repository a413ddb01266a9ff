// JPEG decode split the way a GPU wants it (the decode in `MultimodalDataset.__getitem__`,
// example_scripts/Multimodal_example_task2C.txt:50 / Multimodal_example_task2C.py:270 `Image.open(path).convert("RGB")`):
//
//   host   b200mm_jpeg_parse / b200mm_jpeg_entropy_decode     marker parsing + Huffman decoding (baseline and progressive,
//          restart intervals) into quantised DCT coefficients.  Entropy decoding is a bit-serial chain per scan -- it stays on
//          the host (DataLoader workers; the calls release the GIL), like the "hybrid" back end of vendor decoders.
//   device b200mm_jpeg_reconstruct                            everything after it, for a whole batch in two launches:
//          jpeg_idct_kernel            one thread per 8x8 block: dequantise + 13-bit fixed-point inverse DCT -> component planes
//          jpeg_upsample_color_kernel  one thread per output pixel: triangle-filter chroma up-sampling + YCbCr -> RGB ->
//                                      packed uint8 HWC images, the layout b200mm_preprocess_u8_packed reads
//
// Decoded pixels never cross PCIe and never exist on the host: the batch ships as coefficients.  The integer arithmetic
// (csrc/jpeg_math.cuh) restates libjpeg-turbo's default decode path, so the pixels equal Pillow's bit for bit
// (tests/test_cpu.py on the host build of the same header, tests/test_kernels_gpu.py on the kernels).
//
// Supported: 8-bit Huffman-coded baseline / extended-sequential / progressive files (SOF0 / SOF1 / SOF2), grayscale or
// YCbCr, chroma sub-sampling 4:4:4, 4:2:2 (h2v1), 4:2:0 (h2v2).  Everything else (arithmetic coding, 12-bit, CMYK, RGB-coded
// components, 4:4:0 and exotic sampling factors) returns B200MM_JPEG_UNSUPPORTED and is left to the caller's loader.
#include <cstring>
#include <vector>
#if defined(__SSE2__) && !defined(__CUDA_ARCH__)
#include <emmintrin.h>
#define B200_JPEG_SSE2 1
#endif

#include "common.cuh"
#include "jpeg_math.cuh"

enum : int { B200MM_JPEG_UNSUPPORTED = -10, B200MM_JPEG_CORRUPT = -11 };

namespace {

constexpr int kInfoInts = 32;
// info[]: 0 width, 1 height, 2 components, 3 progressive, 4 hs, 5 vs (chroma sub-sampling ratios, 1 or 2),
//         6..8 blocks per row (padded to whole MCUs) per component, 9..11 block rows (padded), 12..14 real component width,
//         15..17 real component height, 18..20 first coefficient of the component (int16 elements, image-relative),
//         21 coefficients of the image in total, 22 restart interval

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// Bit i of the result: coefficient i (natural order) of the block is non-zero.
inline uint64_t nonzero_mask_natural(const int16_t* blk) {
  uint64_t m = 0;
#ifdef B200_JPEG_SSE2
  const __m128i zero = _mm_setzero_si128();
  for (int i = 0; i < 4; ++i) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(blk + 16 * i));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(blk + 16 * i + 8));
    const __m128i z = _mm_packs_epi16(_mm_cmpeq_epi16(a, zero), _mm_cmpeq_epi16(b, zero));   // 0xFF where the value is 0
    m |= static_cast<uint64_t>(static_cast<uint16_t>(~_mm_movemask_epi8(z))) << (16 * i);
  }
#else
  for (int i = 0; i < 64; ++i) m |= static_cast<uint64_t>(blk[i] != 0) << i;
#endif
  return m;
}

// Natural-order bit mask -> zigzag-order bit mask, one table look-up per byte of the mask.
struct ZigzagMaskTable {
  uint64_t t[8][256];
  ZigzagMaskTable() {
    uint8_t pos_of_natural[64];
    for (int k = 0; k < 64; ++k) pos_of_natural[kZigzag[k]] = static_cast<uint8_t>(k);
    for (int b = 0; b < 8; ++b)
      for (int v = 0; v < 256; ++v) {
        uint64_t m = 0;
        for (int i = 0; i < 8; ++i)
          if (v & (1 << i)) m |= 1ULL << pos_of_natural[8 * b + i];
        t[b][v] = m;
      }
  }
  uint64_t zigzag(uint64_t natural) const {
    uint64_t m = 0;
    for (int b = 0; b < 8; ++b) m |= t[b][(natural >> (8 * b)) & 0xFF];
    return m;
  }
};

inline const ZigzagMaskTable& zigzag_mask_table() {
  static const ZigzagMaskTable table;
  return table;
}

struct HuffTable {
  bool present = false;
  uint8_t bits[17] = {0};
  uint8_t vals[256] = {0};
  // canonical decoding tables (ITU T.81 F.2.2.3) + a 9-bit look-ahead: look[code prefix] = (length << 8) | symbol
  int32_t maxcode[18];
  int32_t valoffset[17];
  uint16_t look[512];
  // AC tables only: when a code AND the value bits behind it fit into the 9-bit look-ahead, the whole coefficient comes out
  // of one lookup: fast[prefix] = (value << 8) | (run << 4) | total bits   (0 = take the two-step path)
  int16_t fast[512];
  void build_fast() {
    for (int i = 0; i < 512; ++i) {
      fast[i] = 0;
      const uint16_t e = look[i];
      if (!e) continue;
      const int l = e >> 8, run = (e >> 4) & 15, s = e & 15;
      if (s == 0 || l + s > 9) continue;
      const int v = ((i << l) & 511) >> (9 - s);
      const int val = v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
      if (val < -128 || val > 127) continue;
      fast[i] = static_cast<int16_t>(val * 256 + run * 16 + (l + s));
    }
  }
  bool build() {
    int32_t code = 0, p = 0;
    std::memset(look, 0, sizeof(look));
    for (int l = 1; l <= 16; ++l) {
      valoffset[l] = p - code;
      for (int i = 0; i < bits[l]; ++i, ++p, ++code) {
        if (p >= 256 || code >= (1 << l)) return false;       // more codes of this length than the code space holds
        if (l <= 9) {
          const int first = code << (9 - l);
          for (int k = 0; k < (1 << (9 - l)); ++k) look[first + k] = static_cast<uint16_t>((l << 8) | vals[p]);
        }
      }
      if (code > (1 << l)) return false;
      maxcode[l] = bits[l] ? code - 1 : -1;
      code <<= 1;
    }
    maxcode[17] = 0x7fffffff;
    return true;
  }
};

struct Component {
  int id = 0, h = 1, v = 1, tq = 0;
  int td = 0, ta = 0;          // tables selected by the current scan
  int wb = 0, hb = 0;          // blocks per row / block rows, padded to whole MCUs
  int rwb = 0, rhb = 0;        // blocks covering the real component size (what a non-interleaved scan codes)
  int cw = 0, ch = 0;          // real size in samples
  int pred = 0;
  long long coef0 = 0;         // first coefficient (int16 elements)
};

struct BitReader {
  const uint8_t* p;
  const uint8_t* end;
  uint64_t acc = 0;      // the next bits of the stream, left-aligned
  int nbits = 0;         // valid bits in acc
  int pad_bits = 0;      // zero bits appended after the segment's end (a marker or the end of the file) since the last reset
  bool hit_marker = false;
  BitReader(const uint8_t* b, const uint8_t* e) : p(b), end(e) {}
  void fill() {
    // fast path: eight bytes without an 0xFF among them -- take as many whole bytes as fit
    if (!hit_marker && end - p >= 8) {
      uint64_t v;
      std::memcpy(&v, p, 8);
      const uint64_t inv = ~v;
      if ((((inv - 0x0101010101010101ULL) & ~inv) & 0x8080808080808080ULL) == 0) {     // no byte of v is 0xFF
        const int k = (64 - nbits) >> 3;
        if (k > 0) {
          v = __builtin_bswap64(v);
          acc |= (k == 8 ? v : (v >> (64 - 8 * k)) << (64 - 8 * k - nbits));
          nbits += 8 * k;
          p += k;
        }
        return;
      }
    }
    while (nbits <= 56) {
      uint32_t byte = 0;
      if (!hit_marker && p < end) {
        byte = *p;
        if (byte == 0xFF) {
          if (p + 1 < end && p[1] == 0x00) {
            p += 2;
          } else {
            hit_marker = true;       // a marker ends the entropy-coded segment: zeros from here on (libjpeg does the same)
            byte = 0;
            pad_bits += 8;
          }
        } else {
          ++p;
        }
      } else {
        hit_marker = true;
        pad_bits += 8;
      }
      acc |= static_cast<uint64_t>(byte) << (56 - nbits);
      nbits += 8;
    }
  }
  // true when bits that were not in the file have been consumed: the segment ended early (truncated / corrupt file)
  bool overran() const { return pad_bits > nbits; }
  inline void need(int n) {
    if (nbits < n) fill();
  }
  inline uint32_t peek(int n) {
    if (nbits < n) fill();
    return static_cast<uint32_t>(acc >> (64 - n));
  }
  inline void skip(int n) {
    acc <<= n;
    nbits -= n;
  }
  inline uint32_t get(int n) {
    if (n == 0) return 0;
    const uint32_t v = peek(n);
    skip(n);
    return v;
  }
  inline int bit() { return static_cast<int>(get(1)); }
  // byte-align and consume an RSTn marker if one is next; returns false when the stream does not continue with one
  bool restart() {
    if (overran()) return false;
    acc = 0;
    nbits = 0;
    pad_bits = 0;
    hit_marker = false;
    while (p + 1 < end) {
      if (p[0] == 0xFF && p[1] >= 0xD0 && p[1] <= 0xD7) {
        p += 2;
        return true;
      }
      if (p[0] == 0xFF && p[1] != 0x00 && p[1] != 0xFF) return false;   // some other marker
      ++p;                                                                  // garbage / fill bytes before the marker
    }
    return false;
  }
};

inline int extend(uint32_t v, int s) { return v < (1u << (s - 1)) ? static_cast<int>(v) - (1 << s) + 1 : static_cast<int>(v); }

// Decodes one Huffman symbol from bits that are ALREADY in the accumulator (the caller made sure of >= 16).
inline int decode_symbol_nofill(BitReader& br, const HuffTable& t) {
  const uint32_t pre = static_cast<uint32_t>(br.acc >> 48);
  const uint16_t e = t.look[pre >> 7];
  if (e) {
    br.skip(e >> 8);
    return e & 0xFF;
  }
  int32_t code = static_cast<int32_t>(pre >> 6);   // 10 bits
  int l = 10;
  while (code > t.maxcode[l]) {
    if (++l > 16) return -1;                       // no code of any length matches: corrupt data
    code = static_cast<int32_t>(pre >> (16 - l));
  }
  br.skip(l);
  const int idx = code + t.valoffset[l];
  return idx >= 0 && idx < 256 ? t.vals[idx] : -1;
}

inline int decode_symbol(BitReader& br, const HuffTable& t) {
  br.need(16);
  return decode_symbol_nofill(br, t);
}

struct Decoder {
  const uint8_t* data;
  long long len;
  int width = 0, height = 0, ncomp = 0, progressive = 0, hmax = 1, vmax = 1, mcux = 0, mcuy = 0;
  int restart_interval = 0;
  bool have_sof = false, jfif = false;
  int adobe_transform = -1;
  Component comp[3];
  uint16_t qt[4][64];
  bool qt_ok[4] = {false, false, false, false};
  HuffTable dc[4], ac[4];
  long long total_coefs = 0;

  static int u16(const uint8_t* p) { return (p[0] << 8) | p[1]; }

  int parse_sof(const uint8_t* s, int n, int marker) {
    if (n < 6) return B200MM_JPEG_CORRUPT;
    if (s[0] != 8) return B200MM_JPEG_UNSUPPORTED;
    height = u16(s + 1);
    width = u16(s + 3);
    ncomp = s[5];
    progressive = marker == 0xC2;
    if (width <= 0 || height <= 0) return B200MM_JPEG_UNSUPPORTED;     // height 0 = DNL-defined: not supported
    if (ncomp != 1 && ncomp != 3) return B200MM_JPEG_UNSUPPORTED;
    if (n < 6 + 3 * ncomp) return B200MM_JPEG_CORRUPT;
    for (int i = 0; i < ncomp; ++i) {
      comp[i].id = s[6 + 3 * i];
      comp[i].h = s[7 + 3 * i] >> 4;
      comp[i].v = s[7 + 3 * i] & 15;
      comp[i].tq = s[8 + 3 * i];
      if (comp[i].tq > 3) return B200MM_JPEG_CORRUPT;
    }
    if (ncomp == 1) {
      comp[0].h = comp[0].v = 1;           // a single component is never sub-sampled whatever the file says
    } else {
      if (comp[1].h != 1 || comp[1].v != 1 || comp[2].h != 1 || comp[2].v != 1) return B200MM_JPEG_UNSUPPORTED;
      if (comp[0].h < 1 || comp[0].h > 2 || comp[0].v < 1 || comp[0].v > 2) return B200MM_JPEG_UNSUPPORTED;
      if (comp[0].h == 1 && comp[0].v == 2) return B200MM_JPEG_UNSUPPORTED;    // 4:4:0 (h1v2 up-sampling)
      if (comp[0].id == 'R' && comp[1].id == 'G' && comp[2].id == 'B') return B200MM_JPEG_UNSUPPORTED;
    }
    hmax = comp[0].h;
    vmax = comp[0].v;
    mcux = (width + 8 * hmax - 1) / (8 * hmax);
    mcuy = (height + 8 * vmax - 1) / (8 * vmax);
    total_coefs = 0;
    for (int i = 0; i < ncomp; ++i) {
      Component& c = comp[i];
      c.cw = (width * c.h + hmax - 1) / hmax;
      c.ch = (height * c.v + vmax - 1) / vmax;
      c.rwb = (c.cw + 7) / 8;
      c.rhb = (c.ch + 7) / 8;
      c.wb = mcux * c.h;
      c.hb = mcuy * c.v;
      c.coef0 = total_coefs;
      total_coefs += static_cast<long long>(c.wb) * c.hb * 64;
    }
    have_sof = true;
    return 0;
  }

  int parse_dqt(const uint8_t* s, int n) {
    while (n > 0) {
      const int pq = s[0] >> 4, tq = s[0] & 15;
      if (tq > 3 || pq > 1) return B200MM_JPEG_CORRUPT;
      const int need = 1 + 64 * (pq + 1);
      if (n < need) return B200MM_JPEG_CORRUPT;
      for (int k = 0; k < 64; ++k) qt[tq][kZigzag[k]] = pq ? u16(s + 1 + 2 * k) : s[1 + k];
      qt_ok[tq] = true;
      s += need;
      n -= need;
    }
    return 0;
  }

  int parse_dht(const uint8_t* s, int n) {
    while (n > 0) {
      if (n < 17) return B200MM_JPEG_CORRUPT;
      const int tc = s[0] >> 4, th = s[0] & 15;
      if (tc > 1 || th > 3) return B200MM_JPEG_CORRUPT;
      HuffTable& t = tc ? ac[th] : dc[th];
      int count = 0;
      t.bits[0] = 0;
      for (int l = 1; l <= 16; ++l) count += (t.bits[l] = s[l]);
      if (count > 256 || n < 17 + count) return B200MM_JPEG_CORRUPT;
      std::memset(t.vals, 0, sizeof(t.vals));
      std::memcpy(t.vals, s + 17, count);
      if (!t.build()) return B200MM_JPEG_CORRUPT;
      if (tc) t.build_fast();
      t.present = true;
      s += 17 + count;
      n -= 17 + count;
    }
    return 0;
  }

  // Walks the markers.  coefs == nullptr: stop after the frame header (sizes only).
  int run(int16_t* coefs) {
    if (len < 4 || data[0] != 0xFF || data[1] != 0xD8) return B200MM_JPEG_CORRUPT;
    const uint8_t* p = data + 2;
    const uint8_t* end = data + len;
    bool any_scan = false, seen_eoi = false;
    while (p + 4 <= end) {
      if (p[0] != 0xFF) { ++p; continue; }
      const int m = p[1];
      if (m == 0xFF) { ++p; continue; }                  // fill byte
      if (m == 0x00 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) { p += 2; continue; }
      if (m == 0xD9) { seen_eoi = true; break; }         // EOI
      const int n = u16(p + 2) - 2;
      const uint8_t* s = p + 4;
      if (n < 0 || s + n > end) return B200MM_JPEG_CORRUPT;
      int rc = 0;
      switch (m) {
        case 0xC0: case 0xC1: case 0xC2:
          if (have_sof) return B200MM_JPEG_CORRUPT;
          rc = parse_sof(s, n, m);
          if (rc == 0 && coefs == nullptr) return 0;
          break;
        case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
          return B200MM_JPEG_UNSUPPORTED;                // lossless, differential, arithmetic coding
        case 0xC4: rc = parse_dht(s, n); break;
        case 0xDB: rc = parse_dqt(s, n); break;
        case 0xDD:
          if (n < 2) return B200MM_JPEG_CORRUPT;
          restart_interval = u16(s);
          break;
        case 0xE0: if (n >= 5 && std::memcmp(s, "JFIF", 5) == 0) jfif = true; break;
        case 0xEE: if (n >= 12 && std::memcmp(s, "Adobe", 5) == 0) adobe_transform = s[11]; break;
        case 0xDA: {
          if (!have_sof) return B200MM_JPEG_CORRUPT;
          if (ncomp == 3 && !jfif && adobe_transform == 0) return B200MM_JPEG_UNSUPPORTED;   // RGB-coded components
          const uint8_t* next = nullptr;
          rc = scan(s, n, end, coefs, &next);
          if (rc) return rc;
          any_scan = true;
          p = next;
          continue;
        }
        default: break;                                  // APPn, COM, DNL ...: skipped
      }
      if (rc) return rc;
      p = s + n;
    }
    if (!have_sof) return B200MM_JPEG_CORRUPT;
    if (coefs != nullptr && !any_scan) return B200MM_JPEG_CORRUPT;
    if (coefs != nullptr && !seen_eoi && !(end - p >= 2 && p[0] == 0xFF && p[1] == 0xD9)) return B200MM_JPEG_CORRUPT;
    return 0;
  }

  int scan(const uint8_t* s, int n, const uint8_t* end, int16_t* coefs, const uint8_t** next) {
    if (n < 1) return B200MM_JPEG_CORRUPT;
    const int ns = s[0];
    if (ns < 1 || ns > ncomp || n < 1 + 2 * ns + 3) return B200MM_JPEG_CORRUPT;
    Component* sc[3];
    for (int i = 0; i < ns; ++i) {
      sc[i] = nullptr;
      for (int c = 0; c < ncomp; ++c)
        if (comp[c].id == s[1 + 2 * i]) sc[i] = &comp[c];
      if (!sc[i]) return B200MM_JPEG_CORRUPT;
      sc[i]->td = s[2 + 2 * i] >> 4;
      sc[i]->ta = s[2 + 2 * i] & 15;
      if (sc[i]->td > 3 || sc[i]->ta > 3) return B200MM_JPEG_CORRUPT;
    }
    const int Ss = s[1 + 2 * ns], Se = s[2 + 2 * ns], Ah = s[3 + 2 * ns] >> 4, Al = s[3 + 2 * ns] & 15;
    if (progressive) {
      if (Ss > Se || Se > 63 || (Ss == 0 && Se != 0) || (Ss > 0 && ns != 1) || Al > 13) return B200MM_JPEG_CORRUPT;
    } else if (Ss != 0 || Se != 63 || Ah != 0 || Al != 0) {
      return B200MM_JPEG_CORRUPT;
    }
    for (int i = 0; i < ns; ++i) {
      const bool need_dc = !progressive || (Ss == 0 && Ah == 0);
      const bool need_ac = !progressive || Ss > 0;
      if (need_dc && !dc[sc[i]->td].present) return B200MM_JPEG_CORRUPT;
      if (need_ac && !ac[sc[i]->ta].present) return B200MM_JPEG_CORRUPT;
      sc[i]->pred = 0;
    }
    BitReader br(s + n, end);
    int eobrun = 0;
    const bool interleaved = ns > 1;
    // a non-interleaved scan codes the blocks that cover the component's REAL size, one block per MCU
    const int units_x = interleaved ? mcux : sc[0]->rwb;
    const int units_y = interleaved ? mcuy : sc[0]->rhb;
    int to_restart = restart_interval;
    for (int uy = 0; uy < units_y; ++uy) {
      for (int ux = 0; ux < units_x; ++ux) {
        if (restart_interval && to_restart == 0) {
          if (!br.restart()) return B200MM_JPEG_CORRUPT;
          for (int i = 0; i < ns; ++i) sc[i]->pred = 0;
          eobrun = 0;
          to_restart = restart_interval;
        }
        for (int i = 0; i < ns; ++i) {
          Component& c = *sc[i];
          const int bh = interleaved ? c.h : 1, bv = interleaved ? c.v : 1;
          for (int v = 0; v < bv; ++v)
            for (int h = 0; h < bh; ++h) {
              const int bx = ux * bh + h, by = uy * bv + v;
              int16_t* blk = coefs + c.coef0 + (static_cast<long long>(by) * c.wb + bx) * 64;
              int rc;
              if (!progressive) rc = block_baseline(br, c, blk);
              else if (Ss == 0) rc = Ah == 0 ? block_dc_first(br, c, blk, Al) : block_dc_refine(br, blk, Al);
              else rc = Ah == 0 ? block_ac_first(br, c, blk, Ss, Se, Al, eobrun) : block_ac_refine(br, c, blk, Ss, Se, Al, eobrun);
              if (rc) return rc;
            }
        }
        --to_restart;
      }
    }
    if (br.overran()) return B200MM_JPEG_CORRUPT;      // the entropy-coded segment ended before its last block
    // the next marker: where the bit reader stopped, or the first marker after the bytes it had not reached yet
    const uint8_t* q = br.p;
    while (q + 1 < end && !(q[0] == 0xFF && q[1] != 0x00 && q[1] != 0xFF && !(q[1] >= 0xD0 && q[1] <= 0xD7))) ++q;
    *next = q;
    return 0;
  }

  int block_baseline(BitReader& br, Component& c, int16_t* blk) {
    br.need(32);                           // a symbol (<= 16 bits) and its value bits (<= 15) per refill check
    int s = decode_symbol_nofill(br, dc[c.td]);
    if (s < 0 || s > 15) return B200MM_JPEG_CORRUPT;
    if (s) {
      c.pred += extend(static_cast<uint32_t>(br.acc >> (64 - s)), s);
      br.skip(s);
    }
    blk[0] = static_cast<int16_t>(c.pred);
    const HuffTable& t = ac[c.ta];
    for (int k = 1; k < 64;) {
      br.need(32);
      const int f = t.fast[br.acc >> 55];
      if (f) {                             // run, size and value from one lookup
        k += (f >> 4) & 15;
        if (k > 63) return B200MM_JPEG_CORRUPT;
        blk[kZigzag[k++]] = static_cast<int16_t>(f >> 8);
        br.skip(f & 15);
        continue;
      }
      const int rs = decode_symbol_nofill(br, t);
      if (rs < 0) return B200MM_JPEG_CORRUPT;
      const int r = rs >> 4;
      s = rs & 15;
      if (s == 0) {
        if (r != 15) break;
        k += 16;
        continue;
      }
      k += r;
      if (k > 63) return B200MM_JPEG_CORRUPT;
      blk[kZigzag[k++]] = static_cast<int16_t>(extend(static_cast<uint32_t>(br.acc >> (64 - s)), s));
      br.skip(s);
    }
    return 0;
  }

  int block_dc_first(BitReader& br, Component& c, int16_t* blk, int Al) {
    const int s = decode_symbol(br, dc[c.td]);
    if (s < 0 || s > 15) return B200MM_JPEG_CORRUPT;
    if (s) c.pred += extend(br.get(s), s);
    blk[0] = static_cast<int16_t>(c.pred * (1 << Al));
    return 0;
  }

  int block_dc_refine(BitReader& br, int16_t* blk, int Al) {
    if (br.bit()) blk[0] = static_cast<int16_t>(blk[0] | (1 << Al));
    return 0;
  }

  int block_ac_first(BitReader& br, Component& c, int16_t* blk, int Ss, int Se, int Al, int& eobrun) {
    if (eobrun > 0) {
      --eobrun;
      return 0;
    }
    const HuffTable& t = ac[c.ta];
    for (int k = Ss; k <= Se; ++k) {
      br.need(32);
      const int f = t.fast[br.acc >> 55];
      if (f) {                             // run, size and value from one lookup
        k += (f >> 4) & 15;
        if (k > 63) return B200MM_JPEG_CORRUPT;
        blk[kZigzag[k]] = static_cast<int16_t>((f >> 8) * (1 << Al));
        br.skip(f & 15);
        continue;
      }
      const int rs = decode_symbol_nofill(br, t);
      if (rs < 0) return B200MM_JPEG_CORRUPT;
      const int r = rs >> 4, s = rs & 15;
      if (s) {
        k += r;
        if (k > 63) return B200MM_JPEG_CORRUPT;
        blk[kZigzag[k]] = static_cast<int16_t>(extend(br.get(s), s) * (1 << Al));
      } else if (r == 15) {
        k += 15;
      } else {
        eobrun = 1 << r;
        if (r) eobrun += static_cast<int>(br.get(r));
        --eobrun;
        break;
      }
    }
    return 0;
  }

  // One correction bit for every coefficient of `seg` (zigzag positions, ascending): |coefficient| grows by 1 << Al when
  // the bit is set and that bit of the magnitude is still clear (ITU T.81 G.1.2.3).
  static inline void refine_nonzeros(BitReader& br, int16_t* blk, uint64_t seg, int p1) {
    while (seg) {
      const int pos = __builtin_ctzll(seg);
      seg &= seg - 1;
      int16_t* coef = blk + kZigzag[pos];
      const int v = *coef;
      const int grow = br.bit() & static_cast<int>((v & p1) == 0);
      *coef = static_cast<int16_t>(v + (v >= 0 ? grow * p1 : -(grow * p1)));
    }
  }

  // Successive-approximation refinement of an AC band -- the hot loop of progressive files (four such scans per file as
  // Pillow writes them).  The reference formulation walks every position of the band and asks "non-zero?" (a branch that
  // mispredicts on textured images); here the block's non-zero positions are one 64-bit mask in zigzag order, a run of r
  // zeros is found with r + 1 count-trailing-zero steps and only the non-zero coefficients in between are visited.
  int block_ac_refine(BitReader& br, Component& c, int16_t* blk, int Ss, int Se, int Al, int& eobrun) {
    const int p1 = 1 << Al, m1 = -(1 << Al);
    const HuffTable& t = ac[c.ta];
    const uint64_t upto_se = Se == 63 ? ~0ULL : (1ULL << (Se + 1)) - 1;
    const uint64_t band = upto_se & ~((1ULL << Ss) - 1);
    const uint64_t nz = zigzag_mask_table().zigzag(nonzero_mask_natural(blk)) & band;   // as the block is on entry
    int k = Ss;
    if (eobrun == 0) {
      while (k <= Se) {
        const int rs = decode_symbol(br, t);
        if (rs < 0) return B200MM_JPEG_CORRUPT;
        int r = rs >> 4, s = rs & 15;
        if (s) {
          s = br.bit() ? p1 : m1;          // the new coefficient's sign; its magnitude is 1 << Al
        } else if (r != 15) {
          eobrun = 1 << r;
          if (r) eobrun += static_cast<int>(br.get(r));
          break;                           // end of band: the rest of this block is handled below
        }
        // the (r + 1)-th still-zero position at or after k (Se + 1 when the band has fewer): where the new coefficient
        // goes (or, for a zero run of 16, the last position skipped); coefficients placed earlier in this call lie below k
        const uint64_t from_k = ~((1ULL << k) - 1);
        uint64_t zeros = ~nz & band & from_k;
        int target = Se + 1;
        for (int i = 0; i <= r && zeros; ++i) {
          if (i == r) target = __builtin_ctzll(zeros);
          zeros &= zeros - 1;
        }
        const uint64_t below_target = target >= 64 ? ~0ULL : (1ULL << target) - 1;
        refine_nonzeros(br, blk, nz & from_k & below_target, p1);
        if (s) {
          if (target > 63) return B200MM_JPEG_CORRUPT;
          blk[kZigzag[target]] = static_cast<int16_t>(s);
        }
        k = target + 1;
      }
    }
    if (eobrun > 0) {
      if (k <= Se) refine_nonzeros(br, blk, nz & ~((1ULL << k) - 1), p1);
      --eobrun;
    }
    return 0;
  }

  void fill_info(int* info) const {
    std::memset(info, 0, kInfoInts * sizeof(int));
    info[0] = width;
    info[1] = height;
    info[2] = ncomp;
    info[3] = progressive;
    info[4] = hmax;
    info[5] = vmax;
    for (int i = 0; i < ncomp; ++i) {
      info[6 + i] = comp[i].wb;
      info[9 + i] = comp[i].hb;
      info[12 + i] = comp[i].cw;
      info[15 + i] = comp[i].ch;
      info[18 + i] = static_cast<int>(comp[i].coef0);
    }
    info[21] = static_cast<int>(total_coefs);
    info[22] = restart_interval;
  }
};

}  // namespace

// ---------------------------------------------------------------------------------------------------- host entry points
// Frame header only: fills info[32] (layout above) so that the caller can size the coefficient buffer (info[21] int16).
B200MM_API int b200mm_jpeg_parse(const void* data, long long len, int* info) {
  if (!data || len <= 0 || !info) return B200MM_ERR_BAD_ARG;
  Decoder d;
  d.data = static_cast<const uint8_t*>(data);
  d.len = len;
  const int rc = d.run(nullptr);
  if (rc) return rc;
  if (d.total_coefs > 0x7fffffffLL) return B200MM_JPEG_UNSUPPORTED;
  d.fill_info(info);
  return B200MM_OK;
}

// Entropy-decodes every scan of the file.  coefs: info[21] int16 (zeroed here), blocks of 64 coefficients in natural
// (row-major) order, component after component, block rows padded to whole MCUs.  qtabs: [3][64] uint16, each
// component's quantisation table in natural order.  Host memory only; no CUDA call is made.
B200MM_API int b200mm_jpeg_entropy_decode(const void* data, long long len, short* coefs, unsigned short* qtabs, int* info) {
  if (!data || len <= 0 || !coefs || !qtabs || !info) return B200MM_ERR_BAD_ARG;
  Decoder d;
  d.data = static_cast<const uint8_t*>(data);
  d.len = len;
  int rc = d.run(nullptr);                       // sizes first: the buffer must be zero before any scan writes into it
  if (rc) return rc;
  if (d.total_coefs > 0x7fffffffLL) return B200MM_JPEG_UNSUPPORTED;
  std::memset(coefs, 0, static_cast<size_t>(d.total_coefs) * sizeof(short));
  Decoder e;
  e.data = d.data;
  e.len = len;
  rc = e.run(coefs);
  if (rc) return rc;
  for (int i = 0; i < 3; ++i) {
    const int c = i < e.ncomp ? i : 0;
    if (!e.qt_ok[e.comp[c].tq]) return B200MM_JPEG_CORRUPT;
    std::memcpy(qtabs + 64 * i, e.qt[e.comp[c].tq], 64 * sizeof(uint16_t));
  }
  e.fill_info(info);
  return B200MM_OK;
}

// The same decode, delivered SPARSE: after the dense decode into `scratch` (info[21] int16, caller-owned and reusable
// across files) the non-zero coefficients are compacted block by block, in the dense layout's block order:
//   block_off [blocks + 1] int32   entry range of block b = [block_off[b], block_off[b + 1])
//   idx [nnz] uint8                position of the coefficient inside its block (natural order, 0..63)
//   val [nnz] int16                its value
// capacity: entries idx / val can hold (info[21] always suffices); *nnz receives the count.  A typical file keeps 5-15 % of
// its coefficients, so the batch that crosses PCIe shrinks accordingly (3 B per non-zero + 4 B per block instead of
// 128 B per block).  Returns B200MM_ERR_BAD_ARG when capacity is too small.
B200MM_API int b200mm_jpeg_entropy_decode_sparse(const void* data, long long len, short* scratch, int* block_off,
                                                 unsigned char* idx, short* val, long long capacity,
                                                 unsigned short* qtabs, int* info, long long* nnz) {
  if (!block_off || !idx || !val || !nnz || capacity < 0) return B200MM_ERR_BAD_ARG;
  const int rc = b200mm_jpeg_entropy_decode(data, len, scratch, qtabs, info);
  if (rc) return rc;
  const long long blocks = info[21] / 64;
  long long n = 0;
  for (long long b = 0; b < blocks; ++b) {
    block_off[b] = static_cast<int>(n);
    const short* blk = scratch + b * 64;
    for (int k = 0; k < 64; k += 4) {
      uint64_t four;
      std::memcpy(&four, blk + k, 8);
      if (four == 0) continue;                       // most groups of four are empty
      for (int j = 0; j < 4; ++j)
        if (blk[k + j] != 0) {
          if (n >= capacity) return B200MM_ERR_BAD_ARG;
          idx[n] = static_cast<unsigned char>(k + j);
          val[n] = blk[k + j];
          ++n;
        }
    }
  }
  block_off[blocks] = static_cast<int>(n);
  *nnz = n;
  return B200MM_OK;
}

// ---------------------------------------------------------------------------------------------------- device side
namespace b200 {

// Per-image row of the batch table (int64 [n][kJpegTableCols]):
//   0 width, 1 height, 2 components, 3 hs, 4 vs, 5..7 blocks per row per component, 8..10 block rows,
//   11..13 real component width, 14..16 real component height, 17..19 first coefficient of the component in the BATCH
//   buffer (int16 elements), 20..22 first byte of the component's plane in the plane scratch, 23 first byte of the image
//   in the packed RGB output, 24 blocks of the image in total
constexpr int kJpegTableCols = 32;

// SPARSE = false: coefs holds 64 int16 per block.  SPARSE = true: the batch ships only its non-zero coefficients
// (sp_off [blocks of the batch + 1], sp_idx, sp_val; table columns 17..19 = first BLOCK of the component in the batch's
// block numbering); a thread scatters its block's entries into its own row of a shared-memory tile (rows 33 words apart:
// the same position in 32 different rows falls into 32 different banks) and reads the row back into registers.
constexpr int kIdctThreads = 128;
constexpr int kSparseRowWords = 33;

template <bool SPARSE>
__global__ void __launch_bounds__(kIdctThreads)
jpeg_idct_kernel(const int16_t* __restrict__ coefs, const int* __restrict__ sp_off, const uint8_t* __restrict__ sp_idx,
                 const int16_t* __restrict__ sp_val, const uint16_t* __restrict__ qtabs,
                 const long long* __restrict__ table, uint8_t* __restrict__ planes) {
  __shared__ uint32_t tile[SPARSE ? kIdctThreads * kSparseRowWords : 1];
  const int img = blockIdx.y;
  const long long* t = table + static_cast<long long>(img) * kJpegTableCols;
  const int total = static_cast<int>(t[24]);
  const int ncomp = static_cast<int>(t[2]);
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < total; b += gridDim.x * blockDim.x) {
    int c = 0, idx = b;
    for (; c < ncomp - 1; ++c) {
      const int nb = static_cast<int>(t[5 + c] * t[8 + c]);
      if (idx < nb) break;
      idx -= nb;
    }
    const int wb = static_cast<int>(t[5 + c]);
    const int by = idx / wb, bx = idx - by * wb;
    int16_t blk[64];
    if (SPARSE) {
      // private to this thread: no synchronisation needed; volatile because the row is written as halves and read as words
      volatile uint32_t* row = tile + threadIdx.x * kSparseRowWords;
#pragma unroll
      for (int k = 0; k < 32; ++k) row[k] = 0u;
      volatile int16_t* row16 = reinterpret_cast<volatile int16_t*>(row);
      const long long gb = t[17 + c] + idx;
      const int e1 = sp_off[gb + 1];
      for (int e = sp_off[gb]; e < e1; ++e) row16[sp_idx[e] & 63] = sp_val[e];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const uint32_t w = row[k];
        blk[2 * k] = static_cast<int16_t>(w & 0xffffu);
        blk[2 * k + 1] = static_cast<int16_t>(w >> 16);
      }
    } else {
      // the block's 128 bytes as eight 16-byte loads; everything in registers
      const uint4* src = reinterpret_cast<const uint4*>(coefs + t[17 + c] + static_cast<long long>(idx) * 64);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint4 v = __ldg(src + k);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          blk[8 * k + 2 * j] = static_cast<int16_t>(w[j] & 0xffffu);
          blk[8 * k + 2 * j + 1] = static_cast<int16_t>(w[j] >> 16);
        }
      }
    }
    uint8_t px[64];
    jpeg::idct_islow_block(blk, qtabs + (static_cast<long long>(img) * 3 + c) * 64, px, 8);
    const int stride = wb * 8;
    uint8_t* dst = planes + t[20 + c] + static_cast<long long>(by) * 8 * stride + bx * 8;   // eight 8-byte stores
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      uint2 o;
      o.x = px[8 * r] | (px[8 * r + 1] << 8) | (px[8 * r + 2] << 16) | (static_cast<uint32_t>(px[8 * r + 3]) << 24);
      o.y = px[8 * r + 4] | (px[8 * r + 5] << 8) | (px[8 * r + 6] << 16) | (static_cast<uint32_t>(px[8 * r + 7]) << 24);
      *reinterpret_cast<uint2*>(dst + r * stride) = o;
    }
  }
}

__global__ void __launch_bounds__(256)
jpeg_upsample_color_kernel(const uint8_t* __restrict__ planes, const long long* __restrict__ table,
                           uint8_t* __restrict__ out) {
  const int img = blockIdx.z;
  const long long* t = table + static_cast<long long>(img) * kJpegTableCols;
  if (t[2] == 0) return;               // an image the caller decoded itself (unsupported file): its pixels are copied in
  const int W = static_cast<int>(t[0]), H = static_cast<int>(t[1]);
  const int stride0 = static_cast<int>(t[5]) * 8;
  const uint8_t* p0 = planes + t[20];
  uint8_t* o = out + t[23];
  if (t[2] == 1) {
    for (int y = blockIdx.y * 8 + (threadIdx.x >> 5); y < H; y += gridDim.y * 8)
      for (int x = blockIdx.x * 32 + (threadIdx.x & 31); x < W; x += gridDim.x * 32) {
        const uint8_t v = p0[y * stride0 + x];               // "L" -> convert("RGB"): the grey value three times
        uint8_t* px = o + (static_cast<long long>(y) * W + x) * 3;
        px[0] = v; px[1] = v; px[2] = v;
      }
    return;
  }
  const int hs = static_cast<int>(t[3]), vs = static_cast<int>(t[4]);
  const int stride1 = static_cast<int>(t[6]) * 8, stride2 = static_cast<int>(t[7]) * 8;
  const uint8_t* p1 = planes + t[21];
  const uint8_t* p2 = planes + t[22];
  const int cw1 = static_cast<int>(t[12]), ch1 = static_cast<int>(t[15]);
  const int cw2 = static_cast<int>(t[13]), ch2 = static_cast<int>(t[16]);
  for (int y = blockIdx.y * 8 + (threadIdx.x >> 5); y < H; y += gridDim.y * 8)
    for (int x = blockIdx.x * 32 + (threadIdx.x & 31); x < W; x += gridDim.x * 32) {
      const int yy = p0[y * stride0 + x];
      const int cb = jpeg::upsampled_sample(p1, stride1, cw1, ch1, hs, vs, x, y);
      const int cr = jpeg::upsampled_sample(p2, stride2, cw2, ch2, hs, vs, x, y);
      jpeg::ycc_to_rgb(yy, cb, cr, o + (static_cast<long long>(y) * W + x) * 3);
    }
}

}  // namespace b200

using namespace b200;

// Reconstructs a batch of entropy-decoded images on the device.  coefs: the batch's coefficients (int16), qtabs:
// [n][3][64] uint16, table: int64 [n][32] (layout above), planes: scratch for the component planes (sum over images and
// components of blocks * 64 bytes), out: packed uint8 RGB, image i = [height_i][width_i][3] at byte table[i][23].
// max_blocks / max_w / max_h: maxima over the batch (grid sizing).  All pointers are device pointers.
B200MM_API int b200mm_jpeg_reconstruct(const short* coefs, const unsigned short* qtabs, const long long* table, int n,
                                       int max_blocks, int max_w, int max_h, void* planes, void* out, void* stream) {
  if (!coefs || !qtabs || !table || !planes || !out || n <= 0 || n > 65535 || max_blocks <= 0 || max_w <= 0 || max_h <= 0)
    return B200MM_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(coefs) & 15) || (reinterpret_cast<uintptr_t>(planes) & 7)) return B200MM_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int gx = ceil_div(max_blocks, 128);
  if (gx > 4096) gx = 4096;
  jpeg_idct_kernel<false><<<dim3(gx, n), kIdctThreads, 0, st>>>(reinterpret_cast<const int16_t*>(coefs), nullptr, nullptr,
                                                                nullptr, qtabs, table, static_cast<uint8_t*>(planes));
  B200MM_CHECK_LAUNCH();
  int cx = ceil_div(max_w, 32), cy = ceil_div(max_h, 8);
  if (cx > 64) cx = 64;
  if (cy > 256) cy = 256;
  jpeg_upsample_color_kernel<<<dim3(cx, cy, n), 256, 0, st>>>(static_cast<const uint8_t*>(planes), table,
                                                              static_cast<uint8_t*>(out));
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

// The same reconstruction from the SPARSE batch (b200mm_jpeg_entropy_decode_sparse): sp_off int32 [blocks of the batch + 1]
// (entry offsets made batch-wide by the caller), sp_idx uint8 [nnz], sp_val int16 [nnz]; table columns 17..19 hold the
// first BLOCK of each component in the batch's block numbering instead of a coefficient offset.
B200MM_API int b200mm_jpeg_reconstruct_sparse(const int* sp_off, const unsigned char* sp_idx, const short* sp_val,
                                              const unsigned short* qtabs, const long long* table, int n, int max_blocks,
                                              int max_w, int max_h, void* planes, void* out, void* stream) {
  if (!sp_off || !sp_idx || !sp_val || !qtabs || !table || !planes || !out || n <= 0 || n > 65535 || max_blocks <= 0 ||
      max_w <= 0 || max_h <= 0)
    return B200MM_ERR_BAD_ARG;
  if (reinterpret_cast<uintptr_t>(planes) & 7) return B200MM_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int gx = ceil_div(max_blocks, kIdctThreads);
  if (gx > 4096) gx = 4096;
  jpeg_idct_kernel<true><<<dim3(gx, n), kIdctThreads, 0, st>>>(nullptr, sp_off, sp_idx, reinterpret_cast<const int16_t*>(sp_val),
                                                               qtabs, table, static_cast<uint8_t*>(planes));
  B200MM_CHECK_LAUNCH();
  int cx = ceil_div(max_w, 32), cy = ceil_div(max_h, 8);
  if (cx > 64) cx = 64;
  if (cy > 256) cy = 256;
  jpeg_upsample_color_kernel<<<dim3(cx, cy, n), 256, 0, st>>>(static_cast<const uint8_t*>(planes), table,
                                                              static_cast<uint8_t*>(out));
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
