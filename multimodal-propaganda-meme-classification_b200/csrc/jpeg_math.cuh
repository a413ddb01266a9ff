// Integer arithmetic of JPEG reconstruction -- dequantisation + inverse DCT, chroma up-sampling, YCbCr -> RGB -- stated
// so that the result is BIT-IDENTICAL to what the reference's loader produces: `Image.open(path).convert("RGB")`
// (example_scripts/Multimodal_example_task2C.txt:50; Multimodal_example_task2C.py:270), i.e. Pillow on libjpeg-turbo with
// its defaults (dct_method = JDCT_ISLOW, do_fancy_upsampling = TRUE, JCS_YCbCr -> JCS_RGB).  libjpeg-turbo is an
// un-vendored dependency of the reference's dependency (Pillow, poetry.lock); what is restated here is its published
// algorithm: the "slow-but-accurate" 13-bit fixed-point Loeffler-Ligtenberg-Moschytz IDCT (jidctint.c), the triangle-
// filter "fancy" up-sampling of h2v1 / h2v2 chroma (jdsample.c) and the 16-bit fixed-point colour conversion (jdcolor.c).
//
// Host-compilable: tests/host/host_jpeg.cpp builds these functions for the CPU and tests/test_cpu.py demands equality
// with Pillow's decode, pixel for pixel, over baseline / progressive files of every supported sampling.  The product
// calls them from jpeg_decode.cu's kernels only.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B200_JHD __host__ __device__ __forceinline__
#else
#define B200_JHD inline
#endif

namespace b200 {
namespace jpeg {

// post-IDCT range limit (jdmaster.c prepare_range_limit_table, indexed with (x & 1023) from the table's centre):
// x + 128 clamped to [0, 255] for every in-range x, and libjpeg's wrap-around for corrupt out-of-range values.
B200_JHD int idct_range_limit(int x) {
  const int i = x & 1023;
  return i < 128 ? i + 128 : i < 512 ? 255 : i < 896 ? 0 : i - 896;
}

B200_JHD int clamp255(int x) { return x < 0 ? 0 : x > 255 ? 255 : x; }

constexpr int kConstBits = 13, kPass1Bits = 2;
constexpr int F_0_298631336 = 2446, F_0_390180644 = 3196, F_0_541196100 = 4433, F_0_765366865 = 6270,
              F_0_899976223 = 7373, F_1_175875602 = 9633, F_1_501321110 = 12299, F_1_847759065 = 15137,
              F_1_961570560 = 16069, F_2_053119869 = 16819, F_2_562915447 = 20995, F_3_072711026 = 25172;

// One 1-D pass of jpeg_idct_islow over eight values; the two passes differ in the final shift only.
// (libjpeg's all-zero-AC column shortcut yields exactly what this computes for such a column, so it is not needed.)
B200_JHD void idct_1d(const int* in, int* out, int shift) {
  int z2 = in[2], z3 = in[6];
  int z1 = (z2 + z3) * F_0_541196100;
  int tmp2 = z1 + z3 * (-F_1_847759065);
  int tmp3 = z1 + z2 * F_0_765366865;
  z2 = in[0];
  z3 = in[4];
  int tmp0 = (z2 + z3) * (1 << kConstBits);
  int tmp1 = (z2 - z3) * (1 << kConstBits);
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  tmp0 = in[7];
  tmp1 = in[5];
  tmp2 = in[3];
  tmp3 = in[1];
  z1 = tmp0 + tmp3;
  z2 = tmp1 + tmp2;
  z3 = tmp0 + tmp2;
  int z4 = tmp1 + tmp3;
  const int z5 = (z3 + z4) * F_1_175875602;
  tmp0 *= F_0_298631336;
  tmp1 *= F_2_053119869;
  tmp2 *= F_3_072711026;
  tmp3 *= F_1_501321110;
  z1 *= -F_0_899976223;
  z2 *= -F_2_562915447;
  z3 *= -F_1_961570560;
  z4 *= -F_0_390180644;
  z3 += z5;
  z4 += z5;
  tmp0 += z1 + z3;
  tmp1 += z2 + z4;
  tmp2 += z2 + z3;
  tmp3 += z1 + z4;
  const int rnd = 1 << (shift - 1);
  out[0] = (tmp10 + tmp3 + rnd) >> shift;
  out[7] = (tmp10 - tmp3 + rnd) >> shift;
  out[1] = (tmp11 + tmp2 + rnd) >> shift;
  out[6] = (tmp11 - tmp2 + rnd) >> shift;
  out[2] = (tmp12 + tmp1 + rnd) >> shift;
  out[5] = (tmp12 - tmp1 + rnd) >> shift;
  out[3] = (tmp13 + tmp0 + rnd) >> shift;
  out[4] = (tmp13 - tmp0 + rnd) >> shift;
}

// coef: 64 quantised coefficients in natural (row-major) order; q: the component's quantisation table, natural order.
// out: 8 rows of 8 samples, `stride` bytes apart.
B200_JHD void idct_islow_block(const int16_t* coef, const uint16_t* q, uint8_t* out, int stride) {
  int ws[64];
#pragma unroll
  for (int c = 0; c < 8; ++c) {            // pass 1: columns
    int in[8], o[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) in[r] = static_cast<int>(coef[r * 8 + c]) * static_cast<int>(q[r * 8 + c]);
    idct_1d(in, o, kConstBits - kPass1Bits);
#pragma unroll
    for (int r = 0; r < 8; ++r) ws[r * 8 + c] = o[r];
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {            // pass 2: rows
    int o[8];
    idct_1d(ws + r * 8, o, kConstBits + kPass1Bits + 3);
#pragma unroll
    for (int c = 0; c < 8; ++c) out[r * stride + c] = static_cast<uint8_t>(idct_range_limit(o[c]));
  }
}

// jdcolor.c ycc_rgb_convert: SCALEBITS = 16, FIX(1.40200) = 91881, FIX(1.77200) = 116130, FIX(0.71414) = 46802,
// FIX(0.34414) = 22554, ONE_HALF = 32768 (arithmetic right shifts).
B200_JHD void ycc_to_rgb(int y, int cb, int cr, uint8_t* rgb) {
  const int xb = cb - 128, xr = cr - 128;
  rgb[0] = static_cast<uint8_t>(clamp255(y + ((91881 * xr + 32768) >> 16)));
  rgb[1] = static_cast<uint8_t>(clamp255(y + ((-22554 * xb + 32768 - 46802 * xr) >> 16)));
  rgb[2] = static_cast<uint8_t>(clamp255(y + ((116130 * xb + 32768) >> 16)));
}

// Chroma sample of OUTPUT pixel (x, y) from a component plane sub-sampled by (hs, vs) in {1, 2}: jdsample.c's
// fullsize / h2v1_fancy / h2v2_fancy up-samplers (h1v2 is not supported).  plane: rows `stride` bytes apart; cw, ch: the
// component's REAL size (ceil of the image size over the sampling ratio) -- the row above the first and below the last real
// row are those rows themselves (jdmainct.c's context rows), columns beyond cw never enter.  A component of width <= 2
// takes libjpeg's plain replication instead (jinit_upsampler: fancy needs downsampled_width > 2).
B200_JHD int upsampled_sample(const uint8_t* plane, int stride, int cw, int ch, int hs, int vs, int x, int y) {
  if (hs == 1 && vs == 1) return plane[y * stride + x];
  const int c = x >> 1, last = cw - 1;
  if (cw <= 2) return plane[(vs == 2 ? (y >> 1) : y) * stride + c];
  if (vs == 1) {                                   // h2v1
    const uint8_t* row = plane + y * stride;
    if ((x & 1) == 0) return c == 0 ? row[0] : (3 * row[c] + row[c - 1] + 1) >> 2;
    return c == last ? row[c] : (3 * row[c] + row[c + 1] + 2) >> 2;
  }
  const int r = y >> 1;                            // h2v2: the nearer row weighs 3, the farther 1, then the same across
  int r1 = (y & 1) ? r + 1 : r - 1;
  r1 = r1 < 0 ? 0 : r1 > ch - 1 ? ch - 1 : r1;
  const uint8_t* row0 = plane + r * stride;
  const uint8_t* row1 = plane + r1 * stride;
  const int cur = 3 * row0[c] + row1[c];
  if ((x & 1) == 0) {
    if (c == 0) return (cur * 4 + 8) >> 4;
    return (cur * 3 + (3 * row0[c - 1] + row1[c - 1]) + 8) >> 4;
  }
  if (c == last) return (cur * 4 + 7) >> 4;
  return (cur * 3 + (3 * row0[c + 1] + row1[c + 1]) + 7) >> 4;
}

}  // namespace jpeg
}  // namespace b200
