// Pillow's 8-bit bilinear resize, restated (the reference's Dataset resizes the PIL image: transforms.Resize(256) /
// Resize((224, 224)) -> Image.resize(..., BILINEAR); example_scripts/Multimodal_example_task2C.txt:37-41,
// Multimodal_example_task2C.py:224).  Pillow is an un-vendored dependency of the reference (poetry.lock); what follows is
// its published algorithm (src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc,
// ImagingResampleHorizontal_8bpc / Vertical_8bpc):
//   * per output index: window [xmin, xmin + n) of the triangle filter of support max(scale, 1), weights in double,
//     normalised by their sum, then rounded to 22-bit fixed point;
//   * every output sample = clip8((2^21 + sum(sample * coefficient)) >> 22);
//   * two passes -- horizontal first, into a uint8 image, then vertical -- so the intermediate is ROUNDED TO uint8.
// That rounding is why the float-path kernel (preprocess.cu) differs from the reference's tensors by up to one uint8
// step; with these functions the resized image is bit-identical to Pillow's (tests/test_cpu.py, host build), and so is
// everything the reference's Dataset hands its loop.
//
// Host-compilable (tests/host/host_resample.cpp); the product calls them from preprocess_pil.cu's kernels only.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B200_RHD __host__ __device__ __forceinline__
#else
#define B200_RHD inline
#endif

namespace b200 {
namespace pil {

constexpr int kPrecisionBits = 32 - 8 - 2;
constexpr int kMaxTaps = 64;      // (int)ceil(support) * 2 + 1 <= 64  <=>  down-scaling by up to 31x

// IEEE double arithmetic, one rounding per operation (no contraction), as the C compiler evaluates Pillow's expressions.
B200_RHD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  volatile double r = a * b;
  return r;
#endif
}
B200_RHD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  volatile double r = a + b;
  return r;
#endif
}
B200_RHD double ddiv(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __ddiv_rn(a, b);
#else
  volatile double r = a / b;
  return r;
#endif
}

B200_RHD bool supported(int in_size, int out_size) {
  if (in_size <= 0 || out_size <= 0) return false;
  const double scale = static_cast<double>(in_size) / out_size;
  const double support = scale < 1.0 ? 1.0 : scale;
  int c = static_cast<int>(support);
  if (static_cast<double>(c) < support) ++c;                        // ceil
  return c * 2 + 1 <= kMaxTaps;
}

// precompute_coeffs + normalize_coeffs_8bpc for ONE output index xx: window start, length, fixed-point coefficients.
// (The weights are evaluated twice -- once for their sum, once to normalise -- instead of being kept in a double array:
// the same expressions give the same values, and a GPU thread keeps kMaxTaps ints instead of kMaxTaps doubles + ints.)
B200_RHD void coefficients(int xx, int in_size, int out_size, int& xmin, int& n, int* kk) {
  const double scale = ddiv(static_cast<double>(in_size), static_cast<double>(out_size));
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = dmul(1.0, filterscale);                     // bilinear: filter support 1.0
  const double center = dmul(dadd(static_cast<double>(xx), 0.5), scale);      // in0 = 0
  const double ss = ddiv(1.0, filterscale);
  xmin = static_cast<int>(dadd(dadd(center, -support), 0.5));
  if (xmin < 0) xmin = 0;
  int xmax = static_cast<int>(dadd(dadd(center, support), 0.5));
  if (xmax > in_size) xmax = in_size;
  n = xmax - xmin;
  if (n > kMaxTaps) n = kMaxTaps;      // memory safety only: callers refuse such sizes up front (supported())
  double ww = 0.0;
  for (int x = 0; x < n; ++x) {
    double a = dmul(dadd(dadd(static_cast<double>(x + xmin), -center), 0.5), ss);
    if (a < 0.0) a = -a;
    ww = dadd(ww, a < 1.0 ? dadd(1.0, -a) : 0.0);
  }
  for (int x = 0; x < n; ++x) {
    double a = dmul(dadd(dadd(static_cast<double>(x + xmin), -center), 0.5), ss);
    if (a < 0.0) a = -a;
    const double w = a < 1.0 ? dadd(1.0, -a) : 0.0;
    const double v = ww != 0.0 ? ddiv(w, ww) : w;
    kk[x] = v < 0 ? static_cast<int>(dadd(-0.5, dmul(v, static_cast<double>(1 << kPrecisionBits))))
                  : static_cast<int>(dadd(0.5, dmul(v, static_cast<double>(1 << kPrecisionBits))));
  }
}

B200_RHD uint8_t clip8(int v) {
  v >>= kPrecisionBits;                                              // arithmetic shift, as in C
  return static_cast<uint8_t>(v < 0 ? 0 : v > 255 ? 255 : v);
}

// One output sample of a pass: `src` points at the window's first sample, consecutive samples `stride` bytes apart.
B200_RHD uint8_t resample(const uint8_t* src, long long stride, int n, const int* kk) {
  int ss = 1 << (kPrecisionBits - 1);
  for (int x = 0; x < n; ++x) ss += static_cast<int>(src[x * stride]) * kk[x];
  return clip8(ss);
}

// torchvision's Resize(int) target size and CenterCrop offsets (functional.py: _compute_resized_output_size, center_crop)
B200_RHD void resized_size(int H, int W, int resize, int square, int crop, int& new_h, int& new_w) {
  if (square) {
    new_h = new_w = crop;
  } else if (H <= W) {
    new_h = resize;
    new_w = static_cast<int>(static_cast<long long>(resize) * W / H);
  } else {
    new_w = resize;
    new_h = static_cast<int>(static_cast<long long>(resize) * H / W);
  }
}

// One output pixel of the reference's PIL transform, as uint8: Resize(resize) -> CenterCrop(crop) (square = 0, .txt:37-38)
// or Resize((crop, crop)) (square = 1, .py:224), then RandomHorizontalFlip when `flip`.  src: [H][W][3] uint8.
// The horizontal pass's uint8 samples are recomputed inside the vertical window (ny x nx multiply-adds per channel, the
// same count as the float kernel's 2-D window) instead of materialising Pillow's intermediate image.
B200_RHD void preprocess_pixel_u8(const uint8_t* src, int H, int W, int resize, int crop, int square, int flip, int ox,
                                  int oy, uint8_t* rgb) {
  int new_h, new_w, top = 0, left = 0;
  resized_size(H, W, resize, square, crop, new_h, new_w);
  if (!square) {
    // center_crop: int(round((size - crop) / 2.0)) with Python's round-half-to-even
    const int dh = new_h - crop, dw = new_w - crop;
    top = (dh >> 1) + ((dh & 1) ? ((dh >> 1) & 1) : 0);
    left = (dw >> 1) + ((dw & 1) ? ((dw >> 1) & 1) : 0);
  }
  const int y = oy + top;
  const int x = flip ? new_w - 1 - (ox + left) : ox + left;
  int kx[kMaxTaps], ky[kMaxTaps];
  int xmin, nx, ymin, ny;
  coefficients(x, W, new_w, xmin, nx, kx);
  coefficients(y, H, new_h, ymin, ny, ky);
  int acc[3] = {1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1)};
  for (int j = 0; j < ny; ++j) {
    const uint8_t* row = src + (static_cast<long long>(ymin + j) * W + xmin) * 3;
    for (int c = 0; c < 3; ++c) acc[c] += static_cast<int>(resample(row + c, 3, nx, kx)) * ky[j];
  }
  for (int c = 0; c < 3; ++c) rgb[c] = clip8(acc[c]);
}

}  // namespace pil
}  // namespace b200
