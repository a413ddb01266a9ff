// The participant script's train-time augmentations with PILLOW'S OWN ARITHMETIC: its Dataset applies ColorJitter and
// RandomRotation to the PIL image (example_scripts/Multimodal_example_task2C.py:224-233), i.e. torchvision's PIL back end:
//   adjust_brightness / contrast / saturation -> ImageEnhance.* = Image.blend(degenerate, image, factor)  (libImaging/Blend.c)
//   grey levels                              -> convert("L"): (R*19595 + G*38470 + B*7471 + 0x8000) >> 16 (Convert.c)
//   adjust_hue                               -> convert("HSV"), uint8 hue += uint8(factor * 255), convert("RGB")  (Convert.c)
//   rotate (NEAREST, expand=False, fill=0)   -> Image.rotate -> affine transform in 16.16 fixed point   (Geometry.c)
// every operator maps uint8 to uint8.  Pillow is an un-vendored dependency of the reference (poetry.lock); what is restated
// is its published algorithm, with each C expression's float / double evaluation reproduced operation by operation.  The
// float-tensor kernels (augment.cu) stay within Pillow's quantisation of this (mean 1.6 uint8 steps); these functions are
// byte-identical to Pillow (tests/test_cpu.py on the host build: every operator over millions of values, the composed
// pipeline against torchvision's PIL ColorJitter / rotate and against a run of the script's own Dataset).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define B200_PHD __host__ __device__ __forceinline__
#else
#define B200_PHD inline
#endif

namespace b200 {
namespace pilaug {

// one IEEE rounding per C operation
B200_PHD float fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b;
  return r;
#endif
}
B200_PHD float fadd(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b;
  return r;
#endif
}
B200_PHD float fdiv(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  volatile float r = a / b;
  return r;
#endif
}
B200_PHD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  volatile double r = a + b;
  return r;
#endif
}
B200_PHD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  volatile double r = a * b;
  return r;
#endif
}
B200_PHD double ddiv(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __ddiv_rn(a, b);
#else
  volatile double r = a / b;
  return r;
#endif
}
B200_PHD int clip8(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }

// convert("L")
B200_PHD int luma(int r, int g, int b) { return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16; }

// ImagingBlend(degenerate, image, alpha): (UINT8)((int)in1 + alpha * ((int)in2 - (int)in1)), clipped when extrapolating
B200_PHD int blend(int degenerate, int image, float alpha) {
  const float t = fadd(static_cast<float>(degenerate), fmul(alpha, static_cast<float>(image - degenerate)));
  if (alpha >= 0.f && alpha <= 1.0f) return static_cast<int>(t) & 255;
  return t <= 0.f ? 0 : t >= 255.f ? 255 : static_cast<int>(t);
}

// rgb2hsv_row: h, the quotients and s are C floats; "2.0 + rc - bc", "h / 6.0 + 1.0" and "* 255.0" evaluate in double
B200_PHD void rgb2hsv(int r, int g, int b, int& uh, int& us, int& uv) {
  const int maxc = r > g ? (r > b ? r : b) : (g > b ? g : b);
  const int minc = r < g ? (r < b ? r : b) : (g < b ? g : b);
  uv = maxc;
  if (minc == maxc) {
    uh = 0;
    us = 0;
    return;
  }
  const float cr = static_cast<float>(maxc - minc);
  const float s = fdiv(cr, static_cast<float>(maxc));
  const float rc = fdiv(static_cast<float>(maxc - r), cr);
  const float gc = fdiv(static_cast<float>(maxc - g), cr);
  const float bc = fdiv(static_cast<float>(maxc - b), cr);
  float h;
  if (r == maxc) h = fadd(bc, -gc);
  else if (g == maxc) h = static_cast<float>(dadd(dadd(2.0, static_cast<double>(rc)), -static_cast<double>(bc)));
  else h = static_cast<float>(dadd(dadd(4.0, static_cast<double>(gc)), -static_cast<double>(rc)));
  const double t = dadd(ddiv(static_cast<double>(h), 6.0), 1.0);          // in [5/6, 11/6]: fmod(t, 1.0) = t - floor(t), exact
  h = static_cast<float>(dadd(t, -floor(t)));
  uh = clip8(static_cast<int>(dmul(static_cast<double>(h), 255.0)));
  us = clip8(static_cast<int>(dmul(static_cast<double>(s), 255.0)));
}

// hsv2rgb_row
B200_PHD void hsv2rgb(int h, int s, int v, int& r, int& g, int& b) {
  if (s == 0) {
    r = g = b = v;
    return;
  }
  const float hf = fdiv(fmul(static_cast<float>(h), 6.0f), 255.0f);
  const float fl = floorf(hf);
  const int i = static_cast<int>(fl);
  const float f = fadd(hf, -fl);
  const float fs = fdiv(static_cast<float>(s), 255.0f);
  const float vf = static_cast<float>(v);
  const float pf = fmul(vf, fadd(1.0f, -fs));
  const float qf = fmul(vf, fadd(1.0f, -fmul(fs, f)));
  const float tf = fmul(vf, fadd(1.0f, -fmul(fs, fadd(1.0f, -f))));
  const int p = clip8(static_cast<int>(floor(dadd(static_cast<double>(pf), 0.5))));     // round(): values are >= 0
  const int q = clip8(static_cast<int>(floor(dadd(static_cast<double>(qf), 0.5))));
  const int t = clip8(static_cast<int>(floor(dadd(static_cast<double>(tf), 0.5))));
  switch (i % 6) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}

// Per-image draw: order = the permutation of the four operators (2 bits each, first applied in the low bits; 0 brightness,
// 1 contrast, 2 saturation, 3 hue); alpha[3] = the three enhancement factors as C floats; hue = uint8(hue_factor * 255)
struct Jitter {
  int order;
  float alpha[3];
  int hue;
};

B200_PHD int contrast_position(int order) {
  for (int k = 0; k < 4; ++k)
    if (((order >> (2 * k)) & 3) == 1) return k;
  return 4;
}

// operators [first, last) of the permutation on one pixel; gray = ImageEnhance.Contrast's degenerate level,
// int(mean(convert("L")) + 0.5) of the image as it is when the contrast operator runs
B200_PHD void jitter_pixel(const Jitter& j, int first, int last, int gray, int& r, int& g, int& b) {
  for (int k = first; k < last; ++k) {
    const int op = (j.order >> (2 * k)) & 3;
    if (op == 0) {
      r = blend(0, r, j.alpha[0]);
      g = blend(0, g, j.alpha[0]);
      b = blend(0, b, j.alpha[0]);
    } else if (op == 1) {
      r = blend(gray, r, j.alpha[1]);
      g = blend(gray, g, j.alpha[1]);
      b = blend(gray, b, j.alpha[1]);
    } else if (op == 2) {
      const int l = luma(r, g, b);
      r = blend(l, r, j.alpha[2]);
      g = blend(l, g, j.alpha[2]);
      b = blend(l, b, j.alpha[2]);
    } else {
      int h, s, v;
      rgb2hsv(r, g, b, h, s, v);
      h = (h + j.hue) & 255;                       // uint8 addition wraps, "as desired" (torchvision)
      hsv2rgb(h, s, v, r, g, b);
    }
  }
}

// ImageStat mean of the grey image, rounded as ImageEnhance.Contrast does: int(sum / count + 0.5) in double
B200_PHD int contrast_gray(unsigned long long luma_sum, long long count) {
  return static_cast<int>(dadd(ddiv(static_cast<double>(luma_sum), static_cast<double>(count)), 0.5));
}

// affine_fixed (Geometry.c): source pixel of output pixel (x, y) under the 16.16 fixed-point matrix a[6] =
// {a0, a1, a2 + half-pixel offsets, a3, a4, a5 + offsets} prepared by the caller exactly as Image.rotate / affine_fixed do.
B200_PHD bool rotate_source(const int* a, int x, int y, int W, int H, int& xin, int& yin) {
  const long long xx = static_cast<long long>(a[2]) + static_cast<long long>(x) * a[0] + static_cast<long long>(y) * a[1];
  const long long yy = static_cast<long long>(a[5]) + static_cast<long long>(x) * a[3] + static_cast<long long>(y) * a[4];
  xin = static_cast<int>(xx >> 16);
  yin = static_cast<int>(yy >> 16);
  return xin >= 0 && xin < W && yin >= 0 && yin < H;
}

}  // namespace pilaug
}  // namespace b200
