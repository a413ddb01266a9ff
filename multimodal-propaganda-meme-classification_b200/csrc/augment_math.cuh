// Per-pixel arithmetic of the HEAD script's train-time augmentations
// (example_scripts/Multimodal_example_task2C.py:224-233: ColorJitter(brightness=.1, contrast=.1, saturation=.1, hue=.1)
// and RandomRotation(degrees=15) between Resize((224, 224)) / RandomHorizontalFlip and ToTensor / Normalize).
//
// Semantics = torchvision's float-tensor path ($SP/torchvision/transforms/_functional_tensor.py: _blend :258-261,
// rgb_to_grayscale :148-168, _rgb2hsv :264-300, _hsv2rgb :303-321, adjust_hue :199-221; rotate = _gen_affine_grid
// :579-602 + grid_sample(nearest, zeros, align_corners=False)), one rounding per torch op (no fused multiply-adds), so
// that the kernel and torchvision agree to the last bit wherever the arithmetic is pointwise (Normalize's final
// division is a multiplication by 1 / std: one ulp).  The reference runs the
// same operators on PIL images, whose uint8 intermediates add up to 1/255 of rounding per operator.
//
// Host-compilable (plain C++ when __CUDACC__ is absent): tests/host/host_augment.cpp builds these functions for the CPU and
// checks them against torchvision where no GPU is present.  The product only ever calls them from augment.cu's kernels.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#else
#define B200_HD inline
#endif

namespace b200 {
namespace aug {

// One rounding per operation, as a chain of separate torch kernels produces.
B200_HD float mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b;
  return r;
#endif
}
B200_HD float add_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b;
  return r;
#endif
}
B200_HD float sub_rn(float a, float b) { return add_rn(a, -b); }
B200_HD float div_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  volatile float r = a / b;
  return r;
#endif
}
B200_HD float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

// rgb_to_grayscale: (0.2989 r + 0.587 g) + 0.114 b
B200_HD float gray(float r, float g, float b) {
  return add_rn(add_rn(mul_rn(0.2989f, r), mul_rn(0.587f, g)), mul_rn(0.114f, b));
}

// _blend(img, other, ratio) = clamp(ratio * img + (1 - ratio) * other, 0, 1); one_minus = fp32(1.0 - ratio)
B200_HD float blend(float v, float other, float ratio, float one_minus) {
  return clamp01(add_rn(mul_rn(ratio, v), mul_rn(one_minus, other)));
}

// torch.fmod(x, 1.0) for x >= 0 and torch.remainder(x, 1.0) for -1 < x: both equal x - floor(x) BIT FOR BIT in these
// ranges (the subtraction is exact for x >= 0; for -1 < x < 0 it is the one rounding of x + 1 that remainder performs),
// so the iterative fmodf is never needed.
B200_HD float frac1(float x) { return sub_rn(x, floorf(x)); }

// adjust_hue: rgb -> hsv, h = (h + shift) mod 1, hsv -> rgb.  _rgb2hsv evaluates three quotients (maxc - c) / cr and
// multiplies two sums by zero; only the two quotients of the branch that survives are computed here -- same values.
B200_HD void hue_shift(float& r, float& g, float& b, float shift) {
  const float maxc = fmaxf(fmaxf(r, g), b);
  const float minc = fminf(fminf(r, g), b);
  const bool eqc = maxc == minc;
  const float cr = sub_rn(maxc, minc);
  const float s = div_rn(cr, eqc ? 1.f : maxc);
  const float cd = eqc ? 1.f : cr;
  // maxc == r: h = bc - gc;  maxc == g (and != r): h = (2 + rc) - bc;  otherwise: h = (4 + gc) - rc
  const bool is_r = maxc == r, is_g = maxc == g;
  const float x1 = is_r ? b : is_g ? r : g;
  const float x2 = is_r ? g : is_g ? b : r;
  const float base = is_r ? 0.f : is_g ? 2.f : 4.f;
  const float q1 = div_rn(sub_rn(maxc, x1), cd);
  const float q2 = div_rn(sub_rn(maxc, x2), cd);
  float h = sub_rn(add_rn(base, q1), q2);
  h = frac1(add_rn(div_rn(h, 6.f), 1.f));          // torch.fmod(h / 6 + 1, 1)
  h = frac1(add_rn(h, shift));                     // (h + hue_factor) % 1.0
  const float v = maxc;
  const float h6 = mul_rn(h, 6.f);
  const float fl = floorf(h6);
  const float f = sub_rn(h6, fl);
  const float p = clamp01(mul_rn(v, sub_rn(1.f, s)));
  const float q = clamp01(mul_rn(v, sub_rn(1.f, mul_rn(s, f))));
  const float t = clamp01(mul_rn(v, sub_rn(1.f, mul_rn(s, sub_rn(1.f, f)))));
  const int i = static_cast<int>(fl) % 6;
  switch (i) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}

// Per-image ColorJitter draw (ColorJitter.get_params, $SP/torchvision/transforms/transforms.py:1237-1266):
//   order : the permutation of the four operators, 2 bits each -- operator applied k-th = (order >> 2k) & 3
//           (0 brightness, 1 contrast, 2 saturation, 3 hue, torchvision's fn_idx numbering)
//   f[4]  : brightness / contrast / saturation factors and the hue shift
struct Jitter {
  int order;
  float f[4];
  float one_minus[3];   // fp32(1.0 - factor) of the three blends
};

B200_HD Jitter make_jitter(int order, const float* f4) {
  Jitter j;
  j.order = order;
  for (int k = 0; k < 4; ++k) j.f[k] = f4[k];
  // Python computes 1.0 - ratio in double and torch rounds it to fp32: the fp32 subtraction rounds the same exact value
  for (int k = 0; k < 3; ++k) j.one_minus[k] = sub_rn(1.f, f4[k]);
  return j;
}

// Position of the contrast operator in the permutation (the operators before it shape the image whose grey mean it needs).
B200_HD int contrast_position(int order) {
  for (int k = 0; k < 4; ++k)
    if (((order >> (2 * k)) & 3) == 1) return k;
  return 4;
}

// Applies operators [first, last) of the permutation to one pixel; `mean` = grey mean of the whole image as it is when
// the contrast operator runs (adjust_contrast: torch.mean(rgb_to_grayscale(img))).
B200_HD void jitter_pixel(const Jitter& j, int first, int last, float mean, float& r, float& g, float& b) {
  for (int k = first; k < last; ++k) {
    const int op = (j.order >> (2 * k)) & 3;
    if (op == 0) {
      // other = zeros: ratio * v + (1 - ratio) * 0
      r = blend(r, 0.f, j.f[0], j.one_minus[0]);
      g = blend(g, 0.f, j.f[0], j.one_minus[0]);
      b = blend(b, 0.f, j.f[0], j.one_minus[0]);
    } else if (op == 1) {
      r = blend(r, mean, j.f[1], j.one_minus[1]);
      g = blend(g, mean, j.f[1], j.one_minus[1]);
      b = blend(b, mean, j.f[1], j.one_minus[1]);
    } else if (op == 2) {
      const float l = gray(r, g, b);
      r = blend(r, l, j.f[2], j.one_minus[2]);
      g = blend(g, l, j.f[2], j.one_minus[2]);
      b = blend(b, l, j.f[2], j.one_minus[2]);
    } else {
      hue_shift(r, g, b, j.f[3]);
    }
  }
}

// RandomRotation(NEAREST, expand=False, fill=0).  m = the first two columns of the inverse affine matrix
// [[m00, m01], [m10, m11]] (translation is zero for a rotation about the centre); _gen_affine_grid divides its rows by half
// the width / height: t[k], the same for every pixel of an image (computed once per CTA, not per thread).
B200_HD float rotate_scale(const float* m, int k, int W, int H) {
  return div_rn(m[k], 0.5f * static_cast<float>(k < 2 ? W : H));
}

// Source pixel of output pixel (ox, oy): grid product, grid_sample's un-normalisation and nearbyint, each in fp32.
// Returns false outside the image (zero fill).
B200_HD bool rotate_source(int ox, int oy, int W, int H, const float* t, int& sx, int& sy) {
  const float bx = static_cast<float>(ox) + (0.5f - 0.5f * static_cast<float>(W));
  const float by = static_cast<float>(oy) + (0.5f - 0.5f * static_cast<float>(H));
  const float gx = add_rn(mul_rn(bx, t[0]), mul_rn(by, t[1]));
  const float gy = add_rn(mul_rn(bx, t[2]), mul_rn(by, t[3]));
  const float ix = mul_rn(sub_rn(mul_rn(add_rn(gx, 1.f), static_cast<float>(W)), 1.f), 0.5f);
  const float iy = mul_rn(sub_rn(mul_rn(add_rn(gy, 1.f), static_cast<float>(H)), 1.f), 0.5f);
  const float rx = nearbyintf(ix), ry = nearbyintf(iy);
  if (!(rx >= 0.f && rx <= static_cast<float>(W - 1) && ry >= 0.f && ry <= static_cast<float>(H - 1))) return false;
  sx = static_cast<int>(rx);
  sy = static_cast<int>(ry);
  return true;
}

// Normalize: (v - mean) * (1 / std) -- within one ulp of torch's (v - mean) / std.
B200_HD float normalize(float v, float mean, float inv_std) { return mul_rn(sub_rn(v, mean), inv_std); }

}  // namespace aug
}  // namespace b200
