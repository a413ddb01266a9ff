// TEST INFRASTRUCTURE (never loaded by the product): csrc/resample_math.cuh -- the arithmetic of preprocess_pil.cu's
// kernels -- compiled for the host, so that tests/test_cpu.py can demand equality with Pillow's Image.resize in a
// container without a GPU.    g++ -O2 -ffp-contract=off -shared -fPIC -I <pkg>/csrc tests/host/host_resample.cpp
#include <vector>

#include "resample_math.cuh"

using namespace b200;

// src [H][W][3] uint8 -> out [new_h][new_w][3] uint8: horizontal pass into a uint8 intermediate, then the vertical pass
extern "C" int host_pil_resize(const unsigned char* src, int H, int W, int new_h, int new_w, unsigned char* out) {
  if (!pil::supported(W, new_w) || !pil::supported(H, new_h)) return -1;
  std::vector<uint8_t> mid(static_cast<size_t>(H) * new_w * 3);
  int kk[pil::kMaxTaps];
  for (int x = 0; x < new_w; ++x) {
    int xmin, n;
    pil::coefficients(x, W, new_w, xmin, n, kk);
    for (int y = 0; y < H; ++y)
      for (int c = 0; c < 3; ++c)
        mid[(static_cast<size_t>(y) * new_w + x) * 3 + c] =
            pil::resample(src + (static_cast<size_t>(y) * W + xmin) * 3 + c, 3, n, kk);
  }
  for (int y = 0; y < new_h; ++y) {
    int ymin, n;
    pil::coefficients(y, H, new_h, ymin, n, kk);
    for (int x = 0; x < new_w; ++x)
      for (int c = 0; c < 3; ++c)
        out[(static_cast<size_t>(y) * new_w + x) * 3 + c] =
            pil::resample(mid.data() + (static_cast<size_t>(ymin) * new_w + x) * 3 + c, static_cast<long long>(new_w) * 3, n, kk);
  }
  return 0;
}

// the whole uint8 part of the transform, pixel by pixel as the kernel computes it: out [crop][crop][3] uint8
extern "C" int host_pil_preprocess(const unsigned char* src, int H, int W, int resize, int crop, int square, int flip,
                                   unsigned char* out) {
  int new_h, new_w;
  pil::resized_size(H, W, resize, square, crop, new_h, new_w);
  if (!pil::supported(W, new_w) || !pil::supported(H, new_h)) return -1;
  for (int oy = 0; oy < crop; ++oy)
    for (int ox = 0; ox < crop; ++ox)
      pil::preprocess_pixel_u8(src, H, W, resize, crop, square, flip, ox, oy, out + (static_cast<size_t>(oy) * crop + ox) * 3);
  return 0;
}
