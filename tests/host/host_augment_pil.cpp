// TEST INFRASTRUCTURE (never loaded by the product): csrc/augment_pil_math.cuh -- the arithmetic of augment_pil.cu's
// kernels -- compiled for the host, so that tests/test_cpu.py can demand byte equality with Pillow's ImageEnhance /
// HSV / rotate in a container without a GPU.   g++ -O2 -ffp-contract=off -shared -fPIC -I <pkg>/csrc tests/host/host_augment_pil.cpp
#include "augment_pil_math.cuh"

using namespace b200;

// in / out: [n][H][W][3] uint8; order [n]; alpha [n][3] float; hue [n] (0..255); affine [n][6] 16.16 fixed point
extern "C" void host_pil_augment(const unsigned char* in, int n, int H, int W, const int* order, const float* alpha,
                                 const int* hue, const int* affine, unsigned char* out) {
  const long long plane = static_cast<long long>(H) * W;
  for (int img = 0; img < n; ++img) {
    const unsigned char* src = in + img * plane * 3;
    pilaug::Jitter j;
    j.order = order[img];
    for (int k = 0; k < 3; ++k) j.alpha[k] = alpha[img * 3 + k];
    j.hue = hue[img];
    const int upto = pilaug::contrast_position(j.order);
    unsigned long long sum = 0;
    for (long long i = 0; i < plane; ++i) {
      int r = src[3 * i], g = src[3 * i + 1], b = src[3 * i + 2];
      pilaug::jitter_pixel(j, 0, upto, 0, r, g, b);
      sum += pilaug::luma(r, g, b);
    }
    const int gray = pilaug::contrast_gray(sum, plane);
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        int r = 0, g = 0, b = 0, xin, yin;
        if (pilaug::rotate_source(affine + img * 6, x, y, W, H, xin, yin)) {
          const unsigned char* p = src + (static_cast<long long>(yin) * W + xin) * 3;
          r = p[0]; g = p[1]; b = p[2];
          pilaug::jitter_pixel(j, 0, 4, gray, r, g, b);
        }
        unsigned char* o = out + (img * plane + static_cast<long long>(y) * W + x) * 3;
        o[0] = static_cast<unsigned char>(r); o[1] = static_cast<unsigned char>(g); o[2] = static_cast<unsigned char>(b);
      }
  }
}

// single operators, for exhaustive checks: mode 0 rgb->hsv, 1 hsv->rgb, 2 luma (into out[0]); in / out: [count][3]
extern "C" void host_pil_convert(const unsigned char* in, long long count, int mode, unsigned char* out) {
  for (long long i = 0; i < count; ++i) {
    int a = in[3 * i], b = in[3 * i + 1], c = in[3 * i + 2], x = 0, y = 0, z = 0;
    if (mode == 0) pilaug::rgb2hsv(a, b, c, x, y, z);
    else if (mode == 1) pilaug::hsv2rgb(a, b, c, x, y, z);
    else x = y = z = pilaug::luma(a, b, c);
    out[3 * i] = static_cast<unsigned char>(x); out[3 * i + 1] = static_cast<unsigned char>(y); out[3 * i + 2] = static_cast<unsigned char>(z);
  }
}
