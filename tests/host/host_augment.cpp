// TEST INFRASTRUCTURE (never loaded by the product): compiles csrc/augment_math.cuh -- the per-pixel arithmetic of
// augment.cu's kernels -- for the host, so that tests/test_cpu.py can check it against torchvision's tensor-path
// ColorJitter / rotate in a container without a GPU.  Built by the test itself with
//   g++ -O2 -ffp-contract=off -shared -fPIC -I <pkg>/csrc tests/host/host_augment.cpp
#include "augment_math.cuh"

using namespace b200;

extern "C" void host_augment_jitter_rotate(const float* img01, const int* order, const float* params, int n, int H,
                                           int W, const float* mean3, const float* std3, float* gray_mean,
                                           float* out) {
  const int plane = H * W;
  for (int img = 0; img < n; ++img) {
    const float* src = img01 + static_cast<long long>(img) * 3 * plane;
    const float* prm = params + img * 8;
    const aug::Jitter j = aug::make_jitter(order[img], prm);
    const int upto = aug::contrast_position(j.order);
    double acc = 0.0;
    for (int i = 0; i < plane; ++i) {
      float r = src[i], g = src[plane + i], b = src[2 * plane + i];
      aug::jitter_pixel(j, 0, upto, 0.f, r, g, b);
      acc += aug::gray(r, g, b);
    }
    gray_mean[img] = static_cast<float>(acc / plane);
    float* dst = out + static_cast<long long>(img) * 3 * plane;
    float t[4];
    for (int k = 0; k < 4; ++k) t[k] = aug::rotate_scale(prm + 4, k, W, H);
    for (int oy = 0; oy < H; ++oy)
      for (int ox = 0; ox < W; ++ox) {
        float r = 0.f, g = 0.f, b = 0.f;
        int sx, sy;
        if (aug::rotate_source(ox, oy, W, H, t, sx, sy)) {
          r = src[sy * W + sx];
          g = src[plane + sy * W + sx];
          b = src[2 * plane + sy * W + sx];
          aug::jitter_pixel(j, 0, 4, gray_mean[img], r, g, b);
        }
        dst[oy * W + ox] = aug::normalize(r, mean3[0], 1.f / std3[0]);
        dst[plane + oy * W + ox] = aug::normalize(g, mean3[1], 1.f / std3[1]);
        dst[2 * plane + oy * W + ox] = aug::normalize(b, mean3[2], 1.f / std3[2]);
      }
  }
}
