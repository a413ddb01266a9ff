// Image preprocessing on the GPU: uint8 HWC (decoded image, any size) -> Resize(shorter side -> R, bilinear with
// antialiasing) -> CenterCrop(S) -> ToTensor (/255) -> Normalize(mean, std) -> fp32 NCHW, one fused pass.
//
// Replaces the per-sample CPU transform of the reference's Dataset
// (example_scripts/Multimodal_example_task2C.txt:37-41: transforms.Resize(256), CenterCrop(224), ToTensor(),
// Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))).  The filter is the separable
// triangle ("bilinear", antialias=True) of torch / torchvision-v2 tensors: support = max(scale, 1), weights
// normalised per output pixel; PIL's own resize additionally rounds the horizontally-resized intermediate to uint8,
// which this kernel does not reproduce (difference < 1/255 per channel).
#include "common.cuh"

namespace b200 {

struct PreprocParams {
  const uint8_t* const* images;  // device array of n pointers to HWC uint8 images
  const int* heights;
  const int* widths;
  int n, resize, crop;
  float mean[3], inv_std[3];
  float* out;                    // [n, 3, crop, crop] fp32
};

__device__ __forceinline__ void aa_window(int o, float scale, int in_size, int& start, int& size, float& center,
                                          float& invscale) {
  const float support = scale >= 1.f ? scale : 1.f;
  invscale = scale >= 1.f ? 1.f / scale : 1.f;
  center = scale * (o + 0.5f);
  start = max(0, static_cast<int>(center - support + 0.5f));
  size = min(in_size, static_cast<int>(center + support + 0.5f)) - start;
}

__global__ void __launch_bounds__(256)
preprocess_kernel(const PreprocParams p) {
  const int img = blockIdx.z;
  const int ox = blockIdx.x * 32 + (threadIdx.x & 31);
  const int oy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (ox >= p.crop || oy >= p.crop) return;
  const int H = p.heights[img], W = p.widths[img];
  const uint8_t* src = p.images[img];
  // torchvision Resize(int): shorter side -> resize, longer side -> int(resize * long / short)
  int new_h, new_w;
  if (H <= W) { new_h = p.resize; new_w = static_cast<int>(static_cast<long long>(p.resize) * W / H); }
  else        { new_w = p.resize; new_h = static_cast<int>(static_cast<long long>(p.resize) * H / W); }
  // CenterCrop: int(round((size - crop) / 2))  (round-half-to-even like Python's round)
  const int top = static_cast<int>(rintf((new_h - p.crop) * 0.5f));
  const int left = static_cast<int>(rintf((new_w - p.crop) * 0.5f));
  const float sy = static_cast<float>(H) / new_h, sx = static_cast<float>(W) / new_w;
  int y0, ny, x0, nx;
  float cy, cx, iy, ix;
  aa_window(oy + top, sy, H, y0, ny, cy, iy);
  aa_window(ox + left, sx, W, x0, nx, cx, ix);
  float acc[3] = {0.f, 0.f, 0.f};
  float wsum_y = 0.f;
  for (int j = 0; j < ny; ++j) {
    const float wy = fmaxf(0.f, 1.f - fabsf((j + y0 - cy + 0.5f) * iy));
    wsum_y += wy;
    const uint8_t* row = src + (static_cast<long long>(y0 + j) * W + x0) * 3;
    float r[3] = {0.f, 0.f, 0.f};
    float wsum_x = 0.f;
    for (int i = 0; i < nx; ++i) {
      const float wx = fmaxf(0.f, 1.f - fabsf((i + x0 - cx + 0.5f) * ix));
      wsum_x += wx;
      r[0] = fmaf(wx, static_cast<float>(row[i * 3 + 0]), r[0]);
      r[1] = fmaf(wx, static_cast<float>(row[i * 3 + 1]), r[1]);
      r[2] = fmaf(wx, static_cast<float>(row[i * 3 + 2]), r[2]);
    }
    const float inv = wsum_x > 0.f ? wy / wsum_x : 0.f;
    acc[0] = fmaf(inv, r[0], acc[0]);
    acc[1] = fmaf(inv, r[1], acc[1]);
    acc[2] = fmaf(inv, r[2], acc[2]);
  }
  const float inv_y = wsum_y > 0.f ? 1.f / wsum_y : 0.f;
  const long long plane = static_cast<long long>(p.crop) * p.crop;
  float* o = p.out + static_cast<long long>(img) * 3 * plane + static_cast<long long>(oy) * p.crop + ox;
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c * plane] = (acc[c] * inv_y * (1.f / 255.f) - p.mean[c]) * p.inv_std[c];
}

}  // namespace b200

using namespace b200;

// images: device array of n device pointers (HWC uint8, 3 channels), heights/widths: device int arrays.
// out: fp32 [n, 3, crop, crop] = Normalize(ToTensor(CenterCrop(crop)(Resize(resize)(img)))).
B200MM_API int b200mm_preprocess_u8(const void* images, const int* heights, const int* widths, int n, int resize,
                                    int crop, const float* mean3, const float* std3, float* out, void* stream) {
  if (n <= 0 || resize <= 0 || crop <= 0 || crop > resize || !mean3 || !std3) return B200MM_ERR_BAD_ARG;
  PreprocParams p{};
  p.images = static_cast<const uint8_t* const*>(images);
  p.heights = heights;
  p.widths = widths;
  p.n = n; p.resize = resize; p.crop = crop;
  for (int c = 0; c < 3; ++c) {
    if (std3[c] == 0.f) return B200MM_ERR_BAD_ARG;
    p.mean[c] = mean3[c];
    p.inv_std[c] = 1.f / std3[c];
  }
  p.out = out;
  dim3 grid(ceil_div(crop, 32), ceil_div(crop, 8), n);
  preprocess_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
