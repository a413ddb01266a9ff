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
  const uint8_t* const* images;  // device array of n pointers to HWC uint8 images (nullptr: packed mode below)
  const uint8_t* packed;         // packed mode: ONE byte buffer holding every image of the batch ...
  const long long* offsets;      // ... image i starts at packed + offsets[i]
  const uint8_t* flip;           // per-image horizontal-flip flags (RandomHorizontalFlip) or nullptr
  const int* heights;
  const int* widths;
  int n, resize, crop;
  int square;                    // 0: Resize(resize) = shorter side -> resize, then CenterCrop(crop)  (.txt:37-41)
                                 // 1: Resize((crop, crop)), no crop  (HEAD script, Multimodal_example_task2C.py:224)
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
  const uint8_t* src = p.images ? p.images[img] : p.packed + p.offsets[img];
  int new_h, new_w, top = 0, left = 0;
  if (p.square) {
    new_h = new_w = p.crop;
  } else {
    // torchvision Resize(int): shorter side -> resize, longer side -> int(resize * long / short)
    if (H <= W) { new_h = p.resize; new_w = static_cast<int>(static_cast<long long>(p.resize) * W / H); }
    else        { new_w = p.resize; new_h = static_cast<int>(static_cast<long long>(p.resize) * H / W); }
    // CenterCrop: int(round((size - crop) / 2))  (round-half-to-even like Python's round)
    top = static_cast<int>(rintf((new_h - p.crop) * 0.5f));
    left = static_cast<int>(rintf((new_w - p.crop) * 0.5f));
  }
  const float sy = static_cast<float>(H) / new_h, sx = static_cast<float>(W) / new_w;
  int y0, ny, x0, nx;
  float cy, cx, iy, ix;
  aa_window(oy + top, sy, H, y0, ny, cy, iy);
  // RandomHorizontalFlip acts on the resized image: output column ox shows resized column new_w-1-(ox+left)
  const int rx = (p.flip && p.flip[img]) ? new_w - 1 - (ox + left) : ox + left;
  aa_window(rx, sx, W, x0, nx, cx, ix);
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

// Fixed-size batch [n, H, W, 3] uint8 (already at network resolution) -> ToTensor -> Normalize -> fp32 NCHW, optional
// per-image horizontal flip: what is left of the transform when the loader ships uint8 pixels (4x fewer PCIe bytes than
// the reference's fp32 tensors, .txt:61-69) -- one coalesced pass, a thread per 4 pixels (12 B in, 3 x 16 B out).
__global__ void __launch_bounds__(256)
u8_normalize_kernel(const uint8_t* __restrict__ src, const uint8_t* __restrict__ flip, int n, int H, int W,
                    float m0, float m1, float m2, float s0, float s1, float s2, float* __restrict__ out) {
  const int W4 = W >> 2;
  const long long total = static_cast<long long>(n) * H * W4;
  const long long plane = static_cast<long long>(H) * W;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x4 = static_cast<int>(i % W4);
    const long long r = i / W4;               // img * H + y
    const int img = static_cast<int>(r / H);
    const bool fl = flip != nullptr && flip[img] != 0;
    const int xs = fl ? W - 4 - 4 * x4 : 4 * x4;          // first source pixel of the 4-pixel group
    const uint32_t* q = reinterpret_cast<const uint32_t*>(src + (r * W + xs) * 3);   // 12 bytes, 4-byte aligned (W % 4 == 0)
    const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);
    // bytes: R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
    float R[4] = {static_cast<float>(w0 & 255u), static_cast<float>(w0 >> 24), static_cast<float>((w1 >> 16) & 255u),
                  static_cast<float>((w2 >> 8) & 255u)};
    float G[4] = {static_cast<float>((w0 >> 8) & 255u), static_cast<float>(w1 & 255u), static_cast<float>(w1 >> 24),
                  static_cast<float>((w2 >> 16) & 255u)};
    float B[4] = {static_cast<float>((w0 >> 16) & 255u), static_cast<float>((w1 >> 8) & 255u), static_cast<float>(w2 & 255u),
                  static_cast<float>(w2 >> 24)};
    if (fl) {
      float t;
      t = R[0]; R[0] = R[3]; R[3] = t; t = R[1]; R[1] = R[2]; R[2] = t;
      t = G[0]; G[0] = G[3]; G[3] = t; t = G[1]; G[1] = G[2]; G[2] = t;
      t = B[0]; B[0] = B[3]; B[3] = t; t = B[1]; B[1] = B[2]; B[2] = t;
    }
    float* o = out + static_cast<long long>(img) * 3 * plane + (r - static_cast<long long>(img) * H) * W + 4 * x4;
    *reinterpret_cast<float4*>(o) = make_float4(fmaf(R[0], s0, m0), fmaf(R[1], s0, m0), fmaf(R[2], s0, m0), fmaf(R[3], s0, m0));
    *reinterpret_cast<float4*>(o + plane) = make_float4(fmaf(G[0], s1, m1), fmaf(G[1], s1, m1), fmaf(G[2], s1, m1), fmaf(G[3], s1, m1));
    *reinterpret_cast<float4*>(o + 2 * plane) = make_float4(fmaf(B[0], s2, m2), fmaf(B[1], s2, m2), fmaf(B[2], s2, m2), fmaf(B[3], s2, m2));
  }
}

}  // namespace b200

using namespace b200;

static int fill_preproc(PreprocParams& p, int n, int resize, int crop, int square, const float* mean3, const float* std3) {
  if (n <= 0 || crop <= 0 || !mean3 || !std3) return B200MM_ERR_BAD_ARG;
  if (!square && (resize <= 0 || crop > resize)) return B200MM_ERR_BAD_ARG;
  p.n = n; p.resize = resize; p.crop = crop; p.square = square;
  for (int c = 0; c < 3; ++c) {
    if (std3[c] == 0.f) return B200MM_ERR_BAD_ARG;
    p.mean[c] = mean3[c];
    p.inv_std[c] = 1.f / std3[c];
  }
  return B200MM_OK;
}

// Packed variant: the whole batch of decoded images sits in ONE device byte buffer (one H2D copy from pinned memory),
// offsets / heights / widths are device arrays of n entries (one more small copy).  square = 1: Resize((crop, crop)) as
// the HEAD script does; flip: per-image RandomHorizontalFlip flags or nullptr.
B200MM_API int b200mm_preprocess_u8_packed(const void* packed, const long long* offsets, const int* heights,
                                           const int* widths, const void* flip, int n, int resize, int crop, int square,
                                           const float* mean3, const float* std3, float* out, void* stream) {
  PreprocParams p{};
  const int rc = fill_preproc(p, n, resize, crop, square, mean3, std3);
  if (rc) return rc;
  if (!packed || !offsets || !heights || !widths || !out) return B200MM_ERR_BAD_ARG;
  p.packed = static_cast<const uint8_t*>(packed);
  p.offsets = offsets;
  p.flip = static_cast<const uint8_t*>(flip);
  p.heights = heights;
  p.widths = widths;
  p.out = out;
  dim3 grid(ceil_div(crop, 32), ceil_div(crop, 8), n);
  preprocess_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

// src [n, H, W, 3] uint8 (W % 4 == 0) -> out [n, 3, H, W] fp32 = (src / 255 - mean) / std, optional per-image flip.
B200MM_API int b200mm_u8_normalize_nchw(const void* src, const void* flip, int n, int H, int W, const float* mean3,
                                        const float* std3, float* out, void* stream) {
  if (n <= 0 || H <= 0 || W <= 0 || (W & 3) || !src || !out || !mean3 || !std3) return B200MM_ERR_BAD_ARG;
  float s[3], m[3];
  for (int c = 0; c < 3; ++c) {
    if (std3[c] == 0.f) return B200MM_ERR_BAD_ARG;
    s[c] = 1.f / (255.f * std3[c]);
    m[c] = -mean3[c] / std3[c];
  }
  const long long total = static_cast<long long>(n) * H * (W >> 2);
  const int blocks = static_cast<int>(total / 256 + 1 < 148 * 16 ? total / 256 + 1 : 148 * 16);
  u8_normalize_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(src), static_cast<const uint8_t*>(flip), n, H, W, m[0], m[1], m[2], s[0], s[1], s[2], out);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

// images: device array of n device pointers (HWC uint8, 3 channels), heights/widths: device int arrays.
// out: fp32 [n, 3, crop, crop] = Normalize(ToTensor(CenterCrop(crop)(Resize(resize)(img)))).
B200MM_API int b200mm_preprocess_u8(const void* images, const int* heights, const int* widths, int n, int resize,
                                    int crop, const float* mean3, const float* std3, float* out, void* stream) {
  PreprocParams p{};
  const int rc = fill_preproc(p, n, resize, crop, 0, mean3, std3);
  if (rc) return rc;
  p.images = static_cast<const uint8_t* const*>(images);
  p.heights = heights;
  p.widths = widths;
  p.out = out;
  dim3 grid(ceil_div(crop, 32), ceil_div(crop, 8), n);
  preprocess_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
