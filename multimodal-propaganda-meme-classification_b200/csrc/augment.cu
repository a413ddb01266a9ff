// Train-time augmentations of the HEAD script on the GPU: ColorJitter + RandomRotation + Normalize over a batch that
// is already resized to network resolution (example_scripts/Multimodal_example_task2C.py:224-233).
//
//   img01 [n, 3, H, W] fp32 in [0, 1]  (output of preprocess_kernel / u8_normalize_kernel with mean 0, std 1: Resize +
//                                       RandomHorizontalFlip + ToTensor already applied)
//   -> jitter_gray_mean_kernel            eight CTAs per image: grey sums of the image as the contrast operator sees it
//                                         (after the operators that precede it in this image's random permutation)
//   -> jitter_rotate_normalize_kernel     one thread per OUTPUT pixel: inverse-rotate to the source pixel (nearest, zero
//                                         fill), apply the four colour operators in the drawn order, Normalize, store
//
// The colour operators are pointwise once the grey mean is known, so "jitter the image, then rotate it" equals "gather
// the rotated source pixel, then jitter it": the jittered image is never materialised.  Both kernels are HBM-bound:
// 12 B read per pixel by the first, 12 B gathered (rotation by <= 15 degrees keeps a warp's 32 sources within two or
// three rows) + 12 B written by the second.  The arithmetic is csrc/augment_math.cuh (torchvision's tensor path).
#include "common.cuh"
#include "augment_math.cuh"

namespace b200 {

constexpr int kMeanThreads = 256;
constexpr int kMeanParts = 8;     // CTAs per image: 8 x 256 images = 2048 CTAs of 256 threads, ~14 resident waves of loads per SM

// partial[img][part] = sum of the grey values of this CTA's slice of the image (after the operators that precede
// contrast).  Fixed slices and a fixed reduction tree: the mean is deterministic, run to run.
__global__ void __launch_bounds__(kMeanThreads)
jitter_gray_mean_kernel(const float* __restrict__ img01, const int* __restrict__ order, const float* __restrict__ params,
                        int H, int W, int vec, float* __restrict__ partial) {
  const int img = blockIdx.y, part = blockIdx.x;
  const aug::Jitter j = aug::make_jitter(order[img], params + img * 8);
  const int upto = aug::contrast_position(j.order);
  const int plane = H * W;
  const float* src = img01 + static_cast<long long>(img) * 3 * plane;
  float acc = 0.f;
  if (vec) {
    // 16-byte loads (plane % 4 == 0 and a 16-byte aligned base: every plane of every image is aligned), three planes per step: 48 B in flight per thread and iteration
    const int q = plane >> 2;
    const int lo = static_cast<int>(static_cast<long long>(q) * part / kMeanParts);
    const int hi = static_cast<int>(static_cast<long long>(q) * (part + 1) / kMeanParts);
    const float4* r4 = reinterpret_cast<const float4*>(src);
    const float4* g4 = reinterpret_cast<const float4*>(src + plane);
    const float4* b4 = reinterpret_cast<const float4*>(src + 2 * plane);
#pragma unroll 2
    for (int i = lo + threadIdx.x; i < hi; i += kMeanThreads) {
      const float4 r = __ldg(r4 + i), g = __ldg(g4 + i), b = __ldg(b4 + i);
      float rr[4] = {r.x, r.y, r.z, r.w}, gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        aug::jitter_pixel(j, 0, upto, 0.f, rr[k], gg[k], bb[k]);
        acc += aug::gray(rr[k], gg[k], bb[k]);
      }
    }
  } else {
    const int lo = static_cast<int>(static_cast<long long>(plane) * part / kMeanParts);
    const int hi = static_cast<int>(static_cast<long long>(plane) * (part + 1) / kMeanParts);
    for (int i = lo + threadIdx.x; i < hi; i += kMeanThreads) {
      float r = __ldg(src + i), g = __ldg(src + plane + i), b = __ldg(src + 2 * plane + i);
      aug::jitter_pixel(j, 0, upto, 0.f, r, g, b);
      acc += aug::gray(r, g, b);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float warp_sum[kMeanThreads / 32];
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < kMeanThreads / 32 ? warp_sum[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) partial[img * kMeanParts + part] = v;
  }
}

__global__ void __launch_bounds__(256)
jitter_rotate_normalize_kernel(const float* __restrict__ img01, const int* __restrict__ order,
                               const float* __restrict__ params, const float* __restrict__ partial, int H, int W,
                               float m0, float m1, float m2, float s0, float s1, float s2, float* __restrict__ gray_mean,
                               float* __restrict__ out) {
  const int img = blockIdx.z;
  const int ox = blockIdx.x * 32 + (threadIdx.x & 31);
  const int oy = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int plane = H * W;
  // the image's grey mean from the eight partial sums, in a fixed order (8 L2-resident words, the same in every thread)
  const float* pp = partial + img * kMeanParts;
  const float mean = (((pp[0] + pp[1]) + (pp[2] + pp[3])) + ((pp[4] + pp[5]) + (pp[6] + pp[7]))) / static_cast<float>(plane);
  if (ox == 0 && oy == 0) gray_mean[img] = mean;
  const float* prm = params + img * 8;
  __shared__ float t[4];              // the image's grid scale, four IEEE divisions per CTA instead of per thread
  if (threadIdx.x < 4) t[threadIdx.x] = aug::rotate_scale(prm + 4, threadIdx.x, W, H);
  __syncthreads();
  if (ox >= W || oy >= H) return;
  float r = 0.f, g = 0.f, b = 0.f;   // RandomRotation's fill
  int sx, sy;
  if (aug::rotate_source(ox, oy, W, H, t, sx, sy)) {
    const float* src = img01 + static_cast<long long>(img) * 3 * plane + sy * W + sx;
    r = __ldg(src);
    g = __ldg(src + plane);
    b = __ldg(src + 2 * plane);
    const aug::Jitter j = aug::make_jitter(order[img], prm);
    aug::jitter_pixel(j, 0, 4, mean, r, g, b);
  }
  float* o = out + static_cast<long long>(img) * 3 * plane + oy * W + ox;
  o[0] = aug::normalize(r, m0, s0);
  o[plane] = aug::normalize(g, m1, s1);
  o[2 * plane] = aug::normalize(b, m2, s2);
}

}  // namespace b200

using namespace b200;

// img01 [n, 3, H, W] fp32 in [0, 1]; order [n] int (2 bits per ColorJitter operator, first applied in the low bits);
// params [n, 8] fp32 = brightness, contrast, saturation factors, hue shift, inverse rotation matrix m00 m01 m10 m11;
// gray_mean [9 n] fp32 scratch: entries [0, n) hold each image's contrast mean afterwards, [n, 9 n) the partial sums;
// out [n, 3, H, W] fp32 normalised.  mean3 / std3: HOST arrays.  All other pointers are device pointers.
B200MM_API int b200mm_augment_jitter_rotate(const float* img01, const int* order, const float* params, int n, int H,
                                            int W, const float* mean3, const float* std3, float* gray_mean, float* out,
                                            void* stream) {
  if (n <= 0 || n > 65535 || H <= 0 || W <= 0 || !img01 || !order || !params || !mean3 || !std3 || !gray_mean || !out)
    return B200MM_ERR_BAD_ARG;
  if (img01 == out) return B200MM_ERR_BAD_ARG;   // the rotation gathers: not an in-place transform
  for (int c = 0; c < 3; ++c)
    if (std3[c] == 0.f) return B200MM_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* partial = gray_mean + n;
  const int vec = ((H * W) & 3) == 0 && (reinterpret_cast<uintptr_t>(img01) & 15) == 0;
  jitter_gray_mean_kernel<<<dim3(kMeanParts, n), kMeanThreads, 0, st>>>(img01, order, params, H, W, vec, partial);
  B200MM_CHECK_LAUNCH();
  dim3 grid(ceil_div(W, 32), ceil_div(H, 8), n);
  jitter_rotate_normalize_kernel<<<grid, 256, 0, st>>>(img01, order, params, partial, H, W, mean3[0], mean3[1], mean3[2],
                                                       1.f / std3[0], 1.f / std3[1], 1.f / std3[2], gray_mean, out);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
