// The participant script's train transform with PILLOW'S OWN ARITHMETIC end to end (csrc/resample_math.cuh,
// csrc/augment_pil_math.cuh): uint8 in, uint8 between the operators, exactly as the PIL image travels through the script's
// Compose (example_scripts/Multimodal_example_task2C.py:222-235).  Opt-in companion of augment.cu (which computes the
// float-tensor operators): GpuImageTransform(resample="pillow", augment=True).
//
//   preprocess_pil_u8_kernel   packed decoded images -> Pillow-exact Resize (+ crop / flip) -> uint8 [n, S, S, 3]
//   pil_luma_sum_kernel        per image: sum of the grey levels of the image as the contrast operator sees it (integer
//                              sum, one atomic per thread: order-independent, hence deterministic)
//   pil_jitter_rotate_kernel   one thread per output pixel: 16.16 fixed-point inverse rotation -> source pixel -> the four
//                              uint8 operators in the drawn order -> ToTensor / Normalize -> fp32 NCHW
//
// Byte-identical to torchvision's PIL back end and to a run of the script's own Dataset on the host build of the headers
// (tests/test_cpu.py); written after the round's GPU budget was spent: first GPU run in tests/test_zz_input_refrun_gpu.py.
#include "common.cuh"
#include "resample_math.cuh"
#include "augment_pil_math.cuh"

namespace b200 {

__global__ void __launch_bounds__(256)
preprocess_pil_u8_kernel(const uint8_t* __restrict__ packed, const long long* __restrict__ offsets,
                         const int* __restrict__ heights, const int* __restrict__ widths,
                         const uint8_t* __restrict__ flip, int resize, int crop, int square, uint8_t* __restrict__ out) {
  const int img = blockIdx.z;
  const int ox = blockIdx.x * 32 + (threadIdx.x & 31);
  const int oy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (ox >= crop || oy >= crop) return;
  uint8_t rgb[3];
  pil::preprocess_pixel_u8(packed + offsets[img], heights[img], widths[img], resize, crop, square,
                           flip != nullptr && flip[img] != 0, ox, oy, rgb);
  uint8_t* o = out + ((static_cast<long long>(img) * crop + oy) * crop + ox) * 3;
  o[0] = rgb[0]; o[1] = rgb[1]; o[2] = rgb[2];
}

constexpr int kLumaParts = 8, kLumaThreads = 256;

__device__ __forceinline__ pilaug::Jitter load_jitter(const int* order, const float* alpha, const int* hue, int img) {
  pilaug::Jitter j;
  j.order = order[img];
  j.alpha[0] = alpha[img * 3];
  j.alpha[1] = alpha[img * 3 + 1];
  j.alpha[2] = alpha[img * 3 + 2];
  j.hue = hue[img];
  return j;
}

__global__ void __launch_bounds__(kLumaThreads)
pil_luma_sum_kernel(const uint8_t* __restrict__ img_u8, const int* __restrict__ order, const float* __restrict__ alpha,
                    const int* __restrict__ hue, int H, int W, unsigned long long* __restrict__ sums) {
  const int img = blockIdx.y, part = blockIdx.x;
  const pilaug::Jitter j = load_jitter(order, alpha, hue, img);
  const int upto = pilaug::contrast_position(j.order);
  const int plane = H * W;
  const uint8_t* src = img_u8 + static_cast<long long>(img) * plane * 3;
  const int lo = static_cast<int>(static_cast<long long>(plane) * part / kLumaParts);
  const int hi = static_cast<int>(static_cast<long long>(plane) * (part + 1) / kLumaParts);
  unsigned long long acc = 0;
  for (int i = lo + threadIdx.x; i < hi; i += kLumaThreads) {
    int r = src[3 * i], g = src[3 * i + 1], b = src[3 * i + 2];
    pilaug::jitter_pixel(j, 0, upto, 0, r, g, b);
    acc += static_cast<unsigned long long>(pilaug::luma(r, g, b));
  }
  // one 64-bit atomic per thread (2 048 per image): integer addition is order-independent, so the sum is deterministic; no
  // data passes between threads, which also lets the host emulation in tests/ run this kernel thread by thread
  if (acc) atomicAdd(sums + img, acc);
}

__global__ void __launch_bounds__(256)
pil_jitter_rotate_kernel(const uint8_t* __restrict__ img_u8, const int* __restrict__ order, const float* __restrict__ alpha,
                         const int* __restrict__ hue, const int* __restrict__ affine,
                         const unsigned long long* __restrict__ sums, int H, int W, float m0, float m1, float m2, float s0,
                         float s1, float s2, float* __restrict__ out) {
  const int img = blockIdx.z;
  const int ox = blockIdx.x * 32 + (threadIdx.x & 31);
  const int oy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (ox >= W || oy >= H) return;
  const long long plane = static_cast<long long>(H) * W;
  int r = 0, g = 0, b = 0, xin, yin;                      // the rotation's fill colour
  if (pilaug::rotate_source(affine + img * 6, ox, oy, W, H, xin, yin)) {
    const uint8_t* p = img_u8 + (static_cast<long long>(img) * plane + static_cast<long long>(yin) * W + xin) * 3;
    r = p[0]; g = p[1]; b = p[2];
    const pilaug::Jitter j = load_jitter(order, alpha, hue, img);
    pilaug::jitter_pixel(j, 0, 4, pilaug::contrast_gray(sums[img], plane), r, g, b);
  }
  float* o = out + static_cast<long long>(img) * 3 * plane + static_cast<long long>(oy) * W + ox;
  // ToTensor: byte / 255; Normalize: (x - mean) / std -- torch's three roundings
  o[0] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(r), 255.f), m0), s0);
  o[plane] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(g), 255.f), m1), s1);
  o[2 * plane] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(b), 255.f), m2), s2);
}

}  // namespace b200

using namespace b200;

// Pillow-exact Resize / CenterCrop / flip of a packed batch, as uint8: out [n, crop, crop, 3].
B200MM_API int b200mm_preprocess_u8_packed_pil_u8(const void* packed, const long long* offsets, const int* heights,
                                                  const int* widths, const void* flip, int n, int resize, int crop,
                                                  int square, void* out, void* stream) {
  if (!packed || !offsets || !heights || !widths || !out || n <= 0 || n > 65535 || crop <= 0) return B200MM_ERR_BAD_ARG;
  if (!square && (resize <= 0 || crop > resize)) return B200MM_ERR_BAD_ARG;
  dim3 grid(ceil_div(crop, 32), ceil_div(crop, 8), n);
  preprocess_pil_u8_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(packed), offsets, heights, widths, static_cast<const uint8_t*>(flip), resize, crop,
      square, static_cast<uint8_t*>(out));
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

// ColorJitter + RandomRotation + ToTensor + Normalize with Pillow's arithmetic.  img_u8 [n, H, W, 3]; order int [n];
// alpha fp32 [n, 3] (brightness, contrast, saturation factors as C floats); hue int [n] = uint8(hue_factor * 255);
// affine int [n, 6] = Image.rotate's matrix in 16.16 fixed point as affine_fixed prepares it (a0, a1, a2', a3, a4, a5');
// sums: n x uint64 scratch; out fp32 [n, 3, H, W].  mean3 / std3: HOST arrays, all other pointers device pointers.
B200MM_API int b200mm_augment_pil(const void* img_u8, const int* order, const float* alpha, const int* hue,
                                  const int* affine, int n, int H, int W, const float* mean3, const float* std3,
                                  void* sums, float* out, void* stream) {
  if (!img_u8 || !order || !alpha || !hue || !affine || !mean3 || !std3 || !sums || !out || n <= 0 || n > 65535 ||
      H <= 0 || W <= 0)
    return B200MM_ERR_BAD_ARG;
  for (int c = 0; c < 3; ++c)
    if (std3[c] == 0.f) return B200MM_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(sums, 0, static_cast<size_t>(n) * sizeof(unsigned long long), st);
  if (e != cudaSuccess) return static_cast<int>(e);
  pil_luma_sum_kernel<<<dim3(kLumaParts, n), kLumaThreads, 0, st>>>(static_cast<const uint8_t*>(img_u8), order, alpha, hue,
                                                                    H, W, static_cast<unsigned long long*>(sums));
  B200MM_CHECK_LAUNCH();
  dim3 grid(ceil_div(W, 32), ceil_div(H, 8), n);
  pil_jitter_rotate_kernel<<<grid, 256, 0, st>>>(static_cast<const uint8_t*>(img_u8), order, alpha, hue, affine,
                                                 static_cast<const unsigned long long*>(sums), H, W, mean3[0], mean3[1],
                                                 mean3[2], std3[0], std3[1], std3[2], out);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
