// The reference Dataset's image transform with PILLOW'S OWN ARITHMETIC (csrc/resample_math.cuh): uint8 HWC (decoded image,
// any size) -> Resize (8-bit two-pass bilinear, 22-bit fixed-point coefficients, uint8 between the passes) -> CenterCrop /
// square -> optional horizontal flip -> ToTensor -> Normalize -> fp32 NCHW.  Where preprocess.cu computes the resize in
// floating point (within one uint8 step of the reference's tensors), this kernel's output equals what
// `transforms.Compose([Resize(256), CenterCrop(224), ToTensor(), Normalize(...)])(Image.open(path).convert("RGB"))`
// returns (example_scripts/Multimodal_example_task2C.txt:37-41, :50) BIT FOR BIT -- together with the split JPEG decode
// the whole input tensor of the training step is the reference's.  Opt-in (GpuImageTransform(resample="pillow")): checked
// against Pillow / the reference Dataset run on the host build of the header; not yet timed on a GPU.
//
// One thread per output pixel: the fixed-point coefficients of its column and row windows in IEEE double (a few dozen
// fp64 operations), then ny x nx integer multiply-adds per channel, the horizontal pass's uint8 rounding applied inside
// the vertical window (no intermediate image).
#include "common.cuh"
#include "resample_math.cuh"

namespace b200 {

struct PilPreprocParams {
  const uint8_t* packed;         // ONE byte buffer holding every image of the batch ...
  const long long* offsets;      // ... image i starts at packed + offsets[i]
  const uint8_t* flip;           // per-image horizontal-flip flags or nullptr
  const int* heights;
  const int* widths;
  int n, resize, crop, square;
  float mean[3], std[3];
  float* out;                    // [n, 3, crop, crop] fp32
};

__global__ void __launch_bounds__(256)
preprocess_pil_kernel(const PilPreprocParams p) {
  const int img = blockIdx.z;
  const int ox = blockIdx.x * 32 + (threadIdx.x & 31);
  const int oy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (ox >= p.crop || oy >= p.crop) return;
  uint8_t rgb[3];
  pil::preprocess_pixel_u8(p.packed + p.offsets[img], p.heights[img], p.widths[img], p.resize, p.crop, p.square,
                           p.flip != nullptr && p.flip[img] != 0, ox, oy, rgb);
  const long long plane = static_cast<long long>(p.crop) * p.crop;
  float* o = p.out + static_cast<long long>(img) * 3 * plane + static_cast<long long>(oy) * p.crop + ox;
#pragma unroll
  for (int c = 0; c < 3; ++c)   // ToTensor: byte / 255; Normalize: (x - mean) / std -- torch's operations, one rounding each
    o[c * plane] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(rgb[c]), 255.f), p.mean[c]), p.std[c]);
}

}  // namespace b200

using namespace b200;

// Same contract as b200mm_preprocess_u8_packed, Pillow-exact resize.  Images whose side shrinks by more than 31x are not
// supported (the coefficient window is capped at 64 taps); callers keep such images on the float kernel.
B200MM_API int b200mm_preprocess_u8_packed_pil(const void* packed, const long long* offsets, const int* heights,
                                               const int* widths, const void* flip, int n, int resize, int crop,
                                               int square, const float* mean3, const float* std3, float* out,
                                               void* stream) {
  if (!packed || !offsets || !heights || !widths || !out || !mean3 || !std3 || n <= 0 || n > 65535 || crop <= 0)
    return B200MM_ERR_BAD_ARG;
  if (!square && (resize <= 0 || crop > resize)) return B200MM_ERR_BAD_ARG;
  PilPreprocParams p{};
  p.packed = static_cast<const uint8_t*>(packed);
  p.offsets = offsets;
  p.flip = static_cast<const uint8_t*>(flip);
  p.heights = heights;
  p.widths = widths;
  p.n = n; p.resize = resize; p.crop = crop; p.square = square;
  for (int c = 0; c < 3; ++c) {
    if (std3[c] == 0.f) return B200MM_ERR_BAD_ARG;
    p.mean[c] = mean3[c];
    p.std[c] = std3[c];
  }
  p.out = out;
  dim3 grid(ceil_div(crop, 32), ceil_div(crop, 8), n);
  preprocess_pil_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
