#include "common.cuh"
#include <cstdlib>
#include <mutex>

namespace b200 {

const DeviceInfo& device_info() {
  static DeviceInfo info;
  static std::once_flag once;
  std::call_once(once, [] {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    cudaDeviceGetAttribute(&info.num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&info.cc_major, cudaDevAttrComputeCapabilityMajor, dev);
    info.ok = info.num_sms > 0;
  });
  return info;
}

static int env_int(const char* name, int dflt) {
  const char* e = std::getenv(name);
  return e != nullptr ? std::atoi(e) : dflt;
}
int g_tune[TUNE_KNOBS] = {8, 1, 0, env_int("B200MM_GEMM_SMEM_FREE_KB", 0)};

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("B200MM_PDL");
    return e != nullptr && e[0] == '1';   // opt-in: measured 30.7 ms (on) vs 30.2 ms (off) per config-2 step
  }();
  return on;
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                      uint32_t box_inner, uint32_t box_outer, uint32_t swizzle_bytes) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return B200MM_ERR_NO_DRIVER;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (row_stride_bytes & 15) || box_inner * 2 > swizzle_bytes ||
      box_outer > 256 || (swizzle_bytes != 128 && swizzle_bytes != 64))
    return B200MM_ERR_BAD_ARG;
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  // 256-byte L2 promotion only when rows are 256-byte aligned: with e.g. the stem's 304-byte rows it makes every
  // 128-byte box row fetch 256 bytes (2.5x DRAM over-fetch, measured 608 us instead of ~220 us on the stem GEMM)
  const CUtensorMapL2promotion promo =
      (row_stride_bytes % 256 == 0) ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, promo,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? B200MM_OK : B200MM_ERR_TENSORMAP;
}

using EncodeIm2colFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeIm2colFn encode_im2col_fn() {
  static EncodeIm2colFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeIm2colFn>(p);
  });
  return fn;
}

int make_tmap_im2col_bf16(CUtensorMap* map, const void* ptr, int N, int H, int W, int C, int ksize, int stride, int pad,
                          uint32_t pixels_per_column) {
  EncodeIm2colFn fn = encode_im2col_fn();
  if (!fn) return B200MM_ERR_NO_DRIVER;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (C & 7) || C < 64 || pixels_per_column > 1024 || stride < 1 ||
      stride > 8)
    return B200MM_ERR_BAD_ARG;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(N)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(W) * C * 2,
                           static_cast<cuuint64_t>(H) * W * C * 2};
  // bounding box of the base pixel (CUTLASS convention, fprop): lower = -pad, upper = pad - (k - 1)
  int lower[2] = {-pad, -pad};
  int upper[2] = {pad - (ksize - 1), pad - (ksize - 1)};
  cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(stride), static_cast<cuuint32_t>(stride), 1};
  const CUtensorMapL2promotion promo =
      ((C * 2) % 256 == 0) ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, lower, upper, 64,
                  pixels_per_column, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? B200MM_OK : B200MM_ERR_TENSORMAP;
}

int make_tmap_3d_bf16(CUtensorMap* map, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                      uint64_t stride2_bytes, uint32_t box0, uint32_t box1) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return B200MM_ERR_NO_DRIVER;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (stride1_bytes & 15) || (stride2_bytes & 15) || box0 * 2 > 128 ||
      box1 > 256)
    return B200MM_ERR_BAD_ARG;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? B200MM_OK : B200MM_ERR_TENSORMAP;
}

int make_tmap_nhwc_bf16(CUtensorMap* map, const void* ptr, int N, int H, int W, int C, uint32_t box_c, uint32_t box_w,
                        uint32_t box_h) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return B200MM_ERR_NO_DRIVER;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (C & 7) || box_c * 2 > 128 || box_w > 256 || box_h > 256 ||
      box_w == 0 || box_h == 0)
    return B200MM_ERR_BAD_ARG;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(N)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(W) * C * 2,
                           static_cast<cuuint64_t>(H) * W * C * 2};
  cuuint32_t box[4] = {box_c, box_w, box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? B200MM_OK : B200MM_ERR_TENSORMAP;
}

}  // namespace b200

namespace b200 {
namespace {
constexpr int MAX_SALT_SETTERS = 32;
int (*g_salt_setters[MAX_SALT_SETTERS])(const unsigned long long*);
int g_num_salt_setters = 0;
}  // namespace
void register_step_salt_setter(int (*setter)(const unsigned long long*)) {
  if (g_num_salt_setters < MAX_SALT_SETTERS) g_salt_setters[g_num_salt_setters++] = setter;
}
}  // namespace b200

// Point every translation unit's dropout seed salt (device_utils.cuh) at `salt` (device memory, one uint64; nullptr
// = no salt).  Synchronous (cudaMemcpyToSymbol): call outside stream capture.
B200MM_API int b200mm_set_step_salt_ptr(const unsigned long long* salt) {
  for (int i = 0; i < b200::g_num_salt_setters; ++i) {
    const int rc = b200::g_salt_setters[i](salt);
    if (rc != 0) return rc;
  }
  return B200MM_OK;
}

// Set dispatch knob `knob` (see g_tune in common.cuh) to `value`.
B200MM_API int b200mm_tune(int knob, int value) {
  if (knob < 0 || knob >= b200::TUNE_KNOBS) return B200MM_ERR_BAD_ARG;
  b200::g_tune[knob] = value;
  return B200MM_OK;
}

B200MM_API int b200mm_version() { return 100; }

// Number of SMs of the current device (0 if no usable device): lets the host side size grids.
B200MM_API int b200mm_num_sms() { return b200::device_info().num_sms; }
