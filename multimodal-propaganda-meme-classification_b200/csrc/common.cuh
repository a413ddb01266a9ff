// Host-side helpers shared by every translation unit of libb200mm: error plumbing, device
// properties, and TMA tensor-map construction through the driver entry point (so the library
// needs no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>

#define B200MM_API extern "C" __attribute__((visibility("default")))

// All C-ABI entry points return 0 on success or a cudaError_t / negative b200mm code.
enum : int {
  B200MM_OK = 0,
  B200MM_ERR_BAD_ARG = -1,      // shape / alignment contract violated
  B200MM_ERR_NO_DRIVER = -2,    // cuTensorMapEncodeTiled not resolvable
  B200MM_ERR_TENSORMAP = -3,    // driver rejected the tensor map
  B200MM_ERR_NOT_SM100 = -4,    // device is not compute capability 10.x
};

#define B200MM_CHECK_LAUNCH()                         \
  do {                                                \
    cudaError_t e__ = cudaGetLastError();             \
    if (e__ != cudaSuccess) return static_cast<int>(e__); \
  } while (0)

namespace b200 {

struct DeviceInfo {
  int num_sms = 0;
  int cc_major = 0;
  bool ok = false;
};
const DeviceInfo& device_info();

// 2-D bf16 tensor map: `inner` contiguous elements per row, `outer` rows `row_stride_bytes` apart,
// box = box_inner x box_outer, SWIZZLE_128B (box_inner * 2 B must be <= 128 B), zero OOB fill.
int make_tmap_2d_bf16(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                      uint32_t box_inner, uint32_t box_outer, uint32_t swizzle_bytes = 128);

// 3-D bf16 tensor map over [d2][d1][d0] (d0 contiguous), box = box0 x box1 x 1, SWIZZLE_128B.
int make_tmap_3d_bf16(CUtensorMap* map, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                      uint64_t stride2_bytes, uint32_t box0, uint32_t box1);

// tiled 4-D bf16 tensor map over an NHWC activation [N][H][W][C] addressed (c, w, h, n): box = box_c x box_w x box_h x 1,
// SWIZZLE_128B, zero OOB fill (halo loads start at w = -1, h = -1).
int make_tmap_nhwc_bf16(CUtensorMap* map, const void* ptr, int N, int H, int W, int C, uint32_t box_c, uint32_t box_w,
                        uint32_t box_h);

// im2col-mode bf16 tensor map over an NHWC activation [N][H][W][C]: square k x k window, symmetric padding,
// traversal stride `stride`; each load delivers pixels_per_column pixels x 64 channels, SWIZZLE_128B.
int make_tmap_im2col_bf16(CUtensorMap* map, const void* ptr, int N, int H, int W, int C, int ksize, int stride, int pad,
                          uint32_t pixels_per_column);

// Dispatch knobs (b200mm_tune; the defaults are the measured optimum, the setter exists for A/B micro-benchmarks and
// for tests that must reach both sides of a dispatch decision):
//   0: smallest reduction depth (k blocks of 64 per tile) the CTA-pair GEMM takes                          (8)
//   1: B-resident GEMM mode for short unsplit K (gemm_kernel.cuh)                                           (1)
//   2: largest tensor in MB whose BatchNorm backward runs as ONE cooperative launch (conv_support.cu); 0 = never (0:
//      measured on config 2 at 52 MB -- ResNet layers 3-4 -- the step takes 30.28 ms against 30.06 ms with two kernels)
//   3: KB of shared memory the GEMM leaves free per SM so that small kernels of a concurrent stream (BatchNorm, LayerNorm,
//      column sums of the other tower) can be co-resident with its persistent CTAs (0; B200MM_GEMM_SMEM_FREE_KB)
constexpr int TUNE_KNOBS = 4;
extern int g_tune[TUNE_KNOBS];

// Programmatic dependent launch (PDL).  A kernel launched through launch_pdl may be scheduled while the kernel before
// it on the stream is still draining (its CTAs start on SMs whose resources are already free and run their prologue);
// it MUST execute pdl_wait() (griddepcontrol.wait: the preceding grid has completed and its writes are visible) before
// its first global-memory access, and should execute pdl_launch_dependents() first thing so that ITS successor may be
// scheduled early in turn.  Inside a stream capture the launches become programmatic graph edges.  Opt-in with
// B200MM_PDL=1: on the config-2 step (one CUDA graph, ~500 kernels) the overlap of set-up with the previous kernel's
// tail did not pay -- 30.7 ms with it against 30.2 ms fully serialised (all parity tests pass either way).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) {
  return (a + b - 1) / b;
}

}  // namespace b200
