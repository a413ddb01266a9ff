// Parameter block and tiling constants shared by the tcgen05 GEMM kernel (gemm_kernel.cuh) and its C-ABI front end.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;              // 64 bf16 = one 128-byte swizzle row
constexpr int GEMM_UMMA_K = 16;
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_THREADS = 64 + GEMM_EPI_WARPS * 32;
constexpr int GEMM_MAX_STAGES = 8;
constexpr int GEMM_BOX_BYTES = 4096;                 // one staged epilogue box = [32 rows x 64 cols] bf16
constexpr int GEMM_SMEM_TOTAL = 227 * 1024 - 1024;   // dynamic smem we ask for (minus alignment slack)

enum EpiMode : int {
  EPI_STORE = 0,       // out = bf16(acc + bias + residual)
  EPI_GELU = 1,        // out = bf16(z = acc + bias); out2 = bf16(gelu(z))
  EPI_DGELU = 2,       // out = bf16(acc * gelu'(aux))
  EPI_F32 = 3,         // out_f32 = acc + bias
  EPI_F32_ATOMIC = 4,  // out_f32 += acc        (split-K partial sums)
  EPI_RELU = 5,        // out = bf16(max(acc + bias + residual, 0))
  EPI_STORE_DROP = 6,  // internal: EPI_STORE with inverted dropout (own instantiation keeps EPI_STORE lean)
  EPI_STORE_MASKRES = 8, // internal: out = bf16(acc + (mask bit ? residual : 0)) -- the identity-branch gradient of a
                         // residual block joins the data gradient without ever being materialised
  EPI_STORE_STATS = 7, // internal: out = bf16(acc) and col_stats[n] += sum_m out, col_stats[N + n] += sum_m out^2
                       // (train-mode BatchNorm statistics of a convolution output, taken on the stored bf16 values)
};

struct GemmParams {
  int M, N, K;
  int a_mn, b_mn;  // operand majorness flags (0 = K-major, 1 = MN-major)
  int m_tiles, n_tiles, splits, k_iters, k_iters_per_split;
  int epi;
  const float* bias;
  const __nv_bfloat16* residual;
  long long ldr;
  const __nv_bfloat16* aux;
  long long ld_aux;
  void* out;
  long long ldc;
  __nv_bfloat16* out2;
  long long ld2;
  const unsigned char* res_mask;  // EPI_STORE_MASKRES: [M, N / 8] bytes, bit i of byte j = element 8 j + i
  long long ld_mask;              // ... row stride in bytes
  float* col_stats;  // EPI_STORE_STATS: fp32 [2N] accumulators (pre-zeroed by the caller)
  int accumulate;    // EPI_STORE with residual == out: out += acc through TMA reduce-add stores (no residual loads)
  // EPI_STORE only: inverted dropout on (acc + bias) before the residual add, mask keyed by (seed, row * N + col)
  float p_drop;
  uint32_t drop_threshold;
  float inv_keep;
  unsigned long long seed;
  // implicit-GEMM convolution: an operand gathered by TMA im2col loads from an NHWC activation instead of a matrix.
  //   a_im2col: A tile = 128 output pixels x 64 channels of tap (kb / cblocks)            (forward / stride-1 dgrad)
  //   b_im2col: B atom = 64 pixels (reduction dim) x 64 channels of tap (col / conv_C)     (wgrad)
  int a_im2col, b_im2col;
  int conv_C, conv_KW, conv_stride, conv_pad, conv_P, conv_Q, conv_cblocks;
  int b_resident;  // single CTA, unsplit short K: the [BN x K] B tile stays in shared memory for all tiles of the CTA
  int num_stages;  // smem ring depth (runtime: deep ring for long-K tiles, ...)
  int nbuf;        // ... or two epilogue staging boxes per warp for short-K, store-heavy tiles
};

}  // namespace b200
