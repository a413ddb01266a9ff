// ResNet stem: 7x7 / stride 2 / pad 3 convolution of the fp32 NCHW image into 64 bf16 NHWC channels, and its weight
// gradient, WITHOUT materialising the im2col matrix (torchvision/models/resnet.py:197 conv1 as instantiated by
// example_scripts/Multimodal_example_task2C.txt:135; 976 MB of columns per step at B = 256 with the lowering this
// replaces).  One work unit = one output row (n, ho): the 7 x 3 input rows it touches are staged in shared memory as
// bf16, the [Wo x 152] patch tile is gathered from them straight into the SWIZZLE_128B layout tcgen05.mma reads, and
//   forward : D[pixel, cout]  = patch[pixel, k] . W[cout, k]^T      (K-major A and B, 12 MMAs of 128 x 64 x 16)
//   wgrad   : D[k, cout]     += patch[pixel, k]^T . dy[pixel, cout] (the same bytes read as an MN-major A; dy arrives
//             by TMA as an MN-major B; the accumulator stays in TMEM for the whole kernel, one red.add pass per CTA).
// k = (kh * 7 + kw) * 3 + c, padded 147 -> 152 (the engine's OHWI weight layout, image_tower.py).
#include "common.cuh"
#include "device_utils.cuh"
#include "ptx.cuh"

namespace b200 {
namespace {

constexpr int ST_KH = 7, ST_KW = 7, ST_CIN = 3, ST_STRIDE = 2, ST_PAD = 3;
static_assert(ST_KH * ST_KW * ST_CIN == 147, "patch rows are 147 taps padded to ST_KP");
constexpr int ST_KP = 152;                     // padded K of the weight / gradient rows
constexpr int ST_COUT = 64;
constexpr int ST_THREADS = 256;
constexpr int ST_MAX_WP = 232;                 // staged row pitch (W + 6 <= 232, i.e. W <= 226)
constexpr int ST_A_BYTES = 3 * 128 * 128;      // patch tile: 3 k-blocks x 128 pixels x 64 k (bf16)
constexpr int ST_ROW_PITCH = ST_CIN * ST_MAX_WP;   // staged row: [wp][c] interleaved, 696 bf16 = 1392 B
constexpr int ST_IN_BYTES = ((ST_KH * ST_ROW_PITCH * 2 + 127) / 128) * 128;

struct StemSmem {
  static constexpr int A = 0;                                // 49152 (1024-aligned)
  static constexpr int X = A + ST_A_BYTES;                   // fwd: weights 3 x 8192; wgrad: dy tile 16384
  static constexpr int OUT = X + 3 * 8192;                   // fwd: staged output tile 128 x 128 B
  static constexpr int IN = OUT + 16384;                     // staged input rows
  static constexpr int BAR = IN + ST_IN_BYTES;               // 2 mbarriers + tmem slot
  static constexpr int TOTAL = BAR + 32;
};
constexpr int ST_SMEM = StemSmem::TOTAL + 1024;              // + alignment slack

struct StemArgs {
  const float* img;
  int N, H, W, Ho, Wo;
  float* col_stats;   // forward: [2 * 64] sum / sum of squares of the bf16 outputs (or nullptr)
  float* dw;          // wgrad: [64, 152] fp32, accumulated
};

// rows 2*ho-3 .. 2*ho+3 of the 3 channels of image n -> sIn[kh][wp][c] (bf16, zero padded, channels
// interleaved so that the 21 taps of one kernel row of one output pixel are 42 contiguous bytes), wp = wi + 3.  Split in a
// register load (24 independent loads per thread, issued a whole tile ahead of their use) and the shared-memory store.
constexpr int ST_ROWS_PER_WARP = (ST_KH * ST_CIN + ST_THREADS / 32 - 1) / (ST_THREADS / 32);   // 3
constexpr int ST_COLS_PER_LANE = (ST_MAX_WP + 31) / 32;                                         // 8
struct StagedRows {
  float v[ST_ROWS_PER_WARP][ST_COLS_PER_LANE];
};
__device__ __forceinline__ void load_rows(const StemArgs& a, int n, int ho, StagedRows& s) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < ST_ROWS_PER_WARP; ++i) {
    const int r = warp + i * (ST_THREADS / 32);
    const int kh = r / ST_CIN, c = r - kh * ST_CIN;
    const int hi = ho * ST_STRIDE - ST_PAD + kh;
    const bool row_ok = r < ST_KH * ST_CIN && hi >= 0 && hi < a.H;
    const float* src = a.img + ((static_cast<long long>(n) * ST_CIN + c) * a.H + (row_ok ? hi : 0)) * a.W;
#pragma unroll
    for (int j = 0; j < ST_COLS_PER_LANE; ++j) {
      const int wi = lane + 32 * j - ST_PAD;
      s.v[i][j] = (row_ok && wi >= 0 && wi < a.W) ? __ldg(src + wi) : 0.f;
    }
  }
}
__device__ __forceinline__ void store_rows(const StagedRows& s, __nv_bfloat16* sIn) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < ST_ROWS_PER_WARP; ++i) {
    const int r = warp + i * (ST_THREADS / 32);
    const int kh = r / ST_CIN, c = r - kh * ST_CIN;
    if (r < ST_KH * ST_CIN) {
#pragma unroll
      for (int j = 0; j < ST_COLS_PER_LANE; ++j) {
        const int wp = lane + 32 * j;
        if (wp < ST_MAX_WP) sIn[kh * ST_ROW_PITCH + wp * ST_CIN + c] = __float2bfloat16(s.v[i][j]);
      }
    }
  }
}

// Patch row of output pixel wo = for kh = 0..6 the 42 bytes at sIn[kh] + 12 * wo, concatenated (k = kh * 21 + kw * 3 + c)
// and zero padded to 152.  One thread per pixel (warps 0..3): 77 conflict-free 4-byte loads (lane stride 12 B), odd kernel
// rows shifted by half a word with a funnel shift, 19 conflict-free 16-byte stores into the SWIZZLE_128B tile.
__device__ __forceinline__ void build_patches(int Wo, const __nv_bfloat16* sIn, uint8_t* sA) {
  const int wo = threadIdx.x;
  if (wo >= Wo) return;
  constexpr int PW = ST_ROW_PITCH / 2;   // row pitch in 32-bit words
  const uint32_t* src = reinterpret_cast<const uint32_t*>(sIn) + 3 * wo;
  uint8_t* dst = sA + wo * 128;
  const int sw = wo & 7;
  uint32_t d[4];
  auto emit = [&](int widx, uint32_t val) {
    d[widx & 3] = val;
    if ((widx & 3) == 3) {
      const int chunk = widx >> 2;
      *reinterpret_cast<uint4*>(dst + (chunk >> 3) * 16384 + (((chunk & 7) ^ sw) << 4)) = make_uint4(d[0], d[1], d[2], d[3]);
    }
  };
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const uint32_t* s = src + (2 * m) * PW;
#pragma unroll
    for (int i = 0; i < 10; ++i) emit(21 * m + i, s[i]);
    const uint32_t s10 = s[10] & 0xFFFFu;
    if (m < 3) {
      const uint32_t* t = src + (2 * m + 1) * PW;
      uint32_t tp = t[0];
      emit(21 * m + 10, s10 | (tp << 16));
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        const uint32_t tn = t[i + 1];
        emit(21 * m + 11 + i, __funnelshift_r(tp, tn, 16));
        tp = tn;
      }
    } else {
      emit(73, s10);
      emit(74, 0u);
      emit(75, 0u);
    }
  }
}

__device__ __forceinline__ void stem_prologue(uint8_t* smem) {
  // zero the patch tile once: rows >= Wo and k-block 2 beyond k = 152 are never written afterwards and must read as 0
  for (int i = threadIdx.x; i < ST_A_BYTES / 16; i += ST_THREADS)
    reinterpret_cast<uint4*>(smem + StemSmem::A)[i] = make_uint4(0, 0, 0, 0);
}

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(ST_THREADS, 2)
stem_fwd_kernel(const __grid_constant__ CUtensorMap tma_w, const __grid_constant__ CUtensorMap tma_out, StemArgs a) {
  // 1024-byte alignment (SWIZZLE_128B atoms) by pointer arithmetic ON the __shared__ array: the compiler keeps the
  // address space and emits LDS / STS (the former round-up through uintptr_t turned every access of the tiles,
  // the staging boxes and the bias rows into generic LD.E / ST.E, which queue with the global loads)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem + StemSmem::A;
  uint8_t* sW = smem + StemSmem::X;
  uint8_t* sOut = smem + StemSmem::OUT;
  __nv_bfloat16* sIn = reinterpret_cast<__nv_bfloat16*>(smem + StemSmem::IN);
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + StemSmem::BAR);
  uint64_t* bar_mma = bar_w + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tma_w);
    tma_prefetch_desc(&tma_out);
  }
  if (warp == 1) tmem_alloc<ST_COUT>(tmem_slot);
  stem_prologue(smem);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_w, 3 * 8192);
#pragma unroll
    for (int kb = 0; kb < 3; ++kb) tma_load_2d(sW + kb * 8192, &tma_w, bar_w, kb * 64, 0);
  }
  const uint32_t idesc = umma_idesc_bf16(128, ST_COUT, 0, 0);
  const uint32_t sA_u = smem_u32(sA), sW_u = smem_u32(sW);
  // BatchNorm column statistics of the bf16 outputs: warps 4..7 re-read the PREVIOUS tile from sOut while warps 0..3
  // build the next patch tile.  Thread -> one 16-byte chunk (8 columns) of every 16th row; partial sums stay in
  // registers for the whole kernel.
  const int st_c16 = threadIdx.x & 7, st_grp = (threadIdx.x >> 3) & 15;
  float st_s[8], st_q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) st_s[i] = st_q[i] = 0.f;
  auto tile_stats = [&]() {
    for (int r = st_grp; r < a.Wo; r += 16) {
      const uint4 u = *reinterpret_cast<const uint4*>(sOut + r * 128 + ((st_c16 ^ (r & 7)) << 4));
      float f[8];
      unpack8(u, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        st_s[i] += f[i];
        st_q[i] = fmaf(f[i], f[i], st_q[i]);
      }
    }
  };
  const bool do_stats = a.col_stats != nullptr && warp >= 4;
  bool have_prev = false;
  uint32_t phase = 0;
  bool w_ready = false;
  const int tiles = a.N * a.Ho;
  StagedRows rows;
  if (static_cast<int>(blockIdx.x) < tiles) load_rows(a, blockIdx.x / a.Ho, blockIdx.x % a.Ho, rows);
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    store_rows(rows, sIn);
    __syncthreads();
    {
      const int tn = t + gridDim.x;
      if (tn < tiles) load_rows(a, tn / a.Ho, tn % a.Ho, rows);   // in flight across this tile's MMA and epilogue
    }
    if (warp < 4) build_patches(a.Wo, sIn, sA);
    else if (do_stats && have_prev) tile_stats();
    have_prev = true;
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0 && elect_one()) {
      if (!w_ready) { mbar_wait(bar_w, 0); w_ready = true; }
      tc_fence_after_sync();
#pragma unroll
      for (int kb = 0; kb < 3; ++kb)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem, umma_desc_sw128(sA_u + kb * 16384 + k * 32, 16, 1024),
                    umma_desc_sw128(sW_u + kb * 8192 + k * 32, 16, 1024), idesc, (kb | k) ? 1u : 0u);
      umma_commit(bar_mma);
    }
    if (threadIdx.x == 0) tma_store_wait_read<0>();   // previous tile's output box has left sOut (same thread stores)
    mbar_wait(bar_mma, phase);
    phase ^= 1;
    tc_fence_after_sync();
    __syncthreads();   // thread 0's store-read wait is visible to every writer of sOut
    {
      // warp w: TMEM lanes 32*(w&3).., columns 32*(w>>2)..
      const int row = (warp & 3) * 32 + lane, half = warp >> 2;
      uint32_t v[32];
      tmem_ld32(tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + half * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
        o.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
        o.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
        o.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
        const int c16 = half * 4 + j;
        *reinterpret_cast<uint4*>(sOut + row * 128 + ((c16 ^ (row & 7)) << 4)) = o;
      }
    }
    tc_fence_before_sync();
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      tma_store_2d(&tma_out, sOut, 0, t * a.Wo);
      tma_store_commit();
    }
  }
  if (do_stats) {
    if (have_prev) tile_stats();   // last tile (sOut is only read from here on)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      st_s[i] += __shfl_xor_sync(0xffffffffu, st_s[i], 8);
      st_q[i] += __shfl_xor_sync(0xffffffffu, st_q[i], 8);
      st_s[i] += __shfl_xor_sync(0xffffffffu, st_s[i], 16);
      st_q[i] += __shfl_xor_sync(0xffffffffu, st_q[i], 16);
    }
    if (lane < 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        atomicAdd(a.col_stats + st_c16 * 8 + i, st_s[i]);
        atomicAdd(a.col_stats + ST_COUT + st_c16 * 8 + i, st_q[i]);
      }
    }
  }
  if (threadIdx.x == 0) tma_store_wait_read<0>();
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<ST_COUT>(tmem);
}

// ------------------------------------------------------------------------------------------------ weight gradient
__global__ void __launch_bounds__(ST_THREADS, 2)
stem_wgrad_kernel(const __grid_constant__ CUtensorMap tma_dy, StemArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* sA = smem + StemSmem::A;
  uint8_t* sDy = smem + StemSmem::X;
  __nv_bfloat16* sIn = reinterpret_cast<__nv_bfloat16*>(smem + StemSmem::IN);
  uint64_t* bar_dy = reinterpret_cast<uint64_t*>(smem + StemSmem::BAR);
  uint64_t* bar_mma = bar_dy + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_dy + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bar_dy, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tma_dy);
  }
  if (warp == 1) tmem_alloc<2 * ST_COUT>(tmem_slot);
  stem_prologue(smem);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  // A operand = patch tile read MN-major (M = k index, K = pixel): 64-wide k atoms are the k-blocks, 16384 B apart.
  // Two accumulators: k 0..127 (k-blocks 0,1) and k 64..191 (k-blocks 1,2); of the second only rows 64.. are used.
  const uint32_t idesc = umma_idesc_bf16(128, ST_COUT, 1, 1);
  const uint32_t sA_u = smem_u32(sA), sDy_u = smem_u32(sDy);
  uint32_t ph_dy = 0, ph_mma = 0;
  bool first = true;
  const int tiles = a.N * a.Ho;
  StagedRows rows;
  if (static_cast<int>(blockIdx.x) < tiles) load_rows(a, blockIdx.x / a.Ho, blockIdx.x % a.Ho, rows);
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    store_rows(rows, sIn);
    {
      const int tn = t + gridDim.x;
      if (tn < tiles) load_rows(a, tn / a.Ho, tn % a.Ho, rows);
    }
    if (!first) {   // the previous tile's MMAs still read sA and sDy
      mbar_wait(bar_mma, ph_mma);
      ph_mma ^= 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(bar_dy, 16384);
      tma_load_2d(sDy, &tma_dy, bar_dy, 0, t * a.Wo);   // 128 rows: the tail belongs to the next row, patch rows are 0
    }
    build_patches(a.Wo, sIn, sA);
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0 && elect_one()) {
      mbar_wait(bar_dy, ph_dy);
      tc_fence_after_sync();
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(tmem + h * ST_COUT, umma_desc_sw128(sA_u + h * 16384 + k * 2048, 16384, 1024),
                    umma_desc_sw128(sDy_u + k * 2048, 8192, 1024), idesc, (first && k == 0) ? 0u : 1u);
      umma_commit(bar_mma);
    }
    ph_dy ^= 1;
    first = false;
  }
  if (!first) {
    mbar_wait(bar_mma, ph_mma);
    tc_fence_after_sync();
    if (warp < 4) {
      const int row = warp * 32 + lane;   // accumulator row
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int kidx = h == 0 ? row : row + 64;
        const bool use = h == 0 ? true : (row >= 64);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v[32];
          tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + h * ST_COUT + half * 32, v);
          tmem_ld_wait();
          if (use && kidx < ST_KP) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              atomicAdd(a.dw + (half * 32 + j) * ST_KP + kidx, __uint_as_float(v[j]));
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<2 * ST_COUT>(tmem);
}

bool stem_shape_ok(int N, int Cin, int H, int W, int Cout, int Kp) {
  if (N <= 0 || Cin != ST_CIN || Cout != ST_COUT || Kp != ST_KP || H < ST_KH || W < ST_KW) return false;
  const int Wo = (W + 2 * ST_PAD - ST_KW) / ST_STRIDE + 1;
  return W + 2 * ST_PAD <= ST_MAX_WP && Wo >= 8 && Wo <= 128;
}

}  // namespace
}  // namespace b200

using namespace b200;

// conv1 forward.  img fp32 [N,3,H,W]; w bf16 [64,152]; out bf16 [N*Ho*Wo, 64] (NHWC); col_stats fp32 [128] or NULL
// (accumulated: zero it first).  Returns B200MM_ERR_BAD_ARG for shapes outside the specialisation (caller lowers).
B200MM_API int b200mm_stem_conv_fwd(const float* img, int N, int Cin, int H, int W, const void* w, int Cout, int Kp,
                                    void* out, float* col_stats, void* stream) {
  if (!stem_shape_ok(N, Cin, H, W, Cout, Kp) || img == nullptr || w == nullptr || out == nullptr)
    return B200MM_ERR_BAD_ARG;
  const DeviceInfo& di = device_info();
  if (!di.ok) return B200MM_ERR_NOT_SM100;
  const int Ho = (H + 2 * ST_PAD - ST_KH) / ST_STRIDE + 1, Wo = (W + 2 * ST_PAD - ST_KW) / ST_STRIDE + 1;
  CUtensorMap tw, to;
  int rc = make_tmap_2d_bf16(&tw, w, ST_KP, ST_COUT, ST_KP * 2, 64, 64);
  if (rc != B200MM_OK) return rc;
  rc = make_tmap_2d_bf16(&to, out, ST_COUT, static_cast<uint64_t>(N) * Ho * Wo, ST_COUT * 2, 64, Wo);
  if (rc != B200MM_OK) return rc;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(stem_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_done = true;
  }
  StemArgs a{img, N, H, W, Ho, Wo, col_stats, nullptr};
  const int tiles = N * Ho;
  const int grid = tiles < 2 * di.num_sms ? tiles : 2 * di.num_sms;
  stem_fwd_kernel<<<grid, ST_THREADS, ST_SMEM, static_cast<cudaStream_t>(stream)>>>(tw, to, a);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

// conv1 weight gradient: dw[64,152] (fp32) += dy[N*Ho*Wo, 64]^T . patches(img).
B200MM_API int b200mm_stem_conv_wgrad(const float* img, int N, int Cin, int H, int W, const void* dy, int Cout,
                                      int Kp, float* dw, void* stream) {
  if (!stem_shape_ok(N, Cin, H, W, Cout, Kp) || img == nullptr || dy == nullptr || dw == nullptr)
    return B200MM_ERR_BAD_ARG;
  const DeviceInfo& di = device_info();
  if (!di.ok) return B200MM_ERR_NOT_SM100;
  const int Ho = (H + 2 * ST_PAD - ST_KH) / ST_STRIDE + 1, Wo = (W + 2 * ST_PAD - ST_KW) / ST_STRIDE + 1;
  CUtensorMap td;
  int rc = make_tmap_2d_bf16(&td, dy, ST_COUT, static_cast<uint64_t>(N) * Ho * Wo, ST_COUT * 2, 64, 128);
  if (rc != B200MM_OK) return rc;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(stem_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_done = true;
  }
  StemArgs a{img, N, H, W, Ho, Wo, nullptr, dw};
  const int tiles = N * Ho;
  const int grid = tiles < 2 * di.num_sms ? tiles : 2 * di.num_sms;
  stem_wgrad_kernel<<<grid, ST_THREADS, ST_SMEM, static_cast<cudaStream_t>(stream)>>>(td, a);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
