// 3x3 / stride 1 / pad 1 convolution with 64 input and 64 output channels (ResNet layer1: torchvision/models/resnet.py
// Bottleneck.conv2 / BasicBlock.conv1,2 at 56 x 56), forward (also the stride-1 data gradient, with rotated weights) and
// weight gradient, with the input halo resident in shared memory.
//
// The generic implicit GEMM (gemm_kernel.cuh, TMA im2col loads) fetches every input byte nine times -- once per filter
// tap -- and at C = 64 that L2 -> SM traffic, not HBM or the tensor pipe, bounds it (186 us for 206 MB of HBM traffic).
// Here one work unit is R output rows of one image, R = floor(128 / (W + 2)): ONE tiled TMA load brings the
// (R + 2) x (W + 2) x 64 halo (zero-filled borders) into shared memory as rows of 128 B, and output position
// r = hl * (W + 2) + w reads, for tap (kh, kw), halo row r + kh * (W + 2) + kw -- a constant row offset per tap, so the
// nine taps are nine UMMA descriptors into the SAME tile (SWIZZLE_128B is a function of the absolute shared-memory
// address: a descriptor may start at any 128-byte row, measured in scripts/desc_shift_test.cu).  Positions with
// w >= W are computed and dropped (3.4 % at W = 56).
//   forward : D[r, cout]       = sum_tap halo[r + off_tap, :] . W_tap[cout, :]^T          (36 MMAs of 128 x 64 x 16)
//   wgrad   : D_pair[cin2, cout] += halo[r + off_tap, cin]^T . dy[r, cout]  (MN-major A whose two 64-channel atoms are
//             two taps, i.e. the same bytes at two row offsets; 5 tap pairs; accumulators live in TMEM all kernel long)
#include "common.cuh"
#include "device_utils.cuh"
#include "ptx.cuh"

namespace b200 {
namespace {

constexpr int C3_C = 64;
constexpr int C3_THREADS = 192;            // warp 0: TMA producer, warp 1: MMA issuer, warps 2..5: epilogue
constexpr int C3_STAGES = 3;
constexpr int C3_STAGE_BYTES = 256 * 128;  // halo rows (R + 2) * (W + 2) <= 256
constexpr int C3_W_BYTES = 9 * 8192;       // nine [64 cout x 64 cin] K-major tiles

struct C3Smem {
  static constexpr int W = 0;
  static constexpr int HALO = W + C3_W_BYTES;
  static constexpr int OUT = HALO + C3_STAGES * C3_STAGE_BYTES;   // 2 x 16 KB staged output tiles
  static constexpr int BAR = OUT + 2 * 16384;
  static constexpr int TOTAL = BAR + 128;
};
constexpr int C3_SMEM = C3Smem::TOTAL + 1024;

struct C3Args {
  int N, H, W, Wp, R, groups;   // groups = ceil(H / R) row groups per image
  float* col_stats;             // forward: [128] or nullptr
  float* dw;                    // wgrad: [64, 576] fp32
};

// ------------------------------------------------------------------------------------------------ forward / dgrad
__global__ void __launch_bounds__(C3_THREADS, 1)
conv3x3_c64_fwd_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_w,
                       const __grid_constant__ CUtensorMap tma_out, C3Args a) {
  // 1024-byte alignment (SWIZZLE_128B atoms) by pointer arithmetic ON the __shared__ array: the compiler keeps the
  // address space and emits LDS / STS (the former round-up through uintptr_t turned every access of the tiles,
  // the staging boxes and the bias rows into generic LD.E / ST.E, which queue with the global loads)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem + C3Smem::W;
  uint8_t* sHalo = smem + C3Smem::HALO;
  uint8_t* sOut = smem + C3Smem::OUT;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C3Smem::BAR);   // [3]
  uint64_t* empty = full + C3_STAGES;                                 // [3]
  uint64_t* tfull = empty + C3_STAGES;                                // [2]
  uint64_t* tempty = tfull + 2;                                       // [2]
  uint64_t* wbar = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C3_STAGES; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull + i, 1); mbar_init(tempty + i, 128); }
    mbar_init(wbar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tma_x);
    tma_prefetch_desc(&tma_w);
    tma_prefetch_desc(&tma_out);
  }
  if (warp == 1) tmem_alloc<2 * C3_C>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const int tiles = a.N * a.groups;
  const uint32_t halo_bytes = static_cast<uint32_t>(a.Wp) * (a.R + 2) * 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(wbar, C3_W_BYTES);
      for (int tap = 0; tap < 9; ++tap) tma_load_2d(sW + tap * 8192, &tma_w, wbar, tap * C3_C, 0);
      int it = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        const int st = it % C3_STAGES, use = it / C3_STAGES;
        if (use > 0) mbar_wait(empty + st, (use - 1) & 1);
        const int n = t / a.groups, h0 = (t - n * a.groups) * a.R;
        mbar_expect_tx(full + st, halo_bytes);
        tma_load_4d(sHalo + st * C3_STAGE_BYTES, &tma_x, full + st, 0, -1, h0 - 1, n);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, C3_C, 0, 0);
      const uint32_t sW_u = smem_u32(sW), sH_u = smem_u32(sHalo);
      mbar_wait(wbar, 0);
      int it = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        const int st = it % C3_STAGES, use = it / C3_STAGES;
        const int buf = it & 1, buse = it >> 1;
        mbar_wait(full + st, use & 1);
        if (buse > 0) mbar_wait(tempty + buf, (buse - 1) & 1);
        tc_fence_after_sync();
        const uint32_t base = sH_u + st * C3_STAGE_BYTES;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t off = static_cast<uint32_t>((tap / 3) * a.Wp + tap % 3) * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + buf * C3_C, umma_desc_sw128(base + off + k * 32, 16, 1024),
                      umma_desc_sw128(sW_u + tap * 8192 + k * 32, 16, 1024), idesc, (tap | k) ? 1u : 0u);
        }
        umma_commit(empty + st);
        umma_commit(tfull + buf);
      }
    }
  } else {
    // epilogue: thread <-> accumulator row r = position hl * Wp + w
    const int et = threadIdx.x - 64;                  // 0..127
    const int lg = warp & 3;                          // TMEM lane group this warp may read
    const int r = lg * 32 + lane;
    const int hl = r / a.Wp, w = r - hl * a.Wp;
    const bool pos_ok = hl < a.R && w < a.W;
    const int rc = hl * a.W + w;                      // compact row in the staged tile
    const int st_c16 = et & 7, st_grp = et >> 3;      // statistics: 16-byte chunk x every 16th row
    float st_s[8], st_q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) st_s[i] = st_q[i] = 0.f;
    int it = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
      const int buf = it & 1, buse = it >> 1;
      const int n = t / a.groups, h0 = (t - n * a.groups) * a.R;
      uint8_t* stage = sOut + buf * 16384;
      mbar_wait(tfull + buf, buse & 1);
      tc_fence_after_sync();
      uint32_t v0[32], v1[32];
      tmem_ld32(tmem + (static_cast<uint32_t>(lg * 32) << 16) + buf * C3_C, v0);
      tmem_ld32(tmem + (static_cast<uint32_t>(lg * 32) << 16) + buf * C3_C + 32, v1);
      tmem_ld_wait();
      tc_fence_before_sync();
      mbar_arrive(tempty + buf);
      if (et == 0) tma_store_wait_read<1>();          // the store that last read this staging buffer is done
      named_bar_sync(1, 128);
      if (pos_ok && h0 + hl < a.H) {
        uint8_t* row = stage + rc * 128;
        const int sw = rc & 7;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(v0[8 * j + 0]), __uint_as_float(v0[8 * j + 1]));
          o.y = pack_bf16x2(__uint_as_float(v0[8 * j + 2]), __uint_as_float(v0[8 * j + 3]));
          o.z = pack_bf16x2(__uint_as_float(v0[8 * j + 4]), __uint_as_float(v0[8 * j + 5]));
          o.w = pack_bf16x2(__uint_as_float(v0[8 * j + 6]), __uint_as_float(v0[8 * j + 7]));
          *reinterpret_cast<uint4*>(row + ((j ^ sw) << 4)) = o;
          o.x = pack_bf16x2(__uint_as_float(v1[8 * j + 0]), __uint_as_float(v1[8 * j + 1]));
          o.y = pack_bf16x2(__uint_as_float(v1[8 * j + 2]), __uint_as_float(v1[8 * j + 3]));
          o.z = pack_bf16x2(__uint_as_float(v1[8 * j + 4]), __uint_as_float(v1[8 * j + 5]));
          o.w = pack_bf16x2(__uint_as_float(v1[8 * j + 6]), __uint_as_float(v1[8 * j + 7]));
          *reinterpret_cast<uint4*>(row + (((4 + j) ^ sw) << 4)) = o;
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (et == 0) {
        tma_store_4d(&tma_out, stage, 0, 0, h0, n);   // rows past H are dropped by the TMA unit
        tma_store_commit();
      }
      if (a.col_stats != nullptr) {
        const int rows = (a.H - h0 < a.R ? a.H - h0 : a.R) * a.W;
        for (int q = st_grp; q < rows; q += 16) {
          const uint4 u = *reinterpret_cast<const uint4*>(stage + q * 128 + ((st_c16 ^ (q & 7)) << 4));
          float f[8];
          unpack8(u, f);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            st_s[i] += f[i];
            st_q[i] = fmaf(f[i], f[i], st_q[i]);
          }
        }
      }
    }
    if (a.col_stats != nullptr) {
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
          atomicAdd(a.col_stats + C3_C + st_c16 * 8 + i, st_q[i]);
        }
      }
    }
    if (et == 0) tma_store_wait_read<0>();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<2 * C3_C>(tmem);
}

// ------------------------------------------------------------------------------------------------ weight gradient
struct C3WSmem {
  static constexpr int HALO = 0;                                   // 3 x 32 KB
  static constexpr int DY = HALO + C3_STAGES * C3_STAGE_BYTES;     // 3 x 16 KB ([128 positions x 64 cout])
  static constexpr int BAR = DY + C3_STAGES * 16384;
  static constexpr int TOTAL = BAR + 128;
};
constexpr int C3W_SMEM = C3WSmem::TOTAL + 1024;
constexpr int C3W_THREADS = 192;

__global__ void __launch_bounds__(C3W_THREADS, 1)
conv3x3_c64_wgrad_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_dy, C3Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* sHalo = smem + C3WSmem::HALO;
  uint8_t* sDy = smem + C3WSmem::DY;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C3WSmem::BAR);
  uint64_t* empty = full + C3_STAGES;
  uint64_t* done = empty + C3_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C3_STAGES; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tma_x);
    tma_prefetch_desc(&tma_dy);
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  // dy rows R * Wp .. 127 are never written by TMA and multiply halo rows that belong to other positions: keep them 0
  for (int i = threadIdx.x; i < C3_STAGES * 16384 / 16; i += C3W_THREADS)
    reinterpret_cast<uint4*>(sDy)[i] = make_uint4(0, 0, 0, 0);
  // ... and the halo rows past the loaded box meet only those zero dy rows, but 0 x NaN garbage would still poison the sum
  for (int i = threadIdx.x; i < C3_STAGES * C3_STAGE_BYTES / 16; i += C3W_THREADS)
    reinterpret_cast<uint4*>(sHalo)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const int tiles = a.N * a.groups;
  const uint32_t halo_bytes = static_cast<uint32_t>(a.Wp) * (a.R + 2) * 128;
  const uint32_t dy_bytes = static_cast<uint32_t>(a.Wp) * a.R * 128;
  const bool any = static_cast<int>(blockIdx.x) < tiles;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        const int st = it % C3_STAGES, use = it / C3_STAGES;
        if (use > 0) mbar_wait(empty + st, (use - 1) & 1);
        const int n = t / a.groups, h0 = (t - n * a.groups) * a.R;
        mbar_expect_tx(full + st, halo_bytes + dy_bytes);
        tma_load_4d(sHalo + st * C3_STAGE_BYTES, &tma_x, full + st, 0, -1, h0 - 1, n);
        tma_load_4d(sDy + st * 16384, &tma_dy, full + st, 0, 0, h0, n);   // w = W, W+1 and rows past H read as 0
      }
    }
  } else if (warp == 1) {
    if (any && elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, C3_C, 1, 1);
      const uint32_t sH_u = smem_u32(sHalo), sD_u = smem_u32(sDy);
      int it = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        const int st = it % C3_STAGES, use = it / C3_STAGES;
        mbar_wait(full + st, use & 1);
        tc_fence_after_sync();
        const uint32_t hb = sH_u + st * C3_STAGE_BYTES, db = sD_u + st * 16384;
#pragma unroll
        for (int pr = 0; pr < 5; ++pr) {
          const int ta = 2 * pr, tb = pr < 4 ? 2 * pr + 1 : 8;
          const uint32_t offa = static_cast<uint32_t>((ta / 3) * a.Wp + ta % 3) * 128;
          const uint32_t offb = static_cast<uint32_t>((tb / 3) * a.Wp + tb % 3) * 128;
          const uint32_t lbo = pr < 4 ? offb - offa : 128u;   // pair 4: second atom is unused (any valid distance)
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(tmem + pr * C3_C, umma_desc_sw128(hb + offa + k * 2048, lbo, 1024),
                      umma_desc_sw128(db + k * 2048, 8192, 1024), idesc, (it | k) ? 1u : 0u);
        }
        umma_commit(empty + st);
      }
      umma_commit(done);
    }
  } else if (any) {
    const int lg = warp & 3;
    const int row = lg * 32 + lane;              // accumulator row = (tap within pair) * 64 + cin
    mbar_wait(done, 0);
    tc_fence_after_sync();
#pragma unroll 1
    for (int pr = 0; pr < 5; ++pr) {
      const int tap = 2 * pr + (row >> 6), cin = row & 63;
      const bool use = pr < 4 || row < 64;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        tmem_ld32(tmem + (static_cast<uint32_t>(lg * 32) << 16) + pr * C3_C + half * 32, v);
        tmem_ld_wait();
        if (use) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            atomicAdd(a.dw + (half * 32 + j) * (9 * C3_C) + tap * C3_C + cin, __uint_as_float(v[j]));
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

bool c3_geometry(int N, int H, int W, C3Args& a) {
  if (N <= 0 || H <= 0 || W < 6 || W + 2 > 64) return false;
  a.N = N; a.H = H; a.W = W; a.Wp = W + 2;
  a.R = 128 / a.Wp;
  if (a.R > H) a.R = H;
  if ((a.R + 2) * a.Wp > 256 || 2 * a.Wp + 2 + 128 > 256) return false;
  a.groups = (H + a.R - 1) / a.R;
  a.col_stats = nullptr;
  a.dw = nullptr;
  return static_cast<long long>(N) * a.groups < 0x7fffffffLL;
}

}  // namespace

bool conv3x3_c64_supported(int N, int H, int W) {
  C3Args a;
  return c3_geometry(N, H, W, a);
}

// out[N,H,W,64] = conv3x3(x[N,H,W,64], w[64, 9*64] (OHWI)); col_stats as in b200mm_conv_fwd.
int conv3x3_c64_fwd(const void* x, int N, int H, int W, const void* w, void* out, float* col_stats, cudaStream_t stream) {
  C3Args a;
  if (!c3_geometry(N, H, W, a)) return B200MM_ERR_BAD_ARG;
  const DeviceInfo& di = device_info();
  if (!di.ok) return B200MM_ERR_NOT_SM100;
  CUtensorMap tx, tw, to;
  int rc = make_tmap_nhwc_bf16(&tx, x, N, H, W, C3_C, 64, a.Wp, a.R + 2);
  if (rc != B200MM_OK) return rc;
  rc = make_tmap_2d_bf16(&tw, w, 9 * C3_C, C3_C, 9 * C3_C * 2, 64, 64);
  if (rc != B200MM_OK) return rc;
  rc = make_tmap_nhwc_bf16(&to, out, N, H, W, C3_C, 64, W, a.R);
  if (rc != B200MM_OK) return rc;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_c64_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C3_SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_done = true;
  }
  a.col_stats = col_stats;
  const int tiles = N * a.groups;
  const int grid = tiles < di.num_sms ? tiles : di.num_sms;
  conv3x3_c64_fwd_kernel<<<grid, C3_THREADS, C3_SMEM, stream>>>(tx, tw, to, a);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

// dw[64, 9*64] (fp32) += dy[N,H,W,64]^T (x) patches(x[N,H,W,64])
int conv3x3_c64_wgrad(const void* dy, const void* x, int N, int H, int W, float* dw, cudaStream_t stream) {
  C3Args a;
  if (!c3_geometry(N, H, W, a)) return B200MM_ERR_BAD_ARG;
  const DeviceInfo& di = device_info();
  if (!di.ok) return B200MM_ERR_NOT_SM100;
  CUtensorMap tx, td;
  int rc = make_tmap_nhwc_bf16(&tx, x, N, H, W, C3_C, 64, a.Wp, a.R + 2);
  if (rc != B200MM_OK) return rc;
  rc = make_tmap_nhwc_bf16(&td, dy, N, H, W, C3_C, 64, a.Wp, a.R);
  if (rc != B200MM_OK) return rc;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e =
        cudaFuncSetAttribute(conv3x3_c64_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C3W_SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_done = true;
  }
  a.dw = dw;
  const int tiles = N * a.groups;
  const int grid = tiles < di.num_sms ? tiles : di.num_sms;
  conv3x3_c64_wgrad_kernel<<<grid, C3W_THREADS, C3W_SMEM, stream>>>(tx, td, a);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

}  // namespace b200
