// Fused multi-head self-attention (head_dim 64, S <= 512) forward and backward on tcgen05 + TMA.
//
// Persistent CTAs of 256 threads; thread (row, half) owns query row `row` (its TMEM lane) and half of the key columns:
//   forward :  S = Q K^T for every key tile (tcgen05.mma, fp32 in TMEM) -> scale + additive key-padding bias -> exact
//              row maximum over all tiles -> P = exp2(S - max) -> Philox dropout -> P (bf16) staged in 128B-swizzled
//              smem -> O += P V (tcgen05.mma) -> 1/rowsum -> staged tile -> TMA store.  LSE kept for backward.
//   backward:  recompute S and dP = dO V^T on the tensor core, P = exp(S - LSE), dS = P o (dP - delta),
//              then dV = P^T dO, dK = dS^T Q, dQ = dS K -- the transposed operands are the SAME smem tiles
//              read through MN-major UMMA descriptors, nothing is transposed in memory.
// Kernels by length: forward S <= 384 (all score tiles resident in TMEM) / longer (two-pass); backward S <= 128
// (pipelined across heads), <= 256 and <= 384 (whole head per CTA, every accumulator in TMEM), longer (two-kind).
// Q/K/V are read straight out of the fused-QKV projection output [B*S, 3*D] (and dQ/dK/dV written into
// the matching [B*S, 3*D] gradient) with 3-D TMA boxes, so no head-major re-layout pass exists.
//
// Replaces: transformers/models/distilbert/modeling_distilbert.py:126-151
// (eager_attention_forward: softmax(QK^T * d^-1/2 + mask) -> dropout -> @V) and its autograd backward.
#include "attention_common.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------ S > 128
// Longer sequences (the reference pads to 512 tokens, example_scripts/Multimodal_example_task2C.txt:14) run the same
// tile maths over several 128-key tiles.  Forward: two passes over the key tiles (pass 1: row maxima; pass 2:
// P = exp(S - max), O += P V accumulated in TMEM) -- no rescaling of the accumulator is ever needed.  Backward: the
// saved LSE makes every (query tile, key tile) pair independent, so one CTA owns a key tile and accumulates dK / dV
// over the query tiles in TMEM, another owns a query tile and accumulates dQ over the key tiles: no atomics.
constexpr int ATT_MAX_S = 512;

__global__ void __launch_bounds__(128, 2)
attn_fwd_multi_kernel(const __grid_constant__ CUtensorMap tma_qkv, const AttnParams p) {
  // 1024-byte alignment (SWIZZLE_128B atoms) by pointer arithmetic ON the __shared__ array: the compiler keeps the
  // address space and emits LDS / STS (the former round-up through uintptr_t turned every access of the tiles,
  // the staging boxes and the bias rows into generic LD.E / ST.E, which queue with the global loads)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATT_TILE_BYTES;
  uint8_t* sV = sK + ATT_TILE_BYTES;
  uint8_t* sP = sV + ATT_TILE_BYTES;  // 32 KB
  float* sBias = reinterpret_cast<float*>(sP + 2 * ATT_TILE_BYTES);   // [ATT_MAX_S]
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(sBias + ATT_MAX_S);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    tma_prefetch_desc(&tma_qkv);
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_S = tmem, t_O = tmem + 128;
  const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
  const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
  const uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
  const bool use_drop = p.p_drop > 0.f;
  const int nt = (p.S + ATT_T - 1) / ATT_T;
  const int s_pad = nt * ATT_T;

  uint32_t ph_load = 0, ph_mma = 0;
  const int items = p.B * p.H * nt;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int qi = item % nt, bh = item / nt;
    const int b = bh / p.H, h = bh - b * p.H;
    const int q0 = qi * ATT_T;
    __syncthreads();   // previous item's readers of sBias are done
    for (int k = tid; k < s_pad; k += 128)
      sBias[k] = k < p.S ? (p.key_bias ? p.key_bias[b * p.S + k] * LOG2E : 0.f) : -INFINITY;
    if (tid == 0) {
      mbar_expect_tx(bar_load, ATT_TILE_BYTES);
      tma_load_3d(sQ, &tma_qkv, bar_load, h * ATT_D, q0, b);
    }
    __syncthreads();
    mbar_wait(bar_load, ph_load);
    ph_load ^= 1;

    // ---- pass 1: row maxima over all key tiles
    float mx = -INFINITY;
    for (int j = 0; j < nt; ++j) {
      if (tid < 32 && elect_one()) {   // one thread of warp 0 issues (ptx.cuh: elect_one)
        mbar_expect_tx(bar_load, ATT_TILE_BYTES);
        tma_load_3d(sK, &tma_qkv, bar_load, p.D + h * ATT_D, j * ATT_T, b);
        mbar_wait(bar_load, ph_load);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(t_S, umma_desc_sw128(smem_u32(sQ) + k * 32, 16, 1024),
                    umma_desc_sw128(smem_u32(sK) + k * 32, 16, 1024), idesc_s, k > 0);
        umma_commit(bar_mma);
      }
      ph_load ^= 1;
      mbar_wait(bar_mma, ph_mma);
      ph_mma ^= 1;
      tc_fence_after_sync();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(t_S + lane_addr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          mx = fmaxf(mx, fmaf(__uint_as_float(v[i]), p.scale_log2, sBias[j * ATT_T + c * 32 + i]));
      }
      tc_fence_before_sync();
      __syncthreads();   // everybody has read S (and sK is free) before the next tile overwrites them
    }
    if (mx == -INFINITY) mx = 0.f;

    // ---- pass 2: P = exp2(S - max), O += P V
    float sum = 0.f;
    for (int j = 0; j < nt; ++j) {
      if (tid < 32 && elect_one()) {   // one thread of warp 0 issues (ptx.cuh: elect_one)
        mbar_expect_tx(bar_load, 2 * ATT_TILE_BYTES);
        tma_load_3d(sK, &tma_qkv, bar_load, p.D + h * ATT_D, j * ATT_T, b);
        tma_load_3d(sV, &tma_qkv, bar_load, 2 * p.D + h * ATT_D, j * ATT_T, b);
        mbar_wait(bar_load, ph_load);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(t_S, umma_desc_sw128(smem_u32(sQ) + k * 32, 16, 1024),
                    umma_desc_sw128(smem_u32(sK) + k * 32, 16, 1024), idesc_s, k > 0);
        umma_commit(bar_mma);
      }
      ph_load ^= 1;
      mbar_wait(bar_mma, ph_mma);
      ph_mma ^= 1;
      tc_fence_after_sync();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(t_S + lane_addr + c * 32, v);
        tmem_ld_wait();
        float x[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          x[i] = exp2f(fmaf(__uint_as_float(v[i]), p.scale_log2, sBias[j * ATT_T + c * 32 + i]) - mx);
          sum += x[i];
        }
        if (use_drop) {
          const uint32_t keep = dropout_keep32(p.seed, drop_chunk(bh, s_pad, q0 + tid, j * ATT_T + c * 32),
                                               p.drop_threshold >> 16);
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = (keep >> i) & 1 ? x[i] * p.inv_keep : 0.f;
        }
        store_row32_sw128(sP, tid, c * 32, x);
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncthreads();
      if (tid < 32 && elect_one()) {   // one thread of warp 0 issues (ptx.cuh: elect_one)
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(t_O, umma_desc_sw128(smem_u32(sP) + (k >> 2) * ATT_TILE_BYTES + (k & 3) * 32, 16, 1024),
                    umma_desc_sw128(smem_u32(sV) + k * 2048, 8192, 1024), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(bar_mma);
      }
      mbar_wait(bar_mma, ph_mma);   // P V done: sP / sK / sV may be overwritten
      ph_mma ^= 1;
      tc_fence_after_sync();
    }

    const float inv_sum = 1.f / sum;
    const bool row_ok = q0 + tid < p.S;
    __nv_bfloat16* orow = p.out + (static_cast<long long>(b) * p.S + q0 + tid) * p.D + h * ATT_D;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld32(t_O + lane_addr + c * 32, v);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(v[q * 8 + 0]) * inv_sum, __uint_as_float(v[q * 8 + 1]) * inv_sum);
          o.y = pack_bf16x2(__uint_as_float(v[q * 8 + 2]) * inv_sum, __uint_as_float(v[q * 8 + 3]) * inv_sum);
          o.z = pack_bf16x2(__uint_as_float(v[q * 8 + 4]) * inv_sum, __uint_as_float(v[q * 8 + 5]) * inv_sum);
          o.w = pack_bf16x2(__uint_as_float(v[q * 8 + 6]) * inv_sum, __uint_as_float(v[q * 8 + 7]) * inv_sum);
          *reinterpret_cast<uint4*>(orow + c * 32 + q * 8) = o;
        }
      }
    }
    if (row_ok && p.lse) p.lse[static_cast<long long>(bh) * p.S + q0 + tid] = (mx + log2f(sum)) * LN2;
    tc_fence_before_sync();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc<256>(tmem);
  }
}

// kind 0: one key tile j, loop over query tiles -> dK_j, dV_j.   kind 1: one query tile i, loop over key tiles -> dQ_i.
__global__ void __launch_bounds__(128, 1)
attn_bwd_multi_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_do,
                      const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATT_TILE_BYTES;
  uint8_t* sV = sK + ATT_TILE_BYTES;
  uint8_t* sdO = sV + ATT_TILE_BYTES;
  uint8_t* sP = sdO + ATT_TILE_BYTES;
  uint8_t* sdS = sP + 2 * ATT_TILE_BYTES;
  float* sBias = reinterpret_cast<float*>(sdS + 2 * ATT_TILE_BYTES);  // [128]: bias of the current key tile
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(sBias + ATT_T);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    tma_prefetch_desc(&tma_qkv);
    tma_prefetch_desc(&tma_do);
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_S = tmem, t_dP = tmem + 128, t_A = tmem + 256, t_B = tmem + 320;  // A: dV or dQ, B: dK
  const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
  const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
  const uint32_t idesc_tt = umma_idesc_bf16(128, 64, 1, 1);
  const uint32_t idesc_nt = umma_idesc_bf16(128, 64, 0, 1);
  const bool use_drop = p.p_drop > 0.f;
  const int nt = (p.S + ATT_T - 1) / ATT_T;
  const int s_pad = nt * ATT_T;

  uint32_t ph_load = 0, ph_mma = 0;
  const int per_kind = p.B * p.H * nt;
  for (int item = blockIdx.x; item < 2 * per_kind; item += gridDim.x) {
    const int kind = item / per_kind;
    const int rest = item - kind * per_kind;
    const int fixed = rest % nt, bh = rest / nt;      // fixed = key tile (kind 0) or query tile (kind 1)
    const int b = bh / p.H, h = bh - b * p.H;
    for (int it = 0; it < nt; ++it) {
      const int qi = kind == 0 ? it : fixed;
      const int kj = kind == 0 ? fixed : it;
      const int q0 = qi * ATT_T, k0 = kj * ATT_T;
      __syncthreads();
      sBias[tid] = k0 + tid < p.S ? (p.key_bias ? p.key_bias[b * p.S + k0 + tid] * LOG2E : 0.f) : -INFINITY;
      if (tid == 0) {
        mbar_expect_tx(bar_load, 4 * ATT_TILE_BYTES);
        tma_load_3d(sQ, &tma_qkv, bar_load, h * ATT_D, q0, b);
        tma_load_3d(sK, &tma_qkv, bar_load, p.D + h * ATT_D, k0, b);
        tma_load_3d(sV, &tma_qkv, bar_load, 2 * p.D + h * ATT_D, k0, b);
        tma_load_3d(sdO, &tma_do, bar_load, h * ATT_D, q0, b);
      }
      const bool qrow_ok = q0 + tid < p.S;
      float delta = 0.f, lse_l2 = INFINITY;
      if (qrow_ok) {
        const long long off = (static_cast<long long>(b) * p.S + q0 + tid) * p.D + h * ATT_D;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float a[8], g[8];
          load8(p.o_in + off + q * 8, a);
          load8(p.do_in + off + q * 8, g);
#pragma unroll
          for (int i = 0; i < 8; ++i) delta = fmaf(a[i], g[i], delta);
        }
        lse_l2 = p.lse[static_cast<long long>(bh) * p.S + q0 + tid] * LOG2E;
      }
      tc_fence_before_sync();
      __syncthreads();
      if (tid < 32 && elect_one()) {   // one thread of warp 0 issues (ptx.cuh: elect_one)
        mbar_wait(bar_load, ph_load);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(t_S, umma_desc_sw128(smem_u32(sQ) + k * 32, 16, 1024),
                    umma_desc_sw128(smem_u32(sK) + k * 32, 16, 1024), idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(t_dP, umma_desc_sw128(smem_u32(sdO) + k * 32, 16, 1024),
                    umma_desc_sw128(smem_u32(sV) + k * 32, 16, 1024), idesc_s, k > 0);
        umma_commit(bar_mma);
      }
      ph_load ^= 1;
      mbar_wait(bar_mma, ph_mma);
      ph_mma ^= 1;
      tc_fence_after_sync();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t vs[32], vp[32];
        tmem_ld32(t_S + lane_addr + c * 32, vs);
        tmem_ld32(t_dP + lane_addr + c * 32, vp);
        tmem_ld_wait();
        float pd[32], ds[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          pd[i] = exp2f(fmaf(__uint_as_float(vs[i]), p.scale_log2, sBias[c * 32 + i]) - lse_l2);
          ds[i] = __uint_as_float(vp[i]);
        }
        if (use_drop) {
          const uint32_t keep =
              dropout_keep32(p.seed, drop_chunk(bh, s_pad, q0 + tid, k0 + c * 32), p.drop_threshold >> 16);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float m = (keep >> i) & 1 ? p.inv_keep : 0.f;
            const float prob = pd[i];
            pd[i] = prob * m;
            ds[i] = prob * (ds[i] * m - delta) * p.scale;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) ds[i] = pd[i] * (ds[i] - delta) * p.scale;
        }
        if (kind == 0) store_row32_sw128(sP, tid, c * 32, pd);
        store_row32_sw128(sdS, tid, c * 32, ds);
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncthreads();
      if (tid < 32 && elect_one()) {   // one thread of warp 0 issues (ptx.cuh: elect_one)
        tc_fence_after_sync();
        const uint32_t accum = it > 0 ? 1u : 0u;
        if (kind == 0) {
#pragma unroll
          for (int k = 0; k < 8; ++k)   // dV_j += P^T dO_i
            umma_bf16(t_A, umma_desc_sw128(smem_u32(sP) + k * 2048, ATT_TILE_BYTES, 1024),
                      umma_desc_sw128(smem_u32(sdO) + k * 2048, 8192, 1024), idesc_tt, (accum || k > 0) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 8; ++k)   // dK_j += dS^T Q_i
            umma_bf16(t_B, umma_desc_sw128(smem_u32(sdS) + k * 2048, ATT_TILE_BYTES, 1024),
                      umma_desc_sw128(smem_u32(sQ) + k * 2048, 8192, 1024), idesc_tt, (accum || k > 0) ? 1u : 0u);
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k)   // dQ_i += dS K_j
            umma_bf16(t_A, umma_desc_sw128(smem_u32(sdS) + (k >> 2) * ATT_TILE_BYTES + (k & 3) * 32, 16, 1024),
                      umma_desc_sw128(smem_u32(sK) + k * 2048, 8192, 1024), idesc_nt, (accum || k > 0) ? 1u : 0u);
        }
        umma_commit(bar_mma);
      }
      mbar_wait(bar_mma, ph_mma);   // operands in smem are free again
      ph_mma ^= 1;
      tc_fence_after_sync();
    }
    // ---- store the accumulated tile(s): rows = keys (kind 0) or queries (kind 1) of the fixed tile
    const int r0 = fixed * ATT_T;
    const bool row_ok = r0 + tid < p.S;
    __nv_bfloat16* grow = p.dqkv + (static_cast<long long>(b) * p.S + r0 + tid) * (3 * p.D) + h * ATT_D;
    const int nout = kind == 0 ? 2 : 1;
#pragma unroll 1
    for (int which = 0; which < nout; ++which) {
      // kind 0: which 0 -> dV (column block 2), which 1 -> dK (block 1); kind 1: dQ (block 0)
      const uint32_t t_src = which == 0 ? t_A : t_B;
      const int blk = kind == 0 ? (which == 0 ? 2 : 1) : 0;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(t_src + lane_addr + c * 32, v);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1]));
            o.y = pack_bf16x2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3]));
            o.z = pack_bf16x2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5]));
            o.w = pack_bf16x2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7]));
            *reinterpret_cast<uint4*>(grow + blk * p.D + c * 32 + q * 8) = o;
          }
        }
      }
    }
    tc_fence_before_sync();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------------------ S <= 384, forward
// All score tiles of a query tile stay in TMEM (NT x 128 columns; NT = ceil(S / 128) <= 3), so the forward is ONE pass:
// Q, every K tile and every V tile arrive with one TMA wave, S_j = Q K_j^T for all j is issued back to back, the exact
// row maximum is taken over the NT x 128 columns, then P_j = exp2(S_j - max) and O += P_j V_j per key tile.  O
// accumulates in the first 64 columns of S_0 (already consumed when the first P V is issued) and P reuses Q's shared
// memory, which keeps the CTA at NT*32 + 32 KB of shared memory and NT*128 TMEM columns: 2 CTAs / SM for the ViT
// lengths (197, 257 -> NT = 2, 3).  Replaces the two-pass kernel above for these lengths (that one recomputed Q K^T
// and reloaded K per pass: 299 us -> see profiles/ for the measured time at B=256, H=12, S=197).
template <int NT>
__global__ void __launch_bounds__(256, NT <= 2 ? 2 : 1)
attn_fwd_tmem_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_out,
                     const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* sQP = smem;                              // Q [128 x 64] first, then P [128 x 128], then the staged O tile
  uint8_t* sK = sQP + 2 * ATT_TILE_BYTES;           // NT tiles
  uint8_t* sV = sK + NT * ATT_TILE_BYTES;           // NT tiles
  float* sBias = reinterpret_cast<float*>(sV + NT * ATT_TILE_BYTES);   // [NT * 128]
  float* sRed = sBias + NT * ATT_T;                 // [2][128]: row maxima / row sums exchanged between the two halves
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(sRed + 2 * ATT_T);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);
  constexpr int TMEM_COLS = NT == 1 ? 128 : (NT == 2 ? 256 : 512);

  // 256 threads: two warps per TMEM lane quarter; thread (row, half) owns 64 of every tile's 128 key columns
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = (warp & 3) * 32 + lane, half = warp >> 2;
  if (tid == 0) {
    tma_prefetch_desc(&tma_qkv);
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_O = tmem;   // aliases columns 0..63 of S_0
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
  const uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
  const bool use_drop = p.p_drop > 0.f;
  const int nt = (p.S + ATT_T - 1) / ATT_T;      // query tiles per head (== NT key tiles)
  const int s_pad = nt * ATT_T;

  uint32_t ph_load = 0, ph_mma = 0;
  const int items = p.B * p.H * nt;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int qi = item % nt, bh = item / nt;
    const int b = bh / p.H, h = bh - b * p.H;
    const int q0 = qi * ATT_T;
    tc_fence_before_sync();
    __syncthreads();   // previous item: everybody has read O / sBias / sRed, the tensor core is done with the tiles
    if (tid == 0) {
      tma_store_wait_read<0>();   // the previous O tile has left sQP
      mbar_expect_tx(bar_load, (1 + 2 * NT) * ATT_TILE_BYTES);
      tma_load_3d(sQP, &tma_qkv, bar_load, h * ATT_D, q0, b);
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        tma_load_3d(sK + j * ATT_TILE_BYTES, &tma_qkv, bar_load, p.D + h * ATT_D, j * ATT_T, b);
        tma_load_3d(sV + j * ATT_TILE_BYTES, &tma_qkv, bar_load, 2 * p.D + h * ATT_D, j * ATT_T, b);
      }
    }
    for (int k = tid; k < NT * ATT_T; k += 256)
      sBias[k] = k < p.S ? (p.key_bias ? p.key_bias[b * p.S + k] * LOG2E : 0.f) : -INFINITY;
    __syncthreads();
    if (tid < 32 && elect_one()) {   // one thread of warp 0 issues (ptx.cuh: elect_one)
      mbar_wait(bar_load, ph_load);
      tc_fence_after_sync();
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem + j * ATT_T, umma_desc_sw128(smem_u32(sQP) + k * 32, 16, 1024),
                    umma_desc_sw128(smem_u32(sK + j * ATT_TILE_BYTES) + k * 32, 16, 1024), idesc_s, k > 0);
      umma_commit(bar_mma);
    }
    ph_load ^= 1;
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tc_fence_after_sync();

    // ---- exact row maximum over all NT x 128 key columns: own 64 columns of every tile, then the partner's maximum
    float mx = -INFINITY;
#pragma unroll 1
    for (int jc = 0; jc < NT * 2; ++jc) {
      const int col = (jc >> 1) * ATT_T + half * 64 + (jc & 1) * 32;
      uint32_t v[32];
      tmem_ld32(tmem + lane_addr + col, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaf(__uint_as_float(v[i]), p.scale_log2, sBias[col + i]));
    }
    sRed[half * ATT_T + row] = mx;
    __syncthreads();
    mx = fmaxf(mx, sRed[(half ^ 1) * ATT_T + row]);
    if (mx == -INFINITY) mx = 0.f;

    // ---- P_j = exp2(S_j - max) -> shared memory -> O += P_j V_j
    float sum = 0.f;
#pragma unroll 1
    for (int j = 0; j < NT; ++j) {
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c = half * 2 + cc;
        uint32_t v[32];
        tmem_ld32(tmem + lane_addr + j * ATT_T + c * 32, v);
        tmem_ld_wait();
        float x[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          x[i] = fast_exp2(fmaf(__uint_as_float(v[i]), p.scale_log2, sBias[j * ATT_T + c * 32 + i]) - mx);
          sum += x[i];
        }
        if (use_drop) {
          const uint32_t keep = dropout_keep32(p.seed, drop_chunk(bh, s_pad, q0 + row, j * ATT_T + c * 32),
                                               p.drop_threshold >> 16);
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = (keep >> i) & 1 ? x[i] * p.inv_keep : 0.f;
        }
        store_row32_sw128(sQP, row, c * 32, x);
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncthreads();   // P_j complete in shared memory; every thread has finished reading S_j from TMEM
      if (tid < 32 && elect_one()) {   // one thread of warp 0 issues (ptx.cuh: elect_one)
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(t_O, umma_desc_sw128(smem_u32(sQP) + (k >> 2) * ATT_TILE_BYTES + (k & 3) * 32, 16, 1024),
                    umma_desc_sw128(smem_u32(sV + j * ATT_TILE_BYTES) + k * 2048, 8192, 1024), idesc_o,
                    (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(bar_mma);
      }
      mbar_wait(bar_mma, ph_mma);   // P V_j done: the P buffer may be overwritten / O may be read
      ph_mma ^= 1;
      tc_fence_after_sync();
    }

    // ---- row sums of the two halves, then O / sum -> staged tile -> one TMA store (rows >= S clipped by the map)
    __syncthreads();               // everybody is past its read of the partner's maximum
    sRed[half * ATT_T + row] = sum;
    __syncthreads();
    sum += sRed[(half ^ 1) * ATT_T + row];
    const float inv_sum = 1.f / sum;
    {
      uint32_t v[32];
      tmem_ld32(t_O + lane_addr + half * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(v[q * 8 + 0]) * inv_sum, __uint_as_float(v[q * 8 + 1]) * inv_sum);
        o.y = pack_bf16x2(__uint_as_float(v[q * 8 + 2]) * inv_sum, __uint_as_float(v[q * 8 + 3]) * inv_sum);
        o.z = pack_bf16x2(__uint_as_float(v[q * 8 + 4]) * inv_sum, __uint_as_float(v[q * 8 + 5]) * inv_sum);
        o.w = pack_bf16x2(__uint_as_float(v[q * 8 + 6]) * inv_sum, __uint_as_float(v[q * 8 + 7]) * inv_sum);
        *reinterpret_cast<uint4*>(sQP + sw128_off(row, half * 4 + q)) = o;
      }
    }
    if (half == 0 && q0 + row < p.S && p.lse) p.lse[static_cast<long long>(bh) * p.S + q0 + row] = (mx + log2f(sum)) * LN2;
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tma_store_3d(&tma_out, sQP, h * ATT_D, q0, b);
      tma_store_commit();
    }
  }
  if (tid == 0) tma_store_wait_read<0>();

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc<TMEM_COLS>(tmem);
  }
}

// ------------------------------------------------------------------------------------------ 128 < S <= 256, backward
// One CTA (256 threads) owns a whole (batch, head): Q, K, V, dO of both 128-row tiles are loaded ONCE (128 KB), and all
// five gradient accumulators live in TMEM next to the score tiles: S | dP | dK_j dV_j | dQ_0 dQ_1 = 512 columns.  For
// each of the four (query tile i, key tile j) pairs: S = Q_i K_j^T and dP = dO_i V_j^T on the tensor core, P and dS in
// registers (two warps per TMEM lane quarter split the 128 key columns -- the backward needs no row reductions: LSE
// and delta = rowsum(dO o O) are per-row scalars), then dV_j += P^T dO_i, dK_j += dS^T Q_i, dQ_i += dS K_j.  Every
// S / dP tile is computed once (the two-kind kernel above computes each twice and reloads its operands per pair:
// 924 us -> see profiles/ for the measured time at B=256, H=12, S=197).
constexpr int ATT_BWD2_THREADS = 256;
constexpr int ATT_BWD2_SMEM = 12 * ATT_TILE_BYTES + 2 * ATT_T * 4 + 64 + 1024;
__global__ void __launch_bounds__(ATT_BWD2_THREADS, 1)
attn_bwd_tmem_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_do,
                     const __grid_constant__ CUtensorMap tma_dqkv, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* sQ = smem;                               // 2 tiles each: [tile][128 x 64]
  uint8_t* sK = sQ + 2 * ATT_TILE_BYTES;
  uint8_t* sV = sK + 2 * ATT_TILE_BYTES;
  uint8_t* sdO = sV + 2 * ATT_TILE_BYTES;
  uint8_t* sP = sdO + 2 * ATT_TILE_BYTES;           // 32 KB
  uint8_t* sdS = sP + 2 * ATT_TILE_BYTES;           // 32 KB
  float* sBias = reinterpret_cast<float*>(sdS + 2 * ATT_TILE_BYTES);   // [256]
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(sBias + 2 * ATT_T);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = (warp & 3) * 32 + lane;   // row of a 128-row tile this thread owns (TMEM lane)
  const int half = warp >> 2;               // which 64 of a tile's 128 key columns / which accumulator it drains
  if (tid == 0) {
    tma_prefetch_desc(&tma_qkv);
    tma_prefetch_desc(&tma_do);
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_S = tmem, t_dP = tmem + 128, t_dK = tmem + 256, t_dV = tmem + 320, t_dQ = tmem + 384;  // dQ_i at +64 i
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
  const uint32_t idesc_tt = umma_idesc_bf16(128, 64, 1, 1);
  const uint32_t idesc_nt = umma_idesc_bf16(128, 64, 0, 1);
  const bool use_drop = p.p_drop > 0.f;
  constexpr int s_pad = 2 * ATT_T;

  auto issue_scores = [&](int i, int j) {   // S = Q_i K_j^T, dP = dO_i V_j^T   (single thread)
    const uint32_t q = smem_u32(sQ + i * ATT_TILE_BYTES), k = smem_u32(sK + j * ATT_TILE_BYTES);
    const uint32_t g = smem_u32(sdO + i * ATT_TILE_BYTES), v = smem_u32(sV + j * ATT_TILE_BYTES);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
      umma_bf16(t_S, umma_desc_sw128(q + kk * 32, 16, 1024), umma_desc_sw128(k + kk * 32, 16, 1024), idesc_s, kk > 0);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
      umma_bf16(t_dP, umma_desc_sw128(g + kk * 32, 16, 1024), umma_desc_sw128(v + kk * 32, 16, 1024), idesc_s, kk > 0);
    umma_commit(bar_mma);
  };

  uint32_t ph_load = 0, ph_mma = 0;
  const int items = p.B * p.H;
  for (int bh = blockIdx.x; bh < items; bh += gridDim.x) {
    const int b = bh / p.H, h = bh - b * p.H;
    tc_fence_before_sync();
    __syncthreads();   // previous head: all shared-memory / TMEM readers are done
    if (tid == 0) {
      mbar_expect_tx(bar_load, 8 * ATT_TILE_BYTES);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        tma_load_3d(sQ + t * ATT_TILE_BYTES, &tma_qkv, bar_load, h * ATT_D, t * ATT_T, b);
        tma_load_3d(sK + t * ATT_TILE_BYTES, &tma_qkv, bar_load, p.D + h * ATT_D, t * ATT_T, b);
        tma_load_3d(sV + t * ATT_TILE_BYTES, &tma_qkv, bar_load, 2 * p.D + h * ATT_D, t * ATT_T, b);
        tma_load_3d(sdO + t * ATT_TILE_BYTES, &tma_do, bar_load, h * ATT_D, t * ATT_T, b);
      }
    }
    sBias[tid] = tid < p.S ? (p.key_bias ? p.key_bias[b * p.S + tid] * LOG2E : 0.f) : -INFINITY;
    // per-row scalars of this thread's row in both query tiles: delta = sum_d dO o O, log2-domain LSE
    float delta[2] = {0.f, 0.f}, lse_l2[2] = {INFINITY, INFINITY};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int qrow = i * ATT_T + row;
      if (qrow < p.S) {
        const long long off = (static_cast<long long>(b) * p.S + qrow) * p.D + h * ATT_D;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float a[8], g[8];
          load8(p.o_in + off + q * 8, a);
          load8(p.do_in + off + q * 8, g);
#pragma unroll
          for (int e = 0; e < 8; ++e) delta[i] = fmaf(a[e], g[e], delta[i]);
        }
        lse_l2[i] = p.lse[static_cast<long long>(bh) * p.S + qrow] * LOG2E;
      }
    }
    __syncthreads();   // sBias visible
    if (tid == 0) {
      mbar_wait(bar_load, ph_load);
      tc_fence_after_sync();
      issue_scores(0, 0);
    }
    ph_load ^= 1;

#pragma unroll 1
    for (int pair = 0; pair < 4; ++pair) {
      const int j = pair >> 1, i = pair & 1;     // key tile outer, query tile inner
      mbar_wait(bar_mma, ph_mma);                // S / dP of this pair are ready (and every earlier MMA has retired)
      ph_mma ^= 1;
      tc_fence_after_sync();
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c = half * 2 + cc;             // 32-column chunk of the 128 key columns
        uint32_t vs[32], vp[32];
        tmem_ld32(t_S + lane_addr + c * 32, vs);
        tmem_ld32(t_dP + lane_addr + c * 32, vp);
        tmem_ld_wait();
        float pd[32], ds[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          pd[e] = fast_exp2(fmaf(__uint_as_float(vs[e]), p.scale_log2, sBias[j * ATT_T + c * 32 + e]) - lse_l2[i]);
          ds[e] = __uint_as_float(vp[e]);
        }
        if (use_drop) {
          const uint32_t keep = dropout_keep32(p.seed, drop_chunk(bh, s_pad, i * ATT_T + row, j * ATT_T + c * 32),
                                               p.drop_threshold >> 16);
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float m = (keep >> e) & 1 ? p.inv_keep : 0.f;
            const float prob = pd[e];
            pd[e] = prob * m;
            ds[e] = prob * (ds[e] * m - delta[i]) * p.scale;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) ds[e] = pd[e] * (ds[e] - delta[i]) * p.scale;
        }
        store_row32_sw128(sP, row, c * 32, pd);
        store_row32_sw128(sdS, row, c * 32, ds);
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncthreads();   // P / dS complete; S / dP fully read
      if (tid < 32 && elect_one()) {   // one thread of warp 0 issues (ptx.cuh: elect_one)
        tc_fence_after_sync();
        const uint32_t q = smem_u32(sQ + i * ATT_TILE_BYTES), k = smem_u32(sK + j * ATT_TILE_BYTES);
        const uint32_t g = smem_u32(sdO + i * ATT_TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)   // dV_j (+)= P^T dO_i   (first query tile overwrites)
          umma_bf16(t_dV, umma_desc_sw128(smem_u32(sP) + kk * 2048, ATT_TILE_BYTES, 1024),
                    umma_desc_sw128(g + kk * 2048, 8192, 1024), idesc_tt, (i > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)   // dK_j (+)= dS^T Q_i
          umma_bf16(t_dK, umma_desc_sw128(smem_u32(sdS) + kk * 2048, ATT_TILE_BYTES, 1024),
                    umma_desc_sw128(q + kk * 2048, 8192, 1024), idesc_tt, (i > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)   // dQ_i (+)= dS K_j     (first key tile overwrites)
          umma_bf16(t_dQ + i * 64,
                    umma_desc_sw128(smem_u32(sdS) + (kk >> 2) * ATT_TILE_BYTES + (kk & 3) * 32, 16, 1024),
                    umma_desc_sw128(k + kk * 2048, 8192, 1024), idesc_nt, (j > 0 || kk > 0) ? 1u : 0u);
        // the next pair's scores go in right behind (S / dP columns are free, the tensor core runs in order) --
        // except across a key-tile boundary, where dK_j / dV_j must first be drained by the epilogue below
        if (pair == 0 || pair == 2) issue_scores(1, j);
        else umma_commit(bar_mma);
      }
      if (i == 1) {
        // ---- key tile j finished: drain dK_j (half 0) / dV_j (half 1) -> dqkv rows j*128 + row
        mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after_sync();
        // staged (swizzled) in the idle P buffer, then two coalesced TMA stores (rows >= S clipped by the map)
        {
          const uint32_t t_src = half == 0 ? t_dK : t_dV;
          uint8_t* dst = sP + half * ATT_TILE_BYTES;
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld32(t_src + lane_addr + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 o;
              o.x = pack_bf16x2(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1]));
              o.y = pack_bf16x2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3]));
              o.z = pack_bf16x2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5]));
              o.w = pack_bf16x2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7]));
              *reinterpret_cast<uint4*>(dst + sw128_off(row, c * 4 + q)) = o;
            }
          }
          fence_proxy_async_smem();
          tc_fence_before_sync();
          __syncthreads();   // tiles staged; dK_j / dV_j drained by everybody (the tensor core may overwrite them)
          if (tid == 0) {
            tma_store_3d(&tma_dqkv, sP, p.D + h * ATT_D, j * ATT_T, b);
            tma_store_3d(&tma_dqkv, sP + ATT_TILE_BYTES, 2 * p.D + h * ATT_D, j * ATT_T, b);
            tma_store_commit();
            if (j == 0) {
              tma_store_wait_read<0>();   // P buffer free again before the next pair's scores (and its P) can exist
              tc_fence_after_sync();
              issue_scores(0, 1);
            }
          }
        }
      }
    }
    // ---- dQ_0 (half 0) / dQ_1 (half 1): the drain above already waited for the last MMA group; staged in dS's buffer
    {
      uint8_t* dst = sdS + half * ATT_TILE_BYTES;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(t_dQ + half * 64 + lane_addr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1]));
          o.y = pack_bf16x2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3]));
          o.z = pack_bf16x2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5]));
          o.w = pack_bf16x2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7]));
          *reinterpret_cast<uint4*>(dst + sw128_off(row, c * 4 + q)) = o;
        }
      }
      fence_proxy_async_smem();
      __syncthreads();
      if (tid == 0) {
        tma_store_3d(&tma_dqkv, sdS, h * ATT_D, 0, b);
        tma_store_3d(&tma_dqkv, sdS + ATT_TILE_BYTES, h * ATT_D, ATT_T, b);
        tma_store_commit();
        tma_store_wait_read<0>();   // P / dS buffers are rewritten by the next head
      }
    }
  }
  if (tid == 0) tma_store_wait_read<0>();

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------------------ 256 < S <= 384, backward
// Three 128-row tiles (ViT-L/14: 257 tokens).  Same whole-head-per-CTA scheme, but 3 dQ accumulators + dK_j + dV_j +
// S + dP would need 576 TMEM columns and P + dS next to 12 operand tiles 256 KB of shared memory, so each (i, j)
// pair runs in two tensor-core phases that SHARE storage:  S -> P (kept packed in registers, staged for dV) | then dP
// into S's columns and dS into P's buffer.   TMEM: S/dP 128 | dK_j dV_j 128 | dQ_0..2 192 = 448 columns;
// shared memory: 12 operand tiles + one 32 KB P/dS buffer = 224 KB.
constexpr int ATT_BWD3_SMEM = 14 * ATT_TILE_BYTES + 3 * ATT_T * 4 + 64 + 1024;
__global__ void __launch_bounds__(256, 1)
attn_bwd_tmem3_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_do,
                      const __grid_constant__ CUtensorMap tma_dqkv, const AttnParams p) {
  constexpr int NT = 3;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* sQ = smem;                                // NT tiles each
  uint8_t* sK = sQ + NT * ATT_TILE_BYTES;
  uint8_t* sV = sK + NT * ATT_TILE_BYTES;
  uint8_t* sdO = sV + NT * ATT_TILE_BYTES;
  uint8_t* sP = sdO + NT * ATT_TILE_BYTES;           // 32 KB: P, then dS (in place), then the staged dK_j | dV_j
  float* sBias = reinterpret_cast<float*>(sP + 2 * ATT_TILE_BYTES);   // [NT * 128]
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(sBias + NT * ATT_T);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = (warp & 3) * 32 + lane, half = warp >> 2;
  if (tid == 0) {
    tma_prefetch_desc(&tma_qkv);
    tma_prefetch_desc(&tma_do);
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_S = tmem, t_dK = tmem + 128, t_dV = tmem + 192, t_dQ = tmem + 256;   // dQ_i at + 64 i
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
  const uint32_t idesc_tt = umma_idesc_bf16(128, 64, 1, 1);
  const uint32_t idesc_nt = umma_idesc_bf16(128, 64, 0, 1);
  const bool use_drop = p.p_drop > 0.f;
  constexpr int s_pad = NT * ATT_T;

  auto issue_s = [&](int i, int j) {   // S = Q_i K_j^T (single thread; no commit)
    const uint32_t q = smem_u32(sQ + i * ATT_TILE_BYTES), k = smem_u32(sK + j * ATT_TILE_BYTES);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
      umma_bf16(t_S, umma_desc_sw128(q + kk * 32, 16, 1024), umma_desc_sw128(k + kk * 32, 16, 1024), idesc_s, kk > 0);
  };

  uint32_t ph_load = 0, ph_mma = 0;
  const int items = p.B * p.H;
  for (int bh = blockIdx.x; bh < items; bh += gridDim.x) {
    const int b = bh / p.H, h = bh - b * p.H;
    tc_fence_before_sync();
    __syncthreads();   // previous head: all shared-memory / TMEM readers are done
    if (tid == 0) {
      tma_store_wait_read<0>();   // its dQ tiles (staged in the Q region) have left shared memory
      mbar_expect_tx(bar_load, 4 * NT * ATT_TILE_BYTES);
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        tma_load_3d(sQ + t * ATT_TILE_BYTES, &tma_qkv, bar_load, h * ATT_D, t * ATT_T, b);
        tma_load_3d(sK + t * ATT_TILE_BYTES, &tma_qkv, bar_load, p.D + h * ATT_D, t * ATT_T, b);
        tma_load_3d(sV + t * ATT_TILE_BYTES, &tma_qkv, bar_load, 2 * p.D + h * ATT_D, t * ATT_T, b);
        tma_load_3d(sdO + t * ATT_TILE_BYTES, &tma_do, bar_load, h * ATT_D, t * ATT_T, b);
      }
    }
    for (int k = tid; k < NT * ATT_T; k += 256)
      sBias[k] = k < p.S ? (p.key_bias ? p.key_bias[b * p.S + k] * LOG2E : 0.f) : -INFINITY;
    float delta[NT], lse_l2[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      delta[i] = 0.f;
      lse_l2[i] = INFINITY;
      const int qrow = i * ATT_T + row;
      if (qrow < p.S) {
        const long long off = (static_cast<long long>(b) * p.S + qrow) * p.D + h * ATT_D;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float a[8], g[8];
          load8(p.o_in + off + q * 8, a);
          load8(p.do_in + off + q * 8, g);
#pragma unroll
          for (int e = 0; e < 8; ++e) delta[i] = fmaf(a[e], g[e], delta[i]);
        }
        lse_l2[i] = p.lse[static_cast<long long>(bh) * p.S + qrow] * LOG2E;
      }
    }
    __syncthreads();   // sBias visible
    if (tid == 0) {
      mbar_wait(bar_load, ph_load);
      tc_fence_after_sync();
      issue_s(0, 0);
      umma_commit(bar_mma);
    }
    ph_load ^= 1;

#pragma unroll 1
    for (int pair = 0; pair < NT * NT; ++pair) {
      const int j = pair / NT, i = pair - j * NT;   // key tile outer, query tile inner
      const float dl = i == 0 ? delta[0] : (i == 1 ? delta[1] : delta[2]);
      const float ll = i == 0 ? lse_l2[0] : (i == 1 ? lse_l2[1] : lse_l2[2]);
      // ---- phase 1: S ready -> P (packed copy stays in registers for phase 2), staged for dV
      mbar_wait(bar_mma, ph_mma);
      ph_mma ^= 1;
      tc_fence_after_sync();
      uint32_t prob_pk[2][16];   // undropped probabilities of this thread's 64 key columns, bf16x2
      uint32_t keep[2] = {0xffffffffu, 0xffffffffu};
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = half * 2 + cc;
        uint32_t vs[32];
        tmem_ld32(t_S + lane_addr + c * 32, vs);
        tmem_ld_wait();
        float pd[32];
#pragma unroll
        for (int e = 0; e < 32; ++e)
          pd[e] = fast_exp2(fmaf(__uint_as_float(vs[e]), p.scale_log2, sBias[j * ATT_T + c * 32 + e]) - ll);
#pragma unroll
        for (int e = 0; e < 16; ++e) prob_pk[cc][e] = pack_bf16x2(pd[2 * e], pd[2 * e + 1]);
        if (use_drop) {
          keep[cc] = dropout_keep32(p.seed, drop_chunk(bh, s_pad, i * ATT_T + row, j * ATT_T + c * 32),
                                    p.drop_threshold >> 16);
#pragma unroll
          for (int e = 0; e < 32; ++e) pd[e] = (keep[cc] >> e) & 1 ? pd[e] * p.inv_keep : 0.f;
        }
        store_row32_sw128(sP, row, c * 32, pd);
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncthreads();   // P staged; S fully read (its columns may take dP now)
      if (tid < 32 && elect_one()) {   // one thread of warp 0 issues (ptx.cuh: elect_one)
        tc_fence_after_sync();
        const uint32_t g = smem_u32(sdO + i * ATT_TILE_BYTES), v = smem_u32(sV + j * ATT_TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)   // dP = dO_i V_j^T into S's columns
          umma_bf16(t_S, umma_desc_sw128(g + kk * 32, 16, 1024), umma_desc_sw128(v + kk * 32, 16, 1024), idesc_s, kk > 0);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)   // dV_j (+)= P^T dO_i
          umma_bf16(t_dV, umma_desc_sw128(smem_u32(sP) + kk * 2048, ATT_TILE_BYTES, 1024),
                    umma_desc_sw128(g + kk * 2048, 8192, 1024), idesc_tt, (i > 0 || kk > 0) ? 1u : 0u);
        umma_commit(bar_mma);
      }
      // ---- phase 2: dP ready (and P consumed) -> dS in place of P
      mbar_wait(bar_mma, ph_mma);
      ph_mma ^= 1;
      tc_fence_after_sync();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = half * 2 + cc;
        uint32_t vp[32];
        tmem_ld32(t_S + lane_addr + c * 32, vp);
        tmem_ld_wait();
        float ds[32];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float2 pr = unpack_bf16x2(prob_pk[cc][e]);
          const float m0 = (keep[cc] >> (2 * e)) & 1 ? (use_drop ? p.inv_keep : 1.f) : 0.f;
          const float m1 = (keep[cc] >> (2 * e + 1)) & 1 ? (use_drop ? p.inv_keep : 1.f) : 0.f;
          ds[2 * e] = pr.x * (__uint_as_float(vp[2 * e]) * m0 - dl) * p.scale;
          ds[2 * e + 1] = pr.y * (__uint_as_float(vp[2 * e + 1]) * m1 - dl) * p.scale;
        }
        store_row32_sw128(sP, row, c * 32, ds);
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncthreads();   // dS staged; dP fully read
      if (tid < 32 && elect_one()) {   // one thread of warp 0 issues (ptx.cuh: elect_one)
        tc_fence_after_sync();
        const uint32_t q = smem_u32(sQ + i * ATT_TILE_BYTES), k = smem_u32(sK + j * ATT_TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)   // dK_j (+)= dS^T Q_i
          umma_bf16(t_dK, umma_desc_sw128(smem_u32(sP) + kk * 2048, ATT_TILE_BYTES, 1024),
                    umma_desc_sw128(q + kk * 2048, 8192, 1024), idesc_tt, (i > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)   // dQ_i (+)= dS K_j
          umma_bf16(t_dQ + i * 64,
                    umma_desc_sw128(smem_u32(sP) + (kk >> 2) * ATT_TILE_BYTES + (kk & 3) * 32, 16, 1024),
                    umma_desc_sw128(k + kk * 2048, 8192, 1024), idesc_nt, (j > 0 || kk > 0) ? 1u : 0u);
        // next pair's scores right behind (S / dP columns are free; in-order tensor pipe)
        if (pair + 1 < NT * NT) issue_s((pair + 1) % NT, (pair + 1) / NT);
        umma_commit(bar_mma);
      }
      if (i == NT - 1) {
        // ---- key tile j finished: drain dK_j (half 0) / dV_j (half 1) through the P buffer
        mbar_wait(bar_mma, ph_mma);     // (also the next pair's S; NOT consumed twice: see below)
        tc_fence_after_sync();
        const uint32_t t_src = half == 0 ? t_dK : t_dV;
        uint8_t* dst = sP + half * ATT_TILE_BYTES;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(t_src + lane_addr + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1]));
            o.y = pack_bf16x2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3]));
            o.z = pack_bf16x2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5]));
            o.w = pack_bf16x2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7]));
            *reinterpret_cast<uint4*>(dst + sw128_off(row, c * 4 + q)) = o;
          }
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        if (tid == 0) {
          tma_store_3d(&tma_dqkv, sP, p.D + h * ATT_D, j * ATT_T, b);
          tma_store_3d(&tma_dqkv, sP + ATT_TILE_BYTES, 2 * p.D + h * ATT_D, j * ATT_T, b);
          tma_store_commit();
          tma_store_wait_read<0>();   // the P buffer is rewritten by the next pair's phase 1
        }
        __syncthreads();
      }
    }
    // ---- dQ_0..2: staged in the (now idle) Q tiles, three TMA stores; the last drain waited for every MMA
    ph_mma ^= 1;   // consume the phase observed (not toggled) by the last drain's wait
#pragma unroll 1
    for (int t = 0; t < NT; ++t) {
      uint32_t v[32];
      tmem_ld32(t_dQ + t * 64 + lane_addr + half * 32, v);
      tmem_ld_wait();
      uint8_t* dst = sQ + t * ATT_TILE_BYTES;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1]));
        o.y = pack_bf16x2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3]));
        o.z = pack_bf16x2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5]));
        o.w = pack_bf16x2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7]));
        *reinterpret_cast<uint4*>(dst + sw128_off(row, half * 4 + q)) = o;
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int t = 0; t < NT; ++t) tma_store_3d(&tma_dqkv, sQ + t * ATT_TILE_BYTES, h * ATT_D, t * ATT_T, b);
      tma_store_commit();
    }
  }
  if (tid == 0) tma_store_wait_read<0>();

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------------------ S <= 128, backward (v2)
// One-tile sequences (the text towers at 128 tokens): recompute S and dP = dO V^T on the tensor core, P = exp(S - LSE),
// dS = P o (dP - delta), then dV = P^T dO, dK = dS^T Q, dQ = dS K -- the transposed operands are the SAME smem tiles
// read through MN-major UMMA descriptors.  256 threads so that two warps per TMEM lane quarter split the key columns of
// P / dS and the three output tiles; a two-stage shared-memory ring so the next head's Q / K / V / dO arrive while the
// current head computes; the next head's score MMAs are issued right behind this head's gradient MMAs; gradient tiles
// leave through swizzled staging + TMA stores.  (The first version -- 128 threads, single-buffered, per-thread row
// stores -- exposed every TMA / MMA latency: 199 us at B=256, H=12, S=128; this one: 150 us.)
constexpr int ATT_BWD1_SMEM = 12 * ATT_TILE_BYTES + 2 * ATT_T * 4 + 64 + 1024;
__global__ void __launch_bounds__(256, 1)
attn_bwd1_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_do,
                 const __grid_constant__ CUtensorMap tma_dqkv, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* sIn = smem;                              // [2 stages][Q | K | V | dO], 64 KB per stage
  uint8_t* sP = sIn + 8 * ATT_TILE_BYTES;           // 32 KB
  uint8_t* sdS = sP + 2 * ATT_TILE_BYTES;           // 32 KB
  float* sBias = reinterpret_cast<float*>(sdS + 2 * ATT_TILE_BYTES);   // [2][128]
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(sBias + 2 * ATT_T); // [2]
  uint64_t* bar_s = bar_load + 2;      // scores (S, dP) of a head are in TMEM
  uint64_t* bar_g = bar_s + 1;         // gradient MMAs of a head have retired
  // (two barriers: the scores of head n+1 are committed right behind the gradients of head n, and a single barrier
  //  could complete two phases before a slow thread has observed the first)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_g + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = (warp & 3) * 32 + lane, half = warp >> 2;
  if (tid == 0) {
    tma_prefetch_desc(&tma_qkv);
    tma_prefetch_desc(&tma_do);
    mbar_init(&bar_load[0], 1);
    mbar_init(&bar_load[1], 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_g, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_S = tmem, t_dP = tmem + 128, t_dK = tmem + 256, t_dV = tmem + 320, t_dQ = tmem + 384;
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
  const uint32_t idesc_tt = umma_idesc_bf16(128, 64, 1, 1);
  const uint32_t idesc_nt = umma_idesc_bf16(128, 64, 0, 1);
  const bool use_drop = p.p_drop > 0.f;
  const int items = p.B * p.H;

  auto issue_loads = [&](int item, int stage) {   // single thread
    const int b = item / p.H, h = item - b * p.H;
    uint8_t* base = sIn + stage * 4 * ATT_TILE_BYTES;
    mbar_expect_tx(&bar_load[stage], 4 * ATT_TILE_BYTES);
    tma_load_3d(base, &tma_qkv, &bar_load[stage], h * ATT_D, 0, b);
    tma_load_3d(base + ATT_TILE_BYTES, &tma_qkv, &bar_load[stage], p.D + h * ATT_D, 0, b);
    tma_load_3d(base + 2 * ATT_TILE_BYTES, &tma_qkv, &bar_load[stage], 2 * p.D + h * ATT_D, 0, b);
    tma_load_3d(base + 3 * ATT_TILE_BYTES, &tma_do, &bar_load[stage], h * ATT_D, 0, b);
  };

  auto issue_scores = [&](int stage) {            // single thread: S = Q K^T, dP = dO V^T
    const uint32_t q = smem_u32(sIn + stage * 4 * ATT_TILE_BYTES), k = q + ATT_TILE_BYTES, v = k + ATT_TILE_BYTES,
                   g = v + ATT_TILE_BYTES;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
      umma_bf16(t_S, umma_desc_sw128(q + kk * 32, 16, 1024), umma_desc_sw128(k + kk * 32, 16, 1024), idesc_s, kk > 0);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
      umma_bf16(t_dP, umma_desc_sw128(g + kk * 32, 16, 1024), umma_desc_sw128(v + kk * 32, 16, 1024), idesc_s, kk > 0);
    umma_commit(bar_s);
  };
  // per-row inputs of delta = rowsum(dO o O) and the LSE, fetched one head AHEAD (they sit on the critical path
  // otherwise: 25 % of the stall samples of the first version)
  uint4 ro[8], rg[8];
  float lse_raw = INFINITY;
  auto prefetch_rows = [&](int item) {
    const int b = item / p.H, h = item - b * p.H;
    if (row < p.S) {
      const long long off = (static_cast<long long>(b) * p.S + row) * p.D + h * ATT_D;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        ro[q] = __ldg(reinterpret_cast<const uint4*>(p.o_in + off + q * 8));
        rg[q] = __ldg(reinterpret_cast<const uint4*>(p.do_in + off + q * 8));
      }
      lse_raw = __ldg(p.lse + static_cast<long long>(item) * p.S + row);
    }
  };
  auto write_bias = [&](int item, int stage) {
    const int b = item / p.H;
    if (tid < ATT_T)
      sBias[stage * ATT_T + tid] = tid < p.S ? (p.key_bias ? p.key_bias[b * p.S + tid] * LOG2E : 0.f) : -INFINITY;
  };

  uint32_t ph_s = 0, ph_g = 0;
  if (static_cast<int>(blockIdx.x) < items) {
    if (tid == 0) {
      issue_loads(blockIdx.x, 0);
      mbar_wait(&bar_load[0], 0);
      tc_fence_after_sync();
      issue_scores(0);
    }
    prefetch_rows(blockIdx.x);
    write_bias(blockIdx.x, 0);
  }
  int n = 0;
  for (int item = blockIdx.x; item < items; item += gridDim.x, ++n) {
    const int stage = n & 1;
    const int next = item + static_cast<int>(gridDim.x);
    const int b = item / p.H, h = item - b * p.H;
    uint8_t* sQ = sIn + stage * 4 * ATT_TILE_BYTES;
    uint8_t* sK = sQ + ATT_TILE_BYTES;
    uint8_t* sdO = sK + 2 * ATT_TILE_BYTES;
    const float* bias = sBias + stage * ATT_T;
    // the other stage was last read by the previous head's MMAs, all retired (bar_g waited): refill it now
    if (tid == 0) {
      if (next < items) issue_loads(next, stage ^ 1);
      tma_store_wait_read<0>();   // the previous head's gradient tiles have left sP / sdS (before the sync below)
    }
    const bool row_ok = row < p.S;
    float delta = 0.f, lse_l2 = INFINITY;
    if (row_ok) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float a[8], g[8];
        unpack8(ro[q], a);
        unpack8(rg[q], g);
#pragma unroll
        for (int e = 0; e < 8; ++e) delta = fmaf(a[e], g[e], delta);
      }
      lse_l2 = lse_raw * LOG2E;
    }
    __syncthreads();   // this head's bias (written one iteration ago) is visible
    mbar_wait(bar_s, ph_s);   // S / dP of this head (issued one iteration ago)
    ph_s ^= 1;
    tc_fence_after_sync();
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      const int c = half * 2 + cc;
      uint32_t vs[32], vp[32];
      tmem_ld32(t_S + lane_addr + c * 32, vs);
      tmem_ld32(t_dP + lane_addr + c * 32, vp);
      tmem_ld_wait();
      float pd[32], ds[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        pd[e] = fast_exp2(fmaf(__uint_as_float(vs[e]), p.scale_log2, bias[c * 32 + e]) - lse_l2);
        ds[e] = __uint_as_float(vp[e]);
      }
      if (use_drop) {
        const uint32_t keep = dropout_keep32(p.seed, (static_cast<uint64_t>(item) * ATT_T + row) * 4 + c,
                                             p.drop_threshold >> 16);   // as the forward
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float m = (keep >> e) & 1 ? p.inv_keep : 0.f;
          const float prob = pd[e];
          pd[e] = prob * m;
          ds[e] = prob * (ds[e] * m - delta) * p.scale;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) ds[e] = pd[e] * (ds[e] - delta) * p.scale;
      }
      store_row32_sw128(sP, row, c * 32, pd);
      store_row32_sw128(sdS, row, c * 32, ds);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();   // P / dS complete, S / dP fully read
    if (tid < 32 && elect_one()) {   // one thread of warp 0 issues (ptx.cuh: elect_one)
      tc_fence_after_sync();
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_bf16(t_dV, umma_desc_sw128(smem_u32(sP) + k * 2048, ATT_TILE_BYTES, 1024),
                  umma_desc_sw128(smem_u32(sdO) + k * 2048, 8192, 1024), idesc_tt, k > 0);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_bf16(t_dK, umma_desc_sw128(smem_u32(sdS) + k * 2048, ATT_TILE_BYTES, 1024),
                  umma_desc_sw128(smem_u32(sQ) + k * 2048, 8192, 1024), idesc_tt, k > 0);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_bf16(t_dQ, umma_desc_sw128(smem_u32(sdS) + (k >> 2) * ATT_TILE_BYTES + (k & 3) * 32, 16, 1024),
                  umma_desc_sw128(smem_u32(sK) + k * 2048, 8192, 1024), idesc_nt, k > 0);
      umma_commit(bar_g);
      // the next head's scores go in right behind (in-order tensor pipe; S / dP are free, its tiles were prefetched)
      if (next < items) {
        mbar_wait(&bar_load[stage ^ 1], ((n + 1) >> 1) & 1);
        tc_fence_after_sync();
        issue_scores(stage ^ 1);
      }
    }
    if (next < items) {   // next head's per-row inputs and key bias travel while the gradient MMAs run
      prefetch_rows(next);
      write_bias(next, stage ^ 1);
    }
    mbar_wait(bar_g, ph_g);   // dV / dK / dQ of this head
    ph_g ^= 1;
    tc_fence_after_sync();
    // ---- drain through shared memory: dQ | dK | dV tiles [128 x 64] are staged (swizzled) in the now idle P / dS
    // buffers and leave with three TMA stores -- coalesced, rows >= S clipped by the tensor map.  (Per-thread row
    // stores were 32 scattered 16-byte sectors per instruction: 15 % of the stall samples of the first version.)
    // half 0 -> dQ (64 cols) + dK cols 0..31 ; half 1 -> dV (64 cols) + dK cols 32..63
    {
      uint8_t* stage_out = sP;    // tiles: 0 = dQ, 1 = dK, 2 = dV (16 KB each; sP and sdS are contiguous)
#pragma unroll 1
      for (int piece = 0; piece < 3; ++piece) {
        const uint32_t t_src = piece < 2 ? ((half == 0 ? t_dQ : t_dV) + piece * 32) : (t_dK + half * 32);
        const int tile = piece < 2 ? (half == 0 ? 0 : 2) : 1;
        const int chunk0 = (piece < 2 ? piece : half) * 4;      // first 16-byte chunk of the 32 columns
        uint32_t v[32];
        tmem_ld32(t_src + lane_addr, v);
        tmem_ld_wait();
        uint8_t* dst = stage_out + tile * ATT_TILE_BYTES;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1]));
          o.y = pack_bf16x2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3]));
          o.z = pack_bf16x2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5]));
          o.w = pack_bf16x2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7]));
          *reinterpret_cast<uint4*>(dst + sw128_off(row, chunk0 + q)) = o;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncthreads();
      if (tid == 0) {
#pragma unroll
        for (int t = 0; t < 3; ++t)
          tma_store_3d(&tma_dqkv, stage_out + t * ATT_TILE_BYTES, t * p.D + h * ATT_D, 0, b);
        tma_store_commit();
      }
    }
  }
  if (tid == 0) tma_store_wait_read<0>();   // shared memory must outlive the last bulk store

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem);
  }
}

template <int NT>
constexpr int att_fwd_tmem_smem() { return (2 + 2 * NT) * ATT_TILE_BYTES + (NT + 2) * ATT_T * 4 + 64 + 1024; }

constexpr int ATT_FWD_MULTI_SMEM = 5 * ATT_TILE_BYTES + ATT_MAX_S * 4 + 64 + 1024;

constexpr int ATT_BWD_SMEM = 8 * ATT_TILE_BYTES + ATT_T * 4 + 64 + 1024;

static int fill_params(AttnParams& p, int B, int H, int S, float p_drop, unsigned long long seed,
                       const float* key_bias) {
  if (B <= 0 || H <= 0 || S <= 0 || S > 512 || p_drop < 0.f || p_drop >= 1.f) return B200MM_ERR_BAD_ARG;
  p.B = B; p.H = H; p.S = S; p.D = H * ATT_D;
  p.scale = 0.125f;  // 64^-1/2
  p.scale_log2 = p.scale * LOG2E;
  p.p_drop = p_drop;
  p.drop_threshold = dropout_threshold(p_drop);
  p.inv_keep = 1.f / (1.f - p_drop);
  p.seed = seed;
  p.key_bias = key_bias;
  return B200MM_OK;
}

}  // namespace b200

using namespace b200;

// O[B*S, H*64] = softmax(Q K^T / 8 + key_bias) (dropout) V, Q/K/V = column blocks of qkv [B*S, 3*H*64].
// lse [B,H,S] (fp32) is written when non-null (required for the backward).  S <= 512 (multi-tile kernels above 128).
static int attention_fwd_impl(const void* qkv, const float* key_bias, void* out, float* lse, int B, int H, int S,
                              float p_drop, unsigned long long seed, void* drop_mask, void* stream) {
  const DeviceInfo& dev = device_info();
  if (!dev.ok || dev.cc_major != 10) return B200MM_ERR_NOT_SM100;
  AttnParams p{};
  int rc = fill_params(p, B, H, S, p_drop, seed, key_bias);
  if (rc) return rc;
  // the keep-bit exchange exists in the one-tile warp-specialised kernels only (forward writes, backward reads)
  p.drop_mask = (S <= ATT_T && attn_ws_enabled() && p_drop > 0.f) ? static_cast<uint32_t*>(drop_mask) : nullptr;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  CUtensorMap tq;
  const uint64_t row = static_cast<uint64_t>(3) * p.D * 2;
  rc = make_tmap_3d_bf16(&tq, qkv, 3 * p.D, S, B, row, row * S, ATT_D, ATT_T);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_FWD_MULTI_SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  if (S <= 3 * ATT_T) {      // up to three key tiles: all score tiles stay in TMEM, one pass
    const int nt = ceil_div(S, ATT_T);
    const int items = B * H * nt;
    static bool configured_tmem = false;
    if (!configured_tmem) {
      cudaError_t e = cudaFuncSetAttribute(attn_fwd_tmem_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           att_fwd_tmem_smem<2>());
      if (e != cudaSuccess) return static_cast<int>(e);
      e = cudaFuncSetAttribute(attn_fwd_tmem_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               att_fwd_tmem_smem<1>());
      if (e != cudaSuccess) return static_cast<int>(e);
      e = cudaFuncSetAttribute(attn_fwd_tmem_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               att_fwd_tmem_smem<3>());
      if (e != cudaSuccess) return static_cast<int>(e);
      configured_tmem = true;
    }
    CUtensorMap to;
    const uint64_t orow = static_cast<uint64_t>(p.D) * 2;
    rc = make_tmap_3d_bf16(&to, out, p.D, S, B, orow, orow * S, ATT_D, ATT_T);
    if (rc) return rc;
    if (nt == 1 && attn_ws_enabled())      // warp-specialised pipeline, four heads in flight per SM (attention_ws.cu)
      return launch_attn_fwd_ws(tq, to, p, dev.num_sms, static_cast<cudaStream_t>(stream));
    if (nt == 1) {
      const int grid = items < 2 * dev.num_sms ? items : 2 * dev.num_sms;
      attn_fwd_tmem_kernel<1><<<grid, 256, att_fwd_tmem_smem<1>(), static_cast<cudaStream_t>(stream)>>>(tq, to, p);
    } else if (nt == 2) {
      const int grid = items < 2 * dev.num_sms ? items : 2 * dev.num_sms;
      attn_fwd_tmem_kernel<2><<<grid, 256, att_fwd_tmem_smem<2>(), static_cast<cudaStream_t>(stream)>>>(tq, to, p);
    } else {
      const int grid = items < dev.num_sms ? items : dev.num_sms;
      attn_fwd_tmem_kernel<3><<<grid, 256, att_fwd_tmem_smem<3>(), static_cast<cudaStream_t>(stream)>>>(tq, to, p);
    }
    B200MM_CHECK_LAUNCH();
    return B200MM_OK;
  }
  if (S > ATT_T) {
    const int items = B * H * ceil_div(S, ATT_T);
    const int grid = items < 2 * dev.num_sms ? items : 2 * dev.num_sms;
    attn_fwd_multi_kernel<<<grid, 128, ATT_FWD_MULTI_SMEM, static_cast<cudaStream_t>(stream)>>>(tq, p);
    B200MM_CHECK_LAUNCH();
    return B200MM_OK;
  }
  return B200MM_ERR_BAD_ARG;   // unreachable: every S <= 512 is served above
}

B200MM_API int b200mm_attention_fwd(const void* qkv, const float* key_bias, void* out, float* lse, int B, int H, int S,
                                    float p_drop, unsigned long long seed, void* stream) {
  return attention_fwd_impl(qkv, key_bias, out, lse, B, H, S, p_drop, seed, nullptr, stream);
}
// Same, and the keep bits of the attention dropout are written to drop_mask ([B*H][4][128] uint32, S <= 128 only;
// ignored for longer sequences) for b200mm_attention_bwd_mask.
B200MM_API int b200mm_attention_fwd_mask(const void* qkv, const float* key_bias, void* out, float* lse, int B, int H,
                                         int S, float p_drop, unsigned long long seed, void* drop_mask, void* stream) {
  return attention_fwd_impl(qkv, key_bias, out, lse, B, H, S, p_drop, seed, drop_mask, stream);
}

// dqkv [B*S, 3*H*64] <- gradients of Q, K, V given dO, using O and the saved LSE.
static int attention_bwd_impl(const void* qkv, const float* key_bias, const void* out, const void* dout,
                              const float* lse, void* dqkv, int B, int H, int S, float p_drop,
                              unsigned long long seed, const void* drop_mask, void* stream) {
  const DeviceInfo& dev = device_info();
  if (!dev.ok || dev.cc_major != 10) return B200MM_ERR_NOT_SM100;
  AttnParams p{};
  int rc = fill_params(p, B, H, S, p_drop, seed, key_bias);
  if (rc) return rc;
  p.drop_mask = (S <= ATT_T && attn_ws_enabled() && p_drop > 0.f)
                    ? const_cast<uint32_t*>(static_cast<const uint32_t*>(drop_mask)) : nullptr;
  if (lse == nullptr) return B200MM_ERR_BAD_ARG;
  p.lse = const_cast<float*>(lse);
  p.o_in = static_cast<const __nv_bfloat16*>(out);
  p.do_in = static_cast<const __nv_bfloat16*>(dout);
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  CUtensorMap tq, td;
  const uint64_t row = static_cast<uint64_t>(3) * p.D * 2;
  rc = make_tmap_3d_bf16(&tq, qkv, 3 * p.D, S, B, row, row * S, ATT_D, ATT_T);
  if (rc) return rc;
  const uint64_t orow = static_cast<uint64_t>(p.D) * 2;
  rc = make_tmap_3d_bf16(&td, dout, p.D, S, B, orow, orow * S, ATT_D, ATT_T);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_BWD_SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  if (S > ATT_T && S <= 2 * ATT_T) {     // whole head per CTA, every accumulator resident in TMEM
    static bool configured2 = false;
    if (!configured2) {
      cudaError_t e = cudaFuncSetAttribute(attn_bwd_tmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           ATT_BWD2_SMEM);
      if (e != cudaSuccess) return static_cast<int>(e);
      configured2 = true;
    }
    const int items = B * H;
    const int grid = items < dev.num_sms ? items : dev.num_sms;
    CUtensorMap tdq2;
    rc = make_tmap_3d_bf16(&tdq2, dqkv, 3 * p.D, S, B, row, row * S, ATT_D, ATT_T);
    if (rc) return rc;
    attn_bwd_tmem_kernel<<<grid, ATT_BWD2_THREADS, ATT_BWD2_SMEM, static_cast<cudaStream_t>(stream)>>>(tq, td, tdq2, p);
    B200MM_CHECK_LAUNCH();
    return B200MM_OK;
  }
  if (S > 2 * ATT_T && S <= 3 * ATT_T) {     // three tiles (ViT-L/14): whole head per CTA, two-phase pairs
    static bool configured3 = false;
    if (!configured3) {
      cudaError_t e = cudaFuncSetAttribute(attn_bwd_tmem3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           ATT_BWD3_SMEM);
      if (e != cudaSuccess) return static_cast<int>(e);
      configured3 = true;
    }
    CUtensorMap tdq3;
    rc = make_tmap_3d_bf16(&tdq3, dqkv, 3 * p.D, S, B, row, row * S, ATT_D, ATT_T);
    if (rc) return rc;
    const int items = B * H;
    const int grid = items < dev.num_sms ? items : dev.num_sms;
    attn_bwd_tmem3_kernel<<<grid, 256, ATT_BWD3_SMEM, static_cast<cudaStream_t>(stream)>>>(tq, td, tdq3, p);
    B200MM_CHECK_LAUNCH();
    return B200MM_OK;
  }
  if (S > ATT_T) {
    const int items = 2 * B * H * ceil_div(S, ATT_T);
    const int grid = items < dev.num_sms ? items : dev.num_sms;
    attn_bwd_multi_kernel<<<grid, 128, ATT_BWD_SMEM, static_cast<cudaStream_t>(stream)>>>(tq, td, p);
    B200MM_CHECK_LAUNCH();
    return B200MM_OK;
  }
  const int items = B * H;
  const int grid = items < dev.num_sms ? items : dev.num_sms;
  static bool configured1 = false;
  if (!configured1) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_BWD1_SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured1 = true;
  }
  CUtensorMap tdq;
  rc = make_tmap_3d_bf16(&tdq, dqkv, 3 * p.D, S, B, row, row * S, ATT_D, ATT_T);
  if (rc) return rc;
  if (attn_ws_enabled())                   // warp-specialised pipeline, two heads in flight per SM (attention_ws.cu)
    return launch_attn_bwd_ws(tq, td, tdq, p, dev.num_sms, static_cast<cudaStream_t>(stream));
  attn_bwd1_kernel<<<grid, 256, ATT_BWD1_SMEM, static_cast<cudaStream_t>(stream)>>>(tq, td, tdq, p);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

B200MM_API int b200mm_attention_bwd(const void* qkv, const float* key_bias, const void* out, const void* dout,
                                    const float* lse, void* dqkv, int B, int H, int S, float p_drop,
                                    unsigned long long seed, void* stream) {
  return attention_bwd_impl(qkv, key_bias, out, dout, lse, dqkv, B, H, S, p_drop, seed, nullptr, stream);
}
// Same, reading the dropout keep bits b200mm_attention_fwd_mask saved (S <= 128; nullptr or longer: regenerated).
B200MM_API int b200mm_attention_bwd_mask(const void* qkv, const float* key_bias, const void* out, const void* dout,
                                         const float* lse, void* dqkv, int B, int H, int S, float p_drop,
                                         unsigned long long seed, const void* drop_mask, void* stream) {
  return attention_bwd_impl(qkv, key_bias, out, dout, lse, dqkv, B, H, S, p_drop, seed, drop_mask, stream);
}
