// Declarations shared by the attention kernels (attention_tcgen05.cu: multi-tile / TMEM-resident kernels;
// attention_ws.cu: the warp-specialised one-tile kernels).
#pragma once
#include <cstdlib>
#include "common.cuh"
#include "device_utils.cuh"
#include "ptx.cuh"

namespace b200 {

constexpr int ATT_T = 128;   // query / key tile (== max sequence length of this kernel)
constexpr int ATT_D = 64;    // head dim
constexpr int ATT_TILE_BYTES = ATT_T * ATT_D * 2;  // 16 KB: one [128 x 64] bf16 tile, 128B rows
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

struct AttnParams {
  int B, H, S, D;  // D = H * 64
  float scale_log2;      // head_dim^-1/2 * log2(e)
  float scale;           // head_dim^-1/2
  float p_drop;
  uint32_t drop_threshold;
  float inv_keep;
  unsigned long long seed;
  const float* key_bias;  // [B, S] additive bias (0 / -inf for padded keys) or nullptr
  __nv_bfloat16* out;     // fwd: O [B*S, D]
  float* lse;             // [B, H, S] natural-log LSE of the scaled+biased scores
  const __nv_bfloat16* o_in;   // bwd: O
  const __nv_bfloat16* do_in;  // bwd: dO [B*S, D]
  __nv_bfloat16* dqkv;         // bwd: [B*S, 3D]
  // one-tile (warp-specialised) kernels: keep bits of the attention dropout, [B*H][4][128] words (word c of row r =
  // keys 32c .. 32c+31, bit i = key 32c+i kept).  The forward writes them when non-null; the backward reads them
  // instead of regenerating the Philox stream (which cost more instructions than the softmax gradient itself).
  uint32_t* drop_mask;
  // debugging aid (B200MM_ATTN_TRACE=1): CTA 0 records %globaltimer at the pipeline's hand-over points,
  // [head][event] with 8 events per head; nullptr in normal operation
  unsigned long long* trace;
};

__device__ __forceinline__ void trace_event(const AttnParams& p, int head, int ev) {
  if (p.trace != nullptr && blockIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[head * 8 + ev] = t;
  }
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// exp2 on the special-function unit (inputs are <= 0 here; ex2.approx maps -inf to +0)
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// byte offset of 16-byte chunk `chunk` (0..7) of row `r` inside a [rows x 64] bf16 SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_off(int r, int chunk) { return r * 128 + ((chunk ^ (r & 7)) << 4); }

// write 32 consecutive bf16 of row r (columns c0..c0+31, c0 % 32 == 0) into a [128 x 128] tile stored as two
// [128 x 64] swizzled blocks
__device__ __forceinline__ void store_row32_sw128(uint8_t* tile, int r, int c0, const float (&x)[32]) {
  uint8_t* blk = tile + (c0 >> 6) * ATT_TILE_BYTES;
  const int chunk0 = (c0 & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 o;
    o.x = pack_bf16x2(x[q * 8 + 0], x[q * 8 + 1]);
    o.y = pack_bf16x2(x[q * 8 + 2], x[q * 8 + 3]);
    o.z = pack_bf16x2(x[q * 8 + 4], x[q * 8 + 5]);
    o.w = pack_bf16x2(x[q * 8 + 6], x[q * 8 + 7]);
    *reinterpret_cast<uint4*>(blk + sw128_off(r, chunk0 + q)) = o;
  }
}

// index of the 32-key chunk starting at `key` (multiple of 32) of query row `qrow`, for dropout_keep32
__device__ __forceinline__ uint64_t drop_chunk(int bh, int s_pad, int qrow, int key) {
  return ((static_cast<uint64_t>(bh) * s_pad + qrow) * s_pad + key) >> 5;
}

// Warp-specialised kernels for one-tile sequences (S <= 128), attention_ws.cu.  Return B200MM_OK or an error.
int launch_attn_fwd_ws(const CUtensorMap& tma_qkv, const CUtensorMap& tma_out, const AttnParams& p, int num_sms,
                       cudaStream_t stream);
int launch_attn_bwd_ws(const CUtensorMap& tma_qkv, const CUtensorMap& tma_do, const CUtensorMap& tma_dqkv,
                       const AttnParams& p, int num_sms, cudaStream_t stream);
// B200MM_ATTN_WS=0 selects the previous (single-role) kernels, for A/B measurements and cross-checks
bool attn_ws_enabled();

}  // namespace b200
