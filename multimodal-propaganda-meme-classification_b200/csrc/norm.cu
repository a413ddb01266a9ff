// LayerNorm / embedding kernels of the text tower (HBM-bound; one warp per token row, 16-byte accesses,
// warp-shuffle reductions, fp32 statistics).
//
// Replaces:
//   transformers/models/distilbert/modeling_distilbert.py:96-122  Embeddings (word + position -> LN -> dropout)
//   ...:257, :261  sa_layer_norm / output_layer_norm  (the residual add is fused into the producing GEMM epilogue)
// and their autograd backward passes.
#include "common.cuh"
#include "device_utils.cuh"
#include "ptx.cuh"

namespace b200 {

constexpr int LN_WARPS = 8;
constexpr int LN_MAX_CHUNKS = 8;  // per lane: 8 chunks x 8 elements x 32 lanes -> D <= 2048

struct DropSpec {
  float p;
  uint32_t threshold;
  float inv_keep;
  unsigned long long seed;
};
static DropSpec make_drop(float p, unsigned long long seed) {
  DropSpec d;
  d.p = p;
  d.threshold = dropout_threshold(p);
  d.inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  d.seed = seed;
  return d;
}
// keep-scale for the 8 consecutive elements starting at flat index `idx` (idx % 8 == 0)
__device__ __forceinline__ void drop_scale8(const DropSpec& d, long long idx, float (&s)[8]) {
  const uint32_t k = dropout_keep8(d.seed, static_cast<uint64_t>(idx >> 3), d.threshold);
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = (k >> i) & 1 ? d.inv_keep : 0.f;
}

// y = LN(x) * gamma + beta, optional dropout on y.  Stats saved for the backward.
// EMBED: x is not read from memory but formed as word[ids[row]] + pos[row % S] (fp32 tables) and also
// written out (bf16) as the saved LayerNorm input.
template <bool EMBED, int NC>
__global__ void __launch_bounds__(LN_WARPS * 32, NC <= 2 ? 6 : (NC <= 4 ? 4 : 2))
layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, const long long* __restrict__ ids,
                     const float* __restrict__ word, const float* __restrict__ pos, const int* __restrict__ pos_ids,
                     const float* __restrict__ type_row, int S, int vocab,
                     __nv_bfloat16* __restrict__ x_saved, const float* __restrict__ gamma,
                     const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out, int M, int D, float eps, DropSpec drop) {
  pdl_launch_dependents();   // see launch_pdl (common.cuh)
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * LN_WARPS + warp;
  if (row >= M) return;
  const int chunks = D >> 3;
  float v[NC][8];   // registers (and with them the rows in flight per SM) scale with D
  float2 sum2 = make_float2(0.f, 0.f);
  const float* wrow = nullptr;
  const float* prow = nullptr;
  if constexpr (EMBED) {
    long long id = ids[row];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    wrow = word + id * D;
    // BERT-family position ids are 0..S-1; RoBERTa / XLM-R derive them from the pad pattern (pos_ids)
    prow = pos + static_cast<long long>(pos_ids != nullptr ? pos_ids[row] : row % S) * D;
  }
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int c = lane + j * 32;
    if (c < chunks) {
      if constexpr (EMBED) {
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(wrow + c * 8));
        const float4 a1 = __ldg(reinterpret_cast<const float4*>(wrow + c * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(prow + c * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(prow + c * 8 + 4));
        v[j][0] = a0.x + b0.x; v[j][1] = a0.y + b0.y; v[j][2] = a0.z + b0.z; v[j][3] = a0.w + b0.w;
        v[j][4] = a1.x + b1.x; v[j][5] = a1.y + b1.y; v[j][6] = a1.z + b1.z; v[j][7] = a1.w + b1.w;
        if (type_row != nullptr) {   // token_type_embeddings[0] (all segment ids are 0 on this path)
          const float4 t0 = __ldg(reinterpret_cast<const float4*>(type_row + c * 8));
          const float4 t1 = __ldg(reinterpret_cast<const float4*>(type_row + c * 8 + 4));
          v[j][0] += t0.x; v[j][1] += t0.y; v[j][2] += t0.z; v[j][3] += t0.w;
          v[j][4] += t1.x; v[j][5] += t1.y; v[j][6] += t1.z; v[j][7] += t1.w;
        }
        store8(x_saved + row * D + c * 8, v[j]);
        // statistics are taken on the bf16-rounded values: that is what the backward will see
        load8(x_saved + row * D + c * 8, v[j]);
      } else {
        load8(x + row * D + c * 8, v[j]);
      }
      // packed fp32 (FADD2 / FFMA2): two elements per issue slot -- the kernel was issue-bound at ~50 % of the HBM rate
      const float2 t = fadd2(fadd2(make_float2(v[j][0], v[j][1]), make_float2(v[j][2], v[j][3])),
                             fadd2(make_float2(v[j][4], v[j][5]), make_float2(v[j][6], v[j][7])));
      sum2 = fadd2(sum2, t);
    }
  }
  const float mean = warp_sum(sum2.x + sum2.y) / D;
  float2 sq2 = make_float2(0.f, 0.f);
  const float2 nmean2 = splat2(-mean);
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int c = lane + j * 32;
    if (c < chunks) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const float2 d = fadd2(make_float2(v[j][i], v[j][i + 1]), nmean2);
        sq2 = ffma2(d, d, sq2);
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq2.x + sq2.y) / D + eps);
  if (lane == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int c = lane + j * 32;
    if (c < chunks) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c * 8 + 4));
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o[8];
      const float2 rstd2 = splat2(rstd);
#pragma unroll
      for (int i = 0; i < 8; i += 2) {      // (x - mean) * rstd * gamma + beta, as the reference orders it
        const float2 h = fmul2(fadd2(make_float2(v[j][i], v[j][i + 1]), nmean2), rstd2);
        const float2 r = ffma2(h, make_float2(g[i], g[i + 1]), make_float2(b[i], b[i + 1]));
        o[i] = r.x;
        o[i + 1] = r.y;
      }
      if (drop.p > 0.f) {
        float s[8];
        drop_scale8(drop, row * D + c * 8, s);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] *= s[i];
      }
      store8(y + row * D + c * 8, o);
    }
  }
}

// dx = LN backward; optional dropout mask on the incoming dy (in_drop: the mask that was applied to the LN output
// in the forward) and optional second output dx2 = dropout-masked dx (out_drop: the mask that was applied to the
// GEMM branch feeding this LN's input).  dgamma/dbeta are accumulated with fp32 atomics (pre-zeroed by caller
// or holding the gradient to accumulate onto).
// Rows reach the warps through a per-warp ring of LN_DEPTH shared-memory slots filled by 1-D bulk copies
// (cp.async.bulk + mbarrier): the prefetch depth no longer costs registers (holding even one extra row of dy / x in
// registers spilled, and with a single row in flight per warp the pass ran at 2.5 TB/s).
constexpr int ln_depth(int nc) { return nc <= 4 ? 3 : 2; }   // D = 2048 with an addend: 2 x 3 rows x 4 KB x 8 warps
template <int NC>
__global__ void __launch_bounds__(LN_WARPS * 32, NC <= 3 ? 2 : 1)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                     const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                     const float* __restrict__ gamma, const __nv_bfloat16* __restrict__ addend,
                     __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dx2, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, int M, int D, int rows_per_cta, DropSpec in_drop, DropSpec out_drop) {
  pdl_launch_dependents();   // see launch_pdl (common.cuh)
  pdl_wait();
  // [LN_WARPS][LN_DEPTH][dy row | x row] bf16 while rows stream; reused as float [LN_WARPS][2][D] for the column sums
  constexpr int LN_DEPTH = ln_depth(NC);
  extern __shared__ __align__(128) uint8_t ln_smem[];
  __shared__ uint64_t bars[LN_WARPS][LN_DEPTH];
  float* red = reinterpret_cast<float*>(ln_smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = D >> 3;
  const uint32_t row_bytes = static_cast<uint32_t>(D) * 2;
  const uint32_t slot_rows = addend != nullptr ? 3u : 2u;   // dy | x (| addend)
  uint8_t* ring = ln_smem + static_cast<size_t>(warp) * LN_DEPTH * slot_rows * row_bytes;
  float ag[NC][8], ab[NC][8];
#pragma unroll
  for (int j = 0; j < NC; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) ag[j][i] = ab[j][i] = 0.f;

  const long long row_begin = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long row_end = min(row_begin + rows_per_cta, static_cast<long long>(M));
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < LN_DEPTH; ++s) mbar_init(&bars[warp][s], 1);
    fence_barrier_init();
  }
  __syncwarp();
  auto issue = [&](long long r, int s) {   // lane 0: both rows of slot s
    uint8_t* dst = ring + static_cast<size_t>(s) * slot_rows * row_bytes;
    mbar_expect_tx(&bars[warp][s], slot_rows * row_bytes);
    bulk_load_1d(dst, dy + r * D, row_bytes, &bars[warp][s]);
    bulk_load_1d(dst + row_bytes, x + r * D, row_bytes, &bars[warp][s]);
    if (addend != nullptr) bulk_load_1d(dst + 2 * row_bytes, addend + r * D, row_bytes, &bars[warp][s]);
  };
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < LN_DEPTH; ++s) {
      const long long r = row_begin + warp + static_cast<long long>(s) * LN_WARPS;
      if (r < row_end) issue(r, s);
    }
  }
  // per-row statistics: lane l holds those of the warp's (32 k + l)-th row, handed out by shuffle (a dependent
  // scalar load at the top of every row cost a DRAM round trip per row)
  float lane_mean = 0.f, lane_rstd = 0.f;
  int it = 0;
  for (long long row = row_begin + warp; row < row_end; row += LN_WARPS, ++it) {
    const int slot = it % LN_DEPTH;
    if ((it & 31) == 0) {
      const long long r = row + static_cast<long long>(lane) * LN_WARPS;
      lane_mean = r < row_end ? mean_in[r] : 0.f;
      lane_rstd = r < row_end ? rstd_in[r] : 0.f;
    }
    const float mean = __shfl_sync(0xffffffffu, lane_mean, it & 31);
    const float rstd = __shfl_sync(0xffffffffu, lane_rstd, it & 31);
    mbar_wait(&bars[warp][slot], (it / LN_DEPTH) & 1);
    const uint8_t* src = ring + static_cast<size_t>(slot) * slot_rows * row_bytes;
    float g_dy[NC][8], xh[NC][8];
    float2 s1v = make_float2(0.f, 0.f), s2v = make_float2(0.f, 0.f);
    const float2 nmean2 = splat2(-mean), rstd2 = splat2(rstd);
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const int c = lane + j * 32;
      if (c < chunks) {
        float d[8], xv[8];
        unpack8(*reinterpret_cast<const uint4*>(src + c * 16), d);
        unpack8(*reinterpret_cast<const uint4*>(src + row_bytes + c * 16), xv);
        if (in_drop.p > 0.f) {
          float s[8];
          drop_scale8(in_drop, row * D + c * 8, s);
#pragma unroll
          for (int i = 0; i < 8; ++i) d[i] *= s[i];
        }
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c * 8 + 4));
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int i = 0; i < 8; i += 2) {     // packed fp32: two elements per FMA-pipe issue slot
          const float2 d2 = make_float2(d[i], d[i + 1]);
          const float2 h = fmul2(fadd2(make_float2(xv[i], xv[i + 1]), nmean2), rstd2);
          xh[j][i] = h.x; xh[j][i + 1] = h.y;
          const float2 a2 = ffma2(d2, h, make_float2(ag[j][i], ag[j][i + 1]));
          ag[j][i] = a2.x; ag[j][i + 1] = a2.y;
          const float2 b2 = fadd2(make_float2(ab[j][i], ab[j][i + 1]), d2);
          ab[j][i] = b2.x; ab[j][i + 1] = b2.y;
          const float2 dg = fmul2(d2, make_float2(g[i], g[i + 1]));
          g_dy[j][i] = dg.x; g_dy[j][i + 1] = dg.y;
          s1v = fadd2(s1v, dg);
          s2v = ffma2(dg, h, s2v);
        }
      }
    }
    const float s1 = warp_sum(s1v.x + s1v.y) / D;
    const float s2 = warp_sum(s2v.x + s2v.y) / D;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const int c = lane + j * 32;
      if (c < chunks) {
        float o[8];
        const float2 ns1 = splat2(-s1), ns2 = splat2(-s2);
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          const float2 t = ffma2(make_float2(xh[j][i], xh[j][i + 1]), ns2,
                                 fadd2(make_float2(g_dy[j][i], g_dy[j][i + 1]), ns1));
          const float2 r = fmul2(t, rstd2);
          o[i] = r.x;
          o[i + 1] = r.y;
        }
        if (addend != nullptr) {   // pre-LN blocks: the residual stream's gradient joins here (dx2 stays LN-only)
          float a[8], t[8];
          unpack8(*reinterpret_cast<const uint4*>(src + 2 * row_bytes + c * 16), a);
#pragma unroll
          for (int i = 0; i < 8; ++i) t[i] = o[i] + a[i];
          store8(dx + row * D + c * 8, t);
        } else {
          store8(dx + row * D + c * 8, o);
        }
        if (dx2 != nullptr) {
          float s[8];
          drop_scale8(out_drop, row * D + c * 8, s);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] *= s[i];
          store8(dx2 + row * D + c * 8, o);
        }
      }
    }
    // all of this row's slot reads are complete (their values were stored above): refill the slot LN_DEPTH rows ahead
    __syncwarp();
    if (lane == 0) {
      const long long nr = row + static_cast<long long>(LN_DEPTH) * LN_WARPS;
      if (nr < row_end) issue(nr, slot);
    }
  }
  // cross-warp reduction of the column sums, then one atomic per column per CTA
  __syncthreads();   // every warp is done with its ring slots: the buffer turns into the float staging area
  float* rg = red + warp * 2 * D;
  float* rb = rg + D;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int c = lane + j * 32;
    if (c < chunks) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        rg[c * 8 + i] = ag[j][i];
        rb[c * 8 + i] = ab[j][i];
      }
    }
  }
  __syncthreads();
  for (int col = threadIdx.x; col < D; col += blockDim.x) {
    float sg = 0.f, sb = 0.f;
#pragma unroll
    for (int w = 0; w < LN_WARPS; ++w) {
      sg += red[w * 2 * D + col];
      sb += red[w * 2 * D + D + col];
    }
    atomicAdd(dgamma + col, sg);
    atomicAdd(dbeta + col, sb);
  }
}

// Embedding backward: scatter-add d(word+pos sum) rows into the fp32 gradient tables.
__global__ void __launch_bounds__(256)
embedding_bwd_kernel(const __nv_bfloat16* __restrict__ dx, const long long* __restrict__ ids,
                     const int* __restrict__ pos_ids, long long pos_padding_idx, int S, int vocab,
                     long long padding_idx, float* __restrict__ dword, float* __restrict__ dpos, int M, int D) {
  const int chunks = D >> 2;
  const long long total = static_cast<long long>(M) * chunks;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / chunks;
    const int c = static_cast<int>(i - row * chunks) * 4;
    long long id = ids[row];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const uint2 r = *reinterpret_cast<const uint2*>(dx + row * D + c);
    const float2 a = unpack_bf16x2_dev(r.x), b = unpack_bf16x2_dev(r.y);
    float* w = dword + id * D + c;
    const long long pid = pos_ids != nullptr ? pos_ids[row] : row % S;
    float* q = dpos + pid * D + c;
    // nn.Embedding(padding_idx=pad_token_id): the pad row never receives a gradient (modeling_distilbert.py:86)
    if (ids[row] != padding_idx)
      asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(w), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y) : "memory");
    // RoBERTa's position table is nn.Embedding(padding_idx=pad): its pad row gets no gradient either
    if (pid != pos_padding_idx)
      asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(q), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y) : "memory");
  }
}

// RoBERTa / XLM-R position ids (transformers/models/xlm_roberta/modeling_xlm_roberta.py
// create_position_ids_from_input_ids): pos = cumsum(ids != pad) * (ids != pad) + pad.  One warp per sequence.
__global__ void __launch_bounds__(128)
position_ids_kernel(const long long* __restrict__ ids, long long pad_id, int B, int S, int* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + warp;
  if (b >= B) return;
  int carry = 0;
  for (int s0 = 0; s0 < S; s0 += 32) {
    const int s = s0 + lane;
    const int real = (s < S && ids[static_cast<long long>(b) * S + s] != pad_id) ? 1 : 0;
    int scan = real;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, scan, o);
      if (lane >= o) scan += t;
    }
    if (s < S) out[static_cast<long long>(b) * S + s] = real ? carry + scan + static_cast<int>(pad_id) : static_cast<int>(pad_id);
    carry += __shfl_sync(0xffffffffu, scan, 31);
  }
}

// ViT token assembly (transformers/models/vit/modeling_vit.py ViTEmbeddings.forward):
//   x[b, 0] = cls + pos[0];  x[b, 1 + i] = patch[b, i] + pos[1 + i]      (fp32 cls / pos, bf16 tokens)
__global__ void __launch_bounds__(256)
vit_assemble_fwd_kernel(const __nv_bfloat16* __restrict__ patch, const float* __restrict__ cls,
                        const float* __restrict__ pos, __nv_bfloat16* __restrict__ x, int B, int P, int D) {
  const int chunks = D >> 3;
  const long long total = static_cast<long long>(B) * (P + 1) * chunks;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / chunks;
    const int c = static_cast<int>(i - row * chunks) * 8;
    const int b = static_cast<int>(row / (P + 1)), t = static_cast<int>(row - static_cast<long long>(b) * (P + 1));
    float v[8];
    if (t == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = cls[c + k];
    } else {
      load8(patch + (static_cast<long long>(b) * P + (t - 1)) * D + c, v);
    }
    const float* pr = pos + static_cast<long long>(t) * D + c;
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] += pr[k];
    store8(x + row * D + c, v);
  }
}

// backward of the assembly: dpatch[b, i] = dx[b, 1 + i];  dpos[t] += sum_b dx[b, t];  dcls += sum_b dx[b, 0].
// One CTA per token position t, thread = 8 channels, loop over the batch (coalesced 16-byte reads).
__global__ void __launch_bounds__(256)
vit_assemble_bwd_kernel(const __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dpatch,
                        float* __restrict__ dcls, float* __restrict__ dpos, int B, int P, int D) {
  const int t = blockIdx.x;
  for (int c = threadIdx.x * 8; c < D; c += blockDim.x * 8) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    constexpr int U = 4;
    for (int b0 = 0; b0 < B; b0 += U) {
      uint4 raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        raw[u] = (b0 + u) < B ? __ldg(reinterpret_cast<const uint4*>(
                                    dx + (static_cast<long long>(b0 + u) * (P + 1) + t) * D + c))
                              : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (b0 + u >= B) break;
        if (t > 0)
          *reinterpret_cast<uint4*>(dpatch + (static_cast<long long>(b0 + u) * P + (t - 1)) * D + c) = raw[u];
        float v[8];
        unpack8(raw[u], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += v[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      dpos[static_cast<long long>(t) * D + c + k] += acc[k];
      if (t == 0) dcls[c + k] += acc[k];
    }
  }
}

// attention_mask (int64, 1 = real token) -> additive key bias (0 / -inf), as HF builds it
// (modeling_distilbert.py:415-419 create_bidirectional_mask)
__global__ void mask_to_bias_kernel(const long long* __restrict__ mask, float* __restrict__ bias, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) bias[i] = mask[i] != 0 ? 0.f : -INFINITY;
}

}  // namespace b200

using namespace b200;

static bool ln_shape_ok(int M, int D) { return M > 0 && D > 0 && (D & 7) == 0 && D <= LN_MAX_CHUNKS * 256; }

// y[M,D] = dropout(LayerNorm(x[M,D])) ; saves mean/rstd (fp32 [M]).
B200MM_API int b200mm_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                    float* rstd, int M, int D, float eps, float p_drop, unsigned long long seed,
                                    void* stream) {
  if (!ln_shape_ok(M, D)) return B200MM_ERR_BAD_ARG;
#define LAUNCH_LN_FWD(NC)                                                                                           \
  launch_pdl(layernorm_fwd_kernel<false, NC>, dim3(ceil_div(M, LN_WARPS)), dim3(LN_WARPS * 32), 0,                  \
             static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(x), nullptr, nullptr, nullptr,      \
             nullptr, nullptr, 1, 1, nullptr, gamma, beta, static_cast<__nv_bfloat16*>(y), mean, rstd, M, D, eps,     \
             make_drop(p_drop, seed))
  const int nc = ceil_div(D, 256);
  if (nc <= 1) LAUNCH_LN_FWD(1);
  else if (nc == 2) LAUNCH_LN_FWD(2);
  else if (nc == 3) LAUNCH_LN_FWD(3);
  else if (nc == 4) LAUNCH_LN_FWD(4);
  else LAUNCH_LN_FWD(8);
#undef LAUNCH_LN_FWD
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

// Embeddings: x_saved = bf16(word[ids] + pos[pos_ids ? pos_ids[m] : m % S] (+ type_row)); y = dropout(LN(x_saved)).
// ids int64 [M = B*S]; pos_ids (nullable) int32 [M]; type_row (nullable) fp32 [D] = token_type_embeddings[0].
B200MM_API int b200mm_embed_layernorm_fwd(const long long* ids, const float* word, const float* pos,
                                          const int* pos_ids, const float* type_row, int S, int vocab,
                                          const float* gamma, const float* beta, void* x_saved, void* y, float* mean,
                                          float* rstd, int M, int D, float eps, float p_drop, unsigned long long seed,
                                          void* stream) {
  if (!ln_shape_ok(M, D) || S <= 0 || vocab <= 0) return B200MM_ERR_BAD_ARG;
#define LAUNCH_LN_EMB(NC)                                                                                          \
  layernorm_fwd_kernel<true, NC><<<ceil_div(M, LN_WARPS), LN_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(  \
      nullptr, ids, word, pos, pos_ids, type_row, S, vocab, static_cast<__nv_bfloat16*>(x_saved), gamma, beta,     \
      static_cast<__nv_bfloat16*>(y), mean, rstd, M, D, eps, make_drop(p_drop, seed))
  const int nc = ceil_div(D, 256);
  if (nc <= 1) LAUNCH_LN_EMB(1);
  else if (nc == 2) LAUNCH_LN_EMB(2);
  else if (nc == 3) LAUNCH_LN_EMB(3);
  else if (nc == 4) LAUNCH_LN_EMB(4);
  else LAUNCH_LN_EMB(8);
#undef LAUNCH_LN_EMB
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

// LayerNorm backward.  dx2 (nullable) receives the LN gradient with the (p_out, seed_out) dropout mask applied;
// addend (nullable, bf16 [M,D]) is added to dx only (pre-LN residual stream).
B200MM_API int b200mm_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd,
                                    const float* gamma, const void* addend, void* dx, void* dx2, float* dgamma,
                                    float* dbeta, int M, int D,
                                    float p_in, unsigned long long seed_in, float p_out, unsigned long long seed_out,
                                    void* stream) {
  if (!ln_shape_ok(M, D)) return B200MM_ERR_BAD_ARG;
  const DeviceInfo& dev = device_info();
  const int nc = ceil_div(D, 256);
  const int nc_inst = nc <= 4 ? nc : 8;
  // ring: LN_WARPS x depth x (dy row + x row (+ addend row)) bf16; >= the 8 D bytes per warp of the float staging
  const size_t smem =
      static_cast<size_t>(LN_WARPS) * ln_depth(nc_inst) * (addend != nullptr ? 3 : 2) * D * sizeof(__nv_bfloat16);
  // one wave of resident CTAs (every CTA ends with 2*D fp32 atomics onto the same 2*D addresses)
#define LAUNCH_LN_BWD(NC)                                                                                          \
  do {                                                                                                             \
    static bool configured = false;                                                                                \
    if (!configured) {                                                                                             \
      cudaError_t e = cudaFuncSetAttribute(layernorm_bwd_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                           LN_WARPS * ln_depth(NC) * 3 * NC * 256 * 2);                            \
      if (e != cudaSuccess) return static_cast<int>(e);                                                            \
      configured = true;                                                                                           \
    }                                                                                                              \
    int resident = 1;                                                                                              \
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, layernorm_bwd_kernel<NC>, LN_WARPS * 32, smem) != \
            cudaSuccess || resident < 1)                                                                           \
      resident = 1;                                                                                                \
    const int target_ctas = (dev.num_sms > 0 ? dev.num_sms : 148) * resident;                                      \
    int rows_per_cta = ceil_div(M, target_ctas);                                                                   \
    rows_per_cta = ceil_div(rows_per_cta, LN_WARPS) * LN_WARPS;                                                    \
    const int grid = ceil_div(M, rows_per_cta);                                                                    \
    launch_pdl(layernorm_bwd_kernel<NC>, dim3(grid), dim3(LN_WARPS * 32), smem, static_cast<cudaStream_t>(stream), \
               static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x), mean, rstd, gamma,     \
               static_cast<const __nv_bfloat16*>(addend), static_cast<__nv_bfloat16*>(dx),                         \
               static_cast<__nv_bfloat16*>(dx2), dgamma, dbeta, M, D, rows_per_cta,                                \
               make_drop(p_in, seed_in), make_drop(dx2 ? p_out : 0.f, seed_out));                                  \
  } while (0)
  if (nc <= 1) LAUNCH_LN_BWD(1);
  else if (nc == 2) LAUNCH_LN_BWD(2);
  else if (nc == 3) LAUNCH_LN_BWD(3);
  else if (nc == 4) LAUNCH_LN_BWD(4);
  else LAUNCH_LN_BWD(8);
#undef LAUNCH_LN_BWD
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

// dword[ids[m]] += dx[m] (skipped where ids[m] == padding_idx; pass -1 for none);
// dpos[pos_ids ? pos_ids[m] : m % S] += dx[m] (skipped where that index == pos_padding_idx; -1 for none)
B200MM_API int b200mm_embedding_bwd(const void* dx, const long long* ids, const int* pos_ids,
                                    long long pos_padding_idx, int S, int vocab, long long padding_idx,
                                    float* dword, float* dpos, int M, int D, void* stream) {
  if (M <= 0 || D <= 0 || (D & 3) || S <= 0) return B200MM_ERR_BAD_ARG;
  const long long total = static_cast<long long>(M) * (D >> 2);
  const int grid = static_cast<int>(total / 256 > 148 * 16 ? 148 * 16 : ceil_div(total, 256LL));
  embedding_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dx), ids, pos_ids, pos_padding_idx, S, vocab, padding_idx, dword, dpos, M, D);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

B200MM_API int b200mm_mask_to_bias(const long long* mask, float* bias, long long n, void* stream) {
  if (n <= 0) return B200MM_ERR_BAD_ARG;
  mask_to_bias_kernel<<<static_cast<int>(ceil_div(n, 256LL)), 256, 0, static_cast<cudaStream_t>(stream)>>>(mask, bias, n);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

// pos_ids[B*S] (int32) = cumsum(ids != pad) * (ids != pad) + pad   (RoBERTa / XLM-R position ids)
B200MM_API int b200mm_position_ids(const long long* ids, long long pad_id, int B, int S, int* out, void* stream) {
  if (B <= 0 || S <= 0) return B200MM_ERR_BAD_ARG;
  position_ids_kernel<<<ceil_div(B, 4), 128, 0, static_cast<cudaStream_t>(stream)>>>(ids, pad_id, B, S, out);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

// ViT embeddings: tokens[B*(P+1), D] = [cls | patch rows] + pos   and the backward of that assembly.
B200MM_API int b200mm_vit_assemble_fwd(const void* patch, const float* cls, const float* pos, void* x, int B, int P,
                                       int D, void* stream) {
  if (B <= 0 || P <= 0 || D <= 0 || (D & 7)) return B200MM_ERR_BAD_ARG;
  const long long total = static_cast<long long>(B) * (P + 1) * (D >> 3);
  const int grid = static_cast<int>(total / 256 > 148 * 16 ? 148 * 16 : ceil_div(total, 256LL));
  vit_assemble_fwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(patch), cls, pos, static_cast<__nv_bfloat16*>(x), B, P, D);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
B200MM_API int b200mm_vit_assemble_bwd(const void* dx, void* dpatch, float* dcls, float* dpos, int B, int P, int D,
                                       void* stream) {
  if (B <= 0 || P <= 0 || D <= 0 || (D & 7)) return B200MM_ERR_BAD_ARG;
  const int threads = D / 8 < 256 ? ((D / 8 + 31) / 32) * 32 : 256;
  vit_assemble_bwd_kernel<<<P + 1, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dx), static_cast<__nv_bfloat16*>(dpatch), dcls, dpos, B, P, D);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
