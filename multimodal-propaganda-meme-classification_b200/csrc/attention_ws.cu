// Warp-specialised fused attention for one-tile sequences (S <= 128, head_dim 64): the text towers of every BASELINE
// configuration (DistilBERT / BERT / XLM-R at 128 tokens).
//
// Replaces: transformers/models/distilbert/modeling_distilbert.py:126-151
// (softmax(Q K^T * d^-1/2 + mask) -> dropout -> @ V) and its autograd backward.
//
// Why: the single-role kernels (attention_tcgen05.cu) run load -> score MMA -> softmax -> value MMA -> store as one
// dependent chain per head and measured 85 us (fwd) / 180 us (bwd) at 3072 heads against 30 / 60 us of HBM time: every
// thread waits for the tensor core and the tensor core waits for every thread.  Here the roles are separate warps that
// only meet at mbarriers, with several heads in flight per SM:
//
//   warp 0       TMA warp      loads Q | K | V (| dO) of head n+SLOTS as soon as slot storage frees up; also issues the
//                              TMA stores of finished tiles and waits for their shared-memory reads (so no compute
//                              thread ever blocks on a bulk-group)
//   warp 1       MMA warp      one elected thread polls "operands landed" / "P staged" barriers and issues tcgen05.mma
//                              for whichever head is ready (score MMAs of later heads overtake the value MMAs of earlier
//                              ones) -- the tensor core is never behind a softmax
//   warps 2..17  softmax warpgroups, one per slot: TMEM -> registers -> exp2 / dropout -> bf16 P (dS) in swizzled shared
//                memory -> mbarrier arrive; later the epilogue of the same head (TMEM -> bf16 -> staged tile)
//
// forward : 4 slots x 128 threads (thread = query row, all 128 key columns: max and sum need no exchange);
//           slot = Q | K | V (48 KB), P overlays Q | K once the score MMA has retired, the O tile is staged over P;
//           TMEM: 128 columns per slot, O accumulates in the first 64 columns of S (read out by then).
// backward: 2 slots x 256 threads (two threads per row split the key columns);
//           slot = Q | K | dO | V | X (112 KB): P = V | X0 (V is dead once S and dP exist), dS = X1 | X2, and the dQ / dK /
//           dV tiles are staged over X while the NEXT head's operands already stream into Q | K | dO | V;
//           TMEM: S | dP per slot; dV / dK alias S's columns and dQ aliases dP's (all read out before the gradient
//           MMAs are issued).
// Dropout: the same Philox stream and element indexing as the single-role kernels (forward and backward of either
// family can be mixed; tests/test_kernels_gpu.py cross-checks them).  The 1 / (1 - p) factor of the forward is folded
// into the final 1 / rowsum scale.
#include "attention_common.cuh"

namespace b200 {

// Polling role loops (TMA warp, MMA warp) must not hang the GPU on a protocol bug: trap after ~2 s without progress,
// like ptx.cuh's bounded mbar_wait.
struct PollWatchdog {
  long long t0;
};
__device__ __forceinline__ void dog_progress(PollWatchdog& d) { d.t0 = clock64(); }
__device__ __forceinline__ void dog_check(const PollWatchdog& d, const char* role) {
  if (clock64() - d.t0 > 4000000000LL) {
    printf("b200mm: attention %s stalled, block %d\n", role, blockIdx.x);
    __trap();
  }
}

bool attn_ws_enabled() {
  const char* e = std::getenv("B200MM_ATTN_WS");
  return !(e && e[0] == '0');
}

// ------------------------------------------------------------------------------------------------ forward
constexpr int WSF_SLOTS = 4;
constexpr int WSF_SLOT_BYTES = 3 * ATT_TILE_BYTES;
constexpr int WSF_THREADS = 64 + WSF_SLOTS * 128;
constexpr int WSF_SMEM = WSF_SLOTS * WSF_SLOT_BYTES + WSF_SLOTS * ATT_T * 4 + 256 + 1024;

__global__ void __launch_bounds__(WSF_THREADS, 1)
attn_fwd_ws_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_out,
                   const AttnParams p) {
  // 1024-byte alignment (SWIZZLE_128B atoms) by pointer arithmetic ON the __shared__ array: the compiler keeps the
  // address space and emits LDS / STS (the former round-up through uintptr_t turned every access of the tiles,
  // the staging boxes and the bias rows into generic LD.E / ST.E, which queue with the global loads)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* slots = smem;
  float* sBias = reinterpret_cast<float*>(slots + WSF_SLOTS * WSF_SLOT_BYTES);   // [SLOTS][128]
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(sBias + WSF_SLOTS * ATT_T);   // operands of the slot's head landed
  uint64_t* bar_s = bar_full + WSF_SLOTS;        // S = Q K^T is in TMEM
  uint64_t* bar_p = bar_s + WSF_SLOTS;           // P staged in shared memory (128 arrivals)
  uint64_t* bar_o = bar_p + WSF_SLOTS;           // O = P V is in TMEM
  uint64_t* bar_staged = bar_o + WSF_SLOTS;      // the bf16 O tile is staged; TMEM of the slot fully read
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_staged + WSF_SLOTS);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_launch_dependents();   // see launch_pdl (common.cuh): the set-up below overlaps the previous kernel's tail
  if (tid == 0) {
    tma_prefetch_desc(&tma_qkv);
    tma_prefetch_desc(&tma_out);
    for (int s = 0; s < WSF_SLOTS; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_s[s], 1);
      mbar_init(&bar_p[s], 128);
      mbar_init(&bar_o[s], 1);
      mbar_init(&bar_staged[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();   // first global-memory access comes after this point

  const int items = p.B * p.H;
  const int n_local = (items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA warp: loads + stores
    if (elect_one()) {
      int nl = 0, nst = 0;
      PollWatchdog dog;
      dog_progress(dog);
      while (nst < n_local) {
        dog_check(dog, "fwd TMA warp");
        if (nl < n_local && nl - WSF_SLOTS < nst) {       // the slot's previous tile has left shared memory
          const int s = nl & (WSF_SLOTS - 1);
          const int item = blockIdx.x + nl * gridDim.x;
          const int b = item / p.H, h = item - b * p.H;
          uint8_t* base = slots + s * WSF_SLOT_BYTES;
          mbar_expect_tx(&bar_full[s], 3 * ATT_TILE_BYTES);
          tma_load_3d(base, &tma_qkv, &bar_full[s], h * ATT_D, 0, b);
          tma_load_3d(base + ATT_TILE_BYTES, &tma_qkv, &bar_full[s], p.D + h * ATT_D, 0, b);
          tma_load_3d(base + 2 * ATT_TILE_BYTES, &tma_qkv, &bar_full[s], 2 * p.D + h * ATT_D, 0, b);
          ++nl;
          dog_progress(dog);
        }
        if (nst < nl) {
          const int s = nst & (WSF_SLOTS - 1);
          if (mbar_try_wait(&bar_staged[s], (nst >> 2) & 1)) {
            const int item = blockIdx.x + nst * gridDim.x;
            const int b = item / p.H, h = item - b * p.H;
            tma_store_3d(&tma_out, slots + s * WSF_SLOT_BYTES, h * ATT_D, 0, b);   // rows >= S clipped by the map
            tma_store_commit();
            tma_store_wait_read<0>();
            ++nst;
            dog_progress(dog);
          }
        }
      }
      tma_store_wait_read<0>();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA warp
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
      int nq = 0, np = 0;
      PollWatchdog dog;
      dog_progress(dog);
      while (np < n_local) {
        dog_check(dog, "fwd MMA warp");
        if (nq < n_local) {
          const int s = nq & (WSF_SLOTS - 1);
          if (mbar_try_wait(&bar_full[s], (nq >> 2) & 1)) {
            tc_fence_after_sync();
            const uint32_t q = smem_u32(slots + s * WSF_SLOT_BYTES), k = q + ATT_TILE_BYTES;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16(tmem + s * ATT_T, umma_desc_sw128(q + kk * 32, 16, 1024), umma_desc_sw128(k + kk * 32, 16, 1024),
                        idesc_s, kk > 0);
            umma_commit(&bar_s[s]);
            ++nq;
            dog_progress(dog);
          }
        }
        if (np < nq) {
          const int s = np & (WSF_SLOTS - 1);
          if (mbar_try_wait(&bar_p[s], (np >> 2) & 1)) {
            tc_fence_after_sync();
            const uint32_t pb = smem_u32(slots + s * WSF_SLOT_BYTES), v = pb + 2 * ATT_TILE_BYTES;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_bf16(tmem + s * ATT_T, umma_desc_sw128(pb + (kk >> 2) * ATT_TILE_BYTES + (kk & 3) * 32, 16, 1024),
                        umma_desc_sw128(v + kk * 2048, 8192, 1024), idesc_o, kk > 0);
            umma_commit(&bar_o[s]);
            ++np;
            dog_progress(dog);
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax warpgroup of slot wg
    const int wg = (warp - 2) >> 2;
    const int wt = tid - 64 - wg * 128;                    // 0..127 inside the warpgroup
    const int row = (warp & 3) * 32 + lane;                // TMEM lane quarter = warp % 4
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t t_S = tmem + wg * ATT_T;
    uint8_t* base = slots + wg * WSF_SLOT_BYTES;
    float* bias = sBias + wg * ATT_T;
    const bool use_drop = p.p_drop > 0.f;
    const float out_scale = use_drop ? p.inv_keep : 1.f;
    const float2 sc2 = splat2(p.scale_log2);
    for (int n = wg; n < n_local; n += WSF_SLOTS) {
      const int k = n >> 2;
      const int item = blockIdx.x + n * gridDim.x;
      const int b = item / p.H;
      // key bias of this head (the previous head's readers all passed the named barrier below)
      bias[wt] = wt < p.S ? (p.key_bias ? __ldg(p.key_bias + b * p.S + wt) * LOG2E : 0.f) : -INFINITY;
      named_bar_sync(1 + wg, 128);
      mbar_wait(&bar_s[wg], k & 1);
      tc_fence_after_sync();
      // ---- exact row maximum
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(t_S + lane_addr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias + c * 32 + i);
          const float2 t0 = ffma2(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2,
                                  make_float2(b4.x, b4.y));
          const float2 t1 = ffma2(make_float2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])), sc2,
                                  make_float2(b4.z, b4.w));
          mx = fmax3(mx, t0.x, t0.y);
          mx = fmax3(mx, t1.x, t1.y);
        }
      }
      if (mx == -INFINITY) mx = 0.f;
      const float2 nmx2 = splat2(-mx);
      // ---- P = exp2(S - max) (dropout) -> shared memory (over Q | K, both dead: the score MMA has retired)
      float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(t_S + lane_addr + c * 32, v);
        tmem_ld_wait();
        float x[32];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {      // packed fp32: two elements per FFMA2 / FADD2 issue slot
          const float4 b4 = *reinterpret_cast<const float4*>(bias + c * 32 + i);
          const float2 t0 = fadd2(ffma2(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2,
                                        make_float2(b4.x, b4.y)), nmx2);
          const float2 t1 = fadd2(ffma2(make_float2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])), sc2,
                                        make_float2(b4.z, b4.w)), nmx2);
          x[i] = fast_exp2(t0.x); x[i + 1] = fast_exp2(t0.y);
          x[i + 2] = fast_exp2(t1.x); x[i + 3] = fast_exp2(t1.y);
          sum2 = fadd2(sum2, fadd2(make_float2(x[i], x[i + 1]), make_float2(x[i + 2], x[i + 3])));
        }
        if (use_drop) {
          const uint32_t keep = dropout_keep32(p.seed, (static_cast<uint64_t>(item) * ATT_T + row) * 4 + c,
                                               p.drop_threshold >> 16);
          if (p.drop_mask != nullptr)           // [item][c][row]: one coalesced 128-byte store per warp
            p.drop_mask[(static_cast<long long>(item) * 4 + c) * ATT_T + row] = keep;
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = (keep >> i) & 1 ? x[i] : 0.f;     // 1 / (1 - p) is applied to O
        }
        store_row32_sw128(base, row, c * 32, x);
      }
      const float sum = sum2.x + sum2.y;
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(&bar_p[wg]);
      // ---- epilogue: O / rowsum -> bf16 tile staged over P's first block
      mbar_wait(&bar_o[wg], k & 1);
      tc_fence_after_sync();
      const float inv = out_scale / sum;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(t_S + lane_addr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 inv2 = splat2(inv);
          const float2 o0 = fmul2(make_float2(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1])), inv2);
          const float2 o1 = fmul2(make_float2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3])), inv2);
          const float2 o2 = fmul2(make_float2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5])), inv2);
          const float2 o3 = fmul2(make_float2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7])), inv2);
          uint4 o;
          o.x = pack_bf16x2(o0.x, o0.y);
          o.y = pack_bf16x2(o1.x, o1.y);
          o.z = pack_bf16x2(o2.x, o2.y);
          o.w = pack_bf16x2(o3.x, o3.y);
          *reinterpret_cast<uint4*>(base + sw128_off(row, c * 4 + q)) = o;
        }
      }
      if (row < p.S && p.lse) p.lse[static_cast<long long>(item) * p.S + row] = (mx + log2f(sum)) * LN2;
      fence_proxy_async_smem();
      tc_fence_before_sync();
      named_bar_sync(1 + wg, 128);     // tile complete, TMEM of the slot read out by all 128 threads
      if (wt == 0) mbar_arrive(&bar_staged[wg]);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------ backward
constexpr int WSB_SLOTS = 2;
constexpr int WSB_SLOT_TILES = 7;     // Q | K | dO | V | X0 | X1 | X2
constexpr int WSB_SLOT_BYTES = WSB_SLOT_TILES * ATT_TILE_BYTES;
constexpr int WSB_THREADS = 64 + WSB_SLOTS * 256 + 32;   // TMA-load + MMA warps | softmax warps | store warp
constexpr int WSB_STORE_WARP = 2 + WSB_SLOTS * 8;
// tiles | key bias [SLOTS][128] | delta exchange [SLOTS][128] | barriers.  No alignment slack: the 14 tiles take 224 of
// the 227 KB, so the kernel relies on the __align__(1024) of the dynamic shared array (and traps if it is not honoured).
constexpr int WSB_SMEM = WSB_SLOTS * WSB_SLOT_BYTES + 2 * WSB_SLOTS * ATT_T * 4 + 128;
static_assert(WSB_SMEM <= 232448, "backward slots must fit the 227 KB of opt-in shared memory");

__global__ void __launch_bounds__(WSB_THREADS, 1)
attn_bwd_ws_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_do,
                   const __grid_constant__ CUtensorMap tma_dqkv, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem_raw) & 1023u) != 0u) {     // SWIZZLE_128B atoms need the 1024-byte base the declaration asks for
    if (threadIdx.x == 0) printf("b200mm: attention backward: dynamic shared memory base is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* slots = smem;
  float* sBias = reinterpret_cast<float*>(slots + WSB_SLOTS * WSB_SLOT_BYTES);   // [SLOTS][128]
  float* sDelta = sBias + WSB_SLOTS * ATT_T;                                      // [SLOTS][128]
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(sDelta + WSB_SLOTS * ATT_T);  // Q, K, V, dO landed
  uint64_t* bar_sc = bar_full + WSB_SLOTS;       // S and dP are in TMEM
  uint64_t* bar_pds = bar_sc + WSB_SLOTS;        // P and dS staged (256 arrivals)
  uint64_t* bar_gd = bar_pds + WSB_SLOTS;        // dV, dK, dQ are in TMEM; the operand tiles are free
  uint64_t* bar_staged = bar_gd + WSB_SLOTS;     // gradient tiles staged; TMEM of the slot fully read
  uint64_t* bar_xfree = bar_staged + WSB_SLOTS;  // the staged tiles have left shared memory (X may be rewritten)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_xfree + WSB_SLOTS);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_launch_dependents();   // see launch_pdl (common.cuh): the set-up below overlaps the previous kernel's tail
  if (tid == 0) {
    tma_prefetch_desc(&tma_qkv);
    tma_prefetch_desc(&tma_do);
    tma_prefetch_desc(&tma_dqkv);
    for (int s = 0; s < WSB_SLOTS; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_sc[s], 1);
      mbar_init(&bar_pds[s], 256);
      mbar_init(&bar_gd[s], 1);
      mbar_init(&bar_staged[s], 1);
      mbar_init(&bar_xfree[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();   // first global-memory access comes after this point

  const int items = p.B * p.H;
  const int n_local = (items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA warp: operand loads
    // (the stores have their own warp: waiting for a store's shared-memory reads used to hold back the next load by
    // 2-3 us per head -- measured with B200MM_ATTN_TRACE)
    if (elect_one()) {
      for (int nl = 0; nl < n_local; ++nl) {
        const int s = nl & 1, k = nl >> 1;
        // operand tiles (and P, which overlays V) of the slot's previous head are free once its gradient MMAs retired
        if (k > 0) mbar_wait(&bar_gd[s], (k - 1) & 1);
        const int item = blockIdx.x + nl * gridDim.x;
        const int b = item / p.H, h = item - b * p.H;
        uint8_t* base = slots + s * WSB_SLOT_BYTES;
        mbar_expect_tx(&bar_full[s], 4 * ATT_TILE_BYTES);
        tma_load_3d(base, &tma_qkv, &bar_full[s], h * ATT_D, 0, b);
        tma_load_3d(base + ATT_TILE_BYTES, &tma_qkv, &bar_full[s], p.D + h * ATT_D, 0, b);
        tma_load_3d(base + 2 * ATT_TILE_BYTES, &tma_do, &bar_full[s], h * ATT_D, 0, b);
        tma_load_3d(base + 3 * ATT_TILE_BYTES, &tma_qkv, &bar_full[s], 2 * p.D + h * ATT_D, 0, b);
        trace_event(p, nl, 0);
      }
    }
  } else if (warp == WSB_STORE_WARP) {
    // ------------------------------------------------------------ store warp: staged dQ | dK | dV tiles -> global
    if (elect_one()) {
      for (int nst = 0; nst < n_local; ++nst) {
        const int s = nst & 1;
        mbar_wait(&bar_staged[s], (nst >> 1) & 1);
        const int item = blockIdx.x + nst * gridDim.x;
        const int b = item / p.H, h = item - b * p.H;
        uint8_t* x = slots + s * WSB_SLOT_BYTES + 4 * ATT_TILE_BYTES;
#pragma unroll
        for (int t = 0; t < 3; ++t) tma_store_3d(&tma_dqkv, x + t * ATT_TILE_BYTES, t * p.D + h * ATT_D, 0, b);
        tma_store_commit();
        trace_event(p, nst, 6);
        tma_store_wait_read<0>();
        trace_event(p, nst, 7);
        mbar_arrive(&bar_xfree[s]);
      }
      tma_store_wait_read<0>();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA warp
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_tt = umma_idesc_bf16(128, 64, 1, 1);
      const uint32_t idesc_nt = umma_idesc_bf16(128, 64, 0, 1);
      int nq = 0, ng = 0;
      PollWatchdog dog;
      dog_progress(dog);
      while (ng < n_local) {
        dog_check(dog, "bwd MMA warp");
        if (nq < n_local) {
          const int s = nq & 1, k = nq >> 1;
          // operands landed AND the slot's TMEM (dV / dK / dQ of its previous head) has been drained
          if (mbar_try_wait(&bar_full[s], k & 1) && (k == 0 || mbar_try_wait(&bar_staged[s], (k - 1) & 1))) {
            tc_fence_after_sync();
            const uint32_t q = smem_u32(slots + s * WSB_SLOT_BYTES), kk_ = q + ATT_TILE_BYTES,
                           g = q + 2 * ATT_TILE_BYTES, v = q + 3 * ATT_TILE_BYTES;
            const uint32_t t_S = tmem + s * 256, t_dP = t_S + 128;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16(t_S, umma_desc_sw128(q + kk * 32, 16, 1024), umma_desc_sw128(kk_ + kk * 32, 16, 1024), idesc_s,
                        kk > 0);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16(t_dP, umma_desc_sw128(g + kk * 32, 16, 1024), umma_desc_sw128(v + kk * 32, 16, 1024), idesc_s,
                        kk > 0);
            umma_commit(&bar_sc[s]);
            trace_event(p, nq, 1);
            ++nq;
            dog_progress(dog);
          }
        }
        if (ng < nq) {
          const int s = ng & 1;
          if (mbar_try_wait(&bar_pds[s], (ng >> 1) & 1)) {
            tc_fence_after_sync();
            const uint32_t q = smem_u32(slots + s * WSB_SLOT_BYTES), kk_ = q + ATT_TILE_BYTES,
                           g = q + 2 * ATT_TILE_BYTES, sP = q + 3 * ATT_TILE_BYTES, sdS = q + 5 * ATT_TILE_BYTES;
            const uint32_t t_dV = tmem + s * 256, t_dK = t_dV + 64, t_dQ = t_dV + 128;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)   // dV = P^T dO
              umma_bf16(t_dV, umma_desc_sw128(sP + kk * 2048, ATT_TILE_BYTES, 1024),
                        umma_desc_sw128(g + kk * 2048, 8192, 1024), idesc_tt, kk > 0);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)   // dK = dS^T Q
              umma_bf16(t_dK, umma_desc_sw128(sdS + kk * 2048, ATT_TILE_BYTES, 1024),
                        umma_desc_sw128(q + kk * 2048, 8192, 1024), idesc_tt, kk > 0);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)   // dQ = dS K
              umma_bf16(t_dQ, umma_desc_sw128(sdS + (kk >> 2) * ATT_TILE_BYTES + (kk & 3) * 32, 16, 1024),
                        umma_desc_sw128(kk_ + kk * 2048, 8192, 1024), idesc_nt, kk > 0);
            umma_commit(&bar_gd[s]);
            trace_event(p, ng, 4);
            ++ng;
            dog_progress(dog);
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax warpgroups (256 threads) of slot s
    const int s = (warp - 2) >> 3;
    const int wt = tid - 64 - s * 256;                     // 0..255 inside the slot's threads
    const int half = ((warp - 2) >> 2) & 1;                // which 64 of the 128 key columns
    const int row = (warp & 3) * 32 + lane;                // TMEM lane quarter = warp % 4
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t t_S = tmem + s * 256, t_dP = t_S + 128;
    uint8_t* base = slots + s * WSB_SLOT_BYTES;
    uint8_t* sP = base + 3 * ATT_TILE_BYTES;               // V | X0
    uint8_t* sdS = base + 5 * ATT_TILE_BYTES;              // X1 | X2
    uint8_t* sX = base + 4 * ATT_TILE_BYTES;               // staged dQ | dK | dV
    float* bias = sBias + s * ATT_T;
    const bool use_drop = p.p_drop > 0.f;
    const float scale = p.scale;
    float lse_next = INFINITY, bias_next = -INFINITY;
    uint32_t keep_next0 = 0xffffffffu, keep_next1 = 0xffffffffu;
    uint4 o_next[4];
    auto fetch_row_scalars = [&](int n2) {
      const int item2 = blockIdx.x + n2 * gridDim.x;
      const int b2 = item2 / p.H, h2 = item2 - b2 * p.H;
      {
        const uint4* po = reinterpret_cast<const uint4*>(
            p.o_in + (static_cast<long long>(b2) * p.S + row) * p.D + h2 * ATT_D + half * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) o_next[q] = row < p.S ? __ldg(po + q) : make_uint4(0u, 0u, 0u, 0u);
      }
      lse_next = row < p.S ? __ldg(p.lse + static_cast<long long>(item2) * p.S + row) * LOG2E : INFINITY;
      if (wt < ATT_T) bias_next = wt < p.S ? (p.key_bias ? __ldg(p.key_bias + b2 * p.S + wt) * LOG2E : 0.f) : -INFINITY;
      if (use_drop && p.drop_mask != nullptr) {      // keep bits of this thread's 64 keys, saved by the forward
        const uint32_t* pm = p.drop_mask + (static_cast<long long>(item2) * 4 + half * 2) * ATT_T + row;
        keep_next0 = __ldg(pm);
        keep_next1 = __ldg(pm + ATT_T);
      }
    };
    if (s < n_local) fetch_row_scalars(s);
    for (int n = s; n < n_local; n += WSB_SLOTS) {
      const int k = n >> 1;
      const int item = blockIdx.x + n * gridDim.x;
      // ---- per-row scalars: delta = rowsum(dO o O), log2-domain LSE (global loads overlap the TMA + score MMAs)
      // per-row scalars of this head were fetched while the previous head of the slot was in its gradient MMAs
      const float lse_l2 = lse_next;
      const uint32_t keep_saved0 = keep_next0, keep_saved1 = keep_next1;
      if (wt < ATT_T) bias[wt] = bias_next;
      // delta = rowsum(dO o O): the two threads of a row each take 32 of the 64 head-dim columns -- O from the
      // registers fetched two heads ago (so its 3-5 us of global latency is off this path), dO from the TMA-loaded tile
      // -- and exchange through shared memory: half 1 posts its partial, half 0 adds its own and posts the total.
      // (Former versions: 16 row-strided global loads per thread right here, 4-6 us per head; an O tile loaded by TMA
      // into X1, which had to wait for the previous head's stores to leave, 1.2 us.)
      const uint4 o0 = o_next[0], o1 = o_next[1], o2 = o_next[2], o3 = o_next[3];
      mbar_wait(&bar_full[s], k & 1);
      float delta;
      {
        const uint8_t* tG = base + 2 * ATT_TILE_BYTES;
        const uint4 g0 = *reinterpret_cast<const uint4*>(tG + sw128_off(row, half * 4 + 0));
        const uint4 g1 = *reinterpret_cast<const uint4*>(tG + sw128_off(row, half * 4 + 1));
        const uint4 g2 = *reinterpret_cast<const uint4*>(tG + sw128_off(row, half * 4 + 2));
        const uint4 g3 = *reinterpret_cast<const uint4*>(tG + sw128_off(row, half * 4 + 3));
        float2 d2 = fmul2(unpack_bf16x2(o0.x), unpack_bf16x2(g0.x));
        d2 = ffma2(unpack_bf16x2(o0.y), unpack_bf16x2(g0.y), d2);
        d2 = ffma2(unpack_bf16x2(o0.z), unpack_bf16x2(g0.z), d2);
        d2 = ffma2(unpack_bf16x2(o0.w), unpack_bf16x2(g0.w), d2);
        d2 = ffma2(unpack_bf16x2(o1.x), unpack_bf16x2(g1.x), d2);
        d2 = ffma2(unpack_bf16x2(o1.y), unpack_bf16x2(g1.y), d2);
        d2 = ffma2(unpack_bf16x2(o1.z), unpack_bf16x2(g1.z), d2);
        d2 = ffma2(unpack_bf16x2(o1.w), unpack_bf16x2(g1.w), d2);
        d2 = ffma2(unpack_bf16x2(o2.x), unpack_bf16x2(g2.x), d2);
        d2 = ffma2(unpack_bf16x2(o2.y), unpack_bf16x2(g2.y), d2);
        d2 = ffma2(unpack_bf16x2(o2.z), unpack_bf16x2(g2.z), d2);
        d2 = ffma2(unpack_bf16x2(o2.w), unpack_bf16x2(g2.w), d2);
        d2 = ffma2(unpack_bf16x2(o3.x), unpack_bf16x2(g3.x), d2);
        d2 = ffma2(unpack_bf16x2(o3.y), unpack_bf16x2(g3.y), d2);
        d2 = ffma2(unpack_bf16x2(o3.z), unpack_bf16x2(g3.z), d2);
        d2 = ffma2(unpack_bf16x2(o3.w), unpack_bf16x2(g3.w), d2);
        delta = d2.x + d2.y;
      }
      float* dx = sDelta + s * ATT_T;
      if (half == 1) dx[row] = delta;
      named_bar_sync(1 + s, 256);      // bias of this head visible too
      if (half == 0) {
        delta += dx[row];
        dx[row] = delta;
      }
      named_bar_sync(1 + s, 256);
      if (half == 1) delta = dx[row];
      const float delta_s = delta * scale;
      mbar_wait(&bar_sc[s], k & 1);
      tc_fence_after_sync();
      if (k > 0) mbar_wait(&bar_xfree[s], (k - 1) & 1);    // previous head's staged tiles have left X
      if (wt == 0) trace_event(p, n, 2);
      const float2 sc2 = splat2(p.scale_log2), nlse2 = splat2(-lse_l2), scl2 = splat2(scale), nds2 = splat2(-delta_s);
      // ---- P = exp2(S - LSE), dS = P o (dP - delta) * scale for this thread's 64 key columns
      uint32_t keep = 0xffffffffu;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int col0 = half * 64 + cc * 16;
        uint32_t vs[16], vp[16];
        tmem_ld16(t_S + lane_addr + col0, vs);
        tmem_ld16(t_dP + lane_addr + col0, vp);
        tmem_ld_wait();
        if (use_drop && (cc & 1) == 0)
          keep = p.drop_mask != nullptr
                     ? (cc >> 1 ? keep_saved1 : keep_saved0)
                     : dropout_keep32(p.seed, (static_cast<uint64_t>(item) * ATT_T + row) * 4 + (col0 >> 5),
                                      p.drop_threshold >> 16);   // as the forward
        const uint32_t kb = use_drop ? (keep >> ((cc & 1) * 16)) : 0xffffu;
        uint32_t pk[8], dk[8];
#pragma unroll
        for (int e = 0; e < 16; e += 2) {      // packed fp32: both elements of a pair share every FMA-pipe issue slot
          const float2 b2 = *reinterpret_cast<const float2*>(bias + col0 + e);
          const float2 t = fadd2(ffma2(make_float2(__uint_as_float(vs[e]), __uint_as_float(vs[e + 1])), sc2, b2), nlse2);
          const float2 prob = make_float2(fast_exp2(t.x), fast_exp2(t.y));
          float2 pr = prob, dp = make_float2(__uint_as_float(vp[e]), __uint_as_float(vp[e + 1]));
          if (use_drop) {
            const float2 m = make_float2((kb >> e) & 1 ? p.inv_keep : 0.f, (kb >> (e + 1)) & 1 ? p.inv_keep : 0.f);
            pr = fmul2(prob, m);
            dp = fmul2(dp, m);
          }
          const float2 ds = fmul2(prob, ffma2(dp, scl2, nds2));
          pk[e >> 1] = pack_bf16x2(pr.x, pr.y);
          dk[e >> 1] = pack_bf16x2(ds.x, ds.y);
        }
        // 16 columns = two 16-byte chunks of this row in the [128 x 128] tile (two [128 x 64] swizzled blocks)
        uint8_t* pblk = sP + (col0 >> 6) * ATT_TILE_BYTES;
        uint8_t* dblk = sdS + (col0 >> 6) * ATT_TILE_BYTES;
        const int chunk0 = (col0 & 63) >> 3;
        *reinterpret_cast<uint4*>(pblk + sw128_off(row, chunk0)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(pblk + sw128_off(row, chunk0 + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        *reinterpret_cast<uint4*>(dblk + sw128_off(row, chunk0)) = make_uint4(dk[0], dk[1], dk[2], dk[3]);
        *reinterpret_cast<uint4*>(dblk + sw128_off(row, chunk0 + 1)) = make_uint4(dk[4], dk[5], dk[6], dk[7]);
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(&bar_pds[s]);
      if (wt == 0) trace_event(p, n, 3);
      if (n + WSB_SLOTS < n_local) fetch_row_scalars(n + WSB_SLOTS);     // latency hides under the gradient MMAs + drain
      // ---- drain: half 0 -> dQ (64 columns) + dK columns 0..31 ; half 1 -> dV (64 columns) + dK columns 32..63
      mbar_wait(&bar_gd[s], k & 1);
      tc_fence_after_sync();
#pragma unroll 1
      for (int piece = 0; piece < 3; ++piece) {
        // TMEM: dV = S[0,64), dK = S[64,128), dQ = dP[0,64);  staged tiles: 0 = dQ, 1 = dK, 2 = dV
        const uint32_t t_src = piece < 2 ? ((half == 0 ? t_dP : t_S) + piece * 32) : (t_S + 64 + half * 32);
        const int tile = piece < 2 ? (half == 0 ? 0 : 2) : 1;
        const int chunk0 = (piece < 2 ? piece : half) * 4;
        uint32_t v[32];
        tmem_ld32(t_src + lane_addr, v);
        tmem_ld_wait();
        uint8_t* dst = sX + tile * ATT_TILE_BYTES;
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
      named_bar_sync(1 + s, 256);      // tiles complete, TMEM of the slot read out by all 256 threads
      if (wt == 0) {
        mbar_arrive(&bar_staged[s]);
        trace_event(p, n, 5);
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem);
  }
}

int launch_attn_fwd_ws(const CUtensorMap& tma_qkv, const CUtensorMap& tma_out, const AttnParams& p, int num_sms,
                       cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WSF_SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  const int items = p.B * p.H;
  const int grid = items < num_sms ? items : num_sms;
  cudaError_t le = launch_pdl(attn_fwd_ws_kernel, dim3(grid), dim3(WSF_THREADS), WSF_SMEM, stream, tma_qkv, tma_out, p);
  if (le != cudaSuccess) return static_cast<int>(le);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

// B200MM_ATTN_TRACE=1: run the backward with CTA 0 time-stamping its hand-over points and print, per head of that
// CTA, the event times relative to the head's load issue (debugging aid; synchronises the stream).
static int trace_bwd_launch(const CUtensorMap& tma_qkv, const CUtensorMap& tma_do, const CUtensorMap& tma_dqkv,
                            AttnParams p, int grid, cudaStream_t stream) {
  const int heads = (p.B * p.H + grid - 1) / grid;
  unsigned long long* dev = nullptr;
  if (cudaMalloc(&dev, sizeof(unsigned long long) * 8 * heads) != cudaSuccess) return B200MM_ERR_BAD_ARG;
  cudaMemsetAsync(dev, 0, sizeof(unsigned long long) * 8 * heads, stream);
  p.trace = dev;
  attn_bwd_ws_kernel<<<grid, WSB_THREADS, WSB_SMEM, stream>>>(tma_qkv, tma_do, tma_dqkv, p);
  cudaStreamSynchronize(stream);
  unsigned long long* host = new unsigned long long[8 * heads];
  cudaMemcpy(host, dev, sizeof(unsigned long long) * 8 * heads, cudaMemcpyDeviceToHost);
  static const char* names[8] = {"load", "scoreMMA", "sc_seen", "pds", "gradMMA", "staged", "store", "store_rd"};
  fprintf(stderr, "attn bwd trace (CTA 0, ns since the first load): ");
  for (int e = 0; e < 8; ++e) fprintf(stderr, "%s ", names[e]);
  fprintf(stderr, "\n");
  for (int h = 0; h < heads; ++h) {
    fprintf(stderr, "  head %2d:", h);
    for (int e = 0; e < 8; ++e) fprintf(stderr, " %7lld", static_cast<long long>(host[h * 8 + e] - host[0]));
    fprintf(stderr, "\n");
  }
  delete[] host;
  cudaFree(dev);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

int launch_attn_bwd_ws(const CUtensorMap& tma_qkv, const CUtensorMap& tma_do, const CUtensorMap& tma_dqkv,
                       const AttnParams& p, int num_sms, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WSB_SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  const int items = p.B * p.H;
  const int grid = items < num_sms ? items : num_sms;
  const char* tr = std::getenv("B200MM_ATTN_TRACE");
  if (tr && tr[0] == '1') return trace_bwd_launch(tma_qkv, tma_do, tma_dqkv, p, grid, stream);
  cudaError_t le = launch_pdl(attn_bwd_ws_kernel, dim3(grid), dim3(WSB_THREADS), WSB_SMEM, stream, tma_qkv, tma_do,
                              tma_dqkv, p);
  if (le != cudaSuccess) return static_cast<int>(le);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

}  // namespace b200
