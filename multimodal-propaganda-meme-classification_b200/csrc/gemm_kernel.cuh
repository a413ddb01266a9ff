// bf16 GEMM on the 5th-generation tensor cores: D[M,N] = A x B with fp32 accumulation in TMEM.
//
// One persistent CTA per SM, 10 warps:
//   warp 0      TMA producer   : cp.async.bulk.tensor -> STAGES-deep smem ring (SWIZZLE_128B)
//   warp 1      MMA issuer     : one thread issues tcgen05.mma (128 x BN x 16), commits to mbarriers
//   warps 2..9  epilogue       : tcgen05.ld accumulator -> bias / GELU / dGELU / residual -> global
// The accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the
// main loop of tile i+1.  Both operands may be K-major ("row-major with the reduction dim contiguous")
// or MN-major (reduction dim strided) -- this is what lets the same kernel serve
//   forward   Y  = X  W^T      (A K-major,  B K-major)      reference: nn.Linear / 1x1 conv / im2col conv
//   dgrad     dX = dY W        (A K-major,  B MN-major)
//   wgrad     dW = dY^T X      (A MN-major, B MN-major, split-K over the token dim, fp32 red.add)
// without transposing anything in HBM.
//
// Replaces: the cuBLASLt calls behind
// transformers/models/distilbert/modeling_distilbert.py q_lin/k_lin/v_lin/out_lin/lin1/lin2 and the
// cuDNN convolutions behind torchvision/models/resnet.py, as driven by
// example_scripts/Multimodal_example_task2C.txt:172-197.
#pragma once
#include "common.cuh"
#include "device_utils.cuh"
#include "ptx.cuh"
#include "gemm_params.cuh"

namespace b200 {

// CTAS = 2: a CTA pair (cluster of two, cta_group::2) computes a 256 x BN tile with ONE MMA stream -- each CTA stages its
// own 128 rows of A and only HALF of the B tile, so the tensor core's shared-memory reads per FLOP drop by a third
// (the single-CTA 128 x 256 tile needs 96 B/clk of the 128 B/clk shared-memory port at full MMA rate).
// EW = epilogue warps: 8 (two column halves per TMEM lane quarter).  16 (four column quarters, one staged box per warp
// and tile) was measured on the short-K convolution shapes and did not pay (-10 % at [200704 x 512 x 128], +15 % at
// [50176 x 1024 x 256], profiles/gemm_launch_cost_r02.log): only the 8-warp form is instantiated.
template <int BN, int CTAS = 1, int EW = GEMM_EPI_WARPS>
struct GemmCfg {
  static constexpr int THREADS = 64 + EW * 32;
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = (BN / CTAS) * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  // ring depth that fits next to `nbuf` staging boxes per epilogue warp and the barrier block
  static constexpr int stages_for(int nbuf, int budget = GEMM_SMEM_TOTAL) {
    const int avail = budget - EW * nbuf * GEMM_BOX_BYTES - 512;
    const int s = avail / STAGE_BYTES;
    return s > GEMM_MAX_STAGES ? GEMM_MAX_STAGES : s;
  }
};

template <int BN, int EPI, int CTAS = 1, int EW = GEMM_EPI_WARPS>
__global__ void __launch_bounds__(64 + EW * 32, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const __grid_constant__ CUtensorMap tma_out, const __grid_constant__ CUtensorMap tma_out2,
                 const GemmParams p) {
  using Cfg = GemmCfg<BN, CTAS, EW>;
  constexpr bool PAIR = CTAS == 2;
  const int cta_rank = PAIR ? static_cast<int>(cluster_ctarank()) : 0;
  const int STAGES = p.num_stages;
  // 1024-byte alignment (SWIZZLE_128B atoms) by pointer arithmetic ON the __shared__ array: the compiler keeps the
  // address space and emits LDS / STS (the former round-up through uintptr_t turned every access of the tiles,
  // the staging boxes and the bias rows into generic LD.E / ST.E, which queue with the global loads)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // B-resident mode (single CTA, short unsplit K, every tile of this CTA in the same column block): the whole B tile
  // [BN x K] is loaded ONCE into a resident area and the ring carries A only (a 128 x 256 x 256 tile then moves 64 KB
  // in + 64 KB out instead of 192 KB in).  Measured on the ResNet 1x1-convolution shapes (L2 flushed): 39.9 -> 37.9 us
  // at [50176 x 1024 x 256], 52.2 -> 38.9 us at the ragged [50000 x 1000 x 200], unchanged elsewhere
  // (profiles/conv_gemm_variants_r02.log).  What paces these shapes is the epilogue / store side: with the epilogue
  // switched off the same launches take 18 us instead of 33, with the MMAs switched off 29 us
  // (profiles/gemm_launch_cost_r02.log); ring depth, CTA pairs, 16 epilogue warps and an L2 prefetch of the next A
  // tile all left the time within 5 %.
  const bool BRES = !PAIR && p.b_resident != 0;
  const int stage_stride = BRES ? Cfg::A_BYTES : Cfg::STAGE_BYTES;
  uint8_t* ring = smem + (BRES ? p.k_iters * Cfg::B_BYTES : 0);
  uint8_t* staging = ring + STAGES * stage_stride;  // 1024-aligned (A_BYTES / B_BYTES are multiples of 1024)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + EW * p.nbuf * GEMM_BOX_BYTES);
  uint64_t* empty_bar = full_bar + GEMM_MAX_STAGES;
  uint64_t* tmem_full = empty_bar + GEMM_MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // programmatic dependent launch: the next kernel on the stream may be scheduled as this one drains, and this CTA's
  // set-up (barriers, TMEM allocation) runs while the previous kernel drains; nothing below touches global memory
  // before pdl_wait()
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], CTAS * EW);   // pair: the leader's copy collects both CTAs' epilogue warps
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot);
    else tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  tc_fence_before_sync();
  if constexpr (PAIR) cluster_sync_all();   // both CTAs' barriers exist before anything may signal them remotely
  else __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // the previous kernel's outputs (this kernel's operands, residual, statistics accumulators) are complete

  // work units: output tiles (single CTA) or pairs of vertically adjacent tiles (CTA pair: rank r takes tile 2u + r)
  const int m_units = PAIR ? (p.m_tiles + 1) / 2 : p.m_tiles;
  const int tiles_mn = m_units * p.n_tiles;
  const int num_work = tiles_mn * p.splits;
  const int w_first = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int w_step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = w_first; w < num_work; w += w_step) {
        const int split = w / tiles_mn;
        const int t = w - split * tiles_mn;
        const int m_unit = t / p.n_tiles;
        const int n_blk = t - m_unit * p.n_tiles;
        const int m_blk = PAIR ? 2 * m_unit + cta_rank : m_unit;
        const int k_begin = split * p.k_iters_per_split;
        const int k_end = min(k_begin + p.k_iters_per_split, p.k_iters);
        int a_n0 = 0, a_p0 = 0, a_q0 = 0;   // first output pixel of this M tile (im2col A operand)
        int a_cb = 0, a_kw = 0, a_kh = 0;   // channel slab / filter tap of the current reduction block
        if (p.a_im2col) {
          const int pq = p.conv_P * p.conv_Q;
          const int m0 = m_blk * GEMM_BM;
          a_n0 = m0 / pq;
          const int r0 = m0 - a_n0 * pq;
          a_p0 = r0 / p.conv_Q;
          a_q0 = r0 - a_p0 * p.conv_Q;
          const int tap = k_begin / p.conv_cblocks;
          a_cb = k_begin - tap * p.conv_cblocks;
          a_kh = tap / p.conv_KW;
          a_kw = tap - a_kh * p.conv_KW;
        }
        // im2col B operand (implicit-GEMM weight gradient): the (tap, channel) of each 64-column atom is fixed for the
        // work unit and the pixel coordinates of the reduction block advance by 64 per iteration -- no divisions in
        // the k loop (with ten of them per iteration this thread, not the tensor pipe, paced the layer 2-4 wgrads)
        int b_c0[BN / 64], b_kw[BN / 64], b_kh[BN / 64], b_atoms = 0;
        int b_n0 = 0, b_p0 = 0, b_q0 = 0;
        if (p.b_im2col) {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) {
            const int col = n_blk * BN + j * 64;
            const int tap = col / p.conv_C;
            b_c0[j] = col - tap * p.conv_C;
            b_kh[j] = tap / p.conv_KW;
            b_kw[j] = tap - b_kh[j] * p.conv_KW;
            if (col < p.N) b_atoms = j + 1;
          }
          const int pq = p.conv_P * p.conv_Q;
          const int pix = k_begin * GEMM_BK;
          b_n0 = pix / pq;
          const int r0 = pix - b_n0 * pq;
          b_p0 = r0 / p.conv_Q;
          b_q0 = r0 - b_p0 * p.conv_Q;
        }
        const bool load_b = !BRES || w == w_first;   // resident B: k block kb of the first tile fills slot kb
        for (int kb = k_begin; kb < k_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = ring + stage * stage_stride;
          uint8_t* sb = BRES ? smem + kb * Cfg::B_BYTES : sa + Cfg::A_BYTES;
          if (p.b_im2col) {
            // wgrad: only the 64-column atoms that exist (N = taps * C may end inside the tile) are loaded
            mbar_expect_tx(&full_bar[stage], Cfg::A_BYTES + b_atoms * 8192);
          } else if (PAIR) {
            if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], CTAS * Cfg::STAGE_BYTES);   // both CTAs' loads count on the leader
          } else {
            mbar_expect_tx(&full_bar[stage], Cfg::A_BYTES + (load_b ? Cfg::B_BYTES : 0));
          }
          if constexpr (PAIR) {
            // (matrix operands only: the convolution shapes stay on the single-CTA kernel)
            const int n0 = n_blk * BN + cta_rank * (BN / 2);
            if (!p.a_mn) {
              tma_load_2d_pair(sa, &tma_a, &full_bar[stage], kb * GEMM_BK, m_blk * GEMM_BM);
            } else {
#pragma unroll
              for (int j = 0; j < GEMM_BM / 64; ++j)
                tma_load_2d_pair(sa + j * 8192, &tma_a, &full_bar[stage], m_blk * GEMM_BM + j * 64, kb * GEMM_BK);
            }
            if (!p.b_mn) {
              tma_load_2d_pair(sb, &tma_b, &full_bar[stage], kb * GEMM_BK, n0);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 128; ++j)
                tma_load_2d_pair(sb + j * 8192, &tma_b, &full_bar[stage], n0 + j * 64, kb * GEMM_BK);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          if (p.a_im2col) {
            tma_load_im2col_4d(sa, &tma_a, &full_bar[stage], a_cb * 64, a_q0 * p.conv_stride - p.conv_pad,
                               a_p0 * p.conv_stride - p.conv_pad, a_n0, static_cast<uint16_t>(a_kw),
                               static_cast<uint16_t>(a_kh));
            if (++a_cb == p.conv_cblocks) {               // next reduction block: next channel slab, then next tap
              a_cb = 0;
              if (++a_kw == p.conv_KW) { a_kw = 0; ++a_kh; }
            }
          } else if (!p.a_mn) {
            tma_load_2d(sa, &tma_a, &full_bar[stage], kb * GEMM_BK, m_blk * GEMM_BM);
          } else {
#pragma unroll
            for (int j = 0; j < GEMM_BM / 64; ++j)
              tma_load_2d(sa + j * 8192, &tma_a, &full_bar[stage], m_blk * GEMM_BM + j * 64, kb * GEMM_BK);
          }
          if (p.b_im2col) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) {
              if (j < b_atoms)
                tma_load_im2col_4d(sb + j * 8192, &tma_b, &full_bar[stage], b_c0[j],
                                   b_q0 * p.conv_stride - p.conv_pad, b_p0 * p.conv_stride - p.conv_pad, b_n0,
                                   static_cast<uint16_t>(b_kw[j]), static_cast<uint16_t>(b_kh[j]));
            }
            b_q0 += GEMM_BK;                              // first output pixel of the next reduction block
            while (b_q0 >= p.conv_Q) { b_q0 -= p.conv_Q; ++b_p0; }
            while (b_p0 >= p.conv_P) { b_p0 -= p.conv_P; ++b_n0; }
          } else if (!load_b) {
            // resident B already holds this k block
          } else if (!p.b_mn) {
            tma_load_2d(sb, &tma_b, &full_bar[stage], kb * GEMM_BK, n_blk * BN);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sb + j * 8192, &tma_b, &full_bar[stage], n_blk * BN + j * 64, kb * GEMM_BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // pair: only the leader issues (and commits to both CTAs' barriers).  elect_one(), not lane == 0: see ptx.cuh
    if (cta_rank == 0 && elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(CTAS * GEMM_BM, BN, p.a_mn, p.b_mn);
      // K-major: 16-element k step = 32 B inside the swizzle row; 8-row groups 1024 B apart.
      // MN-major: 16-element k step = two 8-k-row groups = 2048 B; 64-wide MN atoms one 8 KB box apart.
      const uint32_t a_kstep = p.a_mn ? 2048u : 32u, b_kstep = p.b_mn ? 2048u : 32u;
      const uint32_t a_lbo = p.a_mn ? 8192u : 16u, b_lbo = p.b_mn ? 8192u : 16u;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = w_first; w < num_work; w += w_step) {
        const int split = w / tiles_mn;
        const int k_begin = split * p.k_iters_per_split;
        const int k_end = min(k_begin + p.k_iters_per_split, p.k_iters);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = k_begin; kb < k_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(ring + stage * stage_stride);
          const uint32_t sb = BRES ? smem_u32(smem + kb * Cfg::B_BYTES) : sa + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / GEMM_UMMA_K; ++k) {
            const uint64_t da = umma_desc_sw128(sa + k * a_kstep, a_lbo, 1024);
            const uint64_t db = umma_desc_sw128(sb + k * b_kstep, b_lbo, 1024);
            if constexpr (PAIR) umma_bf16_pair(d_tmem, da, db, idesc, (kb > k_begin || k > 0) ? 1u : 0u);
            else umma_bf16(d_tmem, da, db, idesc, (kb > k_begin || k > 0) ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs have read it
          if constexpr (PAIR) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if constexpr (PAIR) umma_commit_pair(&tmem_full[acc]); else umma_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    // bf16 outputs: TMEM -> registers -> (bias / activation / residual) -> 128B-swizzled smem box -> TMA store
    // (fully coalesced global writes, rows/columns beyond M/N clipped by the tensor map).
    // fp32 outputs (wgrad): direct vector stores / red.global.add straight from registers.
    // The mode is a template parameter: one lean instruction stream per epilogue (the all-modes-in-one version was
    // 4096 SASS instructions, missed the instruction cache and issue-bound the short-K convolution GEMMs).
    const int ew = warp - 2;
    const int quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int half = ew >> 2;              // which column group (half, or quarter with 16 warps) of the BN columns
    constexpr int HALF_COLS = BN / (EW / 4);
    constexpr bool BF16_OUT = EPI != EPI_F32 && EPI != EPI_F32_ATOMIC;
    // store modes that may add a residual: plain / ReLU / dropout, and the ReLU-masked residual of the identity branch
    constexpr bool RES_EPI = EPI == EPI_STORE || EPI == EPI_RELU || EPI == EPI_STORE_DROP || EPI == EPI_STORE_MASKRES;
    int acc = 0;
    uint32_t acc_phase = 0;
    int buf = 0;
    int sv_tile = -1, sv_box = -1;   // which (tile, box) the prefetched side-input registers hold
    // geometry of one staged epilogue box (bf16 outputs)
    constexpr int BOX_W = HALF_COLS < 64 ? HALF_COLS : 64;   // columns per staged box: 64 (SW128) or 32 (SW64)
    constexpr int BOXES = HALF_COLS / BOX_W;
    // Side input of the epilogue (residual for the store modes, the saved pre-activation for dGELU): one
    // [32 rows x BOX_W cols] box per staged output box.  It is read with fully coalesced 16-byte loads (a row
    // segment per 8 / 4 lanes), one box AHEAD of its use so the HBM latency hides under the previous box, and
    // handed to the row-owning lanes through the staging buffer the output is about to overwrite.  (The earlier
    // per-lane row-strided __ldg pattern missed the 28 KB L1 left next to 227 KB of shared memory and made the
    // dGELU / residual epilogues latency-bound.)
    // Only dGELU takes this path: for the residual of the store modes the plain per-lane loads measured faster on the
    // short-K convolution GEMMs (241 vs 283 us at [802816 x 256 x 64]) and equal on the long-K ones, and keeping
    // them out keeps the store kernels' instruction stream short (they are I-cache sensitive).
    constexpr bool HAS_SIDE = EPI == EPI_DGELU;
    constexpr int CPR = BOX_W / 8;            // 16-byte chunks per box row
    constexpr int RPP = 32 / CPR;             // rows covered by the warp per load pass
    constexpr int SIDE_REGS = 32 / RPP;       // passes (uint4 registers) per box
    const __nv_bfloat16* side = nullptr;
    long long side_ld = 0;
    if constexpr (EPI == EPI_DGELU) { side = p.aux; side_ld = p.ld_aux; }
    uint4 sv[SIDE_REGS];
    // tile coordinates advance by a constant (gridDim.x tiles) per round: no integer divisions on the prefetch path
    // (side inputs exist only for the bf16 store modes, which never split K: tile index == w)
    const int step_m = w_step / p.n_tiles, step_n = w_step % p.n_tiles;
    auto advance = [&](int& mu, int& nb) {   // in work units (a unit = one tile, or one tile pair)
      mu += step_m;
      nb += step_n;
      if (nb >= p.n_tiles) { nb -= p.n_tiles; ++mu; }
    };
    auto tile_row = [&](int mu) { return PAIR ? 2 * mu + cta_rank : mu; };
    auto side_fetch = [&](int mb, int nb, int b) {
      const int col = nb * BN + half * HALF_COLS + b * BOX_W + (lane % CPR) * 8;
      const long long r0 = static_cast<long long>(mb) * GEMM_BM + quarter * 32 + lane / CPR;
      const __nv_bfloat16* src = side + r0 * side_ld + col;
#pragma unroll
      for (int j = 0; j < SIDE_REGS; ++j) {
        sv[j] = (r0 + j * RPP < p.M && col < p.N)
                    ? ldg_stream_v4(src + static_cast<long long>(j * RPP) * side_ld)
                    : make_uint4(0u, 0u, 0u, 0u);
      }
    };
    const bool has_bias = p.bias != nullptr;
    // EPI_STORE_STATS: per-lane running column sums (two adjacent columns per staged box), flushed with fp32
    // reductions when the CTA moves to another column block and at the end
    // EPI_DGELU with p.col_stats: column SUMS only, into col_stats[0..N) -- the bias gradient of the linear layer whose
    // output gradient this GEMM produces (the separate column-sum pass re-read the whole [M, N] tensor)
    constexpr bool COLSUM_EPI = EPI == EPI_STORE_STATS || EPI == EPI_DGELU;
    const bool do_colsum = EPI == EPI_STORE_STATS || (EPI == EPI_DGELU && p.col_stats != nullptr);
    float st_s[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, st_q[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    int st_nblk = -1;
    auto stats_flush = [&]() {
      if (st_nblk < 0) return;
      if (BOX_W == 32 && lane >= 16) return;
#pragma unroll
      for (int b = 0; b < BOXES; ++b) {
        const int col = st_nblk * BN + half * HALF_COLS + b * BOX_W + 2 * (BOX_W == 64 ? lane : (lane & 15));
        if (col < p.N) {
          asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p.col_stats + col), "f"(st_s[b][0]), "f"(st_s[b][1]) : "memory");
          if constexpr (EPI == EPI_STORE_STATS)
            asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p.col_stats + p.N + col), "f"(st_q[b][0]), "f"(st_q[b][1]) : "memory");
        }
        st_s[b][0] = st_s[b][1] = st_q[b][0] = st_q[b][1] = 0.f;
      }
    };
    // (m_unit, n_blk) of the work unit advance incrementally: the two runtime divisions per tile were 12 % of the
    // epilogue warps' samples on the short-K shapes (K split never changes the tile an epilogue warp addresses)
    int m_unit, n_blk;
    {
      const int t0 = w_first % tiles_mn;
      m_unit = t0 / p.n_tiles;
      n_blk = t0 - m_unit * p.n_tiles;
    }
    for (int w = w_first; w < num_work; w += w_step) {
      const int m_blk = tile_row(m_unit);
      if constexpr (COLSUM_EPI) {
        if (do_colsum && n_blk != st_nblk) { stats_flush(); st_nblk = n_blk; }
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after_sync();
      const long long row = static_cast<long long>(m_blk) * GEMM_BM + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
      if constexpr (BF16_OUT) {
        constexpr int CHUNKS_PER_BOX = BOX_W / 32;
        constexpr int PASSES = 1;
        uint8_t* stage_base = staging + ew * (p.nbuf * GEMM_BOX_BYTES);
        const __nv_bfloat16* res_row = nullptr;
        if constexpr (RES_EPI)
          if (p.residual != nullptr && row_ok) res_row = p.residual + row * p.ldr;
        if constexpr (HAS_SIDE) {
          if (side != nullptr && (sv_tile != w || sv_box != 0)) { side_fetch(m_blk, n_blk, 0); sv_tile = w; sv_box = 0; }
        }
#pragma unroll 1
        for (int b = 0; b < BOXES; ++b) {
          const int box_col_in_tile = half * HALF_COLS + b * BOX_W;
          const int box_col0 = n_blk * BN + box_col_in_tile;
          if (box_col0 >= p.N) break;
#pragma unroll 1
          for (int pass = 0; pass < PASSES; ++pass) {
            // the store issued from this buffer `nbuf` boxes ago has left shared memory
            if (lane == 0) {
              if (EPI != EPI_GELU && p.nbuf == 2) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
            }
            __syncwarp();
            // EPI_GELU: both boxes of the warp per output box (z -> box 0, gelu(z) -> box 1), filled from ONE read of
            // the accumulator and stored by one bulk group
            uint8_t* stage_buf = stage_base + (EPI == EPI_GELU ? 0 : buf) * GEMM_BOX_BYTES;
            if constexpr (HAS_SIDE) {
              if (side != nullptr) {
#pragma unroll
                for (int j = 0; j < SIDE_REGS; ++j) {
                  const int r = j * RPP + lane / CPR, ch = lane % CPR;
                  const uint32_t off = BOX_W == 64 ? r * 128 + ((ch ^ (r & 7)) << 4)
                                                   : r * 64 + ((ch ^ ((r >> 1) & 3)) << 4);
                  *reinterpret_cast<uint4*>(stage_buf + off) = sv[j];
                }
                // next box of this tile, or the first box of this CTA's next tile
                const bool more = b + 1 < BOXES && n_blk * BN + half * HALF_COLS + (b + 1) * BOX_W < p.N;
                int m1 = m_unit, n1 = n_blk;
                advance(m1, n1);
                {
                  // ... and pull this box of the tile two rounds ahead into L2 (one 128-byte line per lane = one
                  // box row): the register prefetch above then hits L2 instead of waiting ~2 us on HBM.
                  int m2 = m1, n2 = n1;
                  advance(m2, n2);
                  const long long r = static_cast<long long>(tile_row(m2)) * GEMM_BM + quarter * 32 + lane;
                  const int col = n2 * BN + box_col_in_tile;
                  if (w + 2 * w_step < num_work && r < p.M && col < p.N)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(side + r * side_ld + col));
                }
                if (more) { side_fetch(m_blk, n_blk, b + 1); sv_tile = w; sv_box = b + 1; }
                else if (w + w_step < num_work) {
                  side_fetch(tile_row(m1), n1, 0); sv_tile = w + w_step; sv_box = 0;
                }
                __syncwarp();
              }
            }
#pragma unroll 1
            for (int c = 0; c < CHUNKS_PER_BOX; ++c) {
              // residual of this lane's row for the whole 32-column chunk: four independent 16-byte loads issued
              // BEFORE the accumulator read, so their latency overlaps it and each other (one load per 8 columns
              // right before its use serialised four L2 round trips per chunk: 86 vs 40 us at [32768 x 768 x 768])
              uint4 rres[4];
              if constexpr (RES_EPI) {
                if (res_row != nullptr) {
#pragma unroll
                  for (int g = 0; g < 4; ++g) {
                    const int col = box_col0 + c * 32 + g * 8;
                    rres[g] = col < p.N ? __ldg(reinterpret_cast<const uint4*>(res_row + col)) : make_uint4(0u, 0u, 0u, 0u);
                  }
                }
              }
              // EPI_STORE_MASKRES: 1 bit per element of this lane's row (bit set = the ReLU behind the residual's
              // producer passed): 32 columns = one 4-byte word
              uint32_t rmask = 0xffffffffu;
              if constexpr (EPI == EPI_STORE_MASKRES) {
                const int col = box_col0 + c * 32;
                rmask = (row_ok && col < p.N) ? __ldg(reinterpret_cast<const uint32_t*>(p.res_mask + row * p.ld_mask + (col >> 3)))
                                              : 0u;
              }
              uint32_t v[32];
              tmem_ld32(t_base + box_col_in_tile + c * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const int col = box_col0 + c * 32 + g * 8;
                const bool col_ok = col < p.N;
                float x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(v[g * 8 + i]);
                if constexpr (EPI != EPI_STORE_STATS && EPI != EPI_DGELU) {   // conv outputs / dgrads carry no bias
                  if (has_bias && col_ok) {
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
                    const float2 s0 = fadd2(make_float2(x[0], x[1]), make_float2(b0.x, b0.y));
                    const float2 s1 = fadd2(make_float2(x[2], x[3]), make_float2(b0.z, b0.w));
                    const float2 s2 = fadd2(make_float2(x[4], x[5]), make_float2(b1.x, b1.y));
                    const float2 s3 = fadd2(make_float2(x[6], x[7]), make_float2(b1.z, b1.w));
                    x[0] = s0.x; x[1] = s0.y; x[2] = s1.x; x[3] = s1.y;
                    x[4] = s2.x; x[5] = s2.y; x[6] = s3.x; x[7] = s3.y;
                  }
                }
                if constexpr (RES_EPI) {
                  if constexpr (EPI == EPI_STORE_DROP) {
                    const uint32_t keep =
                        dropout_keep8(p.seed, static_cast<uint64_t>(row * p.N + col) >> 3, p.drop_threshold);
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = (keep >> i) & 1 ? x[i] * p.inv_keep : 0.f;
                  }
                  if (res_row != nullptr && col_ok) {
                    // (row-strided per lane: relies on L1 to serve the other 16-byte pieces of each 128-byte line)
                    uint4 r = rres[g];
                    if constexpr (EPI == EPI_STORE_MASKRES) {
                      // keep the residual elements whose mask bit is set: bits 2j, 2j+1 select the halves of word j
                      const uint32_t m8 = rmask >> (8 * g);
                      r.x &= ((m8 & 1u) ? 0x0000ffffu : 0u) | ((m8 & 2u) ? 0xffff0000u : 0u);
                      r.y &= ((m8 & 4u) ? 0x0000ffffu : 0u) | ((m8 & 8u) ? 0xffff0000u : 0u);
                      r.z &= ((m8 & 16u) ? 0x0000ffffu : 0u) | ((m8 & 32u) ? 0xffff0000u : 0u);
                      r.w &= ((m8 & 64u) ? 0x0000ffffu : 0u) | ((m8 & 128u) ? 0xffff0000u : 0u);
                    }
                    const float2 s0 = fadd2(make_float2(x[0], x[1]), unpack_bf16x2(r.x));
                    const float2 s1 = fadd2(make_float2(x[2], x[3]), unpack_bf16x2(r.y));
                    const float2 s2 = fadd2(make_float2(x[4], x[5]), unpack_bf16x2(r.z));
                    const float2 s3 = fadd2(make_float2(x[6], x[7]), unpack_bf16x2(r.w));
                    x[0] = s0.x; x[1] = s0.y; x[2] = s1.x; x[3] = s1.y;
                    x[4] = s2.x; x[5] = s2.y; x[6] = s3.x; x[7] = s3.y;
                  }
                  if constexpr (EPI == EPI_RELU) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = fmaxf(x[i], 0.f);
                  }
                }
                if constexpr (EPI == EPI_DGELU) {
                  if (col_ok) {
                    const int scc = c * 4 + g;
                    const uint4 r = *reinterpret_cast<const uint4*>(
                        stage_buf + (BOX_W == 64 ? lane * 128 + ((scc ^ (lane & 7)) << 4)
                                                 : lane * 64 + ((scc ^ ((lane >> 1) & 3)) << 4)));
                    const float2 g0 = fmul2(make_float2(x[0], x[1]), gelu_erf_grad2(unpack_bf16x2(r.x)));
                    const float2 g1 = fmul2(make_float2(x[2], x[3]), gelu_erf_grad2(unpack_bf16x2(r.y)));
                    const float2 g2 = fmul2(make_float2(x[4], x[5]), gelu_erf_grad2(unpack_bf16x2(r.z)));
                    const float2 g3 = fmul2(make_float2(x[6], x[7]), gelu_erf_grad2(unpack_bf16x2(r.w)));
                    x[0] = g0.x; x[1] = g0.y; x[2] = g1.x; x[3] = g1.y;
                    x[4] = g2.x; x[5] = g2.y; x[6] = g3.x; x[7] = g3.y;
                  }
                }
                uint4 o;
                o.x = pack_bf16x2(x[0], x[1]); o.y = pack_bf16x2(x[2], x[3]);
                o.z = pack_bf16x2(x[4], x[5]); o.w = pack_bf16x2(x[6], x[7]);
                const int cc = c * 4 + g;  // 16-byte chunk index inside the staged row
                const uint32_t off = BOX_W == 64 ? lane * 128 + ((cc ^ (lane & 7)) << 4)
                                                 : lane * 64 + ((cc ^ ((lane >> 1) & 3)) << 4);
                *reinterpret_cast<uint4*>(stage_buf + off) = o;
                if constexpr (EPI == EPI_GELU) {
                  // activation of the STORED (bf16-rounded) pre-activation: forward and backward see the same z
                  const float2 a0 = gelu_erf2(unpack_bf16x2(o.x)), a1 = gelu_erf2(unpack_bf16x2(o.y)),
                               a2 = gelu_erf2(unpack_bf16x2(o.z)), a3 = gelu_erf2(unpack_bf16x2(o.w));
                  uint4 a;
                  a.x = pack_bf16x2(a0.x, a0.y); a.y = pack_bf16x2(a1.x, a1.y);
                  a.z = pack_bf16x2(a2.x, a2.y); a.w = pack_bf16x2(a3.x, a3.y);
                  *reinterpret_cast<uint4*>(stage_buf + GEMM_BOX_BYTES + off) = a;
                }
              }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if constexpr (EPI == EPI_STORE) {
                // in-place residual (out += acc): the L2 performs the add, the SM never loads the residual
                if (p.accumulate) tma_reduce_add_2d(&tma_out, stage_buf, box_col0, m_blk * GEMM_BM + quarter * 32);
                else tma_store_2d(&tma_out, stage_buf, box_col0, m_blk * GEMM_BM + quarter * 32);
              } else {
                tma_store_2d(&tma_out, stage_buf, box_col0, m_blk * GEMM_BM + quarter * 32);
                if constexpr (EPI == EPI_GELU)
                  tma_store_2d(&tma_out2, stage_buf + GEMM_BOX_BYTES, box_col0, m_blk * GEMM_BM + quarter * 32);
              }
              tma_store_commit();
            }
            if constexpr (COLSUM_EPI) if (do_colsum) {
              // column sums over the 32 staged rows: lane l owns the 4-byte word (2 columns) l of every row --
              // conflict-free under either swizzle; rows past M were zero-filled by TMA and add nothing
              float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
              if constexpr (BOX_W == 64) {
                const uint32_t cc = lane >> 2, wo = (lane & 3) * 4;
                // two independent packed (FADD2 / FFMA2) chains over the even and the odd rows
                float2 sa = make_float2(0.f, 0.f), sb = sa, qa = sa, qb = sa;
#pragma unroll
                for (int r = 0; r < 32; r += 2) {
                  const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(
                      stage_buf + r * 128 + ((cc ^ (r & 7)) << 4) + wo));
                  const float2 g = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(
                      stage_buf + (r + 1) * 128 + ((cc ^ ((r + 1) & 7)) << 4) + wo));
                  sa = fadd2(sa, f); qa = ffma2(f, f, qa);
                  sb = fadd2(sb, g); qb = ffma2(g, g, qb);
                }
                sa = fadd2(sa, sb); qa = fadd2(qa, qb);
                s0 = sa.x; s1 = sa.y; q0 = qa.x; q1 = qa.y;
              } else {
                const uint32_t w16 = lane & 15, cc = w16 >> 2, wo = (w16 & 3) * 4, rb = (lane >> 4) * 16;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const uint32_t r = rb + i;
                  const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(
                      stage_buf + r * 64 + ((cc ^ ((r >> 1) & 3)) << 4) + wo));
                  s0 += f.x; s1 += f.y;
                  q0 = fmaf(f.x, f.x, q0); q1 = fmaf(f.y, f.y, q1);
                }
                s0 += __shfl_xor_sync(0xffffffffu, s0, 16); s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
                q0 += __shfl_xor_sync(0xffffffffu, q0, 16); q1 += __shfl_xor_sync(0xffffffffu, q1, 16);
              }
#pragma unroll
              for (int bb = 0; bb < BOXES; ++bb) {
                if (bb == b) { st_s[bb][0] += s0; st_s[bb][1] += s1; st_q[bb][0] += q0; st_q[bb][1] += q1; }
              }
            }
            if (++buf == p.nbuf) buf = 0;
          }
        }
      } else {
        constexpr int CHUNK = HALF_COLS < 32 ? HALF_COLS : 32;
#pragma unroll 1
        for (int c = 0; c < HALF_COLS / CHUNK; ++c) {
          const int col_in_tile = half * HALF_COLS + c * CHUNK;
          uint32_t v[32];
          tmem_ld32(t_base + col_in_tile, v);
          tmem_ld_wait();
          const int col0 = n_blk * BN + col_in_tile;
          if (row_ok && col0 < p.N) {
#pragma unroll
            for (int g = 0; g < CHUNK / 8; ++g) {
              const int col = col0 + g * 8;
              if (col >= p.N) break;
              float x[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(v[g * 8 + i]);
              float* o = static_cast<float*>(p.out) + row * p.ldc + col;
              if constexpr (EPI == EPI_F32) {
                if (has_bias) {
                  const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                  const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
                  x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
                  x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
                }
                *reinterpret_cast<float4*>(o) = make_float4(x[0], x[1], x[2], x[3]);
                *reinterpret_cast<float4*>(o + 4) = make_float4(x[4], x[5], x[6], x[7]);
              } else {  // EPI_F32_ATOMIC
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o), "f"(x[0]), "f"(x[1]), "f"(x[2]),
                             "f"(x[3])
                             : "memory");
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o + 4), "f"(x[4]), "f"(x[5]),
                             "f"(x[6]), "f"(x[7])
                             : "memory");
              }
            }
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_leader(&tmem_empty[acc]); else mbar_arrive(&tmem_empty[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      advance(m_unit, n_blk);
      while (m_unit >= m_units) m_unit -= m_units;   // next K split of the same tile grid
    }
    if constexpr (COLSUM_EPI) {
      if (do_colsum) stats_flush();
    }
    if constexpr (BF16_OUT) {
      // shared memory must outlive the last bulk store's READ of it; the global writes themselves complete with the grid
      // (waiting for them here cost ~1 us of every launch's tail)
      if (lane == 0) tma_store_wait_read<0>();
    }
  }

  tc_fence_before_sync();
  if constexpr (PAIR) cluster_sync_all();   // no CTA may retire while its peer can still signal / read it
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    if constexpr (PAIR) tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
    else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}


// Launch one (BN, EPI, CTAS) instantiation.  Shared memory split: `stages` ring slots + `nbuf` staging boxes per
// epilogue warp.  CTAS == 2 launches clusters of two CTAs (cudaLaunchKernelEx, cluster dimension 2).
// ring depth of the B-resident mode: what is left next to the resident [BN x K] tile and the staging boxes
static inline int b_resident_stages(int bn, int ew, int nbuf, int k_iters, int budget) {
  const int avail = budget - ew * nbuf * GEMM_BOX_BYTES - 512 - k_iters * bn * GEMM_BK * 2;
  const int st = avail / (GEMM_BM * GEMM_BK * 2);
  return st > GEMM_MAX_STAGES ? GEMM_MAX_STAGES : st;
}

template <int BN, int EPI, int CTAS = 1, int EW = GEMM_EPI_WARPS>
static int launch_gemm_inst(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& to2,
                            GemmParams p, int grid, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CTAS, EW>;
  const int grid_in = grid;
  if (p.b_resident && CTAS == 1) grid -= grid % p.n_tiles;   // every tile of a CTA in one column block
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<BN, EPI, CTAS, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         GEMM_SMEM_TOTAL + 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  // long-K tiles live in the main loop: deepest ring, one staging box; short-K tiles are store-bound: two boxes
  // (16 epilogue warps stage ONE box per tile each: a single buffer, free again long before the next tile)
  const int k_per_tile = p.k_iters_per_split;
  if (EW == 16) p.nbuf = 1;
  else if (p.nbuf <= 0 || EPI == EPI_GELU || EPI == EPI_F32 || EPI == EPI_F32_ATOMIC)
    p.nbuf = (EPI == EPI_F32 || EPI == EPI_F32_ATOMIC) ? 0 : (EPI == EPI_GELU || k_per_tile < 8 ? 2 : 1);   // fp32 modes do not stage
  const int budget = GEMM_SMEM_TOTAL - (g_tune[3] > 0 && g_tune[3] <= 64 ? g_tune[3] * 1024 : 0);   // see common.cuh, knob 3
  p.num_stages = Cfg::stages_for(p.nbuf, budget);
  if (p.num_stages < 2) return B200MM_ERR_BAD_ARG;
  static_assert(Cfg::stages_for(2) >= 2, "ring too shallow");
  if (p.b_resident) {
    if (CTAS == 1 && EPI != EPI_GELU) p.nbuf = p.nbuf > 1 ? 1 : p.nbuf;
    const int st = b_resident_stages(BN, EW, p.nbuf, p.k_iters, budget);
    if (CTAS == 1 && st >= 3) p.num_stages = st;
    else { p.b_resident = 0; grid = grid_in; }
  }
  if constexpr (CTAS == 1) {
    cudaError_t e = launch_pdl(gemm_bf16_kernel<BN, EPI, 1, EW>, dim3(grid), dim3(Cfg::THREADS), budget + 1024,
                               stream, ta, tb, to, to2, p);
    if (e != cudaSuccess) return static_cast<int>(e);
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(grid & ~1), 1, 1);
    cfg.blockDim = dim3(Cfg::THREADS, 1, 1);
    cfg.dynamicSmemBytes = budget + 1024;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<BN, EPI, 2, EW>, ta, tb, to, to2, p);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? B200MM_OK : static_cast<int>(e);
}

// CTA-pair instantiations (BN = 256 only; matrix operands, the text / ViT tower shapes)
template <int BN>
int launch_gemm_pair(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& to2,
                     const GemmParams& p, int grid, cudaStream_t stream) {
  switch (p.epi) {
    case EPI_STORE:
      if (p.col_stats != nullptr) return launch_gemm_inst<BN, EPI_STORE_STATS, 2>(ta, tb, to, to2, p, grid, stream);
      if (p.p_drop > 0.f) return launch_gemm_inst<BN, EPI_STORE_DROP, 2>(ta, tb, to, to2, p, grid, stream);
      return launch_gemm_inst<BN, EPI_STORE, 2>(ta, tb, to, to2, p, grid, stream);
    case EPI_GELU: return launch_gemm_inst<BN, EPI_GELU, 2>(ta, tb, to, to2, p, grid, stream);
    case EPI_DGELU: return launch_gemm_inst<BN, EPI_DGELU, 2>(ta, tb, to, to2, p, grid, stream);
    case EPI_F32_ATOMIC: return launch_gemm_inst<BN, EPI_F32_ATOMIC, 2>(ta, tb, to, to2, p, grid, stream);
    default: return B200MM_ERR_BAD_ARG;
  }
}

template <int BN>
int launch_gemm_bn(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& to2,
                   const GemmParams& p, int grid, cudaStream_t stream) {
  switch (p.epi) {
    case EPI_STORE:
      if (p.res_mask != nullptr) return launch_gemm_inst<BN, EPI_STORE_MASKRES>(ta, tb, to, to2, p, grid, stream);
      if (p.col_stats != nullptr) return launch_gemm_inst<BN, EPI_STORE_STATS>(ta, tb, to, to2, p, grid, stream);
      if (p.p_drop > 0.f) return launch_gemm_inst<BN, EPI_STORE_DROP>(ta, tb, to, to2, p, grid, stream);
      return launch_gemm_inst<BN, EPI_STORE>(ta, tb, to, to2, p, grid, stream);
    case EPI_GELU: return launch_gemm_inst<BN, EPI_GELU>(ta, tb, to, to2, p, grid, stream);
    case EPI_DGELU: return launch_gemm_inst<BN, EPI_DGELU>(ta, tb, to, to2, p, grid, stream);
    case EPI_F32: return launch_gemm_inst<BN, EPI_F32>(ta, tb, to, to2, p, grid, stream);
    case EPI_F32_ATOMIC: return launch_gemm_inst<BN, EPI_F32_ATOMIC>(ta, tb, to, to2, p, grid, stream);
    default: return launch_gemm_inst<BN, EPI_RELU>(ta, tb, to, to2, p, grid, stream);
  }
}

}  // namespace b200
