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
#include "gemm_params.cuh"
#include "device_utils.cuh"
#include <cstdlib>

namespace b200 {
// one translation unit per tile width (gemm_bn64.cu / gemm_bn128.cu / gemm_bn256.cu) so they compile in parallel
int launch_gemm_bn64(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const GemmParams&,
                     int, cudaStream_t);
int launch_gemm_bn128(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,
                      const GemmParams&, int, cudaStream_t);
int launch_gemm_bn256(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,
                      const GemmParams&, int, cudaStream_t);
int launch_gemm_bn256_pair(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,
                           const GemmParams&, int, cudaStream_t);
}  // namespace b200


using namespace b200;

namespace {

// Where an operand comes from: a matrix (tiled TMA loads, K-major or MN-major) or an NHWC activation gathered with
// im2col-mode TMA loads (implicit-GEMM convolution).
struct Operand {
  const void* ptr = nullptr;
  int mn = 0;           // matrix: 0 = K-major ([rows, K]), 1 = MN-major ([K, rows])
  long long ld = 0;     // matrix: row stride in elements
  int im2col = 0;       // 1: ptr is an NHWC activation, geometry in ConvGeom
};
struct ConvGeom {
  int N = 0, H = 0, W = 0, C = 0, ksize = 0, stride = 1, pad = 0, P = 0, Q = 0;
};

int run_gemm(const Operand& A, const Operand& B, const ConvGeom& cg, int M, int N, int K, int epi, const float* bias,
             const void* residual, long long ldr, const void* aux, long long ld_aux, void* out, long long ldc,
             void* out2, long long ld2, int splits, int block_n, float p_drop, unsigned long long seed,
             float* col_stats, void* stream, const unsigned char* res_mask = nullptr, long long ld_mask = 0) {
  const DeviceInfo& dev = device_info();
  if (!dev.ok) return B200MM_ERR_NOT_SM100;
  if (dev.cc_major != 10) return B200MM_ERR_NOT_SM100;
  if (M <= 0 || N <= 0 || K <= 0 || (N & 7) || (ldc & 3)) return B200MM_ERR_BAD_ARG;
  if ((!A.im2col && (A.ld & 7)) || (!B.im2col && (B.ld & 7))) return B200MM_ERR_BAD_ARG;
  if (epi != EPI_F32 && epi != EPI_F32_ATOMIC && ((ldc & 7) || (epi == EPI_GELU && (ld2 & 7)))) return B200MM_ERR_BAD_ARG;
  if (epi < EPI_STORE || epi > EPI_RELU) return B200MM_ERR_BAD_ARG;
  if (splits < 1) splits = 1;
  if (splits > 1 && epi != EPI_F32_ATOMIC) return B200MM_ERR_BAD_ARG;
  if (epi == EPI_GELU && out2 == nullptr) return B200MM_ERR_BAD_ARG;
  if (epi == EPI_DGELU && (aux == nullptr || bias != nullptr)) return B200MM_ERR_BAD_ARG;
  // column statistics ride on the plain store epilogue (convolution outputs: no bias, residual or dropout); with the
  // dGELU epilogue col_stats is fp32 [N] and receives the column sums only (bias gradient of the first FFN layer)
  if (col_stats != nullptr && epi != EPI_DGELU &&
      (epi != EPI_STORE || bias != nullptr || residual != nullptr || p_drop > 0.f))
    return B200MM_ERR_BAD_ARG;
  if ((A.im2col || B.im2col) && (cg.C % 64 != 0 || cg.ksize < 1)) return B200MM_ERR_BAD_ARG;
  // masked residual (identity-branch gradient): plain store mode with a residual, 32 columns per mask word
  if (res_mask != nullptr && (epi != EPI_STORE || residual == nullptr || residual == out || bias != nullptr || p_drop > 0.f ||
                              col_stats != nullptr || (N & 31) || (ld_mask & 3) || (reinterpret_cast<uintptr_t>(res_mask) & 3)))
    return B200MM_ERR_BAD_ARG;

  int bn = block_n;
  if (bn == 0) bn = (N % 256 == 0 || N > 512) ? 256 : (N > 64 ? 128 : 64);
  if (bn != 64 && bn != 128 && bn != 256) return B200MM_ERR_BAD_ARG;

  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.a_mn = A.im2col ? 0 : (A.mn ? 1 : 0);
  p.b_mn = B.im2col ? 1 : (B.mn ? 1 : 0);
  p.a_im2col = A.im2col;
  p.b_im2col = B.im2col;
  p.conv_C = cg.C; p.conv_KW = cg.ksize; p.conv_stride = cg.stride; p.conv_pad = cg.pad;
  p.conv_P = cg.P; p.conv_Q = cg.Q; p.conv_cblocks = cg.C / 64;
  p.m_tiles = ceil_div(M, GEMM_BM);
  p.n_tiles = ceil_div(N, bn);
  p.k_iters = ceil_div(K, GEMM_BK);
  if (splits > p.k_iters) splits = p.k_iters;
  p.k_iters_per_split = ceil_div(p.k_iters, splits);
  p.splits = ceil_div(p.k_iters, p.k_iters_per_split);
  p.epi = epi;
  p.bias = bias;
  p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.ldr = ldr;
  // residual aliasing the output (same pointer and stride), plain store mode: accumulate in place with TMA reduce-add
  p.accumulate = (epi == EPI_STORE && residual != nullptr && residual == out && ldr == ldc && p_drop == 0.f) ? 1 : 0;
  if (p.accumulate) p.residual = nullptr;
  else if (residual != nullptr && residual == out) return B200MM_ERR_BAD_ARG;   // aliasing is only defined for that case
  p.aux = static_cast<const __nv_bfloat16*>(aux);
  p.ld_aux = ld_aux;
  p.out = out;
  p.ldc = ldc;
  p.out2 = static_cast<__nv_bfloat16*>(out2);
  p.ld2 = ld2;
  p.col_stats = col_stats;
  p.res_mask = res_mask;
  p.ld_mask = ld_mask;
  if (p_drop < 0.f || p_drop >= 1.f) return B200MM_ERR_BAD_ARG;
  p.p_drop = (epi == EPI_STORE) ? p_drop : 0.f;
  p.drop_threshold = dropout_threshold(p_drop);
  p.inv_keep = 1.f / (1.f - p_drop);
  p.seed = seed;

  // CTA-pair (cta_group::2) kernel: 256 x 256 tile per cluster of two CTAs, for the big matrix-operand shapes of the
  // text / ViT towers (B200MM_GEMM_PAIR=0 switches it off)
  static const bool pair_enabled = [] {
    const char* e = std::getenv("B200MM_GEMM_PAIR");
    return e == nullptr || e[0] != '0';
  }();
  // (the statistics / accumulate epilogues are served too; K < 512 stays on the single-CTA kernel: the pair measured
  // within +-5 % of it on the short-K convolution shapes, profiles/conv_gemm_variants_r02.log)
  const bool pair = pair_enabled && bn == 256 && !A.im2col && !B.im2col && res_mask == nullptr &&
                    (epi == EPI_STORE || epi == EPI_GELU || epi == EPI_DGELU || epi == EPI_F32_ATOMIC) &&
                    p.m_tiles >= 2 && p.k_iters_per_split >= g_tune[0] && (dev.num_sms & 1) == 0;
  p.nbuf = 0;
  CUtensorMap ta, tb;
  int rc;
  if (A.im2col)   rc = make_tmap_im2col_bf16(&ta, A.ptr, cg.N, cg.H, cg.W, cg.C, cg.ksize, cg.stride, cg.pad, GEMM_BM);
  else if (!A.mn) rc = make_tmap_2d_bf16(&ta, A.ptr, K, M, A.ld * 2, GEMM_BK, GEMM_BM);
  else            rc = make_tmap_2d_bf16(&ta, A.ptr, M, K, A.ld * 2, 64, GEMM_BK);
  if (rc) return rc;
  if (B.im2col)   rc = make_tmap_im2col_bf16(&tb, B.ptr, cg.N, cg.H, cg.W, cg.C, cg.ksize, cg.stride, cg.pad, GEMM_BK);
  else if (!B.mn) rc = make_tmap_2d_bf16(&tb, B.ptr, K, N, B.ld * 2, GEMM_BK, pair ? bn / 2 : bn);
  else            rc = make_tmap_2d_bf16(&tb, B.ptr, N, K, B.ld * 2, 64, GEMM_BK);
  if (rc) return rc;

  // output maps for the TMA-store epilogue (bf16 modes): box = [32 rows x 64 cols] SW128, or x 32 cols SW64 for BN=64
  CUtensorMap to = tb, to2 = tb;
  if (epi != EPI_F32 && epi != EPI_F32_ATOMIC) {
    const uint32_t box_w = bn == 64 ? 32 : 64;
    rc = make_tmap_2d_bf16(&to, out, N, M, ldc * 2, box_w, 32, box_w * 2);
    if (rc) return rc;
    to2 = to;
    if (epi == EPI_GELU) {
      rc = make_tmap_2d_bf16(&to2, out2, N, M, ld2 * 2, box_w, 32, box_w * 2);
      if (rc) return rc;
    }
  }

  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (pair) {
    const int units = ((p.m_tiles + 1) / 2) * p.n_tiles * p.splits;
    const int pgrid = 2 * units < dev.num_sms ? 2 * units : dev.num_sms;
    return launch_gemm_bn256_pair(ta, tb, to, to2, p, pgrid, s);
  }
  const int num_work = p.m_tiles * p.n_tiles * p.splits;
  const int grid = num_work < dev.num_sms ? num_work : dev.num_sms;
  // B-resident mode: bf16 output, matrix B operand, unsplit K, several tiles per CTA, n_tiles small enough that a grid
  // of a multiple of n_tiles CTAs keeps (nearly) every SM busy; the launcher checks that the tile and a >= 3-slot
  // ring fit and falls back otherwise
  p.b_resident = (g_tune[1] && !B.im2col && p.splits == 1 && epi != EPI_F32 && epi != EPI_F32_ATOMIC &&
                  p.n_tiles <= 8 && num_work >= 2 * dev.num_sms &&
                  static_cast<long long>(p.k_iters) * bn * GEMM_BK * 2 <= 128 * 1024) ? 1 : 0;
  switch (bn) {
    case 64: return launch_gemm_bn64(ta, tb, to, to2, p, grid, s);
    case 128: return launch_gemm_bn128(ta, tb, to, to2, p, grid, s);
    default: return launch_gemm_bn256(ta, tb, to, to2, p, grid, s);
  }
}

}  // namespace

namespace {

// w_rot[ci][k-1-kh][k-1-kw][co] = w[co][kh][kw][ci]: the weight of the transposed (data-gradient) convolution
__global__ void __launch_bounds__(256)
rotate_conv_weight_kernel(const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ w_rot, int Cout, int Cin,
                          int taps) {
  const long long total = static_cast<long long>(Cout) * taps * Cin;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(i % Cout);           // fastest index of the OUTPUT (coalesced writes)
    long long t = i / Cout;
    const int tap = static_cast<int>(t % taps);
    const int ci = static_cast<int>(t / taps);
    w_rot[i] = w[(static_cast<long long>(co) * taps + (taps - 1 - tap)) * Cin + ci];
  }
}

// all rotations of a step in one launch: blockIdx.y = table entry {w, w_rot, Cout << 32 | Cin, taps}
__global__ void __launch_bounds__(256)
rotate_conv_weight_multi_kernel(const long long* __restrict__ table) {
  const long long* e = table + 4 * blockIdx.y;
  const __nv_bfloat16* w = reinterpret_cast<const __nv_bfloat16*>(e[0]);
  __nv_bfloat16* w_rot = reinterpret_cast<__nv_bfloat16*>(e[1]);
  const int Cout = static_cast<int>(e[2] >> 32), Cin = static_cast<int>(e[2] & 0xffffffffLL), taps = static_cast<int>(e[3]);
  const long long total = static_cast<long long>(Cout) * taps * Cin;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(i % Cout);
    long long t = i / Cout;
    const int tap = static_cast<int>(t % taps);
    const int ci = static_cast<int>(t / taps);
    w_rot[i] = w[(static_cast<long long>(co) * taps + (taps - 1 - tap)) * Cin + ci];
  }
}

}  // namespace

// D[M,N] = A x B (+ epilogue), bf16 operands, fp32 accumulate.
//   a_mn == 0: A is stored [M, K] (row stride lda elements);  a_mn == 1: A is stored [K, M].
//   b_mn == 0: B is stored [N, K] (row stride ldb elements);  b_mn == 1: B is stored [K, N].
//   epi: EpiMode above.  splits > 1 requires epi == EPI_F32_ATOMIC (out must be pre-zeroed or hold the
//   value to accumulate onto).  block_n in {0 (auto), 64, 128, 256}.  p_drop/seed: EPI_STORE dropout (see GemmParams).
//   col_stats (nullable, epi 0 without bias/residual/dropout): fp32 [2N], col_stats[n] += sum_m out[m,n] and
//   col_stats[N + n] += sum_m out[m,n]^2 over the stored bf16 values -- the BatchNorm statistics of a conv output.
// Contract: pointers 16-byte aligned, lda/ldb/ldc/... multiples of 8 elements, N % 8 == 0.
B200MM_API int b200mm_gemm_bf16(const void* A, int a_mn, long long lda, const void* B, int b_mn, long long ldb,
                                int M, int N, int K, int epi, const float* bias, const void* residual,
                                long long ldr, const void* aux, long long ld_aux, void* out, long long ldc,
                                void* out2, long long ld2, int splits, int block_n, float p_drop,
                                unsigned long long seed, float* col_stats, void* stream) {
  Operand a, b;
  a.ptr = A; a.mn = a_mn; a.ld = lda;
  b.ptr = B; b.mn = b_mn; b.ld = ldb;
  return run_gemm(a, b, ConvGeom{}, M, N, K, epi, bias, residual, ldr, aux, ld_aux, out, ldc, out2, ld2, splits,
                  block_n, p_drop, seed, col_stats, stream);
}

// out[M,N] = bf16(A x B + (mask bit ? residual : 0)): the data gradient of a residual block's first convolution joined
// with the identity branch's gradient dz = dout o relu_mask WITHOUT materialising dz -- `residual` is the gradient
// w.r.t. the block's output, `mask` the 1-bit-per-element ReLU mask b200mm_batchnorm_fwd_stats wrote ([M, N / 8]
// bytes).  Saves one full write (and the read-modify-write of an in-place accumulate) per identity block, but the
// per-lane residual loads make the epilogue slower than the TMA reduce-add accumulate it replaces (see image_tower.py:
// off by default).  N % 32 == 0.
B200MM_API int b200mm_gemm_bf16_maskres(const void* A, int a_mn, long long lda, const void* B, int b_mn, long long ldb,
                                        int M, int N, int K, const void* residual, long long ldr,
                                        const unsigned char* mask, long long ld_mask, void* out, long long ldc,
                                        void* stream) {
  if (mask == nullptr) return B200MM_ERR_BAD_ARG;
  Operand a, b;
  a.ptr = A; a.mn = a_mn; a.ld = lda;
  b.ptr = B; b.mn = b_mn; b.ld = ldb;
  return run_gemm(a, b, ConvGeom{}, M, N, K, EPI_STORE, nullptr, residual, ldr, nullptr, 0, out, ldc, nullptr, 0, 1, 0,
                  0.f, 0, nullptr, stream, mask, ld_mask);
}

namespace b200 {
// conv3x3_c64.cu: halo-resident 3x3 / stride 1 / 64 -> 64 channel kernels (B200MM_HALO_CONV=0 falls back to im2col loads)
bool conv3x3_c64_supported(int N, int H, int W);
int conv3x3_c64_fwd(const void* x, int N, int H, int W, const void* w, void* out, float* col_stats, cudaStream_t stream);
int conv3x3_c64_wgrad(const void* dy, const void* x, int N, int H, int W, float* dw, cudaStream_t stream);
}  // namespace b200
static bool halo_conv_enabled() {
  static const bool on = [] {
    const char* e = getenv("B200MM_HALO_CONV");
    return !(e && e[0] == '0');
  }();
  return on;
}

// Implicit-GEMM convolution forward (square k x k window, symmetric padding): out[N*P*Q, Cout] = conv(x, w) with
// x an NHWC bf16 activation [N,H,W,C] (C % 64 == 0) gathered by TMA im2col loads -- no im2col matrix in memory --
// and w the OHWI-flattened weight [Cout, k*k*C].  epi / bias / residual as in b200mm_gemm_bf16 (bf16 output modes).
// The stride-1 data gradient is the same call on dY with the rotated weight (b200mm_conv_weight_rotate).
// Replaces torchvision conv3x3 (torchvision/models/resnet.py:19-31, :118-130) forward / dgrad.
B200MM_API int b200mm_conv_fwd(const void* x, int N, int H, int W, int C, const void* w, int Cout, int ksize,
                               int stride, int pad, int epi, const float* bias, const void* residual, long long ldr,
                               void* out, long long ldc, float* col_stats, void* stream) {
  if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || Cout <= 0 || ksize <= 0 || stride <= 0 || pad < 0)
    return B200MM_ERR_BAD_ARG;
  if (halo_conv_enabled() && ksize == 3 && stride == 1 && pad == 1 && C == 64 && Cout == 64 && epi == EPI_STORE &&
      bias == nullptr && residual == nullptr && ldc == 64 && conv3x3_c64_supported(N, H, W))
    return conv3x3_c64_fwd(x, N, H, W, w, out, col_stats, static_cast<cudaStream_t>(stream));
  ConvGeom cg;
  cg.N = N; cg.H = H; cg.W = W; cg.C = C; cg.ksize = ksize; cg.stride = stride; cg.pad = pad;
  cg.P = (H + 2 * pad - ksize) / stride + 1;
  cg.Q = (W + 2 * pad - ksize) / stride + 1;
  Operand a, b;
  a.ptr = x; a.im2col = 1;
  b.ptr = w; b.mn = 0; b.ld = static_cast<long long>(ksize) * ksize * C;
  const long long M = static_cast<long long>(N) * cg.P * cg.Q;
  if (M > 0x7fffffffLL) return B200MM_ERR_BAD_ARG;
  return run_gemm(a, b, cg, static_cast<int>(M), Cout, ksize * ksize * C, epi, bias, residual, ldr, nullptr, 0, out,
                  ldc, nullptr, 0, 1, 0, 0.f, 0, col_stats, stream);
}

// Implicit-GEMM weight gradient: dw[Cout, k*k*C] (fp32) += dy[N*P*Q, Cout]^T x im2col(x); the im2col operand is
// gathered by TMA, the reduction over output pixels is split across CTAs (fp32 red.global.add).
B200MM_API int b200mm_conv_wgrad(const void* dy, long long ld_dy, const void* x, int N, int H, int W, int C, int Cout,
                                 int ksize, int stride, int pad, float* dw, int splits, void* stream) {
  if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || Cout <= 0 || ksize <= 0 || stride <= 0 || pad < 0)
    return B200MM_ERR_BAD_ARG;
  if (halo_conv_enabled() && ksize == 3 && stride == 1 && pad == 1 && C == 64 && Cout == 64 && ld_dy == 64 &&
      conv3x3_c64_supported(N, H, W))
    return conv3x3_c64_wgrad(dy, x, N, H, W, dw, static_cast<cudaStream_t>(stream));
  ConvGeom cg;
  cg.N = N; cg.H = H; cg.W = W; cg.C = C; cg.ksize = ksize; cg.stride = stride; cg.pad = pad;
  cg.P = (H + 2 * pad - ksize) / stride + 1;
  cg.Q = (W + 2 * pad - ksize) / stride + 1;
  const long long pixels = static_cast<long long>(N) * cg.P * cg.Q;
  if (pixels > 0x7fffffffLL) return B200MM_ERR_BAD_ARG;
  Operand a, b;
  a.ptr = dy; a.mn = 1; a.ld = ld_dy;     // stored [pixels, Cout] = [K, M]
  b.ptr = x; b.im2col = 1;
  const int ncols = ksize * ksize * C;
  return run_gemm(a, b, cg, Cout, ncols, static_cast<int>(pixels), EPI_F32_ATOMIC, nullptr, nullptr, 0, nullptr, 0,
                  dw, ncols, nullptr, 0, splits, 0, 0.f, 0, nullptr, stream);
}

// Every rotation a backward pass needs in ONE launch (ResNet-50: 13 stride-1 3x3 convolutions, 13 launches of ~7 us
// before).  table: DEVICE int64 [n][4] = {w pointer, w_rot pointer, Cout << 32 | Cin, ksize * ksize}.
B200MM_API int b200mm_conv_weight_rotate_multi(const long long* table, int n, void* stream) {
  if (table == nullptr || n <= 0 || n > 65535) return B200MM_ERR_BAD_ARG;
  rotate_conv_weight_multi_kernel<<<dim3(96, n), 256, 0, static_cast<cudaStream_t>(stream)>>>(table);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

// w [Cout, k, k, Cin] -> w_rot [Cin, k, k, Cout] with both spatial axes flipped (bf16)
B200MM_API int b200mm_conv_weight_rotate(const void* w, void* w_rot, int Cout, int Cin, int ksize, void* stream) {
  if (Cout <= 0 || Cin <= 0 || ksize <= 0) return B200MM_ERR_BAD_ARG;
  const long long total = static_cast<long long>(Cout) * ksize * ksize * Cin;
  const int grid = static_cast<int>(total / 256 > 1184 ? 1184 : ceil_div(total, 256LL));
  rotate_conv_weight_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(w), static_cast<__nv_bfloat16*>(w_rot), Cout, Cin, ksize * ksize);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
