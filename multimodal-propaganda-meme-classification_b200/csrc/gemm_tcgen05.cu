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

namespace b200 {
// one translation unit per tile width (gemm_bn64.cu / gemm_bn128.cu / gemm_bn256.cu) so they compile in parallel
int launch_gemm_bn64(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const GemmParams&,
                     int, cudaStream_t);
int launch_gemm_bn128(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,
                      const GemmParams&, int, cudaStream_t);
int launch_gemm_bn256(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,
                      const GemmParams&, int, cudaStream_t);
}  // namespace b200


using namespace b200;

// D[M,N] = A x B (+ epilogue), bf16 operands, fp32 accumulate.
//   a_mn == 0: A is stored [M, K] (row stride lda elements);  a_mn == 1: A is stored [K, M].
//   b_mn == 0: B is stored [N, K] (row stride ldb elements);  b_mn == 1: B is stored [K, N].
//   epi: EpiMode above.  splits > 1 requires epi == EPI_F32_ATOMIC (out must be pre-zeroed or hold the
//   value to accumulate onto).  block_n in {0 (auto), 64, 128, 256}.  p_drop/seed: EPI_STORE dropout (see GemmParams).
// Contract: pointers 16-byte aligned, lda/ldb/ldc/... multiples of 8 elements, N % 8 == 0.
B200MM_API int b200mm_gemm_bf16(const void* A, int a_mn, long long lda, const void* B, int b_mn, long long ldb,
                                int M, int N, int K, int epi, const float* bias, const void* residual,
                                long long ldr, const void* aux, long long ld_aux, void* out, long long ldc,
                                void* out2, long long ld2, int splits, int block_n, float p_drop,
                                unsigned long long seed, void* stream) {
  const DeviceInfo& dev = device_info();
  if (!dev.ok) return B200MM_ERR_NOT_SM100;
  if (dev.cc_major != 10) return B200MM_ERR_NOT_SM100;
  if (M <= 0 || N <= 0 || K <= 0 || (N & 7) || (lda & 7) || (ldb & 7) || (ldc & 3)) return B200MM_ERR_BAD_ARG;
  if (epi != EPI_F32 && epi != EPI_F32_ATOMIC && ((ldc & 7) || (epi == EPI_GELU && (ld2 & 7)))) return B200MM_ERR_BAD_ARG;
  if (epi < EPI_STORE || epi > EPI_RELU) return B200MM_ERR_BAD_ARG;
  if (splits < 1) splits = 1;
  if (splits > 1 && epi != EPI_F32_ATOMIC) return B200MM_ERR_BAD_ARG;
  if (epi == EPI_GELU && out2 == nullptr) return B200MM_ERR_BAD_ARG;
  if (epi == EPI_DGELU && aux == nullptr) return B200MM_ERR_BAD_ARG;

  int bn = block_n;
  if (bn == 0) bn = (N % 256 == 0 || N > 512) ? 256 : (N > 64 ? 128 : 64);
  if (bn != 64 && bn != 128 && bn != 256) return B200MM_ERR_BAD_ARG;

  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.a_mn = a_mn ? 1 : 0;
  p.b_mn = b_mn ? 1 : 0;
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
  p.aux = static_cast<const __nv_bfloat16*>(aux);
  p.ld_aux = ld_aux;
  p.out = out;
  p.ldc = ldc;
  p.out2 = static_cast<__nv_bfloat16*>(out2);
  p.ld2 = ld2;
  if (p_drop < 0.f || p_drop >= 1.f) return B200MM_ERR_BAD_ARG;
  p.p_drop = (epi == EPI_STORE) ? p_drop : 0.f;
  p.drop_threshold = dropout_threshold(p_drop);
  p.inv_keep = 1.f / (1.f - p_drop);
  p.seed = seed;

  CUtensorMap ta, tb;
  int rc;
  if (!p.a_mn) rc = make_tmap_2d_bf16(&ta, A, K, M, lda * 2, GEMM_BK, GEMM_BM);
  else         rc = make_tmap_2d_bf16(&ta, A, M, K, lda * 2, 64, GEMM_BK);
  if (rc) return rc;
  if (!p.b_mn) rc = make_tmap_2d_bf16(&tb, B, K, N, ldb * 2, GEMM_BK, bn);
  else         rc = make_tmap_2d_bf16(&tb, B, N, K, ldb * 2, 64, GEMM_BK);
  if (rc) return rc;

  // output maps for the TMA-store epilogue (bf16 modes): box = [32 rows x 64 cols] SW128, or x 32 cols SW64 for BN=64
  CUtensorMap to = ta, to2 = ta;
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

  const int num_work = p.m_tiles * p.n_tiles * p.splits;
  const int grid = num_work < dev.num_sms ? num_work : dev.num_sms;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (bn) {
    case 64: return launch_gemm_bn64(ta, tb, to, to2, p, grid, s);
    case 128: return launch_gemm_bn128(ta, tb, to, to2, p, grid, s);
    default: return launch_gemm_bn256(ta, tb, to, to2, p, grid, s);
  }
}
