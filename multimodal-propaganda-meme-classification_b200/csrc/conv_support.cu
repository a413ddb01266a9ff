// Image-tower kernels that surround the implicit-GEMM convolutions (all NHWC bf16, HBM-bound, 16-byte
// accesses over the channel dim): train-mode BatchNorm (+ReLU, +residual) forward/backward, max / average
// pooling, and the im2col / col2im lowering that turns 3x3 and 7x7 convolutions into the tcgen05 GEMM.
//
// Replaces: torchvision/models/resnet.py:197-206 (stem), :108-163 (Bottleneck),
// :266-280 (_forward_impl) -- cuDNN convolution / batch-norm and ATen pooling -- and their backward passes.
#include "common.cuh"
#include "device_utils.cuh"
#include <cstdlib>

namespace b200 {

static int grid_for(long long work_items, int threads) {
  const DeviceInfo& dev = device_info();
  const long long max_ctas = static_cast<long long>(dev.num_sms > 0 ? dev.num_sms : 148) * 8;
  const long long want = ceil_div(work_items, static_cast<long long>(threads));
  return static_cast<int>(want < 1 ? 1 : (want > max_ctas ? max_ctas : want));
}

// Column-reduction epilogue shared by the BatchNorm reductions.  Thread (ty, g) holds partial sums a[8], b[8] of
// channels g*8..g*8+7.  (1) CTA-level reduction over ty in shared memory; (2) one fp32 atomic per channel and CTA
// into one of BN_REPLICAS replica rows (CTA index mod R) -- spreading the same-address traffic that otherwise
// saturates a handful of L2 atomic units (measured: 45 % of the kernel with a single row, 1184 CTAs);
// (3) the last CTA to finish (ticket counter) sums the replicas into the final row.
// scratch layout (floats): [R][2C] replicas | [2C] final | 1 ticket counter (all zeroed by the caller).
constexpr int BN_REPLICAS = 8;
__device__ __forceinline__ void column_reduce_finish(float (&red)[2][256][8], float (&a)[8], float (&b)[8], int G,
                                                     int g, int ty, int rpp, int C, float* __restrict__ scratch) {
  __shared__ unsigned int s_ticket;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[0][threadIdx.x][i] = a[i];
    red[1][threadIdx.x][i] = b[i];
  }
  __syncthreads();
  if (ty == 0) {
    for (int t = 1; t < rpp; ++t)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a[i] += red[0][t * G + g][i];
        b[i] += red[1][t * G + g][i];
      }
    float* rep = scratch + static_cast<size_t>(blockIdx.x % BN_REPLICAS) * 2 * C;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(rep + g * 8 + i, a[i]);
      atomicAdd(rep + C + g * 8 + i, b[i]);
    }
  }
  __threadfence();
  __syncthreads();
  unsigned int* counter = reinterpret_cast<unsigned int*>(scratch + static_cast<size_t>(BN_REPLICAS + 1) * 2 * C);
  if (threadIdx.x == 0) s_ticket = atomicAdd(counter, 1u);
  __syncthreads();
  if (s_ticket == gridDim.x - 1) {
    __threadfence();
    float* fin = scratch + static_cast<size_t>(BN_REPLICAS) * 2 * C;
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < BN_REPLICAS; ++r) t += __ldcg(scratch + static_cast<size_t>(r) * 2 * C + c);
      fin[c] = t;
    }
  }
}

// ------------------------------------------------------------------ BatchNorm2d (training mode)
// Layout: x [M = N*H*W, C]; G = C/8 channel groups (power of two, <= 256); a CTA of 256 threads covers
// 256/G rows per pass, thread (ty, g) owns channels g*8..g*8+7.
__global__ void __launch_bounds__(256)
bn_stats_kernel(const __nv_bfloat16* __restrict__ x, long long M, int C, int rows_per_cta,
                float* __restrict__ scratch) {
  __shared__ float red[2][256][8];
  const int G = C >> 3;
  const int g = threadIdx.x % G, ty = threadIdx.x / G, rpp = 256 / G;
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long r1 = min(r0 + rows_per_cta, M);
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  constexpr int U = 8;  // independent 16-byte loads in flight per thread (the loop is latency-bound otherwise)
  for (long long r = r0 + ty; r < r1; r += static_cast<long long>(rpp) * U) {
    uint4 raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + static_cast<long long>(u) * rpp;
      raw[u] = rr < r1 ? __ldg(reinterpret_cast<const uint4*>(x + rr * C + g * 8)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float v[8];
      unpack8(raw[u], v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += v[i];
        q[i] = fmaf(v[i], v[i], q[i]);
      }
    }
  }
  column_reduce_finish(red, s, q, G, g, ty, rpp, C, scratch);
}

// out = act(x * scale + shift (+ residual)); CTA 0 also records mean / rstd and updates the running statistics
// (momentum update with the unbiased variance, as nn.BatchNorm2d does).
template <bool HAS_RES>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ residual, long long M, int C,
                int rows_per_cta, const float* __restrict__ sum, const float* __restrict__ sumsq,
                const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum, int relu,
                __nv_bfloat16* __restrict__ out, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                float* __restrict__ running_mean, float* __restrict__ running_var,
                unsigned char* __restrict__ relu_mask) {
  pdl_launch_dependents();   // see launch_pdl (common.cuh)
  pdl_wait();
  const int G = C >> 3;
  const int g = threadIdx.x % G, ty = threadIdx.x / G, rpp = 256 / G;
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = g * 8 + i;
    const float mean = sum[c] / M;
    const float var = fmaxf(sumsq[c] / M - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    sc[i] = gamma[c] * rstd;
    sh[i] = fmaf(-mean, sc[i], beta[c]);   // explicit fma: the backward recomputes the ReLU mask with the same ops
    if (blockIdx.x == 0 && ty == 0) {
      mean_out[c] = mean;
      rstd_out[c] = rstd;
      if (running_mean != nullptr) {
        const float unbiased = M > 1 ? var * (static_cast<float>(M) / static_cast<float>(M - 1)) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
      }
    }
  }
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long r1 = min(r0 + rows_per_cta, M);
  // independent 16-byte loads in flight per thread: 8 on the single input stream, 4 + 4 with a residual (what is
  // outstanding per SM sets the bandwidth of this pass: 65 KB gave 4.4 TB/s)
  constexpr int U = HAS_RES ? 4 : 8;
  for (long long r = r0 + ty; r < r1; r += static_cast<long long>(rpp) * U) {
    uint4 rx[U], rr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = r + static_cast<long long>(u) * rpp;
      const bool ok = row < r1;
      rx[u] = ok ? __ldg(reinterpret_cast<const uint4*>(x + row * C + g * 8)) : make_uint4(0, 0, 0, 0);
      rr[u] = (HAS_RES && ok) ? __ldg(reinterpret_cast<const uint4*>(residual + row * C + g * 8))
                              : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = r + static_cast<long long>(u) * rpp;
      if (row >= r1) break;
      float v[8], rv[8];
      unpack8(rx[u], v);
      unpack8(rr[u], rv);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], sc[i], sh[i]) + rv[i];
      if (relu) {
        if (relu_mask != nullptr) {   // 1 bit per element: what the backward needs of a residual BatchNorm's output
          unsigned int m = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) m |= (v[i] > 0.f ? 1u : 0u) << i;
          relu_mask[row * G + g] = static_cast<unsigned char>(m);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
      }
      store8(out + row * C + g * 8, v);
    }
  }
}

// eval-mode BN: running statistics instead of batch statistics
template <bool HAS_RES>
__global__ void __launch_bounds__(256)
bn_eval_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ residual, long long M, int C,
               int rows_per_cta, const float* __restrict__ running_mean, const float* __restrict__ running_var,
               const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int relu,
               __nv_bfloat16* __restrict__ out) {
  const int G = C >> 3;
  const int g = threadIdx.x % G, ty = threadIdx.x / G, rpp = 256 / G;
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = g * 8 + i;
    sc[i] = gamma[c] * rsqrtf(running_var[c] + eps);
    sh[i] = beta[c] - running_mean[c] * sc[i];
  }
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long r1 = min(r0 + rows_per_cta, M);
  // independent 16-byte loads in flight per thread: 8 on the single input stream, 4 + 4 with a residual (what is
  // outstanding per SM sets the bandwidth of this pass: 65 KB gave 4.4 TB/s)
  constexpr int U = HAS_RES ? 4 : 8;
  for (long long r = r0 + ty; r < r1; r += static_cast<long long>(rpp) * U) {
    uint4 rx[U], rr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = r + static_cast<long long>(u) * rpp;
      const bool ok = row < r1;
      rx[u] = ok ? __ldg(reinterpret_cast<const uint4*>(x + row * C + g * 8)) : make_uint4(0, 0, 0, 0);
      rr[u] = (HAS_RES && ok) ? __ldg(reinterpret_cast<const uint4*>(residual + row * C + g * 8))
                              : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = r + static_cast<long long>(u) * rpp;
      if (row >= r1) break;
      float v[8], rv[8];
      unpack8(rx[u], v);
      unpack8(rr[u], rv);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], sc[i], sh[i]) + rv[i];
      if (relu) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
      }
      store8(out + row * C + g * 8, v);
    }
  }
}

// backward pass 1: dbeta = sum dz, dgamma = sum dz * xhat, dz = dout o (out > 0) when relu
template <int MASK_SRC>
__global__ void __launch_bounds__(256, 2)
bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ out,
                     const unsigned char* __restrict__ relu_mask,
                     const __nv_bfloat16* __restrict__ x, long long M, int C, int rows_per_cta,
                     const float* __restrict__ mean, const float* __restrict__ rstd, int relu,
                     const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ scratch) {
  pdl_launch_dependents();   // see launch_pdl (common.cuh)
  pdl_wait();
  __shared__ float red[2][256][8];
  const int G = C >> 3;
  const int g = threadIdx.x % G, ty = threadIdx.x / G, rpp = 256 / G;
  // out == nullptr with relu: the mask is recomputed from x with the forward's own fma (saves one full read of the
  // activation; only possible when no residual entered the ReLU)
  constexpr bool HAS_OUT = MASK_SRC == 1;
  const bool recompute = relu && MASK_SRC == 0;
  float mu[8], rs[8], sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mu[i] = mean[g * 8 + i];
    rs[i] = rstd[g * 8 + i];
    sc[i] = recompute ? gamma[g * 8 + i] * rs[i] : 0.f;
    sh[i] = recompute ? fmaf(-mu[i], sc[i], beta[g * 8 + i]) : 0.f;
  }
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long r1 = min(r0 + rows_per_cta, M);
  float sg[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sb[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // bytes in flight per SM decide the bandwidth of these streaming passes: with two input streams instead of three
  // the unroll goes up so that the same ~100 KB stay outstanding
  constexpr int U = HAS_OUT ? 4 : 6;
  for (long long r = r0 + ty; r < r1; r += static_cast<long long>(rpp) * U) {
    uint4 rd[U], rx[U], ro[U];
    unsigned int rm[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + static_cast<long long>(u) * rpp;
      const bool ok = rr < r1;
      const long long off = rr * C + g * 8;
      rd[u] = ok ? __ldg(reinterpret_cast<const uint4*>(dout + off)) : make_uint4(0, 0, 0, 0);
      rx[u] = ok ? __ldg(reinterpret_cast<const uint4*>(x + off)) : make_uint4(0, 0, 0, 0);
      if constexpr (HAS_OUT)
        ro[u] = (ok && relu) ? __ldg(reinterpret_cast<const uint4*>(out + off)) : make_uint4(0, 0, 0, 0);
      if constexpr (MASK_SRC == 2) rm[u] = ok ? __ldg(relu_mask + rr * G + g) : 0u;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float d[8], xv[8];
      unpack8(rd[u], d);
      unpack8(rx[u], xv);
      if constexpr (MASK_SRC == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = (rm[u] >> i) & 1u ? d[i] : 0.f;
      } else if (recompute) {
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = fmaf(xv[i], sc[i], sh[i]) > 0.f ? d[i] : 0.f;
      } else if (HAS_OUT && relu) {
        float o[8];
        unpack8(ro[u], o);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = o[i] > 0.f ? d[i] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sb[i] += d[i];
        sg[i] = fmaf(d[i], (xv[i] - mu[i]) * rs[i], sg[i]);
      }
    }
  }
  column_reduce_finish(red, sg, sb, G, g, ty, rpp, C, scratch);
}

// backward pass 2: dx = gamma * rstd * (dz - dbeta/M - xhat * dgamma/M); optional dz copy for the identity branch;
// CTA 0 accumulates the parameter gradients.
template <int MASK_SRC>
__global__ void __launch_bounds__(256, MASK_SRC == 1 ? 3 : 2)
bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ out,
                    const unsigned char* __restrict__ relu_mask,
                    const __nv_bfloat16* __restrict__ x, long long M, int C, int rows_per_cta,
                    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                    const float* __restrict__ beta, int relu, const float* __restrict__ dgamma_sum,
                    const float* __restrict__ dbeta_sum,
                    __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dz_out, float* __restrict__ dgamma,
                    float* __restrict__ dbeta) {
  pdl_launch_dependents();   // see launch_pdl (common.cuh)
  pdl_wait();
  const int G = C >> 3;
  const int g = threadIdx.x % G, ty = threadIdx.x / G, rpp = 256 / G;
  float mu[8], rs[8], k0[8], k1[8], k2[8], sh[8];
  const float invM = 1.f / static_cast<float>(M);
  constexpr bool HAS_OUT = MASK_SRC == 1;
  const bool recompute = relu && MASK_SRC == 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = g * 8 + i;
    mu[i] = mean[c];
    rs[i] = rstd[c];
    const float dg = dgamma_sum[c], db = dbeta_sum[c];
    k0[i] = gamma[c] * rs[i];
    sh[i] = recompute ? fmaf(-mu[i], k0[i], beta[c]) : 0.f;
    k1[i] = db * invM;
    k2[i] = dg * invM;
    if (blockIdx.x == 0 && ty == 0) {
      dgamma[c] += dg;
      dbeta[c] += db;
    }
  }
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long r1 = min(r0 + rows_per_cta, M);
  constexpr int U = HAS_OUT ? 2 : (MASK_SRC == 2 ? 4 : 6);   // 16-byte loads in flight: 3 CTAs x 2 x 3 streams, or 2 CTAs x 4..6 x 2
  for (long long r = r0 + ty; r < r1; r += static_cast<long long>(rpp) * U) {
    uint4 rd[U], rx[U], ro[U];
    unsigned int rm[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + static_cast<long long>(u) * rpp;
      const bool ok = rr < r1;
      const long long off = rr * C + g * 8;
      rd[u] = ok ? __ldg(reinterpret_cast<const uint4*>(dout + off)) : make_uint4(0, 0, 0, 0);
      rx[u] = ok ? __ldg(reinterpret_cast<const uint4*>(x + off)) : make_uint4(0, 0, 0, 0);
      if constexpr (HAS_OUT)
        ro[u] = (ok && relu) ? __ldg(reinterpret_cast<const uint4*>(out + off)) : make_uint4(0, 0, 0, 0);
      if constexpr (MASK_SRC == 2) rm[u] = ok ? __ldg(relu_mask + rr * G + g) : 0u;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + static_cast<long long>(u) * rpp;
      if (rr >= r1) break;
      float d[8], xv[8];
      unpack8(rd[u], d);
      unpack8(rx[u], xv);
      if constexpr (MASK_SRC == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = (rm[u] >> i) & 1u ? d[i] : 0.f;
      } else if (recompute) {
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = fmaf(xv[i], k0[i], sh[i]) > 0.f ? d[i] : 0.f;
      } else if (HAS_OUT && relu) {
        float o[8];
        unpack8(ro[u], o);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = o[i] > 0.f ? d[i] : 0.f;
      }
      if (dz_out != nullptr) store8(dz_out + rr * C + g * 8, d);
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = k0[i] * (d[i] - k1[i] - (xv[i] - mu[i]) * rs[i] * k2[i]);
      store8(dx + rr * C + g * 8, o);
    }
  }
}

// Both backward passes in ONE launch for tensors the L2 holds (dout + x <= ~100 MB: ResNet layers 3-4, the narrow
// layer-2 tensors): every CTA reduces its rows, publishes its partial sums, waits on a grid-wide counter, then applies
// to the SAME rows -- the second read of dout and x comes from L2 instead of DRAM, and one launch (plus the gap and
// the cold start of a second kernel over a 13-50 MB tensor) disappears.  Launched cooperatively (all CTAs co-resident:
// grid = SMs x occupancy).  scratch: the replica rows and the counter of column_reduce_finish, zeroed by the caller.
template <int MASK_SRC>
__global__ void __launch_bounds__(256, 2)
bn_bwd_fused_kernel(const __nv_bfloat16* __restrict__ dout, const unsigned char* __restrict__ relu_mask,
                    const __nv_bfloat16* __restrict__ x, long long M, int C, int rows_per_cta,
                    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                    const float* __restrict__ beta, int relu, float* __restrict__ scratch,
                    __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dz_out, float* __restrict__ dgamma,
                    float* __restrict__ dbeta) {
  static_assert(MASK_SRC == 0 || MASK_SRC == 2, "mask from x (recomputed) / none, or the 1-bit mask");
  __shared__ float red[2][256][8];
  const int G = C >> 3;
  const int g = threadIdx.x % G, ty = threadIdx.x / G, rpp = 256 / G;
  const bool recompute = relu && MASK_SRC == 0;
  float mu[8], rs[8], k0[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mu[i] = mean[g * 8 + i];
    rs[i] = rstd[g * 8 + i];
    k0[i] = gamma[g * 8 + i] * rs[i];
    sh[i] = recompute ? fmaf(-mu[i], k0[i], beta[g * 8 + i]) : 0.f;
  }
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long r1 = min(r0 + rows_per_cta, M);
  // ---- pass 1: dbeta = sum dz, dgamma = sum dz * xhat over this CTA's rows
  float sg[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sb[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  constexpr int U = 6;
  for (long long r = r0 + ty; r < r1; r += static_cast<long long>(rpp) * U) {
    uint4 rd[U], rx[U];
    unsigned int rm[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + static_cast<long long>(u) * rpp;
      const bool ok = rr < r1;
      const long long off = rr * C + g * 8;
      rd[u] = ok ? __ldcg(reinterpret_cast<const uint4*>(dout + off)) : make_uint4(0, 0, 0, 0);
      rx[u] = ok ? __ldcg(reinterpret_cast<const uint4*>(x + off)) : make_uint4(0, 0, 0, 0);
      if constexpr (MASK_SRC == 2) rm[u] = ok ? __ldg(relu_mask + rr * G + g) : 0u;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float d[8], xv[8];
      unpack8(rd[u], d);
      unpack8(rx[u], xv);
      if constexpr (MASK_SRC == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = (rm[u] >> i) & 1u ? d[i] : 0.f;
      } else if (recompute) {
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = fmaf(xv[i], k0[i], sh[i]) > 0.f ? d[i] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sb[i] += d[i];
        sg[i] = fmaf(d[i], (xv[i] - mu[i]) * rs[i], sg[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[0][threadIdx.x][i] = sg[i];
    red[1][threadIdx.x][i] = sb[i];
  }
  __syncthreads();
  if (ty == 0) {
    for (int t = 1; t < rpp; ++t)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sg[i] += red[0][t * G + g][i];
        sb[i] += red[1][t * G + g][i];
      }
    float* rep = scratch + static_cast<size_t>(blockIdx.x % BN_REPLICAS) * 2 * C;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(rep + g * 8 + i, sg[i]);
      atomicAdd(rep + C + g * 8 + i, sb[i]);
    }
  }
  // ---- grid-wide barrier (all CTAs are co-resident: cooperative launch)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* counter = reinterpret_cast<unsigned int*>(scratch + static_cast<size_t>(BN_REPLICAS + 1) * 2 * C);
    atomicAdd(counter, 1u);
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
      if (seen < gridDim.x) __nanosleep(64);
    } while (seen < gridDim.x);
  }
  __syncthreads();
  // ---- pass 2: dx = gamma * rstd * (dz - dbeta/M - xhat * dgamma/M) over the same rows (L2-resident now)
  const float invM = 1.f / static_cast<float>(M);
  float k1[8], k2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = g * 8 + i;
    float dg = 0.f, db = 0.f;
#pragma unroll
    for (int r = 0; r < BN_REPLICAS; ++r) {
      dg += __ldcg(scratch + static_cast<size_t>(r) * 2 * C + c);
      db += __ldcg(scratch + static_cast<size_t>(r) * 2 * C + C + c);
    }
    k1[i] = db * invM;
    k2[i] = dg * invM;
    if (blockIdx.x == 0 && ty == 0) {
      dgamma[c] += dg;
      dbeta[c] += db;
    }
  }
  for (long long r = r0 + ty; r < r1; r += static_cast<long long>(rpp) * U) {
    uint4 rd[U], rx[U];
    unsigned int rm[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + static_cast<long long>(u) * rpp;
      const bool ok = rr < r1;
      const long long off = rr * C + g * 8;
      rd[u] = ok ? __ldcg(reinterpret_cast<const uint4*>(dout + off)) : make_uint4(0, 0, 0, 0);
      rx[u] = ok ? __ldcg(reinterpret_cast<const uint4*>(x + off)) : make_uint4(0, 0, 0, 0);
      if constexpr (MASK_SRC == 2) rm[u] = ok ? __ldg(relu_mask + rr * G + g) : 0u;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + static_cast<long long>(u) * rpp;
      if (rr >= r1) break;
      float d[8], xv[8];
      unpack8(rd[u], d);
      unpack8(rx[u], xv);
      if constexpr (MASK_SRC == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = (rm[u] >> i) & 1u ? d[i] : 0.f;
      } else if (recompute) {
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = fmaf(xv[i], k0[i], sh[i]) > 0.f ? d[i] : 0.f;
      }
      if (dz_out != nullptr) store8(dz_out + rr * C + g * 8, d);
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = k0[i] * (d[i] - k1[i] - (xv[i] - mu[i]) * rs[i] * k2[i]);
      store8(dx + rr * C + g * 8, o);
    }
  }
}

// ------------------------------------------------------------------ pooling
// 3x3 / stride 2 / pad 1 max pooling; argmax (0..8, first maximum in (kh,kw) scan order like ATen) kept for bwd.
// All nine window loads are issued before the first compare (predicated, no data-dependent control flow): the earlier
// loop-with-continue form exposed one load latency per tap (237 us for 565 MB).
__global__ void __launch_bounds__(128)
maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, int N, int H, int W, int C, int Ho, int Wo,
                   __nv_bfloat16* __restrict__ out, uint8_t* __restrict__ argmax) {
  // blockIdx.x = output row (n, ho); blockIdx.y * blockDim.x + threadIdx.x = wo * G + g: one 32-bit division per thread
  const int G = C >> 3;
  const int col = blockIdx.y * blockDim.x + threadIdx.x;
  if (col < Wo * G) {
    const int wo = col / G, g = col - wo * G;
    const int n = blockIdx.x / Ho, ho = blockIdx.x - n * Ho;
    uint4 raw[9];
    bool ok[9];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int hi = ho * 2 - 1 + kh;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int wi = wo * 2 - 1 + kw;
        ok[kh * 3 + kw] = hi >= 0 && hi < H && wi >= 0 && wi < W;
        raw[kh * 3 + kw] = ok[kh * 3 + kw]
            ? __ldg(reinterpret_cast<const uint4*>(x + ((static_cast<long long>(n) * H + hi) * W + wi) * C + g * 8))
            : make_uint4(0, 0, 0, 0);
      }
    }
    float best[8];
    int arg[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { best[k] = -INFINITY; arg[k] = 0; }
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      float v[8];
      unpack8(raw[tap], v);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (ok[tap] && v[k] > best[k]) { best[k] = v[k]; arg[k] = tap; }
    }
    const long long o = ((static_cast<long long>(n) * Ho + ho) * Wo + wo) * C + g * 8;
    store8(out + o, best);
    uint2 a;
    a.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
    a.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
    *reinterpret_cast<uint2*>(argmax + o) = a;
  }
}
// Stem tail in one pass: train-mode BatchNorm (statistics from the stem convolution's epilogue) + ReLU + the 3x3 / 2
// max pooling, straight from the convolution output.  The normalised 112 x 112 activation (411 MB at batch 256) is never
// written or re-read -- nothing in the backward needs it (BatchNorm's backward recomputes the ReLU mask from x, the
// pooling's backward needs the argmax only): 30.24 -> 30.0 ms per config-2 step.  Values are rounded to bf16
// before the comparison, so pooled values AND argmax are bit-identical to b200mm_batchnorm_fwd_stats followed by
// b200mm_maxpool3x3s2_fwd.
__global__ void __launch_bounds__(128)
bn_relu_maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, int N, int H, int W, int C, int Ho, int Wo,
                           const float* __restrict__ sum, const float* __restrict__ sumsq,
                           const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
                           __nv_bfloat16* __restrict__ out, uint8_t* __restrict__ argmax, float* __restrict__ mean_out,
                           float* __restrict__ rstd_out, float* __restrict__ running_mean,
                           float* __restrict__ running_var) {
  const int G = C >> 3;
  const int col = blockIdx.y * blockDim.x + threadIdx.x;
  if (col < Wo * G) {
    const int wo = col / G, g = col - wo * G;
    const int n = blockIdx.x / Ho, ho = blockIdx.x - n * Ho;
    const long long M = static_cast<long long>(N) * H * W;
    float sc[8], sh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = g * 8 + i;
      const float mean = sum[c] / M;
      const float var = fmaxf(sumsq[c] / M - mean * mean, 0.f);
      const float rstd = rsqrtf(var + eps);
      sc[i] = gamma[c] * rstd;
      sh[i] = fmaf(-mean, sc[i], beta[c]);   // same operations as bn_apply_kernel: the backward recomputes the mask
      if (blockIdx.x == 0 && wo == 0) {
        mean_out[c] = mean;
        rstd_out[c] = rstd;
        if (running_mean != nullptr) {
          const float unbiased = M > 1 ? var * (static_cast<float>(M) / static_cast<float>(M - 1)) : var;
          running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
          running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
        }
      }
    }
    uint4 raw[9];
    bool ok[9];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int hi = ho * 2 - 1 + kh;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int wi = wo * 2 - 1 + kw;
        ok[kh * 3 + kw] = hi >= 0 && hi < H && wi >= 0 && wi < W;
        raw[kh * 3 + kw] = ok[kh * 3 + kw]
            ? __ldg(reinterpret_cast<const uint4*>(x + ((static_cast<long long>(n) * H + hi) * W + wi) * C + g * 8))
            : make_uint4(0, 0, 0, 0);
      }
    }
    float best[8];
    int arg[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { best[k] = -INFINITY; arg[k] = 0; }
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      float v[8];
      unpack8(raw[tap], v);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float a = __bfloat162float(__float2bfloat16_rn(fmaxf(fmaf(v[k], sc[k], sh[k]), 0.f)));
        if (ok[tap] && a > best[k]) { best[k] = a; arg[k] = tap; }
      }
    }
    const long long o = ((static_cast<long long>(n) * Ho + ho) * Wo + wo) * C + g * 8;
    store8(out + o, best);
    uint2 a;
    a.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
    a.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
    *reinterpret_cast<uint2*>(argmax + o) = a;
  }
}
// Backward as a gather.  Input pixel (hi, wi) lies in the windows of output rows {hi/2} (hi even, tap row 1) or
// {(hi+1)/2, (hi-1)/2} (hi odd, tap rows 0 and 2), same for columns.  One thread owns the 2 x 2 input block
// (2a..2a+1, 2b..2b+1) x 8 channels: the four windows (a..a+1, b..b+1) cover all nine (pixel, window) pairs, so each
// gradient / argmax vector is fetched once per block instead of 2.25 times per pixel (the pass was L2-traffic bound).
__global__ void __launch_bounds__(128)
maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const uint8_t* __restrict__ argmax, int N, int H, int W,
                   int C, int Ho, int Wo, __nv_bfloat16* __restrict__ dx) {
  const int G = C >> 3;
  const int W2 = (W + 1) >> 1, H2 = (H + 1) >> 1;
  const int col = blockIdx.y * blockDim.x + threadIdx.x;
  if (col < W2 * G) {
    const int bq = col / G, g = col - bq * G;
    const int n = blockIdx.x / H2, aq = blockIdx.x - n * H2;
    uint4 rd[2][2];
    uint2 ra[2][2];
    bool ok[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        ok[i][j] = aq + i < Ho && bq + j < Wo;
        const long long o = ((static_cast<long long>(n) * Ho + aq + i) * Wo + bq + j) * C + g * 8;
        rd[i][j] = ok[i][j] ? __ldg(reinterpret_cast<const uint4*>(dout + o)) : make_uint4(0, 0, 0, 0);
        ra[i][j] = ok[i][j] ? __ldg(reinterpret_cast<const uint2*>(argmax + o)) : make_uint2(0xffffffffu, 0xffffffffu);
      }
    float d[2][2][8];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) unpack8(rd[i][j], d[i][j]);
#pragma unroll
    for (int ph = 0; ph < 2; ++ph)
#pragma unroll
      for (int pw = 0; pw < 2; ++pw) {
        const int hi = 2 * aq + ph, wi = 2 * bq + pw;
        if (hi >= H || wi >= W) continue;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            // window (aq + i, bq + j) reaches pixel (hi, wi) through tap kh = hi + 1 - 2 (aq + i) = ph + 1 - 2 i
            const int kh = ph + 1 - 2 * i, kw = pw + 1 - 2 * j;
            if (kh < 0 || kw < 0) continue;   // compile-time after unrolling
            const int code = kh * 3 + kw;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int ak = ((k < 4 ? ra[i][j].x : ra[i][j].y) >> ((k & 3) * 8)) & 0xff;
              if (ak == code) acc[k] += d[i][j][k];
            }
          }
        store8(dx + ((static_cast<long long>(n) * H + hi) * W + wi) * C + g * 8, acc);
      }
  }
}

// global average pool [N, HW, C] -> [N, C] and its backward
__global__ void __launch_bounds__(256)
avgpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, int N, int HW, int C, __nv_bfloat16* __restrict__ out) {
  const int G = C >> 3;
  const long long total = static_cast<long long>(N) * G;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % G);
    const long long n = i / G;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = 0; p < HW; ++p) {
      float v[8];
      load8(x + (n * HW + p) * C + g * 8, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += v[k];
    }
    const float inv = 1.f / HW;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] *= inv;
    store8(out + n * C + g * 8, acc);
  }
}
__global__ void __launch_bounds__(256)
avgpool_bwd_kernel(const __nv_bfloat16* __restrict__ dout, int N, int HW, int C, __nv_bfloat16* __restrict__ dx) {
  const int G = C >> 3;
  const long long total = static_cast<long long>(N) * HW * G;
  const float inv = 1.f / HW;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % G);
    const long long n = i / G / HW;
    float v[8];
    load8(dout + n * C + g * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] *= inv;
    store8(dx + (i / G) * C + g * 8, v);
  }
}

// ------------------------------------------------------------------ im2col / col2im (NHWC)
// cols[m, (kh*KW + kw)*C + c] = x[n, ho*stride - pad + kh, wo*stride - pad + kw, c]  (0 outside)
__global__ void __launch_bounds__(256)
im2col_kernel(const __nv_bfloat16* __restrict__ x, int N, int H, int W, int C, int KH, int KW, int stride, int pad,
              int Ho, int Wo, __nv_bfloat16* __restrict__ cols) {
  const int G = C >> 3;
  const int taps = KH * KW;
  const long long total = static_cast<long long>(N) * Ho * Wo * G;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % G);
    long long t = i / G;
    const long long m = t;
    const int wo = static_cast<int>(t % Wo); t /= Wo;
    const int ho = static_cast<int>(t % Ho);
    const int n = static_cast<int>(t / Ho);
    // all taps of this (pixel, 8-channel chunk): the loads are independent, issue them before the stores
    for (int tap0 = 0; tap0 < taps; tap0 += 9) {
      uint4 v[9];
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const int tap = tap0 + j;
        const int hi = ho * stride - pad + tap / KW, wi = wo * stride - pad + tap % KW;
        v[j] = make_uint4(0, 0, 0, 0);
        if (tap < taps && hi >= 0 && hi < H && wi >= 0 && wi < W)
          v[j] = __ldg(reinterpret_cast<const uint4*>(x + ((static_cast<long long>(n) * H + hi) * W + wi) * C + g * 8));
      }
#pragma unroll
      for (int j = 0; j < 9; ++j)
        if (tap0 + j < taps) *reinterpret_cast<uint4*>(cols + (m * taps + tap0 + j) * C + g * 8) = v[j];
    }
  }
}
// dx[n,hi,wi,c] = sum over taps of dcols[m(ho,wo), tap, c] (+ addend)
// 3x3 / pad 1 fast path, stride as a template parameter (no integer divisions in the tap loop), all 9 gathers of a
// thread issued before the first use.
template <int STRIDE>
__global__ void __launch_bounds__(256, 3)
col2im3x3_kernel(const __nv_bfloat16* __restrict__ dcols, const __nv_bfloat16* __restrict__ addend, int N, int H,
                 int W, int C, int Ho, int Wo, __nv_bfloat16* __restrict__ dx) {
  // blockIdx.x = input row (n, hi); blockIdx.y * blockDim.x + threadIdx.x = wi * G + g (32-bit index math only)
  const int G = C >> 3;
  const int col = blockIdx.y * blockDim.x + threadIdx.x;
  if (col < W * G) {
    const int wi = col / G, g = col - wi * G;
    const int n = blockIdx.x / H, hi = blockIdx.x - n * H;
    const long long pix = static_cast<long long>(blockIdx.x) * W + wi;
    const long long img_base = static_cast<long long>(n) * Ho * Wo;
    uint4 v[9];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int th = hi + 1 - kh;
      const bool hok = th >= 0 && (STRIDE == 1 || (th & 1) == 0) && (th / STRIDE) < Ho;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int tw = wi + 1 - kw;
        const bool ok = hok && tw >= 0 && (STRIDE == 1 || (tw & 1) == 0) && (tw / STRIDE) < Wo;
        v[kh * 3 + kw] = make_uint4(0, 0, 0, 0);
        if (ok) {
          const long long m = img_base + static_cast<long long>(th / STRIDE) * Wo + tw / STRIDE;
          v[kh * 3 + kw] = __ldg(reinterpret_cast<const uint4*>(dcols + (m * 9 + kh * 3 + kw) * C + g * 8));
        }
      }
    }
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (addend != nullptr) load8(addend + pix * C + g * 8, acc);
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      float f[8];
      unpack8(v[j], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += f[k];
    }
    store8(dx + pix * C + g * 8, acc);
  }
}
// generic fallback (any kernel size / stride / padding)
__global__ void __launch_bounds__(256)
col2im_kernel(const __nv_bfloat16* __restrict__ dcols, const __nv_bfloat16* __restrict__ addend, int N, int H, int W,
              int C, int KH, int KW, int stride, int pad, int Ho, int Wo, __nv_bfloat16* __restrict__ dx) {
  const int G = C >> 3;
  const int taps = KH * KW;
  const long long total = static_cast<long long>(N) * H * W * G;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % G);
    long long t = i / G;
    const long long pix = t;
    const int wi = static_cast<int>(t % W); t /= W;
    const int hi = static_cast<int>(t % H);
    const int n = static_cast<int>(t / H);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (addend != nullptr) load8(addend + pix * C + g * 8, acc);
    for (int kh = 0; kh < KH; ++kh) {
      const int th = hi + pad - kh;
      if (th < 0 || th % stride || th / stride >= Ho) continue;
      for (int kw = 0; kw < KW; ++kw) {
        const int tw = wi + pad - kw;
        if (tw < 0 || tw % stride || tw / stride >= Wo) continue;
        const long long m = (static_cast<long long>(n) * Ho + th / stride) * Wo + tw / stride;
        float v[8];
        load8(dcols + (m * taps + kh * KW + kw) * C + g * 8, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += v[k];
      }
    }
    store8(dx + pix * C + g * 8, acc);
  }
}
// Stem lowering straight from the fp32 NCHW image the reference's DataLoader yields:
// cols[m, (kh*KW+kw)*Cin + c] (row length Kp >= KH*KW*Cin, zero padded).
// One CTA per output row (n, ho): the KH x Cin input rows it needs are staged in shared memory with coalesced reads
// (as bf16), then the Wo x Kp output row block -- contiguous in memory -- is written with 16-byte stores.
__global__ void __launch_bounds__(256)
im2col_nchw_f32_kernel(const float* __restrict__ img, int N, int Cin, int H, int W, int KH, int KW, int stride,
                       int pad, int Ho, int Wo, int Kp, __nv_bfloat16* __restrict__ cols) {
  extern __shared__ uint8_t smem_raw[];
  const int Wp = W + 2 * pad;
  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(smem_raw);          // [KH][Cin][Wp]
  int* lut = reinterpret_cast<int*>(smem_raw + ((static_cast<size_t>(KH) * Cin * Wp * 2 + 15) & ~size_t(15)));  // [Kp]
  const int n = blockIdx.x / Ho, ho = blockIdx.x - n * Ho;
  const int K = KH * KW * Cin;
  for (int idx = threadIdx.x; idx < KH * Cin * Wp; idx += blockDim.x) {
    const int wp = idx % Wp;
    const int c = (idx / Wp) % Cin;
    const int kh = idx / (Wp * Cin);
    const int hi = ho * stride - pad + kh, wi = wp - pad;
    float v = 0.f;
    if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = __ldg(img + ((static_cast<long long>(n) * Cin + c) * H + hi) * W + wi);
    tile[idx] = __float2bfloat16(v);
  }
  for (int kk = threadIdx.x; kk < Kp; kk += blockDim.x) {
    int off = -1;
    if (kk < K) {
      const int c = kk % Cin, tap = kk / Cin;
      off = ((tap / KW) * Cin + c) * Wp + tap % KW;
    }
    lut[kk] = off;
  }
  __syncthreads();
  const int chunks = Kp >> 3;
  __nv_bfloat16* out = cols + (static_cast<long long>(n) * Ho + ho) * Wo * Kp;
  for (int idx = threadIdx.x; idx < Wo * chunks; idx += blockDim.x) {
    const int wo = idx / chunks, ch = idx - wo * chunks;
    const int base = wo * stride;
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int o0 = lut[ch * 8 + 2 * k], o1 = lut[ch * 8 + 2 * k + 1];
      const uint16_t lo = o0 >= 0 ? *reinterpret_cast<const uint16_t*>(tile + o0 + base) : uint16_t(0);
      const uint16_t hi = o1 >= 0 ? *reinterpret_cast<const uint16_t*>(tile + o1 + base) : uint16_t(0);
      w[k] = static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
    }
    *reinterpret_cast<uint4*>(out + static_cast<long long>(wo) * Kp + ch * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
// spatial subsampling for stride-2 1x1 convolutions: out[n,ho,wo,:] = x[n,ho*s,wo*s,:]; and its transpose (zero fill)
__global__ void __launch_bounds__(256)
subsample_kernel(const __nv_bfloat16* __restrict__ x, int N, int H, int W, int C, int stride, int Ho, int Wo,
                 __nv_bfloat16* __restrict__ out) {
  const int G = C >> 3;
  const long long total = static_cast<long long>(N) * Ho * Wo * G;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % G);
    long long t = i / G;
    const int wo = static_cast<int>(t % Wo); t /= Wo;
    const int ho = static_cast<int>(t % Ho);
    const int n = static_cast<int>(t / Ho);
    *reinterpret_cast<uint4*>(out + (i / G) * C + g * 8) = *reinterpret_cast<const uint4*>(
        x + ((static_cast<long long>(n) * H + ho * stride) * W + wo * stride) * C + g * 8);
  }
}
// dx[n,hi,wi,:] = addend[n,hi,wi,:] + (hi,wi on the stride grid ? dsub[n,hi/s,wi/s,:] : 0)
__global__ void __launch_bounds__(256)
upsample_add_kernel(const __nv_bfloat16* __restrict__ dsub, const __nv_bfloat16* __restrict__ addend, int N, int H,
                    int W, int C, int stride, int Ho, int Wo, __nv_bfloat16* __restrict__ dx) {
  // blockIdx.x = input row (n, hi); blockIdx.y * blockDim.x + threadIdx.x = wi * G + g; both loads issued together
  const int G = C >> 3;
  const int col = blockIdx.y * blockDim.x + threadIdx.x;
  if (col < W * G) {
    const int wi = col / G, g = col - wi * G;
    const int n = blockIdx.x / H, hi = blockIdx.x - n * H;
    const long long off = (static_cast<long long>(blockIdx.x) * W + wi) * C + g * 8;
    const bool on_grid = hi % stride == 0 && wi % stride == 0 && hi / stride < Ho && wi / stride < Wo;
    const uint4 ra = addend != nullptr ? __ldg(reinterpret_cast<const uint4*>(addend + off)) : make_uint4(0, 0, 0, 0);
    const uint4 rs = on_grid ? __ldg(reinterpret_cast<const uint4*>(
                                   dsub + ((static_cast<long long>(n) * Ho + hi / stride) * Wo + wi / stride) * C + g * 8))
                             : make_uint4(0, 0, 0, 0);
    float acc[8], v[8];
    unpack8(ra, acc);
    unpack8(rs, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += v[k];
    store8(dx + off, acc);
  }
}

static bool bn_shape_ok(long long M, int C) {
  const int G = C >> 3;
  return M > 0 && C >= 8 && (C & 7) == 0 && G <= 256 && (G & (G - 1)) == 0;
}
// floats of scratch a BatchNorm launch needs: replicas + final row + ticket (rounded up)
static size_t bn_scratch_floats(int C) { return static_cast<size_t>(BN_REPLICAS + 1) * 2 * C + 32; }
// resident 256-thread CTAs per SM of a kernel (register-limited: 2..3 for the apply kernels).  Grids are sized to ONE
// wave of resident CTAs: with more, a [12544 x 512] tensor paid the per-CTA prologue and a DRAM round trip 3.5 times
// (20 us for 26 MB of traffic).
template <typename Kernel>
static int bn_occupancy(Kernel kernel) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, 256, 0) != cudaSuccess || n < 1) n = 1;
  return n;
}
static int bn_rows_per_cta(long long M, int C, int* grid, int ctas_per_sm) {
  const DeviceInfo& dev = device_info();
  const int rpp = 256 / (C >> 3);
  const long long target = static_cast<long long>(dev.num_sms > 0 ? dev.num_sms : 148) * ctas_per_sm;
  long long rows = ceil_div(M, target);
  rows = ceil_div(rows, static_cast<long long>(rpp)) * rpp;
  *grid = static_cast<int>(ceil_div(M, rows));
  return static_cast<int>(rows);
}

}  // namespace b200

using namespace b200;

// Training-mode BatchNorm over x [M, C] (NHWC flattened): out = act(BN(x) (+ residual)).
// scratch: fp32 workspace of at least 18*C + 32 floats (zeroed here).  mean_out / rstd_out: fp32 [C] saved for the backward.
// running_mean / running_var (nullable) receive the momentum update.
B200MM_API int b200mm_batchnorm_fwd(const void* x, const void* residual, long long M, int C, const float* gamma,
                                    const float* beta, float eps, float momentum, int relu, void* out, float* mean_out,
                                    float* rstd_out, float* running_mean, float* running_var, float* scratch,
                                    void* stream) {
  if (!bn_shape_ok(M, C)) return B200MM_ERR_BAD_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(float) * bn_scratch_floats(C), s);
  if (e != cudaSuccess) return static_cast<int>(e);
  int grid, rgrid;
  static const int occ_res = bn_occupancy(bn_apply_kernel<true>), occ_plain = bn_occupancy(bn_apply_kernel<false>);
  const int rows = bn_rows_per_cta(M, C, &grid, residual != nullptr ? occ_res : occ_plain);
  const int rrows = bn_rows_per_cta(M, C, &rgrid, 3);   // reductions: fewer, fatter CTAs (8 loads in flight / thread)
  const float* fin = scratch + static_cast<size_t>(BN_REPLICAS) * 2 * C;
  bn_stats_kernel<<<rgrid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), M, C, rrows, scratch);
  B200MM_CHECK_LAUNCH();
  if (residual != nullptr)
    bn_apply_kernel<true><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x),
                                               static_cast<const __nv_bfloat16*>(residual), M, C, rows, fin, fin + C,
                                               gamma, beta, eps, momentum, relu, static_cast<__nv_bfloat16*>(out),
                                               mean_out, rstd_out, running_mean, running_var, nullptr);
  else
    bn_apply_kernel<false><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), nullptr, M, C, rows, fin,
                                                fin + C, gamma, beta, eps, momentum, relu,
                                                static_cast<__nv_bfloat16*>(out), mean_out, rstd_out, running_mean,
                                                running_var, nullptr);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

// Same, with the column statistics already accumulated by the producing convolution's epilogue
// (b200mm_gemm_bf16 / b200mm_conv_fwd col_stats): col_stats = [sum(C) | sum of squares(C)], fp32.  One pass over x.
B200MM_API int b200mm_batchnorm_fwd_stats(const void* x, const void* residual, long long M, int C,
                                          const float* col_stats, const float* gamma, const float* beta, float eps,
                                          float momentum, int relu, void* out, float* mean_out, float* rstd_out,
                                          float* running_mean, float* running_var, unsigned char* relu_mask,
                                          void* stream) {
  if (!bn_shape_ok(M, C) || col_stats == nullptr) return B200MM_ERR_BAD_ARG;
  int grid;
  static const int occ_res = bn_occupancy(bn_apply_kernel<true>), occ_plain = bn_occupancy(bn_apply_kernel<false>);
  const int rows = bn_rows_per_cta(M, C, &grid, residual != nullptr ? occ_res : occ_plain);
  // launched as a programmatic dependent of the producing convolution: its CTAs are resident (waiting in
  // griddepcontrol.wait) when the convolution's last tile retires
  cudaError_t e;
  if (residual != nullptr)
    e = launch_pdl(bn_apply_kernel<true>, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream),
                   static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(residual), M, C, rows,
                   col_stats, col_stats + C, gamma, beta, eps, momentum, relu, static_cast<__nv_bfloat16*>(out),
                   mean_out, rstd_out, running_mean, running_var, relu_mask);
  else
    e = launch_pdl(bn_apply_kernel<false>, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream),
                   static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(nullptr), M, C, rows,
                   col_stats, col_stats + C, gamma, beta, eps, momentum, relu, static_cast<__nv_bfloat16*>(out),
                   mean_out, rstd_out, running_mean, running_var, relu_mask);
  if (e != cudaSuccess) return static_cast<int>(e);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

B200MM_API int b200mm_batchnorm_eval(const void* x, const void* residual, long long M, int C, const float* gamma,
                                     const float* beta, const float* running_mean, const float* running_var, float eps,
                                     int relu, void* out, void* stream) {
  if (!bn_shape_ok(M, C)) return B200MM_ERR_BAD_ARG;
  int grid;
  static const int occ_res = bn_occupancy(bn_eval_kernel<true>), occ_plain = bn_occupancy(bn_eval_kernel<false>);
  const int rows = bn_rows_per_cta(M, C, &grid, residual != nullptr ? occ_res : occ_plain);
  if (residual != nullptr)
    bn_eval_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(residual), M, C, rows, running_mean,
        running_var, gamma, beta, eps, relu, static_cast<__nv_bfloat16*>(out));
  else
    bn_eval_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), nullptr, M, C, rows, running_mean, running_var, gamma, beta, eps, relu,
        static_cast<__nv_bfloat16*>(out));
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

// BatchNorm backward.  dout: gradient w.r.t. the (post-activation) output; out: that output (ReLU mask; may be null
// when relu == 0); x: the BN input.  dx: gradient w.r.t. x; dz_out (nullable): gradient w.r.t. the pre-activation
// sum, i.e. what flows into the residual branch.  dgamma / dbeta accumulate.  With relu != 0 and out == nullptr the
// ReLU mask is recomputed from x, mean, rstd, gamma and beta (valid only if no residual was added before the ReLU).
B200MM_API int b200mm_batchnorm_bwd(const void* dout, const void* out, const void* x, long long M, int C,
                                    const float* mean, const float* rstd, const float* gamma, const float* beta,
                                    const unsigned char* relu_mask, int relu, void* dx, void* dz_out, float* dgamma,
                                    float* dbeta, float* scratch, void* stream) {
  if (!bn_shape_ok(M, C) || (relu && out == nullptr && beta == nullptr && relu_mask == nullptr))
    return B200MM_ERR_BAD_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(float) * bn_scratch_floats(C), s);
  if (e != cudaSuccess) return static_cast<int>(e);
  int grid, rgrid;
  static const int occ_apply[3] = {bn_occupancy(bn_bwd_apply_kernel<0>), bn_occupancy(bn_bwd_apply_kernel<1>),
                                   bn_occupancy(bn_bwd_apply_kernel<2>)};
  const int src = (relu && relu_mask != nullptr) ? 2 : (relu && out != nullptr) ? 1 : 0;
  // L2-resident tensors: one cooperative launch does both passes (b200mm_tune knob 2: largest tensor in MB, 0 = off)
  const long long fused_bytes = static_cast<long long>(g_tune[2]) << 20;
  if (src != 1 && M * C * 2 <= fused_bytes) {
    static const int occ_fused[2] = {bn_occupancy(bn_bwd_fused_kernel<0>), bn_occupancy(bn_bwd_fused_kernel<2>)};
    int fgrid;
    const int frows = bn_rows_per_cta(M, C, &fgrid, occ_fused[src == 2 ? 1 : 0] < 2 ? occ_fused[src == 2 ? 1 : 0] : 2);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(fgrid);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    static const bool coop = [] { const char* e = std::getenv("B200MM_BN_COOP"); return e == nullptr || e[0] != '0'; }();
    cfg.numAttrs = coop ? 1 : 0;
    const __nv_bfloat16* d_ = static_cast<const __nv_bfloat16*>(dout);
    const __nv_bfloat16* xx = static_cast<const __nv_bfloat16*>(x);
    __nv_bfloat16* dxp = static_cast<__nv_bfloat16*>(dx);
    __nv_bfloat16* dzp = static_cast<__nv_bfloat16*>(dz_out);
    cudaError_t le = src == 2
        ? cudaLaunchKernelEx(&cfg, bn_bwd_fused_kernel<2>, d_, relu_mask, xx, M, C, frows, mean, rstd, gamma, beta, relu,
                             scratch, dxp, dzp, dgamma, dbeta)
        : cudaLaunchKernelEx(&cfg, bn_bwd_fused_kernel<0>, d_, relu_mask, xx, M, C, frows, mean, rstd, gamma, beta, relu,
                             scratch, dxp, dzp, dgamma, dbeta);
    if (le != cudaSuccess) return static_cast<int>(le);
    B200MM_CHECK_LAUNCH();
    return B200MM_OK;
  }
  const int rows = bn_rows_per_cta(M, C, &grid, occ_apply[src]);
  const int rrows = bn_rows_per_cta(M, C, &rgrid, 2);
  const float* fin = scratch + static_cast<size_t>(BN_REPLICAS) * 2 * C;
  const __nv_bfloat16* dout_ = static_cast<const __nv_bfloat16*>(dout);
  const __nv_bfloat16* out_ = static_cast<const __nv_bfloat16*>(out);
  const __nv_bfloat16* x_ = static_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* dx_ = static_cast<__nv_bfloat16*>(dx);
  __nv_bfloat16* dz_ = static_cast<__nv_bfloat16*>(dz_out);
#define BN_BWD(SRC)                                                                                                 \
  do {                                                                                                              \
    bn_bwd_reduce_kernel<SRC><<<rgrid, 256, 0, s>>>(dout_, out_, relu_mask, x_, M, C, rrows, mean, rstd, relu, gamma, \
                                                    beta, scratch);                                                 \
    B200MM_CHECK_LAUNCH();                                                                                          \
    e = launch_pdl(bn_bwd_apply_kernel<SRC>, dim3(grid), dim3(256), 0, s, dout_, out_, relu_mask, x_, M, C, rows, mean, \
                   rstd, gamma, beta, relu, fin, fin + C, dx_, dz_, dgamma, dbeta);                                 \
    if (e != cudaSuccess) return static_cast<int>(e);                                                               \
    B200MM_CHECK_LAUNCH();                                                                                          \
  } while (0)
  if (src == 2) BN_BWD(2);
  else if (src == 1) BN_BWD(1);
  else BN_BWD(0);
#undef BN_BWD
  return B200MM_OK;
}

B200MM_API int b200mm_maxpool3x3s2_fwd(const void* x, int N, int H, int W, int C, void* out, void* argmax,
                                       void* stream) {
  if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 7)) return B200MM_ERR_BAD_ARG;
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  if (static_cast<long long>(N) * H > 0x7fffffffLL) return B200MM_ERR_BAD_ARG;
  maxpool_fwd_kernel<<<dim3(N * Ho, (Wo * (C >> 3) + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x), N, H, W, C, Ho, Wo,
                                                            static_cast<__nv_bfloat16*>(out),
                                                            static_cast<uint8_t*>(argmax));
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
// out[N, Ho, Wo, C] = maxpool3x3s2(relu(BN_train(x))) for x [N, H, W, C] with the column statistics of x already
// accumulated (col_stats = [sum(C) | sum of squares(C)]); argmax as b200mm_maxpool3x3s2_fwd; mean / rstd / running
// statistics as b200mm_batchnorm_fwd_stats.  Replaces bn1 + relu + maxpool of torchvision/models/resnet.py:268-271.
B200MM_API int b200mm_bn_relu_maxpool_fwd(const void* x, int N, int H, int W, int C, const float* col_stats,
                                          const float* gamma, const float* beta, float eps, float momentum, void* out,
                                          void* argmax, float* mean_out, float* rstd_out, float* running_mean,
                                          float* running_var, void* stream) {
  if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 7) || col_stats == nullptr) return B200MM_ERR_BAD_ARG;
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  bn_relu_maxpool_fwd_kernel<<<dim3(N * Ho, (Wo * (C >> 3) + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), N, H, W, C, Ho, Wo, col_stats, col_stats + C, gamma, beta, eps, momentum,
      static_cast<__nv_bfloat16*>(out), static_cast<uint8_t*>(argmax), mean_out, rstd_out, running_mean, running_var);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

B200MM_API int b200mm_maxpool3x3s2_bwd(const void* dout, const void* argmax, int N, int H, int W, int C, void* dx,
                                       void* stream) {
  if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 7)) return B200MM_ERR_BAD_ARG;
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  if (static_cast<long long>(N) * H > 0x7fffffffLL) return B200MM_ERR_BAD_ARG;
  maxpool_bwd_kernel<<<dim3(N * ((H + 1) / 2), (((W + 1) / 2) * (C >> 3) + 127) / 128), 128, 0,
                       static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(dout),
                                                            static_cast<const uint8_t*>(argmax), N, H, W, C, Ho, Wo,
                                                            static_cast<__nv_bfloat16*>(dx));
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
B200MM_API int b200mm_avgpool_fwd(const void* x, int N, int HW, int C, void* out, void* stream) {
  if (N <= 0 || HW <= 0 || C <= 0 || (C & 7)) return B200MM_ERR_BAD_ARG;
  avgpool_fwd_kernel<<<grid_for(static_cast<long long>(N) * (C >> 3), 256), 256, 0,
                       static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x), N, HW, C,
                                                            static_cast<__nv_bfloat16*>(out));
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
B200MM_API int b200mm_avgpool_bwd(const void* dout, int N, int HW, int C, void* dx, void* stream) {
  if (N <= 0 || HW <= 0 || C <= 0 || (C & 7)) return B200MM_ERR_BAD_ARG;
  avgpool_bwd_kernel<<<grid_for(static_cast<long long>(N) * HW * (C >> 3), 256), 256, 0,
                       static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(dout), N, HW, C,
                                                            static_cast<__nv_bfloat16*>(dx));
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

B200MM_API int b200mm_im2col_nhwc(const void* x, int N, int H, int W, int C, int KH, int KW, int stride, int pad,
                                  void* cols, void* stream) {
  if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 7) || KH <= 0 || KW <= 0 || stride <= 0 || pad < 0)
    return B200MM_ERR_BAD_ARG;
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  im2col_kernel<<<grid_for(static_cast<long long>(N) * Ho * Wo * (C >> 3), 256), 256, 0,
                  static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x), N, H, W, C, KH, KW, stride,
                                                       pad, Ho, Wo, static_cast<__nv_bfloat16*>(cols));
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
B200MM_API int b200mm_col2im_nhwc(const void* dcols, const void* addend, int N, int H, int W, int C, int KH, int KW,
                                  int stride, int pad, void* dx, void* stream) {
  if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 7) || KH <= 0 || KW <= 0 || stride <= 0 || pad < 0)
    return B200MM_ERR_BAD_ARG;
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  const int grid = grid_for(static_cast<long long>(N) * H * W * (C >> 3), 256);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* dc = static_cast<const __nv_bfloat16*>(dcols);
  const __nv_bfloat16* ad = static_cast<const __nv_bfloat16*>(addend);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(dx);
  const dim3 grid2(N * H, (W * (C >> 3) + 255) / 256);
  if (static_cast<long long>(N) * H > 0x7fffffffLL) return B200MM_ERR_BAD_ARG;
  if (KH == 3 && KW == 3 && pad == 1 && stride == 1)
    col2im3x3_kernel<1><<<grid2, 256, 0, s>>>(dc, ad, N, H, W, C, Ho, Wo, o);
  else if (KH == 3 && KW == 3 && pad == 1 && stride == 2)
    col2im3x3_kernel<2><<<grid2, 256, 0, s>>>(dc, ad, N, H, W, C, Ho, Wo, o);
  else
    col2im_kernel<<<grid, 256, 0, s>>>(dc, ad, N, H, W, C, KH, KW, stride, pad, Ho, Wo, o);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
B200MM_API int b200mm_im2col_nchw_f32(const float* img, int N, int Cin, int H, int W, int KH, int KW, int stride,
                                      int pad, int Kp, void* cols, void* stream) {
  if (N <= 0 || Cin <= 0 || H <= 0 || W <= 0 || KH <= 0 || KW <= 0 || stride <= 0 || pad < 0 || (Kp & 7) ||
      Kp < KH * KW * Cin)
    return B200MM_ERR_BAD_ARG;
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  const size_t smem = ((static_cast<size_t>(KH) * Cin * (W + 2 * pad) * 2 + 15) & ~size_t(15)) + sizeof(int) * Kp;
  if (smem > 48 * 1024) return B200MM_ERR_BAD_ARG;
  im2col_nchw_f32_kernel<<<N * Ho, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      img, N, Cin, H, W, KH, KW, stride, pad, Ho, Wo, Kp, static_cast<__nv_bfloat16*>(cols));
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
B200MM_API int b200mm_subsample_nhwc(const void* x, int N, int H, int W, int C, int stride, void* out, void* stream) {
  if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 7) || stride <= 0) return B200MM_ERR_BAD_ARG;
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  subsample_kernel<<<grid_for(static_cast<long long>(N) * Ho * Wo * (C >> 3), 256), 256, 0,
                     static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x), N, H, W, C, stride, Ho,
                                                          Wo, static_cast<__nv_bfloat16*>(out));
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
B200MM_API int b200mm_upsample_add_nhwc(const void* dsub, const void* addend, int N, int H, int W, int C, int stride,
                                        void* dx, void* stream) {
  if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 7) || stride <= 0) return B200MM_ERR_BAD_ARG;
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  if (static_cast<long long>(N) * H > 0x7fffffffLL) return B200MM_ERR_BAD_ARG;
  upsample_add_kernel<<<dim3(N * H, (W * (C >> 3) + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(dsub),
                                                             static_cast<const __nv_bfloat16*>(addend), N, H, W, C,
                                                             stride, Ho, Wo, static_cast<__nv_bfloat16*>(dx));
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
