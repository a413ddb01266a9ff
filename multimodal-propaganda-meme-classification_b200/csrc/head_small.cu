// Head-side kernels of the HEAD-script model (example_scripts/Multimodal_example_task2C.py:476-499, 562-685): the
// batch is the only long dimension here (B rows of 512..1536 features), so these are small, latency-bound kernels
// in fp32 arithmetic on bf16 activations:
//   BatchNorm1d (+ReLU) forward / backward            text_fc / caption_text_fc / ConcatAttention3  (:599-601, :479-490)
//   softmax gate  y = softmax(a) * x  fwd / bwd       ConcatAttention3.forward                      (:492-499)
//   Linear(512,1) + BatchNorm1d(1) + sigmoid focal    output_fc + criterion, one fused kernel       (:641-643, :167, :711)
//   ReLU backward mask                                CustomDenseNet161.fine_tune                   (:571-574)
#include "common.cuh"
#include "device_utils.cuh"

namespace b200 {

// BatchNorm input element: bf16, or fp32 where the signal is a small variation on a large offset (the 1536 -> 512
// "reduce" projection of ConcatAttention3 sees inputs scaled by a softmax over 1536 features: its bias would swamp
// the batch variation at bf16 resolution, and the BatchNorm behind it would amplify the rounding noise)
__device__ __forceinline__ float ld_elem(const void* x, long long i, int f32) {
  return f32 ? static_cast<const float*>(x)[i] : __bfloat162float(static_cast<const __nv_bfloat16*>(x)[i]);
}

// ------------------------------------------------------------------ BatchNorm1d over [B, C]
// One thread per channel, looping over the batch (consecutive threads read consecutive channels: coalesced).
__global__ void __launch_bounds__(128)
bn1d_fwd_kernel(const void* __restrict__ x, int x_f32, long long ldx, int B, int C, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, float momentum, int relu, int train,
                __nv_bfloat16* __restrict__ out, long long ldo, float* __restrict__ mean_out,
                float* __restrict__ rstd_out, float* __restrict__ running_mean, float* __restrict__ running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, rstd;
  if (train) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += ld_elem(x, b * ldx + c, x_f32);
    mean = s / B;
    float q = 0.f;
    for (int b = 0; b < B; ++b) {
      const float d = ld_elem(x, b * ldx + c, x_f32) - mean;
      q = fmaf(d, d, q);
    }
    const float var = q / B;   // biased variance normalises; the unbiased one feeds the running statistic
    rstd = rsqrtf(var + eps);
    if (mean_out) { mean_out[c] = mean; rstd_out[c] = rstd; }
    if (running_mean) {
      const float unbiased = B > 1 ? var * (static_cast<float>(B) / static_cast<float>(B - 1)) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
  } else {
    mean = running_mean[c];
    rstd = rsqrtf(running_var[c] + eps);
  }
  const float sc = gamma[c] * rstd, sh = fmaf(-mean, sc, beta[c]);
  for (int b = 0; b < B; ++b) {
    float v = fmaf(ld_elem(x, b * ldx + c, x_f32), sc, sh);
    if (relu) v = fmaxf(v, 0.f);
    out[b * ldo + c] = __float2bfloat16(v);
  }
}

// dx = gamma rstd (dz - mean(dz) - xhat mean(dz xhat)), dz = dout o (out > 0) when relu; dgamma / dbeta accumulate.
__global__ void __launch_bounds__(128)
bn1d_bwd_kernel(const __nv_bfloat16* __restrict__ dout, long long ldd, const __nv_bfloat16* __restrict__ out,
                long long ldo, const void* __restrict__ x, int x_f32, long long ldx, int B, int C,
                const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                int relu, __nv_bfloat16* __restrict__ dx, long long lddx, float* __restrict__ dgamma,
                float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mu = mean[c], rs = rstd[c];
  float sb = 0.f, sg = 0.f;
  for (int b = 0; b < B; ++b) {
    float d = __bfloat162float(dout[b * ldd + c]);
    if (relu && !(__bfloat162float(out[b * ldo + c]) > 0.f)) d = 0.f;
    sb += d;
    sg = fmaf(d, (ld_elem(x, b * ldx + c, x_f32) - mu) * rs, sg);
  }
  dgamma[c] += sg;
  dbeta[c] += sb;
  const float k0 = gamma[c] * rs, k1 = sb / B, k2 = sg / B;
  for (int b = 0; b < B; ++b) {
    float d = __bfloat162float(dout[b * ldd + c]);
    if (relu && !(__bfloat162float(out[b * ldo + c]) > 0.f)) d = 0.f;
    const float xh = (ld_elem(x, b * ldx + c, x_f32) - mu) * rs;
    dx[b * lddx + c] = __float2bfloat16(k0 * (d - k1 - xh * k2));
  }
}

// ------------------------------------------------------------------ softmax gate: y = softmax(a) * x, one warp per row
__global__ void __launch_bounds__(256)
softmax_gate_fwd_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ x, int B, int C,
                        float* __restrict__ w_out, __nv_bfloat16* __restrict__ y) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= B) return;
  const long long off = static_cast<long long>(row) * C;
  float mx = -INFINITY;
  for (int j = lane; j < C; j += 32) mx = fmaxf(mx, __bfloat162float(a[off + j]));
  mx = warp_max(mx);
  float s = 0.f;
  for (int j = lane; j < C; j += 32) s += __expf(__bfloat162float(a[off + j]) - mx);
  const float inv = 1.f / warp_sum(s);
  for (int j = lane; j < C; j += 32) {
    const float w = __expf(__bfloat162float(a[off + j]) - mx) * inv;
    w_out[off + j] = w;
    y[off + j] = __float2bfloat16(w * __bfloat162float(x[off + j]));
  }
}
// dy -> dx_direct = dy * w (gradient through the multiplicand), da = w (dw - sum_j dw_j w_j) with dw = dy * x
__global__ void __launch_bounds__(256)
softmax_gate_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ w,
                        const __nv_bfloat16* __restrict__ x, int B, int C, __nv_bfloat16* __restrict__ da,
                        __nv_bfloat16* __restrict__ dx_direct) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= B) return;
  const long long off = static_cast<long long>(row) * C;
  float dot = 0.f;
  for (int j = lane; j < C; j += 32)
    dot = fmaf(__bfloat162float(dy[off + j]) * __bfloat162float(x[off + j]), w[off + j], dot);
  dot = warp_sum(dot);
  for (int j = lane; j < C; j += 32) {
    const float g = __bfloat162float(dy[off + j]), wj = w[off + j];
    da[off + j] = __float2bfloat16(wj * (g * __bfloat162float(x[off + j]) - dot));
    dx_direct[off + j] = __float2bfloat16(g * wj);
  }
}

__global__ void relu_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y, long long n,
                                __nv_bfloat16* __restrict__ dx) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dx[i] = __bfloat162float(y[i]) > 0.f ? dy[i] : __float2bfloat16(0.f);
}

// ------------------------------------------------------------------ Linear(F,1) + BatchNorm1d(1) + sigmoid focal loss
// ONE CTA (1024 threads) -- the batch statistic of the single logit couples every sample.  B <= HEADBN_MAX_B.
//   z_i = feat_i . W + b ;  y_i = (z_i - mu) rstd g + beta  (train: batch mu / var, eval: running) ;  loss = mean focal(y_i)
// Outputs: logits y (fp32 [B]), loss (sum_i loss_i / B), correct (sigmoid(y) > 0.5 == label).  train: dfeat (bf16 [B,F]),
// dW / db / dg / dbeta accumulate, running stats updated.  dlogits_in (nullable): external dL/dy instead of the focal loss.
constexpr int HEADBN_MAX_B = 4096;
__global__ void __launch_bounds__(1024)
head_bn_focal_kernel(const __nv_bfloat16* __restrict__ feat, const float* __restrict__ W, const float* __restrict__ bias,
                     const float* __restrict__ bn_g, const float* __restrict__ bn_b, float* __restrict__ running_mean,
                     float* __restrict__ running_var, const long long* __restrict__ labels, int B, int F, float eps,
                     float momentum, float alpha, float gamma, int train, int bn_train,
                     const float* __restrict__ dlogits_in, float* __restrict__ logits, float* __restrict__ loss_out,
                     int* __restrict__ correct, __nv_bfloat16* __restrict__ dfeat, float* __restrict__ dW,
                     float* __restrict__ dbias, float* __restrict__ dg, float* __restrict__ dbeta) {
  __shared__ float z[HEADBN_MAX_B];    // z, then dL/dy, then dz
  __shared__ float xhs[HEADBN_MAX_B];  // normalised logit (z - mu) rstd
  __shared__ float red[32];
  __shared__ float stat[4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  auto block_sum = [&](float v) {
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = lane < nwarps ? red[lane] : 0.f;
    return warp_sum(t);   // every warp ends up with the total
  };
  for (int i = warp; i < B; i += nwarps) {
    const __nv_bfloat16* f = feat + static_cast<long long>(i) * F;
    float acc = 0.f;
    for (int j = lane; j < F; j += 32) acc = fmaf(__bfloat162float(f[j]), __ldg(W + j), acc);
    acc = warp_sum(acc);
    if (lane == 0) z[i] = acc + bias[0];
  }
  __syncthreads();
  float mu, rstd;
  if (bn_train) {
    float s = 0.f;
    for (int i = tid; i < B; i += blockDim.x) s += z[i];
    mu = block_sum(s) / B;
    float q = 0.f;
    for (int i = tid; i < B; i += blockDim.x) q = fmaf(z[i] - mu, z[i] - mu, q);
    const float var = block_sum(q) / B;
    rstd = rsqrtf(var + eps);
    if (tid == 0 && running_mean) {
      const float unbiased = B > 1 ? var * (static_cast<float>(B) / static_cast<float>(B - 1)) : var;
      running_mean[0] = (1.f - momentum) * running_mean[0] + momentum * mu;
      running_var[0] = (1.f - momentum) * running_var[0] + momentum * unbiased;
    }
  } else {
    mu = running_mean[0];
    rstd = rsqrtf(running_var[0] + eps);
  }
  const float g = bn_g[0], be = bn_b[0];
  // loss + dL/dy
  float lsum = 0.f, sdy = 0.f, sdyx = 0.f;
  int ok = 0;
  for (int i = tid; i < B; i += blockDim.x) {
    const float xh = (z[i] - mu) * rstd;
    const float x = fmaf(xh, g, be);
    logits[i] = x;
    float dy = 0.f;
    if (dlogits_in != nullptr) {
      dy = dlogits_in[i];
    } else if (labels != nullptr) {
      const float y = static_cast<float>(labels[i]);
      const float pr = 1.f / (1.f + expf(-x));
      const float ce = fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));
      const float pt = pr * y + (1.f - pr) * (1.f - y);
      const float one_m = 1.f - pt;
      const float mod = powf(one_m, gamma);
      const float at = alpha >= 0.f ? alpha * y + (1.f - alpha) * (1.f - y) : 1.f;
      lsum += at * ce * mod;
      const float dce = pr - y;
      const float dpt = (2.f * y - 1.f) * pr * (1.f - pr);
      const float dmod = one_m > 0.f ? -gamma * powf(one_m, gamma - 1.f) * dpt : 0.f;
      dy = at * (dce * mod + ce * dmod) / B;
      ok += ((pr > 0.5f) == (labels[i] != 0)) ? 1 : 0;
    }
    sdy += dy;
    sdyx = fmaf(dy, xh, sdyx);
    xhs[i] = xh;
    z[i] = dy;
  }
  lsum = block_sum(lsum);
  const float tot_ok = block_sum(static_cast<float>(ok));
  if (tid == 0 && labels != nullptr && dlogits_in == nullptr) {
    loss_out[0] += lsum / B;
    correct[0] += static_cast<int>(tot_ok + 0.5f);
  }
  if (!train) return;
  sdy = block_sum(sdy);
  sdyx = block_sum(sdyx);
  if (tid == 0) {
    dg[0] += sdyx;
    dbeta[0] += sdy;
  }
  // dz_i = g rstd (dy_i - mean(dy) - xh_i mean(dy xh))   (bn_train)   |   g rstd dy_i   (eval statistics)
  const float k0 = g * rstd, k1 = bn_train ? sdy / B : 0.f, k2 = bn_train ? sdyx / B : 0.f;
  float sdz = 0.f;
  __syncthreads();
  for (int i = tid; i < B; i += blockDim.x) {
    const float dz = k0 * (z[i] - k1 - xhs[i] * k2);
    z[i] = dz;
    sdz += dz;
  }
  sdz = block_sum(sdz);
  if (tid == 0) dbias[0] += sdz;
  __syncthreads();
  // dfeat_i = dz_i W ;  dW = sum_i dz_i feat_i
  for (int j = tid; j < F; j += blockDim.x) {
    const float wj = W[j];
    float acc = 0.f;
    for (int i = 0; i < B; ++i) {
      const float dz = z[i];
      acc = fmaf(dz, __bfloat162float(feat[static_cast<long long>(i) * F + j]), acc);
      dfeat[static_cast<long long>(i) * F + j] = __float2bfloat16(dz * wj);
    }
    dW[j] += acc;
  }
}

}  // namespace b200

using namespace b200;

// BatchNorm1d (+ReLU) over x [B, C] (bf16, or fp32 when x_f32 != 0; row strides ldx / ldo elements).  train != 0: batch statistics, saved in
// mean / rstd (nullable) and folded into the running statistics; train == 0: running statistics.
B200MM_API int b200mm_bn1d_fwd(const void* x, int x_f32, long long ldx, int B, int C, const float* gamma, const float* beta,
                               float eps, float momentum, int relu, int train, void* out, long long ldo, float* mean,
                               float* rstd, float* running_mean, float* running_var, void* stream) {
  if (B <= 0 || C <= 0 || (!train && (running_mean == nullptr || running_var == nullptr))) return B200MM_ERR_BAD_ARG;
  bn1d_fwd_kernel<<<ceil_div(C, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      x, x_f32, ldx, B, C, gamma, beta, eps, momentum, relu, train,
      static_cast<__nv_bfloat16*>(out), ldo, mean, rstd, running_mean, running_var);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
B200MM_API int b200mm_bn1d_bwd(const void* dout, long long ldd, const void* out, long long ldo, const void* x,
                               int x_f32, long long ldx, int B, int C, const float* mean, const float* rstd, const float* gamma,
                               int relu, void* dx, long long lddx, float* dgamma, float* dbeta, void* stream) {
  if (B <= 0 || C <= 0 || (relu && out == nullptr)) return B200MM_ERR_BAD_ARG;
  bn1d_bwd_kernel<<<ceil_div(C, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dout), ldd, static_cast<const __nv_bfloat16*>(out), ldo,
      x, x_f32, ldx, B, C, mean, rstd, gamma, relu, static_cast<__nv_bfloat16*>(dx), lddx,
      dgamma, dbeta);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
// y[B,C] = softmax(a[B,C], dim=1) * x[B,C]; the weights are kept (fp32) for the backward.
B200MM_API int b200mm_softmax_gate_fwd(const void* a, const void* x, int B, int C, float* w, void* y, void* stream) {
  if (B <= 0 || C <= 0) return B200MM_ERR_BAD_ARG;
  softmax_gate_fwd_kernel<<<ceil_div(B, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(x), B, C, w,
      static_cast<__nv_bfloat16*>(y));
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
B200MM_API int b200mm_softmax_gate_bwd(const void* dy, const float* w, const void* x, int B, int C, void* da,
                                       void* dx_direct, void* stream) {
  if (B <= 0 || C <= 0) return B200MM_ERR_BAD_ARG;
  softmax_gate_bwd_kernel<<<ceil_div(B, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dy), w, static_cast<const __nv_bfloat16*>(x), B, C,
      static_cast<__nv_bfloat16*>(da), static_cast<__nv_bfloat16*>(dx_direct));
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
B200MM_API int b200mm_relu_bwd(const void* dy, const void* y, long long n, void* dx, void* stream) {
  if (n <= 0) return B200MM_ERR_BAD_ARG;
  relu_bwd_kernel<<<static_cast<int>(ceil_div(n, 256LL)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(y), n, static_cast<__nv_bfloat16*>(dx));
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
// output_fc of the HEAD script (Linear(F,1) + BatchNorm1d(1), squeeze) fused with sigmoid_focal_loss(alpha, gamma,
// 'mean') and its backward.  loss / correct accumulate (zero them first).  B <= 4096.
B200MM_API int b200mm_head_bn_focal(const void* feat, const float* W, const float* bias, const float* bn_gamma,
                                    const float* bn_beta, float* running_mean, float* running_var,
                                    const long long* labels, int B, int F, float eps, float momentum, float alpha,
                                    float gamma, int train, int bn_train, const float* dlogits_in, float* logits,
                                    float* loss, int* correct, void* dfeat, float* dW, float* dbias, float* dg,
                                    float* dbeta, void* stream) {
  if (B <= 0 || B > HEADBN_MAX_B || F <= 0) return B200MM_ERR_BAD_ARG;
  if (train && (dfeat == nullptr || dW == nullptr || dbias == nullptr || dg == nullptr || dbeta == nullptr))
    return B200MM_ERR_BAD_ARG;
  if (!bn_train && (running_mean == nullptr || running_var == nullptr)) return B200MM_ERR_BAD_ARG;
  head_bn_focal_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(feat), W, bias, bn_gamma, bn_beta, running_mean, running_var, labels, B, F,
      eps, momentum, alpha, gamma, train, bn_train, dlogits_in, logits, loss, correct,
      static_cast<__nv_bfloat16*>(dfeat), dW, dbias, dg, dbeta);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
