// Small HBM-bound kernels around the GEMMs: bias-gradient column sums, the classification head's
// output layer fused with the cross-entropy (and sigmoid-focal) loss forward+backward, the fused Adam step
// with gradient-norm clipping, and dtype casts.
//
// Replaces:
//   example_scripts/Multimodal_example_task2C.txt:195 (output_fc) + :214/:248 (nn.CrossEntropyLoss, mean)
//   example_scripts/Multimodal_example_task2C.txt:249, :217 (optim.Adam(lr=2e-5).step(); defaults
//       betas=(0.9,0.999), eps=1e-8, no weight decay)   -- SURVEY.md §2.2 K12/K13/K14
//   example_scripts/Multimodal_example_task2C.py:713-715 (clip_grad_norm_) ; :167 sigmoid_focal_loss
//       (torchvision/ops/focal_loss.py:41-54)
#include "common.cuh"
#include "device_utils.cuh"

namespace b200 {

// ------------------------------------------------------------------ column sums (bias gradients)
// out[n] += sum_m x[m, n]; x bf16 [M, N] with row stride ld.  block = (32 column groups of 8) x 8 row lanes.
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ x, long long ld, int M, int N, int rows_per_cta,
              float* __restrict__ out) {
  pdl_launch_dependents();   // see launch_pdl (common.cuh)
  pdl_wait();
  __shared__ float red[8][256 + 8];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + tx) * 8;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_cta;
  const long long r1 = min(r0 + rows_per_cta, static_cast<long long>(M));
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < N) {
    constexpr int U = 8;   // independent 16-byte loads in flight per thread (a single one left the loop latency-bound)
    for (long long r = r0 + ty; r < r1; r += 8 * U) {
      uint4 raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long rr = r + 8 * u;
        raw[u] = rr < r1 ? __ldg(reinterpret_cast<const uint4*>(x + rr * ld + col)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float v[8];
        unpack8(raw[u], v);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += v[i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[ty][tx * 8 + i] = acc[i];
  __syncthreads();
  const int c = threadIdx.x;  // 256 columns of this CTA
  const int gcol = blockIdx.x * 256 + c;
  if (gcol < N) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][c];
    atomicAdd(out + gcol, s);
  }
}

// ------------------------------------------------------------------ output layer + loss
// One warp per sample.  feat bf16 [B, F]; W fp32 [C, F]; bias fp32 [C]; C <= 8.
// loss_kind 0: softmax cross-entropy (labels int64 class ids), mean over B.
// loss_kind 1: sigmoid focal loss on logit[:,0] (alpha, gamma), labels int64 in {0,1}, mean over B; C must be 1.
// loss_kind 2: no loss here -- dL/dlogits comes from dlogits_in (fp32 [B,C]) and only the backward part runs.
// Writes logits (fp32 [B,C]); accumulates loss_sum (sum_i loss_i / B) and correct (argmax == label, or
// sigmoid > 0.5 == label); when train != 0 also dfeat (bf16 [B,F]) and accumulates dW, dbias.
constexpr int HEAD_MAX_C = 8;
__global__ void __launch_bounds__(256)
head_loss_kernel(const __nv_bfloat16* __restrict__ feat, const float* __restrict__ W, const float* __restrict__ bias,
                 const long long* __restrict__ labels, int B, int F, int C, int loss_kind, float alpha, float gamma,
                 int train, const float* __restrict__ dlogits_in, float* __restrict__ logits,
                 float* __restrict__ loss_sum, int* __restrict__ correct,
                 __nv_bfloat16* __restrict__ dfeat, float* __restrict__ dW, float* __restrict__ dbias) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= B) return;
  const __nv_bfloat16* f = feat + static_cast<long long>(row) * F;
  float z[HEAD_MAX_C];
#pragma unroll
  for (int c = 0; c < HEAD_MAX_C; ++c) z[c] = 0.f;
  for (int j = lane; j < F; j += 32) {
    const float fv = __bfloat162float(f[j]);
#pragma unroll
    for (int c = 0; c < HEAD_MAX_C; ++c)
      if (c < C) z[c] = fmaf(fv, __ldg(W + c * F + j), z[c]);
  }
#pragma unroll
  for (int c = 0; c < HEAD_MAX_C; ++c)
    if (c < C) z[c] = warp_sum(z[c]) + bias[c];
  const long long label = labels ? labels[row] : 0;
  float dz[HEAD_MAX_C];
  float loss = 0.f;
  int ok = 0;
  if (loss_kind == 2) {
    // external loss: the caller (autograd of an arbitrary criterion) supplies dL/dlogits
#pragma unroll
    for (int c = 0; c < HEAD_MAX_C; ++c) dz[c] = (c < C && dlogits_in) ? dlogits_in[row * C + c] : 0.f;
  } else if (loss_kind == 0) {
    float mx = z[0];
    int arg = 0;
#pragma unroll
    for (int c = 1; c < HEAD_MAX_C; ++c)
      if (c < C && z[c] > mx) { mx = z[c]; arg = c; }
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < HEAD_MAX_C; ++c)
      if (c < C) se += expf(z[c] - mx);
    const float lse = mx + logf(se);
#pragma unroll
    for (int c = 0; c < HEAD_MAX_C; ++c) {
      dz[c] = 0.f;
      if (c < C) {
        const float pc = expf(z[c] - lse);
        dz[c] = (pc - (c == label ? 1.f : 0.f)) / B;
        if (c == label) loss = lse - z[c];
      }
    }
    ok = (arg == label);
  } else {
    // torchvision.ops.sigmoid_focal_loss: p = sigmoid(x); ce = BCEWithLogits; p_t = p y + (1-p)(1-y);
    // loss = alpha_t * ce * (1 - p_t)^gamma
    const float x = z[0], y = static_cast<float>(label);
    const float pr = 1.f / (1.f + expf(-x));
    const float ce = fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));
    const float pt = pr * y + (1.f - pr) * (1.f - y);
    const float one_m = 1.f - pt;
    const float mod = powf(one_m, gamma);
    const float at = alpha >= 0.f ? alpha * y + (1.f - alpha) * (1.f - y) : 1.f;
    loss = at * ce * mod;
    // d/dx: dce = p - y ; dpt = (2y - 1) p (1-p) ; d(mod) = -gamma (1-pt)^(gamma-1) dpt
    const float dce = pr - y;
    const float dpt = (2.f * y - 1.f) * pr * (1.f - pr);
    const float dmod = one_m > 0.f ? -gamma * powf(one_m, gamma - 1.f) * dpt : 0.f;
#pragma unroll
    for (int c = 0; c < HEAD_MAX_C; ++c) dz[c] = 0.f;
    dz[0] = at * (dce * mod + ce * dmod) / B;
    ok = ((pr > 0.5f) == (label != 0));
  }
  if (lane == 0) {
    for (int c = 0; c < C; ++c) logits[row * C + c] = z[c];
    if (labels && loss_kind != 2) {
      atomicAdd(loss_sum, loss / B);
      if (ok) atomicAdd(correct, 1);
    }
  }
  if (train) {
    for (int j = lane; j < F; j += 32) {
      const float fv = __bfloat162float(f[j]);
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < HEAD_MAX_C; ++c)
        if (c < C) {
          d = fmaf(dz[c], __ldg(W + c * F + j), d);
          atomicAdd(dW + c * F + j, dz[c] * fv);
        }
      dfeat[static_cast<long long>(row) * F + j] = __float2bfloat16(d);
    }
    if (lane == 0)
      for (int c = 0; c < C; ++c) atomicAdd(dbias + c, dz[c]);
  }
}

// ------------------------------------------------------------------ optimizer
// sum of squares of a flat fp32 buffer -> *out (accumulate)
// four consecutive gradient elements as fp32, from an fp32 or a bf16 buffer (the data-parallel payload is bf16)
__device__ __forceinline__ float4 load_grad4(const float* g, long long i) {
  return __ldg(reinterpret_cast<const float4*>(g) + i);
}
__device__ __forceinline__ float4 load_grad4(const __nv_bfloat16* g, long long i) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(g) + i);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float grad_as_float(float x) { return x; }
__device__ __forceinline__ float grad_as_float(__nv_bfloat16 x) { return __bfloat162float(x); }

template <typename G>
__global__ void __launch_bounds__(256)
sumsq_kernel(const G* __restrict__ g, long long n, float* __restrict__ out) {
  float s = 0.f;
  const long long n4 = n >> 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = load_grad4(g, i);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 << 2; i < n; ++i) s += grad_as_float(g[i]) * grad_as_float(g[i]);
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    s = red[threadIdx.x];
    s += __shfl_xor_sync(0xffu, s, 4);
    s += __shfl_xor_sync(0xffu, s, 2);
    s += __shfl_xor_sync(0xffu, s, 1);
    if (threadIdx.x == 0) atomicAdd(out, s);
  }
}

// torch.optim.Adam semantics (no amsgrad, L2 weight decay folded into the gradient):
//   g = grad * clip ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2
//   p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
// clip = min(1, max_norm / (sqrt(*gradsq) + 1e-6)) when gradsq != nullptr (torch.nn.utils.clip_grad_norm_).
// Also refreshes the bf16 shadow copy the GEMMs read (shadow may be nullptr).
template <typename G>
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const G* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            __nv_bfloat16* __restrict__ shadow, long long n, float step_size, float inv_sqrt_bc2, float b1, float b2,
            float eps, float weight_decay, const float* __restrict__ gradsq, float max_norm, float grad_scale,
            const float* __restrict__ lr_dev, const int* __restrict__ step_dev) {
  if (step_dev != nullptr) {
    // graph-captured step: learning rate and step count live in device memory (the launch parameters are frozen at
    // capture).  bias corrections 1 - beta^t through expm1 (no cancellation at small t).
    const float t = static_cast<float>(*step_dev);
    const float bc1 = -expm1f(t * logf(b1)), bc2 = -expm1f(t * logf(b2));
    step_size = *lr_dev / bc1;
    inv_sqrt_bc2 = rsqrtf(bc2);
  }
  float clip = grad_scale;
  if (gradsq != nullptr) {
    const float norm = sqrtf(*gradsq) * grad_scale;
    clip *= fminf(1.f, max_norm / (norm + 1e-6f));
  }
  const long long n4 = n >> 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = load_grad4(g, i);
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float pa[4] = {pp.x, pp.y, pp.z, pp.w};
    const float ga[4] = {gg.x, gg.y, gg.z, gg.w};
    float ma[4] = {mm.x, mm.y, mm.z, mm.w};
    float va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gk = ga[k] * clip;
      if (weight_decay != 0.f) gk = fmaf(weight_decay, pa[k], gk);
      ma[k] = b1 * ma[k] + (1.f - b1) * gk;
      va[k] = b2 * va[k] + (1.f - b2) * gk * gk;
      pa[k] -= step_size * ma[k] / (sqrtf(va[k]) * inv_sqrt_bc2 + eps);
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(va[0], va[1], va[2], va[3]);
    if (shadow != nullptr) {
      uint2 o;
      o.x = pack_bf16x2_dev(pa[0], pa[1]);
      o.y = pack_bf16x2_dev(pa[2], pa[3]);
      reinterpret_cast<uint2*>(shadow)[i] = o;
    }
  }
}

__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n, float scale) {
  const long long n4 = n >> 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    uint2 o;
    o.x = pack_bf16x2_dev(v.x * scale, v.y * scale);
    o.y = pack_bf16x2_dev(v.z * scale, v.w * scale);
    reinterpret_cast<uint2*>(y)[i] = o;
  }
}

// rows gather: out[i, :] = x[(i * stride_rows + offset_rows), :]  (used to pull h[:, -1, :] / h[:, 0, :] out of the
// token matrix, and its backward scatter with zero fill is done by scatter_rows)
__global__ void __launch_bounds__(256)
gather_rows_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int rows, int D,
                   long long stride_rows, long long offset_rows, float p_drop, uint32_t threshold, float inv_keep,
                   unsigned long long seed) {
  const int chunks = D >> 3;
  const long long total = static_cast<long long>(rows) * chunks;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / chunks;
    const int c = static_cast<int>(i - r * chunks) * 8;
    float v[8];
    load8(x + (r * stride_rows + offset_rows) * D + c, v);
    if (p_drop > 0.f) {
      const uint32_t keep = dropout_keep8(seed, static_cast<uint64_t>(r * D + c) >> 3, threshold);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = (keep >> k) & 1 ? v[k] * inv_keep : 0.f;
    }
    store8(out + r * D + c, v);
  }
}
// dx[M, D] = 0 except rows (i * stride_rows + offset_rows) = dropout-masked dpooled[i]
__global__ void __launch_bounds__(256)
scatter_rows_kernel(const __nv_bfloat16* __restrict__ dpooled, __nv_bfloat16* __restrict__ dx, long long M, int D,
                    long long stride_rows, long long offset_rows, float p_drop, uint32_t threshold, float inv_keep,
                    unsigned long long seed) {
  const int chunks = D >> 3;
  const long long total = M * chunks;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / chunks;
    const int c = static_cast<int>(i - row * chunks) * 8;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const long long rel = row - offset_rows;
    if (rel >= 0 && rel % stride_rows == 0) {
      const long long r = rel / stride_rows;
      load8(dpooled + r * D + c, v);
      if (p_drop > 0.f) {
        const uint32_t keep = dropout_keep8(seed, static_cast<uint64_t>(r * D + c) >> 3, threshold);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = (keep >> k) & 1 ? v[k] * inv_keep : 0.f;
      }
    }
    store8(dx + row * D + c, v);
  }
}

// last node of a captured train step: the next replay sees a new dropout salt and Adam step count
__global__ void step_advance_kernel(unsigned long long* salt, int* adam_step) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (salt != nullptr) *salt += 0x9E3779B97F4A7C15ull;
    if (adam_step != nullptr) *adam_step += 1;
  }
}

static int grid_for(long long work_items, int threads) {
  const DeviceInfo& dev = device_info();
  const long long max_ctas = static_cast<long long>(dev.num_sms > 0 ? dev.num_sms : 148) * 8;
  const long long want = ceil_div(work_items, static_cast<long long>(threads));
  return static_cast<int>(want < 1 ? 1 : (want > max_ctas ? max_ctas : want));
}

}  // namespace b200

using namespace b200;

// out[n] += sum_m x[m,n]   (x bf16 [M,N], row stride ld; out fp32 [N], accumulate)
B200MM_API int b200mm_colsum_bf16(const void* x, long long ld, int M, int N, float* out, void* stream) {
  if (M <= 0 || N <= 0 || (N & 7) || (ld & 7)) return B200MM_ERR_BAD_ARG;
  const DeviceInfo& dev = device_info();
  const int gx = ceil_div(N, 256);
  int gy = ceil_div((dev.num_sms > 0 ? dev.num_sms : 148) * 4, gx);
  int rows_per_cta = ceil_div(M, gy);
  rows_per_cta = ceil_div(rows_per_cta, 8) * 8;
  gy = ceil_div(M, rows_per_cta);
  launch_pdl(colsum_kernel, dim3(gx, gy), dim3(256), 0, static_cast<cudaStream_t>(stream),
             static_cast<const __nv_bfloat16*>(x), ld, M, N, rows_per_cta, out);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

B200MM_API int b200mm_head_loss(const void* feat, const float* W, const float* bias, const long long* labels, int B,
                                int F, int C, int loss_kind, float alpha, float gamma, int train,
                                const float* dlogits_in, float* logits, float* loss_sum, int* correct, void* dfeat,
                                float* dW, float* dbias, void* stream) {
  if (B <= 0 || F <= 0 || C <= 0 || C > HEAD_MAX_C || (loss_kind == 1 && C != 1) || loss_kind < 0 || loss_kind > 2)
    return B200MM_ERR_BAD_ARG;
  if (train && (!dfeat || !dW || !dbias)) return B200MM_ERR_BAD_ARG;
  if (train && loss_kind != 2 && !labels) return B200MM_ERR_BAD_ARG;
  if (train && loss_kind == 2 && !dlogits_in) return B200MM_ERR_BAD_ARG;
  head_loss_kernel<<<ceil_div(B, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(feat), W, bias, labels, B, F, C, loss_kind, alpha, gamma, train, dlogits_in,
      logits, loss_sum, correct, static_cast<__nv_bfloat16*>(dfeat), dW, dbias);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

B200MM_API int b200mm_sumsq_f32(const float* g, long long n, float* out, void* stream) {
  if (n <= 0) return B200MM_ERR_BAD_ARG;
  sumsq_kernel<float><<<grid_for(n >> 2, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(g, n, out);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
// same over a bf16 buffer (the all-reduced data-parallel gradient payload)
B200MM_API int b200mm_sumsq_bf16(const void* g, long long n, float* out, void* stream) {
  if (n <= 0) return B200MM_ERR_BAD_ARG;
  sumsq_kernel<__nv_bfloat16><<<grid_for(n >> 2, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(g), n, out);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

// One fused Adam step over a flat fp32 parameter segment (n % 4 == 0, 16-byte aligned).
B200MM_API int b200mm_adam_step(float* p, const float* g, float* m, float* v, void* shadow_bf16, long long n, float lr,
                                float beta1, float beta2, float eps, float weight_decay, int step,
                                const float* gradsq, float max_norm, float grad_scale, void* stream) {
  if (n <= 0 || (n & 3) || step < 1) return B200MM_ERR_BAD_ARG;
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  adam_kernel<float><<<grid_for(n >> 2, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, static_cast<__nv_bfloat16*>(shadow_bf16), n, static_cast<float>(lr / bc1),
      static_cast<float>(1.0 / sqrt(bc2)), beta1, beta2, eps, weight_decay, gradsq, max_norm, grad_scale, nullptr,
      nullptr);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
// The same step reading the gradient from a bf16 buffer: under data parallelism the all-reduced payload is bf16
// (half the NVLink bytes) while parameters and both moments stay fp32 -- 28 B / parameter instead of 30.
B200MM_API int b200mm_adam_step_g16(float* p, const void* g_bf16, float* m, float* v, void* shadow_bf16, long long n,
                                    float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                                    const float* gradsq, float max_norm, float grad_scale, void* stream) {
  if (n <= 0 || (n & 3) || step < 1) return B200MM_ERR_BAD_ARG;
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  adam_kernel<__nv_bfloat16><<<grid_for(n >> 2, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, static_cast<const __nv_bfloat16*>(g_bf16), m, v, static_cast<__nv_bfloat16*>(shadow_bf16), n,
      static_cast<float>(lr / bc1), static_cast<float>(1.0 / sqrt(bc2)), beta1, beta2, eps, weight_decay, gradsq,
      max_norm, grad_scale, nullptr, nullptr);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

// Capturable Adam step: learning rate (*lr_dev) and step count (*step_dev, >= 1) are read on the device, so ONE captured
// launch serves every replay of a CUDA graph (the scheduler rewrites *lr_dev, b200mm_step_advance increments *step_dev).
// g_is_bf16: the gradient buffer is the bf16 data-parallel payload.
B200MM_API int b200mm_adam_step_dyn(float* p, const void* g, int g_is_bf16, float* m, float* v, void* shadow_bf16,
                                    long long n, const float* lr_dev, float beta1, float beta2, float eps,
                                    float weight_decay, const int* step_dev, const float* gradsq, float max_norm,
                                    float grad_scale, void* stream) {
  if (n <= 0 || (n & 3) || !lr_dev || !step_dev) return B200MM_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (g_is_bf16)
    adam_kernel<__nv_bfloat16><<<grid_for(n >> 2, 256), 256, 0, st>>>(
        p, static_cast<const __nv_bfloat16*>(g), m, v, static_cast<__nv_bfloat16*>(shadow_bf16), n, 0.f, 1.f, beta1,
        beta2, eps, weight_decay, gradsq, max_norm, grad_scale, lr_dev, step_dev);
  else
    adam_kernel<float><<<grid_for(n >> 2, 256), 256, 0, st>>>(
        p, static_cast<const float*>(g), m, v, static_cast<__nv_bfloat16*>(shadow_bf16), n, 0.f, 1.f, beta1, beta2, eps,
        weight_decay, gradsq, max_norm, grad_scale, lr_dev, step_dev);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

// salt += odd constant, adam_step += 1 (either may be nullptr): the last node of a captured step.
B200MM_API int b200mm_step_advance(unsigned long long* salt, int* adam_step, void* stream) {
  step_advance_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(salt, adam_step);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

B200MM_API int b200mm_cast_f32_to_bf16(const float* x, void* y, long long n, void* stream) {
  if (n <= 0 || (n & 3)) return B200MM_ERR_BAD_ARG;
  cast_f32_bf16_kernel<<<grid_for(n >> 2, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<__nv_bfloat16*>(y), n, 1.f);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
// y = bf16(x * scale): packs a gradient range for the data-parallel all-reduce (scale = 1 / world size, applied
// BEFORE the sum so the bf16 partial sums stay in range)
B200MM_API int b200mm_scale_cast_f32_to_bf16(const float* x, void* y, long long n, float scale, void* stream) {
  if (n <= 0 || (n & 3)) return B200MM_ERR_BAD_ARG;
  cast_f32_bf16_kernel<<<grid_for(n >> 2, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<__nv_bfloat16*>(y), n, scale);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

// out[i,:] = dropout(x[i*stride_rows + offset_rows, :])  -- pooled token (h[:, -1] for the baseline, h[:, 0] CLS)
B200MM_API int b200mm_gather_rows(const void* x, void* out, int rows, int D, long long stride_rows,
                                  long long offset_rows, float p_drop, unsigned long long seed, void* stream) {
  if (rows <= 0 || D <= 0 || (D & 7) || stride_rows <= 0 || p_drop < 0.f || p_drop >= 1.f) return B200MM_ERR_BAD_ARG;
  gather_rows_kernel<<<grid_for(static_cast<long long>(rows) * (D >> 3), 256), 256, 0,
                       static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), rows, D, stride_rows, offset_rows,
      p_drop, dropout_threshold(p_drop), 1.f / (1.f - p_drop), seed);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
B200MM_API int b200mm_scatter_rows(const void* dpooled, void* dx, long long M, int D, long long stride_rows,
                                   long long offset_rows, float p_drop, unsigned long long seed, void* stream) {
  if (M <= 0 || D <= 0 || (D & 7) || stride_rows <= 0 || p_drop < 0.f || p_drop >= 1.f) return B200MM_ERR_BAD_ARG;
  scatter_rows_kernel<<<grid_for(M * (D >> 3), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dpooled), static_cast<__nv_bfloat16*>(dx), M, D, stride_rows, offset_rows,
      p_drop, dropout_threshold(p_drop), 1.f / (1.f - p_drop), seed);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
