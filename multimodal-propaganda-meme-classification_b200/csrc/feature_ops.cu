// Kernels of the feature-extraction path that the classifier step does not need: the depthwise
// 7x7 convolution of ConvNeXt blocks and the tanh of BERT's pooler.
//
// Replaces: torchvision/models/convnext.py CNBlock's nn.Conv2d(dim, dim, kernel_size=7, padding=3, groups=dim) and
// transformers/models/bert/modeling_bert.py BertPooler's nn.Tanh, as called by baselines/extract_feat.py:52-67
// (img_model.avgpool(img_model.features(images)), text_model(text_tokens).pooler_output).
#include "common.cuh"
#include "device_utils.cuh"
#include "ptx.cuh"

namespace b200 {

// y[n, h, w, c] = bias[c] + sum_{kh, kw} wt[kh * 7 + kw][c] * x[n, h + kh - 3, w + kw - 3, c]   (zero padding), NHWC bf16.
// A thread owns 8 channels (one 16-byte vector) x 4 consecutive output columns: the 10 input vectors of a kernel row
// are loaded once and slide under the 7 taps, so an output costs 17.5 loads instead of 49; fp32 accumulation in packed
// FFMA2.  HBM traffic is |x| + |y| (the 7-row window of a thread column stays in L1 / L2).
constexpr int DW_OUT = 4;
__global__ void __launch_bounds__(256)
dwconv7x7_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ wt,
                 const float* __restrict__ bias, __nv_bfloat16* __restrict__ y, int N, int H, int W, int C) {
  const int G = C >> 3;
  const int WB = ceil_div(W, DW_OUT);
  const long long total = static_cast<long long>(N) * H * WB * G;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % G);
    long long r = idx / G;
    const int wb = static_cast<int>(r % WB);
    r /= WB;
    const int h = static_cast<int>(r % H);
    const int n = static_cast<int>(r / H);
    const int w0 = wb * DW_OUT, c0 = g * 8;
    float2 acc[DW_OUT][4];
    {
      const float4 b0 = bias ? __ldg(reinterpret_cast<const float4*>(bias + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 b1 = bias ? __ldg(reinterpret_cast<const float4*>(bias + c0 + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int o = 0; o < DW_OUT; ++o) {
        acc[o][0] = make_float2(b0.x, b0.y); acc[o][1] = make_float2(b0.z, b0.w);
        acc[o][2] = make_float2(b1.x, b1.y); acc[o][3] = make_float2(b1.z, b1.w);
      }
    }
#pragma unroll 1
    for (int kh = 0; kh < 7; ++kh) {
      const int ih = h + kh - 3;
      if (ih < 0 || ih >= H) continue;
      const __nv_bfloat16* xrow = x + (static_cast<long long>(n) * H + ih) * W * C + c0;
      uint4 xin[DW_OUT + 6];
#pragma unroll
      for (int j = 0; j < DW_OUT + 6; ++j) {
        const int iw = w0 + j - 3;
        xin[j] = (iw >= 0 && iw < W) ? __ldg(reinterpret_cast<const uint4*>(xrow + static_cast<long long>(iw) * C))
                                     : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int kw = 0; kw < 7; ++kw) {
        const uint4 wv = __ldg(reinterpret_cast<const uint4*>(wt + static_cast<long long>(kh * 7 + kw) * C + c0));
        const float2 w0f = unpack_bf16x2(wv.x), w1f = unpack_bf16x2(wv.y), w2f = unpack_bf16x2(wv.z),
                     w3f = unpack_bf16x2(wv.w);
#pragma unroll
        for (int o = 0; o < DW_OUT; ++o) {
          const uint4 xv = xin[o + kw];
          acc[o][0] = ffma2(unpack_bf16x2(xv.x), w0f, acc[o][0]);
          acc[o][1] = ffma2(unpack_bf16x2(xv.y), w1f, acc[o][1]);
          acc[o][2] = ffma2(unpack_bf16x2(xv.z), w2f, acc[o][2]);
          acc[o][3] = ffma2(unpack_bf16x2(xv.w), w3f, acc[o][3]);
        }
      }
    }
    __nv_bfloat16* yrow = y + ((static_cast<long long>(n) * H + h) * W + w0) * C + c0;
#pragma unroll
    for (int o = 0; o < DW_OUT; ++o) {
      if (w0 + o < W) {
        uint4 ov;
        ov.x = pack_bf16x2(acc[o][0].x, acc[o][0].y);
        ov.y = pack_bf16x2(acc[o][1].x, acc[o][1].y);
        ov.z = pack_bf16x2(acc[o][2].x, acc[o][2].y);
        ov.w = pack_bf16x2(acc[o][3].x, acc[o][3].y);
        *reinterpret_cast<uint4*>(yrow + static_cast<long long>(o) * C) = ov;
      }
    }
  }
}

__global__ void __launch_bounds__(256)
tanh_f32_kernel(float* __restrict__ x, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    x[i] = tanhf(x[i]);
}

}  // namespace b200

using namespace b200;

// Depthwise 7x7 / stride 1 / pad 3 convolution, NHWC bf16; wt = [49][C] bf16 (tap-major), bias fp32 [C] or nullptr.
B200MM_API int b200mm_dwconv7x7_nhwc(const void* x, const void* wt, const float* bias, void* y, int N, int H, int W,
                                     int C, void* stream) {
  if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 7) || !x || !wt || !y) return B200MM_ERR_BAD_ARG;
  const long long total = static_cast<long long>(N) * H * ceil_div(W, DW_OUT) * (C >> 3);
  const DeviceInfo& dev = device_info();
  const long long max_ctas = static_cast<long long>(dev.num_sms > 0 ? dev.num_sms : 148) * 8;
  const long long want = ceil_div(total, 256LL);
  const int grid = static_cast<int>(want > max_ctas ? max_ctas : (want < 1 ? 1 : want));
  dwconv7x7_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(wt), bias,
      static_cast<__nv_bfloat16*>(y), N, H, W, C);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}

// x[i] = tanh(x[i]) in place, fp32 (BertPooler activation; exact tanhf: the pooled vector is the product here).
B200MM_API int b200mm_tanh_f32(float* x, long long n, void* stream) {
  if (n <= 0 || !x) return B200MM_ERR_BAD_ARG;
  const long long want = ceil_div(n, 256LL);
  tanh_f32_kernel<<<static_cast<int>(want > 1184 ? 1184 : want), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n);
  B200MM_CHECK_LAUNCH();
  return B200MM_OK;
}
