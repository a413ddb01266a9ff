// Device-side helpers shared by the non-GEMM kernels: warp/block reductions, Philox counter RNG,
// 16-byte vector <-> bf16x8 conversion.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace b200 {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Per-step salt of every dropout seed, read from DEVICE memory.  The host derives one seed per dropout site from
// (model seed, step, layer, site) and passes it by value; a CUDA graph freezes those values at capture, so a graphed
// train step additionally adds *c_step_salt -- a device counter its last node advances -- and every replay draws fresh
// masks (forward and backward of one replay read the same value).  nullptr (eager mode): seeds are used as passed.
// One copy of the pointer per translation unit (no relocatable device code in this build): each unit registers a
// setter, b200mm_set_step_salt_ptr (common.cu) calls them all.
static __constant__ const unsigned long long* c_step_salt = nullptr;
void register_step_salt_setter(int (*setter)(const unsigned long long*));
static int set_step_salt_this_unit(const unsigned long long* ptr) {
  return static_cast<int>(cudaMemcpyToSymbol(c_step_salt, &ptr, sizeof(ptr)));
}
struct StepSaltRegistration {
  explicit StepSaltRegistration(int (*setter)(const unsigned long long*)) { register_step_salt_setter(setter); }
};
static StepSaltRegistration step_salt_registration(&set_step_salt_this_unit);
__device__ __forceinline__ uint64_t step_seed(uint64_t seed) {
  const unsigned long long* sp = c_step_salt;
  return sp != nullptr ? seed + __ldg(sp) : seed;
}

// Philox4x32-10: stateless counter RNG so that forward and backward (and every data-parallel replica,
// given its own seed) regenerate identical dropout masks from (seed, element index).
__device__ __forceinline__ uint4 philox4x32(uint64_t seed, uint64_t ctr) {
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  uint32_t c0 = static_cast<uint32_t>(ctr), c1 = static_cast<uint32_t>(ctr >> 32), c2 = 0x1BD11BDAu, c3 = 0x5851F42Du;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// Same generator with a compile-time round count (7 rounds already pass BigCrush, Salmon et al. 2011; the attention
// kernels draw 128 x 128 Bernoulli variables per head and were spending more instructions on the generator than on the
// softmax) and 16-bit fields: one call yields 8 keep decisions.
template <int ROUNDS>
__device__ __forceinline__ uint4 philox4x32_r(uint64_t seed, uint64_t ctr) {
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  uint32_t c0 = static_cast<uint32_t>(ctr), c1 = static_cast<uint32_t>(ctr >> 32), c2 = 0x1BD11BDAu, c3 = 0x5851F42Du;
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// 32 keep bits (bit i set = element i kept) for one 32-element chunk; thr16 = p_drop * 2^16.
__device__ __forceinline__ uint32_t dropout_keep32(uint64_t seed, uint64_t chunk_idx, uint32_t thr16) {
  seed = step_seed(seed);
  uint32_t bits = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 r = philox4x32_r<7>(seed, chunk_idx * 4 + q);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      bits |= ((w[i] & 0xFFFFu) >= thr16 ? 1u : 0u) << (q * 8 + 2 * i);
      bits |= ((w[i] >> 16) >= thr16 ? 1u : 0u) << (q * 8 + 2 * i + 1);
    }
  }
  return bits;
}
// keep-mask for 8 consecutive elements (bit i set = element i kept) from ONE Philox4x32-7 call (16-bit fields);
// group8_idx = flat element index / 8, threshold = p_drop * 2^32 (the upper 16 bits are compared).  Every elementwise
// dropout of the engine (GEMM epilogue, LayerNorm / embedding kernels, pooled-feature gather / scatter) draws its mask
// here, forward and backward alike -- two 10-round calls per 8 elements made the LayerNorm backward ALU-bound.
__device__ __forceinline__ uint32_t dropout_keep8(uint64_t seed, uint64_t group8_idx, uint32_t threshold) {
  const uint32_t thr16 = threshold >> 16;
  const uint4 r = philox4x32_r<7>(step_seed(seed), group8_idx);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t bits = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    bits |= ((w[i] & 0xFFFFu) >= thr16 ? 1u : 0u) << (2 * i);
    bits |= ((w[i] >> 16) >= thr16 ? 1u : 0u) << (2 * i + 1);
  }
  return bits;
}
// keep-mask for 4 consecutive elements: bit i set = element kept. threshold = p_drop * 2^32.
__device__ __forceinline__ uint32_t dropout_keep4(uint64_t seed, uint64_t group_idx, uint32_t threshold) {
  const uint4 r = philox4x32(step_seed(seed), group_idx);
  return (r.x >= threshold ? 1u : 0u) | (r.y >= threshold ? 2u : 0u) | (r.z >= threshold ? 4u : 0u) |
         (r.w >= threshold ? 8u : 0u);
}
__host__ __device__ inline uint32_t dropout_threshold(float p) {
  const double t = static_cast<double>(p) * 4294967296.0;
  return t >= 4294967295.0 ? 0xFFFFFFFFu : static_cast<uint32_t>(t);
}

__device__ __forceinline__ float2 unpack_bf16x2_dev(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ uint32_t pack_bf16x2_dev(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

struct alignas(16) bf16x8 {
  uint32_t u[4];
};
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&x)[8]) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&u[i]);
    const float2 f = __bfloat1622float2(v);
    x[2 * i] = f.x;
    x[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void unpack8(const uint4& r, float (&x)[8]) {
  const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&u[i]);
    const float2 f = __bfloat1622float2(v);
    x[2 * i] = f.x;
    x[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&x)[8]) {
  uint4 o;
  uint32_t* u = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 v = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
    u[i] = *reinterpret_cast<uint32_t*>(&v);
  }
  *reinterpret_cast<uint4*>(p) = o;
}

}  // namespace b200
