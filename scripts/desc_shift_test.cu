// Experiment: does a SWIZZLE_128B UMMA descriptor whose start address is shifted by whole 128-byte rows (not a multiple
// of 1024 B) read the rows TMA-style (absolute-address swizzle) data correctly with base_offset = 0?
#include "../multimodal-propaganda-meme-classification_b200/csrc/ptx.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace b200;

__global__ void __launch_bounds__(128) k(int shift, int mode, int boff, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;                 // 384 rows x 128 B, value X[r][c] = (r % 61) - 30 + c/64.0 (exact in bf16? use ints)
  uint8_t* sB = smem + 384 * 128;     // 64 x 64 identity (symmetric: K-major == MN-major)
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 384 * 64; i += 128) {
    const int r = i >> 6, c = i & 63;
    const float v = static_cast<float>(((r * 7 + c * 3) % 127) - 63);
    const int byte = r * 128 + ((((c >> 3) ^ (r & 7)) << 4)) + (c & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sX + byte) = __float2bfloat16(v);
  }
  for (int i = threadIdx.x; i < 64 * 64; i += 128) {
    const int r = i >> 6, c = i & 63;
    const int byte = r * 128 + ((((c >> 3) ^ (r & 7)) << 4)) + (c & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sB + byte) = __float2bfloat16(r == c ? 1.f : 0.f);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<64>(&slot);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t a0 = smem_u32(sX) + shift * 128, b0 = smem_u32(sB);
    const uint64_t bo = static_cast<uint64_t>(boff & 7) << 49;
    if (mode == 0) {   // K-major A: D[r][n] = X[shift + r][n]
      const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
      for (int kk = 0; kk < 4; ++kk)
        umma_bf16(tmem, umma_desc_sw128(a0 + kk * 32, 16, 1024) | bo, umma_desc_sw128(b0 + kk * 32, 16, 1024), idesc, kk ? 1u : 0u);
    } else {           // MN-major A (M = channel, K = row): atom 1 = one row further (LBO = 128 B); B MN-major identity
      const uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
      for (int kk = 0; kk < 4; ++kk)
        umma_bf16(tmem, umma_desc_sw128(a0 + kk * 2048, 128, 1024) | bo, umma_desc_sw128(b0 + kk * 2048, 8192, 1024), idesc, kk ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int h = 0; h < 2; ++h) {
    uint32_t v[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + h * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + h * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<64>(tmem);
}

static float X(int r, int c) { return static_cast<float>(((r * 7 + c * 3) % 127) - 63); }

int main() {
  float* d;
  cudaMalloc(&d, 128 * 64 * 4);
  const int smem = 384 * 128 + 8192 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> h(128 * 64);
  const int shifts[] = {0, 1, 3, 8, 58, 59, 117};
  for (int mode = 0; mode < 2; ++mode)
    for (int s : shifts)
      for (int bo = 0; bo < 2; ++bo) {
        const int boff = bo ? (s & 7) : 0;
        if (bo && boff == 0) continue;
        k<<<1, 128, smem>>>(s, mode, boff, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d shift %d: %s\n", mode, s, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 64; ++n) {
            float ref;
            if (mode == 0) ref = X(s + m, n);
            else ref = m < 64 ? X(s + n, m) : X(s + 1 + n, m - 64);
            if (h[m * 64 + n] != ref) ++bad;
          }
        printf("mode %d shift %3d base_offset %d: %s (%d mismatches)\n", mode, s, boff, bad ? "WRONG" : "ok", bad);
      }
  return 0;
}
