// TEST INFRASTRUCTURE: the handful of CUDA names a simple one-thread-per-element kernel uses, defined for a plain C++
// compiler, so that the KERNEL BODIES of preprocess_pil.cu / augment_pil.cu (indexing included, not only their arithmetic
// headers) can be executed thread by thread on the host by tests/test_cpu.py.  Only kernels whose threads do not exchange
// data can run this way (integer atomicAdd is fine: sequential addition gives the same sum); the shuffle below is a dummy.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

struct EmuDim3 { unsigned x = 1, y = 1, z = 1; };
static EmuDim3 blockIdx, threadIdx, blockDim, gridDim;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __ldg(p) (*(p))
#define __shared__ static
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }                 // not executed
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p += v; return o; }

// runs `kernel(args...)` for every thread of a (gx, gy, gz) grid of `threads` threads, one after the other
template <typename K, typename... A>
static void emu_launch(K kernel, unsigned gx, unsigned gy, unsigned gz, unsigned threads, A... args) {
  gridDim.x = gx; gridDim.y = gy; gridDim.z = gz;
  blockDim.x = threads;
  for (unsigned z = 0; z < gz; ++z)
    for (unsigned y = 0; y < gy; ++y)
      for (unsigned x = 0; x < gx; ++x)
        for (unsigned t = 0; t < threads; ++t) {
          blockIdx.x = x; blockIdx.y = y; blockIdx.z = z;
          threadIdx.x = t;
          kernel(args...);
        }
}
