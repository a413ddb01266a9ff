// TEST INFRASTRUCTURE (never loaded by the product): csrc/jpeg_math.cuh -- the integer arithmetic of jpeg_decode.cu's
// kernels -- compiled for the host, so that tests/test_cpu.py can demand pixel equality with Pillow's decode in a
// container without a GPU.  Input: the coefficients / tables the product's own host entry point
// (b200mm_jpeg_entropy_decode) produced.    g++ -O2 -shared -fPIC -I <pkg>/csrc tests/host/host_jpeg.cpp
#include <vector>

#include "jpeg_math.cuh"

using namespace b200;

// info: the int[32] of b200mm_jpeg_parse; out: [height][width][3] uint8
extern "C" void host_jpeg_reconstruct(const short* coefs, const unsigned short* qtabs, const int* info,
                                      unsigned char* out) {
  const int W = info[0], H = info[1], ncomp = info[2], hs = info[4], vs = info[5];
  std::vector<std::vector<uint8_t>> plane(ncomp);
  for (int c = 0; c < ncomp; ++c) {
    const int wb = info[6 + c], hb = info[9 + c], stride = wb * 8;
    plane[c].assign(static_cast<size_t>(stride) * hb * 8, 0);
    for (int by = 0; by < hb; ++by)
      for (int bx = 0; bx < wb; ++bx)
        jpeg::idct_islow_block(coefs + info[18 + c] + (static_cast<long long>(by) * wb + bx) * 64, qtabs + 64 * c,
                               plane[c].data() + static_cast<size_t>(by) * 8 * stride + bx * 8, stride);
  }
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      unsigned char* px = out + (static_cast<long long>(y) * W + x) * 3;
      const int yy = plane[0][static_cast<size_t>(y) * info[6] * 8 + x];
      if (ncomp == 1) {
        px[0] = px[1] = px[2] = static_cast<unsigned char>(yy);
        continue;
      }
      const int cb = jpeg::upsampled_sample(plane[1].data(), info[7] * 8, info[13], info[16], hs, vs, x, y);
      const int cr = jpeg::upsampled_sample(plane[2].data(), info[8] * 8, info[14], info[17], hs, vs, x, y);
      jpeg::ycc_to_rgb(yy, cb, cr, px);
    }
}
