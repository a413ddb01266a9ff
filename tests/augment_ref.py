"""torchvision's float-tensor ColorJitter / RandomRotation / Normalize applied with GIVEN per-image draws: the checker of
csrc/augment.cu (GPU tests) and of csrc/augment_math.cuh compiled for the host (CPU test).  The reference transform is
example_scripts/Multimodal_example_task2C.py:224-233."""
import os
import subprocess

import torch
import torchvision.transforms.functional as TF
from torchvision.transforms import InterpolationMode

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def torchvision_augment(img01, perm, factors, angles, mean, std):
    """img01 [n, 3, H, W] in [0, 1] (any device) -> what ColorJitter.forward (transforms.py:1276-1290) followed by
    RandomRotation.forward (NEAREST, expand=False, fill=0) and Normalize produce for these draws."""
    out = torch.empty_like(img01)
    for i in range(img01.shape[0]):
        b, c, s, h = (float(v) for v in factors[i])
        x = img01[i]
        for fn in perm[i].tolist():
            if fn == 0:
                x = TF.adjust_brightness(x, b)
            elif fn == 1:
                x = TF.adjust_contrast(x, c)
            elif fn == 2:
                x = TF.adjust_saturation(x, s)
            else:
                x = TF.adjust_hue(x, h)
        x = TF.rotate(x, float(angles[i]), InterpolationMode.NEAREST, False, None, [0.0, 0.0, 0.0])
        out[i] = TF.normalize(x, list(mean), list(std))
    return out


def build_host_harness(out_dir):
    """g++ build of tests/host/host_augment.cpp (the kernels' per-pixel header compiled for the CPU)."""
    lib = os.path.join(str(out_dir), "libhost_augment.so")
    csrc = os.path.join(ROOT, "multimodal-propaganda-meme-classification_b200", "csrc")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", csrc,
                    os.path.join(ROOT, "tests", "host", "host_augment.cpp"), "-o", lib], check=True)
    return lib


def build_host_jpeg_harness(out_dir):
    """g++ build of tests/host/host_jpeg.cpp (csrc/jpeg_math.cuh -- the kernels' integer arithmetic -- for the CPU)."""
    lib = os.path.join(str(out_dir), "libhost_jpeg.so")
    csrc = os.path.join(ROOT, "multimodal-propaganda-meme-classification_b200", "csrc")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-I", csrc, os.path.join(ROOT, "tests", "host", "host_jpeg.cpp"),
                    "-o", lib], check=True)
    return lib


def jpeg_cases(seed=0):
    """(description, file bytes) over the supported set: baseline / progressive, 4:4:4 / 4:2:2 / 4:2:0 / grayscale,
    optimised tables, restart intervals, sizes that are not multiples of the MCU, one-pixel images, extreme qualities."""
    import io

    import numpy as np
    import torch.nn.functional as F
    from PIL import Image
    rng = np.random.default_rng(seed)

    def photo(h, w):
        base = torch.from_numpy(rng.random((1, 3, 9, 11), dtype=np.float32))
        im = F.interpolate(base, size=(h, w), mode="bicubic", align_corners=False)[0]
        im = im + 0.08 * torch.from_numpy(rng.standard_normal((3, h, w)).astype(np.float32))
        return (im.clamp(0, 1) * 255).byte().permute(1, 2, 0).numpy()

    def enc(arr, gray=False, **kw):
        b = io.BytesIO()
        (Image.fromarray(arr).convert("L") if gray else Image.fromarray(arr)).save(b, "JPEG", **kw)
        return b.getvalue()

    cases = []
    for (h, w) in [(64, 64), (37, 53), (1, 1), (8, 17), (200, 301), (2, 3), (33, 4), (5, 3)]:
        for sub in (0, 1, 2):
            for prog in (False, True):
                q = (30, 75, 95)[(h + sub + prog) % 3]
                cases.append((f"{h}x{w} sub{sub} prog{prog} q{q}", enc(photo(h, w), quality=q, subsampling=sub,
                                                                       progressive=prog, optimize=bool((h + sub) & 1))))
        cases.append((f"{h}x{w} gray", enc(photo(h, w), gray=True, quality=80)))
        cases.append((f"{h}x{w} gray prog", enc(photo(h, w), gray=True, quality=60, progressive=True)))
    noise = rng.integers(0, 256, (123, 211, 3), dtype=np.uint8)
    for kw in (dict(quality=100, subsampling=0), dict(quality=1, subsampling=2),
               dict(quality=100, subsampling=2, progressive=True), dict(quality=90, subsampling=2, restart_marker_blocks=3),
               dict(quality=90, subsampling=1, restart_marker_rows=1),
               dict(quality=85, subsampling=2, progressive=True, restart_marker_rows=2),
               dict(quality=50, subsampling=0, restart_marker_blocks=1)):
        cases.append((f"noise {kw}", enc(noise, **kw)))
        cases.append((f"photo {kw}", enc(photo(240, 320), **kw)))
    return cases


def build_host_resample_harness(out_dir):
    """g++ build of tests/host/host_resample.cpp (csrc/resample_math.cuh -- Pillow's 8-bit resize restated -- for the CPU)."""
    lib = os.path.join(str(out_dir), "libhost_resample.so")
    csrc = os.path.join(ROOT, "multimodal-propaganda-meme-classification_b200", "csrc")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", csrc,
                    os.path.join(ROOT, "tests", "host", "host_resample.cpp"), "-o", lib], check=True)
    return lib


def build_host_augment_pil_harness(out_dir):
    """g++ build of tests/host/host_augment_pil.cpp (csrc/augment_pil_math.cuh -- Pillow's ColorJitter / rotate arithmetic)."""
    lib = os.path.join(str(out_dir), "libhost_augment_pil.so")
    csrc = os.path.join(ROOT, "multimodal-propaganda-meme-classification_b200", "csrc")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", csrc,
                    os.path.join(ROOT, "tests", "host", "host_augment_pil.cpp"), "-o", lib], check=True)
    return lib


def build_emulated_pil_kernels(out_dir):
    """The kernel functions of preprocess_pil.cu and augment_pil.cu (text between their `namespace b200 {` and the C-ABI
    entry points) compiled for the host behind tests/host/cuda_emulation_shim.h, plus two drivers that run them thread by
    thread over the same grids the entry points launch."""
    csrc = os.path.join(ROOT, "multimodal-propaganda-meme-classification_b200", "csrc")

    def kernels(name):
        text = open(os.path.join(csrc, name)).read()
        return text[text.index("namespace b200 {"):text.index("using namespace b200;")]

    src = os.path.join(str(out_dir), "emulated_pil_kernels.cpp")
    with open(src, "w") as f:
        f.write('#include "cuda_emulation_shim.h"\n#include "resample_math.cuh"\n#include "augment_pil_math.cuh"\n')
        f.write(kernels("preprocess_pil.cu"))
        f.write(kernels("augment_pil.cu"))
        f.write(r'''
using namespace b200;
static unsigned cdiv(int a, int b) { return static_cast<unsigned>((a + b - 1) / b); }
extern "C" void emu_preprocess_pil(const unsigned char* packed, const long long* offsets, const int* heights,
                                   const int* widths, const unsigned char* flip, int n, int resize, int crop, int square,
                                   const float* mean3, const float* std3, float* out) {
  PilPreprocParams p{};
  p.packed = packed; p.offsets = offsets; p.flip = flip; p.heights = heights; p.widths = widths;
  p.n = n; p.resize = resize; p.crop = crop; p.square = square; p.out = out;
  for (int c = 0; c < 3; ++c) { p.mean[c] = mean3[c]; p.std[c] = std3[c]; }
  emu_launch(preprocess_pil_kernel, cdiv(crop, 32), cdiv(crop, 8), n, 256, p);
}
extern "C" void emu_train_transform_pil(const unsigned char* packed, const long long* offsets, const int* heights,
                                        const int* widths, const unsigned char* flip, int n, int resize, int crop,
                                        const int* order, const float* alpha, const int* hue, const int* affine,
                                        const float* mean3, const float* std3, unsigned char* u8, float* out) {
  emu_launch(preprocess_pil_u8_kernel, cdiv(crop, 32), cdiv(crop, 8), n, 256, packed, offsets, heights, widths, flip, resize,
             crop, 1, u8);
  unsigned long long* sums = new unsigned long long[n]();
  emu_launch(pil_luma_sum_kernel, kLumaParts, n, 1, kLumaThreads, u8, order, alpha, hue, crop, crop, sums);
  emu_launch(pil_jitter_rotate_kernel, cdiv(crop, 32), cdiv(crop, 8), n, 256, u8, order, alpha, hue, affine, sums, crop, crop,
             mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2], out);
  delete[] sums;
}
''')
    lib = os.path.join(str(out_dir), "libemulated_pil_kernels.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "tests", "host"),
                    "-I", csrc, src, "-o", lib], check=True)
    return lib


def build_emulated_jpeg_kernels(out_dir):
    """The device kernels of jpeg_decode.cu (dense and sparse IDCT, up-sampling + colour conversion) compiled for the host
    behind tests/host/cuda_emulation_shim.h, with a driver that runs them thread by thread over the entry points' grids."""
    csrc = os.path.join(ROOT, "multimodal-propaganda-meme-classification_b200", "csrc")
    text = open(os.path.join(csrc, "jpeg_decode.cu")).read()
    dev = text[text.index("// ---------------------------------------------------------------------------------------------------- device side"):]
    dev = dev[dev.index("namespace b200 {"):dev.index("using namespace b200;")]
    src = os.path.join(str(out_dir), "emulated_jpeg_kernels.cpp")
    with open(src, "w") as f:
        f.write('#include "cuda_emulation_shim.h"\n#include "jpeg_math.cuh"\n')
        f.write(dev)
        f.write(r"""
using namespace b200;
static unsigned cdiv(int a, int b) { return static_cast<unsigned>((a + b - 1) / b); }
// sparse != 0: sp_off / sp_idx / sp_val carry the batch; else coefs does
extern "C" void emu_jpeg_reconstruct(const short* coefs, const int* sp_off, const unsigned char* sp_idx, const short* sp_val,
                                     int sparse, const unsigned short* qtabs, const long long* table, int n, int max_blocks,
                                     int max_w, int max_h, unsigned char* planes, unsigned char* out) {
  unsigned gx = cdiv(max_blocks, kIdctThreads);
  if (gx > 4096) gx = 4096;
  if (sparse)
    emu_launch(jpeg_idct_kernel<true>, gx, n, 1, kIdctThreads, static_cast<const int16_t*>(nullptr), sp_off, sp_idx, sp_val,
               qtabs, table, planes);
  else
    emu_launch(jpeg_idct_kernel<false>, gx, n, 1, kIdctThreads, coefs, static_cast<const int*>(nullptr),
               static_cast<const uint8_t*>(nullptr), static_cast<const int16_t*>(nullptr), qtabs, table, planes);
  unsigned cx = cdiv(max_w, 32), cy = cdiv(max_h, 8);
  if (cx > 64) cx = 64;
  if (cy > 256) cy = 256;
  emu_launch(jpeg_upsample_color_kernel, cx, cy, n, 256, static_cast<const uint8_t*>(planes), table, out);
}
""")
    lib = os.path.join(str(out_dir), "libemulated_jpeg_kernels.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "tests", "host"), "-I", csrc, src, "-o", lib],
                   check=True)
    return lib
