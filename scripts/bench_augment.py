"""Micro-benchmark of the train-time augmentation kernels (csrc/augment.cu) at the headline batch: 256 images of
3 x 224 x 224, ColorJitter + RandomRotation + Normalize behind the resize.  CUDA events on the launching stream, input and
output (154 MB each) larger than the 126 MB L2.  Writes gpurun_out/augment_bench_r02.json.

    python scripts/bench_augment.py [--n 256] [--iters 20]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200mm  # noqa: E402
from b200mm import ops  # noqa: E402
from b200mm.data import GpuImageTransform  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    n, H, W = a.n, 224, 224
    img = torch.rand(n, 3, H, W, device=dev)
    u8 = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device=dev)
    tr = GpuImageTransform("square", train=True, augment=True, seed=1)
    order, params = (t.to(dev) for t in tr.draw_augment(n))

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.iters * 1e3          # us per call

    px = n * H * W
    t_aug = timed(lambda: ops.augment_jitter_rotate(img, order, params))
    t_norm = timed(lambda: ops.u8_normalize(u8))
    t_full = timed(lambda: tr.fixed(u8))
    # the Pillow-exact pair (uint8 operators, fixed-point rotation) on the same batch
    perm, factors, angles = tr.draw_raw(n)
    o2, alpha, hue, affine = (t.to(dev) for t in GpuImageTransform.pack_augment_pil(perm, factors, angles, W, H))
    t_pil = timed(lambda: ops.augment_pil(u8, o2, alpha, hue, affine))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    out = {"images": n, "pixels": px,
           "augment_jitter_rotate_us": t_aug, "augment_bytes": 36 * px,
           "augment_gbs": 36 * px / t_aug / 1e3,
           "note": "two kernels: grey-mean pass reads 12 B/pixel; gather pass reads 12 B/pixel and writes 12 B/pixel",
           "u8_normalize_us": t_norm, "u8_normalize_gbs": 15 * px / t_norm / 1e3,
           "full_train_transform_fixed_us": t_full, "augment_pil_us": t_pil,
           "full_note": "u8_normalize to [0, 1] + parameter draw on the host + two small H2D copies + the two kernels",
           "measured_peaks": peaks}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "augment_bench_r02.json"), "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
