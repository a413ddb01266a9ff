"""Throughput of the resize / crop / normalise kernel (csrc/preprocess.cu) on a packed batch of decoded images, and of the
whole device-side input path from entropy-decoded JPEGs (reconstruct + resize) -- CUDA events on the launching stream,
operands resident, 256 images of ~500 x 600 px (the batch's pixels, 232 MB, exceed the 126 MB L2).
Writes gpurun_out/preprocess_bench_r02.json.

    python scripts/bench_preprocess.py [--n 256] [--iters 10]
"""
import argparse
import io
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200mm  # noqa: E402
from b200mm import jpeg, ops  # noqa: E402
from b200mm.data import GpuImageTransform  # noqa: E402


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--resample", default="float", choices=("float", "pillow"), help="pillow: the Pillow-exact kernel")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    uniq = []
    for i in range(16):
        h, w = 440 + 8 * i, 560 + 12 * (i % 8)
        base = torch.from_numpy(rng.random((1, 3, 14, 18), dtype=np.float32))
        im = F.interpolate(base, size=(h, w), mode="bicubic", align_corners=False)[0].clamp(0, 1)
        b = io.BytesIO()
        Image.fromarray((im * 255).byte().permute(1, 2, 0).numpy()).save(b, "JPEG", quality=85, subsampling=2)
        uniq.append(b.getvalue())
    batch = jpeg.pack_jpeg_batch([uniq[i % len(uniq)] for i in range(a.n)])
    dbatch = dict(batch, jpeg_coefs=batch["jpeg_coefs"].to(dev), jpeg_qtabs=batch["jpeg_qtabs"].to(dev))
    packed, table = jpeg.reconstruct_batch(dbatch, dev)
    src_bytes = int((table[1] * table[2] * 3).sum())
    out_bytes = a.n * 3 * 224 * 224 * 4
    t_crop = timed(lambda: ops.preprocess_u8_packed(packed, table, resample=a.resample), a.iters)
    t_square = timed(lambda: ops.preprocess_u8_packed(packed, table, square=True, resample=a.resample), a.iters)
    tr = GpuImageTransform("center_crop", resample=a.resample)
    t_path = timed(lambda: tr.packed(*jpeg.reconstruct_batch(dbatch, dev)), a.iters)
    out = {"images": a.n, "resample": a.resample, "source_megapixels": src_bytes / 3e6, "source_bytes": src_bytes, "output_bytes": out_bytes,
           "resize256_centercrop224_normalize_ms": t_crop, "resize256_centercrop224_gbs": (src_bytes + out_bytes) / t_crop / 1e6,
           "resize224x224_normalize_ms": t_square, "resize224x224_gbs": (src_bytes + out_bytes) / t_square / 1e6,
           "jpeg_reconstruct_plus_resize_ms": t_path,
           "note": "GB/s = (decoded source bytes + fp32 NCHW output bytes) / time; the centre crop reads only the part of "
                   "the source under the crop window, so its true traffic is lower than the bytes counted"}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"preprocess_bench_{a.resample}.json" if a.resample != "float" else "preprocess_bench_r02.json"), "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
