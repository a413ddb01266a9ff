"""Host half of the split JPEG decode against Pillow's full decode, one core, no GPU needed: per-file times over
smooth / noisy x baseline / progressive synthetic photos of ~0.3 MP (quality 85, 4:2:0).  Writes
profiles/jpeg_host_bench_r02.json when run with --save.

    python scripts/bench_jpeg_host.py [--n 32] [--save]
"""
import argparse
import io
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200mm  # noqa: E402
from b200mm import jpeg  # noqa: E402


def make(n, noise, prog):
    rng = np.random.default_rng(0)
    out = []
    for i in range(n):
        h, w = 440 + 8 * (i % 16), 560 + 12 * (i % 8)
        im = F.interpolate(torch.from_numpy(rng.random((1, 3, 14, 18), dtype=np.float32)), size=(h, w), mode="bicubic",
                           align_corners=False)[0]
        im = im + noise * torch.from_numpy(rng.standard_normal((3, h, w)).astype(np.float32))
        b = io.BytesIO()
        Image.fromarray((im.clamp(0, 1) * 255).byte().permute(1, 2, 0).numpy()).save(b, "JPEG", quality=85, subsampling=2,
                                                                                     progressive=prog)
        out.append(b.getvalue())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=32)
    ap.add_argument("--save", action="store_true")
    a = ap.parse_args()
    res = {"cpu": open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0].strip(": \t") if os.path.exists(
        "/proc/cpuinfo") else "?", "files_per_set": a.n, "sets": {}}
    buf = np.empty(4_000_000, dtype=np.int16)
    for name, noise, prog in (("smooth baseline", 0.0, False), ("noisy baseline", 0.03, False),
                              ("smooth progressive", 0.0, True), ("noisy progressive", 0.03, True)):
        files = make(a.n, noise, prog)
        best_p = best_h = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            for f in files:
                np.asarray(Image.open(io.BytesIO(f)).convert("RGB"))
            best_p = min(best_p, (time.perf_counter() - t0) / len(files) * 1e3)
            t0 = time.perf_counter()
            for f in files:
                jpeg.entropy_decode(f, out=buf)
            best_h = min(best_h, (time.perf_counter() - t0) / len(files) * 1e3)
        dense = jpeg.pack_jpeg_batch(files, pin=False)
        sparse = jpeg.pack_jpeg_batch(files, pin=False, sparse=True)
        nbytes = lambda b: sum(v.numel() * v.element_size() for v in b.values() if isinstance(v, torch.Tensor))  # noqa: E731
        res["sets"][name] = {"file_kb": sum(map(len, files)) / len(files) / 1024, "pillow_ms_per_file": best_p,
                             "host_half_ms_per_file": best_h, "host_time_saved_frac": 1 - best_h / best_p,
                             "ship_mb_dense": nbytes(dense) / 1e6, "ship_mb_sparse": nbytes(sparse) / 1e6,
                             "ship_mb_pixels": int((dense["jpeg_table"][:, 0] * dense["jpeg_table"][:, 1] * 3).sum()) / 1e6}
    print(json.dumps(res, indent=1))
    if a.save:
        json.dump(res, open(os.path.join(ROOT, "profiles", "jpeg_host_bench_r02.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
