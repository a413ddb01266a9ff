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
