"""Input pipeline with the image transform on the GPU.

The reference's ``MultimodalDataset.__getitem__`` (example_scripts/Multimodal_example_task2C.txt:46-71;
HEAD script Multimodal_example_task2C.py:262-304) tokenises the text and runs the whole torchvision transform on the
host for every sample of every epoch, and ships fp32 pixels to the device: its loop is input-bound (31.6 s/it on a T4,
SURVEY.md §3.1).  Here the host only does what must stay there:

  * ``read_data``         the JSON reader of the driver scripts (.txt:20-35, .py:96-112), without pandas;
  * ``TokenCache``        every text is tokenised ONCE (ids + mask as int64 [N, S]); ``__getitem__`` slices the cache;
  * ``MemeDataset``       returns the reference's batch dict, but ``image`` is the DECODED uint8 HWC image (any size);
  * ``collate_packed``    packs the batch's images into ONE pinned byte buffer + an offset / height / width table
                          (ops.pack_images): two host->device copies per batch whatever its size, 4x fewer bytes than
                          fp32 pixels even before counting the resize;
  * ``file_bytes_loader`` + ``jpeg.collate_jpeg``   (optional) the worker only Huffman-decodes the JPEG; the rest of the
                          decode runs on the GPU, bit-identical to Pillow (csrc/jpeg_decode.cu);
  * ``GpuImageTransform`` Resize / CenterCrop (or Resize((224, 224))) / RandomHorizontalFlip / ToTensor / Normalize as ONE
                          kernel on the copy stream (csrc/preprocess.cu), called by ``loop.DevicePrefetcher``; with
                          ``augment=True`` the HEAD script's whole train transform (.py:222-235): + ColorJitter(.1, .1,
                          .1, .1) + RandomRotation(15) as two more kernels (csrc/augment.cu).

The augmentations follow torchvision's float-tensor
operators with per-image parameters drawn as ``ColorJitter.get_params`` / ``RandomRotation.get_params`` draw them; the
reference applies the same operators to PIL images (uint8 intermediates: up to 1/255 of rounding per operator).
"""
from __future__ import annotations

import json
import math

import torch
from torch.utils.data import Dataset

from . import ops

L2ID = {"not_propaganda": 0, "propaganda": 1}     # .txt:17 / .py:114


def read_data(fpath, is_test: bool = False):
    """Reference: Multimodal_example_task2C.py:96-112 (same keys; a dict of lists instead of a DataFrame)."""
    data = {"id": [], "text": [], "image": []} if is_test else {"id": [], "text": [], "image": [], "label": []}
    with open(fpath, encoding="utf-8") as f:
        for obj in json.load(f):
            data["id"].append(obj["id"])
            data["image"].append(obj["img_path"])
            data["text"].append(obj["text"])
            if not is_test:
                data["label"].append(obj["class_label"])
    return data


class TokenCache:
    """ids / attention mask of every text, tokenised once.  ``tokenizer(text) -> (ids, mask)`` lists or tensors of
    length <= max_len; shorter ones are padded with ``pad_id`` ("padding='max_length'", .txt:54-56)."""

    def __init__(self, texts, tokenizer, max_len: int, pad_id: int = 0):
        n = len(texts)
        self.ids = torch.full((n, max_len), pad_id, dtype=torch.int64)
        self.mask = torch.zeros(n, max_len, dtype=torch.int64)
        for i, t in enumerate(texts):
            ids, mask = tokenizer(t)
            ids = torch.as_tensor(ids, dtype=torch.int64)[:max_len]
            mask = torch.as_tensor(mask, dtype=torch.int64)[:max_len]
            self.ids[i, :ids.numel()] = ids
            self.mask[i, :mask.numel()] = mask

    def __getitem__(self, i):
        return self.ids[i], self.mask[i]


def pil_loader(path):
    """Image.open(path).convert('RGB') as a uint8 [H, W, 3] tensor (.txt:50; .py:270)."""
    import numpy as np
    from PIL import Image
    with Image.open(path) as im:
        return torch.from_numpy(np.asarray(im.convert("RGB"), dtype=np.uint8).copy())


def file_bytes_loader(path):
    """The image FILE as a uint8 1-D tensor: for ``collate_fn=jpeg.collate_jpeg``, which Huffman-decodes the batch in the
    loader worker and leaves inverse DCT / up-sampling / colour conversion to the GPU (pixels equal ``pil_loader``'s)."""
    with open(path, "rb") as f:
        return torch.frombuffer(bytearray(f.read()), dtype=torch.uint8)


class MemeDataset(Dataset):
    """``MultimodalDataset`` with the per-sample work hoisted: the batch dict has the reference's keys (``id``, ``text``,
    ``text_mask``, ``image``, ``label``; + ``caption_text`` / ``caption_text_mask`` when caption tokens are given,
    .py:293-303), ``image`` being the decoded uint8 image.  ``image_loader(path) -> uint8 [H, W, 3]``."""

    def __init__(self, ids, text_data, image_data, labels, *, tokenizer, max_len: int = 128, pad_id: int = 0,
                 image_loader=pil_loader, is_test: bool = False, captions=None, caption_tokenizer=None,
                 caption_pad_id: int = 1):
        self.ids = list(ids)
        self.image_data = list(image_data)
        self.labels = None if labels is None else [L2ID[l] if isinstance(l, str) else int(l) for l in labels]
        self.is_test = is_test
        self.tokens = TokenCache(list(text_data), tokenizer, max_len, pad_id)
        self.caption_tokens = None
        if captions is not None:
            self.caption_tokens = TokenCache(list(captions), caption_tokenizer, max_len, caption_pad_id)
        self.image_loader = image_loader

    def __len__(self):
        return len(self.ids)

    def __getitem__(self, index):
        ids, mask = self.tokens[index]
        fdata = {"id": self.ids[index], "text": ids, "text_mask": mask,
                 "image": self.image_loader(self.image_data[index])}
        if self.caption_tokens is not None:
            c, cm = self.caption_tokens[index]
            fdata["caption_text"], fdata["caption_text_mask"] = c, cm
        if not self.is_test and self.labels is not None:
            fdata["label"] = torch.tensor(self.labels[index], dtype=torch.long)
        return fdata


def collate_packed(samples, pin: bool = True):
    """DataLoader ``collate_fn``: stacks the token tensors / labels and PACKS the variable-size uint8 images
    (``image_packed`` byte buffer + ``image_table`` int64 [3, n]); ``loop.DevicePrefetcher`` turns the pair into the
    normalised [n, 3, 224, 224] ``image`` on the device.  Images that already share one shape are stacked to a
    [n, H, W, 3] uint8 ``image`` instead (the fixed-size fast path)."""
    out = {"id": [s["id"] for s in samples]}
    for k in ("text", "text_mask", "caption_text", "caption_text_mask", "label"):
        if k in samples[0]:
            t = torch.stack([s[k] for s in samples])
            out[k] = t.pin_memory() if pin else t
    imgs = [s["image"] for s in samples]
    if all(im.shape == imgs[0].shape for im in imgs) and imgs[0].shape[1] % 4 == 0:
        t = torch.stack(imgs)
        out["image"] = t.pin_memory() if pin else t
    else:
        out["image_packed"], out["image_table"] = ops.pack_images(imgs, pin=pin)
    return out


class GpuImageTransform:
    """The reference's image transform as device kernels, applied to a batch that has just crossed PCIe.

    mode 'center_crop': Resize(256) -> CenterCrop(224) -> ToTensor -> Normalize          (.txt:37-41)
    mode 'square'     : Resize((224, 224)) [-> RandomHorizontalFlip] -> ToTensor -> Normalize   (.py:222-235)
    ``train=True`` draws the flip flags (p = 0.5) from a seeded host generator, one per image.
    ``augment=True`` (with ``train=True``, mode 'square') completes the HEAD script's train transform (.py:224-233):
    ColorJitter(brightness, contrast, saturation, hue) in a per-image random operator order and
    RandomRotation(degrees) (nearest, zero fill) between the flip and ToTensor / Normalize.
    ``rng='batched'`` (default) draws a batch's parameters in a few vectorised calls from the transform's own seeded
    generator; ``rng='torchvision'`` draws them image by image from torch's GLOBAL generator with the very calls and in
    the very order the script's Compose makes them (``torch.rand(1)`` for the flip, ``ColorJitter.get_params``,
    ``RandomRotation.get_params``), so that after the same ``torch.manual_seed`` the device transform takes the decisions
    the script's Dataset would take for those samples (tests/test_cpu.py replays a run of the script's own Dataset).
    ``resample='pillow'`` (packed batches) resizes with Pillow's own 8-bit two-pass arithmetic instead of the float
    kernel: the tensor equals the one the script's Dataset builds from the PIL image bit for bit (csrc/preprocess_pil.cu;
    verified on the host build against Pillow and a run of the script's Dataset, not yet timed on a GPU).  With
    ``augment=True`` the colour operators and the rotation switch to Pillow's uint8 arithmetic as well
    (csrc/augment_pil.cu): for the same draws (``rng='torchvision'`` + the same seed) the tensor equals the script's."""

    def __init__(self, mode: str = "center_crop", *, resize: int = 256, crop: int = 224, train: bool = False,
                 seed: int = 0, mean=ops.IMAGENET_MEAN, std=ops.IMAGENET_STD, augment: bool = False,
                 brightness: float = 0.1, contrast: float = 0.1, saturation: float = 0.1, hue: float = 0.1,
                 degrees: float = 15.0, rng: str = "batched", resample: str = "float"):
        if mode not in ("center_crop", "square"):
            raise ValueError(f"unknown image transform mode {mode!r}")
        if augment and not (train and mode == "square"):
            raise ValueError("augment=True is the HEAD script's TRAIN transform: it needs train=True and mode='square'")
        if rng not in ("batched", "torchvision"):
            raise ValueError("rng must be 'batched' or 'torchvision'")
        if resample not in ("float", "pillow"):
            raise ValueError("resample must be 'float' or 'pillow'")
        self.resample = resample
        if not 0.0 <= hue <= 0.5 or min(brightness, contrast, saturation, degrees) < 0.0:
            raise ValueError("ColorJitter / RandomRotation ranges must be non-negative (hue <= 0.5)")
        self.mode, self.resize, self.crop, self.train = mode, resize, crop, train
        self.mean, self.std = tuple(mean), tuple(std)
        self.augment = augment
        self.jitter = (brightness, contrast, saturation, hue)
        self.degrees = degrees
        self.gen = torch.Generator().manual_seed(seed)
        self.rng = rng

    def _flip(self, n, device):
        if not (self.train and self.mode == "square"):
            return None
        f = (torch.rand(n, generator=self.gen) < 0.5).to(torch.uint8)
        return f.pin_memory().to(device, non_blocking=True)

    def draw_raw(self, n):
        """Per-image draws of ColorJitter.get_params ($SP/torchvision/transforms/transforms.py:1237-1266: a permutation of
        the four operators, factors ~ U[max(0, 1 - x), 1 + x], hue shift ~ U[-hue, hue]) and RandomRotation.get_params
        (:1354-1361: angle ~ U[-degrees, degrees]): perm int64 [n, 4], factors fp32 [n, 4], angles fp64 [n] (degrees)."""
        perm = torch.rand(n, 4, generator=self.gen).argsort(1)      # n uniform permutations of the four operators, one call
        u = torch.rand(n, 5, generator=self.gen)
        factors = torch.empty(n, 4)
        for k, x in enumerate(self.jitter[:3]):
            lo = max(0.0, 1.0 - x)
            factors[:, k] = lo + u[:, k] * (1.0 + x - lo)
        factors[:, 3] = (2.0 * u[:, 3] - 1.0) * self.jitter[3]
        angles = ((2.0 * u[:, 4] - 1.0) * self.degrees).double()
        return perm, factors, angles

    @staticmethod
    def pack_augment(perm, factors, angles):
        """The kernel's parameter tables (csrc/augment.cu): order int32 [n] -- 2 bits per operator, first applied in the
        low bits -- and params fp32 [n, 8] = factors | inverse rotation matrix.  torchvision's ``rotate`` hands -angle to
        ``_get_inverse_affine_matrix`` (functional.py:1131); for a rotation about the centre that matrix is
        [[cos t, sin t, 0], [-sin t, cos t, 0]] with t = radians(-angle), computed in double and rounded to fp32."""
        order = (perm << torch.tensor([0, 2, 4, 6])).sum(1).to(torch.int32)
        t = angles.double().neg() * (math.pi / 180.0)
        rot = torch.stack([torch.cos(t), torch.sin(t), torch.sin(t).neg(), torch.cos(t)], dim=1).float()
        return order, torch.cat([factors.float(), rot], dim=1).contiguous()

    def draw_augment(self, n):
        return self.pack_augment(*self.draw_raw(n))

    @staticmethod
    def pack_augment_pil(perm, factors, angles, width, height):
        """The parameter tables of the Pillow-exact kernels (csrc/augment_pil.cu), built the way Pillow / torchvision's PIL
        back end build theirs: alpha = the three factors as C floats; hue = ``np.array(hue_factor * 255).astype(np.uint8)``
        (functional_pil.adjust_hue); affine = ``Image.rotate``'s matrix -- cos / sin rounded to 15 digits, centre
        (w / 2, h / 2) -- converted to 16.16 fixed point with the half-pixel offsets of libImaging's affine_fixed."""
        import numpy as np
        order = (perm << torch.tensor([0, 2, 4, 6])).sum(1).to(torch.int32)
        alpha = factors[:, :3].float().contiguous()
        with np.errstate(invalid="ignore"):
            hue = torch.tensor([int(np.array(float(h) * 255).astype(np.uint8)) for h in factors[:, 3].tolist()],
                               dtype=torch.int32)

        def fix(v):
            v = v * 65536.0 + 0.5
            return int(v) if v >= 0.0 else int(math.floor(v))

        rows = []
        for angle in angles.tolist():
            a = -math.radians(angle % 360.0)
            m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0, round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]
            cx, cy = width / 2, height / 2
            m[2] = m[0] * (-cx) + m[1] * (-cy) + m[2]
            m[5] = m[3] * (-cx) + m[4] * (-cy) + m[5]
            m[2] += cx
            m[5] += cy
            rows.append([fix(m[0]), fix(m[1]), fix(m[2] + m[0] * 0.5 + m[1] * 0.5),
                         fix(m[3]), fix(m[4]), fix(m[5] + m[3] * 0.5 + m[4] * 0.5)])
        return order, alpha, hue, torch.tensor(rows, dtype=torch.int32)

    def draw_torchvision(self, n):
        """flips bool [n], perm, factors, angles (as ``draw_raw``) from torch's global generator, per image, in the order
        of the script's Compose (.py:222-233): RandomHorizontalFlip.forward, ColorJitter.forward, RandomRotation.forward."""
        import torchvision.transforms as T
        b, c, s, h = self.jitter
        rng = lambda x: [max(0.0, 1.0 - x), 1.0 + x]          # noqa: E731  (ColorJitter._check_input)
        flips, perms, factors, angles = [], [], [], []
        for _ in range(n):
            flips.append(bool(torch.rand(1) < 0.5))
            fn_idx, fb, fc, fs, fh = T.ColorJitter.get_params(rng(b), rng(c), rng(s), [-h, h])
            perms.append(fn_idx)
            factors.append([fb, fc, fs, fh])
            angles.append(T.RandomRotation.get_params([-float(self.degrees), float(self.degrees)]))
        return (torch.tensor(flips), torch.stack(perms), torch.tensor(factors, dtype=torch.float32),
                torch.tensor(angles, dtype=torch.float64))

    def _draws(self, n, device):
        """(flip flags on the device or None, (order, params) host tensors or None) for a batch of n images."""
        if self.augment and self.rng == "torchvision":
            flips, perm, factors, angles = self.draw_torchvision(n)
            return flips.to(torch.uint8).pin_memory().to(device, non_blocking=True), self.pack_augment(perm, factors, angles)
        return self._flip(n, device), None

    def _finish(self, img01, drawn=None):
        order, params = drawn if drawn is not None else self.draw_augment(img01.shape[0])
        order = order.pin_memory().to(img01.device, non_blocking=True)
        params = params.pin_memory().to(img01.device, non_blocking=True)
        return ops.augment_jitter_rotate(img01, order, params, mean=self.mean, std=self.std)[0]

    def _packed_pil_augment(self, packed, table):
        """resample='pillow' with augment: uint8 all the way, Pillow's arithmetic in every operator."""
        n, dev = table.shape[1], packed.device
        if self.rng == "torchvision":
            flips, perm, factors, angles = self.draw_torchvision(n)
            flip = flips.to(torch.uint8).pin_memory().to(dev, non_blocking=True)
        else:
            flip = self._flip(n, dev)
            perm, factors, angles = self.draw_raw(n)
        u8 = ops.preprocess_u8_packed_pil_u8(packed, table, resize=self.resize, crop=self.crop, square=True, flip=flip)
        tables = self.pack_augment_pil(perm, factors, angles, self.crop, self.crop)
        order, alpha, hue, affine = (t.pin_memory().to(dev, non_blocking=True) for t in tables)
        return ops.augment_pil(u8, order, alpha, hue, affine, mean=self.mean, std=self.std)

    def packed(self, packed, table):
        if self.augment and self.resample == "pillow":
            return self._packed_pil_augment(packed, table)
        flip, drawn = self._draws(table.shape[1], packed.device)
        square = self.mode == "square"
        if self.augment:    # resize + flip + ToTensor to [0, 1] here, Normalize at the end of the augmentation kernel
            return self._finish(ops.preprocess_u8_packed(packed, table, resize=self.resize, crop=self.crop, square=square,
                                                         flip=flip, mean=(0.0, 0.0, 0.0), std=(1.0, 1.0, 1.0),
                                                         resample=self.resample), drawn)
        return ops.preprocess_u8_packed(packed, table, resize=self.resize, crop=self.crop, square=square, flip=flip,
                                        mean=self.mean, std=self.std, resample=self.resample)

    def fixed(self, images):
        if self.augment and self.resample == "pillow":
            raise ValueError("resample='pillow' with augment=True takes packed batches (data.collate_packed / "
                             "jpeg.collate_jpeg): the Pillow-exact path starts from the decoded image")
        flip, drawn = self._draws(images.shape[0], images.device)
        if self.augment:
            return self._finish(ops.u8_normalize(images, flip=flip, mean=(0.0, 0.0, 0.0), std=(1.0, 1.0, 1.0)), drawn)
        return ops.u8_normalize(images, flip=flip, mean=self.mean, std=self.std)
