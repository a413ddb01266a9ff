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
  * ``GpuImageTransform`` Resize / CenterCrop (or Resize((224, 224))) / RandomHorizontalFlip / ToTensor / Normalize as ONE
                          kernel on the copy stream (csrc/preprocess.cu), called by ``loop.DevicePrefetcher``.

JPEG decode stays on the host (PIL, as in the reference); ColorJitter / RandomRotation of the HEAD script's train
transform (.py:224-233) are not reproduced (documented in DESIGN.md; the parity runs use fixed inputs).
"""
from __future__ import annotations

import json

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
    """The reference's image transform as one device kernel, applied to a batch that has just crossed PCIe.

    mode 'center_crop': Resize(256) -> CenterCrop(224) -> ToTensor -> Normalize          (.txt:37-41)
    mode 'square'     : Resize((224, 224)) [-> RandomHorizontalFlip] -> ToTensor -> Normalize   (.py:222-235)
    ``train=True`` draws the flip flags (p = 0.5) from a seeded host generator, one per image."""

    def __init__(self, mode: str = "center_crop", *, resize: int = 256, crop: int = 224, train: bool = False,
                 seed: int = 0, mean=ops.IMAGENET_MEAN, std=ops.IMAGENET_STD):
        if mode not in ("center_crop", "square"):
            raise ValueError(f"unknown image transform mode {mode!r}")
        self.mode, self.resize, self.crop, self.train = mode, resize, crop, train
        self.mean, self.std = tuple(mean), tuple(std)
        self.gen = torch.Generator().manual_seed(seed)

    def _flip(self, n, device):
        if not (self.train and self.mode == "square"):
            return None
        f = (torch.rand(n, generator=self.gen) < 0.5).to(torch.uint8)
        return f.pin_memory().to(device, non_blocking=True)

    def packed(self, packed, table):
        return ops.preprocess_u8_packed(packed, table, resize=self.resize, crop=self.crop, square=self.mode == "square",
                                        flip=self._flip(table.shape[1], packed.device), mean=self.mean, std=self.std)

    def fixed(self, images):
        return ops.u8_normalize(images, flip=self._flip(images.shape[0], images.device), mean=self.mean, std=self.std)
