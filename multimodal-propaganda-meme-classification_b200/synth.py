"""Synthetic inputs of the benchmark / parity harness: what the reference's
``MultimodalDataset.__getitem__`` (example_scripts/Multimodal_example_task2C.txt:46-71) hands to the loop, minus the
disk: N(0,1) "normalised" pixels [B,3,H,W] fp32, int64 token ids with random real lengths and pad id 0 beyond,
the matching attention mask, and labels drawn with the train split's class prior."""
from __future__ import annotations

import torch

PAD_ID = 0
TRAIN_PRIOR = 603 / 2143


def synthetic_batch(batch: int, seq_len: int, *, vocab_size: int = 119547, image_size: int = 224, seed: int = 1234,
                    pad_id: int = PAD_ID):
    g = torch.Generator().manual_seed(seed)
    image = torch.randn(batch, 3, image_size, image_size, generator=g)
    lo = min(1000, vocab_size // 2)
    ids = torch.randint(lo, vocab_size, (batch, seq_len), generator=g)
    lengths = torch.randint(min(8, seq_len), seq_len + 1, (batch,), generator=g)
    lengths[0] = seq_len
    pos = torch.arange(seq_len).unsqueeze(0)
    mask = (pos < lengths.unsqueeze(1)).long()
    ids = ids * mask + pad_id * (1 - mask)
    labels = (torch.rand(batch, generator=g) < TRAIN_PRIOR).long()
    return {"text": ids, "text_mask": mask, "image": image, "label": labels}
