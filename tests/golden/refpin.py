"""Shared by tests/golden/make_reference_golden.py (build container, needs /root/reference) and tests/test_cpu.py
(anywhere): the small configuration, the name-keyed weight initialisation and the synthetic batches of the
REFERENCE-RUN golden vectors.

The generator executes the reference's own model classes and loop functions (source text taken verbatim from
/root/reference at generation time); only the three calls that need the network are replaced: ``AutoModel.from_pretrained``
and ``models.resnet50(pretrained=True)`` / ``timm.create_model(..., pretrained=True)`` return from-config modules of the
same architecture.  The classes hard-code 768-wide text features and a 1000-way ResNet output, so the small configuration
keeps those and shrinks depth, vocabulary, FFN width, sequence length and image size.

Weights are a pure function of (parameter name, seed) -- not of construction order -- so the oracle, built here without the
reference, carries exactly the weights the reference classes carried when the vectors were made.
"""
import zlib

import torch
import torch.nn as nn

SEQ = 16
IMG = 64
VOCAB = 600
TEXT_LAYERS = 2
TEXT_FFN = 1024
MAX_POS = 64


def distilbert():
    """Stands in for AutoModel.from_pretrained('distilbert-base-multilingual-cased') (.txt:158)."""
    from transformers import DistilBertConfig, DistilBertModel
    hf = DistilBertConfig(vocab_size=VOCAB, max_position_embeddings=MAX_POS, dim=768, n_layers=TEXT_LAYERS, n_heads=12,
                          hidden_dim=TEXT_FFN)                 # dropout / attention_dropout stay at the library's 0.1
    hf._attn_implementation = "eager"
    return DistilBertModel(hf)


def bert(arch: str):
    """Stands in for AutoModel.from_pretrained(text_model / english_text_model) (.py:317): BERT ('bert') or
    RoBERTa ('roberta': pad id 1, position ids offset by the pad id)."""
    if arch == "bert":
        from transformers import BertConfig as C, BertModel as M
        kw = dict(pad_token_id=0, layer_norm_eps=1e-12)
    else:
        from transformers import XLMRobertaConfig as C, XLMRobertaModel as M
        kw = dict(pad_token_id=1, layer_norm_eps=1e-5, type_vocab_size=1)
    hf = C(vocab_size=VOCAB, hidden_size=768, num_hidden_layers=TEXT_LAYERS, num_attention_heads=12,
           intermediate_size=TEXT_FFN, max_position_embeddings=MAX_POS, **kw)
    hf._attn_implementation = "eager"
    return M(hf)


def resnet50_small():
    """Stands in for torchvision.models.resnet50(pretrained=True) (.txt:164): Bottleneck ResNet, one block per stage."""
    from torchvision.models.resnet import Bottleneck, ResNet
    return ResNet(Bottleneck, [1, 1, 1, 1], num_classes=1000)


def timm_resnet18():
    """Stands in for timm.create_model('resnet18', pretrained=True) (.py:569-570; timm is not installed): torchvision's
    BasicBlock ResNet -- same state-dict keys as timm's -- with timm's ``reset_classifier``."""
    from torchvision.models.resnet import BasicBlock, ResNet

    class TimmLikeResNet(ResNet):
        def reset_classifier(self, num_classes):
            assert num_classes == 0
            self.fc = nn.Identity()

    return TimmLikeResNet(BasicBlock, [1, 1, 1, 1], num_classes=1000)


def reseed_by_name(model: nn.Module, seed: int = 0) -> nn.Module:
    """Every floating-point entry of the state dict becomes a function of (its name, seed)."""
    with torch.no_grad():
        for name, t in model.state_dict().items():
            if not t.is_floating_point():
                continue
            g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) & 0x7FFFFFFF)
            r = torch.randn(t.shape, generator=g)
            if name.endswith("running_var"):
                t.copy_(1.0 + 0.2 * r.abs())
            elif name.endswith("running_mean"):
                t.copy_(0.1 * r)
            elif t.dim() >= 2:
                t.copy_(r / (t.numel() / t.shape[0]) ** 0.5)
            elif name.endswith("weight"):                       # LayerNorm / BatchNorm scale
                t.copy_(1.0 + 0.1 * r)
            else:
                t.copy_(0.1 * r)
    return model


def batches(n_batches: int, batch: int, *, seed: int, captions: bool, pad_id: int = 0, caption_pad_id: int = 1):
    """The reference's batch dicts (.txt:61-69; .py:293-303) with synthetic content; row 0 of every batch is full length."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for b in range(n_batches):
        def tokens(pad):
            ids = torch.randint(5, VOCAB, (batch, SEQ), generator=g)
            lens = torch.randint(3, SEQ + 1, (batch,), generator=g)
            lens[0] = SEQ
            mask = (torch.arange(SEQ).unsqueeze(0) < lens.unsqueeze(1)).long()
            return torch.where(mask.bool(), ids, torch.full_like(ids, pad)), mask
        text, mask = tokens(pad_id)
        d = {"id": [f"data/arabic_memes_fb_insta_pinterest/x/img_{b}_{i}.jpg" for i in range(batch)],
             "text": text, "text_mask": mask, "image": torch.randn(batch, 3, IMG, IMG, generator=g),
             "label": (torch.rand(batch, generator=g) < 0.4).long()}
        d["label"][0], d["label"][1] = 1, 0                   # both classes in every batch (ROC / F1 well defined)
        if captions:
            d["caption_text"], d["caption_text_mask"] = tokens(caption_pad_id)
        out.append(d)
    return out


class ListLoader(list):
    """A 'DataLoader' that is a list of ready batches; ``.dataset`` has the length the loops divide by."""

    @property
    def dataset(self):
        return range(sum(len(b["id"]) for b in self))


def param_norms(model: nn.Module, grads: bool = False):
    """name -> L2 norm of the parameter (or of its gradient; None where autograd produced none, e.g. an unused pooler)."""
    def norm(p):
        t = p.grad if grads else p.detach()
        return None if t is None else float(t.double().norm())
    return {n: norm(p) for n, p in model.named_parameters()}


# ------------------------------------------------------------------------------------------------- Dataset run
DATASET_TEXTS = ["the memes! هذا ميم propaganda", "propaganda? the meme.", "ميم"]
DATASET_LABELS = [1, 0, 1]
DATASET_SIZES = [(300, 400), (420, 310), (256, 256)]          # (height, width) of the synthetic JPEG files
DATASET_AUG_SEED = 77                                        # torch.manual_seed(DATASET_AUG_SEED + i) before sample i of the participant Dataset


class EncodePlusTokenizer:
    """A tokenizer built offline from a 130-entry WordPiece vocabulary, with the ``encode_plus`` entry point the reference's
    Dataset calls (.txt:54-56; transformers 5.x dropped the name: for one text it is ``__call__``)."""

    def __init__(self, workdir):
        import os
        from transformers import DistilBertTokenizer
        ar = list("ابتثجحخدذرزسشصضطظعغفقكلمنهوي")
        lat = list("abcdefghijklmnopqrstuvwxyz")
        vocab = (["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + lat + ["##" + c for c in lat] +
                 ["meme", "propaganda", "the", "##s", "!", "?", "."] + ar + ["##" + c for c in ar])
        path = os.path.join(str(workdir), "vocab.txt")
        with open(path, "w", encoding="utf-8") as f:
            f.write("\n".join(vocab))
        self.tok = DistilBertTokenizer(path, do_lower_case=False)

    def encode_plus(self, text, **kw):
        return self.tok(text, **kw)


def dataset_jpegs():
    """The Dataset run's image files: deterministic photo-like content, encoded by Pillow (baseline, progressive, 4:4:4)."""
    import io

    import numpy as np
    import torch.nn.functional as F
    from PIL import Image
    rng = np.random.default_rng(2024)
    files = []
    for i, (h, w) in enumerate(DATASET_SIZES):
        base = torch.from_numpy(rng.random((1, 3, 7, 9), dtype=np.float32))
        im = F.interpolate(base, size=(h, w), mode="bicubic", align_corners=False)[0]
        im = im + 0.04 * torch.from_numpy(rng.standard_normal((3, h, w)).astype(np.float32))
        b = io.BytesIO()
        Image.fromarray((im.clamp(0, 1) * 255).byte().permute(1, 2, 0).numpy()).save(
            b, "JPEG", quality=(85, 70, 92)[i], subsampling=(2, 1, 0)[i], progressive=(i == 1))
        files.append(b.getvalue())
    return files


# ------------------------------------------------------------------------------------------------- SVM-baseline consumer run
def svm_feature_corpus():
    """Synthetic ids / labels / feature vectors of the feature-extraction contract (baselines/extract_feat.py:52-67 ->
    baselines/subtask_2c.py:74-95): 40 train + 12 dev items, 8-d image and 6-d text features with a weak class signal."""
    g = torch.Generator().manual_seed(99)
    out = {}
    for split, n in (("train", 40), ("dev", 12)):
        ids = [f"data/x/{split}_{i}.jpg" for i in range(n)]
        y = (torch.rand(n, generator=g) < 0.4).long()
        y[0], y[1] = 1, 0
        img = torch.randn(n, 8, generator=g) + 0.8 * y[:, None]
        txt = torch.randn(n, 6, generator=g) - 0.5 * y[:, None]
        out[split] = {"id": ids, "class_label": ["propaganda" if v else "not_propaganda" for v in y.tolist()],
                      "imgfeats": {i: img[k].tolist() for k, i in enumerate(ids)},
                      "textfeats": {i: txt[k].tolist() for k, i in enumerate(ids)}}
    return out
