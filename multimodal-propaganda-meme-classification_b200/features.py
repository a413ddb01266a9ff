"""Feature-extraction service of the SVM baseline, on the engine's kernels.

Reference contract -- baselines/extract_feat.py:52-67, :103-112:

    img_features  = img_model.avgpool(img_model.features(images))          ConvNeXt-tiny, 768-d
    text_features = text_model(text_tokens).pooler_output                  AraBERTv2 (BERT-base), 768-d
    json.dump({"imgfeats": {id: [float, ...]}, "textfeats": {id: [...]}}, features/<split>_feats.json)

consumed by baselines/subtask_2c.py:74-95 (``tr_feats["imgfeats"][id] + tr_feats["textfeats"][id]`` -> linear SVM).

``ConvNeXtTiny``: torchvision/models/convnext.py (stem 4x4/4 conv + LayerNorm2d; stages of CNBlocks = depthwise 7x7 ->
LayerNorm -> Linear 4x -> GELU -> Linear -> layer scale -> residual; LayerNorm2d + 2x2/2 conv between stages;
global average pool), inference only, NHWC bf16:
    stem / down-sampling convolutions   im2col with stride = kernel (a re-layout, no duplication) + tcgen05 GEMM
    depthwise 7x7                       csrc/feature_ops.cu
    LayerNorm over channels             norm.cu (a pixel row of the NHWC matrix is the normalised axis)
    MLP                                 tcgen05 GEMM with the GELU epilogue, then GEMM with the residual epilogue; the layer
                                        scale is folded into the second projection's weights and bias (gamma * (W a + b))
``BertPoolerModel``: the engine's BERT tower (text_tower.py, arch 'bert') + BertPooler (dense on the [CLS] row + tanh,
transformers/models/bert/modeling_bert.py BertPooler) -> ``pooler_output`` fp32 [B, D].
State-dict key names are the libraries' own, so pretrained checkpoints load with ``load_state_dict``.
"""
from __future__ import annotations

import json
import os

import torch

from . import _lib, ops
from .params import ParamStore
from .text_tower import TextConfig, TextTower

CONVNEXT_TINY = dict(depths=(3, 3, 9, 3), dims=(96, 192, 384, 768))


def _need_cuda():
    if not torch.cuda.is_available():
        raise _lib.B200MMError("b200mm needs a CUDA device (sm_100a); there is no CPU fallback")
    _lib.load()


class ConvNeXtTiny:
    """``avgpool(features(x))`` of torchvision's ``convnext_tiny`` (eval mode).  ``__call__(images fp32 NCHW)`` ->
    bf16 [N, 768]; ``features`` / ``avgpool`` mirror the two attributes the reference script calls."""

    def __init__(self, device=None, depths=CONVNEXT_TINY["depths"], dims=CONVNEXT_TINY["dims"], layer_norm_eps=1e-6,
                 seed: int = 0):
        _need_cuda()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.depths, self.dims, self.eps = tuple(depths), tuple(dims), layer_norm_eps
        st = self.store = ParamStore(self.device)
        # ---- fp32-read parameters first (biases, LayerNorm affine, layer scales), then the bf16-shadowed weights
        st.add("features.0.0.bias", (dims[0],), shadow=False)
        st.add("features.0.1.weight", (dims[0],), shadow=False)
        st.add("features.0.1.bias", (dims[0],), shadow=False)
        for si, (depth, dim) in enumerate(zip(depths, dims)):
            f = 2 * si + 1
            for j in range(depth):
                b = f"features.{f}.{j}"
                st.add(f"{b}.block.0.bias", (dim,), shadow=False)
                st.add(f"{b}.block.2.weight", (dim,), shadow=False)
                st.add(f"{b}.block.2.bias", (dim,), shadow=False)
                st.add(f"{b}.block.3.bias", (4 * dim,), shadow=False)
                st.add(f"{b}.block.5.bias", (dim,), shadow=False)
                st.add(f"{b}.layer_scale", (dim,), shadow=False)
            if si + 1 < len(dims):
                d = f"features.{f + 1}"
                st.add(f"{d}.0.weight", (dim,), shadow=False)
                st.add(f"{d}.0.bias", (dim,), shadow=False)
                st.add(f"{d}.1.bias", (dims[si + 1],), shadow=False)
        st.add("features.0.0.weight", (dims[0], 48))                   # OHWI-flattened [96, 4*4*3]
        for si, (depth, dim) in enumerate(zip(depths, dims)):
            f = 2 * si + 1
            for j in range(depth):
                b = f"features.{f}.{j}"
                st.add(f"{b}.block.0.weight", (49, dim))               # depthwise taps, tap-major [7*7, C]
                st.add(f"{b}.block.3.weight", (4 * dim, dim))
                st.add(f"{b}.block.5.weight", (dim, 4 * dim))
            if si + 1 < len(dims):
                st.add(f"features.{2 * si + 2}.1.weight", (dims[si + 1], 4 * dim))     # OHWI-flattened [2C, 2*2*C]
        st.finalize()
        self._folded = None
        self.reset_parameters(seed)

    # ------------------------------------------------------------------ parameters
    @torch.no_grad()
    def reset_parameters(self, seed: int = 0):
        """torchvision's init: trunc_normal(0.02) weights, zero biases, LayerNorm (1, 0), layer scale 1e-6."""
        g = torch.Generator(device=self.device).manual_seed(seed)
        st = self.store
        for name in st.names():
            t = st.p(name)
            if name.endswith("layer_scale"):
                t.fill_(1e-6)
            elif name.endswith(".bias"):
                t.zero_()
            elif t.dim() == 1:
                t.fill_(1.0)
            else:
                t.normal_(0.0, 0.02, generator=g).clamp_(-0.04, 0.04)
        self._refresh()

    def _refresh(self):
        self.store.refresh_shadow()
        self._folded = None

    @torch.no_grad()
    def load_state_dict(self, sd: dict, strict: bool = True):
        """torchvision ``convnext_tiny().state_dict()`` (conv weights OIHW) -> engine layouts."""
        st = self.store
        for name in st.names():
            src = sd[name].to(self.device, torch.float32)
            dst = st.p(name)
            if src.dim() == 4 and src.shape[1] == 1:                  # depthwise [C, 1, 7, 7] -> [49, C]
                dst.copy_(src.reshape(src.shape[0], 49).t())
            elif src.dim() == 4:                                      # OIHW -> OHWI-flattened
                dst.copy_(src.permute(0, 2, 3, 1).reshape(dst.shape))
            else:
                dst.copy_(src.reshape(dst.shape))
        self._refresh()

    def _fold_layer_scales(self):
        """gamma * (W2 a + b2) == (gamma[:, None] * W2) a + gamma * b2: the second projection absorbs the layer scale."""
        st, out = self.store, {}
        for si, depth in enumerate(self.depths):
            for j in range(depth):
                b = f"features.{2 * si + 1}.{j}"
                gamma = st.p(f"{b}.layer_scale")
                out[b] = ((gamma[:, None] * st.p(f"{b}.block.5.weight")).to(torch.bfloat16).contiguous(),
                          (gamma * st.p(f"{b}.block.5.bias")).contiguous())
        return out

    def eval(self):
        return self

    def to(self, *_a, **_k):
        return self

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def features(self, images: torch.Tensor):
        """fp32 NCHW [N, 3, H, W] (H, W multiples of 32) -> (NHWC bf16 matrix [N*h*w, 768], N, h, w)."""
        if images.device != self.device or images.dtype != torch.float32:
            raise _lib.B200MMError("images must be fp32 on the model's CUDA device")
        st, eps = self.store, self.eps
        if self._folded is None:
            self._folded = self._fold_layer_scales()
        N = images.shape[0]
        cols, H, W = ops.im2col_nchw_f32(images, 4, 4, 0, 48)
        x = ops.linear_fwd(cols, st.s("features.0.0.weight"), st.p("features.0.0.bias"))
        x, _, _ = ops.layernorm_fwd(x, st.p("features.0.1.weight"), st.p("features.0.1.bias"), eps)
        for si, (depth, dim) in enumerate(zip(self.depths, self.dims)):
            f = 2 * si + 1
            for j in range(depth):
                b = f"features.{f}.{j}"
                y = ops.dwconv7x7(x, st.s(f"{b}.block.0.weight"), st.p(f"{b}.block.0.bias"), N, H, W, dim)
                y, _, _ = ops.layernorm_fwd(y, st.p(f"{b}.block.2.weight"), st.p(f"{b}.block.2.bias"), eps)
                _, a = ops.linear_gelu_fwd(y, st.s(f"{b}.block.3.weight"), st.p(f"{b}.block.3.bias"))
                w2, b2 = self._folded[b]
                x = ops.linear_fwd(a, w2, b2, residual=x)
            if si + 1 < len(self.dims):
                d = f"features.{f + 1}"
                y, _, _ = ops.layernorm_fwd(x, st.p(f"{d}.0.weight"), st.p(f"{d}.0.bias"), eps)
                cols, H, W = ops.im2col(y, N, H, W, dim, 2, 2, 0)
                x = ops.linear_fwd(cols, st.s(f"{d}.1.weight"), st.p(f"{d}.1.bias"))
        return x, N, H, W

    @torch.no_grad()
    def avgpool(self, feats):
        x, N, H, W = feats
        return ops.avgpool_fwd(x, N, H * W, self.dims[-1])

    def __call__(self, images):
        return self.avgpool(self.features(images))


class BertPoolerModel:
    """``AutoModel.from_pretrained('aubmindlab/bert-base-arabertv2')``'s forward as far as the script uses it:
    ``model(ids [, mask]).pooler_output`` (fp32 [B, D]) and ``.last_hidden_state``."""

    class Output:
        def __init__(self, last_hidden_state, pooler_output):
            self.last_hidden_state, self.pooler_output = last_hidden_state, pooler_output

    def __init__(self, config: TextConfig | None = None, device=None, seed: int = 0):
        _need_cuda()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        cfg = config or TextConfig.bert_base()
        self.cfg = cfg = TextConfig(**{**cfg.__dict__, "prefix": "bert", "dropout": 0.0, "attention_dropout": 0.0})
        st = self.store = ParamStore(self.device)
        self.text = TextTower(cfg, st)
        self.text.register_noshadow()
        st.add("bert.pooler.dense.bias", (cfg.dim,), shadow=False)
        self.text.register_shadowed()
        st.add("bert.pooler.dense.weight", (cfg.dim, cfg.dim))
        st.finalize()
        self.text.bind()
        with torch.no_grad():
            g = torch.Generator(device=self.device).manual_seed(seed)
            self.text.init_parameters(g)
            st.p("bert.pooler.dense.weight").normal_(0.0, 0.02, generator=g)
            st.p("bert.pooler.dense.bias").zero_()
        st.refresh_shadow()

    @torch.no_grad()
    def load_state_dict(self, sd: dict, strict: bool = True):
        """``BertModel.state_dict()`` keys (no ``bert.`` prefix) or the classifier's ``bert.*`` keys."""
        st = self.store
        for name in st.names():
            key = name if name in sd else name[len("bert."):]
            st.p(name).copy_(sd[key].to(self.device, torch.float32).reshape(st.p(name).shape))
        st.refresh_shadow()

    def eval(self):
        return self

    def to(self, *_a, **_k):
        return self

    @torch.no_grad()
    def __call__(self, input_ids, attention_mask=None):
        ids = input_ids.to(self.device)
        mask = torch.ones_like(ids) if attention_mask is None else attention_mask.to(self.device)   # script passes none
        B, S = ids.shape
        h = self.text.forward(ids, mask, training=False)
        cls = ops.gather_rows(h, B, S, 0)
        pooled = ops.linear_fwd_f32(cls, self.store.s("bert.pooler.dense.weight"), self.store.p("bert.pooler.dense.bias"))
        return self.Output(h.view(B, S, -1), ops.tanh_(pooled))


def get_features(loader, img_model, text_model, device=None):
    """baselines/extract_feat.py:52-67, same batch format (``tweet_ids, images, text_tokens``) and return value."""
    device = torch.device(device) if device is not None else img_model.device
    img_feats, text_feats = {}, {}
    for tweet_ids, images, text_tokens in loader:
        images = images.to(device, non_blocking=True).float()
        text_tokens = text_tokens.to(device, non_blocking=True)
        with torch.no_grad():
            img_features = img_model.avgpool(img_model.features(images)).float().cpu().numpy()
            text_features = text_model(text_tokens).pooler_output.float().cpu().numpy()
        for twt_id, img_ft, text_ft in zip(tweet_ids, img_features, text_features):
            img_feats[twt_id] = img_ft.flatten().tolist()
            text_feats[twt_id] = text_ft.flatten().tolist()
    return img_feats, text_feats


def write_features_json(out_path, img_feats: dict, text_feats: dict) -> str:
    """``json.dump({"imgfeats": ..., "textfeats": ...}, open(features/<name>, "w"))`` (extract_feat.py:107-111)."""
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    with open(out_path, "w") as f:
        json.dump({"imgfeats": img_feats, "textfeats": text_feats}, f)
    return out_path


def load_concat_features(path, ids):
    """The consumer's view (baselines/subtask_2c.py:74-84): ``imgfeats[id] + textfeats[id]`` rows for ``ids``."""
    import numpy as np
    with open(path) as f:
        feats = json.load(f)
    return np.array([feats["imgfeats"][i] + feats["textfeats"][i] for i in ids])
