"""``MultimodalClassifierHEAD``: the HEAD script's three-tower model on the sm_100a engine.

Reference: example_scripts/Multimodal_example_task2C.py:587-685 (``MultimodalClassifier(fusion_method)``), with
``LLMWithClassificationHead`` (CLS pooling, :307-360), ``CustomDenseNet161`` (timm ResNet-18 + ``fine_tune`` MLP,
:562-585) and ``ConcatAttention3`` (:476-499).  Only ``fusion_method="concatenation"`` is live code in the reference
(the other fusion classes are called with the wrong arity), so that is what exists here; anything
else raises the reference's own ``ValueError``.

    text     = text_fc(dropout_.3(AraBERT(text, mask)[:, 0]))            Linear 768->512 + BatchNorm1d + ReLU
    caption  = caption_text_fc(dropout_.3(RoBERTa(caption, mask)[:, 0]))  (same)
    image    = fine_tune(ResNet18(image))                                  Linear-ReLU-Dropout(.35)-Linear on 512
    x        = cat(text, image, caption)                                   [B, 1536]
    w        = softmax(relu(BN(Linear_1536(x))));  fused = relu(BN(Linear_512(w * x)))
    logit    = BatchNorm1d(1)(Linear(512, 1)(fused)).squeeze(1)            -> sigmoid focal loss

State-dict keys are the reference module's (``text_model.model.*``, ``caption_text_model.model.*``,
``image_model.image_model.*``, ``image_model.fine_tune.{0,3}.*``, ``text_fc.{0,1}.*``, ``fusion_layer.*``,
``output_fc.{0,1}.*``), so ``get_params(lr)`` groups parameters exactly as :645-664 does (including the substring
quirk that puts the caption tower in the 0.8 x lr text group).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, ops
from .image_tower import ImageConfig, ImageTower
from .model import MultimodalClassifier, _EngineFunction
from .params import ParamStore
from .text_tower import TextConfig, TextTower, _mix

_BN1D = (("text_fc.1", 512), ("caption_text_fc.1", 512), ("fusion_layer.attention_layer.1", 1536),
         ("fusion_layer.reduce.1", 512), ("output_fc.1", 1))
_LINEAR = (("text_fc.0", 512, None), ("caption_text_fc.0", 512, None), ("image_model.fine_tune.0", 512, 512),
           ("image_model.fine_tune.3", 512, 512), ("fusion_layer.attention_layer.0", 1536, 1536),
           ("fusion_layer.reduce.0", 512, 1536))


class MultimodalClassifierHEAD(MultimodalClassifier):
    def __init__(self, fusion_method: str = "concatenation", *, text_config: TextConfig | None = None,
                 caption_config: TextConfig | None = None, image_config: ImageConfig | None = None, device=None,
                 seed: int = 42, init: bool = True, text_dropout: float = 0.3, image_dropout: float = 0.35):
        nn.Module.__init__(self)
        if fusion_method != "concatenation":
            # 'mca' / 'cross_modal' / 'self_attention' cannot run in the reference either (wrong call arity, :678)
            raise ValueError(f"Unsupported fusion method: {fusion_method}")
        if not torch.cuda.is_available():
            raise _lib.B200MMError("b200mm needs a CUDA device (sm_100a); there is no CPU fallback")
        _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.fusion_method = fusion_method
        self.num_classes, self.squeeze_output, self.pooling = 1, True, "cls"
        self.head_dropout, self.image_dropout = text_dropout, image_dropout
        self.seed, self._step = seed, 0
        self.tcfg = text_config or TextConfig.bert_base()
        self.ccfg = caption_config or TextConfig.roberta_base()
        self.icfg = image_config or ImageConfig.resnet18()
        self.tcfg.prefix, self.ccfg.prefix = "text_model.model", "caption_text_model.model"
        self.icfg.prefix = "image_model.image_model"
        with torch.cuda.device(self.device):
            st = ParamStore(self.device)
            self.store = st
            self.text = TextTower(self.tcfg, st)
            self.caption = TextTower(self.ccfg, st)
            self.img = ImageTower(self.icfg, st)
            assert self.img.out_dim == 512, "fine_tune is hard-wired to 512 features in the reference (:571)"
            for t in (self.text, self.caption, self.img):
                t.register_noshadow()
            for name, n_out, _ in _LINEAR:
                st.add(f"{name}.bias", (n_out,), shadow=False)
            for name, c in _BN1D:
                st.add(f"{name}.weight", (c,), shadow=False)
                st.add(f"{name}.bias", (c,), shadow=False)
            st.add("output_fc.0.weight", (1, 512), shadow=False)     # read in fp32 by the fused output kernel
            st.add("output_fc.0.bias", (1,), shadow=False)
            for t in (self.text, self.caption, self.img):
                t.register_shadowed()
            for name, n_out, n_in in _LINEAR:
                n_in = n_in or (self.tcfg.dim if name.startswith("text_fc") else self.ccfg.dim)
                st.add(f"{name}.weight", (n_out, n_in))
            st.finalize()
            for t in (self.text, self.caption, self.img):
                t.bind()
            # BatchNorm1d running statistics (buffers, outside the optimizer's flat storage)
            self.bn_buffers = {}
            for name, c in _BN1D:
                self.bn_buffers[name] = (torch.zeros(c, device=self.device), torch.ones(c, device=self.device))
            if init:
                self.reset_parameters()
        self._param_names = st.names()
        self._params = nn.ParameterList([nn.Parameter(st.p(n), requires_grad=True) for n in self._param_names])
        st.owner = self
        for n, p in zip(self._param_names, self._params):
            spec = st.specs[n]
            p._b200mm_store, p._b200mm_offset = st, spec.offset
            p._b200mm_padded = (spec.numel + 63) // 64 * 64
        self._anchor = torch.zeros((), device=self.device, requires_grad=True)
        self._attach_grads()
        self._saved = None

    # ------------------------------------------------------------------ init / state dict
    @torch.no_grad()
    def reset_parameters(self):
        g = torch.Generator(device=self.device)
        g.manual_seed(self.seed)
        for t in (self.text, self.caption, self.img):
            t.init_parameters(g)
        st = self.store
        lin = [(n, st.specs[f"{n}.weight"].shape[1]) for n, _, _ in _LINEAR] + [("output_fc.0", 512)]
        for name, fan_in in lin:
            bound = 1.0 / (fan_in ** 0.5)       # nn.Linear default init
            st.p(f"{name}.weight").uniform_(-bound, bound, generator=g)
            st.p(f"{name}.bias").uniform_(-bound, bound, generator=g)
        for name, _ in _BN1D:
            st.p(f"{name}.weight").fill_(1.0)
            st.p(f"{name}.bias").zero_()
        st.refresh_shadow()

    @torch.no_grad()
    def load_reference_state_dict(self, sd: dict):
        used = super().load_reference_state_dict(sd)
        for name, (rm, rv) in self.bn_buffers.items():
            rm.copy_(sd[f"{name}.running_mean"])
            rv.copy_(sd[f"{name}.running_var"])
        return used

    @torch.no_grad()
    def reference_state_dict(self, _grads: bool = False) -> dict:
        out = super().reference_state_dict(_grads)
        if not _grads:
            for name, (rm, rv) in self.bn_buffers.items():
                out[f"{name}.running_mean"], out[f"{name}.running_var"] = rm.clone(), rv.clone()
        return out

    def get_params(self, lr):
        """Reference :645-664 (same substring rules, same group order)."""
        from .loop_head import param_group_index
        groups = ([], [], [])
        for name, param in self.named_parameters():
            groups[param_group_index(name)].append(param)
        return [{"params": groups[0], "lr": lr}, {"params": groups[1], "lr": lr * 0.8},
                {"params": groups[2], "lr": lr * 0.8}]

    def enable_data_parallel(self, group=None, bucket_elems: int = 64 * 1024 * 1024, payload: str = "bf16"):
        from .ddp import GradSync
        is_text = lambda n: n.startswith("text_model.") or n.startswith("caption_text_model.")   # noqa: E731
        self.grad_sync = GradSync(self.store, group, bucket_elems, payload=payload,
                                  phases=[("text", is_text)] + self.img.grad_phases() + [("rest", lambda n: True)])
        bufs = [self.img.buffers] + [t for pair in self.bn_buffers.values() for t in pair]
        self.grad_sync.broadcast_parameters(bufs)
        self.store.refresh_shadow()
        return self.grad_sync

    # ------------------------------------------------------------------ forward
    def _bn1d(self, name, x, training, out=None):
        st = self.store
        rm, rv = self.bn_buffers[name]
        return ops.bn1d_fwd(x, st.p(f"{name}.weight"), st.p(f"{name}.bias"), rm, rv, relu=True, train=training,
                            out=out)

    def _text_branch(self, tower, ids, mask, fc, cat_slice, training, seed, site):
        st = self.store
        B, S = ids.shape
        h = tower.forward(ids, mask, training=training, seed=seed, step=self._step)
        pd = self.head_dropout if training else 0.0
        s_drop = _mix(seed, self._step, 254, site)
        pooled = ops.gather_rows(h, B, S, 0, p_drop=pd, seed=s_drop)                 # dropout(h[:, 0]), :667-668
        lin = ops.linear_fwd(pooled, st.s(f"{fc}.0.weight"), st.p(f"{fc}.0.bias"))
        _, mean, rstd = self._bn1d(f"{fc}.1", lin, training, out=cat_slice)             # BN + ReLU, :599-601
        return (B, S, pd, s_drop, pooled, lin, mean, rstd)

    def _features(self, text, image, mask, caption_text, caption_text_mask, training):
        st = self.store
        B = text.shape[0]
        cat = torch.empty(B, 1536, device=self.device, dtype=torch.bfloat16)            # (text | image | caption), :493
        sv_t = self._text_branch(self.text, text, mask, "text_fc", cat[:, :512], training, self.seed, 0)
        sv_c = self._text_branch(self.caption, caption_text, caption_text_mask, "caption_text_fc", cat[:, 1024:],
                                 training, self.seed + 1, 1)
        feat = self.img.forward(image, training=training)                               # [B, 512]
        f1 = ops.linear_fwd(feat, st.s("image_model.fine_tune.0.weight"), st.p("image_model.fine_tune.0.bias"),
                            relu=True)
        pi = self.image_dropout if training else 0.0
        s_img = _mix(self.seed, self._step, 253, 0)
        f1d = ops.gather_rows(f1, B, 1, 0, p_drop=pi, seed=s_img) if pi > 0 else f1     # Dropout(.35), :573
        ops.linear_fwd(f1d, st.s("image_model.fine_tune.3.weight"), st.p("image_model.fine_tune.3.bias"),
                       out=cat[:, 512:1024])
        a0 = "fusion_layer.attention_layer"
        z_att = ops.linear_fwd(cat, st.s(f"{a0}.0.weight"), st.p(f"{a0}.0.bias"))
        a, m_a, r_a = self._bn1d(f"{a0}.1", z_att, training)                            # Linear-BN-ReLU, :479-483
        y, w = ops.softmax_gate_fwd(a, cat)                                             # Softmax(dim=1) * x, :484, :496
        # fp32 result: y carries the signal at 1/1536 scale, the bias would swamp it at bf16 resolution and the
        # BatchNorm behind it would amplify the rounding noise
        r_lin = ops.linear_fwd_f32(y, st.s("fusion_layer.reduce.0.weight"), st.p("fusion_layer.reduce.0.bias"))
        fused, m_r, r_r = self._bn1d("fusion_layer.reduce.1", r_lin, training)          # :486-490
        if training:
            self._saved = (B, cat, sv_t, sv_c, feat, f1, f1d, pi, s_img, z_att, a, m_a, r_a, y, w, r_lin, m_r, r_r,
                           fused)
        return fused

    def _output(self, fused, labels, training, *, alpha=0.25, gamma=2.0, dlogits=None, bn_train=None):
        st = self.store
        rm, rv = self.bn_buffers["output_fc.1"]
        g = st.g if training else (lambda n: None)
        return ops.head_bn_focal(fused, st.p("output_fc.0.weight"), st.p("output_fc.0.bias"), st.p("output_fc.1.weight"),
                                 st.p("output_fc.1.bias"), rm, rv, labels, alpha=alpha, gamma=gamma, train=training,
                                 bn_train=self.training if bn_train is None else bn_train, dlogits=dlogits,
                                 dW=g("output_fc.0.weight"), dbias=g("output_fc.0.bias"), dg=g("output_fc.1.weight"),
                                 dbeta=g("output_fc.1.bias"))

    def _prep(self):
        if not getattr(self, "_shadow_fresh", False):
            self.store.refresh_shadow()
        self._shadow_fresh = False

    def _engine_forward(self, text, image, mask, caption_text, caption_text_mask, training):
        if text.device != self.device:
            raise _lib.B200MMError("inputs must already be on the model's CUDA device (the loop does .to(device))")
        with torch.no_grad():
            self._prep()
            fused = self._features(text, image, mask, caption_text, caption_text_mask, training)
            # logits only (no loss): BatchNorm1d(1) uses batch statistics in train mode (and updates the running ones)
            logits, _, _, _ = self._output(fused, None, False, bn_train=training)
            if training:
                self._step += 1
        return logits

    # ------------------------------------------------------------------ backward
    def _text_branch_bwd(self, tower, sv, fc, d_slice, cat_slice):
        st = self.store
        B, S, pd, s_drop, pooled, lin, mean, rstd = sv
        d_lin = ops.bn1d_bwd(d_slice, cat_slice, lin, mean, rstd, st.p(f"{fc}.1.weight"), st.g(f"{fc}.1.weight"),
                             st.g(f"{fc}.1.bias"), relu=True)
        ops.linear_wgrad(d_lin, pooled, st.g(f"{fc}.0.weight"))
        ops.colsum(d_lin, st.g(f"{fc}.0.bias"))
        d_pooled = ops.linear_dgrad(d_lin, st.s(f"{fc}.0.weight"))
        tower.backward(ops.scatter_rows(d_pooled, B * S, S, 0, p_drop=pd, seed=s_drop))

    def _backward_from_dfused(self, dfused):
        st = self.store
        (B, cat, sv_t, sv_c, feat, f1, f1d, pi, s_img, z_att, a, m_a, r_a, y, w, r_lin, m_r, r_r, fused) = self._saved
        r0, a0 = "fusion_layer.reduce", "fusion_layer.attention_layer"
        d_rlin = ops.bn1d_bwd(dfused, fused, r_lin, m_r, r_r, st.p(f"{r0}.1.weight"), st.g(f"{r0}.1.weight"),
                              st.g(f"{r0}.1.bias"), relu=True)
        ops.linear_wgrad(d_rlin, y, st.g(f"{r0}.0.weight"))
        ops.colsum(d_rlin, st.g(f"{r0}.0.bias"))
        dy = ops.linear_dgrad(d_rlin, st.s(f"{r0}.0.weight"))
        da, dcat_direct = ops.softmax_gate_bwd(dy, w, cat)
        d_zatt = ops.bn1d_bwd(da, a, z_att, m_a, r_a, st.p(f"{a0}.1.weight"), st.g(f"{a0}.1.weight"),
                              st.g(f"{a0}.1.bias"), relu=True)
        ops.linear_wgrad(d_zatt, cat, st.g(f"{a0}.0.weight"))
        ops.colsum(d_zatt, st.g(f"{a0}.0.bias"))
        dcat = ops.linear_dgrad(d_zatt, st.s(f"{a0}.0.weight"), residual=dcat_direct)       # [B, 1536]
        self._text_branch_bwd(self.text, sv_t, "text_fc", dcat[:, :512], cat[:, :512])
        self._text_branch_bwd(self.caption, sv_c, "caption_text_fc", dcat[:, 1024:], cat[:, 1024:])
        sync = getattr(self, "grad_sync", None)
        if sync is not None:
            sync.ready("text")
        d_i = dcat[:, 512:1024]
        ops.linear_wgrad(d_i, f1d, st.g("image_model.fine_tune.3.weight"))
        ops.colsum(d_i, st.g("image_model.fine_tune.3.bias"))
        d_f1 = ops.linear_dgrad(d_i, st.s("image_model.fine_tune.3.weight"))
        if pi > 0:
            d_f1 = ops.scatter_rows(d_f1, B, 1, 0, p_drop=pi, seed=s_img)                      # dropout backward
        d_f1 = ops.relu_bwd(d_f1, f1)
        ops.linear_wgrad(d_f1, feat, st.g("image_model.fine_tune.0.weight"))
        ops.colsum(d_f1, st.g("image_model.fine_tune.0.bias"))
        self.img.backward(ops.linear_dgrad(d_f1, st.s("image_model.fine_tune.0.weight")),
                          on_grads_ready=sync.ready if sync is not None else None)
        if sync is not None:
            sync.ready("rest")
            if not getattr(self, "_defer_grad_sync", False):
                sync.finish_into_grad()
        self._saved = None
        self._attach_grads()

    def _engine_backward_from_dlogits(self, dlogits):
        with torch.no_grad():
            fused = self._saved[-1]
            # BatchNorm1d(1) statistics were already folded into the running buffers by the forward: keep them
            rm, rv = self.bn_buffers["output_fc.1"]
            keep = (rm.clone(), rv.clone())
            _, _, _, dfused = self._output(fused, None, True, dlogits=dlogits.reshape(-1).float().contiguous(),
                                           bn_train=True)
            rm.copy_(keep[0])
            rv.copy_(keep[1])
            self._backward_from_dfused(dfused)

    # ------------------------------------------------------------------ public call contract (reference :666-685)
    def forward(self, text=None, image=None, mask=None, caption_text=None, caption_text_mask=None):
        if any(t is None for t in (text, image, mask, caption_text, caption_text_mask)):
            raise TypeError("forward(text, image, mask, caption_text, caption_text_mask)")
        if self.training and torch.is_grad_enabled():
            return _EngineFunction.apply(self._anchor, self, text, image, mask, caption_text, caption_text_mask)
        return self._engine_forward(text, image, mask, caption_text, caption_text_mask, training=False)

    def train_step_fused(self, text, image, mask, caption_text, caption_text_mask, labels, *, alpha=0.25, gamma=2.0,
                         **_):
        """forward + sigmoid focal loss + backward, output_fc / BatchNorm1d(1) / loss fused in one kernel.
        Returns (logits fp32 [B], loss fp32 [1] (mean), correct int32 [1])."""
        with torch.no_grad():
            self._prep()
            fused = self._features(text, image, mask, caption_text, caption_text_mask, True)
            logits, loss, correct, dfused = self._output(fused, labels.contiguous(), True, alpha=alpha, gamma=gamma,
                                                         bn_train=True)
            self._step += 1
            self._backward_from_dfused(dfused)
        return logits, loss, correct

    def eval_step_fused(self, text, image, mask, caption_text, caption_text_mask, labels, *, alpha=0.25, gamma=2.0,
                        **_):
        with torch.no_grad():
            if not getattr(self, "_shadow_fresh", False):
                self.store.refresh_shadow()
            fused = self._features(text, image, mask, caption_text, caption_text_mask, False)
            logits, loss, correct, _ = self._output(fused, labels.contiguous(), False, alpha=alpha, gamma=gamma,
                                                    bn_train=False)
        return logits, loss, correct
