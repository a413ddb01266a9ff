"""``MultimodalClassifier``: the drop-in for the reference's late-fusion module, running on the sm_100a engine.

Same constructor, attribute vocabulary and call contract as the reference class
(example_scripts/Multimodal_example_task2C.txt:152-197):

    model = MultimodalClassifier(num_classes=2); model.to(device)
    output = model(text, image, mask)            # -> fp32 logits [B, num_classes]
    loss = criterion(output, labels); loss.backward(); optimizer.step()

``forward(input_ids, attention_mask, pixel_values)`` keyword aliases (BASELINE.json north_star) are accepted
too.  ``loss.backward()`` reaches the hand-written backward through a single autograd node; parameter
gradients land in the flat gradient buffer and are exposed as ``param.grad`` views, so both
``torch.optim.Adam(model.parameters())`` (the reference's optimizer) and ``b200mm.FusedAdam`` work.

There is no CPU path: constructing the model without a CUDA device, or with libb200mm.so missing, raises.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import _lib, ops
from .image_tower import ImageConfig, ImageTower
from .params import ParamStore
from .text_tower import TextConfig, TextTower, _mix
from .vit_tower import ViTConfig, ViTTower

POOL_LAST, POOL_CLS = "last", "cls"


_TOWER_OVERLAP = os.environ.get("B200MM_TOWER_OVERLAP", "1") != "0"

class _EngineFunction(torch.autograd.Function):
    """Single autograd node: forward = engine forward, backward = engine backward (grads written in place)."""

    @staticmethod
    def forward(ctx, anchor, model, *inputs):
        ctx.model, ctx.n_inputs = model, len(inputs)
        logits = model._engine_forward(*inputs, training=True)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        ctx.model._engine_backward_from_dlogits(dlogits.reshape(dlogits.shape[0], -1).contiguous())
        return (torch.zeros((), device=dlogits.device), None) + (None,) * ctx.n_inputs


class MultimodalClassifier(nn.Module):
    def __init__(self, num_classes: int = 2, *, text_config: TextConfig | None = None,
                 image_config: ImageConfig | ViTConfig | None = None, device=None, head_dropout: float = 0.3,
                 pooling: str | None = None, seed: int = 42, init: bool = True, squeeze_output: bool = False):
        super().__init__()
        if not torch.cuda.is_available():
            raise _lib.B200MMError("b200mm needs a CUDA device (sm_100a); there is no CPU fallback")
        _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.num_classes = num_classes
        self.head_dropout = head_dropout       # self.bert_drop = nn.Dropout(0.3), .txt:160
        self.tcfg = text_config or TextConfig()
        self.icfg = image_config or ImageConfig()
        if pooling is None:     # 'last' = bert_output[0][:, -1, :], .txt:178 ; 'cls' = HEAD script (.py:359-360)
            pooling = POOL_LAST if self.tcfg.arch == "distilbert" else POOL_CLS
        if pooling not in (POOL_LAST, POOL_CLS):
            raise ValueError(f"Unsupported pooling method: {pooling}")     # HEAD script :352
        self.pooling = pooling
        self.squeeze_output = squeeze_output   # single-logit head: return [B] like output.squeeze_(1) (HEAD :683)
        self.seed = seed
        self._step = 0
        with torch.cuda.device(self.device):
            st = ParamStore(self.device)
            self.store = st
            self.text = TextTower(self.tcfg, st)
            self.img = ViTTower(self.icfg, st) if self.icfg.arch == "vit" else ImageTower(self.icfg, st)
            D = self.tcfg.dim
            # --- shadow-less (fp32-read) parameters first
            self.text.register_noshadow()
            self.img.register_noshadow()
            for name, n in (("bert_fc.bias", 512), ("resnet_fc.bias", 512), ("fusion_fc.bias", 512),
                            ("output_fc.weight", num_classes * 512), ("output_fc.bias", num_classes)):
                st.add(name, (num_classes, 512) if name == "output_fc.weight" else (n,), shadow=False)
            # --- bf16-shadowed GEMM weights
            self.text.register_shadowed()
            self.img.register_shadowed()
            st.add("bert_fc.weight", (512, D))                          # .txt:161
            st.add("resnet_fc.weight", (512, self.img.out_dim))         # .txt:165 (1000 for the ResNet, dim for ViT)
            st.add("fusion_fc.weight", (512, 1024))                     # .txt:168
            st.finalize()
            self.text.bind()
            self.img.bind()
            if init:
                self.reset_parameters()
        # expose parameters under the reference's names (dots are not allowed in register_parameter)
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
        self.last_aux = None

    # ------------------------------------------------------------------ nn.Module plumbing
    def named_parameters(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True):
        for n, p in zip(self._param_names, self._params):
            yield (prefix + ("." if prefix else "") + n, p)

    def parameters(self, recurse: bool = True):
        return iter(self._params)

    def to(self, *args, **kwargs):
        dev = None
        for a in args:
            if isinstance(a, (str, torch.device)):
                dev = torch.device(a)
        dev = kwargs.get("device", dev)
        if dev is not None:
            dev = torch.device(dev)
            if dev.type != "cuda" or (dev.index is not None and dev.index != self.device.index):
                raise _lib.B200MMError(f"model lives on {self.device}; b200mm cannot move it to {dev}")
        return self

    def cuda(self, device=None):
        return self

    def _attach_grads(self):
        for n, p in zip(self._param_names, self._params):
            p.grad = self.store.g(n)

    def zero_grad(self, set_to_none: bool = False):
        self.store.zero_grad()
        self._attach_grads()

    @torch.no_grad()
    def reset_parameters(self):
        g = torch.Generator(device=self.device)
        g.manual_seed(self.seed)
        self.text.init_parameters(g)
        self.img.init_parameters(g)
        st = self.store
        for name, fan_in in (("bert_fc", self.tcfg.dim), ("resnet_fc", self.img.out_dim), ("fusion_fc", 1024),
                             ("output_fc", 512)):
            bound = 1.0 / (fan_in ** 0.5)   # nn.Linear default init
            st.p(f"{name}.weight").uniform_(-bound, bound, generator=g)
            st.p(f"{name}.bias").uniform_(-bound, bound, generator=g)
        st.refresh_shadow()

    # ------------------------------------------------------------------ state-dict exchange with the reference module
    @torch.no_grad()
    def load_reference_state_dict(self, sd: dict):
        """Copy weights from the reference/oracle module's ``state_dict()`` (torch layouts) into the engine layout:
        conv OIHW -> OHWI-flattened (K zero-padded to a multiple of 8); everything else is a straight copy.  Keys
        the engine does not hold (BERT ``pooler.*``, which the path never reads; ``position_ids`` buffers) are
        ignored."""
        st = self.store
        used = set()
        img_prefix = self.icfg.prefix + "."
        for name in st.names():
            dst = st.p(name)
            src = sd[name].to(self.device, torch.float32)
            used.add(name)
            if name.startswith(img_prefix):
                self.img.import_param(name, src, dst)
            else:
                dst.copy_(src.view(dst.shape))
        self.img.load_buffers(sd)
        st.refresh_shadow()
        return used

    @torch.no_grad()
    def reference_grad_dict(self) -> dict:
        """Gradients in the reference module's parameter layouts (for parity checks against autograd)."""
        return self.reference_state_dict(_grads=True)

    @torch.no_grad()
    def reference_state_dict(self, _grads: bool = False) -> dict:
        """Inverse of ``load_reference_state_dict``: a state dict the reference module can ``load_state_dict``
        (``strict=False`` for BERT-family towers, whose unused pooler the engine does not carry)."""
        st = self.store
        out = {}
        img_prefix = self.icfg.prefix + "."
        for name in st.names():
            t = (st.g(name) if _grads else st.p(name)).detach().clone()
            out[name] = self.img.export_param(name, t) if name.startswith(img_prefix) else t
        if not _grads:
            self.img.export_buffers(out)
        return out

    def state_dict(self, *args, **kwargs):
        return self.reference_state_dict()

    def load_state_dict(self, sd, strict: bool = True):
        self.load_reference_state_dict(sd)

    # ------------------------------------------------------------------ engine forward / backward
    def _features(self, text, image, mask, training):
        """Towers + fusion layers up to the 512-d fused feature (input of output_fc)."""
        st, S = self.store, text.shape[1]
        B = text.shape[0]
        # The two towers are independent until the concat: the image tower is issued on a side stream (a parallel branch
        # of the step's CUDA graph), so its HBM-bound BatchNorm / pooling kernels share the SMs with the text tower's
        # tensor-bound GEMMs instead of alternating with them (B200MM_TOWER_OVERLAP=0: one stream)
        main, side = self._tower_streams()
        if side is not None:
            side.wait_stream(main)
            image.record_stream(side)
            with torch.cuda.stream(side):
                r1000 = self.img.forward(image, training=training, seed=self.seed, step=self._step)
            r1000.record_stream(main)
        h = self.text.forward(text, mask, training=training, seed=self.seed, step=self._step)      # [B*S, D]
        off = S - 1 if self.pooling == POOL_LAST else 0
        s_head = _mix(self.seed, self._step, 254, 0)
        pd = self.head_dropout if training else 0.0
        pooled = ops.gather_rows(h, B, S, off, p_drop=pd, seed=s_head)                             # bert_drop(h[:, -1])
        if side is not None:
            main.wait_stream(side)
        else:
            r1000 = self.img.forward(image, training=training, seed=self.seed, step=self._step)    # [B, 1000 | dim]
        cat = torch.empty(B, 1024, device=self.device, dtype=torch.bfloat16)
        ops.linear_fwd(pooled, st.s("bert_fc.weight"), st.p("bert_fc.bias"), out=cat[:, :512])      # .txt:179
        ops.linear_fwd(r1000, st.s("resnet_fc.weight"), st.p("resnet_fc.bias"), out=cat[:, 512:])   # .txt:184, 190
        fused = ops.linear_fwd(cat, st.s("fusion_fc.weight"), st.p("fusion_fc.bias"))               # .txt:193
        if training:
            self._saved = (B, S, off, pd, s_head, pooled, r1000, cat, fused)
        return fused

    def _tower_streams(self):
        """(current stream, side stream for the image tower) -- side is None when the overlap is switched off."""
        main = torch.cuda.current_stream(self.device)
        if not _TOWER_OVERLAP:
            return main, None
        side = getattr(self, "_side_stream", None)
        if side is None:
            side = self._side_stream = torch.cuda.Stream(device=self.device)
        return main, side

    def _engine_forward(self, text, image, mask, training):
        st = self.store
        if text.device != self.device:
            raise _lib.B200MMError("inputs must already be on the model's CUDA device (the loop does .to(device))")
        with torch.no_grad():
            if not getattr(self, "_shadow_fresh", False):
                st.refresh_shadow()
            self._shadow_fresh = False
            fused = self._features(text, image, mask, training)
            logits, _, _, _ = ops.head_loss(fused, st.p("output_fc.weight"), st.p("output_fc.bias"), None, train=False)
            if training:
                self._step += 1
        return logits

    def _backward_from_dfused(self, dfused):
        st = self.store
        B, S, off, pd, s_head, pooled, r1000, cat, fused = self._saved
        ops.linear_wgrad(dfused, cat, st.g("fusion_fc.weight"))
        ops.colsum(dfused, st.g("fusion_fc.bias"))
        dcat = ops.linear_dgrad(dfused, st.s("fusion_fc.weight"))                                   # [B, 1024]
        d_t, d_r = dcat[:, :512], dcat[:, 512:]
        ops.linear_wgrad(d_t, pooled, st.g("bert_fc.weight"))
        ops.colsum(d_t, st.g("bert_fc.bias"))
        ops.linear_wgrad(d_r, r1000, st.g("resnet_fc.weight"))
        ops.colsum(d_r, st.g("resnet_fc.bias"))
        d_pooled = ops.linear_dgrad(d_t, st.s("bert_fc.weight"))
        d_r1000 = ops.linear_dgrad(d_r, st.s("resnet_fc.weight"))
        dh = ops.scatter_rows(d_pooled, B * S, S, off, p_drop=pd, seed=s_head)
        # data parallel: the towers announce each parameter group whose gradients are final (ddp.GradSync.ready), so
        # its all-reduce runs on NCCL's stream under the rest of the backward
        sync = getattr(self, "grad_sync", None)
        cb = sync.ready if sync is not None else None
        main, side = self._tower_streams()
        if side is not None:
            side.wait_stream(main)
            d_r1000.record_stream(side)
            with torch.cuda.stream(side):
                self.img.backward(d_r1000, on_grads_ready=cb)
            self.text.backward(dh, on_grads_ready=cb)
            main.wait_stream(side)
        else:
            self.text.backward(dh, on_grads_ready=cb)
            self.img.backward(d_r1000, on_grads_ready=cb)
        if sync is not None:
            sync.ready("rest")
            if not getattr(self, "_defer_grad_sync", False):
                sync.finish_into_grad()      # an optimizer that reads param.grad (torch.optim.*) needs the mean NOW
        self._saved = None
        self._attach_grads()

    def _engine_backward_from_dlogits(self, dlogits):
        st = self.store
        with torch.no_grad():
            fused = self._saved[-1]
            _, _, _, dfused = ops.head_loss(fused, st.p("output_fc.weight"), st.p("output_fc.bias"), None,
                                            loss_kind=ops.LOSS_EXTERNAL, dW=st.g("output_fc.weight"),
                                            dbias=st.g("output_fc.bias"), dlogits=dlogits.float())
            self._backward_from_dfused(dfused)

    def enable_data_parallel(self, group=None, bucket_elems: int = 64 * 1024 * 1024, payload: str = "bf16"):
        """One process per GPU: broadcast rank 0's parameters / BN buffers and all-reduce gradients every step, in
        phases that follow the order in which the backward finishes them (ddp.py)."""
        from .ddp import GradSync
        phases = self.text.grad_phases() + self.img.grad_phases() + [("rest", lambda n: True)]
        self.grad_sync = GradSync(self.store, group, bucket_elems, phases=phases, payload=payload)
        self.grad_sync.broadcast_parameters([self.img.buffers] if self.img.buffers is not None else [])
        self.store.refresh_shadow()
        return self.grad_sync

    # ------------------------------------------------------------------ public call contract
    def forward(self, text=None, image=None, mask=None, *, input_ids=None, attention_mask=None, pixel_values=None):
        text = input_ids if text is None else text
        image = pixel_values if image is None else image
        mask = attention_mask if mask is None else mask
        if text is None or image is None or mask is None:
            raise TypeError("forward(text, image, mask) / forward(input_ids=, attention_mask=, pixel_values=)")
        if self.training and torch.is_grad_enabled():
            out = _EngineFunction.apply(self._anchor, self, text, image, mask)
        else:
            out = self._engine_forward(text, image, mask, training=False)
        return out.squeeze(1) if (self.squeeze_output and self.num_classes == 1) else out

    def eval_step_fused(self, text, image, mask, labels, *, loss_kind=ops.LOSS_CE, alpha=0.25, gamma=2.0):
        """Eval-mode forward with output_fc + loss + accuracy count fused. Returns (logits, loss_sum, correct)."""
        st = self.store
        with torch.no_grad():
            if not getattr(self, "_shadow_fresh", False):
                st.refresh_shadow()
            fused = self._features(text, image, mask, False)
            logits, loss, correct, _ = ops.head_loss(fused, st.p("output_fc.weight"), st.p("output_fc.bias"),
                                                     labels.contiguous(), loss_kind=loss_kind, alpha=alpha,
                                                     gamma=gamma, train=False)
        return logits, loss, correct

    # ------------------------------------------------------------------ fused step (used by b200mm.loop.train)
    def train_step_fused(self, text, image, mask, labels, *, loss_kind=ops.LOSS_CE, alpha=0.25, gamma=2.0):
        """forward + loss + backward in one go, with output_fc + loss fused in one kernel.
        Returns (logits fp32 [B,C], loss_sum fp32 [1], correct int32 [1]) -- all still on the device."""
        st = self.store
        with torch.no_grad():
            if not getattr(self, "_shadow_fresh", False):
                st.refresh_shadow()
            self._shadow_fresh = False
            fused = self._features(text, image, mask, True)
            logits, loss, correct, dfused = ops.head_loss(
                fused, st.p("output_fc.weight"), st.p("output_fc.bias"), labels.contiguous(), loss_kind=loss_kind,
                alpha=alpha, gamma=gamma, train=True, dW=st.g("output_fc.weight"), dbias=st.g("output_fc.bias"))
            self._step += 1
            self._backward_from_dfused(dfused)
        return logits, loss, correct
