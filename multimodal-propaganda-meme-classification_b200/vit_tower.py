"""Image tower, ViT flavour: patch-embed conv-as-GEMM + pre-LN transformer encoder with explicit forward / backward
on the sm_100a kernels (BASELINE.json configs 3-5: ViT-B/16, ViT-L/14).

The reference only *names* ViT image towers (``vit_base_patch16_224``, example_scripts/Multimodal_example_task2C.py:82;
``ViTModel`` feature extractor, mm_model_mm_example_task2C.py:66-67), it never runs one on the 2C path, so the
architecture is taken from the library it would have called: transformers' ``ViTModel(add_pooling_layer=False)``
(transformers/models/vit/modeling_vit.py: ViTEmbeddings, ViTLayer, final ``layernorm``); the pooled feature is the
CLS token of ``last_hidden_state`` -- what timm's ``reset_classifier(0)`` model returns as well.  Parity for this
tower is pinned by the in-repo oracle only.

    cols  = patches of the fp32 NCHW image as a bf16 [B*P, 3*p*p] matrix         im2col_nchw_f32 (stride = kernel)
    patch = cols Wp^T + bp                                                       tcgen05 GEMM   
    x     = [cls | patch] + pos                                                  vit_assemble_fwd
    per layer (pre-LN):
      qkv = LN1(x) Wqkv^T + b ; ctx = softmax(q k^T / 8) v                       LN, GEMM, tcgen05 attention (no mask)
      x2  = ctx Wo^T + bo + x                                                    GEMM (+residual epilogue)
      x3  = gelu(LN2(x2) W1^T + b1) W2^T + b2 + x2                               LN, GEMM (+GELU), GEMM (+residual)
    feat  = LN_f(x[:, 0])                                                        gather_rows, LN  (only the CLS row is used)
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import ops
from .params import ParamStore
from .text_tower import _mix


@dataclass
class ViTConfig:
    image_size: int = 224
    patch_size: int = 16
    dim: int = 768
    n_layers: int = 12
    n_heads: int = 12
    hidden_dim: int = 3072
    layer_norm_eps: float = 1e-12
    attention_dropout: float = 0.0   # ViTConfig.attention_probs_dropout_prob default
    prefix: str = "resnet"           # the reference's attribute name for the image tower (.txt:164)
    arch: str = "vit"

    @staticmethod
    def vit_b16(**kw) -> "ViTConfig":
        return ViTConfig(**kw)

    @staticmethod
    def vit_l14(**kw) -> "ViTConfig":
        return ViTConfig(patch_size=14, dim=1024, n_layers=24, n_heads=16, hidden_dim=4096, **kw)

    @property
    def num_patches(self) -> int:
        return (self.image_size // self.patch_size) ** 2

    @property
    def patch_k(self) -> int:
        return 3 * self.patch_size * self.patch_size

    @property
    def patch_kp(self) -> int:
        return (self.patch_k + 7) // 8 * 8   # TMA rows must be multiples of 16 bytes (588 -> 592 for p = 14)


class ViTTower:
    def __init__(self, cfg: ViTConfig, store: ParamStore):
        assert cfg.dim == cfg.n_heads * 64, "attention kernel is specialised for head_dim 64"
        assert cfg.num_patches + 1 <= 512, "attention kernels cover up to 512 tokens"
        self.cfg = cfg
        self.store = store
        self.out_dim = cfg.dim
        self.buffers = None
        self._saved = None
        self.capture = None

    # ------------------------------------------------------------------ parameters (transformers ViTModel key names)
    def _layer_names(self, i: int) -> dict:
        L = f"{self.cfg.prefix}.encoder.layer.{i}"
        return {"q": f"{L}.attention.attention.query", "k": f"{L}.attention.attention.key",
                "v": f"{L}.attention.attention.value", "o": f"{L}.attention.output.dense",
                "ln1": f"{L}.layernorm_before", "ln2": f"{L}.layernorm_after",
                "f1": f"{L}.intermediate.dense", "f2": f"{L}.output.dense"}

    def register_noshadow(self):
        c, st, p = self.cfg, self.store, self.cfg.prefix
        st.add(f"{p}.embeddings.cls_token", (c.dim,), shadow=False)
        st.add(f"{p}.embeddings.position_embeddings", (c.num_patches + 1, c.dim), shadow=False)
        st.add(f"{p}.embeddings.patch_embeddings.projection.bias", (c.dim,), shadow=False)
        for i in range(c.n_layers):
            n = self._layer_names(i)
            for k in ("q", "k", "v"):
                st.add(f"{n[k]}.bias", (c.dim,), shadow=False)
            st.add(f"{n['o']}.bias", (c.dim,), shadow=False)
            st.add(f"{n['f1']}.bias", (c.hidden_dim,), shadow=False)
            st.add(f"{n['f2']}.bias", (c.dim,), shadow=False)
            for ln in ("ln1", "ln2"):
                st.add(f"{n[ln]}.weight", (c.dim,), shadow=False)
                st.add(f"{n[ln]}.bias", (c.dim,), shadow=False)
        st.add(f"{p}.layernorm.weight", (c.dim,), shadow=False)
        st.add(f"{p}.layernorm.bias", (c.dim,), shadow=False)

    def register_shadowed(self):
        c, st, p = self.cfg, self.store, self.cfg.prefix
        # conv weight stored OHWI-flattened [D, p*p*3] (K padded to a multiple of 8), like every conv of the engine
        st.add(f"{p}.embeddings.patch_embeddings.projection.weight", (c.dim, c.patch_kp))
        for i in range(c.n_layers):
            n = self._layer_names(i)
            for k in ("q", "k", "v"):
                st.add(f"{n[k]}.weight", (c.dim, c.dim))
            st.add(f"{n['o']}.weight", (c.dim, c.dim))
            st.add(f"{n['f1']}.weight", (c.hidden_dim, c.dim))
            st.add(f"{n['f2']}.weight", (c.dim, c.hidden_dim))

    def bind(self):
        c, st, p = self.cfg, self.store, self.cfg.prefix
        D = c.dim
        e = f"{p}.embeddings"
        self.cls, self.dcls = st.p(f"{e}.cls_token"), st.g(f"{e}.cls_token")
        self.pos, self.dpos = st.p(f"{e}.position_embeddings"), st.g(f"{e}.position_embeddings")
        pw = f"{e}.patch_embeddings.projection"
        self.wp, self.dwp = st.s(f"{pw}.weight"), st.g(f"{pw}.weight")
        self.bp, self.dbp = st.p(f"{pw}.bias"), st.g(f"{pw}.bias")
        self.gf, self.dgf = st.p(f"{p}.layernorm.weight"), st.g(f"{p}.layernorm.weight")
        self.bf, self.dbf = st.p(f"{p}.layernorm.bias"), st.g(f"{p}.layernorm.bias")
        self.layers = []
        for i in range(c.n_layers):
            n = self._layer_names(i)
            d = {}
            d["wqkv"] = st.span(st.shadow, f"{n['q']}.weight", f"{n['v']}.weight", (3 * D, D))
            d["dwqkv"] = st.span(st.grad, f"{n['q']}.weight", f"{n['v']}.weight", (3 * D, D))
            d["bqkv"] = st.span(st.master, f"{n['q']}.bias", f"{n['v']}.bias", (3 * D,))
            d["dbqkv"] = st.span(st.grad, f"{n['q']}.bias", f"{n['v']}.bias", (3 * D,))
            for short, name in (("wo", f"{n['o']}.weight"), ("w1", f"{n['f1']}.weight"), ("w2", f"{n['f2']}.weight")):
                d[short], d["d" + short] = st.s(name), st.g(name)
            for short, name in (("bo", f"{n['o']}.bias"), ("b1", f"{n['f1']}.bias"), ("b2", f"{n['f2']}.bias"),
                                ("g1", f"{n['ln1']}.weight"), ("be1", f"{n['ln1']}.bias"),
                                ("g2", f"{n['ln2']}.weight"), ("be2", f"{n['ln2']}.bias")):
                d[short], d["d" + short] = st.p(name), st.g(name)
            self.layers.append(d)

    def init_parameters(self, generator=None):
        """transformers' ViT init: trunc_normal(0, 0.02) weights / cls / pos, zero biases, LN = (1, 0)."""
        st, pre = self.store, self.cfg.prefix + "."
        for name in st.names():
            if not name.startswith(pre):
                continue
            t = st.p(name)
            if "layernorm" in name and name.endswith(".weight"):
                t.fill_(1.0)
            elif name.endswith(".bias"):
                t.zero_()
            else:
                t.normal_(0.0, 0.02, generator=generator).clamp_(-0.04, 0.04)
        pw = st.p(f"{self.cfg.prefix}.embeddings.patch_embeddings.projection.weight")
        pw[:, self.cfg.patch_k:].zero_()

    # ------------------------------------------------------------------ state-dict layout exchange (see model.py)
    def import_param(self, name: str, src: torch.Tensor, dst: torch.Tensor) -> None:
        if src.dim() == 4:      # Conv2d OIHW -> OHWI-flattened, K zero-padded
            co, ci, kh, kw = src.shape
            flat = src.permute(0, 2, 3, 1).reshape(co, kh * kw * ci)
            dst.zero_()
            dst[:, :flat.shape[1]].copy_(flat)
        else:
            dst.copy_(src.reshape(dst.shape))

    def export_param(self, name: str, t: torch.Tensor) -> torch.Tensor:
        c = self.cfg
        if name.endswith("patch_embeddings.projection.weight"):
            p = c.patch_size
            return t[:, :c.patch_k].reshape(c.dim, p, p, 3).permute(0, 3, 1, 2).contiguous()
        if name.endswith("cls_token"):
            return t.reshape(1, 1, c.dim)
        if name.endswith("position_embeddings"):
            return t.reshape(1, c.num_patches + 1, c.dim)
        return t

    def load_buffers(self, sd: dict) -> None:
        pass

    def export_buffers(self, out: dict) -> None:
        pass

    # ------------------------------------------------------------------ forward
    def forward(self, image: torch.Tensor, *, training: bool, seed: int = 0, step: int = 0):
        """image: fp32 NCHW [B, 3, H, W]. Returns the CLS feature after the final LayerNorm, bf16 [B, dim]."""
        c = self.cfg
        B, Cin, Hh, Ww = image.shape
        if Cin != 3 or Hh != c.image_size or Ww != c.image_size:
            raise ValueError(f"ViT tower expects [B, 3, {c.image_size}, {c.image_size}] images")
        P, T, H = c.num_patches, c.num_patches + 1, c.n_heads
        pa = c.attention_dropout if training else 0.0
        cols, _, _ = ops.im2col_nchw_f32(image, c.patch_size, c.patch_size, 0, c.patch_kp)      # [B*P, Kp]
        patch = ops.linear_fwd(cols, self.wp, self.bp)                                           # [B*P, D]
        x = ops.vit_assemble_fwd(patch, self.cls, self.pos, B, P)                                # [B*T, D]
        sv = {"B": B, "cols": cols, "layers": []} if training else None
        if self.capture is not None:
            self.capture.append(x)
        for li, L in enumerate(self.layers):
            s_att = _mix(seed, step, 128 + li, 1)
            h1, m1, r1 = ops.layernorm_fwd(x, L["g1"], L["be1"], c.layer_norm_eps)
            qkv = ops.linear_fwd(h1, L["wqkv"], L["bqkv"])
            ctx, lse = ops.attention_fwd(qkv, None, B, H, T, p_drop=pa, seed=s_att, need_lse=training)
            x2 = ops.linear_fwd(ctx, L["wo"], L["bo"], residual=x)
            h2, m2, r2 = ops.layernorm_fwd(x2, L["g2"], L["be2"], c.layer_norm_eps)
            z, a = ops.linear_gelu_fwd(h2, L["w1"], L["b1"])
            x3 = ops.linear_fwd(a, L["w2"], L["b2"], residual=x2)
            if training:
                sv["layers"].append((x, m1, r1, h1, qkv, ctx, lse, x2, m2, r2, h2, z, a, pa, s_att))
            x = x3
            if self.capture is not None:
                self.capture.append(x)
        cls_rows = ops.gather_rows(x, B, T, 0)                                                    # x[:, 0]
        feat, mf, rf = ops.layernorm_fwd(cls_rows, self.gf, self.bf, c.layer_norm_eps)
        if training:
            sv["tail"] = (cls_rows, mf, rf)
        self._saved = sv
        return feat

    # ------------------------------------------------------------------ data-parallel gradient phases
    GRAD_GROUPS = 4

    def _group_of(self, li: int) -> int:
        per = -(-self.cfg.n_layers // self.GRAD_GROUPS)
        return li // per

    def grad_phases(self):
        """Ordered (tag, predicate) list for ddp.GradSync: encoder layers in GRAD_GROUPS groups, top group first; the
        patch embedding, cls / position tables and the final LayerNorm fall to the model's catch-all last phase."""
        pre = self.cfg.prefix
        out = []
        for g in reversed(range(self._group_of(self.cfg.n_layers - 1) + 1)):
            stems = tuple(f"{pre}.encoder.layer.{i}." for i in range(self.cfg.n_layers) if self._group_of(i) == g)
            out.append((f"{pre}.g{g}", lambda n, stems=stems: n.startswith(stems)))
        return out

    # ------------------------------------------------------------------ backward
    def backward(self, dfeat: torch.Tensor, on_grads_ready=None):
        """dfeat: bf16 [B, dim]. Accumulates parameter gradients (the image itself needs none).
        on_grads_ready(tag): called as soon as every gradient of a ``grad_phases`` group is final."""
        sv = self._saved
        assert sv is not None, "backward() without a training-mode forward()"
        c = self.cfg
        B, P, T, H = sv["B"], c.num_patches, c.num_patches + 1, c.n_heads
        cls_rows, mf, rf = sv["tail"]
        d_cls, _ = ops.layernorm_bwd(dfeat, cls_rows, mf, rf, self.gf, self.dgf, self.dbf)
        dx = ops.scatter_rows(d_cls, B * T, T, 0)
        if getattr(self, "_wq", None) is None:
            self._wq = ops.SideQueue(dfeat.device)
        wq = self._wq       # weight / bias gradients run beside the data-gradient chain (ops.SideQueue)
        for li in reversed(range(len(self.layers))):
            L = self.layers[li]
            x, m1, r1, h1, qkv, ctx, lse, x2, m2, r2, h2, z, a, pa, s_att = sv["layers"][li]
            # x3 = a W2^T + b2 + x2
            wq.run(lambda: (ops.linear_wgrad(dx, a, L["dw2"]), ops.colsum(dx, L["db2"])), dx, a)
            dz = ops.linear_dgrad(dx, L["w2"], gelu_z=z, bias_grad=L["db1"])   # + column sums = fc1's bias gradient
            wq.run(lambda: ops.linear_wgrad(dz, h2, L["dw1"]), dz, h2)
            dh2 = ops.linear_dgrad(dz, L["w1"])
            dx2, _ = ops.layernorm_bwd(dh2, x2, m2, r2, L["g2"], L["dg2"], L["dbe2"], addend=dx)
            # x2 = ctx Wo^T + bo + x
            wq.run(lambda: (ops.linear_wgrad(dx2, ctx, L["dwo"]), ops.colsum(dx2, L["dbo"])), dx2, ctx)
            dctx = ops.linear_dgrad(dx2, L["wo"])
            dqkv = ops.attention_bwd(qkv, None, ctx, dctx, lse, B, H, T, p_drop=pa, seed=s_att)
            wq.run(lambda: (ops.linear_wgrad(dqkv, h1, L["dwqkv"]), ops.colsum(dqkv, L["dbqkv"])), dqkv, h1)
            dh1 = ops.linear_dgrad(dqkv, L["wqkv"])
            dx, _ = ops.layernorm_bwd(dh1, x, m1, r1, L["g1"], L["dg1"], L["dbe1"], addend=dx2)
            if on_grads_ready is not None and (li == 0 or self._group_of(li - 1) != self._group_of(li)):
                wq.join()
                on_grads_ready(f"{c.prefix}.g{self._group_of(li)}")
        dpatch = ops.vit_assemble_bwd(dx, self.dcls, self.dpos, B, P)
        ops.linear_wgrad(dpatch, sv["cols"], self.dwp)
        ops.colsum(dpatch, self.dbp)
        wq.join()
        self._saved = None
