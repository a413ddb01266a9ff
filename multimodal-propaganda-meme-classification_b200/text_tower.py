"""Text tower: post-LN BERT-family encoder (DistilBERT layout) with explicit forward / backward on the
sm_100a kernels.

Mirrors, op for op, what the reference's ``self.bert(text, attention_mask=mask)`` executes
(example_scripts/Multimodal_example_task2C.txt:158, :175 -> transformers DistilBertModel:
modeling_distilbert.py:83-122 embeddings, :126-207 attention, :210-228 FFN, :231-263 block):

    x   = dropout(LN(word[ids] + pos))                                   embed_layernorm_fwd
    per layer:
      qkv = x Wqkv^T + b            (q_lin | k_lin | v_lin fused, N = 3D)  tcgen05 GEMM
      ctx = softmax(q k^T / 8 + mask) v   (dropout on probs)              tcgen05 attention
      y   = LN(ctx Wo^T + b + x)                                          GEMM (+residual epilogue), LN
      a   = gelu(y W1^T + b)                                              GEMM (+bias+GELU epilogue)
      out = LN(dropout(a W2^T + b) + y)                                   GEMM (+dropout+residual), LN
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import ops
from .params import ParamStore


@dataclass
class TextConfig:
    vocab_size: int = 119547
    max_position_embeddings: int = 512
    dim: int = 768
    n_layers: int = 6
    n_heads: int = 12
    hidden_dim: int = 3072
    dropout: float = 0.1
    attention_dropout: float = 0.1
    layer_norm_eps: float = 1e-12
    pad_token_id: int = 0      # nn.Embedding(padding_idx=pad_token_id): no gradient for the pad row
    prefix: str = "bert"


def _mix(seed: int, step: int, layer: int, site: int) -> int:
    x = (seed * 0x9E3779B97F4A7C15 + step * 0xD1B54A32D192ED03 + layer * 0x100 + site + 1) & 0xFFFFFFFFFFFFFFFF
    x ^= x >> 31
    x = (x * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    x ^= x >> 29
    return x


class TextTower:
    def __init__(self, cfg: TextConfig, store: ParamStore):
        assert cfg.dim == cfg.n_heads * 64, "attention kernel is specialised for head_dim 64"
        self.cfg = cfg
        self.store = store
        self._saved = None
        self.capture = None   # set to a list to record the embedding output and every layer's output

    # ------------------------------------------------------------------ parameter registration (reference key names)
    def register_noshadow(self):
        c, st, p = self.cfg, self.store, self.cfg.prefix
        st.add(f"{p}.embeddings.word_embeddings.weight", (c.vocab_size, c.dim), shadow=False)
        st.add(f"{p}.embeddings.position_embeddings.weight", (c.max_position_embeddings, c.dim), shadow=False)
        st.add(f"{p}.embeddings.LayerNorm.weight", (c.dim,), shadow=False)
        st.add(f"{p}.embeddings.LayerNorm.bias", (c.dim,), shadow=False)
        for i in range(c.n_layers):
            L = f"{p}.transformer.layer.{i}"
            for n in ("q_lin", "k_lin", "v_lin"):
                st.add(f"{L}.attention.{n}.bias", (c.dim,), shadow=False)
            st.add(f"{L}.attention.out_lin.bias", (c.dim,), shadow=False)
            st.add(f"{L}.sa_layer_norm.weight", (c.dim,), shadow=False)
            st.add(f"{L}.sa_layer_norm.bias", (c.dim,), shadow=False)
            st.add(f"{L}.ffn.lin1.bias", (c.hidden_dim,), shadow=False)
            st.add(f"{L}.ffn.lin2.bias", (c.dim,), shadow=False)
            st.add(f"{L}.output_layer_norm.weight", (c.dim,), shadow=False)
            st.add(f"{L}.output_layer_norm.bias", (c.dim,), shadow=False)

    def register_shadowed(self):
        c, st, p = self.cfg, self.store, self.cfg.prefix
        for i in range(c.n_layers):
            L = f"{p}.transformer.layer.{i}"
            for n in ("q_lin", "k_lin", "v_lin"):
                st.add(f"{L}.attention.{n}.weight", (c.dim, c.dim))
            st.add(f"{L}.attention.out_lin.weight", (c.dim, c.dim))
            st.add(f"{L}.ffn.lin1.weight", (c.hidden_dim, c.dim))
            st.add(f"{L}.ffn.lin2.weight", (c.dim, c.hidden_dim))

    def bind(self):
        """Resolve views once the store is finalized."""
        c, st, p = self.cfg, self.store, self.cfg.prefix
        D = c.dim
        e = f"{p}.embeddings"
        self.word, self.dword = st.p(f"{e}.word_embeddings.weight"), st.g(f"{e}.word_embeddings.weight")
        self.pos, self.dpos = st.p(f"{e}.position_embeddings.weight"), st.g(f"{e}.position_embeddings.weight")
        self.eg, self.deg = st.p(f"{e}.LayerNorm.weight"), st.g(f"{e}.LayerNorm.weight")
        self.eb, self.deb = st.p(f"{e}.LayerNorm.bias"), st.g(f"{e}.LayerNorm.bias")
        self.layers = []
        for i in range(c.n_layers):
            L = f"{p}.transformer.layer.{i}"
            a = f"{L}.attention"
            d = {}
            d["wqkv"] = st.span(st.shadow, f"{a}.q_lin.weight", f"{a}.v_lin.weight", (3 * D, D))
            d["dwqkv"] = st.span(st.grad, f"{a}.q_lin.weight", f"{a}.v_lin.weight", (3 * D, D))
            d["bqkv"] = st.span(st.master, f"{a}.q_lin.bias", f"{a}.v_lin.bias", (3 * D,))
            d["dbqkv"] = st.span(st.grad, f"{a}.q_lin.bias", f"{a}.v_lin.bias", (3 * D,))
            for short, name in (("wo", f"{a}.out_lin.weight"), ("w1", f"{L}.ffn.lin1.weight"),
                                ("w2", f"{L}.ffn.lin2.weight")):
                d[short], d["d" + short] = st.s(name), st.g(name)
            for short, name in (("bo", f"{a}.out_lin.bias"), ("b1", f"{L}.ffn.lin1.bias"), ("b2", f"{L}.ffn.lin2.bias"),
                                ("g1", f"{L}.sa_layer_norm.weight"), ("be1", f"{L}.sa_layer_norm.bias"),
                                ("g2", f"{L}.output_layer_norm.weight"), ("be2", f"{L}.output_layer_norm.bias")):
                d[short], d["d" + short] = st.p(name), st.g(name)
            self.layers.append(d)

    def init_parameters(self, generator=None):
        """transformers' DistilBERT init: N(0, 0.02) for linear / embedding weights, zeros for biases, LN = (1, 0)."""
        st = self.store
        for name in st.names():
            if not name.startswith(self.cfg.prefix + "."):
                continue
            t = st.p(name)
            if "LayerNorm.weight" in name or "layer_norm.weight" in name:
                t.fill_(1.0)
            elif name.endswith(".bias"):
                t.zero_()
            else:
                t.normal_(0.0, 0.02, generator=generator)

    # ------------------------------------------------------------------ forward
    def forward(self, ids: torch.Tensor, mask: torch.Tensor, *, training: bool, seed: int = 0, step: int = 0):
        """ids, mask: int64 [B, S].  Returns the last hidden state as a bf16 [B*S, D] token matrix."""
        c = self.cfg
        B, S = ids.shape
        if S > 512:
            raise NotImplementedError("attention kernels cover sequence lengths up to 512 (the reference's maximum)")
        if S > c.max_position_embeddings:
            raise ValueError("sequence longer than the position table")
        ids = ids.contiguous()
        H = c.n_heads
        pd = c.dropout if training else 0.0
        pa = c.attention_dropout if training else 0.0
        key_bias = ops.mask_to_bias(mask)
        s_emb = _mix(seed, step, 255, 0)
        x, x_emb, e_mean, e_rstd = ops.embed_layernorm_fwd(ids, self.word, self.pos, self.eg, self.eb,
                                                           c.layer_norm_eps, p_drop=pd, seed=s_emb)
        saved = {"ids": ids, "key_bias": key_bias, "B": B, "S": S, "emb": (x_emb, e_mean, e_rstd, pd, s_emb),
                 "layers": []} if training else None
        if self.capture is not None:
            self.capture.append(x)
        for li, L in enumerate(self.layers):
            s_att, s_ffn = _mix(seed, step, li, 1), _mix(seed, step, li, 2)
            qkv = ops.linear_fwd(x, L["wqkv"], L["bqkv"])
            ctx, lse = ops.attention_fwd(qkv, key_bias, B, H, S, p_drop=pa, seed=s_att, need_lse=training)
            y_pre = ops.linear_fwd(ctx, L["wo"], L["bo"], residual=x)
            y, m1, r1 = ops.layernorm_fwd(y_pre, L["g1"], L["be1"], c.layer_norm_eps)
            z, a = ops.linear_gelu_fwd(y, L["w1"], L["b1"])
            o_pre = ops.linear_fwd(a, L["w2"], L["b2"], residual=y, p_drop=pd, seed=s_ffn)
            out, m2, r2 = ops.layernorm_fwd(o_pre, L["g2"], L["be2"], c.layer_norm_eps)
            if training:
                saved["layers"].append((x, qkv, ctx, lse, y_pre, m1, r1, y, z, a, o_pre, m2, r2, pa, s_att, pd, s_ffn))
            x = out
            if self.capture is not None:
                self.capture.append(x)
        self._saved = saved
        return x

    # ------------------------------------------------------------------ backward
    def backward(self, dh: torch.Tensor):
        """dh: gradient w.r.t. the returned token matrix, bf16 [B*S, D]. Accumulates into the store's grad buffer."""
        sv = self._saved
        assert sv is not None, "backward() without a training-mode forward()"
        B, S, H = sv["B"], sv["S"], self.cfg.n_heads
        kb = sv["key_bias"]
        d_out = dh
        for li in reversed(range(len(self.layers))):
            L = self.layers[li]
            x, qkv, ctx, lse, y_pre, m1, r1, y, z, a, o_pre, m2, r2, pa, s_att, pd, s_ffn = sv["layers"][li]
            # out = LN2(o_pre);  o_pre = dropout(a W2^T + b2) + y
            d_opre, d_opre_m = ops.layernorm_bwd(d_out, o_pre, m2, r2, L["g2"], L["dg2"], L["dbe2"],
                                                 p_out=pd, seed_out=s_ffn)
            d_lin2 = d_opre_m if d_opre_m is not None else d_opre
            ops.linear_wgrad(d_lin2, a, L["dw2"])
            ops.colsum(d_lin2, L["db2"])
            dz = ops.linear_dgrad(d_lin2, L["w2"], gelu_z=z)               # (d_lin2 W2) * gelu'(z)
            ops.linear_wgrad(dz, y, L["dw1"])
            ops.colsum(dz, L["db1"])
            dy = ops.linear_dgrad(dz, L["w1"], residual=d_opre)            # + residual path of LN2's input
            # y = LN1(y_pre);  y_pre = ctx Wo^T + bo + x
            d_ypre, _ = ops.layernorm_bwd(dy, y_pre, m1, r1, L["g1"], L["dg1"], L["dbe1"])
            ops.linear_wgrad(d_ypre, ctx, L["dwo"])
            ops.colsum(d_ypre, L["dbo"])
            dctx = ops.linear_dgrad(d_ypre, L["wo"])
            dqkv = ops.attention_bwd(qkv, kb, ctx, dctx, lse, B, H, S, p_drop=pa, seed=s_att)
            ops.linear_wgrad(dqkv, x, L["dwqkv"])
            ops.colsum(dqkv, L["dbqkv"])
            d_out = ops.linear_dgrad(dqkv, L["wqkv"], residual=d_ypre)     # + residual path of LN1's input
        x_emb, e_mean, e_rstd, pd, s_emb = sv["emb"]
        d_emb, _ = ops.layernorm_bwd(d_out, x_emb, e_mean, e_rstd, self.eg, self.deg, self.deb, p_in=pd, seed_in=s_emb)
        ops.embedding_bwd(d_emb, sv["ids"], self.dword, self.dpos, padding_idx=self.cfg.pad_token_id)
        self._saved = None
