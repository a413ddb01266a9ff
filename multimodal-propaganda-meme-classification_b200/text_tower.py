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

``TextConfig.arch`` selects the member of the family (BASELINE configs 3-5 and the HEAD
script's AraBERT / RoBERTa towers, example_scripts/Multimodal_example_task2C.py:76-80, 317, 337):
  * ``distilbert``  -- as above (parameter names of transformers' DistilBertModel);
  * ``bert``        -- + token_type_embeddings[0], dropout after the attention output projection
                       (transformers/models/bert/modeling_bert.py:53-113, 287-298, 330-356, 456-469);
  * ``roberta``     -- ``bert`` with RoBERTa / XLM-R position ids (cumsum over non-pad tokens, offset pad_id + 1;
                       transformers/models/xlm_roberta/modeling_xlm_roberta.py:56-159).
The ``pooler.dense`` parameters of BERT / XLM-R are not part of the engine: the reference reads
``last_hidden_state`` only, so they never receive a gradient.
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
    arch: str = "distilbert"   # 'distilbert' | 'bert' | 'roberta' (RoBERTa / XLM-R)
    type_vocab_size: int = 2   # bert / roberta only (xlm-roberta checkpoints use 1)

    @staticmethod
    def bert_base(vocab_size: int = 64000, **kw) -> "TextConfig":
        """BERT-base / AraBERTv2-shaped (BASELINE configs 3, 5; HEAD script text tower)."""
        return TextConfig(vocab_size=vocab_size, n_layers=12, arch="bert", **kw)

    @staticmethod
    def roberta_base(vocab_size: int = 50265, **kw) -> "TextConfig":
        """RoBERTa-base (HEAD script caption tower, Multimodal_example_task2C.py:603-612)."""
        return TextConfig(vocab_size=vocab_size, max_position_embeddings=514, n_layers=12, layer_norm_eps=1e-5,
                          pad_token_id=1, arch="roberta", type_vocab_size=1, **kw)

    @staticmethod
    def xlmr_large(vocab_size: int = 250002, **kw) -> "TextConfig":
        """XLM-R-large (BASELINE config 4)."""
        return TextConfig(vocab_size=vocab_size, max_position_embeddings=514, dim=1024, n_layers=24, n_heads=16,
                          hidden_dim=4096, layer_norm_eps=1e-5, pad_token_id=1, arch="roberta", type_vocab_size=1,
                          **kw)


def _layer_names(cfg: TextConfig, i: int) -> dict:
    """Parameter-name stems of encoder layer i under the reference libraries' module names."""
    p = cfg.prefix
    if cfg.arch == "distilbert":
        L = f"{p}.transformer.layer.{i}"
        return {"q": f"{L}.attention.q_lin", "k": f"{L}.attention.k_lin", "v": f"{L}.attention.v_lin",
                "o": f"{L}.attention.out_lin", "ln1": f"{L}.sa_layer_norm", "f1": f"{L}.ffn.lin1",
                "f2": f"{L}.ffn.lin2", "ln2": f"{L}.output_layer_norm"}
    L = f"{p}.encoder.layer.{i}"
    return {"q": f"{L}.attention.self.query", "k": f"{L}.attention.self.key", "v": f"{L}.attention.self.value",
            "o": f"{L}.attention.output.dense", "ln1": f"{L}.attention.output.LayerNorm",
            "f1": f"{L}.intermediate.dense", "f2": f"{L}.output.dense", "ln2": f"{L}.output.LayerNorm"}


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
        if c.arch not in ("distilbert", "bert", "roberta"):
            raise ValueError(f"unknown text tower arch {c.arch!r}")
        st.add(f"{p}.embeddings.word_embeddings.weight", (c.vocab_size, c.dim), shadow=False)
        st.add(f"{p}.embeddings.position_embeddings.weight", (c.max_position_embeddings, c.dim), shadow=False)
        if c.arch != "distilbert":
            st.add(f"{p}.embeddings.token_type_embeddings.weight", (c.type_vocab_size, c.dim), shadow=False)
        st.add(f"{p}.embeddings.LayerNorm.weight", (c.dim,), shadow=False)
        st.add(f"{p}.embeddings.LayerNorm.bias", (c.dim,), shadow=False)
        for i in range(c.n_layers):
            n = _layer_names(c, i)
            for k in ("q", "k", "v"):
                st.add(f"{n[k]}.bias", (c.dim,), shadow=False)
            st.add(f"{n['o']}.bias", (c.dim,), shadow=False)
            st.add(f"{n['ln1']}.weight", (c.dim,), shadow=False)
            st.add(f"{n['ln1']}.bias", (c.dim,), shadow=False)
            st.add(f"{n['f1']}.bias", (c.hidden_dim,), shadow=False)
            st.add(f"{n['f2']}.bias", (c.dim,), shadow=False)
            st.add(f"{n['ln2']}.weight", (c.dim,), shadow=False)
            st.add(f"{n['ln2']}.bias", (c.dim,), shadow=False)

    def register_shadowed(self):
        c, st = self.cfg, self.store
        for i in range(c.n_layers):
            n = _layer_names(c, i)
            for k in ("q", "k", "v"):
                st.add(f"{n[k]}.weight", (c.dim, c.dim))
            st.add(f"{n['o']}.weight", (c.dim, c.dim))
            st.add(f"{n['f1']}.weight", (c.hidden_dim, c.dim))
            st.add(f"{n['f2']}.weight", (c.dim, c.hidden_dim))

    def bind(self):
        """Resolve views once the store is finalized."""
        c, st, p = self.cfg, self.store, self.cfg.prefix
        D = c.dim
        e = f"{p}.embeddings"
        self.word, self.dword = st.p(f"{e}.word_embeddings.weight"), st.g(f"{e}.word_embeddings.weight")
        self.pos, self.dpos = st.p(f"{e}.position_embeddings.weight"), st.g(f"{e}.position_embeddings.weight")
        self.type0 = self.dtype0 = None
        if c.arch != "distilbert":
            self.type0 = st.p(f"{e}.token_type_embeddings.weight")[0]
            self.dtype0 = st.g(f"{e}.token_type_embeddings.weight")[0]
        self.eg, self.deg = st.p(f"{e}.LayerNorm.weight"), st.g(f"{e}.LayerNorm.weight")
        self.eb, self.deb = st.p(f"{e}.LayerNorm.bias"), st.g(f"{e}.LayerNorm.bias")
        self.layers = []
        for i in range(c.n_layers):
            n = _layer_names(c, i)
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
        """transformers' DistilBERT init: N(0, 0.02) for linear / embedding weights, zeros for biases, LN = (1, 0)."""
        st = self.store
        for name in st.names():
            if not name.startswith(self.cfg.prefix + "."):
                continue
            t = st.p(name)
            if "LayerNorm.weight" in name or "layer_norm.weight" in name:
                t.fill_(1.0)
            elif "LayerNorm.bias" in name or "layer_norm.bias" in name:
                t.zero_()
            elif name.endswith(".bias"):
                t.zero_()
            else:
                t.normal_(0.0, 0.02, generator=generator)
        if self.cfg.arch != "distilbert":   # nn.Embedding(padding_idx=...) rows start at zero in BERT / RoBERTa
            self.word[self.cfg.pad_token_id].zero_()
            if self.cfg.arch == "roberta":
                self.pos[self.cfg.pad_token_id].zero_()

    # ------------------------------------------------------------------ forward
    def forward(self, ids: torch.Tensor, mask: torch.Tensor, *, training: bool, seed: int = 0, step: int = 0):
        """ids, mask: int64 [B, S].  Returns the last hidden state as a bf16 [B*S, D] token matrix."""
        c = self.cfg
        B, S = ids.shape
        if S > 512:
            raise NotImplementedError("attention kernels cover sequence lengths up to 512 (the reference's maximum)")
        roberta = c.arch == "roberta"
        if S + (c.pad_token_id + 1 if roberta else 0) > c.max_position_embeddings:
            raise ValueError("sequence longer than the position table")
        ids = ids.contiguous()
        pos_ids = ops.position_ids(ids, c.pad_token_id) if roberta else None
        p_out = (c.dropout if training else 0.0) if c.arch != "distilbert" else 0.0   # BertSelfOutput.dropout
        H = c.n_heads
        pd = c.dropout if training else 0.0
        pa = c.attention_dropout if training else 0.0
        key_bias = ops.mask_to_bias(mask)
        s_emb = _mix(seed, step, 255, 0)
        x, x_emb, e_mean, e_rstd = ops.embed_layernorm_fwd(ids, self.word, self.pos, self.eg, self.eb,
                                                           c.layer_norm_eps, p_drop=pd, seed=s_emb, pos_ids=pos_ids,
                                                           type_row=self.type0)
        saved = {"ids": ids, "pos_ids": pos_ids, "key_bias": key_bias, "B": B, "S": S,
                 "emb": (x_emb, e_mean, e_rstd, pd, s_emb), "layers": []} if training else None
        if self.capture is not None:
            self.capture.append(x)
        for li, L in enumerate(self.layers):
            s_att, s_ffn, s_out = _mix(seed, step, li, 1), _mix(seed, step, li, 2), _mix(seed, step, li, 3)
            qkv = ops.linear_fwd(x, L["wqkv"], L["bqkv"])
            if training:   # the forward hands its dropout keep bits to the backward (one-tile sequences)
                ctx, lse, amask = ops.attention_fwd(qkv, key_bias, B, H, S, p_drop=pa, seed=s_att, save_mask=True)
            else:
                (ctx, lse), amask = ops.attention_fwd(qkv, key_bias, B, H, S, p_drop=pa, seed=s_att, need_lse=False), None
            y_pre = ops.linear_fwd(ctx, L["wo"], L["bo"], residual=x, p_drop=p_out, seed=s_out)
            y, m1, r1 = ops.layernorm_fwd(y_pre, L["g1"], L["be1"], c.layer_norm_eps)
            z, a = ops.linear_gelu_fwd(y, L["w1"], L["b1"])
            o_pre = ops.linear_fwd(a, L["w2"], L["b2"], residual=y, p_drop=pd, seed=s_ffn)
            out, m2, r2 = ops.layernorm_fwd(o_pre, L["g2"], L["be2"], c.layer_norm_eps)
            if training:
                saved["layers"].append((x, qkv, ctx, lse, y_pre, m1, r1, y, z, a, o_pre, m2, r2, pa, s_att, pd, s_ffn,
                                        p_out, s_out, amask))
            x = out
            if self.capture is not None:
                self.capture.append(x)
        self._saved = saved
        return x

    # ------------------------------------------------------------------ data-parallel gradient phases
    GRAD_GROUPS = 3

    def _layer_stem(self, i: int) -> str:
        mid = "transformer" if self.cfg.arch == "distilbert" else "encoder"
        return f"{self.cfg.prefix}.{mid}.layer.{i}."

    def _group_of(self, li: int) -> int:
        per = -(-self.cfg.n_layers // self.GRAD_GROUPS)
        return li // per

    def grad_phases(self):
        """Ordered (tag, predicate) list for ddp.GradSync: encoder layers in GRAD_GROUPS groups, top group first (the
        order the backward finishes them), then everything else of this tower (embedding tables, embedding LN)."""
        pre = self.cfg.prefix
        out = []
        for g in reversed(range(self._group_of(self.cfg.n_layers - 1) + 1)):
            stems = tuple(self._layer_stem(i) for i in range(self.cfg.n_layers) if self._group_of(i) == g)
            out.append((f"{pre}.g{g}", lambda n, stems=stems: n.startswith(stems)))
        out.append((f"{pre}.tail", lambda n, pre=pre: n.startswith(pre + ".")))
        return out

    # ------------------------------------------------------------------ backward
    def backward(self, dh: torch.Tensor, on_grads_ready=None):
        """dh: gradient w.r.t. the returned token matrix, bf16 [B*S, D]. Accumulates into the store's grad buffer.
        on_grads_ready(tag): called as soon as every gradient of a ``grad_phases`` group is final."""
        sv = self._saved
        assert sv is not None, "backward() without a training-mode forward()"
        B, S, H = sv["B"], sv["S"], self.cfg.n_heads
        kb = sv["key_bias"]
        d_out = dh
        if getattr(self, "_wq", None) is None:
            self._wq = ops.SideQueue(dh.device)
        wq = self._wq
        for li in reversed(range(len(self.layers))):
            L = self.layers[li]
            x, qkv, ctx, lse, y_pre, m1, r1, y, z, a, o_pre, m2, r2, pa, s_att, pd, s_ffn, p_out, s_out, amask = \
                sv["layers"][li]
            # out = LN2(o_pre);  o_pre = dropout(a W2^T + b2) + y
            d_opre, d_opre_m = ops.layernorm_bwd(d_out, o_pre, m2, r2, L["g2"], L["dg2"], L["dbe2"],
                                                 p_out=pd, seed_out=s_ffn)
            d_lin2 = d_opre_m if d_opre_m is not None else d_opre
            wq.run(lambda: (ops.linear_wgrad(d_lin2, a, L["dw2"]), ops.colsum(d_lin2, L["db2"])), d_lin2, a)
            # (d_lin2 W2) * gelu'(z); the epilogue also accumulates the result's column sums = lin1's bias gradient
            dz = ops.linear_dgrad(d_lin2, L["w2"], gelu_z=z, bias_grad=L["db1"])
            wq.run(lambda: ops.linear_wgrad(dz, y, L["dw1"]), dz, y)
            dy = ops.linear_dgrad(dz, L["w1"], residual=d_opre)            # + residual path of LN2's input
            # y = LN1(y_pre);  y_pre = dropout(ctx Wo^T + bo) + x   (the dropout exists in BERT / RoBERTa only)
            d_ypre, d_ypre_m = ops.layernorm_bwd(dy, y_pre, m1, r1, L["g1"], L["dg1"], L["dbe1"],
                                                 p_out=p_out, seed_out=s_out)
            d_o = d_ypre_m if d_ypre_m is not None else d_ypre
            wq.run(lambda: (ops.linear_wgrad(d_o, ctx, L["dwo"]), ops.colsum(d_o, L["dbo"])), d_o, ctx)
            dctx = ops.linear_dgrad(d_o, L["wo"])
            dqkv = ops.attention_bwd(qkv, kb, ctx, dctx, lse, B, H, S, p_drop=pa, seed=s_att, drop_mask=amask)
            wq.run(lambda: (ops.linear_wgrad(dqkv, x, L["dwqkv"]), ops.colsum(dqkv, L["dbqkv"])), dqkv, x)
            d_out = ops.linear_dgrad(dqkv, L["wqkv"], residual=d_ypre)     # + residual path of LN1's input
            if on_grads_ready is not None and (li == 0 or self._group_of(li - 1) != self._group_of(li)):
                wq.join()
                on_grads_ready(f"{self.cfg.prefix}.g{self._group_of(li)}")
        x_emb, e_mean, e_rstd, pd, s_emb = sv["emb"]
        d_emb, _ = ops.layernorm_bwd(d_out, x_emb, e_mean, e_rstd, self.eg, self.deg, self.deb, p_in=pd, seed_in=s_emb)
        roberta = self.cfg.arch == "roberta"
        ops.embedding_bwd(d_emb, sv["ids"], self.dword, self.dpos, padding_idx=self.cfg.pad_token_id,
                          pos_ids=sv["pos_ids"], pos_padding_idx=self.cfg.pad_token_id if roberta else -1)
        if self.dtype0 is not None:
            ops.colsum(d_emb, self.dtype0)      # every token has segment id 0
        wq.join()
        if on_grads_ready is not None:
            on_grads_ready(f"{self.cfg.prefix}.tail")
        self._saved = None
