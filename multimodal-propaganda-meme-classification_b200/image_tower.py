"""Image tower: torchvision-style ResNet (Bottleneck, v1.5) with explicit forward / backward on the sm_100a
kernels.  Activations are NHWC bf16 token matrices [N*H*W, C]; every convolution runs on tcgen05
(1x1: the GEMM directly on the activation matrix; 3x3: implicit GEMM, the activation operand gathered by im2col-mode
TMA loads -- at 64 channels the halo-resident kernels of csrc/conv3x3_c64.cu, picked inside b200mm_conv_fwd /
b200mm_conv_wgrad; 7x7 stem: csrc/stem_conv.cu straight from the fp32 image, no im2col matrix; the three stride-2
data gradients: explicit col2im); BatchNorm runs in training mode with per-replica batch statistics exactly like the
reference.

Mirrors ``self.resnet(image)`` of example_scripts/Multimodal_example_task2C.txt:164, :183 ->
torchvision/models/resnet.py:108-163 (Bottleneck), :197-213 (stem, init), :266-280 (_forward_impl).
The 1000-way ImageNet ``fc`` is kept, as the reference keeps it (.txt:164-165).
"""
from __future__ import annotations

from dataclasses import dataclass

import os

import torch

from . import ops
from .params import ParamStore

STEM_K = 7 * 7 * 3
STEM_KP = 152  # K padded to a multiple of 8 elements (TMA row stride must be a multiple of 16 bytes)


# B200MM_MASKRES=1: never materialise the identity-branch gradient dz -- the first convolution's data gradient adds
# (mask ? d_out : 0) in its epilogue (b200mm_gemm_bf16_maskres).  Off by default: measured on config 2 the BatchNorm
# backward gains 0.42 ms (one write less per block) but the residual-loading epilogue costs the short-K GEMMs 0.65 ms
# against the TMA reduce-add accumulate it replaces (30.58 vs 30.34 ms per step, profiles/step_profile_r02_maskres*.json).
_MASKRES = os.environ.get("B200MM_MASKRES", "0") == "1"


@dataclass
class ImageConfig:
    layers: tuple = (3, 4, 6, 3)   # ResNet-50
    width: int = 64
    block: str = "bottleneck"      # 'bottleneck' (ResNet-50/101/152) | 'basic' (ResNet-18/34, the HEAD script's tower)
    num_outputs: int = 1000        # 0: no fc, the pooled feature is returned (timm reset_classifier(0))
    bn_eps: float = 1e-5
    bn_momentum: float = 0.1
    prefix: str = "resnet"
    arch: str = "resnet"

    @staticmethod
    def resnet18(**kw) -> "ImageConfig":
        """timm.create_model('resnet18') + reset_classifier(0) (Multimodal_example_task2C.py:84, 569-570)."""
        return ImageConfig(layers=(2, 2, 2, 2), block="basic", num_outputs=0, **kw)


class _Conv:
    """One conv (+ its BatchNorm): parameter views and geometry."""

    def __init__(self, name, bn_name, cin, cout, k, stride, pad):
        self.name, self.bn_name = name, bn_name
        self.cin, self.cout, self.k, self.stride, self.pad = cin, cout, k, stride, pad
        self.kdim = STEM_KP if cin == 3 else k * k * cin


class ImageTower:
    def __init__(self, cfg: ImageConfig, store: ParamStore):
        self.cfg = cfg
        self.store = store
        p = cfg.prefix
        self.stem = _Conv(f"{p}.conv1", f"{p}.bn1", 3, cfg.width, 7, 2, 3)
        self.blocks = []
        inplanes = cfg.width
        if cfg.block not in ("bottleneck", "basic"):
            raise ValueError(f"unknown ResNet block {cfg.block!r}")
        self.basic = cfg.block == "basic"
        expansion = 1 if self.basic else 4
        for li, nblocks in enumerate(cfg.layers):
            planes = cfg.width * (2 ** li)
            for bi in range(nblocks):
                stride = 2 if (li > 0 and bi == 0) else 1
                base = f"{p}.layer{li + 1}.{bi}"
                if self.basic:     # torchvision BasicBlock (resnet.py:59-105): 3x3 (stride) -> 3x3
                    blk = {
                        "c1": _Conv(f"{base}.conv1", f"{base}.bn1", inplanes, planes, 3, stride, 1),
                        "c2": _Conv(f"{base}.conv2", f"{base}.bn2", planes, planes, 3, 1, 1),
                        "c3": None, "ds": None, "stride": stride,
                    }
                else:
                    blk = {
                        "c1": _Conv(f"{base}.conv1", f"{base}.bn1", inplanes, planes, 1, 1, 0),
                        "c2": _Conv(f"{base}.conv2", f"{base}.bn2", planes, planes, 3, stride, 1),
                        "c3": _Conv(f"{base}.conv3", f"{base}.bn3", planes, planes * 4, 1, 1, 0),
                        "ds": None,
                        "stride": stride,
                    }
                if stride != 1 or inplanes != planes * expansion:
                    blk["ds"] = _Conv(f"{base}.downsample.0", f"{base}.downsample.1", inplanes, planes * expansion, 1,
                                      stride, 0)
                self.blocks.append(blk)
                inplanes = planes * expansion
        self.feat_dim = inplanes
        self.out_dim = cfg.num_outputs or inplanes   # the 1000-way ImageNet fc is kept (.txt:164-165) unless 0
        self._convs = [self.stem] + [c for b in self.blocks for c in (b["c1"], b["c2"], b["c3"], b["ds"]) if c]
        self._conv_by_weight = {c.name + ".weight": c for c in self._convs}
        off = 0
        for c in self._convs:
            c.stats_off = off
            off += 2 * c.cout
        self._stats_total, self._stats = off, None
        self._saved = None
        self.buffers = None
        self.capture = None   # set to a list to record every block's output (per-layer parity checks)

    # ------------------------------------------------------------------ parameters
    def register_noshadow(self):
        st = self.store
        for c in self._convs:
            st.add(f"{c.bn_name}.weight", (c.cout,), shadow=False)
            st.add(f"{c.bn_name}.bias", (c.cout,), shadow=False)
        if self.cfg.num_outputs:
            st.add(f"{self.cfg.prefix}.fc.bias", (self.cfg.num_outputs,), shadow=False)

    def register_shadowed(self):
        st = self.store
        for c in self._convs:
            st.add(f"{c.name}.weight", (c.cout, c.kdim))   # OHWI flattened: [Cout, kh*kw*Cin] (stem padded to 152)
        if self.cfg.num_outputs:
            st.add(f"{self.cfg.prefix}.fc.weight", (self.cfg.num_outputs, self.feat_dim))

    def bind(self):
        st = self.store
        dev = st.device
        # running statistics live outside the optimizer's flat buffer
        total = sum(2 * c.cout for c in self._convs)
        self.buffers = torch.zeros(total, device=dev, dtype=torch.float32)
        off = 0
        for c in self._convs:
            c.w, c.dw = st.s(f"{c.name}.weight"), st.g(f"{c.name}.weight")
            c.g, c.dg = st.p(f"{c.bn_name}.weight"), st.g(f"{c.bn_name}.weight")
            c.b, c.db = st.p(f"{c.bn_name}.bias"), st.g(f"{c.bn_name}.bias")
            c.rm = self.buffers[off:off + c.cout]
            c.rv = self.buffers[off + c.cout:off + 2 * c.cout]
            c.rv.fill_(1.0)
            off += 2 * c.cout
        p = self.cfg.prefix
        if self.cfg.num_outputs:
            self.fc_w, self.dfc_w = st.s(f"{p}.fc.weight"), st.g(f"{p}.fc.weight")
            self.fc_b, self.dfc_b = st.p(f"{p}.fc.bias"), st.g(f"{p}.fc.bias")
        self.num_batches_tracked = 0

    def init_parameters(self, generator=None):
        """torchvision init (resnet.py:208-213): kaiming_normal_(fan_out, relu) convs, BN = (1, 0); nn.Linear default
        (kaiming_uniform a=sqrt(5)) for fc."""
        st = self.store
        for c in self._convs:
            w = st.p(f"{c.name}.weight")
            std = (2.0 / (c.cout * c.k * c.k)) ** 0.5
            w.normal_(0.0, std, generator=generator)
            if c.cin == 3:
                w[:, STEM_K:].zero_()
            st.p(f"{c.bn_name}.weight").fill_(1.0)
            st.p(f"{c.bn_name}.bias").zero_()
        p = self.cfg.prefix
        if self.cfg.num_outputs:
            bound = 1.0 / (self.feat_dim ** 0.5)
            st.p(f"{p}.fc.weight").uniform_(-bound, bound, generator=generator)
            st.p(f"{p}.fc.bias").uniform_(-bound, bound, generator=generator)

    # ------------------------------------------------------------------ state-dict layout exchange (see model.py)
    def import_param(self, name: str, src: torch.Tensor, dst: torch.Tensor) -> None:
        """torch layout -> engine layout: conv OIHW -> OHWI-flattened (stem K padded 147 -> 152)."""
        if src.dim() == 4:
            co, ci, kh, kw = src.shape
            flat = src.permute(0, 2, 3, 1).reshape(co, kh * kw * ci)
            dst.zero_()
            dst[:, :flat.shape[1]].copy_(flat)
        else:
            dst.copy_(src.view(dst.shape))

    def export_param(self, name: str, t: torch.Tensor) -> torch.Tensor:
        c = self._conv_by_weight.get(name)
        if c is None:
            return t
        return t[:, :c.k * c.k * c.cin].reshape(c.cout, c.k, c.k, c.cin).permute(0, 3, 1, 2).contiguous()

    def load_buffers(self, sd: dict) -> None:
        for c in self._convs:
            c.rm.copy_(sd[f"{c.bn_name}.running_mean"])
            c.rv.copy_(sd[f"{c.bn_name}.running_var"])

    def export_buffers(self, out: dict) -> None:
        for c in self._convs:
            out[f"{c.bn_name}.running_mean"] = c.rm.clone()
            out[f"{c.bn_name}.running_var"] = c.rv.clone()
            out[f"{c.bn_name}.num_batches_tracked"] = torch.tensor(self.num_batches_tracked)

    # ------------------------------------------------------------------ building blocks
    def _stats_begin(self, training):
        """One zeroed fp32 arena per step: [2*cout] column sums / sums of squares per convolution, filled by the
        convolutions' own epilogues (so BatchNorm never re-reads its input for the statistics)."""
        self._stats = torch.zeros(self._stats_total, device=self.store.device, dtype=torch.float32) if training else None

    def _stats_of(self, c):
        return None if self._stats is None else self._stats[c.stats_off:c.stats_off + 2 * c.cout]

    def _bn(self, c, x, training, residual=None, relu=True, want_mask=False):
        """want_mask: also return the 1-bit ReLU mask (4th value) -- what the backward of a residual BatchNorm reads
        instead of the full output tensor."""
        cfg = self.cfg
        if training:
            return ops.batchnorm_fwd(x, c.g, c.b, c.rm, c.rv, residual=residual, relu=relu, eps=cfg.bn_eps,
                                     momentum=cfg.bn_momentum, col_stats=self._stats_of(c), want_mask=want_mask)
        out = ops.batchnorm_eval(x, c.g, c.b, c.rm, c.rv, residual=residual, relu=relu, eps=cfg.bn_eps)
        return (out, None, None, None) if want_mask else (out, None, None)

    # ------------------------------------------------------------------ forward
    def forward(self, image: torch.Tensor, *, training: bool, seed: int = 0, step: int = 0):
        """image: fp32 NCHW [N, 3, H, W] (what the reference's transforms produce). Returns bf16 [N, 1000].
        (seed / step: unused -- the ResNet has no dropout; kept for the tower interface shared with ViTTower.)"""
        N, Cin, H, W = image.shape
        assert Cin == 3
        sv = {"N": N, "blocks": []} if training else None
        st = self.stem
        self._stats_begin(training)
        direct = ops.stem_conv_supported(image, st.w)
        if direct:   # csrc/stem_conv.cu: patches gathered on the fly, no [N*Ho*Wo, 152] matrix
            c0, H1, W1 = ops.stem_conv_fwd(image, st.w, col_stats=self._stats_of(st))
            cols = image
        else:        # other stem widths / very wide images: explicit lowering
            cols, H1, W1 = ops.im2col_nchw_f32(image, 7, 2, 3, STEM_KP)
            c0 = ops.linear_fwd(cols, st.w, col_stats=self._stats_of(st))
        if training:
            # bn1 + relu + maxpool in one pass: the 112 x 112 activation between them is not needed again (the
            # BatchNorm backward recomputes the ReLU mask from c0) and is never written
            x, arg, H2, W2, m0, r0 = ops.bn_relu_maxpool_fwd(c0, N, H1, W1, st.cout, st.g, st.b, st.rm, st.rv,
                                                             self._stats_of(st), eps=self.cfg.bn_eps,
                                                             momentum=self.cfg.bn_momentum)
            sv["stem"] = (cols, direct, c0, None, m0, r0, arg, H1, W1)
        else:
            a0, m0, r0 = self._bn(st, c0, training)
            x, arg, H2, W2 = ops.maxpool_fwd(a0, N, H1, W1, st.cout)
        Hc, Wc = H2, W2
        for blk in self.blocks:
            c1, c2, c3, ds, stride = blk["c1"], blk["c2"], blk["c3"], blk["ds"], blk["stride"]
            if self.basic:
                y1, Ho, Wo = ops.conv_fwd(x, N, Hc, Wc, c1.cin, c1.w, 3, stride, 1, col_stats=self._stats_of(c1))
                a1, m1, r1 = self._bn(c1, y1, training)
                y2, _, _ = ops.conv_fwd(a1, N, Ho, Wo, c2.cin, c2.w, 3, 1, 1, col_stats=self._stats_of(c2))
                xs = yd = md = rd = None
                if ds is not None:
                    xs = x if stride == 1 else ops.subsample(x, N, Hc, Wc, ds.cin, stride)[0]
                    yd = ops.linear_fwd(xs, ds.w, col_stats=self._stats_of(ds))
                    idn, md, rd = self._bn(ds, yd, training, relu=False)
                else:
                    idn = x
                out, m2, r2, msk = self._bn(c2, y2, training, residual=idn, relu=True, want_mask=True)
                if training:
                    sv["blocks"].append((x, y1, a1, m1, r1, y2, m2, r2, xs, yd, md, rd, msk, Hc, Wc, Ho, Wo))
                x, Hc, Wc = out, Ho, Wo
                if self.capture is not None:
                    self.capture.append((x, N, Hc, Wc))
                continue
            y1 = ops.linear_fwd(x, c1.w, col_stats=self._stats_of(c1))
            a1, m1, r1 = self._bn(c1, y1, training)
            y2, Ho, Wo = ops.conv_fwd(a1, N, Hc, Wc, c2.cin, c2.w, 3, stride, 1,      # implicit GEMM (TMA im2col)
                                      col_stats=self._stats_of(c2))
            a2, m2, r2 = self._bn(c2, y2, training)
            y3 = ops.linear_fwd(a2, c3.w, col_stats=self._stats_of(c3))
            xs = yd = md = rd = None
            if ds is not None:
                xs = x if stride == 1 else ops.subsample(x, N, Hc, Wc, ds.cin, stride)[0]
                yd = ops.linear_fwd(xs, ds.w, col_stats=self._stats_of(ds))
                idn, md, rd = self._bn(ds, yd, training, relu=False)
            else:
                idn = x
            out, m3, r3, msk = self._bn(c3, y3, training, residual=idn, relu=True, want_mask=True)
            if training:
                sv["blocks"].append((x, y1, a1, m1, r1, y2, a2, m2, r2, y3, m3, r3, xs, yd, md, rd, msk, Hc, Wc,
                                     Ho, Wo))
            x, Hc, Wc = out, Ho, Wo
            if self.capture is not None:
                self.capture.append((x, N, Hc, Wc))
        pooled = ops.avgpool_fwd(x, N, Hc * Wc, self.feat_dim)
        logits = ops.linear_fwd(pooled, self.fc_w, self.fc_b) if self.cfg.num_outputs else pooled
        if training:
            sv["tail"] = (pooled, Hc, Wc)
            self.num_batches_tracked += 1
        self._saved = sv
        return logits

    # ------------------------------------------------------------------ data-parallel gradient phases
    def grad_phases(self):
        """Ordered (tag, predicate) list for ddp.GradSync: the convolution weights of layer4 (+ fc), layer3, layer2,
        layer1 -- the order the backward finishes them.  The stem and the BatchNorm affine parameters (tiny, and stored
        in the shadow-less prefix of the flat buffer) fall to the model's catch-all last phase."""
        pre = self.cfg.prefix
        n_stage = len(self.cfg.layers)
        out = []
        for li in reversed(range(n_stage)):
            stem = f"{pre}.layer{li + 1}."
            fc = f"{pre}.fc.weight" if li == n_stage - 1 else None
            out.append((f"{pre}.layer{li + 1}",
                        lambda n, stem=stem, fc=fc: (n.startswith(stem) and n.endswith(".weight")
                                                     and (".conv" in n or ".downsample.0." in n)) or n == fc))
        return out

    def _rotate_weights(self):
        """Rotated (transposed-convolution) copies of every stride-1 3x3 weight, refreshed from the bf16 shadow in ONE
        launch at the top of the backward (they change once per optimizer step; 13 launches in ResNet-50 before)."""
        convs = [c for c in self._convs if c.k == 3 and c.stride == 1 and c.cin != 3]
        key = tuple(c.w.data_ptr() for c in convs)           # (the flat shadow buffer may have been re-created)
        if getattr(self, "_rot_key", None) != key:
            self._rot_key = key
            total = sum(c.cin * 9 * c.cout for c in convs)
            dev = self.store.device
            self._rot_arena = torch.empty(max(total, 1), device=dev, dtype=torch.bfloat16)
            rows, off = [], 0
            for c in convs:
                c.w_rot = self._rot_arena[off:off + c.cin * 9 * c.cout].view(c.cin, 9 * c.cout)
                rows.append([c.w.data_ptr(), c.w_rot.data_ptr(), (c.cout << 32) | c.cin, 9])
                off += c.cin * 9 * c.cout
            self._rot_n = len(rows)
            self._rot_table = torch.tensor(rows if rows else [[0, 0, 0, 0]], dtype=torch.int64).to(dev)
        if self._rot_n:
            ops.conv_weight_rotate_multi(self._rot_table, self._rot_n)

    # ------------------------------------------------------------------ backward
    def backward(self, dlogits: torch.Tensor, on_grads_ready=None):
        """dlogits: bf16 [N, 1000]. Accumulates parameter gradients (the image itself needs none).
        on_grads_ready(tag): called as soon as every gradient of a ``grad_phases`` group is final."""
        sv = self._saved
        assert sv is not None, "backward() without a training-mode forward()"
        N = sv["N"]
        pooled, Hc, Wc = sv["tail"]
        if self.cfg.num_outputs:
            ops.linear_wgrad(dlogits, pooled, self.dfc_w)
            ops.colsum(dlogits, self.dfc_b)
            dpooled = ops.linear_dgrad(dlogits, self.fc_w)
        else:
            dpooled = dlogits.contiguous()
        d_out = ops.avgpool_bwd(dpooled, N, Hc * Wc, self.feat_dim)
        stage_of = [li for li, nb in enumerate(self.cfg.layers) for _ in range(nb)]
        self._rotate_weights()
        if getattr(self, "_wq", None) is None:
            self._wq = ops.SideQueue(dlogits.device)
        wq = self._wq       # weight gradients run beside the data-gradient chain (ops.SideQueue)
        for bi, blk, s in zip(reversed(range(len(self.blocks))), reversed(self.blocks), reversed(sv["blocks"])):
            if on_grads_ready is not None and bi + 1 < len(self.blocks) and stage_of[bi + 1] != stage_of[bi]:
                wq.join()
                on_grads_ready(f"{self.cfg.prefix}.layer{stage_of[bi + 1] + 1}")     # the stage above is complete
            c1, c2, c3, ds, stride = blk["c1"], blk["c2"], blk["c3"], blk["ds"], blk["stride"]
            if self.basic:
                x, y1, a1, m1, r1, y2, m2, r2, xs, yd, md, rd, msk, Hi, Wi, Ho, Wo = s
                # out = relu(bn2(y2) + idn)
                d_y2, dz = ops.batchnorm_bwd(d_out, None, y2, m2, r2, c2.g, c2.dg, c2.db, relu=True, need_dz=True,
                                             mask=msk)
                ops.conv_wgrad(d_y2, a1, N, Ho, Wo, c2.cin, 3, 1, 1, c2.dw)
                d_a1, _, _ = ops.conv_fwd(d_y2, N, Ho, Wo, c2.cout, c2.w_rot,
                                          3, 1, 1)
                d_y1, _ = ops.batchnorm_bwd(d_a1, None, y1, m1, r1, c1.g, c1.dg, c1.db, relu=True, beta=c1.b)
                ops.conv_wgrad(d_y1, x, N, Hi, Wi, c1.cin, 3, stride, 1, c1.dw)
                if ds is None:      # stride 1: data gradient + identity branch in one epilogue
                    d_out, _, _ = ops.conv_fwd(d_y1, N, Ho, Wo, c1.cout,
                                               c1.w_rot, 3, 1, 1, residual=dz)
                else:
                    d_yd, _ = ops.batchnorm_bwd(dz, None, yd, md, rd, ds.g, ds.dg, ds.db, relu=False)
                    ops.linear_wgrad(d_yd, xs, ds.dw)
                    d_xs = ops.linear_dgrad(d_yd, ds.w)
                    if stride == 1:
                        d_x1, _, _ = ops.conv_fwd(d_y1, N, Ho, Wo, c1.cout,
                                                  c1.w_rot, 3, 1, 1,
                                                  residual=d_xs)
                        d_out = d_x1
                    else:
                        d_cols = ops.linear_dgrad(d_y1, c1.w)
                        d_x1 = ops.col2im(d_cols, N, Hi, Wi, c1.cin, 3, stride, 1)
                        d_out = ops.upsample_add(d_xs, d_x1, N, Hi, Wi, ds.cin, stride)
                continue
            x, y1, a1, m1, r1, y2, a2, m2, r2, y3, m3, r3, xs, yd, md, rd, msk, Hi, Wi, Ho, Wo = s
            # out = relu(bn3(y3) + idn).  The gradient of the pre-activation sum, dz = d_out o relu_mask, is what both
            # branches receive; it is never written out: the consumers below take (d_out, mask) instead (one full
            # write + read of the block's output size saved per block)
            fused_id = _MASKRES
            d_y3, dz = ops.batchnorm_bwd(d_out, None, y3, m3, r3, c3.g, c3.dg, c3.db, relu=True, mask=msk,
                                         need_dz=not fused_id)
            wq.run(lambda: ops.linear_wgrad(d_y3, a2, c3.dw), d_y3, a2)
            d_a2 = ops.linear_dgrad(d_y3, c3.w)
            d_y2, _ = ops.batchnorm_bwd(d_a2, None, y2, m2, r2, c2.g, c2.dg, c2.db, relu=True, beta=c2.b)
            wq.run(lambda: ops.conv_wgrad(d_y2, a1, N, Hi, Wi, c2.cin, 3, stride, 1, c2.dw), d_y2, a1)
            if stride == 1:
                # data gradient = the same implicit-GEMM convolution applied to dY with the rotated weight
                d_a1, _, _ = ops.conv_fwd(d_y2, N, Ho, Wo, c2.cout, c2.w_rot, 3, 1, 1)
            else:
                d_cols2 = ops.linear_dgrad(d_y2, c2.w)
                d_a1 = ops.col2im(d_cols2, N, Hi, Wi, c2.cin, 3, stride, 1)
            d_y1, _ = ops.batchnorm_bwd(d_a1, None, y1, m1, r1, c1.g, c1.dg, c1.db, relu=True, beta=c1.b)
            wq.run(lambda: ops.linear_wgrad(d_y1, x, c1.dw), d_y1, x)
            if ds is None:
                if fused_id:
                    # identity branch joins the data gradient in the epilogue: d_out' = d_y1 W1 + (mask ? d_out : 0)
                    d_out = ops.linear_dgrad(d_y1, c1.w, residual=d_out, residual_mask=msk)
                else:
                    # dz += d_y1 W1 in place (TMA reduce-add stores; the residual is never loaded)
                    d_out = ops.linear_dgrad(d_y1, c1.w, residual=dz, out=dz)
            else:
                if fused_id:
                    # downsample branch: its BatchNorm has no ReLU of its own, the block's mask turns d_out into dz on load
                    d_yd, _ = ops.batchnorm_bwd(d_out, None, yd, md, rd, ds.g, ds.dg, ds.db, relu=True, mask=msk)
                else:
                    d_yd, _ = ops.batchnorm_bwd(dz, None, yd, md, rd, ds.g, ds.dg, ds.db, relu=False)
                wq.run(lambda: ops.linear_wgrad(d_yd, xs, ds.dw), d_yd, xs)
                d_x1 = ops.linear_dgrad(d_y1, c1.w)
                if stride == 1:
                    d_out = ops.linear_dgrad(d_yd, ds.w, residual=d_x1)
                else:
                    d_xs = ops.linear_dgrad(d_yd, ds.w)
                    d_out = ops.upsample_add(d_xs, d_x1, N, Hi, Wi, ds.cin, stride)
        wq.join()
        if on_grads_ready is not None:
            on_grads_ready(f"{self.cfg.prefix}.layer1")
        cols, direct, c0, a0, m0, r0, arg, H1, W1 = sv["stem"]
        st = self.stem
        d_a0 = ops.maxpool_bwd(d_out, arg, N, H1, W1, st.cout)
        d_c0, _ = ops.batchnorm_bwd(d_a0, None, c0, m0, r0, st.g, st.dg, st.db, relu=True, beta=st.b)
        if direct:
            ops.stem_conv_wgrad(cols, d_c0, st.dw)
        else:
            ops.linear_wgrad(d_c0, cols, st.dw)
        self._saved = None
