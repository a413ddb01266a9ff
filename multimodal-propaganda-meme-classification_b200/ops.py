"""Torch-tensor front end of the C-ABI kernels (device pointers + sizes go down, nothing else).

Every function launches on torch's current CUDA stream and allocates its outputs with ``torch.empty`` (the
caching allocator), so a whole step is CUDA-graph capturable.  There is no CPU path: tensors must live on
a CUDA device and libb200mm.so must load, otherwise these raise.
"""
from __future__ import annotations

import os

import torch

from . import _lib

bf16 = torch.bfloat16
f32 = torch.float32

EPI_STORE, EPI_GELU, EPI_DGELU, EPI_F32, EPI_F32_ATOMIC, EPI_RELU = range(6)
LOSS_CE, LOSS_FOCAL, LOSS_EXTERNAL = 0, 1, 2

_num_sms = None


def num_sms() -> int:
    global _num_sms
    if _num_sms is None:
        _num_sms = _lib.load().b200mm_num_sms() or 148
    return _num_sms


def _s():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, dtype, name="tensor"):
    if not t.is_cuda:
        raise _lib.B200MMError(f"{name} must be a CUDA tensor (b200mm has no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if t.dim() >= 1 and t.stride(-1) != 1:
        raise ValueError(f"{name}: innermost dimension must be contiguous")


# ----------------------------------------------------------------------------------------------- GEMM
_gemm_profile = None   # when a list: (start_event, end_event, flops) per GEMM launch (see profile_gemm)


def profile_gemm(step_fn, steps: int = 2, ridge: float = 208.0):
    """Time every tcgen05 GEMM launch (plain and implicit-conv) of ``steps`` calls of ``step_fn`` with CUDA events on
    the launching stream.  Returns (ms per step, FLOPs per step, launches per step, per-roofline detail): launches whose
    algorithmic intensity (FLOP / byte of operands + output) is below ``ridge`` are HBM-bound, the rest tensor-bound."""
    global _gemm_profile
    torch.cuda.synchronize()
    _gemm_profile = []
    try:
        for _ in range(steps):
            step_fn()
        torch.cuda.synchronize()
        rec = _gemm_profile
    finally:
        _gemm_profile = None
    ms = sum(r[0].elapsed_time(r[1]) for r in rec)
    fl = sum(r[2] for r in rec)
    # split by the roofline that bounds each launch: arithmetic intensity above / below the ridge point
    split = {"tensor": [0.0, 0.0, 0.0, 0, 0.0], "hbm": [0.0, 0.0, 0.0, 0, 0.0]}
    for a, b, f, by, wby in rec:
        k = "tensor" if f / by >= ridge else "hbm"
        split[k][0] += a.elapsed_time(b)
        split[k][1] += f
        split[k][2] += by
        split[k][3] += 1
        split[k][4] += wby
    detail = {k: {"ms": v[0] / steps, "flops": v[1] / steps, "bytes": v[2] / steps, "launches": v[3] // steps,
                  "bytes_written": v[4] / steps}
              for k, v in split.items()}
    return ms / steps, fl / steps, len(rec) // steps, detail


def gemm_raw(A, a_mn, B, b_mn, M, N, K, out, *, epi=EPI_STORE, bias=None, residual=None, aux=None, out2=None,
             splits=1, block_n=0, p_drop=0.0, seed=0, col_stats=None):
    """D[M,N] = A x B (+epilogue). A: [M,K] (a_mn False) or [K,M] (a_mn True); B: [N,K] or [K,N]."""
    if _gemm_profile is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call("b200mm_gemm_bf16", _p(A), int(a_mn), A.stride(0), _p(B), int(b_mn), B.stride(0), M, N, K, epi,
                  _p(bias), _p(residual), residual.stride(0) if residual is not None else 0,
                  _p(aux), aux.stride(0) if aux is not None else 0, _p(out), out.stride(0),
                  _p(out2), out2.stride(0) if out2 is not None else 0, splits, block_n, float(p_drop), int(seed),
                  _p(col_stats), _s())
        e1.record()
        nbytes = 2.0 * (M * K + N * K) + M * N * (8.0 if epi in (EPI_F32, EPI_F32_ATOMIC) else (4.0 if epi == EPI_GELU else 2.0))
        if residual is not None or aux is not None:
            nbytes += 2.0 * M * N
        _gemm_profile.append((e0, e1, 2.0 * M * N * K, nbytes,
                              M * N * (8.0 if epi in (EPI_F32, EPI_F32_ATOMIC) else (4.0 if epi == EPI_GELU else 2.0))))
        return out
    _lib.call("b200mm_gemm_bf16", _p(A), int(a_mn), A.stride(0), _p(B), int(b_mn), B.stride(0), M, N, K, epi,
              _p(bias), _p(residual), residual.stride(0) if residual is not None else 0,
              _p(aux), aux.stride(0) if aux is not None else 0, _p(out), out.stride(0),
              _p(out2), out2.stride(0) if out2 is not None else 0, splits, block_n, float(p_drop), int(seed),
              _p(col_stats), _s(), key=(M, N, K, int(a_mn), int(b_mn), epi, splits))
    return out


def linear_fwd(x, w, bias=None, *, residual=None, out=None, relu=False, p_drop=0.0, seed=0, col_stats=None):
    """y = x @ w.T + bias (+dropout) (+residual) (relu).  x [M,K] bf16, w [N,K] bf16, bias fp32 [N].
    col_stats (fp32 [2N], pre-zeroed): also accumulate sum / sum of squares of every output column (BatchNorm)."""
    M, K = x.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, device=x.device, dtype=bf16)
    return gemm_raw(x, False, w, False, M, N, K, out, epi=EPI_RELU if relu else EPI_STORE, bias=bias,
                    residual=residual, p_drop=p_drop, seed=seed, col_stats=col_stats)


def linear_fwd_f32(x, w, bias=None):
    """y = x @ w.T + bias with an fp32 result (for the rare spot where bf16 storage would lose the signal)."""
    M, K = x.shape
    N = w.shape[0]
    out = torch.empty(M, N, device=x.device, dtype=f32)
    return gemm_raw(x, False, w, False, M, N, K, out, epi=EPI_F32, bias=bias)


def linear_gelu_fwd(x, w, bias):
    """z = x @ w.T + bias ; a = gelu(z) (exact erf).  Returns (z, a), both bf16 [M,N]."""
    M, K = x.shape
    N = w.shape[0]
    z = torch.empty(M, N, device=x.device, dtype=bf16)
    a = torch.empty(M, N, device=x.device, dtype=bf16)
    gemm_raw(x, False, w, False, M, N, K, z, epi=EPI_GELU, bias=bias, out2=a)
    return z, a


FOLD_BIAS_GRAD = os.environ.get("B200MM_FOLD_BIAS_GRAD", "1") != "0"
WGRAD_OVERLAP = os.environ.get("B200MM_WGRAD_OVERLAP", "0") == "1"


class SideQueue:
    """Work that is OFF the backward's critical chain -- weight-gradient GEMMs and bias-gradient column sums -- issued on
    a side stream (a parallel branch of the step's CUDA graph): the data-gradient chain does not wait for it.  What it is
    worth depends on the batch: at the reference's batch 16 the graph replays in 5.30 instead of 5.62 ms (batch 8:
    4.56 vs 4.90, batch 32: 6.78 vs 7.05), at batch 256 it is neutral on top of the two-stream tower overlap (29.93 vs
    29.92 ms; config 3: 54.7 vs 54.5) -- every persistent GEMM CTA takes a whole SM's shared memory, so another branch
    only ever fills kernel-boundary gaps.  Opt-in (B200MM_WGRAD_OVERLAP=1): with it on by default one run of the full
    GPU suite failed test_head_three_tower_model_matches_oracle (not reproduced in four reruns of that file, cause not
    found before the GPU budget ran out), so the default stays the configuration every full run has passed with.  ``run(fn, *tensors)`` orders ``fn`` after everything
    issued so far on the current stream; ``tensors`` are the operands it reads (kept alive for the side stream);
    ``join()`` makes the current stream wait for all of it (before the gradients are announced / consumed)."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device) if WGRAD_OVERLAP else None

    def run(self, fn, *tensors):
        if self.stream is None:
            fn()
            return
        main = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(main)
        for t in tensors:
            t.record_stream(self.stream)
        with torch.cuda.stream(self.stream):
            fn()

    def join(self):
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)


def linear_dgrad(dy, w, *, residual=None, gelu_z=None, out=None, residual_mask=None, bias_grad=None):
    """dx = dy @ w (+residual), or dx = (dy @ w) * gelu'(gelu_z).  dy [M,N], w [N,K] -> dx [M,K].
    residual_mask (uint8 [M, K/8], 1 bit per element, from batchnorm_fwd(want_mask=True)): only the residual elements
    whose bit is set are added -- the identity-branch gradient dout o relu_mask of a residual block, never materialised."""
    M, N = dy.shape
    K = w.shape[1]
    if out is None:
        out = torch.empty(M, K, device=dy.device, dtype=bf16)
    if gelu_z is not None:
        # bias_grad (fp32 [K], accumulated): column sums of the result = the bias gradient of the layer that produced z
        if bias_grad is not None and not FOLD_BIAS_GRAD:      # A/B switch: the separate column-sum pass
            gemm_raw(dy, False, w, True, M, K, N, out, epi=EPI_DGELU, aux=gelu_z)
            colsum(out, bias_grad)
            return out
        return gemm_raw(dy, False, w, True, M, K, N, out, epi=EPI_DGELU, aux=gelu_z, col_stats=bias_grad)
    assert bias_grad is None
    if residual_mask is not None:
        assert residual is not None and residual.data_ptr() != out.data_ptr()
        args = ("b200mm_gemm_bf16_maskres", _p(dy), 0, dy.stride(0), _p(w), 1, w.stride(0), M, K, N, _p(residual),
                residual.stride(0), _p(residual_mask), residual_mask.stride(0), _p(out), out.stride(0), _s())
        if _gemm_profile is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.call(*args)
            e1.record()
            _gemm_profile.append((e0, e1, 2.0 * M * N * K, 2.0 * (M * N + N * K + 2 * M * K) + M * K / 8.0, 2.0 * M * K))
        else:
            _lib.call(*args, key=(M, K, N, 0, 1, "maskres", 1))
        return out
    return gemm_raw(dy, False, w, True, M, K, N, out, epi=EPI_STORE, residual=residual)


def _wgrad_splits(n_out, k_out, m_red):
    """Split-K factor of a weight-gradient GEMM (dw[n_out, k_out], reduction over m_red rows): the one that minimises
    (waves of work units over the SMs) x (k iterations per unit), with a small charge per split for the fp32
    red.add traffic of the partial sums.  Work units follow the kernel's own geometry: 128 x 256 tiles on 148 CTAs, or
    256 x 256 tile pairs on 74 clusters where the CTA-pair kernel applies (see b200mm_gemm_bf16)."""
    m_tiles = (n_out + 127) // 128
    bn = 256 if (k_out % 256 == 0 or k_out > 512) else (128 if k_out > 64 else 64)
    n_tiles = (k_out + bn - 1) // bn
    k_iters = (m_red + 63) // 64
    pair = bn == 256 and m_tiles >= 2 and num_sms() % 2 == 0
    units = ((m_tiles + 1) // 2 if pair else m_tiles) * n_tiles
    slots = num_sms() // 2 if pair else num_sms()
    best, best_cost = 1, None
    for s in range(1, max(1, min(k_iters // 8, 4 * slots)) + 1):
        waves = (units * s + slots - 1) // slots
        cost = waves * ((k_iters + s - 1) // s + 6) + 0.5 * s      # + pipeline fill per unit, + partial-sum traffic
        if best_cost is None or cost < best_cost - 1e-9:
            best, best_cost = s, cost
    return best


def linear_wgrad(dy, x, dw):
    """dw[N,K] (fp32) += dy[M,N].T @ x[M,K]   (split-K over M, fp32 red.add accumulation)."""
    M, N = dy.shape
    K = x.shape[1]
    return gemm_raw(dy, True, x, True, N, K, M, dw, epi=EPI_F32_ATOMIC, splits=_wgrad_splits(N, K, M))


def _halo_conv_taken(H, W, C, Cout, ksize, stride, pad) -> bool:
    """Mirror of the dispatch in b200mm_conv_fwd / b200mm_conv_wgrad (csrc/gemm_tcgen05.cu, csrc/conv3x3_c64.cu
    c3_geometry): those launches run conv3x3_c64_*_kernel, not gemm_bf16_kernel, and are kept out of the GEMM
    kernel's roofline profile."""
    if os.environ.get("B200MM_HALO_CONV", "1").startswith("0"):
        return False
    if not (ksize == 3 and stride == 1 and pad == 1 and C == 64 and Cout == 64 and 6 <= W <= 61):
        return False
    R = min(128 // (W + 2), H)
    return (R + 2) * (W + 2) <= 256


def conv_fwd(x, N, H, W, C, w, ksize, stride, pad, *, residual=None, relu=False, out=None, col_stats=None):
    """Implicit-GEMM convolution: x NHWC bf16 [N*H*W, C] (C % 64 == 0), w [Cout, k*k*C] -> ([N*P*Q, Cout], P, Q)."""
    Cout = w.shape[0]
    P, Q = conv_out_hw(H, W, ksize, stride, pad)
    if not (x.is_contiguous() and w.is_contiguous()):
        raise ValueError("conv_fwd: x must be a contiguous NHWC [N*H*W, C] tensor and w a contiguous [Cout, k*k*C] one")
    if out is None:
        out = torch.empty(N * P * Q, Cout, device=x.device, dtype=bf16)
    prof = _gemm_profile
    if prof is not None and residual is None and not relu and out.stride(0) == 64 and \
            _halo_conv_taken(H, W, C, Cout, ksize, stride, pad):
        prof = None
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _lib.call("b200mm_conv_fwd", _p(x), N, H, W, C, _p(w), Cout, ksize, stride, pad, EPI_RELU if relu else EPI_STORE,
              None, _p(residual), residual.stride(0) if residual is not None else 0, _p(out), out.stride(0),
              _p(col_stats), _s(), key=("conv_fwd", N * P * Q, Cout, ksize * ksize * C, stride))
    if prof is not None:
        e1.record()
        M_, K_ = N * P * Q, ksize * ksize * C
        prof.append((e0, e1, 2.0 * M_ * Cout * K_, 2.0 * (N * H * W * C + Cout * K_ + M_ * Cout), 2.0 * M_ * Cout))
    return out, P, Q


def conv_wgrad(dy, x, N, H, W, C, ksize, stride, pad, dw):
    """dw[Cout, k*k*C] (fp32) += dy[N*P*Q, Cout]^T im2col(x)  -- im2col operand gathered by TMA."""
    Cout = dy.shape[1]
    pixels = dy.shape[0]
    if not x.is_contiguous() or dy.stride(1) != 1:
        raise ValueError("conv_wgrad: x must be contiguous NHWC and dy row-major")
    prof = _gemm_profile
    if prof is not None and dy.stride(0) == 64 and _halo_conv_taken(H, W, C, Cout, ksize, stride, pad):
        prof = None
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _lib.call("b200mm_conv_wgrad", _p(dy), dy.stride(0), _p(x), N, H, W, C, Cout, ksize, stride, pad, _p(dw),
              _wgrad_splits(Cout, ksize * ksize * C, pixels), _s(),
              key=("conv_wgrad", Cout, ksize * ksize * C, pixels, stride))
    if prof is not None:
        e1.record()
        K_ = ksize * ksize * C
        prof.append((e0, e1, 2.0 * pixels * Cout * K_, 2.0 * (pixels * Cout + N * H * W * C) + 8.0 * Cout * K_,
                     8.0 * Cout * K_))
    return dw


def conv_weight_rotate_multi(table, n):
    """All weight rotations of a backward pass in one launch; table: device int64 [n, 4] (see the C header)."""
    _lib.call("b200mm_conv_weight_rotate_multi", _p(table), int(n), _s())


def conv_weight_rotate(w, Cout, Cin, ksize, out=None):
    if out is None:
        out = torch.empty(Cin, ksize * ksize * Cout, device=w.device, dtype=bf16)
    _lib.call("b200mm_conv_weight_rotate", _p(w), _p(out), Cout, Cin, ksize, _s())
    return out


def colsum(x, out):
    """out[N] (fp32) += x[M,N].sum(0)"""
    M, N = x.shape
    _lib.call("b200mm_colsum_bf16", _p(x), x.stride(0), M, N, _p(out), _s())
    return out


# ----------------------------------------------------------------------------------------------- attention
def mask_to_bias(mask):
    """int64 attention_mask (1 = token) -> fp32 additive key bias (0 / -inf)."""
    _chk(mask, torch.int64, "attention_mask")
    mask = mask.contiguous()
    bias = torch.empty(mask.shape, device=mask.device, dtype=f32)
    _lib.call("b200mm_mask_to_bias", _p(mask), _p(bias), mask.numel(), _s())
    return bias


def attention_fwd(qkv, key_bias, B, H, S, *, p_drop=0.0, seed=0, need_lse=True, save_mask=False):
    """save_mask (training, S <= 128, p_drop > 0): also returns the dropout keep bits ([B*H, 4, 128] int32) for
    ``attention_bwd(..., drop_mask=)`` -- the backward then reads 2 KB per head instead of regenerating the Philox
    stream.  Returns (out, lse) or (out, lse, mask)."""
    out = torch.empty(B * S, H * 64, device=qkv.device, dtype=bf16)
    lse = torch.empty(B, H, S, device=qkv.device, dtype=f32) if need_lse else None
    if save_mask:
        mask = torch.empty(B * H, 4, 128, device=qkv.device, dtype=torch.int32) if (S <= 128 and p_drop > 0) else None
        _lib.call("b200mm_attention_fwd_mask", _p(qkv), _p(key_bias), _p(out), _p(lse), B, H, S, float(p_drop),
                  int(seed), _p(mask), _s())
        return out, lse, mask
    _lib.call("b200mm_attention_fwd", _p(qkv), _p(key_bias), _p(out), _p(lse), B, H, S, float(p_drop), int(seed), _s())
    return out, lse


def attention_bwd(qkv, key_bias, out, dout, lse, B, H, S, *, p_drop=0.0, seed=0, drop_mask=None):
    dqkv = torch.empty_like(qkv)
    if drop_mask is not None:
        _lib.call("b200mm_attention_bwd_mask", _p(qkv), _p(key_bias), _p(out), _p(dout), _p(lse), _p(dqkv), B, H, S,
                  float(p_drop), int(seed), _p(drop_mask), _s())
        return dqkv
    _lib.call("b200mm_attention_bwd", _p(qkv), _p(key_bias), _p(out), _p(dout), _p(lse), _p(dqkv), B, H, S,
              float(p_drop), int(seed), _s())
    return dqkv


# ----------------------------------------------------------------------------------------------- norms / embeddings
def layernorm_fwd(x, gamma, beta, eps, *, p_drop=0.0, seed=0):
    M, D = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(M, device=x.device, dtype=f32)
    rstd = torch.empty(M, device=x.device, dtype=f32)
    _lib.call("b200mm_layernorm_fwd", _p(x), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd), M, D, float(eps),
              float(p_drop), int(seed), _s())
    return y, mean, rstd


def embed_layernorm_fwd(ids, word, pos, gamma, beta, eps, *, p_drop=0.0, seed=0, pos_ids=None, type_row=None):
    B, S = ids.shape
    V, D = word.shape
    M = B * S
    x_saved = torch.empty(M, D, device=ids.device, dtype=bf16)
    y = torch.empty(M, D, device=ids.device, dtype=bf16)
    mean = torch.empty(M, device=ids.device, dtype=f32)
    rstd = torch.empty(M, device=ids.device, dtype=f32)
    _lib.call("b200mm_embed_layernorm_fwd", _p(ids), _p(word), _p(pos), _p(pos_ids), _p(type_row), S, V, _p(gamma),
              _p(beta), _p(x_saved), _p(y),
              _p(mean), _p(rstd), M, D, float(eps), float(p_drop), int(seed), _s())
    return y, x_saved, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, dgamma, dbeta, *, p_in=0.0, seed_in=0, p_out=0.0, seed_out=0,
                  addend=None):
    """Returns (dx, dx_masked) where dx_masked is None unless p_out > 0.  ``addend`` (bf16 [M,D]) is added to dx
    only (the residual-stream gradient of a pre-LN block)."""
    M, D = x.shape
    dx = torch.empty_like(x)
    dx2 = torch.empty_like(x) if p_out > 0 else None
    _lib.call("b200mm_layernorm_bwd", _p(dy), _p(x), _p(mean), _p(rstd), _p(gamma), _p(addend), _p(dx), _p(dx2),
              _p(dgamma),
              _p(dbeta), M, D, float(p_in), int(seed_in), float(p_out), int(seed_out), _s())
    return dx, dx2


def embedding_bwd(dx, ids, dword, dpos, padding_idx=-1, *, pos_ids=None, pos_padding_idx=-1):
    B, S = ids.shape
    V, D = dword.shape
    _lib.call("b200mm_embedding_bwd", _p(dx), _p(ids), _p(pos_ids), int(pos_padding_idx), S, V, int(padding_idx),
              _p(dword), _p(dpos), B * S, D, _s())


def position_ids(ids, pad_id):
    """RoBERTa / XLM-R position ids (int32 [B, S]): cumsum(ids != pad) * (ids != pad) + pad."""
    _chk(ids, torch.int64, "input_ids")
    B, S = ids.shape
    out = torch.empty(B, S, device=ids.device, dtype=torch.int32)
    _lib.call("b200mm_position_ids", _p(ids), int(pad_id), B, S, _p(out), _s())
    return out


def vit_assemble_fwd(patch, cls, pos, B, P):
    D = patch.shape[1]
    x = torch.empty(B * (P + 1), D, device=patch.device, dtype=bf16)
    _lib.call("b200mm_vit_assemble_fwd", _p(patch), _p(cls), _p(pos), _p(x), B, P, D, _s())
    return x


def vit_assemble_bwd(dx, dcls, dpos, B, P):
    D = dx.shape[1]
    dpatch = torch.empty(B * P, D, device=dx.device, dtype=bf16)
    _lib.call("b200mm_vit_assemble_bwd", _p(dx), _p(dpatch), _p(dcls), _p(dpos), B, P, D, _s())
    return dpatch


def gather_rows(x, rows, stride_rows, offset_rows, *, p_drop=0.0, seed=0):
    D = x.shape[1]
    out = torch.empty(rows, D, device=x.device, dtype=bf16)
    _lib.call("b200mm_gather_rows", _p(x), _p(out), rows, D, stride_rows, offset_rows, float(p_drop), int(seed), _s())
    return out


def scatter_rows(dpooled, M, stride_rows, offset_rows, *, p_drop=0.0, seed=0):
    D = dpooled.shape[1]
    dx = torch.empty(M, D, device=dpooled.device, dtype=bf16)
    _lib.call("b200mm_scatter_rows", _p(dpooled), _p(dx), M, D, stride_rows, offset_rows, float(p_drop), int(seed),
              _s())
    return dx


# ----------------------------------------------------------------------------------------------- head / loss / optim
def head_loss(feat, W, bias, labels, *, loss_kind=LOSS_CE, alpha=0.25, gamma=2.0, train=True, dW=None, dbias=None,
              dlogits=None):
    """Output layer + loss (+ its backward when train).  Returns (logits fp32 [B,C], loss_sum fp32 [1],
    correct int32 [1], dfeat bf16 [B,F] | None)."""
    B, F = feat.shape
    C = W.shape[0]
    dev = feat.device
    logits = torch.empty(B, C, device=dev, dtype=f32)
    loss = torch.zeros(1, device=dev, dtype=f32)
    correct = torch.zeros(1, device=dev, dtype=torch.int32)
    dfeat = torch.empty(B, F, device=dev, dtype=bf16) if train else None
    _lib.call("b200mm_head_loss", _p(feat), _p(W), _p(bias), _p(labels), B, F, C, loss_kind, float(alpha),
              float(gamma), int(train), _p(dlogits), _p(logits), _p(loss), _p(correct), _p(dfeat), _p(dW), _p(dbias),
              _s())
    return logits, loss, correct, dfeat


# ----------------------------------------------------------------------------------------------- HEAD-script head
def bn1d_fwd(x, gamma, beta, running_mean, running_var, *, relu=False, train=True, eps=1e-5, momentum=0.1, out=None):
    """BatchNorm1d (+ReLU) over x [B, C] bf16 (may be a column slice).  Returns (out, mean, rstd)."""
    B, C = x.shape
    if out is None:
        out = torch.empty(B, C, device=x.device, dtype=bf16)
    mean = torch.empty(C, device=x.device, dtype=f32) if train else None
    rstd = torch.empty(C, device=x.device, dtype=f32) if train else None
    _lib.call("b200mm_bn1d_fwd", _p(x), int(x.dtype == f32), x.stride(0), B, C, _p(gamma), _p(beta), float(eps), float(momentum), int(relu),
              int(train), _p(out), out.stride(0), _p(mean), _p(rstd), _p(running_mean), _p(running_var), _s())
    return out, mean, rstd


def bn1d_bwd(dout, out, x, mean, rstd, gamma, dgamma, dbeta, *, relu=False):
    B, C = x.shape
    dx = torch.empty(B, C, device=x.device, dtype=bf16)
    _lib.call("b200mm_bn1d_bwd", _p(dout), dout.stride(0), _p(out), out.stride(0) if out is not None else 0, _p(x),
              int(x.dtype == f32), x.stride(0), B, C, _p(mean), _p(rstd), _p(gamma), int(relu), _p(dx), dx.stride(0), _p(dgamma),
              _p(dbeta), _s())
    return dx


def softmax_gate_fwd(a, x):
    """y = softmax(a, dim=1) * x.  Returns (y bf16, w fp32)."""
    B, C = a.shape
    w = torch.empty(B, C, device=a.device, dtype=f32)
    y = torch.empty(B, C, device=a.device, dtype=bf16)
    _lib.call("b200mm_softmax_gate_fwd", _p(a), _p(x), B, C, _p(w), _p(y), _s())
    return y, w


def softmax_gate_bwd(dy, w, x):
    """Returns (da, dx_direct)."""
    B, C = w.shape
    da = torch.empty(B, C, device=w.device, dtype=bf16)
    dxd = torch.empty(B, C, device=w.device, dtype=bf16)
    _lib.call("b200mm_softmax_gate_bwd", _p(dy), _p(w), _p(x), B, C, _p(da), _p(dxd), _s())
    return da, dxd


def relu_bwd(dy, y):
    dx = torch.empty_like(y)
    _lib.call("b200mm_relu_bwd", _p(dy), _p(y), y.numel(), _p(dx), _s())
    return dx


def head_bn_focal(feat, W, bias, bn_g, bn_b, running_mean, running_var, labels, *, alpha=0.25, gamma=2.0, train=True,
                  bn_train=True, eps=1e-5, momentum=0.1, dlogits=None, dW=None, dbias=None, dg=None, dbeta=None):
    """Linear(F,1) + BatchNorm1d(1) + sigmoid focal loss (+ backward when train).
    Returns (logits fp32 [B], loss fp32 [1], correct int32 [1], dfeat bf16 [B,F] | None)."""
    B, F = feat.shape
    dev = feat.device
    logits = torch.empty(B, device=dev, dtype=f32)
    loss = torch.zeros(1, device=dev, dtype=f32)
    correct = torch.zeros(1, device=dev, dtype=torch.int32)
    dfeat = torch.empty(B, F, device=dev, dtype=bf16) if train else None
    _lib.call("b200mm_head_bn_focal", _p(feat), _p(W), _p(bias), _p(bn_g), _p(bn_b), _p(running_mean),
              _p(running_var), _p(labels), B, F, float(eps), float(momentum), float(alpha), float(gamma), int(train),
              int(bn_train), _p(dlogits), _p(logits), _p(loss), _p(correct), _p(dfeat), _p(dW), _p(dbias), _p(dg),
              _p(dbeta), _s())
    return logits, loss, correct, dfeat


def sumsq(g, out):
    """out[0] += sum(g^2); g fp32 or bf16 (the all-reduced data-parallel payload)."""
    _lib.call("b200mm_sumsq_bf16" if g.dtype == bf16 else "b200mm_sumsq_f32", _p(g), g.numel(), _p(out), _s())
    return out


def adam_step(p, g, m, v, shadow, *, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, step=1, gradsq=None,
              max_norm=0.0, grad_scale=1.0):
    _lib.call("b200mm_adam_step_g16" if g.dtype == bf16 else "b200mm_adam_step",
              _p(p), _p(g), _p(m), _p(v), _p(shadow), p.numel(), float(lr), float(beta1),
              float(beta2), float(eps), float(weight_decay), int(step), _p(gradsq), float(max_norm),
              float(grad_scale), _s())


def dwconv7x7(x, wt, bias, N, H, W, C):
    """Depthwise 7x7 / pad 3 convolution on an NHWC bf16 activation [N*H*W, C]; wt bf16 [49, C], bias fp32 [C]."""
    _chk(x, bf16, "x"); _chk(wt, bf16, "weight")
    y = torch.empty_like(x)
    _lib.call("b200mm_dwconv7x7_nhwc", _p(x), _p(wt), _p(bias), _p(y), N, H, W, C, _s())
    return y


def tanh_(x):
    _chk(x, f32, "x")
    _lib.call("b200mm_tanh_f32", _p(x), x.numel(), _s())
    return x


def adam_step_dyn(p, g, m, v, shadow, *, lr_dev, step_dev, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0,
                  gradsq=None, max_norm=0.0, grad_scale=1.0):
    """adam_step with the learning rate (fp32 device scalar) and step count (int32 device scalar) read on the device:
    the launch can be captured in a CUDA graph and replayed while both change."""
    _lib.call("b200mm_adam_step_dyn", _p(p), _p(g), int(g.dtype == bf16), _p(m), _p(v), _p(shadow), p.numel(),
              _p(lr_dev), float(beta1), float(beta2), float(eps), float(weight_decay), _p(step_dev), _p(gradsq),
              float(max_norm), float(grad_scale), _s())


def set_step_salt(salt):
    """salt: uint64-sized device tensor (int64 [1]) added to every dropout seed by the kernels, or None."""
    _lib.call("b200mm_set_step_salt_ptr", _p(salt))


def step_advance(salt, adam_step):
    _lib.call("b200mm_step_advance", _p(salt), _p(adam_step), _s())


def cast_to_bf16(x, y, scale: float = 1.0):
    if scale == 1.0:
        _lib.call("b200mm_cast_f32_to_bf16", _p(x), _p(y), x.numel(), _s())
    else:
        _lib.call("b200mm_scale_cast_f32_to_bf16", _p(x), _p(y), x.numel(), float(scale), _s())
    return y


# ----------------------------------------------------------------------------------------------- image tower pieces
class BNScratch:
    """fp32 workspace (>= 18*C + 32 floats) shared by every BatchNorm launch of one device/stream."""
    _buf = {}

    @classmethod
    def get(cls, device, C):
        key = (device.index, torch.cuda.current_stream().cuda_stream)
        b = cls._buf.get(key)
        if b is None or b.numel() < 18 * C + 32:
            b = torch.empty(max(18 * C + 32, 18 * 2048 + 32), device=device, dtype=f32)
            cls._buf[key] = b
        return b


def batchnorm_fwd(x, gamma, beta, running_mean, running_var, *, residual=None, relu=True, eps=1e-5, momentum=0.1,
                  col_stats=None, want_mask=False):
    """Train-mode BatchNorm.  col_stats (fp32 [2C] = sums | sums of squares, from the producing convolution's
    epilogue): skip the statistics pass."""
    M, C = x.shape
    out = torch.empty_like(x)
    mean = torch.empty(C, device=x.device, dtype=f32)
    rstd = torch.empty(C, device=x.device, dtype=f32)
    if col_stats is not None:
        # want_mask: also emit the 1-bit ReLU mask [M, C/8] (returned as a 4th value) for the backward
        mask = torch.empty(M, C // 8, device=x.device, dtype=torch.uint8) if (want_mask and relu) else None
        _lib.call("b200mm_batchnorm_fwd_stats", _p(x), _p(residual), M, C, _p(col_stats), _p(gamma), _p(beta),
                  float(eps), float(momentum), int(relu), _p(out), _p(mean), _p(rstd), _p(running_mean),
                  _p(running_var), _p(mask), _s(), key=("bn_fwd", M, C, int(residual is not None), int(mask is not None)))
        return (out, mean, rstd, mask) if want_mask else (out, mean, rstd)
    _lib.call("b200mm_batchnorm_fwd", _p(x), _p(residual), M, C, _p(gamma), _p(beta), float(eps), float(momentum),
              int(relu), _p(out), _p(mean), _p(rstd), _p(running_mean), _p(running_var),
              _p(BNScratch.get(x.device, C)), _s())
    return out, mean, rstd


def batchnorm_eval(x, gamma, beta, running_mean, running_var, *, residual=None, relu=True, eps=1e-5):
    M, C = x.shape
    out = torch.empty_like(x)
    _lib.call("b200mm_batchnorm_eval", _p(x), _p(residual), M, C, _p(gamma), _p(beta), _p(running_mean),
              _p(running_var), float(eps), int(relu), _p(out), _s())
    return out


BN_FUSED_MB = [0]     # mirror of b200mm_tune knob 2 (set both through set_bn_fused_mb)
_bn_fused_env = [os.environ.get("B200MM_BN_FUSED_MB")]   # A/B switch, applied at the first call


def set_bn_fused_mb(mb: int) -> None:
    """Largest tensor (MB) whose BatchNorm backward runs as one cooperative launch; 0 = always the two-kernel path."""
    _lib.load().b200mm_tune(2, int(mb))
    BN_FUSED_MB[0] = int(mb)


def batchnorm_bwd(dout, out, x, mean, rstd, gamma, dgamma, dbeta, *, relu=True, need_dz=False, beta=None, mask=None):
    """ReLU mask: ``mask`` (1 bit / element, from batchnorm_fwd(want_mask=True)), else ``out``, else recomputed from x
    (pass beta; only valid without a residual)."""
    M, C = x.shape
    if _bn_fused_env[0] is not None:
        set_bn_fused_mb(int(_bn_fused_env[0]))
        _bn_fused_env[0] = None
    dx = torch.empty_like(x)
    dz = torch.empty_like(x) if need_dz else None
    if (out is None or not relu or mask is not None) and M * C * 2 <= (BN_FUSED_MB[0] << 20):
        _lib.LAUNCHES[0] -= 1      # one cooperative launch instead of reduce + apply (mirror of b200mm_batchnorm_bwd)
    _lib.call("b200mm_batchnorm_bwd", _p(dout), _p(out), _p(x), M, C, _p(mean), _p(rstd), _p(gamma), _p(beta),
              _p(mask), int(relu),
              _p(dx), _p(dz), _p(dgamma), _p(dbeta), _p(BNScratch.get(x.device, C)), _s(),
              key=("bn_bwd", M, C, int(need_dz), 0 if mask is None else 1))
    return dx, dz


def maxpool_fwd(x, N, H, W, C):
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = torch.empty(N * Ho * Wo, C, device=x.device, dtype=bf16)
    arg = torch.empty(N * Ho * Wo, C, device=x.device, dtype=torch.uint8)
    _lib.call("b200mm_maxpool3x3s2_fwd", _p(x), N, H, W, C, _p(out), _p(arg), _s())
    return out, arg, Ho, Wo


def bn_relu_maxpool_fwd(x, N, H, W, C, gamma, beta, running_mean, running_var, col_stats, *, eps=1e-5, momentum=0.1):
    """maxpool3x3s2(relu(BN_train(x))) in one pass from the convolution output x [N*H*W, C] and its column statistics;
    returns (pooled, argmax, Ho, Wo, mean, rstd) -- the normalised activation itself is never written."""
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = torch.empty(N * Ho * Wo, C, device=x.device, dtype=bf16)
    arg = torch.empty(N * Ho * Wo, C, device=x.device, dtype=torch.uint8)
    mean = torch.empty(C, device=x.device, dtype=f32)
    rstd = torch.empty(C, device=x.device, dtype=f32)
    _lib.call("b200mm_bn_relu_maxpool_fwd", _p(x), N, H, W, C, _p(col_stats), _p(gamma), _p(beta), float(eps),
              float(momentum), _p(out), _p(arg), _p(mean), _p(rstd), _p(running_mean), _p(running_var), _s())
    return out, arg, Ho, Wo, mean, rstd


def maxpool_bwd(dout, arg, N, H, W, C):
    dx = torch.empty(N * H * W, C, device=dout.device, dtype=bf16)
    _lib.call("b200mm_maxpool3x3s2_bwd", _p(dout), _p(arg), N, H, W, C, _p(dx), _s())
    return dx


def avgpool_fwd(x, N, HW, C):
    out = torch.empty(N, C, device=x.device, dtype=bf16)
    _lib.call("b200mm_avgpool_fwd", _p(x), N, HW, C, _p(out), _s())
    return out


def avgpool_bwd(dout, N, HW, C):
    dx = torch.empty(N * HW, C, device=dout.device, dtype=bf16)
    _lib.call("b200mm_avgpool_bwd", _p(dout), N, HW, C, _p(dx), _s())
    return dx


def conv_out_hw(H, W, k, stride, pad):
    return (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1


def im2col(x, N, H, W, C, k, stride, pad):
    Ho, Wo = conv_out_hw(H, W, k, stride, pad)
    cols = torch.empty(N * Ho * Wo, k * k * C, device=x.device, dtype=bf16)
    _lib.call("b200mm_im2col_nhwc", _p(x), N, H, W, C, k, k, stride, pad, _p(cols), _s())
    return cols, Ho, Wo


def col2im(dcols, N, H, W, C, k, stride, pad, addend=None):
    dx = torch.empty(N * H * W, C, device=dcols.device, dtype=bf16)
    _lib.call("b200mm_col2im_nhwc", _p(dcols), _p(addend), N, H, W, C, k, k, stride, pad, _p(dx), _s())
    return dx


def im2col_nchw_f32(img, k, stride, pad, Kp):
    _chk(img, f32, "pixel_values")
    img = img.contiguous()
    N, Cin, H, W = img.shape
    Ho, Wo = conv_out_hw(H, W, k, stride, pad)
    cols = torch.empty(N * Ho * Wo, Kp, device=img.device, dtype=bf16)
    _lib.call("b200mm_im2col_nchw_f32", _p(img), N, Cin, H, W, k, k, stride, pad, Kp, _p(cols), _s())
    return cols, Ho, Wo


def stem_conv_supported(img, w) -> bool:
    """Shapes the direct stem kernels are specialised for (csrc/stem_conv.cu: 7x7/2/3, 3 -> 64, K padded to 152)."""
    N, Cin, H, W = img.shape
    return Cin == 3 and tuple(w.shape) == (64, 152) and W <= 226 and 8 <= (W - 1) // 2 + 1 <= 128 and H >= 7


def stem_conv_fwd(img, w, col_stats=None):
    """conv1 (7x7 / 2 / pad 3) of the fp32 NCHW image -> bf16 NHWC [N*Ho*Wo, 64], no im2col matrix
    (torchvision/models/resnet.py:197).  col_stats: fp32 [128] accumulator of column sum / sum of squares."""
    _chk(img, f32, "pixel_values")
    _chk(w, bf16, "conv1.weight")
    img = img.contiguous()
    N, Cin, H, W = img.shape
    Ho, Wo = conv_out_hw(H, W, 7, 2, 3)
    out = torch.empty(N * Ho * Wo, w.shape[0], device=img.device, dtype=bf16)
    _lib.call("b200mm_stem_conv_fwd", _p(img), N, Cin, H, W, _p(w), w.shape[0], w.shape[1], _p(out),
              _p(col_stats) if col_stats is not None else None, _s())
    return out, Ho, Wo


def stem_conv_wgrad(img, dy, dw):
    """dw[64,152] (fp32) += dy^T . patches(img) for conv1, patches gathered on the fly."""
    _chk(img, f32, "pixel_values")
    _chk(dy, bf16, "dy")
    _chk(dw, f32, "dw")
    img = img.contiguous()
    N, Cin, H, W = img.shape
    _lib.call("b200mm_stem_conv_wgrad", _p(img), N, Cin, H, W, _p(dy), dw.shape[0], dw.shape[1], _p(dw), _s())


def subsample(x, N, H, W, C, stride):
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    out = torch.empty(N * Ho * Wo, C, device=x.device, dtype=bf16)
    _lib.call("b200mm_subsample_nhwc", _p(x), N, H, W, C, stride, _p(out), _s())
    return out, Ho, Wo


def upsample_add(dsub, addend, N, H, W, C, stride):
    dx = torch.empty(N * H * W, C, device=dsub.device, dtype=bf16)
    _lib.call("b200mm_upsample_add_nhwc", _p(dsub), _p(addend), N, H, W, C, stride, _p(dx), _s())
    return dx


# ----------------------------------------------------------------------------------------------- preprocessing
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def _c3(v):
    import ctypes
    return ctypes.cast((ctypes.c_float * 3)(*v), ctypes.c_void_p)


def pack_images(images, pin: bool = True):
    """Host side of the packed image batch: a list of uint8 CPU tensors [H, W, 3] (decoded, any size) -> ONE uint8
    buffer (pinned) + an int64 table [3, n] = (byte offset | height | width), so a batch crosses PCIe in two copies."""
    n = len(images)
    table = torch.empty(3, n, dtype=torch.int64)
    off = 0
    for i, im in enumerate(images):
        if im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3:
            raise ValueError("images must be [H, W, 3] uint8")
        table[0, i], table[1, i], table[2, i] = off, im.shape[0], im.shape[1]
        off += (im.numel() + 15) // 16 * 16
    buf = torch.empty(off, dtype=torch.uint8, pin_memory=pin)
    for i, im in enumerate(images):
        o = int(table[0, i])
        buf[o:o + im.numel()].copy_(im.reshape(-1))
    if pin:
        table = table.pin_memory()
    return buf, table


def preprocess_u8_packed(packed, table, *, resize=256, crop=224, square=False, flip=None, mean=IMAGENET_MEAN,
                         std=IMAGENET_STD, resample="float"):
    """packed: uint8 CUDA buffer, table: int64 CUDA [3, n] (offset | height | width) from ``pack_images``.
    Returns fp32 [n, 3, crop, crop]: Resize(resize) -> CenterCrop(crop) (square=False, .txt:37-41) or
    Resize((crop, crop)) (square=True, HEAD script :224), optional per-image horizontal flip (uint8 flags [n]),
    then ToTensor -> Normalize -- antialiased bilinear like torchvision's tensor path (``resample='float'``), or with
    Pillow's own 8-bit two-pass arithmetic (``resample='pillow'``: bit-identical to what the reference's Dataset computes
    on the PIL image; sides shrinking by more than 31x are not supported)."""
    if resample not in ("float", "pillow"):
        raise ValueError("resample must be 'float' or 'pillow'")
    _chk(packed, torch.uint8, "packed images")
    n = table.shape[1]
    hw = table[1:3].to(torch.int32).contiguous()            # [2, n] int32 heights | widths (tiny device op)
    out = torch.empty(n, 3, crop, crop, device=packed.device, dtype=f32)
    entry = "b200mm_preprocess_u8_packed_pil" if resample == "pillow" else "b200mm_preprocess_u8_packed"
    _lib.call(entry, _p(packed), _p(table[0]), _p(hw[0]), _p(hw[1]), _p(flip), n, int(resize),
              int(crop), int(square), _c3(mean), _c3(std), _p(out), _s())
    return out


def preprocess_u8_packed_pil_u8(packed, table, *, resize=256, crop=224, square=False, flip=None):
    """Pillow-exact Resize / CenterCrop (or square resize) / flip of a packed batch, kept as uint8 [n, crop, crop, 3] --
    the PIL image as it enters the script's ColorJitter (``augment_pil``)."""
    _chk(packed, torch.uint8, "packed images")
    n = table.shape[1]
    hw = table[1:3].to(torch.int32).contiguous()
    out = torch.empty(n, crop, crop, 3, device=packed.device, dtype=torch.uint8)
    _lib.call("b200mm_preprocess_u8_packed_pil_u8", _p(packed), _p(table[0]), _p(hw[0]), _p(hw[1]), _p(flip), n,
              int(resize), int(crop), int(square), _p(out), _s())
    return out


def augment_pil(img_u8, order, alpha, hue, affine, *, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """ColorJitter + RandomRotation + ToTensor + Normalize with Pillow's own uint8 arithmetic (csrc/augment_pil.cu).
    img_u8: uint8 CUDA [n, H, W, 3]; order int32 [n]; alpha fp32 [n, 3]; hue int32 [n]; affine int32 [n, 6]
    (data.GpuImageTransform.pack_augment_pil builds the four tables).  Returns fp32 [n, 3, H, W]."""
    _chk(img_u8, torch.uint8, "img_u8")
    _chk(order, torch.int32, "order")
    _chk(alpha, f32, "alpha")
    _chk(hue, torch.int32, "hue")
    _chk(affine, torch.int32, "affine")
    if img_u8.dim() != 4 or img_u8.shape[3] != 3 or not img_u8.is_contiguous():
        raise ValueError("img_u8 must be a contiguous [n, H, W, 3] uint8 tensor")
    n, H, W, _ = img_u8.shape
    if tuple(order.shape) != (n,) or tuple(alpha.shape) != (n, 3) or tuple(hue.shape) != (n,) or \
            tuple(affine.shape) != (n, 6) or not alpha.is_contiguous() or not affine.is_contiguous():
        raise ValueError("order / hue must be [n], alpha a contiguous [n, 3], affine a contiguous [n, 6]")
    out = torch.empty(n, 3, H, W, device=img_u8.device, dtype=f32)
    sums = torch.empty(n, device=img_u8.device, dtype=torch.int64)
    _lib.call("b200mm_augment_pil", _p(img_u8), _p(order), _p(alpha), _p(hue), _p(affine), n, H, W, _c3(mean), _c3(std),
              _p(sums), _p(out), _s())
    return out


def u8_normalize(images, *, flip=None, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """images: uint8 CUDA [n, H, W, 3] already at network resolution -> fp32 [n, 3, H, W] = Normalize(ToTensor(img)),
    optional per-image horizontal flip (uint8 flags [n])."""
    _chk(images, torch.uint8, "images")
    if images.dim() != 4 or images.shape[3] != 3 or not images.is_contiguous():
        raise ValueError("images must be a contiguous [n, H, W, 3] uint8 tensor")
    n, H, W, _ = images.shape
    out = torch.empty(n, 3, H, W, device=images.device, dtype=f32)
    _lib.call("b200mm_u8_normalize_nchw", _p(images), _p(flip), n, H, W, _c3(mean), _c3(std), _p(out), _s())
    return out


def augment_jitter_rotate(img01, order, params, *, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """ColorJitter + RandomRotation + Normalize of the HEAD script's train transform (.py:224-233) over a batch already
    resized / flipped / scaled to [0, 1].  img01: fp32 CUDA [n, 3, H, W]; order: int32 CUDA [n] (2 bits per operator,
    first applied in the low bits; 0 brightness, 1 contrast, 2 saturation, 3 hue); params: fp32 CUDA [n, 8] (the three
    factors, the hue shift, the inverse rotation matrix m00 m01 m10 m11).  Returns (normalised fp32 [n, 3, H, W],
    per-image contrast means [n]).  torchvision's float-tensor semantics (csrc/augment_math.cuh)."""
    _chk(img01, f32, "img01")
    _chk(order, torch.int32, "order")
    _chk(params, f32, "params")
    if img01.dim() != 4 or img01.shape[1] != 3 or not img01.is_contiguous():
        raise ValueError("img01 must be a contiguous [n, 3, H, W] fp32 tensor")
    n, _, H, W = img01.shape
    if tuple(order.shape) != (n,) or tuple(params.shape) != (n, 8) or not params.is_contiguous():
        raise ValueError("order must be [n] and params a contiguous [n, 8]")
    out = torch.empty_like(img01)
    scratch = torch.empty(9 * n, device=img01.device, dtype=f32)      # [n] means | [n, 8] partial sums
    _lib.call("b200mm_augment_jitter_rotate", _p(img01), _p(order), _p(params), n, H, W, _c3(mean), _c3(std),
              _p(scratch), _p(out), _s())
    return out, scratch[:n]


def preprocess_u8(images, *, resize=256, crop=224, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """images: list of uint8 CUDA tensors [H, W, 3] (decoded, any size).  Returns fp32 [n, 3, crop, crop] exactly as
    Resize(resize) -> CenterCrop(crop) -> ToTensor -> Normalize(mean, std) would (antialiased bilinear).
    (Device-resident inputs; a loader that starts from host images uses pack_images + preprocess_u8_packed.)"""
    n = len(images)
    dev = images[0].device
    imgs = [im.contiguous() for im in images]
    for im in imgs:
        _chk(im, torch.uint8, "image")
        if im.dim() != 3 or im.shape[2] != 3:
            raise ValueError("images must be [H, W, 3] uint8")
    # one pinned table [3, n] (pointer | height | width) and one asynchronous copy instead of three pageable ones
    table = torch.tensor([[im.data_ptr() for im in imgs], [im.shape[0] for im in imgs], [im.shape[1] for im in imgs]],
                         dtype=torch.int64).pin_memory().to(dev, non_blocking=True)
    hw = table[1:3].to(torch.int32).contiguous()
    out = torch.empty(n, 3, crop, crop, device=dev, dtype=f32)
    _lib.call("b200mm_preprocess_u8", _p(table[0]), _p(hw[0]), _p(hw[1]), n, resize, crop, _c3(mean), _c3(std),
              _p(out), _s())
    out._keepalive = (imgs, table)
    return out
