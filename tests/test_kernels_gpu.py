"""Per-kernel parity on the B200: every C-ABI kernel vs a plain PyTorch fp32 reference of the same op."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

bf16 = torch.bfloat16


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


@pytest.fixture(scope="module")
def ops(cuda_device):
    from b200mm import ops as o
    return o


# ------------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (384, 768, 768), (1000, 512, 200), (130, 64, 152), (256, 1000, 2048)])
def test_gemm_forward_epilogues(ops, cuda_device, M, N, K):
    torch.manual_seed(0)
    x = torch.randn(M, K, device=cuda_device).to(bf16)
    w = torch.randn(N, K, device=cuda_device).to(bf16)
    b = torch.randn(N, device=cuda_device)
    res = torch.randn(M, N, device=cuda_device).to(bf16)
    ref = x.float() @ w.float().t() + b
    assert rel(ops.linear_fwd(x, w, b), ref) < 1e-2
    assert rel(ops.linear_fwd(x, w, b, residual=res), ref + res.float()) < 1e-2
    assert rel(ops.linear_fwd(x, w, b, relu=True), torch.relu(ref)) < 1e-2
    xs = (x.float() * 0.05).to(bf16)
    refs = xs.float() @ w.float().t() + b
    z, a = ops.linear_gelu_fwd(xs, w, b)
    assert rel(z, refs) < 1e-2 and rel(a, F.gelu(refs)) < 1e-2


@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (4096, 768, 3072), (1000, 512, 200)])
def test_gemm_dgrad_wgrad(ops, cuda_device, M, N, K):
    torch.manual_seed(1)
    x = torch.randn(M, K, device=cuda_device).to(bf16)
    w = torch.randn(N, K, device=cuda_device).to(bf16)
    dy = torch.randn(M, N, device=cuda_device).to(bf16)
    res = torch.randn(M, K, device=cuda_device).to(bf16)
    assert rel(ops.linear_dgrad(dy, w), dy.float() @ w.float()) < 1e-2
    assert rel(ops.linear_dgrad(dy, w, residual=res), dy.float() @ w.float() + res.float()) < 1e-2
    acc = res.clone()                                            # residual aliasing the output: in-place reduce-add
    got = ops.linear_dgrad(dy, w, residual=acc, out=acc)
    assert got.data_ptr() == acc.data_ptr() and rel(acc, dy.float() @ w.float() + res.float()) < 1e-2
    z = torch.randn(M, K, device=cuda_device).to(bf16)
    zf = z.float().requires_grad_(True)
    F.gelu(zf).sum().backward()
    assert rel(ops.linear_dgrad(dy, w, gelu_z=z), (dy.float() @ w.float()) * zf.grad) < 1e-2
    # the same epilogue accumulating the bias gradient of the layer behind z: column sums of the STORED result
    dbz = torch.full((K,), 0.5, device=cuda_device)
    dz = ops.linear_dgrad(dy, w, gelu_z=z, bias_grad=dbz)
    assert torch.equal(dz, ops.linear_dgrad(dy, w, gelu_z=z))
    ref_db = dz.float().sum(0) + 0.5
    assert rel(dbz, ref_db) < 1e-4 and (dbz - ref_db).abs().max().item() < 1e-3 * (1.0 + ref_db.abs().max().item())
    dw = torch.ones(N, K, device=cuda_device)
    ops.linear_wgrad(dy, x, dw)
    assert rel(dw, dy.float().t() @ x.float() + 1.0) < 1e-3
    db = torch.zeros(N, device=cuda_device)
    ops.colsum(dy, db)
    assert rel(db, dy.float().sum(0)) < 1e-3


@pytest.mark.parametrize("M,N,ld", [(5000, 768, 768), (4100, 2304, 2304), (8192, 3072, 3072), (4097, 264, 264),
                                    (6000, 768, 2304), (300, 768, 768), (20000, 64, 64)])
def test_colsum_shapes(ops, cuda_device, M, N, ld):
    """bias-gradient column sums: tall / wide / narrow shapes, strided rows (a column slice of a wider matrix),
    accumulation onto an existing value."""
    torch.manual_seed(3)
    full = torch.randn(M, ld, device=cuda_device).to(bf16)
    x = full[:, :N]
    out = torch.full((N,), 2.0, device=cuda_device)
    ops.colsum(x, out)
    assert rel(out, x.float().sum(0) + 2.0) < 1e-3


def test_gemm_dropout_epilogue(ops, cuda_device):
    torch.manual_seed(2)
    M, N, K = 512, 768, 256
    x = torch.randn(M, K, device=cuda_device).to(bf16)
    w = torch.randn(N, K, device=cuda_device).to(bf16)
    ref = x.float() @ w.float().t()
    y = ops.linear_fwd(x, w, None, p_drop=0.25, seed=77).float()
    kept = y != 0
    frac = kept.float().mean().item()
    assert abs(frac - 0.75) < 0.01
    assert rel(y[kept], (ref / 0.75)[kept]) < 1e-2
    # the LayerNorm backward regenerates the same mask from (seed, row * N + col)
    g = torch.ones(N, device=cuda_device)
    xin = torch.randn(M, N, device=cuda_device).to(bf16)
    _, mean, rstd = ops.layernorm_fwd(xin, g, torch.zeros_like(g), 1e-5)
    dg, db = torch.zeros_like(g), torch.zeros_like(g)
    dx, dxm = ops.layernorm_bwd(torch.randn(M, N, device=cuda_device).to(bf16), xin, mean, rstd, g, dg, db,
                                p_out=0.25, seed_out=77)
    # the masked gradient is non-zero exactly where the forward kept the element and the gradient itself is non-zero
    # (the former two-way "or" of this assertion was implied by this predicate alone: where dx != 0 both said
    # "dxm != 0 <=> kept", where dx == 0 the masked value is 0 and only this form holds)
    assert ((dxm.float() != 0) == (kept & (dx.float() != 0))).all()


def test_gelu_epilogue_tails(ops, cuda_device):
    """The epilogue's GELU is erf evaluated as tanh(x(A + Bx^2)) with tanh.approx (a deliberate departure from erff,
    |dPhi| < 4e-4): checked where the approximation is weakest relative to the value -- the tails |x| > 4 -- and across
    the whole range, forward and derivative, through an identity weight so that z is exactly the input."""
    torch.manual_seed(14)
    M = N = K = 256
    x = torch.cat([torch.linspace(-9, 9, M * K // 2), (torch.rand(M * K // 2) * 2 - 1) * 9]).view(M, K)
    x = x.to(cuda_device).to(bf16)
    eye = torch.eye(N, K, device=cuda_device).to(bf16)
    zero = torch.zeros(N, device=cuda_device)
    z, a = ops.linear_gelu_fwd(x, eye, zero)
    xf = x.float()
    assert torch.equal(z.float(), xf)
    ref = F.gelu(xf)
    err = (a.float() - ref).abs()
    assert (err <= 4e-4 * xf.abs() + 2.0 ** -8 * ref.abs() + 1e-6).all(), err.max().item()
    tail = xf.abs() > 4
    assert (a.float()[tail & (xf < 0)].abs() < 2e-3).all()                 # gelu(x < -4) -> 0 (|exact| < 1.3e-4)
    assert rel(a.float()[tail & (xf > 0)], ref[tail & (xf > 0)]) < 3e-3      # gelu(x > 4) -> x
    # derivative: dz = (dy W) * gelu'(z) with dy = ones through the identity -> gelu'(z)
    ones = torch.ones(M, N, device=cuda_device).to(bf16)
    dz = ops.linear_dgrad(ones, eye, gelu_z=z)
    xr = xf.clone().requires_grad_(True)
    F.gelu(xr).sum().backward()
    derr = (dz.float() - xr.grad).abs()
    assert (derr <= 3e-3 + 2.0 ** -8 * xr.grad.abs()).all(), derr.max().item()
    assert (dz.float()[tail & (xf < 0)].abs() < 3e-3).all() and rel(dz.float()[tail & (xf > 0)], xr.grad[tail & (xf > 0)]) < 5e-3


# ------------------------------------------------------------------------------------------------- attention
def _attn_ref(qkv, key_bias, B, H, S):
    D = H * 64
    q, k, v = qkv.float().view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) / 8.0
    if key_bias is not None:
        s = s + key_bias.view(B, 1, 1, S)
    p = torch.softmax(s, -1)
    return (p @ v).permute(0, 2, 1, 3).reshape(B * S, D)


@pytest.mark.parametrize("B,H,S", [(2, 2, 128), (3, 12, 128), (2, 4, 64), (2, 3, 100), (2, 2, 256), (2, 3, 197),
                                   (1, 4, 512), (2, 2, 300), (3, 2, 257), (2, 2, 384), (5, 12, 197),
                                   # more (batch, head) items than CTAs: the persistent / cross-head pipelined paths
                                   (40, 12, 128), (30, 12, 197), (26, 12, 64), (20, 16, 257), (2, 2, 300)])
def test_attention_fwd_bwd(ops, cuda_device, B, H, S):
    torch.manual_seed(3)
    D = H * 64
    qkv = torch.randn(B * S, 3 * D, device=cuda_device).to(bf16)
    lengths = torch.randint(4, S + 1, (B,), device=cuda_device)
    lengths[0] = S
    mask = (torch.arange(S, device=cuda_device)[None] < lengths[:, None]).long()
    bias = ops.mask_to_bias(mask)
    out, lse = ops.attention_fwd(qkv, bias, B, H, S)
    qf = qkv.float().requires_grad_(True)
    ref = _attn_ref(qf, bias, B, H, S)
    assert rel(out, ref) < 1e-2
    dout = torch.randn(B * S, D, device=cuda_device).to(bf16)
    ref.backward(dout.float())
    dqkv = ops.attention_bwd(qkv, bias, out, dout, lse, B, H, S)
    assert rel(dqkv, qf.grad) < 2e-2


def test_attention_dropout_is_consistent(ops, cuda_device):
    """With p > 0 the forward is an unbiased estimate of the p = 0 output and the backward uses the same mask:
    checked through the directional derivative d<O, dO>/d eps along a random direction."""
    torch.manual_seed(4)
    B, H, S = 2, 2, 128
    D = H * 64
    qkv = (torch.randn(B * S, 3 * D, device=cuda_device) * 0.5).to(bf16)
    o0, _ = ops.attention_fwd(qkv, None, B, H, S)
    acc = torch.zeros_like(o0, dtype=torch.float32)
    n = 64
    for i in range(n):
        o, _ = ops.attention_fwd(qkv, None, B, H, S, p_drop=0.1, seed=1000 + i)
        acc += o.float()
    assert rel(acc / n, o0) < 0.06
    out, lse = ops.attention_fwd(qkv, None, B, H, S, p_drop=0.1, seed=5)
    out2, _ = ops.attention_fwd(qkv, None, B, H, S, p_drop=0.1, seed=5)
    assert torch.equal(out, out2)
    dout = torch.randn(B * S, D, device=cuda_device).to(bf16)
    dqkv = ops.attention_bwd(qkv, None, out, dout, lse, B, H, S, p_drop=0.1, seed=5).float()
    direction = torch.randn_like(dqkv)
    eps = 0.05
    op, _ = ops.attention_fwd((qkv.float() + eps * direction).to(bf16), None, B, H, S, p_drop=0.1, seed=5)
    om, _ = ops.attention_fwd((qkv.float() - eps * direction).to(bf16), None, B, H, S, p_drop=0.1, seed=5)
    fd = ((op.float() - om.float()) * dout.float()).sum().item() / (2 * eps)
    an = (dqkv * direction).sum().item()
    assert abs(fd - an) / (abs(an) + 1e-6) < 0.1


@pytest.mark.parametrize("B,H,S", [(40, 12, 128), (5, 2, 77), (300, 2, 128)])
def test_attention_saved_dropout_mask_equals_regenerated(ops, cuda_device, B, H, S):
    """The forward's saved keep bits ([B*H, 4, 128] words) are exactly the Philox decisions the backward would
    regenerate: the gradient computed from the saved mask is bit-identical to the regenerated one, the bits'
    density is 1 - p, and padded rows / keys carry no influence."""
    torch.manual_seed(12)
    D = H * 64
    qkv = (torch.randn(B * S, 3 * D, device=cuda_device) * 0.7).to(bf16)
    lengths = torch.randint(3, S + 1, (B,), device=cuda_device)
    lengths[0] = S
    bias = ops.mask_to_bias((torch.arange(S, device=cuda_device)[None] < lengths[:, None]).long())
    dout = torch.randn(B * S, D, device=cuda_device).to(bf16)
    out, lse, mask = ops.attention_fwd(qkv, bias, B, H, S, p_drop=0.1, seed=77, save_mask=True)
    out_plain, lse_plain = ops.attention_fwd(qkv, bias, B, H, S, p_drop=0.1, seed=77)
    assert mask is not None and mask.shape == (B * H, 4, 128)
    assert torch.equal(out, out_plain) and torch.equal(lse, lse_plain)
    g_saved = ops.attention_bwd(qkv, bias, out, dout, lse, B, H, S, p_drop=0.1, seed=77, drop_mask=mask)
    g_regen = ops.attention_bwd(qkv, bias, out, dout, lse, B, H, S, p_drop=0.1, seed=77)
    assert torch.equal(g_saved, g_regen)
    bits = torch.stack([(mask >> i) & 1 for i in range(32)], -1).float()        # [BH, 4, 128 rows, 32]
    assert abs(bits.mean().item() - 0.9) < 5e-3
    # no mask is produced (or needed) without dropout or beyond one tile
    assert ops.attention_fwd(qkv, bias, B, H, S, p_drop=0.0, save_mask=True)[2] is None


@pytest.mark.parametrize("B,H,S,p", [(3, 4, 128, 0.0), (40, 12, 128, 0.1), (5, 2, 77, 0.1), (700, 2, 128, 0.1)])
def test_attention_warp_specialised_matches_single_role(ops, cuda_device, B, H, S, p):
    """attention_ws.cu (TMA warp / MMA warp / softmax warpgroups, several heads in flight) against the single-role
    kernels it replaces for one-tile sequences: same Philox stream and element indexing, so with the SAME seed the two
    families produce the same dropout mask -- outputs, LSE and all three gradients agree to bf16 rounding, element by
    element (max-abs, not only the L2 ratio: a wrong edge row or a bad last head would hide in a norm).  The last
    shape gives every CTA 9-10 heads, i.e. several trips around the 4-slot / 2-slot rings."""
    import os
    torch.manual_seed(11)
    D = H * 64
    qkv = (torch.randn(B * S, 3 * D, device=cuda_device) * 0.7).to(bf16)
    lengths = torch.randint(3, S + 1, (B,), device=cuda_device)
    lengths[0] = S
    bias = ops.mask_to_bias((torch.arange(S, device=cuda_device)[None] < lengths[:, None]).long())
    dout = torch.randn(B * S, D, device=cuda_device).to(bf16)
    res = {}
    try:
        for ws in ("0", "1"):
            os.environ["B200MM_ATTN_WS"] = ws
            out, lse = ops.attention_fwd(qkv, bias, B, H, S, p_drop=p, seed=99)
            dqkv = ops.attention_bwd(qkv, bias, out, dout, lse, B, H, S, p_drop=p, seed=99)
            res[ws] = (out.float(), lse.clone(), dqkv.float())
        # mixed: warp-specialised backward on the single-role forward's output / LSE
        os.environ["B200MM_ATTN_WS"] = "1"
        dq_mixed = ops.attention_bwd(qkv, bias, res["0"][0].to(bf16), dout, res["0"][1], B, H, S, p_drop=p, seed=99).float()
    finally:
        os.environ["B200MM_ATTN_WS"] = "1"
    o0, l0, g0 = res["0"]
    o1, l1, g1 = res["1"]
    assert torch.isfinite(o1).all() and torch.isfinite(g1).all()
    assert rel(o1, o0) < 4e-3 and (o1 - o0).abs().max().item() < 2e-2 * o0.abs().max().item() + 1e-3
    assert (l1 - l0).abs().max().item() < 1e-4
    for name, a, b in (("ws", g1, g0), ("mixed", dq_mixed, g0)):
        assert rel(a, b) < 1e-2, name
        assert (a - b).abs().max().item() < 3e-2 * b.abs().max().item() + 1e-3, name
        # per-row check of the last query row of the last head and of the rows past a short sequence's end
        last = slice((B - 1) * S, B * S)
        assert rel(a[last, -64:], b[last, -64:]) < 2e-2, name
    if p == 0.0:
        qf = qkv.float().requires_grad_(True)
        ref = _attn_ref(qf, bias, B, H, S)
        assert rel(o1, ref) < 1e-2
        ref.backward(dout.float())
        assert rel(g1, qf.grad) < 2e-2


# ------------------------------------------------------------------------------------------------- LayerNorm / embeddings
@pytest.mark.parametrize("M,D", [(64, 128), (1000, 768), (300, 1024), (96, 2048)])
def test_layernorm_fwd_bwd(ops, cuda_device, M, D):
    torch.manual_seed(5)
    x = torch.randn(M, D, device=cuda_device).to(bf16)
    g = torch.randn(D, device=cuda_device)
    b = torch.randn(D, device=cuda_device)
    y, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-12)
    xf = x.float().requires_grad_(True)
    gf = g.clone().requires_grad_(True)
    bf = b.clone().requires_grad_(True)
    ref = F.layer_norm(xf, (D,), gf, bf, 1e-12)
    assert rel(y, ref) < 1e-2
    dy = torch.randn(M, D, device=cuda_device).to(bf16)
    ref.backward(dy.float())
    dg = torch.zeros(D, device=cuda_device)
    db = torch.zeros(D, device=cuda_device)
    dx, dx2 = ops.layernorm_bwd(dy, x, mean, rstd, g, dg, db)
    assert dx2 is None
    assert rel(dx, xf.grad) < 1e-2
    assert rel(dg, gf.grad) < 1e-2 and rel(db, bf.grad) < 1e-2


def test_embedding_fwd_bwd(ops, cuda_device):
    torch.manual_seed(6)
    B, S, V, D = 4, 32, 500, 128
    ids = torch.randint(0, V, (B, S), device=cuda_device)
    word = torch.randn(V, D, device=cuda_device)
    pos = torch.randn(64, D, device=cuda_device)
    g = torch.rand(D, device=cuda_device) + 0.5
    b = torch.randn(D, device=cuda_device)
    y, xs, mean, rstd = ops.embed_layernorm_fwd(ids, word, pos, g, b, 1e-12)
    ref_x = word[ids] + pos[:S][None]
    assert rel(xs, ref_x.view(B * S, D)) < 1e-2
    assert rel(y, F.layer_norm(ref_x, (D,), g, b, 1e-12).view(B * S, D)) < 1e-2
    dx = torch.randn(B * S, D, device=cuda_device).to(bf16)
    dword = torch.zeros(V, D, device=cuda_device)
    dpos = torch.zeros(64, D, device=cuda_device)
    ops.embedding_bwd(dx, ids, dword, dpos)
    ref_w = torch.zeros(V, D, device=cuda_device).index_add_(0, ids.view(-1), dx.float())
    pad = int(ids[0, 0])
    dword_p = torch.zeros(V, D, device=cuda_device)
    ops.embedding_bwd(dx, ids, dword_p, torch.zeros(64, D, device=cuda_device), padding_idx=pad)
    ref_wp = ref_w.clone()
    ref_wp[pad] = 0
    assert rel(dword_p, ref_wp) < 1e-5 and dword_p[pad].abs().sum().item() == 0
    ref_p = torch.zeros(64, D, device=cuda_device)
    ref_p[:S] = dx.float().view(B, S, D).sum(0)
    assert rel(dword, ref_w) < 1e-5 and rel(dpos, ref_p) < 1e-5


def test_gather_scatter_rows(ops, cuda_device):
    torch.manual_seed(7)
    B, S, D = 5, 16, 64
    h = torch.randn(B * S, D, device=cuda_device).to(bf16)
    last = ops.gather_rows(h, B, S, S - 1)
    assert torch.equal(last, h.view(B, S, D)[:, -1])
    cls = ops.gather_rows(h, B, S, 0)
    assert torch.equal(cls, h.view(B, S, D)[:, 0])
    d = torch.randn(B, D, device=cuda_device).to(bf16)
    dx = ops.scatter_rows(d, B * S, S, S - 1).view(B, S, D)
    assert torch.equal(dx[:, -1], d) and dx[:, :-1].abs().sum().item() == 0
    dropped = ops.gather_rows(h, B, S, S - 1, p_drop=0.3, seed=9).float()
    back = ops.scatter_rows(torch.ones(B, D, device=cuda_device).to(bf16), B * S, S, S - 1, p_drop=0.3, seed=9)
    assert torch.equal(dropped != 0, (back.view(B, S, D)[:, -1].float() != 0) & (last.float() != 0))


# ------------------------------------------------------------------------------------------------- BatchNorm / pools / conv lowering
@pytest.mark.parametrize("M,C,relu,use_res", [(4096, 64, True, False), (1568, 256, True, True), (392, 2048, False, False)])
def test_batchnorm_fwd_bwd(ops, cuda_device, M, C, relu, use_res):
    torch.manual_seed(8)
    x = (torch.randn(M, C, device=cuda_device) * 2 + 0.5).to(bf16)
    res = torch.randn(M, C, device=cuda_device).to(bf16) if use_res else None
    g = torch.rand(C, device=cuda_device) + 0.5
    b = torch.randn(C, device=cuda_device)
    rm = torch.zeros(C, device=cuda_device)
    rv = torch.ones(C, device=cuda_device)
    out, mean, rstd = ops.batchnorm_fwd(x, g, b, rm, rv, residual=res, relu=relu)
    xf = x.float().requires_grad_(True)
    gf = g.clone().requires_grad_(True)
    bf = b.clone().requires_grad_(True)
    rm2, rv2 = torch.zeros(C, device=cuda_device), torch.ones(C, device=cuda_device)
    ref = F.batch_norm(xf, rm2, rv2, gf, bf, True, 0.1, 1e-5)
    resf = res.float().requires_grad_(True) if use_res else None
    if use_res:
        ref = ref + resf
    if relu:
        ref = torch.relu(ref)
    assert rel(out, ref) < 1e-2
    assert rel(rm, rm2) < 1e-3 and rel(rv, rv2) < 1e-3
    dout = torch.randn(M, C, device=cuda_device).to(bf16)
    ref.backward(dout.float())
    dg, db = torch.zeros(C, device=cuda_device), torch.zeros(C, device=cuda_device)
    dx, dz = ops.batchnorm_bwd(dout, out, x, mean, rstd, g, dg, db, relu=relu, need_dz=use_res)
    assert rel(dx, xf.grad) < 2e-2
    assert rel(dg, gf.grad) < 2e-2 and rel(db, bf.grad) < 2e-2
    if use_res:
        assert rel(dz, resf.grad) < 2e-2
    if relu and not use_res:
        # ReLU mask recomputed from x instead of read from the saved output: same gradients up to the summation
        # order of the column reductions (the two variants unroll differently)
        dg2, db2 = torch.zeros(C, device=cuda_device), torch.zeros(C, device=cuda_device)
        dx2, _ = ops.batchnorm_bwd(dout, None, x, mean, rstd, g, dg2, db2, relu=True, beta=b)
        assert rel(dg2, dg) < 1e-5 and rel(db2, db) < 1e-5 and rel(dx2, dx) < 1e-3
    if relu:
        # 1-bit ReLU mask written by the statistics-fed forward, read by the backward instead of the output tensor
        st = torch.zeros(2 * C, device=cuda_device)
        st[:C], st[C:] = x.float().sum(0), (x.float() ** 2).sum(0)
        rm3, rv3 = torch.zeros(C, device=cuda_device), torch.ones(C, device=cuda_device)
        out3, mean3, rstd3, msk = ops.batchnorm_fwd(x, g, b, rm3, rv3, residual=res, relu=True, col_stats=st,
                                                    want_mask=True)
        bits = ((msk.unsqueeze(-1) >> torch.arange(8, device=cuda_device, dtype=torch.uint8)) & 1).view(M, C).bool()
        assert torch.equal(bits, out3 > 0)
        dg3, db3 = torch.zeros(C, device=cuda_device), torch.zeros(C, device=cuda_device)
        dx3, dz3 = ops.batchnorm_bwd(dout, None, x, mean3, rstd3, g, dg3, db3, relu=True, need_dz=use_res, mask=msk)
        assert rel(dx3, xf.grad) < 2e-2 and rel(dg3, gf.grad) < 2e-2 and rel(db3, bf.grad) < 2e-2
        if use_res:
            assert rel(dz3, resf.grad) < 2e-2
    ev = ops.batchnorm_eval(x, g, b, rm2, rv2, residual=res, relu=relu)
    ref_ev = F.batch_norm(x.float(), rm2, rv2, g, b, False, 0.1, 1e-5)
    if use_res:
        ref_ev = ref_ev + res.float()
    if relu:
        ref_ev = torch.relu(ref_ev)
    assert rel(ev, ref_ev) < 1e-2


@pytest.mark.parametrize("M,C,mode", [(50176, 256, "recompute"), (12544, 512, "mask"), (3001, 64, "recompute"),
                                      (777, 2048, "mask"), (20000, 128, "plain")])
def test_batchnorm_bwd_single_launch_matches_two_pass(ops, cuda_device, M, C, mode):
    """BatchNorm backward as ONE cooperative launch (reduce -> grid barrier -> apply, L2-resident tensors) against the
    two-kernel path and against autograd: same dx / dz element for element up to the summation order of the column
    sums, last rows included."""
    torch.manual_seed(17)
    x = (torch.randn(M, C, device=cuda_device) * 1.5 + 0.3).to(bf16)
    g = torch.rand(C, device=cuda_device) + 0.5
    b = torch.randn(C, device=cuda_device) * 0.3
    res = torch.randn(M, C, device=cuda_device).to(bf16) if mode == "mask" else None
    st = torch.zeros(2 * C, device=cuda_device)
    st[:C], st[C:] = x.float().sum(0), (x.float() ** 2).sum(0)
    rm, rv = torch.zeros(C, device=cuda_device), torch.ones(C, device=cuda_device)
    relu = mode != "plain"
    out, mean, rstd, msk = ops.batchnorm_fwd(x, g, b, rm, rv, residual=res, relu=relu, col_stats=st, want_mask=True)
    dout = torch.randn(M, C, device=cuda_device).to(bf16)
    kw = dict(relu=relu, need_dz=mode == "mask", mask=msk if mode == "mask" else None,
              beta=b if mode == "recompute" else None)
    got = {}
    try:
        for name, mb in (("fused", 52), ("two_pass", 0)):
            ops.set_bn_fused_mb(mb)
            dg, db = torch.zeros(C, device=cuda_device), torch.zeros(C, device=cuda_device)
            dx, dz = ops.batchnorm_bwd(dout, None, x, mean, rstd, g, dg, db, **kw)
            got[name] = (dx, dz, dg, db)
    finally:
        ops.set_bn_fused_mb(0)
    (dx, dz, dg, db), (dx2, dz2, dg2, db2) = got["fused"], got["two_pass"]
    assert rel(dg, dg2) < 1e-5 and rel(db, db2) < 1e-5
    assert rel(dx, dx2) < 1e-3 and (dx.float() - dx2.float()).abs().max().item() < 0.05
    assert (dx[-1].float() - dx2[-1].float()).abs().max().item() < 0.05
    if dz is not None:
        assert torch.equal(dz, dz2)
    # autograd reference
    xf = x.float().requires_grad_(True)
    gf, bff = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = F.batch_norm(xf, None, None, gf, bff, True, 0.1, 1e-5)
    if res is not None:
        y = y + res.float()
    if relu:
        y = torch.relu(y)
    y.backward(dout.float())
    assert rel(dx, xf.grad) < 2e-2 and rel(dg, gf.grad) < 2e-2 and rel(db, bff.grad) < 2e-2


@pytest.mark.parametrize("M,N,K", [(4096, 64, 152), (3000, 256, 64), (1000, 512, 256), (777, 2048, 512), (5000, 128, 1152)])
def test_gemm_column_statistics_feed_batchnorm(ops, cuda_device, M, N, K):
    """The convolution epilogue's column sums (over the stored bf16 outputs) and the one-pass BatchNorm they feed,
    against the separate statistics pass and against F.batch_norm."""
    torch.manual_seed(21)
    x = torch.randn(M, K, device=cuda_device).to(bf16)
    w = (torch.randn(N, K, device=cuda_device) / K ** 0.5).to(bf16)
    stats = torch.zeros(2 * N, device=cuda_device)
    y = ops.linear_fwd(x, w, col_stats=stats)
    y_plain = ops.linear_fwd(x, w)
    assert torch.equal(y, y_plain)
    yf = y.float()
    assert rel(stats[:N], yf.sum(0)) < 1e-4 and rel(stats[N:], (yf * yf).sum(0)) < 1e-4
    g = torch.rand(N, device=cuda_device) + 0.5
    b = torch.randn(N, device=cuda_device)
    rm, rv = torch.zeros(N, device=cuda_device), torch.ones(N, device=cuda_device)
    rm2, rv2 = torch.zeros(N, device=cuda_device), torch.ones(N, device=cuda_device)
    out, mean, rstd = ops.batchnorm_fwd(y, g, b, rm, rv, col_stats=stats)
    out2, mean2, rstd2 = ops.batchnorm_fwd(y, g, b, rm2, rv2)
    assert rel(mean, mean2) < 1e-4 and rel(rstd, rstd2) < 1e-4 and rel(out, out2) < 5e-3
    assert rel(rm, rm2) < 1e-4 and rel(rv, rv2) < 1e-4
    ref = torch.relu(F.batch_norm(yf, None, None, g, b, True, 0.1, 1e-5))
    assert rel(out, ref) < 1e-2


@pytest.mark.parametrize("M,N,K", [(3000, 256, 64), (1000, 512, 128), (20000, 1024, 256), (777, 2048, 512), (300, 64, 256)])
def test_gemm_masked_residual_joins_identity_gradient(ops, cuda_device, M, N, K):
    """dx = dy W + (ReLU-mask bit ? d_out : 0): the identity-branch gradient of a residual block is added in the data
    gradient's epilogue from (d_out, 1-bit mask) -- exact element selection, checked on every element incl. the last
    rows / columns, and against the two-step path (BatchNorm backward's dz, then the in-place accumulate)."""
    torch.manual_seed(31)
    dy = torch.randn(M, K, device=cuda_device).to(bf16)              # gradient w.r.t. the first convolution's output
    w = (torch.randn(K, N, device=cuda_device) / K ** 0.5).to(bf16)   # its weight [Cout = K, Cin = N]
    d_out = torch.randn(M, N, device=cuda_device).to(bf16)
    keep = torch.rand(M, N, device=cuda_device) < 0.6
    bits = keep.view(M, N // 8, 8).to(torch.uint8)
    mask = (bits << torch.arange(8, device=cuda_device, dtype=torch.uint8)).sum(-1).to(torch.uint8).contiguous()
    got = ops.linear_dgrad(dy, w, residual=d_out, residual_mask=mask)
    ref = dy.float() @ w.float() + torch.where(keep, d_out.float(), torch.zeros((), device=cuda_device))
    assert rel(got, ref) < 4e-3
    err = (got.float() - ref).abs()
    assert err.max().item() < 0.06 and err[-1].max().item() < 0.06 and err[:, -1].max().item() < 0.06
    # where dy W is exactly representable (dy = 0) the selection itself is exact
    got0 = ops.linear_dgrad(torch.zeros_like(dy), w, residual=d_out, residual_mask=mask)
    assert torch.equal(got0, torch.where(keep, d_out, torch.zeros((), device=cuda_device, dtype=bf16)))
    # the path it replaces: dz materialised, then accumulated into in place
    dz = torch.where(keep, d_out, torch.zeros((), device=cuda_device, dtype=bf16)).contiguous()
    two_step = ops.linear_dgrad(dy, w, residual=dz, out=dz)
    assert rel(got, two_step) < 6e-3      # (the in-place path rounds acc to bf16 before the add in L2)


def _nhwc(x):  # NCHW fp32 -> [N*H*W, C] bf16
    N, C, H, W = x.shape
    return x.permute(0, 2, 3, 1).reshape(N * H * W, C).to(bf16).contiguous()   # (N == 1: the reshape is a strided view)


def _nchw(x, N, H, W):
    return x.float().view(N, H, W, -1).permute(0, 3, 1, 2)


@pytest.mark.parametrize("H,W", [(18, 18), (17, 21), (112, 112)])
def test_pools(ops, cuda_device, H, W):
    torch.manual_seed(9)
    N, C = 3, 64
    x = torch.randn(N, C, H, W, device=cuda_device).to(bf16).float()
    out, arg, Ho, Wo = ops.maxpool_fwd(_nhwc(x), N, H, W, C)
    xf = x.clone().requires_grad_(True)
    ref = F.max_pool2d(xf, 3, 2, 1)
    assert (Ho, Wo) == tuple(ref.shape[2:])
    assert torch.equal(_nchw(out, N, Ho, Wo), ref.detach())
    dout = torch.randn_like(ref).to(bf16).float()
    ref.backward(dout)
    dx = ops.maxpool_bwd(_nhwc(dout), arg, N, H, W, C)
    assert rel(_nchw(dx, N, H, W), xf.grad) < 1e-2
    a = ops.avgpool_fwd(_nhwc(x), N, H * W, C)
    assert rel(a, x.mean((2, 3))) < 1e-2
    d = torch.randn(N, C, device=cuda_device).to(bf16)
    da = ops.avgpool_bwd(d, N, H * W, C)
    assert rel(_nchw(da, N, H, W), (d.float() / (H * W))[:, :, None, None].expand(N, C, H, W)) < 1e-2


def test_conv_weight_rotate_multi_matches_single(ops, cuda_device):
    """All rotated (transposed-convolution) weights of a backward pass from one launch == one launch per weight."""
    torch.manual_seed(5)
    shapes = [(64, 64), (128, 128), (256, 192), (512, 512)]
    ws = [torch.randn(co, 9 * ci, device=cuda_device).to(bf16) for co, ci in shapes]
    outs = [torch.empty(ci, 9 * co, device=cuda_device, dtype=bf16) for co, ci in shapes]
    table = torch.tensor([[w.data_ptr(), o.data_ptr(), (co << 32) | ci, 9] for w, o, (co, ci) in zip(ws, outs, shapes)],
                         dtype=torch.int64).to(cuda_device)
    ops.conv_weight_rotate_multi(table, len(shapes))
    for w, o, (co, ci) in zip(ws, outs, shapes):
        assert torch.equal(o, ops.conv_weight_rotate(w, co, ci, 3))
        ref = w.view(co, 3, 3, ci).flip(1, 2).permute(3, 1, 2, 0).reshape(ci, 9 * co)
        assert torch.equal(o, ref)


@pytest.mark.parametrize("N,H,W", [(2, 112, 112), (3, 17, 21), (1, 8, 8)])
def test_stem_tail_bn_relu_maxpool_one_pass(ops, cuda_device, N, H, W):
    """bn1 + relu + maxpool in one pass from the convolution output: bit-identical pooled values, argmax, statistics and
    running-statistics update to the BatchNorm kernel followed by the pooling kernel."""
    torch.manual_seed(23)
    C = 64
    x = (torch.randn(N * H * W, C, device=cuda_device) * 2.0 + 0.2).to(bf16)
    g = torch.rand(C, device=cuda_device) + 0.5
    b = torch.randn(C, device=cuda_device) * 0.5
    st = torch.zeros(2 * C, device=cuda_device)
    st[:C], st[C:] = x.float().sum(0), (x.float() ** 2).sum(0)
    rm, rv = torch.zeros(C, device=cuda_device), torch.ones(C, device=cuda_device)
    rm2, rv2 = torch.zeros(C, device=cuda_device), torch.ones(C, device=cuda_device)
    a, mean, rstd = ops.batchnorm_fwd(x, g, b, rm, rv, relu=True, col_stats=st)
    ref, ref_arg, Ho, Wo = ops.maxpool_fwd(a, N, H, W, C)
    got, arg, Ho2, Wo2, mean2, rstd2 = ops.bn_relu_maxpool_fwd(x, N, H, W, C, g, b, rm2, rv2, st)
    assert (Ho, Wo) == (Ho2, Wo2)
    assert torch.equal(got, ref) and torch.equal(arg, ref_arg)
    assert torch.equal(mean, mean2) and torch.equal(rstd, rstd2) and torch.equal(rm, rm2) and torch.equal(rv, rv2)


@pytest.mark.parametrize("k,stride,pad,Cin,Cout,H", [(3, 1, 1, 64, 64, 14), (3, 2, 1, 128, 128, 16), (1, 1, 0, 64, 256, 8)])
def test_conv_as_gemm(ops, cuda_device, k, stride, pad, Cin, Cout, H):
    torch.manual_seed(10)
    N, W = 4, H
    x = torch.randn(N, Cin, H, W, device=cuda_device).to(bf16).float()
    w = (torch.randn(Cout, Cin, k, k, device=cuda_device) * 0.05).to(bf16).float()
    xf = x.clone().requires_grad_(True)
    wf = w.clone().requires_grad_(True)
    ref = F.conv2d(xf, wf, None, stride, pad)
    w_ohwi = w.permute(0, 2, 3, 1).reshape(Cout, k * k * Cin).to(bf16).contiguous()
    xn = _nhwc(x)
    if k == 1:
        cols, Ho, Wo = xn, H, W
    else:
        cols, Ho, Wo = ops.im2col(xn, N, H, W, Cin, k, stride, pad)
    y = ops.linear_fwd(cols, w_ohwi)
    assert rel(_nchw(y, N, Ho, Wo), ref) < 1e-2
    dy = torch.randn_like(ref).to(bf16).float()
    ref.backward(dy)
    dyn = _nhwc(dy)
    dw = torch.zeros(Cout, k * k * Cin, device=cuda_device)
    ops.linear_wgrad(dyn, cols, dw)
    assert rel(dw.view(Cout, k, k, Cin).permute(0, 3, 1, 2), wf.grad) < 1e-2
    dcols = ops.linear_dgrad(dyn, w_ohwi)
    dx = dcols if k == 1 else ops.col2im(dcols, N, H, W, Cin, k, stride, pad)
    assert rel(_nchw(dx, N, H, W), xf.grad) < 2e-2


def test_stem_lowering_and_subsample(ops, cuda_device):
    torch.manual_seed(11)
    N, H, W = 2, 32, 32
    img = torch.randn(N, 3, H, W, device=cuda_device)
    w = (torch.randn(64, 3, 7, 7, device=cuda_device) * 0.05).to(bf16).float()
    ref = F.conv2d(img.to(bf16).float(), w, None, 2, 3)
    Kp = 152
    cols, Ho, Wo = ops.im2col_nchw_f32(img, 7, 2, 3, Kp)
    wp = torch.zeros(64, Kp, device=cuda_device, dtype=bf16)
    wp[:, :147] = w.permute(0, 2, 3, 1).reshape(64, 147).to(bf16)
    y = ops.linear_fwd(cols, wp)
    assert rel(_nchw(y, N, Ho, Wo), ref) < 1e-2
    x = torch.randn(N, 64, 8, 8, device=cuda_device).to(bf16).float()
    sub, ho, wo = ops.subsample(_nhwc(x), N, 8, 8, 64, 2)
    assert torch.equal(_nchw(sub, N, ho, wo), x[:, :, ::2, ::2])
    add = torch.randn(N, 64, 8, 8, device=cuda_device).to(bf16).float()
    up = ops.upsample_add(sub, _nhwc(add), N, 8, 8, 64, 2)
    ref_up = add.clone()
    ref_up[:, :, ::2, ::2] += x[:, :, ::2, ::2]
    assert rel(_nchw(up, N, 8, 8), ref_up) < 1e-2


@pytest.mark.parametrize("N,H,W", [(2, 8, 8), (3, 10, 12), (2, 7, 30), (2, 56, 56), (12, 56, 56), (3, 33, 62), (2, 9, 61),
                                   (1, 3, 6)])
def test_conv3x3_c64_halo_resident(ops, cuda_device, N, H, W):
    """csrc/conv3x3_c64.cu (taken by conv_fwd / conv_wgrad for 3x3 / 1 / 1, 64 -> 64): forward + BN statistics, data
    gradient through the rotated weights, weight gradient -- against autograd on the same bf16-rounded operands.
    (12, 56, 56) is 336 row groups > 148 CTAs (several tiles per CTA, ring and TMEM double buffer wrap); (3, 10, 12),
    (2, 7, 30) and (3, 33, 62) have a partial last row group / a row count R that does not divide H."""
    torch.manual_seed(31)
    C = 64
    x = torch.randn(N, C, H, W, device=cuda_device).to(bf16).float()
    w = (torch.randn(C, C, 3, 3, device=cuda_device) * 0.05).to(bf16).float()
    xf, wf = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    ref = F.conv2d(xf, wf, None, 1, 1)
    w_ohwi = w.permute(0, 2, 3, 1).reshape(C, 9 * C).to(bf16).contiguous()
    xn = _nhwc(x)
    stats = torch.zeros(2 * C, device=cuda_device)
    y, P, Q = ops.conv_fwd(xn, N, H, W, C, w_ohwi, 3, 1, 1, col_stats=stats)
    assert (P, Q) == (H, W)
    assert rel(_nchw(y, N, H, W), ref) < 1e-2
    yf = y.float()
    assert rel(stats[:C], yf.sum(0)) < 2e-3 and rel(stats[C:], (yf * yf).sum(0)) < 2e-3
    dy = torch.randn_like(ref).to(bf16).float()
    ref.backward(dy)
    dyn = _nhwc(dy)
    dw = torch.full((C, 9 * C), 0.25, device=cuda_device)
    ops.conv_wgrad(dyn, xn, N, H, W, C, 3, 1, 1, dw)
    assert rel((dw - 0.25).view(C, 3, 3, C).permute(0, 3, 1, 2), wf.grad) < 1e-2
    w_rot = ops.conv_weight_rotate(w_ohwi, C, C, 3)
    dx, _, _ = ops.conv_fwd(dyn, N, H, W, C, w_rot, 3, 1, 1)
    assert rel(_nchw(dx, N, H, W), xf.grad) < 1e-2


@pytest.mark.parametrize("N,H,W", [(2, 32, 32), (3, 64, 48), (2, 224, 224), (40, 224, 224), (1, 16, 226), (2, 15, 17)])
def test_stem_direct_conv(ops, cuda_device, N, H, W):
    """csrc/stem_conv.cu against F.conv2d (same bf16-rounded operands) and against the im2col lowering it replaces;
    N = 40 at 224 x 224 gives 4480 output rows > 2 x 148 CTAs, so every CTA loops over several rows."""
    torch.manual_seed(21)
    img = torch.randn(N, 3, H, W, device=cuda_device)
    w = (torch.randn(64, 3, 7, 7, device=cuda_device) * 0.05).to(bf16).float()
    wp = torch.zeros(64, 152, device=cuda_device, dtype=bf16)
    wp[:, :147] = w.permute(0, 2, 3, 1).reshape(64, 147).to(bf16)
    assert ops.stem_conv_supported(img, wp)
    ref = F.conv2d(img.to(bf16).float(), w, None, 2, 3)
    stats = torch.zeros(128, device=cuda_device)
    y, Ho, Wo = ops.stem_conv_fwd(img, wp, col_stats=stats)
    assert (Ho, Wo) == tuple(ref.shape[2:])
    assert rel(_nchw(y, N, Ho, Wo), ref) < 1e-2
    cols, _, _ = ops.im2col_nchw_f32(img, 7, 2, 3, 152)
    y_low = ops.linear_fwd(cols, wp)
    assert rel(y, y_low) < 5e-3
    yf = y.float()
    assert rel(stats[:64], yf.sum(0)) < 1e-3 and rel(stats[64:], (yf * yf).sum(0)) < 1e-3
    # weight gradient, accumulated on top of an existing value
    dy = torch.randn(N * Ho * Wo, 64, device=cuda_device).to(bf16)
    dw = torch.full((64, 152), 0.5, device=cuda_device)
    ops.stem_conv_wgrad(img, dy, dw)
    ref_dw = dy.float().t() @ cols.float() + 0.5
    assert rel(dw, ref_dw) < 1e-3
    assert torch.equal(dw[:, 147:], torch.full((64, 5), 0.5, device=cuda_device))


# ------------------------------------------------------------------------------------------------- head / loss / optimizer
def test_head_cross_entropy(ops, cuda_device):
    torch.manual_seed(12)
    B, Fd, C = 37, 512, 2
    feat = torch.randn(B, Fd, device=cuda_device).to(bf16)
    W = torch.randn(C, Fd, device=cuda_device) * 0.05
    b = torch.randn(C, device=cuda_device)
    labels = torch.randint(0, C, (B,), device=cuda_device)
    dW, db = torch.zeros_like(W), torch.zeros_like(b)
    logits, loss, correct, dfeat = ops.head_loss(feat, W, b, labels, dW=dW, dbias=db)
    ff = feat.float().requires_grad_(True)
    Wf = W.clone().requires_grad_(True)
    bf = b.clone().requires_grad_(True)
    ref_logits = ff @ Wf.t() + bf
    ref_loss = F.cross_entropy(ref_logits, labels)
    ref_loss.backward()
    assert rel(logits, ref_logits) < 1e-5
    assert abs(loss.item() - ref_loss.item()) < 1e-5
    assert correct.item() == (ref_logits.argmax(1) == labels).sum().item()
    assert rel(dfeat, ff.grad) < 1e-2 and rel(dW, Wf.grad) < 1e-4 and rel(db, bf.grad) < 1e-4
    # external-dlogits path == what autograd of an arbitrary criterion would hand back
    dlog = torch.autograd.grad(F.cross_entropy(ref_logits.detach().requires_grad_(True), labels),
                               [], allow_unused=True) if False else None
    lg = ref_logits.detach().requires_grad_(True)
    F.cross_entropy(lg, labels).backward()
    dW2, db2 = torch.zeros_like(W), torch.zeros_like(b)
    _, _, _, dfeat2 = ops.head_loss(feat, W, b, None, loss_kind=ops.LOSS_EXTERNAL, dW=dW2, dbias=db2,
                                    dlogits=lg.grad.contiguous())
    assert rel(dfeat2, ff.grad) < 1e-2 and rel(dW2, Wf.grad) < 1e-4


def test_head_focal(ops, cuda_device):
    from torchvision.ops import sigmoid_focal_loss
    torch.manual_seed(13)
    B, Fd = 64, 512
    feat = torch.randn(B, Fd, device=cuda_device).to(bf16)
    W = torch.randn(1, Fd, device=cuda_device) * 0.1
    b = torch.randn(1, device=cuda_device)
    labels = torch.randint(0, 2, (B,), device=cuda_device)
    dW, db = torch.zeros_like(W), torch.zeros_like(b)
    logits, loss, correct, dfeat = ops.head_loss(feat, W, b, labels, loss_kind=ops.LOSS_FOCAL, alpha=0.25, gamma=2.0,
                                                 dW=dW, dbias=db)
    ff = feat.float().requires_grad_(True)
    Wf = W.clone().requires_grad_(True)
    ref_logits = (ff @ Wf.t() + b).squeeze(1)
    ref = sigmoid_focal_loss(ref_logits, labels.float(), alpha=0.25, gamma=2.0, reduction="mean")
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-5
    assert rel(dfeat, ff.grad) < 1e-2 and rel(dW, Wf.grad) < 1e-4


def test_adam_matches_torch(ops, cuda_device):
    torch.manual_seed(14)
    n = 4096 * 33
    p = torch.randn(n, device=cuda_device)
    ref_p = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref_p], lr=2e-5)
    m, v = torch.zeros(n, device=cuda_device), torch.zeros(n, device=cuda_device)
    shadow = torch.empty(n, device=cuda_device, dtype=bf16)
    for step in range(1, 6):
        g = torch.randn(n, device=cuda_device) * (0.1 if step % 2 else 10)
        ref_p.grad = g.clone()
        opt.step()
        ops.adam_step(p, g, m, v, shadow, lr=2e-5, step=step)
    assert (p - ref_p.detach()).abs().max().item() < 1e-6
    assert torch.equal(shadow, p.to(bf16))
    # clipping: same as clip_grad_norm_ followed by Adam
    g = torch.randn(n, device=cuda_device) * 3
    ref_p.grad = g.clone()
    torch.nn.utils.clip_grad_norm_([ref_p], 1.0)
    opt.step()
    sq = torch.zeros(1, device=cuda_device)
    ops.sumsq(g, sq)
    assert abs(sq.item() - (g.double() ** 2).sum().item()) / sq.item() < 1e-4
    ops.adam_step(p, g, m, v, shadow, lr=2e-5, step=6, gradsq=sq, max_norm=1.0)
    assert (p - ref_p.detach()).abs().max().item() < 1e-6


# ------------------------------------------------------------------------------------------------- implicit-GEMM convolution
@pytest.mark.parametrize("N,H,Cin,Cout,stride", [(4, 14, 64, 64, 1), (2, 16, 128, 128, 2), (8, 28, 128, 256, 1),
                                                   (3, 7, 512, 512, 1), (5, 56, 64, 64, 1)])
def test_implicit_gemm_conv3x3(ops, cuda_device, N, H, Cin, Cout, stride):
    """TMA-im2col convolution (no lowering matrix in memory): forward, stride-1 data gradient through the rotated
    weight, and the weight gradient, all vs F.conv2d / autograd."""
    torch.manual_seed(20)
    W = H
    x = torch.randn(N, Cin, H, W, device=cuda_device).to(bf16).float()
    w = (torch.randn(Cout, Cin, 3, 3, device=cuda_device) * 0.05).to(bf16).float()
    xf = x.clone().requires_grad_(True)
    wf = w.clone().requires_grad_(True)
    ref = F.conv2d(xf, wf, None, stride, 1)
    w_ohwi = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).to(bf16).contiguous()
    xn = _nhwc(x)
    y, P, Q = ops.conv_fwd(xn, N, H, W, Cin, w_ohwi, 3, stride, 1)
    assert (P, Q) == tuple(ref.shape[2:])
    assert rel(_nchw(y, N, P, Q), ref) < 1e-2
    dy = torch.randn_like(ref).to(bf16).float()
    ref.backward(dy)
    dyn = _nhwc(dy)
    dw = torch.zeros(Cout, 9 * Cin, device=cuda_device)
    ops.conv_wgrad(dyn, xn, N, H, W, Cin, 3, stride, 1, dw)
    assert rel(dw.view(Cout, 3, 3, Cin).permute(0, 3, 1, 2), wf.grad) < 1e-2
    if stride == 1:
        w_rot = ops.conv_weight_rotate(w_ohwi, Cout, Cin, 3)
        dx, _, _ = ops.conv_fwd(dyn, N, P, Q, Cout, w_rot, 3, 1, 1)
        assert rel(_nchw(dx, N, H, W), xf.grad) < 1e-2


def test_preprocess_u8_matches_torch_transforms(ops, cuda_device):
    """uint8 HWC -> Resize(256) -> CenterCrop(224) -> /255 -> Normalize, vs torch's antialiased bilinear resize."""
    torch.manual_seed(30)
    sizes = [(300, 400), (768, 512), (224, 224), (1000, 333), (257, 900)]
    imgs = [torch.randint(0, 256, (h, w, 3), dtype=torch.uint8, device=cuda_device) for h, w in sizes]
    out = ops.preprocess_u8(imgs)
    mean = torch.tensor(ops.IMAGENET_MEAN, device=cuda_device).view(3, 1, 1)
    std = torch.tensor(ops.IMAGENET_STD, device=cuda_device).view(3, 1, 1)
    for k, (im, (h, w)) in enumerate(zip(imgs, sizes)):
        nh, nw = (256, int(256 * w / h)) if h <= w else (int(256 * h / w), 256)
        x = im.permute(2, 0, 1).float().unsqueeze(0)
        r = F.interpolate(x, size=(nh, nw), mode="bilinear", antialias=True, align_corners=False)[0]
        top, left = int(round((nh - 224) / 2.0)), int(round((nw - 224) / 2.0))
        ref = (r[:, top:top + 224, left:left + 224] / 255.0 - mean) / std
        assert (out[k] - ref).abs().max().item() < 2e-4, (k, (out[k] - ref).abs().max().item())


def _torch_transform(im, *, resize=256, crop=224, square=False, flip=False):
    """fp32 torch restatement of the reference's tensor transform for one uint8 [H, W, 3] image (on its device)."""
    from b200mm import ops as O
    h, w = im.shape[:2]
    mean = torch.tensor(O.IMAGENET_MEAN, device=im.device).view(3, 1, 1)
    std = torch.tensor(O.IMAGENET_STD, device=im.device).view(3, 1, 1)
    x = im.permute(2, 0, 1).float().unsqueeze(0)
    if square:
        r = F.interpolate(x, size=(crop, crop), mode="bilinear", antialias=True, align_corners=False)[0]
    else:
        nh, nw = (resize, int(resize * w / h)) if h <= w else (int(resize * h / w), resize)
        r = F.interpolate(x, size=(nh, nw), mode="bilinear", antialias=True, align_corners=False)[0]
        top, left = int(round((nh - crop) / 2.0)), int(round((nw - crop) / 2.0))
        r = r[:, top:top + crop, left:left + crop]
    if flip:
        r = r.flip(-1)
    return (r / 255.0 - mean) / std


@pytest.mark.parametrize("square", [False, True])
def test_preprocess_u8_packed_flip_square(ops, cuda_device, square):
    """One pinned byte buffer + offset/height/width table -> the whole transform in one kernel; Resize(256)+CenterCrop
    (.txt:37-41) and Resize((224, 224)) + RandomHorizontalFlip (HEAD script :222-235) variants."""
    torch.manual_seed(31)
    sizes = [(300, 400), (640, 480), (224, 224), (97, 1001), (513, 259), (256, 256)]
    imgs = [torch.randint(0, 256, (h, w, 3), dtype=torch.uint8) for h, w in sizes]
    buf, table = ops.pack_images(imgs)
    assert buf.is_pinned() and table.shape == (3, len(sizes)) and int(table[0, 1]) % 16 == 0
    flip = torch.tensor([0, 1, 1, 0, 1, 0], dtype=torch.uint8, device=cuda_device)
    out = ops.preprocess_u8_packed(buf.to(cuda_device), table.to(cuda_device), square=square, flip=flip)
    assert out.shape == (len(sizes), 3, 224, 224)
    for k, im in enumerate(imgs):
        ref = _torch_transform(im.to(cuda_device), square=square, flip=bool(flip[k]))
        err = (out[k] - ref).abs().max().item()
        assert err < 2e-4, (k, square, err)


def test_u8_normalize_is_exact_and_flips(ops, cuda_device):
    """uint8 [n, H, W, 3] at network resolution -> fp32 NCHW (ToTensor + Normalize), per-image horizontal flip."""
    torch.manual_seed(32)
    x = torch.randint(0, 256, (5, 224, 224, 3), dtype=torch.uint8, device=cuda_device)
    flip = torch.tensor([1, 0, 1, 0, 0], dtype=torch.uint8, device=cuda_device)
    mean = torch.tensor(ops.IMAGENET_MEAN, device=cuda_device).view(1, 3, 1, 1)
    std = torch.tensor(ops.IMAGENET_STD, device=cuda_device).view(1, 3, 1, 1)
    ref = (x.permute(0, 3, 1, 2).float() / 255.0 - mean) / std
    got = ops.u8_normalize(x)
    assert (got - ref).abs().max().item() < 2e-6
    got = ops.u8_normalize(x, flip=flip)
    ref_f = torch.where(flip.view(-1, 1, 1, 1).bool(), ref.flip(-1), ref)
    assert (got - ref_f).abs().max().item() < 2e-6
    with pytest.raises(ValueError):
        ops.u8_normalize(x[:, :, :222].contiguous().permute(0, 2, 1, 3))      # not contiguous HWC


def test_prefetcher_runs_gpu_transform_on_uint8_batches(cuda_device):
    """loop.DevicePrefetcher: uint8 batches (fixed-size and packed) come out as the fp32 tensor the reference's CPU
    transform would have produced; fp32 batches pass through untouched."""
    from b200mm import data as D
    from b200mm.loop import DevicePrefetcher
    torch.manual_seed(33)
    S = 16

    def sample(i, h, w):
        return {"id": f"img_{i}", "text": torch.randint(0, 100, (S,)), "text_mask": torch.ones(S, dtype=torch.long),
                "image": torch.randint(0, 256, (h, w, 3), dtype=torch.uint8), "label": torch.tensor(i & 1)}

    fixed = [sample(i, 224, 224) for i in range(4)]
    ragged = [sample(i, 200 + 37 * i, 300 - 11 * i) for i in range(4)]
    batches = [D.collate_packed(fixed), D.collate_packed(ragged)]
    assert "image" in batches[0] and batches[0]["image"].dtype == torch.uint8 and batches[0]["image"].is_pinned()
    assert "image_packed" in batches[1] and batches[1]["image_table"].shape == (3, 4)
    got = list(DevicePrefetcher(batches, cuda_device))
    assert len(got) == 2
    for (text, image, mask, labels, raw), src in zip(got, (fixed, ragged)):
        assert image.dtype == torch.float32 and image.shape == (4, 3, 224, 224) and image.is_cuda
        assert torch.equal(text.cpu(), torch.stack([s["text"] for s in src]))
        assert raw["id"] == [s["id"] for s in src]
        for k, s in enumerate(src):
            im = s["image"].to(cuda_device)
            ref = _torch_transform(im) if im.shape[0] != 224 else _torch_transform(im, resize=224, crop=224)
            assert (image[k] - ref).abs().max().item() < 2e-4


@pytest.mark.parametrize("H,W", [(224, 224), (96, 130), (45, 67)])
def test_augment_jitter_rotate_matches_torchvision(ops, cuda_device, H, W):
    """ColorJitter (random operator order) + RandomRotation(15) + Normalize of the HEAD script's train transform
    (.py:224-233) as two kernels vs torchvision's tensor operators run on the same GPU with the same draws: element-wise,
    every pixel (nearest-neighbour rotation: a wrong source pixel would be an O(1) error), partial tiles included."""
    from augment_ref import torchvision_augment
    from b200mm.data import GpuImageTransform
    torch.manual_seed(40)
    n = 9
    img = torch.rand(n, 3, H, W, device=cuda_device)
    img[1, :, 10:50, 20:80] = 0.5            # grey patch: hue's max == min branch
    img[2] = (img[2] * 255).round() / 255
    img[3] = 0.0
    img[4] = 1.0
    tr = GpuImageTransform("square", train=True, augment=True, seed=7)
    perm, factors, angles = tr.draw_raw(n)
    angles[5], angles[6], angles[7] = 0.0, 15.0, -15.0
    order, params = tr.pack_augment(perm, factors, angles)
    out, gm = ops.augment_jitter_rotate(img, order.to(cuda_device), params.to(cuda_device))
    ref = torchvision_augment(img, perm, factors, angles, ops.IMAGENET_MEAN, ops.IMAGENET_STD)
    d = (out - ref).abs()
    # A source coordinate that lands within an ulp of x.5 may round to the other neighbour when the grid product is
    # fused differently (cuBLAS bmm vs the kernel's one-rounding-per-op chain): a handful of pixels per million, each
    # then off by a neighbouring pixel's value.  Everything else agrees to the contrast mean's last bits.
    bad = (d > 2e-5).any(dim=1)
    assert bad.float().mean().item() < 1e-4, bad.sum().item()
    assert d[5].max().item() < 2e-5 and d[3].max().item() < 2e-5 and d[4].max().item() < 2e-5   # angle 0 / flat images
    # last row / last column (partial 32 x 8 tiles at 96 x 130) and the zero-filled corners
    assert (d[:, :, -1, :] > 2e-5).float().mean().item() < 5e-3 and (d[:, :, :, -1] > 2e-5).float().mean().item() < 5e-3
    corner = torch.tensor([(0 - m) / s for m, s in zip(ops.IMAGENET_MEAN, ops.IMAGENET_STD)], device=cuda_device)
    assert torch.allclose(out[6, :, 0, 0], corner, atol=1e-6) and torch.allclose(out[7, :, 0, -1], corner, atol=1e-6)
    # the contrast mean is the grey mean of the image after the operators that precede contrast
    for i in (0, 5):
        x = img[i]
        import torchvision.transforms.functional as TF
        for fn in perm[i].tolist():
            if fn == 1:
                break
            x = [lambda v: TF.adjust_brightness(v, float(factors[i, 0])), None,
                 lambda v: TF.adjust_saturation(v, float(factors[i, 2])),
                 lambda v: TF.adjust_hue(v, float(factors[i, 3]))][fn](x)
        assert abs(gm[i].item() - TF.rgb_to_grayscale(x).mean().item()) < 1e-6
    with pytest.raises(ValueError):
        ops.augment_jitter_rotate(img, order.to(cuda_device)[:3], params.to(cuda_device))


def test_gpu_transform_full_head_train_pipeline(cuda_device):
    """data.GpuImageTransform(augment=True): Resize((224, 224)) + RandomHorizontalFlip + ColorJitter + RandomRotation +
    ToTensor + Normalize on packed ragged uint8 images, against the torch restatement of every stage with the
    transform's own draws (a second transform with the same seed replays them)."""
    from augment_ref import torchvision_augment
    from b200mm import ops as O
    from b200mm.data import GpuImageTransform
    torch.manual_seed(41)
    sizes = [(300, 400), (640, 480), (224, 224), (97, 1001), (513, 259)]
    imgs = [torch.randint(0, 256, (h, w, 3), dtype=torch.uint8) for h, w in sizes]
    n = len(imgs)
    buf, table = O.pack_images(imgs)
    tr = GpuImageTransform("square", train=True, augment=True, seed=123)
    out = tr.packed(buf.to(cuda_device), table.to(cuda_device))
    assert out.shape == (n, 3, 224, 224) and out.dtype == torch.float32
    replay = GpuImageTransform("square", train=True, augment=True, seed=123)
    flip = torch.rand(n, generator=replay.gen) < 0.5
    perm, factors, angles = replay.draw_raw(n)
    mean = torch.tensor(O.IMAGENET_MEAN, device=cuda_device).view(3, 1, 1)
    std = torch.tensor(O.IMAGENET_STD, device=cuda_device).view(3, 1, 1)
    img01 = torch.stack([_torch_transform(im.to(cuda_device), square=True, flip=bool(flip[k])) * std + mean
                         for k, im in enumerate(imgs)]).clamp(0, 1)
    ref = torchvision_augment(img01, perm, factors, angles, O.IMAGENET_MEAN, O.IMAGENET_STD)
    # the resize stage agrees to 2e-4 (test_preprocess_u8_packed_flip_square); the colour operators are 1.1-Lipschitz
    # and the hue operator's slope is bounded by ~6, so a handful of pixels may move by a few 1e-3
    # (and, as in the kernel test above, a source coordinate within an ulp of x.5 may pick the other neighbour: such a
    # pixel is off by a neighbouring pixel's value -- counted, not bounded)
    d = (out - ref).abs()
    off = (d > 2e-2).any(dim=1).float().mean().item()
    assert off < 1e-4 and d.mean().item() < 2e-4, (off, d.max().item(), d.mean().item())
    # fixed-size batches take the u8_normalize route
    x = torch.randint(0, 256, (4, 224, 224, 3), dtype=torch.uint8, device=cuda_device)
    tr2 = GpuImageTransform("square", train=True, augment=True, seed=9)
    out2 = tr2.fixed(x)
    replay2 = GpuImageTransform("square", train=True, augment=True, seed=9)
    flip2 = torch.rand(4, generator=replay2.gen) < 0.5
    perm2, factors2, angles2 = replay2.draw_raw(4)
    x01 = x.permute(0, 3, 1, 2).float() / 255.0
    x01 = torch.where(flip2.view(-1, 1, 1, 1).to(cuda_device), x01.flip(-1), x01)
    ref2 = torchvision_augment(x01, perm2, factors2, angles2, O.IMAGENET_MEAN, O.IMAGENET_STD)
    d2 = (out2 - ref2).abs()
    assert (d2 > 1e-4).any(dim=1).float().mean().item() < 1e-4 and d2.mean().item() < 1e-5


@pytest.mark.parametrize("sparse", [False, True])
def test_jpeg_reconstruct_is_bit_identical_to_pillow(cuda_device, sparse):
    """Host Huffman decode + the two device kernels (dequantise + IDCT, up-sampling + colour conversion) over a batch of
    files of every supported coding = ``Image.open(...).convert("RGB")`` PIXEL FOR PIXEL (the reference's loader,
    .txt:50; .py:270); one batch, ragged sizes, images one pixel wide included."""
    import io
    import numpy as np
    from PIL import Image
    from augment_ref import jpeg_cases
    from b200mm import jpeg
    cases = jpeg_cases()
    got = jpeg.decode_jpeg([d for _, d in cases], device=cuda_device, sparse=sparse)   # sparse: only non-zero coefficients ship
    torch.cuda.synchronize()
    assert len(got) == len(cases)
    for (name, data), g in zip(cases, got):
        ref = torch.from_numpy(np.asarray(Image.open(io.BytesIO(data)).convert("RGB")).copy())
        assert g.shape == ref.shape and g.dtype == torch.uint8, name
        assert torch.equal(g.cpu(), ref), (name, (g.cpu().int() - ref.int()).abs().max().item())


def test_jpeg_batches_through_the_prefetcher(cuda_device):
    """jpeg.collate_jpeg batches (coefficients cross PCIe) come out of loop.DevicePrefetcher as the tensor the reference's
    Dataset would have produced from Pillow's pixels; with unsupported='pil' a CMYK file and a PNG ride along."""
    import functools
    import io
    import numpy as np
    from PIL import Image
    from b200mm import jpeg
    from b200mm.loop import DevicePrefetcher
    rng = np.random.default_rng(5)
    S = 16

    def enc(arr, fmt="JPEG", mode=None, **kw):
        b = io.BytesIO()
        im = Image.fromarray(arr)
        (im.convert(mode) if mode else im).save(b, fmt, **kw)
        return b.getvalue()

    def picture(h, w):            # a gradient with mild noise (Pillow's encoder refuses incompressible progressive input)
        g = np.linspace(0, 200, h * w * 3).reshape(h, w, 3) + rng.integers(0, 40, (h, w, 3))
        return g.astype(np.uint8)

    files = [enc(picture(300 + 31 * i, 280 + 17 * i), quality=70 + 5 * i, subsampling=i % 3, progressive=bool(i & 1))
             for i in range(4)]
    files.append(enc(picture(260, 300), mode="CMYK", quality=60))
    files.append(enc(picture(270, 290), "PNG"))
    samples = [{"id": f"img_{i}", "text": torch.randint(0, 100, (S,)), "text_mask": torch.ones(S, dtype=torch.long),
                "image": torch.frombuffer(bytearray(f), dtype=torch.uint8), "label": torch.tensor(i & 1)}
               for i, f in enumerate(files)]
    with pytest.raises(jpeg.UnsupportedJpeg):
        jpeg.collate_jpeg(samples)
    strict = jpeg.collate_jpeg(samples[:4])
    mixed = functools.partial(jpeg.collate_jpeg, unsupported="pil")(samples)
    assert strict["jpeg_coefs"].is_pinned() and strict["jpeg_coefs"].dtype == torch.int16 and not strict["jpeg_raw"]
    assert sorted(i for i, _ in mixed["jpeg_raw"]) == [4, 5]
    got = list(DevicePrefetcher([strict, mixed], cuda_device))
    for (text, image, mask, labels, raw), src in zip(got, (samples[:4], samples)):
        assert image.shape == (len(src), 3, 224, 224) and image.dtype == torch.float32
        for k, smp in enumerate(src):
            px = np.asarray(Image.open(io.BytesIO(bytes(smp["image"].numpy()))).convert("RGB"))
            ref = _torch_transform(torch.from_numpy(px.copy()).to(cuda_device))
            assert (image[k] - ref).abs().max().item() < 2e-4, k
    # the synchronous route of the evaluation loops (loop._to_device) finishes the decode the same way
    from b200mm.loop import _to_device
    _, image2, _, _ = _to_device(strict, cuda_device)
    assert torch.equal(image2, got[0][1])
    # sparse batches (non-zero coefficients only) give the same tensors, Pillow-decoded stragglers included
    for src, want in ((samples[:4], got[0][1]), (samples, got[1][1])):
        sp = jpeg.collate_jpeg(src, unsupported="pil", sparse=True, threads=2)
        assert "jpeg_coefs" not in sp and sp["jpeg_sp_off"].dtype == torch.int32 and sp["jpeg_sp_idx"].is_pinned()
        _, image3, _, _ = _to_device(sp, cuda_device)
        assert torch.equal(image3, want)


# ------------------------------------------------------------------ BERT / RoBERTa / ViT support kernels
def test_position_ids_match_transformers(ops, cuda_device):
    """RoBERTa / XLM-R position ids: cumsum(ids != pad) * (ids != pad) + pad
    (transformers/models/xlm_roberta/modeling_xlm_roberta.py create_position_ids_from_input_ids)."""
    g = torch.Generator().manual_seed(3)
    B, S, pad = 7, 77, 1
    ids = torch.randint(2, 500, (B, S), generator=g)
    lens = torch.randint(1, S + 1, (B,), generator=g)
    ids[torch.arange(S).unsqueeze(0) >= lens.unsqueeze(1)] = pad
    ids[2, 5] = pad        # a pad in the middle of real tokens
    mask = ids.ne(pad).int()
    ref = (torch.cumsum(mask, dim=1) * mask).long() + pad
    got = ops.position_ids(ids.to(cuda_device), pad)
    assert torch.equal(got.cpu().long(), ref)


def test_embedding_variants_fwd_bwd(ops, cuda_device):
    """word + pos[pos_ids] + type[0] -> LN, and the matching scatter-add backward with both padding rows skipped."""
    g = torch.Generator().manual_seed(5)
    B, S, V, D, pad = 4, 24, 300, 128, 1
    ids = torch.randint(2, V, (B, S), generator=g)
    ids[:, 18:] = pad
    word = torch.randn(V, D, generator=g)
    pos = torch.randn(S + 2, D, generator=g)
    typ = torch.randn(D, generator=g)
    gamma, beta = torch.rand(D, generator=g) + 0.5, torch.randn(D, generator=g)
    dev = cuda_device
    pos_ids = ops.position_ids(ids.to(dev), pad)
    y, x_saved, mean, rstd = ops.embed_layernorm_fwd(ids.to(dev), word.to(dev), pos.to(dev), gamma.to(dev),
                                                     beta.to(dev), 1e-5, pos_ids=pos_ids, type_row=typ.to(dev))
    x_ref = word[ids] + pos[pos_ids.cpu().long()] + typ
    y_ref = torch.nn.functional.layer_norm(x_ref, (D,), gamma, beta, 1e-5)
    assert rel(x_saved.view(B, S, D), x_ref) < 5e-3
    assert rel(y.view(B, S, D), y_ref) < 1e-2
    dx = torch.randn(B * S, D, generator=g).to(dev).to(torch.bfloat16)
    dword = torch.zeros(V, D, device=dev)
    dpos = torch.zeros(S + 2, D, device=dev)
    ops.embedding_bwd(dx, ids.to(dev), dword, dpos, padding_idx=pad, pos_ids=pos_ids, pos_padding_idx=pad)
    dxf = dx.float().cpu().view(B, S, D)
    dword_ref = torch.zeros(V, D).index_add_(0, ids.flatten(), dxf.view(-1, D))
    dword_ref[pad] = 0
    dpos_ref = torch.zeros(S + 2, D).index_add_(0, pos_ids.cpu().long().flatten(), dxf.view(-1, D))
    dpos_ref[pad] = 0
    assert rel(dword, dword_ref) < 1e-5 and rel(dpos, dpos_ref) < 1e-5


@pytest.mark.parametrize("M,D", [(300, 256), (20000, 768), (1100, 1024), (96, 2048)])
def test_layernorm_bwd_addend(ops, cuda_device, M, D):
    """pre-LN residual: dx = LN_bwd(dy) + addend.  (20000 rows: > 32 rows per warp, the per-lane row statistics are
    reloaded; 2048: the two-slot ring.)"""
    g = torch.Generator().manual_seed(9)
    dev = cuda_device
    x = torch.randn(M, D, generator=g).to(dev).to(torch.bfloat16)
    dy = torch.randn(M, D, generator=g).to(dev).to(torch.bfloat16)
    add = torch.randn(M, D, generator=g).to(dev).to(torch.bfloat16)
    gamma = (torch.rand(D, generator=g) + 0.5).to(dev)
    beta = torch.zeros(D, device=dev)
    y, mean, rstd = ops.layernorm_fwd(x, gamma, beta, 1e-6)
    dg0, db0 = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    dx0, _ = ops.layernorm_bwd(dy, x, mean, rstd, gamma, dg0, db0)
    dg1, db1 = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    dx1, _ = ops.layernorm_bwd(dy, x, mean, rstd, gamma, dg1, db1, addend=add)
    xr = x.float().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (D,), gamma, beta, 1e-6).backward(dy.float())
    assert rel(dx0, xr.grad) < 1e-2
    assert rel(dx1, xr.grad + add.float()) < 1e-2
    assert torch.equal(dg0, dg1) or rel(dg0, dg1) < 1e-6


def test_vit_assemble_fwd_bwd(ops, cuda_device):
    g = torch.Generator().manual_seed(11)
    B, P, D = 5, 16, 128
    dev = cuda_device
    patch = torch.randn(B * P, D, generator=g).to(dev).to(torch.bfloat16)
    cls = torch.randn(D, generator=g).to(dev)
    pos = torch.randn(P + 1, D, generator=g).to(dev)
    x = ops.vit_assemble_fwd(patch, cls, pos, B, P)
    ref = torch.cat([cls.view(1, 1, D).expand(B, 1, D), patch.float().view(B, P, D)], 1) + pos.view(1, P + 1, D)
    assert rel(x.view(B, P + 1, D), ref) < 5e-3
    dx = torch.randn(B * (P + 1), D, generator=g).to(dev).to(torch.bfloat16)
    dcls = torch.zeros(D, device=dev)
    dpos = torch.zeros(P + 1, D, device=dev)
    dpatch = ops.vit_assemble_bwd(dx, dcls, dpos, B, P)
    dxf = dx.float().view(B, P + 1, D)
    assert torch.equal(dpatch.view(B, P, D), dx.view(B, P + 1, D)[:, 1:])
    assert rel(dpos, dxf.sum(0)) < 1e-5 and rel(dcls, dxf[:, 0].sum(0)) < 1e-5


# ------------------------------------------------------------------ HEAD-script head kernels
@pytest.mark.parametrize("B,C", [(16, 512), (37, 1536), (8, 8)])
def test_bn1d_fwd_bwd(ops, cuda_device, B, C):
    torch.manual_seed(31)
    dev = cuda_device
    x = (torch.randn(B, C, device=dev) * 1.5 + 0.3).to(bf16)
    g = torch.rand(C, device=dev) + 0.5
    b = torch.randn(C, device=dev) * 0.3
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    out, mean, rstd = ops.bn1d_fwd(x, g, b, rm, rv, relu=True, train=True)
    xf = x.float().requires_grad_(True)
    gf, bf_ = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    rm2, rv2 = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    ref = torch.relu(F.batch_norm(xf, rm2, rv2, gf, bf_, True, 0.1, 1e-5))
    assert rel(out, ref) < 1e-2 and rel(rm, rm2) < 1e-4 and rel(rv, rv2) < 1e-4
    dout = torch.randn(B, C, device=dev).to(bf16)
    ref.backward(dout.float())
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    dx = ops.bn1d_bwd(dout, out, x, mean, rstd, g, dg, db, relu=True)
    assert rel(dx, xf.grad) < 3e-2 and rel(dg, gf.grad) < 2e-2 and rel(db, bf_.grad) < 2e-2
    ev, _, _ = ops.bn1d_fwd(x, g, b, rm2, rv2, relu=True, train=False)
    assert rel(ev, torch.relu(F.batch_norm(x.float(), rm2, rv2, g, b, False, 0.1, 1e-5))) < 1e-2


def test_softmax_gate_and_relu_bwd(ops, cuda_device):
    torch.manual_seed(32)
    dev = cuda_device
    B, C = 19, 1536
    a = torch.relu(torch.randn(B, C, device=dev)).to(bf16)
    x = torch.randn(B, C, device=dev).to(bf16)
    y, w = ops.softmax_gate_fwd(a, x)
    af, xf = a.float().requires_grad_(True), x.float().requires_grad_(True)
    ref = torch.softmax(af, 1) * xf
    assert rel(y, ref) < 1e-2 and rel(w, torch.softmax(a.float(), 1)) < 1e-4
    dy = torch.randn(B, C, device=dev).to(bf16)
    ref.backward(dy.float())
    da, dxd = ops.softmax_gate_bwd(dy, w, x)
    assert rel(da, af.grad) < 2e-2 and rel(dxd, xf.grad) < 1e-2
    yv = torch.randn(B, C, device=dev).to(bf16)
    assert torch.equal(ops.relu_bwd(dy, yv), torch.where(yv.float() > 0, dy, torch.zeros_like(dy)))


@pytest.mark.parametrize("B", [16, 300])
def test_head_bn_focal(ops, cuda_device, B):
    """Linear(512,1) + BatchNorm1d(1) + sigmoid_focal_loss, fused, against torch / torchvision."""
    from torchvision.ops import sigmoid_focal_loss
    torch.manual_seed(33)
    dev = cuda_device
    Fdim = 512
    feat = torch.randn(B, Fdim, device=dev).to(bf16)
    W = (torch.randn(1, Fdim, device=dev) / Fdim ** 0.5)
    bias = torch.randn(1, device=dev)
    g, be = torch.tensor([1.3], device=dev), torch.tensor([-0.2], device=dev)
    rm, rv = torch.zeros(1, device=dev), torch.ones(1, device=dev)
    labels = (torch.rand(B, device=dev) < 0.3).long()
    dW, dbias, dg, dbe = (torch.zeros(1, Fdim, device=dev), torch.zeros(1, device=dev), torch.zeros(1, device=dev),
                          torch.zeros(1, device=dev))
    logits, loss, correct, dfeat = ops.head_bn_focal(feat, W, bias, g, be, rm, rv, labels, train=True, bn_train=True,
                                                     dW=dW, dbias=dbias, dg=dg, dbeta=dbe)
    ff = feat.float().requires_grad_(True)
    Wf, bf_, gf, bef = (t.clone().requires_grad_(True) for t in (W, bias, g, be))
    rm2, rv2 = torch.zeros(1, device=dev), torch.ones(1, device=dev)
    z = ff @ Wf.t() + bf_
    yref = F.batch_norm(z, rm2, rv2, gf, bef, True, 0.1, 1e-5).squeeze(1)
    lref = sigmoid_focal_loss(yref, labels.float(), alpha=0.25, gamma=2.0, reduction="mean")
    lref.backward()
    assert rel(logits, yref) < 1e-4 and abs(loss.item() - lref.item()) / lref.item() < 1e-4
    assert correct.item() == ((torch.sigmoid(yref) > 0.5) == (labels != 0)).sum().item()
    assert rel(dfeat, ff.grad) < 1e-2 and rel(dW, Wf.grad) < 1e-3 and rel(dg, gf.grad) < 1e-3
    assert abs(dbe.item() - bef.grad.item()) < 1e-5 + 1e-3 * abs(bef.grad.item())
    assert rel(rm, rm2) < 1e-5 and rel(rv, rv2) < 1e-5
    # eval statistics, no backward
    lg2, loss2, _, _ = ops.head_bn_focal(feat, W, bias, g, be, rm2, rv2, labels, train=False, bn_train=False)
    y2 = F.batch_norm(feat.float() @ W.t() + bias, rm2, rv2, g, be, False, 0.1, 1e-5).squeeze(1)
    assert rel(lg2, y2) < 1e-4


# ------------------------------------------------------------------ CTA-pair (cta_group::2) GEMM
@pytest.mark.parametrize("M,N,K", [(2048, 768, 768), (1157, 768, 1024), (4096, 3072, 768), (3000, 2304, 768)])
def test_gemm_cta_pair_shapes(ops, cuda_device, M, N, K):
    """Shapes big enough for the 256 x 256 CTA-pair kernel (incl. an odd number of 128-row tiles: the second CTA of
    the last pair works on a phantom tile): every operand layout and epilogue it serves, vs fp32 torch."""
    torch.manual_seed(17)
    dev = cuda_device
    x = torch.randn(M, K, device=dev).to(bf16)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(bf16)
    b = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev).to(bf16)
    ref = x.float() @ w.float().t() + b
    assert rel(ops.linear_fwd(x, w, b), ref) < 1e-2
    assert rel(ops.linear_fwd(x, w, b, residual=r), ref + r.float()) < 1e-2
    assert rel(ops.linear_fwd(x, w, b, residual=r, p_drop=0.0), ref + r.float()) < 1e-2
    z, a = ops.linear_gelu_fwd(x, w, b)
    assert rel(z, ref) < 1e-2 and rel(a, F.gelu(ref)) < 1e-2
    dy = torch.randn(M, N, device=dev).to(bf16)
    assert rel(ops.linear_dgrad(dy, w), dy.float() @ w.float()) < 1e-2
    zz = torch.randn(M, K, device=dev).to(bf16)
    zf = zz.float().requires_grad_(True)
    F.gelu(zf).sum().backward()
    assert rel(ops.linear_dgrad(dy, w, gelu_z=zz), (dy.float() @ w.float()) * zf.grad) < 1e-2
    dw = torch.zeros(N, K, device=dev)
    ops.linear_wgrad(dy, x, dw)
    assert rel(dw, dy.float().t() @ x.float()) < 1e-3
    # dropout epilogue keeps the expected scale
    yd = ops.linear_fwd(x, w, b, p_drop=0.25, seed=3)
    kept = (yd != 0).float().mean().item()
    assert abs(kept - 0.75) < 0.02
