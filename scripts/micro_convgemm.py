"""1x1-convolution GEMM shapes of ResNet-50 (config 2, batch 256): CTA-pair vs single-CTA kernel, tile widths, with a
correctness check of every variant (output max-abs / column statistics vs torch fp32).

    python scripts/micro_convgemm.py            # timing table (L2 flushed between launches)
    python scripts/micro_convgemm.py ncu        # a few launches of the L2-bound shapes for an ncu capture
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200mm import ops, _lib
dev = torch.device("cuda:0"); bf = torch.bfloat16
torch.manual_seed(0)
lib = _lib.load()
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def timeit(fn, iters=8):
    for _ in range(2): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e3


def tune(knob, v):
    assert lib.b200mm_tune(knob, v) == 0


# (M, N, K, b_mn, mode): mode "stats" = forward conv (column statistics), "plain" = dgrad store, "acc" = dgrad accumulating
# into the residual gradient (TMA reduce-add), "res" = store + residual read
SHAPES = [
    (50000, 1000, 200, 0, "stats"), (50000, 1000, 200, 1, "res"),      # ragged: partial row / column / k tiles
    (50176, 1024, 256, 0, "stats"), (50176, 1024, 256, 1, "plain"), (50176, 1024, 256, 1, "acc"),
    (50176, 256, 1024, 0, "stats"), (50176, 256, 1024, 1, "plain"),
    (12544, 2048, 512, 0, "stats"), (12544, 2048, 512, 1, "acc"),
    (12544, 512, 2048, 0, "stats"), (12544, 512, 2048, 1, "plain"),
    (200704, 512, 128, 0, "stats"), (200704, 512, 128, 1, "acc"),
    (200704, 128, 512, 0, "stats"), (200704, 128, 512, 1, "plain"),
    (802816, 256, 64, 0, "stats"), (802816, 256, 64, 1, "acc"),
    (802816, 64, 256, 0, "stats"), (802816, 64, 256, 1, "plain"),
    (200704, 256, 512, 0, "stats"), (50176, 512, 1024, 0, "stats"), (50176, 1024, 512, 0, "stats"),
    (12544, 2048, 1024, 0, "stats"),
]
if "ncu" in sys.argv:
    SHAPES = [(50176, 1024, 256, 0, "stats"), (200704, 512, 128, 1, "acc"), (12544, 512, 2048, 0, "stats")]

bad = 0
for (M, N, K, b_mn, mode) in SHAPES:
    A = torch.randn(M, K, device=dev).to(bf)
    Bm = ((torch.randn(N, K, device=dev) if not b_mn else torch.randn(K, N, device=dev)) / K ** 0.5).to(bf)
    out = torch.zeros(M, N, device=dev, dtype=bf)
    base = torch.randn(M, N, device=dev).to(bf) if mode in ("acc", "res") else None
    st = torch.zeros(2 * N, device=dev) if mode == "stats" else None
    rows = torch.cat([torch.arange(0, 300, device=dev), torch.arange(M - 300, M, device=dev)])
    Bf = Bm.float() if b_mn else Bm.float().t()
    ref_rows = A[rows].float() @ Bf
    if base is not None:
        ref_rows = ref_rows + base[rows].float()

    def run(bn=0):
        if mode == "stats":
            ops.gemm_raw(A, False, Bm, bool(b_mn), M, N, K, out, epi=0, col_stats=st, block_n=bn)
        elif mode == "acc":
            ops.gemm_raw(A, False, Bm, bool(b_mn), M, N, K, out, epi=0, residual=out, block_n=bn)
        elif mode == "res":
            ops.gemm_raw(A, False, Bm, bool(b_mn), M, N, K, out, epi=0, residual=base, block_n=bn)
        else:
            ops.gemm_raw(A, False, Bm, bool(b_mn), M, N, K, out, epi=0, block_n=bn)

    def check(tag):
        global bad
        if mode == "acc":
            out.copy_(base)
        if mode == "stats":
            st.zero_()
        run()
        torch.cuda.synchronize()
        err = (out[rows].float() - ref_rows).abs().max().item()
        tol = 0.06 if base is None else 0.12
        msg = f"max-abs {err:.3f}"
        ok = err < tol
        if mode == "stats":
            o = out.float()
            s_err = ((st[:N] - o.sum(0)).abs().max() / (o.sum(0).abs().max() + 1e-6)).item()
            q_err = ((st[N:] - (o * o).sum(0)).abs().max() / (o * o).sum(0).abs().max()).item()
            msg += f" stats {s_err:.1e}/{q_err:.1e}"
            ok = ok and s_err < 1e-3 and q_err < 1e-3
        if not ok:
            bad += 1
        return ("ok " if ok else "FAIL ") + msg

    hbm = 2.0 * (M * K + N * K + M * N * (2 if base is not None else 1))
    line = f"[{M:7d} x{N:5d} x{K:5d}] b_mn{b_mn} {mode:5s}"
    # (pair min k, B-resident mode)
    configs = {"base": (8, 0), "pair": (1, 0), "bres": (8, 1)}
    if "ncu" in sys.argv:
        configs = {"base": (8, 0), "bres": (8, 1)}
    res = []
    for name, kn in configs.items():
        for k, v in enumerate(kn):
            tune(k, v)
        c = check(name)
        t = 0.0 if "ncu" in sys.argv else timeit(run)
        res.append(f"{name} {t:6.1f}" + ("" if c.startswith("ok") else " " + c))
    print(line, " | ".join(res), f"| hbm floor {hbm / 6553.6e3:5.1f} us", flush=True)
    del A, Bm, out, base, st
for k, v in enumerate((8, 1)):
    tune(k, v)
sys.exit(1 if bad else 0)
