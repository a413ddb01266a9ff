"""GPU bring-up check for b200mm_gemm_bf16: every operand layout / epilogue vs torch fp32 matmul."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200mm import _lib as L

dev = torch.device("cuda:0")
torch.manual_seed(0)
stream = torch.cuda.current_stream().cuda_stream
P = lambda t: t.data_ptr() if t is not None else None


def gemm(A, a_mn, B, b_mn, M, N, K, epi=0, bias=None, residual=None, aux=None, out=None, out2=None, splits=1, bn=0):
    L.call("b200mm_gemm_bf16", P(A), a_mn, A.stride(0), P(B), b_mn, B.stride(0), M, N, K, epi,
           P(bias), P(residual), residual.stride(0) if residual is not None else 0,
           P(aux), aux.stride(0) if aux is not None else 0, P(out), out.stride(0),
           P(out2), out2.stride(0) if out2 is not None else 0, splits, bn, 0.0, 0, None, stream)


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


fails = 0
def report(name, err, tol=1e-2):
    global fails
    ok = err < tol
    fails += (not ok)
    print(f"{'OK  ' if ok else 'FAIL'} {name}: rel_err={err:.3e}", flush=True)


for (M, N, K) in [(128, 256, 64), (256, 256, 128), (384, 768, 768), (1000, 512, 200), (4096, 2304, 768), (130, 64, 152)]:
    for bn in (0, 64, 128, 256):
        A = torch.randn(M, K, device=dev).bfloat16()
        W = torch.randn(N, K, device=dev).bfloat16()
        bias = torch.randn(N, device=dev)
        ref = A.float() @ W.float().t()
        out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        gemm(A, 0, W, 0, M, N, K, 0, None, None, None, out, bn=bn)
        torch.cuda.synchronize()
        report(f"TN plain M{M} N{N} K{K} bn{bn}", rel(out, ref))
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    res = torch.randn(M, N, device=dev).bfloat16()
    gemm(A, 0, W, 0, M, N, K, 0, bias, res, None, out)
    report(f"TN bias+res M{M} N{N} K{K}", rel(out, ref + bias + res.float()))
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    gemm(A, 0, W, 0, M, N, K, 5, bias, None, None, out)
    report(f"TN relu M{M} N{N} K{K}", rel(out, torch.relu(ref + bias)))
    # GELU dual output
    A2 = (A.float() * 0.05).bfloat16()
    ref2 = A2.float() @ W.float().t() + bias
    z = torch.zeros(M, N, device=dev, dtype=torch.bfloat16); g = torch.zeros_like(z)
    gemm(A2, 0, W, 0, M, N, K, 1, bias, None, None, z, g)
    report(f"TN gelu z M{M} N{N} K{K}", rel(z, ref2))
    report(f"TN gelu a M{M} N{N} K{K}", rel(g, torch.nn.functional.gelu(ref2)))
    # dGELU epilogue
    zz = torch.randn(M, N, device=dev).bfloat16()
    d = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    gemm(A, 0, W, 0, M, N, K, 2, None, None, zz, d)
    zf = zz.float().requires_grad_(True); torch.nn.functional.gelu(zf).sum().backward()
    report(f"TN dgelu M{M} N{N} K{K}", rel(d, ref * zf.grad))
    # fp32 output
    o32 = torch.zeros(M, N, device=dev)
    gemm(A, 0, W, 0, M, N, K, 3, bias, None, None, o32)
    report(f"TN f32 M{M} N{N} K{K}", rel(o32, ref + bias), 1e-5)
    # dgrad: dX[M,K'] = dY[M,N'] W[N',K']  -> A=dY K-major, B=W stored [N',K'] = [K_red, N_out] MN-major
    if K % 8 == 0:
        dY = torch.randn(M, N, device=dev).bfloat16()
        dX = torch.zeros(M, K, device=dev, dtype=torch.bfloat16)
        gemm(dY, 0, W, 1, M, K, N, 0, None, None, None, dX)
        report(f"NN dgrad M{M} N{K} K{N}", rel(dX, dY.float() @ W.float()))
        # wgrad: dW[N,K] = dY^T[N,M] X[M,K] : A = dY stored [M,N] = [K_red, M_out] MN-major; B = X stored [M,K] MN-major
        for splits in (1, 3, 8):
            dW = torch.zeros(N, K, device=dev)
            gemm(dY, 1, A, 1, N, K, M, 4 if splits > 1 else 3, None, None, None, dW, splits=splits)
            report(f"NT wgrad N{N} K{K} M{M} splits{splits}", rel(dW, dY.float().t() @ A.float()), 1e-3)
torch.cuda.synchronize()

# ---- timing at the config-2 shapes
def bench(M, N, K, a_mn=0, b_mn=0, epi=0, splits=1, iters=20):
    if a_mn == 0:
        A = torch.randn(M, K, device=dev).bfloat16()
    else:
        A = torch.randn(K, M, device=dev).bfloat16()
    B = torch.randn(N, K, device=dev).bfloat16() if b_mn == 0 else torch.randn(K, N, device=dev).bfloat16()
    out = torch.zeros(M, N, device=dev, dtype=torch.float32 if epi in (3, 4) else torch.bfloat16)
    out2 = torch.zeros(M, N, device=dev, dtype=torch.bfloat16) if epi == 1 else None
    bias = torch.zeros(N, device=dev)
    for _ in range(3):
        gemm(A, a_mn, B, b_mn, M, N, K, epi, bias if epi != 4 else None, None, None, out, out2, splits=splits)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        gemm(A, a_mn, B, b_mn, M, N, K, epi, bias if epi != 4 else None, None, None, out, out2, splits=splits)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # cuBLAS for comparison
    if a_mn == 0 and b_mn == 0:
        for _ in range(3): torch.matmul(A, B.t())
        e0.record()
        for _ in range(iters): torch.matmul(A, B.t())
        e1.record(); torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / iters
        print(f"bench M{M} N{N} K{K} a_mn{a_mn} b_mn{b_mn} epi{epi} splits{splits}: {ms:.3f} ms {tf:.0f} TF/s | cuBLAS {ms2:.3f} ms {2.0*M*N*K/ms2/1e9:.0f} TF/s", flush=True)
    else:
        print(f"bench M{M} N{N} K{K} a_mn{a_mn} b_mn{b_mn} epi{epi} splits{splits}: {ms:.3f} ms {tf:.0f} TF/s", flush=True)

bench(32768, 2304, 768)
bench(32768, 768, 768)
bench(32768, 3072, 768, epi=1)
bench(32768, 768, 3072)
bench(32768, 768, 3072, b_mn=1)            # dgrad lin1-like
bench(32768, 3072, 768, b_mn=1, epi=0)     # dgrad lin2-like
bench(3072, 768, 32768, a_mn=1, b_mn=1, epi=4, splits=8)   # wgrad lin1
bench(768, 768, 32768, a_mn=1, b_mn=1, epi=4, splits=16)
bench(8192, 8192, 8192)
print("FAILS", fails)
sys.exit(1 if fails else 0)
