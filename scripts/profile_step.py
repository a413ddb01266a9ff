"""Per-op CUDA-event profile of the real (pipelined, warm-L2) train step (PCFG = BASELINE config, default 2).
Aggregates by C-ABI entry point, and by shape for the GEMM.  Writes profiles/step_profile_<tag>.json."""
import json, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, b200mm
import b200mm.model as _m
_m._TOWER_OVERLAP = False     # per-op events need one stream (the two-stream tower overlap interleaves kernels)
from b200mm import _lib
from b200mm.synth import synthetic_batch
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
import bench
CFG = int(os.environ.get("PCFG", 2))
B = int(os.environ.get("PB", bench.WORKLOADS[CFG]["batch"]))
S = bench.WORKLOADS[CFG]["seq"]
dev = torch.device("cuda:0")
model, synth_kw = bench.build_model(CFG, dev); model.train()
opt = b200mm.FusedAdam(model.parameters(), lr=2e-5)
d = {k: v.to(dev) for k, v in synthetic_batch(B, S, **synth_kw).items()}
def step():
    opt.zero_grad(); model.train_step_fused(d["text"], d["image"], d["text_mask"], d["label"]); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): step()
e1.record(); torch.cuda.synchronize()
plain = e0.elapsed_time(e1) / 3
_lib.PROFILE = []
STEPS = 2
e0.record()
for _ in range(STEPS): step()
e1.record(); torch.cuda.synchronize()
prof_ms = e0.elapsed_time(e1) / STEPS
rec, _lib.PROFILE = _lib.PROFILE, None
by_fn = collections.defaultdict(lambda: [0, 0.0]); by_gemm = collections.defaultdict(lambda: [0, 0.0])
for name, key, a, b in rec:
    ms = a.elapsed_time(b)
    by_fn[name][0] += 1; by_fn[name][1] += ms
    if key is not None: by_gemm[key][0] += 1; by_gemm[key][1] += ms
tot = sum(v[1] for v in by_fn.values()) / STEPS
print(f"step {plain:.2f} ms (with events {prof_ms:.2f} ms); sum of op times {tot:.2f} ms")
rep = {"config": CFG, "batch": B, "ms_per_step": plain, "ms_per_step_profiled": prof_ms, "ops": {}, "gemm": []}
for k, v in sorted(by_fn.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:32s} n={v[0]//STEPS:4d} {v[1]/STEPS:8.3f} ms")
    rep["ops"][k] = {"calls": v[0] // STEPS, "ms": v[1] / STEPS}
print("GEMM by shape (M,N,K,a_mn,b_mn,epi,splits):")
for k, v in sorted(by_gemm.items(), key=lambda kv: -kv[1][1])[:90]:
    if isinstance(k[0], str) and k[0].startswith("bn_"):
        M, C = k[1:3]
        by = (4.0 + 2.0 * k[3]) * M * C if k[0] == "bn_fwd" else (10.0 + 2.0 * k[3]) * M * C
        print(f"  {str(k):46s} n={v[0]//STEPS:3d} {v[1]/STEPS:7.3f} ms  {by*v[0]/v[1]/1e6:7.0f} GB/s (algorithmic)")
        rep["gemm"].append({"key": list(k), "calls": v[0] // STEPS, "ms": v[1] / STEPS, "gbs": by * v[0] / v[1] / 1e6})
        continue
    if isinstance(k[0], str):
        M, N, K = k[1:4]
        by = 2.0 * (M * N + N * K + M * K / 9.0) if k[0] == "conv_fwd" else 2.0 * (K * M + K * N / 9.0 + 2 * M * N)
    else:
        M, N, K = k[:3]
        by = 2.0 * (M * K + N * K + M * N * (2 if k[5] in (3, 4) else 1))
    fl = 2.0 * M * N * K * v[0]
    print(f"  {str(k):46s} n={v[0]//STEPS:3d} {v[1]/STEPS:7.3f} ms  {fl/v[1]/1e9:7.0f} TF/s  {by*v[0]/v[1]/1e6:7.0f} GB/s")
    rep["gemm"].append({"key": list(k), "calls": v[0] // STEPS, "ms": v[1] / STEPS, "tflops": fl / v[1] / 1e9, "gbs": by * v[0] / v[1] / 1e6})
os.makedirs("profiles", exist_ok=True)
json.dump(rep, open(f"profiles/step_profile_{tag}.json", "w"), indent=1)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rep, open(f"gpurun_out/step_profile_{tag}.json", "w"), indent=1)
