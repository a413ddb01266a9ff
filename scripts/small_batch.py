"""Step latency at the batch sizes the reference actually trains with (16: .txt:18 / HEAD :73; 8; 32), config-2 model:
eager launches vs the whole step as one CUDA graph (b200mm.GraphedTrainStep).  Writes gpurun_out/small_batch_r02b.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200mm
from b200mm import _lib
from b200mm.synth import synthetic_batch
dev = torch.device("cuda:0")
STEPS = 30
rows = []
for B in [int(a) for a in sys.argv[1:]] or [8, 16, 32, 256]:
    for mode in ("eager", "graph"):
        model = b200mm.MultimodalClassifier(2, device=dev, seed=42)
        model.train()
        opt = b200mm.FusedAdam(model.parameters(), lr=2e-5)
        crit = b200mm.CrossEntropyLoss()
        d = {k: v.to(dev) for k, v in synthetic_batch(B, 128, seed=1).items()}
        if mode == "graph":
            g = b200mm.GraphedTrainStep(model, opt, crit)
            fn = lambda: g(d["text"], d["image"], d["text_mask"], d["label"])
        else:
            def fn():
                opt.zero_grad()
                out = model.train_step_fused(d["text"], d["image"], d["text_mask"], d["label"])
                opt.step()
                return out
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        _lib.LAUNCHES[0] = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(STEPS):
            out = fn()
        e1.record()
        host_ms = (time.perf_counter() - t0) * 1e3 / STEPS      # host time to ENQUEUE a step
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / STEPS
        rows.append({"batch": B, "mode": mode, "ms_per_step": ms, "samples_per_s": B / ms * 1e3,
                     "host_enqueue_ms_per_step": host_ms, "launch_calls_per_step": _lib.LAUNCHES[0] / STEPS,
                     "loss": float(out[1].item())})
        print(json.dumps(rows[-1]), flush=True)
        if mode == "graph":
            g.close()
        del model, opt
        torch.cuda.empty_cache()
json.dump({"what": "config-2 train step (ResNet-50 + DistilBERT, seq 128, dropout on) at small batch: eager vs CUDA graph",
           "rows": rows}, open("gpurun_out/small_batch_r02b.json", "w"), indent=1)
