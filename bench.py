#!/usr/bin/env python
"""Headline benchmark: train samples/s of the task-2C late-fusion classifier step (BASELINE.json configs[1]:
ResNet-50 + DistilBERT-multilingual, bf16, batch 256 per GPU, 224 px images + 128-token text).

    python bench.py --gpus N --steps K --warmup W            # this engine (one process per GPU under torchrun)
    python bench.py --impl reference ...                      # the reference's CPU train step (oracle) on host cores

One "step" = zero_grad -> forward -> cross-entropy -> backward -> Adam over one batch of synthetic inputs
(example_scripts/Multimodal_example_task2C.txt:204-220).  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_TRAIN_PER_SAMPLE = 57.8   # 3 x (8.18 ResNet-50 + 11.17 DistilBERT-S128) dense GFLOP
METRIC = "train samples/s (224px img + 128-tok text)"

# The headline (default) workload is BASELINE configs[1]; --config 3 / 4 / 5 run the other configurations of
# BASELINE.json ("same head, towers swapped") for the record -- the driver only runs the default.
# gflop = dense train GFLOP / sample (fwd x 3) resp. forward GFLOP / ensembled sample for config 5.
WORKLOADS = {
    2: dict(name="ResNet-50 + DistilBERT-multilingual late fusion train step (fwd + CE + bwd + Adam), "
                 "BASELINE configs[1]", batch=256, seq=128, gflop=57.8),
    3: dict(name="ViT-B/16 + BERT-base (vocab 64000) late fusion train step (fwd + CE + bwd + Adam), "
                 "BASELINE configs[2]", batch=256, seq=128, gflop=172.2),
    4: dict(name="ViT-L/14 + XLM-R-large late fusion train step (fwd + CE + bwd + Adam), seq 256, "
                 "BASELINE configs[3]", batch=64, seq=256, gflop=969.0),
    5: dict(name="5-fold ensemble inference (5 x ViT-B/16 + BERT-base forward, mean prob), BASELINE configs[4]",
            batch=1024, seq=128, gflop=287.4),
}


def build_model(config: int, dev, seed: int = 42, num_classes: int = 2):
    import b200mm
    if config == 2:
        return b200mm.MultimodalClassifier(num_classes, device=dev, seed=seed), dict(vocab_size=119547, pad_id=0)
    if config in (3, 5):
        t = b200mm.TextConfig.bert_base()
        v = b200mm.ViTConfig.vit_b16()
    else:
        t = b200mm.TextConfig.xlmr_large()
        v = b200mm.ViTConfig.vit_l14()
    m = b200mm.MultimodalClassifier(num_classes, text_config=t, image_config=v, device=dev, seed=seed)
    return m, dict(vocab_size=t.vocab_size, pad_id=t.pad_token_id)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(WORKLOADS),
                    help="BASELINE.json configuration (1-based; default 2 = the headline)")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (0 = the configuration's)")
    ap.add_argument("--seq", type=int, default=0)
    ap.add_argument("--cpu-batch", type=int, default=16, help="batch of the CPU reference step (BASELINE config 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true",
                    help="launch the step's kernels one by one instead of replaying the captured CUDA graph (N = 1)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_tflops": p.get("bf16_tflops"), "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "hbm_gbs": p.get("hbm_gbs"), "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU implementation of the step (the oracle module = the reference's module code on the
    stock PyTorch/transformers/torchvision CPU path), all host threads, bounded sample (batch 16 per step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import reference_model as R
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(42)
    model = R.MultimodalClassifier(2)
    model.train()
    crit = torch.nn.CrossEntropyLoss()
    opt = torch.optim.Adam(model.parameters(), lr=2e-5)
    data = R.synthetic_batch(args.cpu_batch, args.seq)
    # the driver's --steps / --warmup are honoured; only capped (20 / 2) so that the CPU arm -- about one second per
    # batch-16 step on 16 cores -- stays within a few minutes whatever it asks for
    steps, warmup = min(max(1, args.steps), 20), min(max(0, args.warmup), 2)
    for _ in range(warmup):
        R.train_step(model, data, crit, opt)
    t0 = time.perf_counter()
    for _ in range(steps):
        R.train_step(model, data, crit, opt)
    dt = (time.perf_counter() - t0) / steps
    v = args.cpu_batch / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ResNet-50 + DistilBERT-multilingual late fusion train step, 224px, seq 128 "
                               "(BASELINE configs[1] graph; CPU sample = batch %d per step)" % args.cpu_batch},
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": threads, "kind": "port",
                         "sample": f"{steps} steps of batch {args.cpu_batch} after {warmup} warm-up, fp32, "
                                   f"torch CPU, oracle/reference_model.py"},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ------------------------------------------------------------------------------------------------- config 5
def run_ensemble(args, dev, world, rank):
    """BASELINE configs[4]: 5-fold ensemble inference.  Five fold-models (same architecture, different weights) are
    resident on every GPU; a step = the five eval-mode forward passes over one batch, sigmoid of the single logit
    (HEAD script, Multimodal_example_task2C.py:808-810, 843-851) and the per-id mean over folds
    (combine_preds.py:29-31).  Samples are sharded across ranks; value = ensembled samples/s."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import b200mm
    from b200mm import _lib, ensemble
    from b200mm.synth import synthetic_batch
    B, S, wl = args.batch, args.seq, WORKLOADS[5]
    folds = 5
    models = []
    for k in range(folds):
        t, v = b200mm.TextConfig.bert_base(), b200mm.ViTConfig.vit_b16()
        m = b200mm.MultimodalClassifier(1, text_config=t, image_config=v, device=dev, seed=42 + k, pooling="cls",
                                        squeeze_output=True)
        m.eval()
        models.append(m)
    host = synthetic_batch(B, S, seed=1234 + rank, vocab_size=64000)
    host = {k: v.pin_memory() for k, v in host.items()}
    devd = {k: v.to(dev) for k, v in host.items()}
    ids = [f"img_{rank}_{i}" for i in range(B)]
    h2d = sum(host[k].numel() * host[k].element_size() for k in ("text", "text_mask", "image"))
    logits_dev = torch.empty(folds, B, device=dev, dtype=torch.float32)
    logits_host = torch.empty(folds, B, dtype=torch.float32).pin_memory()

    def forward_all(d):
        with torch.no_grad():
            for k, m in enumerate(models):
                logits_dev[k].copy_(m(d["text"], d["image"], d["text_mask"]))

    def step_resident():
        forward_all(devd)

    # multi-GPU tail: every rank's per-fold logits are gathered on rank 0, which averages the folds'
    # probabilities per id (combine_preds.py:29-31) and, after the run, writes the five per-fold probability TSVs
    # (Multimodal_example_task2C.py:869-879) and the ensembled submission through b200mm.combine_folds
    gathered = torch.empty(world, folds, B, device=dev, dtype=torch.float32) if rank == 0 else None
    gathered_host = torch.empty(world, folds, B, dtype=torch.float32).pin_memory() if rank == 0 else None
    all_ids = [f"data/synth/img_{r}_{i}.jpg" for r in range(world) for i in range(B)]
    kept = {}

    def step_e2e():
        d = {k: host[k].to(dev, non_blocking=True) for k in ("text", "text_mask", "image")}
        forward_all(d)
        if world > 1:
            dist.gather(logits_dev, gather_list=list(gathered.unbind(0)) if rank == 0 else None, dst=0)
        elif rank == 0:
            gathered[0].copy_(logits_dev)
        if rank != 0:
            return 0
        gathered_host.copy_(gathered, non_blocking=False)
        probs = 1.0 / (1.0 + np.exp(-gathered_host.numpy().astype(np.float64)))     # [world, folds, B]
        per_fold = [probs[:, k, :].reshape(-1) for k in range(folds)]
        _, mean_prob = ensemble.average_probability([all_ids] * folds, per_fold)
        kept["per_fold"] = per_fold
        return int((mean_prob > 0.5).sum())

    def write_tail(out_dir):
        """rank 0, after the timed region: per-fold prob TSVs of the last batch + the ensembled label TSV."""
        from b200mm import combine_folds, tsv
        t0 = time.perf_counter()
        paths = []
        for k, p in enumerate(kept["per_fold"]):
            path = os.path.join(out_dir, f"task2C_bench_probs_fold_{k}.tsv")
            labels = ["propaganda" if x > 0.5 else "not_propaganda" for x in p]
            tsv.write_prob_tsv(path, all_ids, labels, p, "bench_vit-b16_bert-base")
            paths.append(path)
        out = os.path.join(out_dir, "task2C_bench_ensemble.tsv")
        ids, mean_prob, labels, _, _ = combine_folds(paths, None, out_path=out, log=lambda *_: None)
        assert len(ids) == world * B and tsv.check_label_tsv(out)
        return {"rows": len(ids), "files": len(paths) + 1, "seconds": time.perf_counter() - t0}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(dev.index)
    sampler.start()
    _lib.LAUNCHES[0] = 0
    ms = timed(step_resident, args.steps) / args.steps
    launches = _lib.LAUNCHES[0]
    clocks = sampler.stop()
    value = world * B / (ms * 1e-3)
    pk = peaks()
    tfl = value / world * wl["gflop"] / 1e3
    roofline = {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05+TMA)", "achieved": tfl,
                "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": tfl / pk["bf16_tflops_sustained"],
                "traffic": None, "peak_source": pk["source"],
                "note": "whole-forward dense FLOPs (5 x 57.5 GFLOP per ensembled sample) / step time"}
    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            step_e2e()
        ms_e = timed(step_e2e, args.steps) / args.steps
        e2e = {"value": world * B / (ms_e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": folds * B * 4 * world, "ms_per_step": ms_e,
               "path": "per-rank forward of 5 resident fold models, logits gathered on rank 0 (NCCL gather), "
                       "sigmoid + per-id mean over folds on the host every step"}
        if rank == 0:
            import tempfile
            with tempfile.TemporaryDirectory() as td:
                e2e["tsv_tail"] = write_tail(td)
    if rank == 0:
        print(json.dumps({
            "metric": "ensembled inference samples/s (5 folds, 224px img + 128-tok text)", "value": value,
            "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": wl["name"], "per_gpu_batch": B, "global_batch": B * world, "seq_len": S,
                       "image": "3x224x224", "parallelism": f"dp{world}", "folds": folds,
                       "l2": "inputs+activations per step >> 126 MB L2 (no flush needed)"},
            "roofline": roofline, "cpu_baseline": None, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }), flush=True)
    if world > 1:
        _shutdown(dist)


def _shutdown(dist, gstep=None):
    """Tear the process group down.  A captured graph that holds NCCL launches must die before its communicator
    (destroying the group first hung the ranks); a watchdog ends the process if the teardown still blocks -- the
    result line is already out."""
    import gc
    import torch
    sys.stdout.flush()
    killer = threading.Timer(20.0, lambda: os._exit(0))
    killer.daemon = True
    killer.start()
    if gstep is not None:
        gstep.close()
        gstep.static_out = gstep.static_in = None
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    killer.cancel()


# ------------------------------------------------------------------------------------------------- engine arm
def run_engine(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import b200mm
    from b200mm import _lib, ops
    from b200mm.synth import synthetic_batch

    B, S = args.batch, args.seq
    wl = WORKLOADS[args.config]
    if args.config == 5:
        return run_ensemble(args, dev, world, rank)
    model, synth_kw = build_model(args.config, dev)
    if world > 1 and not os.environ.get("B200MM_NO_SYNC"):   # B200MM_NO_SYNC=1: N independent replicas (diagnostic only)
        model.enable_data_parallel()
    model.train()
    crit = b200mm.CrossEntropyLoss()
    opt = b200mm.FusedAdam(model.parameters(), lr=2e-5)

    host = synthetic_batch(B, S, seed=1234 + rank, **synth_kw)
    host = {k: v.pin_memory() for k, v in host.items()}
    devd = {k: v.to(dev) for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    def step_eager():
        opt.zero_grad()
        _, loss, ok = model.train_step_fused(devd["text"], devd["image"], devd["text_mask"], devd["label"])
        opt.step()
        return loss

    # one process: the loop body (zero_grad / forward / loss / backward / Adam) is captured once as a CUDA graph and
    # replayed -- the same ~500 kernels, enqueued by one cudaGraphLaunch instead of 500 host calls (at batch 256 the
    # host needs 30 of the step's 31 ms to enqueue them one by one: profiles/small_batch_r02.json).  Data parallel:
    # the NCCL all-reduces of the gradient phases are captured with the step (forked / joined by events).
    gstep = b200mm.GraphedTrainStep(model, opt, crit) if not args.no_graph else None

    def step_resident():
        if gstep is None:
            return step_eager()
        return gstep(devd["text"], devd["image"], devd["text_mask"], devd["label"])[1]

    # end-to-end: b200mm.train() ITSELF over a torch DataLoader(pin_memory=True), the call a user of the reference
    # makes (.txt:251-257).  The dataset hands out what a decoder produces -- uint8 HWC pixels at network resolution +
    # cached token ids (b200mm.data) -- so a step's batch crosses PCIe as 38.5 MB instead of the reference's 154 MB of
    # fp32 pixels; train()'s DevicePrefetcher copies batch i+1 on a side stream under step i and runs ToTensor +
    # Normalize there as one kernel; every step's loss / correct count is read back (one step late, .txt:218-220).
    import itertools
    import shutil
    from torch.utils.data import DataLoader, Dataset

    class SyntheticMemes(Dataset):
        def __init__(self, n):
            g = torch.Generator().manual_seed(99 + rank)
            self.n = n
            self.pixels = torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8, generator=g)

        def __len__(self):
            return self.n

        def __getitem__(self, i):
            j = i % B
            return {"id": f"img_{i}", "text": host["text"][j], "text_mask": host["text_mask"][j],
                    "image": self.pixels[j], "label": host["label"][j]}

    class Primed:
        """One epoch of the DataLoader, consumed by successive train() calls (warm-up, then the timed K steps): the
        worker processes are already prefetching when the timed call starts, as they are in the middle of an epoch."""

        def __init__(self, loader):
            self.it, self.k = iter(loader), 0

        def take(self, k):
            self.k = k
            return self

        def __iter__(self):
            return itertools.islice(self.it, self.k)

    e2e_warm = int(os.environ.get("B200MM_E2E_WARM", "8"))
    e2e_log = []

    def make_e2e_loader(workers, steps):
        kw = dict(prefetch_factor=2, persistent_workers=False) if workers else {}
        dl = DataLoader(SyntheticMemes(steps * B), batch_size=B, shuffle=False, drop_last=True,
                        num_workers=workers, pin_memory=True, **kw)
        return Primed(dl)

    def pick_e2e_workers():
        """DataLoader worker count from a short measurement on THIS host: with the step replayed as one CUDA graph the
        launching thread is idle, and what decides is whether the loader processes keep up without fighting the
        pin-memory thread for cores (16-core box, graph mode: 0 / 1 / 2 / 3 workers -> 31.7 / 30.6 / 38.7 / 34.7 ms per
        step against 30.2 resident, profiles/e2e_workers_r02.log).  Every rank runs the same number of probe steps."""
        if "B200MM_E2E_WORKERS" in os.environ:
            return int(os.environ["B200MM_E2E_WORKERS"]), None
        shm_free = shutil.disk_usage("/dev/shm").free if os.path.isdir("/dev/shm") else 0
        cands = [1, 2, 0] if shm_free > (4 << 30) and (os.cpu_count() or 2) >= 4 * world else [0]
        probe = {}
        for w in cands:
            # (8 untimed steps first: the worker processes' start-up and the first, empty-queue batches are not what
            # is being compared -- with 3 the same box reported 30 and 53 ms for one worker on two runs)
            primed = make_e2e_loader(w, 8 + 8)
            e2e_train(primed, 8)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e2e_train(primed, 8)
            torch.cuda.synchronize()
            probe[w] = (time.perf_counter() - t0) / 8 * 1e3
            del primed
        return min(probe, key=probe.get), {str(k): round(v, 2) for k, v in probe.items()}

    def e2e_train(primed, k):
        return b200mm.train(model, primed.take(k), crit, opt, dev, on_step=lambda loss, bs: e2e_log.append(loss),
                            graph=gstep)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local_rank)
    sampler.start()
    _lib.LAUNCHES[0] = 0
    ms = timed(step_resident, args.steps)
    launches = _lib.LAUNCHES[0] if gstep is None else gstep.launches_per_replay * args.steps
    clocks = sampler.stop()
    ms_step = ms / args.steps
    value = world * B / (ms_step * 1e-3)
    sync_info = None
    gs = getattr(model, "grad_sync", None)
    if world > 1 and gs is not None:
        # exposed tail of the gradient exchange: device time the compute stream waited in GradSync.finish() (inside
        # FusedAdam.step) during the LAST timed step, max over ranks
        w = torch.tensor([gs.exposed_wait_ms()], device=dev)
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
        sync_info = {"payload": gs.payload, "bytes_per_step_per_rank": gs.bytes_per_step,
                     "phases": {k: len(v) for k, v in gs.phases.items()}, "exposed_wait_ms_last_step": w.item() if gstep is None else None,
                     "note": None if gstep is None else "captured in the step graph: no timing events inside a capture "
                                                        "(run with --no-graph for the exposed-wait measurement)"}

    # ---- roofline of the dominant kernel (the tcgen05 GEMM): per-launch CUDA events on the launching stream
    pk = peaks()
    ridge = pk["bf16_tflops_sustained"] * 1e12 / (pk["hbm_gbs"] * 1e9)        # FLOP per byte
    # per-launch events need the launches back to back on ONE stream: the two-stream tower overlap is switched off for
    # this (untimed) profiling pass, otherwise a GEMM's interval also covers kernels of the other tower
    from b200mm import model as _model_mod
    _overlap, _model_mod._TOWER_OVERLAP = _model_mod._TOWER_OVERLAP, False
    _wq_on, ops.WGRAD_OVERLAP = ops.WGRAD_OVERLAP, False
    for _tw in (getattr(model, "text", None), getattr(model, "img", None)):
        if _tw is not None and getattr(_tw, "_wq", None) is not None:
            _tw._wq = None                      # (a side queue created earlier keeps its stream: rebuild it serial)
    gemm_ms, gemm_flops, n_gemm, det = ops.profile_gemm(step_eager if gstep is None else (lambda: gstep.eager(devd["text"], devd["image"], devd["text_mask"], devd["label"])), steps=2, ridge=ridge)
    _model_mod._TOWER_OVERLAP = _overlap
    ops.WGRAD_OVERLAP = _wq_on
    for _tw in (getattr(model, "text", None), getattr(model, "img", None)):
        if _tw is not None:
            _tw._wq = None
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    t, h = det["tensor"], det["hbm"]
    # DRAM traffic of the same kernel from the committed ncu pass (dram__bytes_read.sum + dram__bytes_write.sum, average
    # per launch over every GEMM launch of a step, cold caches under ncu): profiles/ncu_traffic_r02b.json
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic_r02b.json")
    if args.config == 2 and os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f)["gemm"]["dram_bytes_per_launch"]
        traffic_src = "profiles/ncu_traffic_r02b.json (ncu, cold cache, mean over the GEMM launches of one step)"
    algo_bytes_per_launch = (t["bytes"] + h["bytes"]) / max(n_gemm, 1)
    roofline = {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05+TMA; linear layers and implicit-GEMM convolutions)",
                "achieved": achieved, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["bf16_tflops_sustained"], "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": algo_bytes_per_launch, "peak_source": pk["source"],
                "note": "all launches of the kernel; it is HBM-bound on the short-K convolution shapes, split below",
                "launches_per_step": n_gemm, "gemm_ms_per_step": gemm_ms,
                "gemm_share_of_step": gemm_ms / ms_step,
                "tensor_bound_launches": {"launches": t["launches"], "ms": t["ms"],
                                          "achieved_tflops": t["flops"] / (t["ms"] * 1e-3) / 1e12 if t["ms"] else None,
                                          "frac_of_tensor_peak": (t["flops"] / (t["ms"] * 1e-3) / 1e12
                                                                  / pk["bf16_tflops_sustained"]) if t["ms"] else None},
                "hbm_bound_launches": {"launches": h["launches"], "ms": h["ms"],
                                       "achieved_gbs": h["bytes"] / (h["ms"] * 1e-3) / 1e9 if h["ms"] else None,
                                       "frac_of_hbm_peak": (h["bytes"] / (h["ms"] * 1e-3) / 1e9 / pk["hbm_gbs"])
                                       if h["ms"] else None, "ridge_flop_per_byte": ridge,
                                       "bytes_written_share": h["bytes_written"] / h["bytes"] if h["bytes"] else None},
                "whole_step_tflops": value / world * wl["gflop"] / 1e3,
                "whole_step_frac_of_peak": value / world * wl["gflop"] / 1e3 / pk["bf16_tflops_sustained"]}

    e2e = None
    if not args.no_e2e:
        workers, worker_probe = pick_e2e_workers()
        primed = make_e2e_loader(workers, e2e_warm + args.steps)
        e2e_train(primed, e2e_warm)
        e2e_log.clear()
        ms_e = timed(lambda: e2e_train(primed, args.steps), 1) / args.steps
        assert len(e2e_log) == args.steps, "every timed step's loss / correct count must have reached the host"
        h2d_e = B * 224 * 224 * 3 + sum(host[k].numel() * host[k].element_size() for k in ("text", "text_mask", "label"))
        e2e = {"value": world * B / (ms_e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d_e * world,
               "d2h_bytes_per_step": 8 * world, "ms_per_step": ms_e,
               "path": "b200mm.train(model, DataLoader(pin_memory=True, num_workers=%d), CrossEntropyLoss, FusedAdam, "
                       "device, graph=GraphedTrainStep): uint8 HWC pixels + token ids from pinned memory, ToTensor/Normalize on the GPU copy "
                       "stream" % workers,
               "readback": "loss + correct count of every step, asynchronous to pinned memory, consumed one step later",
               "loader_workers_probe_ms": worker_probe, "last_loss": e2e_log[-1]}
        del primed

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.config == 2:
        from oracle import reference_model as R      # cpu_baseline leg: the one place this arm touches the oracle
        r = R.cpu_train_throughput(batch=args.cpu_batch, seq_len=S, steps=3, warmup=1)
        cpu = {"value": r["samples_per_s"], "unit": "samples/s", "cores": r["cores"], "kind": "port",
               "sample": f"3 steps of batch {args.cpu_batch} (BASELINE config 1) after 1 warm-up, fp32 torch CPU, "
                         f"median {r['median_s']:.2f} s/step"}
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": wl["name"], "per_gpu_batch": B, "global_batch": B * world, "seq_len": S, "image": "3x224x224",
                       "parallelism": f"dp{world}", "l2": "inputs+activations per step >> 126 MB L2 (no flush needed)",
                       "dropout": "on (0.1 / 0.1 / 0.3, Philox)",
                       "launch": "eager" if gstep is None else "cuda_graph (one replay per step, NCCL all-reduces included; "
                                                                 "the image tower is a parallel branch of the graph)"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "grad_sync": sync_info,
        }), flush=True)
    if world > 1:
        _shutdown(dist, gstep)


if __name__ == "__main__":
    a = parse()
    a.batch = a.batch or WORKLOADS[a.config]["batch"]
    a.seq = a.seq or WORKLOADS[a.config]["seq"]
    if a.impl == "reference":
        run_reference(a)
    else:
        run_engine(a)
