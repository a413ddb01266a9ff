"""The whole train step as ONE CUDA graph (small-batch regime).

The reference trains with batch 16 (example_scripts/Multimodal_example_task2C.txt:18, 142-143; HEAD script :73) and
its loop body is ``zero_grad -> forward -> loss -> backward -> optimizer.step`` (.txt:204-217).  At that batch the
engine's ~500 kernels per step are a few microseconds each: the step is bound by the host launching them, not by the
GPU.  ``GraphedTrainStep`` captures the loop body once (``torch.cuda.graph`` around the engine's own launches -- every
kernel goes to torch's current stream, activations come from the graph's private memory pool, the TMA descriptors are
by-value launch parameters pointing into that pool) and replays it with one ``cudaGraphLaunch`` per step.

What a capture freezes, and where the per-step values live instead:

  * dropout seeds (host: hash of (seed, step, layer, site), passed by value)  -> the kernels add a device-resident
    salt (csrc/device_utils.cuh ``step_seed``) that the graph's last node advances; forward and backward of one
    replay see the same salt, consecutive replays different ones;
  * Adam's step count and learning rate (bias corrections were computed on the host)  -> ``b200mm_adam_step_dyn``
    reads both from device memory; ``FusedAdam.sync_lr()`` pushes a scheduler's new rates before the replay;
  * the batch  -> copied into static input buffers on the replay stream.

Shapes are static: one graph per (batch, sequence length); a different shape (the last, short batch of an epoch)
falls back to the eager step.  Under data parallelism (nccl) the gradient all-reduces are captured with the step: they
are launches on NCCL's stream, forked from and joined to the capturing stream by events.
"""
from __future__ import annotations

import torch

from . import ops
from .optim import FusedAdam


class GraphedTrainStep:
    """``step = GraphedTrainStep(model, optimizer, criterion)``; ``logits, loss, ok = step(text, image, mask, labels)``
    == ``optimizer.zero_grad(); model.train_step_fused(...); optimizer.step()`` (+ ``scheduler.step()`` by the caller).
    The returned tensors are the graph's static outputs: read (or copy) them before the next call."""

    def __init__(self, model, optimizer, criterion=None, warmup: int = 2):
        if not isinstance(optimizer, FusedAdam):
            raise TypeError("GraphedTrainStep needs b200mm.FusedAdam (its step is the capturable one)")
        sync = getattr(model, "grad_sync", None)
        if sync is not None and sync.world > 1:
            # data parallel: the phase all-reduces are NCCL launches on NCCL's stream, forked from / joined to the
            # capturing stream by events -- all of it becomes part of the graph.  Needs the nccl backend (gloo has no
            # stream semantics) and no timing events inside the captured region.
            import torch.distributed as dist
            if dist.get_backend(sync.group) != "nccl":
                raise RuntimeError("GraphedTrainStep under data parallelism needs the nccl backend")
            sync._ev = None
        self.model, self.optimizer, self.warmup = model, optimizer, warmup
        self.loss_kw = {}
        if criterion is not None:
            self.loss_kw = dict(loss_kind=criterion.loss_kind, alpha=criterion.alpha, gamma=criterion.gamma)
        self.device = model.device
        self.salt = torch.zeros(1, device=self.device, dtype=torch.int64)
        self.graph = None
        self.key = None
        self.static_in = None
        self.static_out = None
        self.replays = 0
        self.launches_per_replay = 0

    # ------------------------------------------------------------------ the loop body (eager)
    def _body(self, text, image, mask, labels):
        self.optimizer.zero_grad()
        out = self.model.train_step_fused(text, image, mask, labels, **self.loss_kw)
        self.optimizer.step()
        return out

    def _capture(self, text, image, mask, labels):
        model, opt = self.model, self.optimizer
        model.train()
        self.static_in = tuple(t.clone() for t in (text, image, mask, labels))
        # the warm-up steps below are real optimizer steps: snapshot everything they touch and put it back, so that
        # the first replay is the run's first step (drop-in behaviour of the eager loop)
        st = model.store
        bufs = [b for b in (getattr(getattr(model, "img", None), "buffers", None),) if b is not None]
        snap = [t.clone() for t in (st.master, opt.exp_avg, opt.exp_avg_sq, *bufs)]
        host = (opt._step, model._step, getattr(getattr(model, "img", None), "num_batches_tracked", 0))
        # warm-up on a side stream (torch's capture rule): first-use configuration (cudaFuncSetAttribute, tensor-map
        # entry points, lazily created buffers) must happen outside the capture
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self._body(*self.static_in)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        opt.make_capturable()
        ops.set_step_salt(self.salt)
        self.graph = torch.cuda.CUDAGraph()
        from . import _lib
        before = _lib.LAUNCHES[0]
        with torch.cuda.graph(self.graph):
            self.static_out = self._body(*self.static_in)
            ops.step_advance(self.salt, opt._step_dev)
        self.launches_per_replay = _lib.LAUNCHES[0] - before      # kernel nodes of ours in the graph
        # the capture only recorded work; the warm-up steps did run: restore parameters, moments, BatchNorm buffers
        # and the host / device step counters
        for dst, src in zip((st.master, opt.exp_avg, opt.exp_avg_sq, *bufs), snap):
            dst.copy_(src)
        st.refresh_shadow()
        model._shadow_fresh = True
        opt._step, model._step = host[0], host[1]
        if hasattr(getattr(model, "img", None), "num_batches_tracked"):
            model.img.num_batches_tracked = host[2]
        opt._step_dev.fill_(opt._step + 1)
        self.salt.zero_()
        self.key = self._key(text, image, mask, labels)

    @staticmethod
    def _key(text, image, mask, labels):
        return (tuple(text.shape), tuple(image.shape), image.dtype, tuple(labels.shape), labels.dtype)

    # ------------------------------------------------------------------ public
    def __call__(self, text, image, mask, labels):
        if self.graph is None:
            self._capture(text, image, mask, labels)
        if self._key(text, image, mask, labels) != self.key:
            return self.eager(text, image, mask, labels)
        for dst, src in zip(self.static_in, (text, image, mask, labels)):
            dst.copy_(src, non_blocking=True)
        self.optimizer.sync_lr()
        self.graph.replay()
        self.optimizer._step += 1
        self.replays += 1
        return self.static_out

    def eager(self, text, image, mask, labels):
        """The same step without the graph (odd-shaped batches).  Keeps the device-side counters in step."""
        opt = self.optimizer
        if opt._step_dev is not None:
            opt.sync_lr()
        out = self._body(text, image, mask, labels)
        if opt._step_dev is not None:
            ops.step_advance(self.salt, opt._step_dev)
        return out

    def close(self):
        """Detach the salt from the kernels (eager steps then use their seeds as passed)."""
        ops.set_step_salt(None)
        self.graph = None
