"""``FusedAdam``: torch.optim.Adam semantics (reference: optim.Adam(model.parameters(), lr=2e-5),
example_scripts/Multimodal_example_task2C.txt:249; HEAD script param groups + clip_grad_norm_,
Multimodal_example_task2C.py:168, :645-664, :713-715) as ONE streaming kernel per contiguous parameter range:
read p, g, m, v (fp32), write p, m, v and the bf16 shadow the GEMMs consume -- 30 B / parameter of HBM traffic,
nothing else.  Optional global-norm clipping is folded in (a sum-of-squares pre-pass, then the clip factor is
applied inside the Adam kernel; no separate scaling pass over the gradients).
"""
from __future__ import annotations

import torch

from . import ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=None):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.max_grad_norm = max_grad_norm
        stores = {id(getattr(p, "_b200mm_store", None)): getattr(p, "_b200mm_store", None)
                  for g in self.param_groups for p in g["params"]}
        if len(stores) != 1 or None in stores.values():
            raise ValueError("FusedAdam drives the parameters of exactly one b200mm model")
        self.store = next(iter(stores.values()))
        self.model = self.store.owner
        # data parallel: this optimizer finishes the gradient all-reduce itself (inside step(), so the exchange overlaps
        # whatever the host does between backward and step); without this flag the model finishes it at the end of
        # backward so that ANY optimizer reading param.grad sees averaged gradients (ddp.GradSync.finish_into_grad)
        self.model._defer_grad_sync = True
        n = self.store.numel
        self.exp_avg = torch.zeros(n, device=self.store.device, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(n, device=self.store.device, dtype=torch.float32)
        self._gradsq = torch.zeros(1, device=self.store.device, dtype=torch.float32)
        self._step = 0
        self.last_grad_norm = None
        # capturable mode (graph.GraphedTrainStep): step count and per-group learning rates in device memory
        self._step_dev = None
        self._lr_dev = None
        self._lr_host = None
        # contiguous element ranges per param group (alignment padding between parameters is all-zero: safe to sweep)
        self._ranges = []
        for gi, g in enumerate(self.param_groups):
            spans = sorted((p._b200mm_offset, p._b200mm_offset + p._b200mm_padded) for p in g["params"])
            merged = []
            for a, b in spans:
                if merged and merged[-1][1] == a:
                    merged[-1][1] = b
                else:
                    merged.append([a, b])
            self._ranges.append(merged)

    def zero_grad(self, set_to_none: bool = False):
        self.model.zero_grad()

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        assert closure is None
        st = self.store
        self._step += 1
        sync = getattr(self.model, "grad_sync", None)
        grad = st.grad
        if sync is not None and sync.world > 1:
            grad_scale = grad_scale * sync.finish()     # fp32 payload: summed -> mean, folded into the Adam kernel
            grad = sync.grad_buffer()                   # bf16 payload: the averaged gradients live in the comm buffer
        gradsq = None
        if self.max_grad_norm is not None:
            self._gradsq.zero_()
            ops.sumsq(grad, self._gradsq)
            gradsq = self._gradsq
            # device scalar: SQUARED norm of the mean gradient (.sqrt().item() when logging)
            self.last_grad_norm = self._gradsq if grad_scale == 1.0 else self._gradsq * (grad_scale * grad_scale)
        s0 = st.shadow_start
        for gi, (g, ranges) in enumerate(zip(self.param_groups, self._ranges)):
            b1, b2 = g["betas"]
            for a, b in ranges:
                # split at the shadow boundary: embedding tables / norm params have no bf16 shadow
                for lo, hi, sh in ((a, min(b, s0), False), (max(a, s0), b, True)):
                    if hi <= lo:
                        continue
                    if self._step_dev is not None:
                        ops.adam_step_dyn(st.master[lo:hi], grad[lo:hi], self.exp_avg[lo:hi], self.exp_avg_sq[lo:hi],
                                          st.shadow[lo:hi] if sh else None, lr_dev=self._lr_dev[gi:gi + 1],
                                          step_dev=self._step_dev, beta1=b1, beta2=b2, eps=g["eps"],
                                          weight_decay=g["weight_decay"], gradsq=gradsq,
                                          max_norm=self.max_grad_norm or 0.0, grad_scale=grad_scale)
                        continue
                    ops.adam_step(st.master[lo:hi], grad[lo:hi], self.exp_avg[lo:hi], self.exp_avg_sq[lo:hi],
                                  st.shadow[lo:hi] if sh else None, lr=g["lr"], beta1=b1, beta2=b2, eps=g["eps"],
                                  weight_decay=g["weight_decay"], step=self._step, gradsq=gradsq,
                                  max_norm=self.max_grad_norm or 0.0, grad_scale=grad_scale)
        self.model._shadow_fresh = True

    # ---- CUDA-graph support (graph.GraphedTrainStep)
    def make_capturable(self):
        """Move the step count and the learning rates to device memory (read by b200mm_adam_step_dyn): a captured
        ``step()`` then serves every replay.  ``_step_dev`` holds the count the NEXT step uses; the graph's last node
        (ops.step_advance) increments it, ``sync_lr()`` pushes the param groups' current learning rates."""
        dev = self.store.device
        self._step_dev = torch.full((1,), self._step + 1, device=dev, dtype=torch.int32)
        self._lr_host = [float(g["lr"]) for g in self.param_groups]
        self._lr_dev = torch.tensor(self._lr_host, device=dev, dtype=torch.float32)
        return self

    def sync_lr(self):
        """Push learning rates a scheduler changed since the last call (one fill kernel per changed group: the value
        travels in the launch, so nothing on the host must outlive the call)."""
        for gi, g in enumerate(self.param_groups):
            lr = float(g["lr"])
            if lr != self._lr_host[gi]:
                self._lr_dev[gi:gi + 1].fill_(lr)
                self._lr_host[gi] = lr

    # ---- checkpoint / resume: the moments and the step count live outside Optimizer.state (flat buffers)
    def state_dict(self):
        sd = super().state_dict()
        sd["b200mm"] = {"exp_avg": self.exp_avg.detach().cpu(), "exp_avg_sq": self.exp_avg_sq.detach().cpu(),
                        "step": self._step, "max_grad_norm": self.max_grad_norm}
        return sd

    def load_state_dict(self, state_dict):
        extra = state_dict.get("b200mm")
        super().load_state_dict({k: v for k, v in state_dict.items() if k != "b200mm"})
        if extra is not None:
            self.exp_avg.copy_(extra["exp_avg"])
            self.exp_avg_sq.copy_(extra["exp_avg_sq"])
            self._step = int(extra["step"])
            self.max_grad_norm = extra.get("max_grad_norm", self.max_grad_norm)


def get_linear_schedule_with_warmup(optimizer, num_warmup_steps, num_training_steps, last_epoch=-1):
    """Same schedule as transformers.get_linear_schedule_with_warmup (used at Multimodal_example_task2C.py:172-174)."""
    def lr_lambda(step):
        if step < num_warmup_steps:
            return float(step) / float(max(1, num_warmup_steps))
        return max(0.0, float(num_training_steps - step) / float(max(1, num_training_steps - num_warmup_steps)))
    return torch.optim.lr_scheduler.LambdaLR(optimizer, lr_lambda, last_epoch)
