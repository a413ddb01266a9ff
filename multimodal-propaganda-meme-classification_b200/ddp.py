"""Data-parallel gradient synchronisation for the one-process-per-GPU launch (new work: the reference is
single-process).

Pure data parallelism: every rank holds the full parameter store and processes its own shard of samples; the
only exchange is one gradient all-reduce per step.  Because parameters, gradients and optimizer state are flat,
contiguous buffers (params.ParamStore), a "bucket" is simply an element range of the flat gradient buffer -- no
flatten / unflatten copies.  Overlap with backward comes from the order in which the engine finishes gradients:

    head -> text tower, top layer group first ... embeddings last -> image tower, last stage first ... stem last

Every tower announces (``grad_phases``) the parameter groups whose gradients become final at successive points of
its backward and calls back as each point is passed, so each group's all-reduce is launched on NCCL's stream while
the rest of the backward is still running; only the very last group (stem / patch embedding + the small norm / bias
parameters) is an exposed tail.

Payload: bf16 by default (SURVEY.md §8e "bf16 payload -> fp32 master update").  A ready range is packed into a
persistent bf16 buffer, pre-scaled by 1 / world size (so the partial sums stay in range), all-reduced there, and the
optimizer reads the averaged gradient straight from that buffer (``b200mm_adam_step_g16``): half the NVLink bytes of
the fp32 exchange, no un-pack pass.  ``payload="fp32"`` all-reduces the fp32 gradient buffer in place (the 1 / world
size averaging is then folded into the Adam kernel's ``grad_scale``).  BatchNorm statistics stay per replica, exactly
as N independent runs of the reference would behave (no SyncBN in the reference).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def contiguous_ranges(specs, names, predicate, align=64):
    """Merge the (padded) element ranges of all parameters whose name satisfies ``predicate``."""
    spans = []
    for n in names:
        if predicate(n):
            s = specs[n]
            spans.append([s.offset, s.offset + (s.numel + align - 1) // align * align])
    spans.sort()
    merged = []
    for a, b in spans:
        if merged and merged[-1][1] == a:
            merged[-1][1] = b
        else:
            merged.append([a, b])
    return [tuple(m) for m in merged]


def split_ranges(ranges, max_elems):
    out = []
    for a, b in ranges:
        while b - a > max_elems:
            out.append((a, a + max_elems))
            a += max_elems
        if b > a:
            out.append((a, b))
    return out


def default_phases():
    """Two phases (round-1 plan): the text tower, then everything else."""
    return [("text", lambda n: n.startswith("bert.")), ("rest", lambda n: True)]


class GradSync:
    """All-reduce of the flat gradient buffer in phases that follow the backward's completion order.

    ``phases``: ordered ``[(tag, predicate)]``; a parameter belongs to the FIRST phase whose predicate accepts its
    name (so a catch-all ``lambda n: True`` at the end covers whatever the earlier ones left)."""

    def __init__(self, store, group=None, bucket_elems: int = 64 * 1024 * 1024, phases=None, payload: str = "bf16",
                 pack=None, phase_predicates=None):
        self.store = store
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if payload not in ("bf16", "fp32"):
            raise ValueError(f"payload must be 'bf16' or 'fp32', not {payload!r}")
        self.payload = payload
        names = store.names()
        if phase_predicates is not None:            # round-1 signature: dict tag -> predicate (order = dict order)
            phases = list(phase_predicates.items())
        if phases is None:
            phases = default_phases()
        claimed = set()
        self.phases = {}
        for tag, pred in phases:
            mine = [n for n in names if n not in claimed and pred(n)]
            claimed.update(mine)
            mine_set = set(mine)
            self.phases[tag] = split_ranges(contiguous_ranges(store.specs, names, mine_set.__contains__), bucket_elems)
        if len(claimed) != len(names):
            missing = [n for n in names if n not in claimed]
            raise ValueError(f"gradient phases do not cover {missing[:4]} ... ({len(missing)} parameters)")
        self._pending = []
        self._launched = set()
        self.comm = None          # bf16 payload buffer, same element offsets as store.grad
        if payload == "bf16" and self.world > 1:
            self.comm = torch.zeros(store.numel, device=store.device, dtype=torch.bfloat16)
        if pack is None:
            from . import ops
            pack = ops.cast_to_bf16
        self._pack = pack
        # exposed tail of the exchange: device time between "the backward's last kernel" and "every all-reduce done",
        # from two events recorded on the compute stream around the waits in finish() (read with exposed_wait_ms())
        self._ev = None
        if self.world > 1 and torch.device(store.device).type == "cuda":
            self._ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        self.bytes_per_step = sum(b - a for rs in self.phases.values() for a, b in rs) * (2 if payload == "bf16" else 4)

    def covered(self):
        return sorted(r for rs in self.phases.values() for r in rs)

    def grad_buffer(self):
        """The buffer holding the rank-AVERAGED (bf16) or rank-SUMMED (fp32) gradients after ``finish``."""
        return self.comm if self.comm is not None else self.store.grad

    def ready(self, phase: str):
        """Called by the engine as soon as every gradient of ``phase`` is final: launch its all-reduces."""
        if self.world == 1 or phase in self._launched:
            return
        self._launched.add(phase)
        for a, b in self.phases.get(phase, ()):
            if self.comm is not None:
                self._pack(self.store.grad[a:b], self.comm[a:b], 1.0 / self.world)
                buf = self.comm[a:b]
            else:
                buf = self.store.grad[a:b]
            self._pending.append(dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self) -> float:
        """Launch whatever phase the backward did not announce, block the current stream on the outstanding
        all-reduces, and return the scale that turns ``grad_buffer()`` into the mean gradient (1 for the pre-scaled
        bf16 payload, 1 / world for the summed fp32 one) -- to be folded into the optimizer step."""
        if self.world == 1:
            return 1.0
        for tag in self.phases:
            self.ready(tag)
        if self._ev is not None:
            self._ev[0].record()
        for w in self._pending:
            w.wait()
        if self._ev is not None:
            self._ev[1].record()
        self._pending.clear()
        self._launched.clear()
        return 1.0 if self.comm is not None else 1.0 / self.world

    def exposed_wait_ms(self) -> float:
        """Device time the compute stream stalled in the last ``finish()`` (synchronises; for the benchmark record)."""
        if self._ev is None:
            return 0.0
        self._ev[1].synchronize()
        return self._ev[0].elapsed_time(self._ev[1])

    def finish_into_grad(self):
        """For optimizers that read ``param.grad`` (torch.optim.*): finish, then leave the MEAN gradient in the fp32
        gradient buffer the ``.grad`` views alias."""
        if self.world == 1:
            return
        scale = self.finish()
        if self.comm is not None:
            self.store.grad.copy_(self.comm)
        elif scale != 1.0:
            self.store.grad.mul_(scale)

    def broadcast_parameters(self, buffers=()):
        if self.world == 1:
            return
        dist.broadcast(self.store.master, src=0, group=self.group)
        for b in buffers:
            dist.broadcast(b, src=0, group=self.group)
