"""Flat parameter storage: every trainable tensor of the model is a view into ONE fp32 master buffer, with a
matching fp32 gradient buffer and a bf16 shadow the tensor-core kernels read.

Why flat: the optimizer (reference: ``optim.Adam(model.parameters(), lr=2e-5)``,
example_scripts/Multimodal_example_task2C.txt:249) becomes a single HBM-streaming kernel over 161.7 M
contiguous elements, ``zero_grad`` a single memset, and the data-parallel gradient all-reduce a handful of
contiguous NCCL buckets -- no per-tensor launches, no flatten/unflatten copies.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

from . import ops

ALIGN = 64  # elements; keeps every view 128-byte aligned in bf16 and 256-byte aligned in fp32


@dataclass
class ParamSpec:
    name: str
    shape: tuple
    offset: int
    numel: int
    shadow: bool


class ParamStore:
    def __init__(self, device):
        self.device = torch.device(device)
        self.specs: dict[str, ParamSpec] = {}
        self._order: list[str] = []
        self._size = 0
        self._no_shadow_end = 0
        self._finalized = False
        self.master = self.grad = self.shadow = None

    def add(self, name: str, shape, shadow: bool = True) -> None:
        """Register a parameter. Parameters without a bf16 shadow (embedding tables, norm scales, biases that the
        kernels read in fp32) must be added before any shadowed one."""
        assert not self._finalized and name not in self.specs
        numel = int(math.prod(shape))
        if not shadow:
            assert self._no_shadow_end == self._size, "add shadow-less parameters first"
        self.specs[name] = ParamSpec(name, tuple(shape), self._size, numel, shadow)
        self._order.append(name)
        self._size += (numel + ALIGN - 1) // ALIGN * ALIGN
        if not shadow:
            self._no_shadow_end = self._size

    def finalize(self) -> None:
        n = self._size
        self.master = torch.zeros(n, device=self.device, dtype=torch.float32)
        self.grad = torch.zeros(n, device=self.device, dtype=torch.float32)
        self.shadow = torch.zeros(n, device=self.device, dtype=torch.bfloat16)
        self._finalized = True

    # ---- views
    def _view(self, buf, name):
        s = self.specs[name]
        return buf[s.offset:s.offset + s.numel].view(s.shape)

    def p(self, name):
        return self._view(self.master, name)

    def g(self, name):
        return self._view(self.grad, name)

    def s(self, name):
        return self._view(self.shadow, name)

    def span(self, buf, first: str, last: str, shape):
        """One view covering consecutive parameters first..last (e.g. q_lin|k_lin|v_lin as a fused [3D, D])."""
        a, b = self.specs[first], self.specs[last]
        n = int(math.prod(shape))
        assert b.offset + b.numel - a.offset == n, "parameters are not densely consecutive"
        return buf[a.offset:a.offset + n].view(shape)

    def names(self):
        return list(self._order)

    @property
    def numel(self):
        return self._size

    @property
    def shadow_start(self):
        return self._no_shadow_end

    # ---- whole-buffer operations
    def refresh_shadow(self):
        n0 = self._no_shadow_end
        if self._size > n0:
            ops.cast_to_bf16(self.master[n0:], self.shadow[n0:])

    def zero_grad(self):
        self.grad.zero_()
