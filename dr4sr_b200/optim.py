"""Fused dense Adam over flat parameter buffers (reference: torch.optim.Adam at model/basemodel.py:85-86).

One kernel pass per buffer (p, g, m, v read; p, m, v written; optionally g cleared in the same pass),
identical arithmetic to torch's single-tensor Adam.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch

from . import engine as _engine


class FlatGroup:
    """A contiguous fp32 parameter buffer with its gradient and Adam state."""

    def __init__(self, param: torch.Tensor, grad: torch.Tensor, name: str, zero_grad_in_step: bool = False) -> None:
        self.name = name
        self.param, self.grad = param, grad
        self.m = torch.zeros_like(param)
        self.v = torch.zeros_like(param)
        self.zero_grad_in_step = zero_grad_in_step
        self.dirty = False          # grad buffer holds an un-consumed gradient


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Optimizer façade (param_groups / zero_grad / state_dict shape) over FlatGroups."""

    def __init__(self, params, groups: List[FlatGroup], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, on_step: Optional[Callable[[], None]] = None) -> None:
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.flat_groups = groups
        self.step_count = 0
        self._on_step = on_step
        self.pending_backwards = 0      # backward() calls since the last step() / zero_grad() (see BaseModel: at most one)

    def zero_grad(self, set_to_none: bool = True) -> None:
        """No work to do, by construction: every backward OVERWRITES the flat encoder gradient, and the table gradient
        accumulator is cleared by the Adam pass that consumes it (zero_grad_in_step), so nothing can carry over from the
        previous step.  `.grad` keeps pointing at those persistent buffers (torch's loop over ~30 parameters setting
        them to None costs 25 us of host time per step and would be re-done by the next backward anyway).
        Unlike torch, gradients do NOT accumulate across backward() calls: a second backward() without a step() or
        zero_grad() in between raises (dr4sr_b200/model/basemodel.py::_TrainStep.backward) instead of silently overwriting."""
        self.pending_backwards = 0
        return None

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self.pending_backwards = 0
        self.step_count += 1
        h = self.param_groups[0]
        for g in self.flat_groups:
            _engine.adam_step(g.param, g.grad, g.m, g.v, self.step_count, h['lr'], h['betas'][0], h['betas'][1], h['eps'],
                              h['weight_decay'], zero_grad=g.zero_grad_in_step)
            if g.zero_grad_in_step:
                g.dirty = False
        if self._on_step is not None:
            self._on_step()
        return loss
