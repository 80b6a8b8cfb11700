"""MSE loss and the Adam update on the device (train_model.py:636-637,421,424)."""
from __future__ import annotations

from ctypes import c_float, c_int
from typing import Iterable, List

import torch

from . import _lib
from ._lib import ADAM_CHUNK, call, ptr, stream


def mse_loss(outputs: torch.Tensor, targets: torch.Tensor, want_grad: bool = False):
    """torch.nn.MSELoss() (mean) on device; returns loss[1] (and dL/d outputs when ``want_grad``)."""
    _lib.require_device(outputs, torch.float32, "outputs")
    _lib.require_device(targets, torch.float32, "targets")
    if outputs.numel() != targets.numel():
        raise _lib.CtkError("outputs and targets differ in size")
    n = outputs.numel()
    loss = torch.empty(1, device=outputs.device, dtype=torch.float32)
    grad = torch.empty_like(outputs) if want_grad else None
    call("ctk_mse_loss", ptr(outputs), ptr(targets), c_int(n), ptr(loss), ptr(grad), stream())
    return (loss, grad) if want_grad else loss


class _MseFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, outputs, targets):
        loss, grad = mse_loss(outputs.contiguous(), targets.contiguous(), want_grad=True)
        ctx.save_for_backward(grad)
        ctx.shape = outputs.shape
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dloss):
        (grad,) = ctx.saved_tensors
        dloss = dloss.reshape(1).contiguous().float()
        _lib.require_device(dloss, torch.float32, "loss gradient")
        out = torch.empty_like(grad)
        call("ctk_scale_by_scalar", ptr(grad), ptr(dloss), c_int(grad.numel()), ptr(out), stream())
        return out.view(ctx.shape), None


class MSELoss(torch.nn.Module):
    """``torch.nn.MSELoss()`` (reduction='mean', train_model.py:636) as ONE libctk launch: the loss value and
    dL/d outputs come out of the same kernel, the backward just hands the stored gradient on (scaled by the incoming
    gradient, which ``loss.backward()`` seeds with 1)."""

    def forward(self, outputs: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        return _MseFunction.apply(outputs, targets)


class Adam(torch.optim.Optimizer):
    """Drop-in for ``optim.Adam(params, lr, weight_decay=1e-4)`` (train_model.py:637): coupled L2, betas
    (0.9, 0.999), eps 1e-8 -- one multi-tensor kernel launch per step over every parameter of the group.

    ``state_dict()`` / ``param_groups[0]['lr']`` behave like torch.optim.Adam's (state keys ``step``,
    ``exp_avg``, ``exp_avg_sq``), so LR schedulers and checkpoints keep working.
    """

    def __init__(self, params: Iterable, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._tables = {}

    def _table(self, gi: int, plist: List[torch.Tensor]):
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr()) for p in plist)
        tab = self._tables.get(gi)
        if tab is not None and tab["key"] == key:
            return tab
        dev = plist[0].device
        i64 = lambda xs: torch.tensor(xs, dtype=torch.int64, device=dev)      # noqa: E731
        bt, bc = [], []
        for ti, p in enumerate(plist):
            for c in range((p.numel() + ADAM_CHUNK - 1) // ADAM_CHUNK):
                bt.append(ti)
                bc.append(c)
        tab = dict(key=key,
                   p=i64([p.data_ptr() for p in plist]), g=i64([p.grad.data_ptr() for p in plist]),
                   m=i64([self.state[p]["exp_avg"].data_ptr() for p in plist]),
                   v=i64([self.state[p]["exp_avg_sq"].data_ptr() for p in plist]),
                   n=i64([p.numel() for p in plist]),
                   bt=torch.tensor(bt, dtype=torch.int32, device=dev), bc=torch.tensor(bc, dtype=torch.int32, device=dev),
                   blocks=len(bt))
        self._tables[gi] = tab
        return tab

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            for p in plist:
                _lib.require_device(p, torch.float32, "parameter")
                _lib.require_device(p.grad, torch.float32, "gradient")
                st = self.state[p]
                if not st:
                    st["step"] = torch.zeros((), dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            steps = {int(self.state[p]["step"]) for p in plist}
            if len(steps) != 1:
                raise _lib.CtkError("parameters of one group must share a step count")
            t = steps.pop() + 1
            tab = self._table(gi, plist)
            b1, b2 = group["betas"]
            call("ctk_adam_multi", ptr(tab["p"]), ptr(tab["g"]), ptr(tab["m"]), ptr(tab["v"]), ptr(tab["n"]),
                 ptr(tab["bt"]), ptr(tab["bc"]), c_int(tab["blocks"]), c_float(group["lr"]), c_float(b1), c_float(b2),
                 c_float(group["eps"]), c_float(group["weight_decay"]), c_int(t), c_float(grad_scale), stream())
            for p in plist:
                self.state[p]["step"] += 1
            # the kernel updated the parameters through raw pointers: bump their version counters so that derived caches
            # keyed on them (InferenceEngine's packed weights / folded BatchNorm) are rebuilt at the next eval forward
            torch.autograd.graph.increment_version(plist)
        return loss
