"""Fused training-step tail: clip_grad_norm_ + SGD(momentum, weight decay) + the
linear-decay learning-rate schedule of var_updown/scripts/train.py:126-134,173-176, as three kernel
launches over all parameter tensors instead of ~100 eager launches."""
import ctypes as C
from typing import Iterable

import torch

from . import _lib


class FusedClipSGD:
    """`step()` == clip_grad_norm_(params, max_norm); SGD.step(); LambdaLR.step() of the reference loop.

    Parameters that currently have no gradient are left untouched, as torch's SGD does for
    `p.grad is None` (frozen embedding; decoder LSTM under the freeze schedule, train.py:156-161).

    `legacy_zero_grad=True` reproduces the reference's PINNED torch 1.1.0 (requirements.txt:12) instead:
    there `optimizer.zero_grad()` leaves zero tensors, not None, so a parameter that has had a gradient once
    and is then frozen keeps being updated with a zero gradient (weight decay + momentum coasting). With it,
    `zero_grad()` zeroes nothing on the device: such parameters are handed to the kernel with a NULL gradient."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr=0.015, momentum=0.9, weight_decay=0.001,
                 max_norm=12.5, num_iterations=70000, legacy_zero_grad=False):
        self.params = [p for p in params]
        self.base_lr, self.momentum, self.weight_decay, self.max_norm = lr, momentum, weight_decay, max_norm
        self.num_iterations = num_iterations
        self.legacy_zero_grad = legacy_zero_grad
        self.iteration = 0
        self._mom = {}
        self._partial = None

    @property
    def lr(self) -> float:
        return self.base_lr * (1 - self.iteration / self.num_iterations)      # train.py:132-134

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0):
        """`grad_scale` multiplies every gradient before clipping and the update. Data-parallel training passes
        1 / world_size together with `BucketedGradReducer.reduce(..., average=False)`: the all-reduce leaves the SUM in
        the buckets and no extra pass divides them."""
        L = _lib.lib()
        ps = [p for p in self.params if p.grad is not None or (self.legacy_zero_grad and id(p) in self._mom)]
        if not ps:
            return
        dev = ps[0].device
        if not ps[0].is_cuda:
            raise RuntimeError("FusedClipSGD runs only on CUDA parameters; there is no CPU fallback")
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        grads = [None if p.grad is None else p.grad.contiguous() for p in ps]
        firsts = []
        for p in ps:
            first = id(p) not in self._mom
            if first:
                self._mom[id(p)] = torch.empty_like(p)
            firsts.append(int(first))
        # all tensors in three launches: squared-norm partials over a fixed chunking, their ordered sum, the update
        if len(ps) > 32:
            raise RuntimeError("FusedClipSGD handles at most 32 parameter tensors (the captioner has 21)")
        n = len(ps)
        chunks = sum((p.numel() + 16383) // 16384 for p in ps)
        if self._partial is None or self._partial.numel() < chunks + 1 or self._partial.device != dev:
            self._partial = torch.empty(chunks + 1, dtype=torch.float32, device=dev)
        vp = C.c_void_p
        _lib.check(L.sscvae_sgd_step_multi(
            n, (vp * n)(*[p.data_ptr() for p in ps]), (vp * n)(*[None if g is None else g.data_ptr() for g in grads]),
            (vp * n)(*[self._mom[id(p)].data_ptr() for p in ps]), (C.c_uint64 * n)(*[p.numel() for p in ps]),
            (C.c_int32 * n)(*firsts), float(self.max_norm), float(self.lr), float(self.momentum), float(self.weight_decay),
            float(grad_scale),
            _lib.ptr(self._partial), self._partial.numel(), stream))
        for p in ps:
            # the kernel wrote through raw pointers: tell autograd (and the captioner's packed-weight cache, which is
            # keyed on the version counters) that the parameters changed
            torch.autograd.graph.increment_version(p)
        self.iteration += 1
