"""Fused training-step tail: clip_grad_norm_ + SGD(momentum, weight decay) + the
linear-decay learning-rate schedule of var_updown/scripts/train.py:126-134,173-176, as three kernel
launches over flat buffers instead of ~100 eager launches."""
import ctypes as C
from typing import Iterable

import torch

from . import _lib


class FusedClipSGD:
    """`step()` == clip_grad_norm_(params, max_norm); SGD.step(); LambdaLR.step() of the reference loop.

    Parameters that currently have no gradient are left untouched, as torch's SGD does for
    `p.grad is None` (frozen embedding; decoder LSTM under the freeze schedule, train.py:156-161)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr=0.015, momentum=0.9, weight_decay=0.001,
                 max_norm=12.5, num_iterations=70000):
        self.params = [p for p in params]
        self.base_lr, self.momentum, self.weight_decay, self.max_norm = lr, momentum, weight_decay, max_norm
        self.num_iterations = num_iterations
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
    def step(self):
        L = _lib.lib()
        ps = [p for p in self.params if p.grad is not None]
        if not ps:
            return
        dev = ps[0].device
        if not ps[0].is_cuda:
            raise RuntimeError("FusedClipSGD runs only on CUDA parameters; there is no CPU fallback")
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        if self._partial is None:
            self._partial = torch.empty(2048, dtype=torch.float32, device=dev)
        sq = self._partial[1024:1025]
        # total squared norm = sum over tensors (each reduced deterministically), accumulated on device
        total = torch.zeros(1, dtype=torch.float32, device=dev)
        for p in ps:
            g = p.grad.contiguous()
            _lib.check(L.sscvae_grad_sqnorm(_lib.ptr(g), g.numel(), _lib.ptr(self._partial), _lib.ptr(sq), stream))
            total += sq
        for p in ps:
            first = id(p) not in self._mom
            if first:
                self._mom[id(p)] = torch.empty_like(p)
            g = p.grad.contiguous()
            _lib.check(L.sscvae_sgd_step(_lib.ptr(p), _lib.ptr(g), _lib.ptr(self._mom[id(p)]), p.numel(), _lib.ptr(total),
                                         float(self.max_norm), float(self.lr), float(self.momentum),
                                         float(self.weight_decay), int(first), stream))
            # the kernel wrote through a raw pointer: tell autograd (and the captioner's packed-weight cache,
            # which is keyed on the version counters) that the parameter changed
            torch.autograd.graph.increment_version(p)
        self.iteration += 1
