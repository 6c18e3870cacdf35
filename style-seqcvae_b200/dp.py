"""Data-parallel training wrapper: one process per GPU, bucketed gradient all-reduce overlapped
with the tail of backward.

The reference's only multi-GPU mechanism is `nn.DataParallel` (var_updown/scripts/train.py:123-124):
single process, parameters re-broadcast every iteration, gradients reduced to GPU 0. Here every rank
holds a replica, runs the same kernels on its shard of the batch, and the gradients are summed with
`torch.distributed.all_reduce` (NCCL over NVLink on GPUs; gloo in the CPU tests) in SSCVAE_GRAD_GROUPS
buckets. `sscvae_train_backward` records one CUDA event per bucket as soon as that bucket's
gradients are final (head first, attention last), so a side stream can reduce bucket g while the
weight-gradient GEMMs of bucket g+1 are still running.
"""
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

# parameter-name prefixes of each gradient group, in the order backward finalises them
GROUP_PREFIXES = [
    ("_output_projection", "_output_layer"),                                       # 0 head
    ("_updown_cell._language_lstm_cell_decoder",),                                 # 1 decoder LSTM
    ("_updown_cell._language_lstm_cell_encoder", "_updown_cell.fc_"),              # 2 encoder LSTM + latent heads
    ("_updown_cell._attention_lstm_cell", "_embedding_layer"),                     # 3 attention LSTM (+ learned embedding)
    ("_updown_cell._butd_attention",),                                             # 4 attention module
]


def group_of(name: str) -> int:
    for g, prefixes in enumerate(GROUP_PREFIXES):
        if any(name.startswith(p) for p in prefixes):
            return g
    raise KeyError(name)


def shard_batch(n_items: int, rank: int, world: int) -> slice:
    """Contiguous shard of the global batch owned by `rank` (rows are independent captions)."""
    per = (n_items + world - 1) // world
    return slice(min(rank * per, n_items), min((rank + 1) * per, n_items))


class BucketedGradReducer:
    """Averages `.grad` of the given named parameters across ranks, one all-reduce per group.

    Parameters' gradients are flattened per group into one contiguous buffer (so a group is one
    collective), reduced, divided by world size and copied back. Works on CPU tensors with the gloo
    backend (tests) and on CUDA tensors with NCCL, where `events` (one per group, recorded by the
    backward kernels) let the reduction of early groups overlap the rest of backward."""

    def __init__(self, named_params: Sequence, process_group=None):
        self.pg = process_group
        self.groups: List[List[torch.nn.Parameter]] = [[] for _ in GROUP_PREFIXES]
        seen = set()
        for name, p in named_params:
            if id(p) in seen:
                continue
            seen.add(id(p))
            self.groups[group_of(name)].append(p)
        self._flat: Dict[int, torch.Tensor] = {}
        self._side_stream = None

    def world(self) -> int:
        return dist.get_world_size(self.pg) if dist.is_initialized() else 1

    def reduce(self, events: Optional[Sequence] = None, buckets: Optional[Dict] = None, average: bool = True):
        """`buckets` = UpDownCaptioner.grad_buckets(): when every gradient of a bucket is the captioner's own view into
        the bucket's flat buffer (the normal case) the bucket is all-reduced IN PLACE: no flatten / copy-back.
        `average=False` leaves the SUM over ranks (FusedClipSGD.step(grad_scale=1/world) takes the mean on the fly and
        saves one pass over every bucket)."""
        world = self.world()
        if world == 1:
            return
        self.last_comm_done = None
        if buckets and self._reduce_in_place(buckets, world, events, average):
            return
        cuda = any(p.is_cuda for g in self.groups for p in g)
        if cuda and self._side_stream is None:
            self._side_stream = torch.cuda.Stream()
        handles = []
        for g, params in enumerate(self.groups):
            ps = [p for p in params if p.grad is not None]
            if not ps:
                continue
            n = sum(p.grad.numel() for p in ps)
            flat = self._flat.get(g)
            if flat is None or flat.numel() != n or flat.device != ps[0].grad.device:
                flat = torch.empty(n, dtype=ps[0].grad.dtype, device=ps[0].grad.device)
                self._flat[g] = flat
            ctx = torch.cuda.stream(self._side_stream) if cuda else _NullCtx()
            with ctx:
                if cuda:
                    self._wait_for_group(events, g)
                off = 0
                for p in ps:
                    flat[off:off + p.grad.numel()].copy_(p.grad.reshape(-1))
                    off += p.grad.numel()
                h = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
                handles.append((h, flat, ps))
        for h, flat, ps in handles:
            h.wait()
            ctx = torch.cuda.stream(self._side_stream) if cuda else _NullCtx()
            with ctx:
                if average:
                    flat.div_(world)
                off = 0
                for p in ps:
                    p.grad.copy_(flat[off:off + p.grad.numel()].view_as(p.grad))
                    off += p.grad.numel()
        if cuda:
            torch.cuda.current_stream().wait_stream(self._side_stream)


    def _wait_for_group(self, events, g):
        """Order the side stream behind the backward kernels that produce bucket `g`. An event that was never recorded
        has no CUDA handle and wait_event() on it is a no-op: fall back to waiting for the whole backward stream."""
        if events is not None and getattr(events[g], "cuda_event", 0):
            self._side_stream.wait_event(events[g])
        else:
            self._side_stream.wait_stream(torch.cuda.current_stream())

    def _reduce_in_place(self, buckets: Dict, world: int, events, average: bool = True) -> bool:
        for g, (flat, pairs) in buckets.items():
            for p, v in pairs:
                if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                    return False
        cuda = any(flat.is_cuda for flat, _ in buckets.values())
        if cuda and self._side_stream is None:
            self._side_stream = torch.cuda.Stream()
        ctx = torch.cuda.stream(self._side_stream) if cuda else _NullCtx()
        handles = []
        with ctx:
            for g in sorted(buckets):
                flat, pairs = buckets[g]
                if not any(p.grad is not None for p, _ in pairs):
                    continue
                if cuda:
                    self._wait_for_group(events, g)
                handles.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=True), flat))
            for h, flat in handles:
                h.wait()
                if average:
                    flat.div_(world)
            if cuda:                                       # when the last collective finished (bench.py: comm_exposed_ms)
                self.last_comm_done = torch.cuda.Event(enable_timing=True)
                self.last_comm_done.record(self._side_stream)
        if cuda:
            torch.cuda.current_stream().wait_stream(self._side_stream)
        return True


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def global_grad_norm(params) -> torch.Tensor:
    """L2 norm over all gradients; identical on every rank after `reduce` (clip_grad_norm_(12.5),
    var_updown/scripts/train.py:173)."""
    sq = [p.grad.float().pow(2).sum() for p in params if p.grad is not None]
    return torch.stack(sq).sum().sqrt()
