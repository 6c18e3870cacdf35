"""CPU, world_size 2, gloo: the host-side logic of the data-parallel path: batch sharding,
gradient-group assignment, bucketed all-reduce averaging and the rank-identical global grad norm."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import sscvae
from style_seqcvae_b200 import dp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


NAMES = [
    "_updown_cell._attention_lstm_cell.weight_ih", "_updown_cell._attention_lstm_cell.bias_hh",
    "_updown_cell._butd_attention._attention_layer.weight", "_updown_cell._language_lstm_cell_encoder.weight_hh",
    "_updown_cell._language_lstm_cell_decoder.weight_ih", "_updown_cell.fc_mean.weight", "_updown_cell.fc_log_var.bias",
    "_output_projection.0.weight", "_output_layer.weight", "_embedding_layer.weight",
]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    params = [(n, torch.nn.Parameter(torch.randn(7, 5, generator=g))) for n in NAMES]
    # per-rank gradients: rank r holds grad = (r+1) * base; the decoder LSTM is "frozen" (no grad)
    for i, (n, p) in enumerate(params):
        if "decoder" in n:
            continue
        p.grad = torch.full_like(p, float(i + 1)) * (rank + 1)
    red = sscvae.BucketedGradReducer(params)
    red.reduce()
    mean_factor = sum(r + 1 for r in range(world)) / world
    ok = True
    for i, (n, p) in enumerate(params):
        if "decoder" in n:
            ok &= p.grad is None
        else:
            ok &= bool(torch.allclose(p.grad, torch.full_like(p, float(i + 1)) * mean_factor))
    # in-place path: gradients that are views into per-bucket flat buffers (UpDownCaptioner.grad_buckets)
    buckets = {}
    for i, (n, p) in enumerate(params):
        if "decoder" in n:
            continue
        buckets.setdefault(dp.group_of(n), []).append((i, p))
    flat = {}
    for gidx, items in buckets.items():
        buf = torch.zeros(sum(p.numel() for _, p in items))
        off, pairs = 0, []
        for i, p in items:
            v = buf[off:off + p.numel()].view_as(p)
            v.copy_(torch.full_like(p, float(i + 1)) * (rank + 1))
            p.grad = v
            pairs.append((p, v))
            off += p.numel()
        flat[gidx] = (buf, pairs)
    red.reduce(buckets=flat)
    for i, (n, p) in enumerate(params):
        if "decoder" not in n:
            ok &= bool(torch.allclose(p.grad, torch.full_like(p, float(i + 1)) * mean_factor))
            ok &= p.grad.data_ptr() == [v for q, v in flat[dp.group_of(n)][1] if q is p][0].data_ptr()
    norm = sscvae.global_grad_norm([p for _, p in params])
    norms = [torch.zeros(()) for _ in range(world)]
    dist.all_gather(norms, norm)
    ok &= all(torch.equal(norms[0], n) for n in norms)       # identical clip coefficient on every rank
    if rank == 0:
        out.put(bool(ok))
    dist.destroy_process_group()


def test_bucketed_reducer_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True


def test_group_assignment_covers_every_trainable_parameter():
    from helpers import StubVocabulary
    for E in (600, 48):
        m = sscvae.UpDownCaptioner(StubVocabulary(50), 32, E, 16, 8, z_space=8, prior_std=1.0, latent_embedding="glove",
                                   sentiment_vae=1, senti_prior_multip=0.5, cbs_simple=True, use_cbs=(E == 600))
        groups = {n: dp.group_of(n) for n, p in m.named_parameters()}
        assert set(groups.values()) == set(range(5))
        assert groups["_updown_cell._language_lstm_cell_decoder.weight_ih"] == 1
        assert groups["_updown_cell.fc_mean.bias"] == 2


def test_shard_batch_partitions_rows():
    for n, w in ((256, 8), (10, 4), (3, 8)):
        idx = []
        for r in range(w):
            sl = sscvae.shard_batch(n, r, w)
            idx += list(range(n))[sl]
        assert idx == list(range(n))
