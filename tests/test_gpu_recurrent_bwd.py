"""The persistent BPTT kernel (csrc/recurrent_bwd.cu: one cooperative launch for all T reverse timesteps) against the
per-launch backward it replaces (api_train.cu: ten kernels per step), on identical inputs.

Both paths compute the same math with the same operand rounding (bf16 gate gradients, fp32 accumulation); they differ
in summation order (split-K slots) only, so the agreement is much tighter than the oracle tolerances. The oracle /
reference parity of whichever path is the default is covered by test_gpu_train.py (golden sizes) and
test_gpu_fullsize.py (dims Y and D).
"""
import pytest
import torch

from conftest import load_golden
from helpers import module_from_cfg, rel_err
import sscvae
from test_gpu_fullsize import DIMS_Y, DIMS_D, _batch

pytestmark = pytest.mark.gpu

TOL_BETWEEN_PATHS = 4e-3     # max-abs difference of a gradient relative to its largest entry


def _lib():
    return sscvae._lib


def _grads(m, feats, toks, sent, eps, persistent):
    L = _lib()
    L.check(L.lib().sscvae_set_option(m._handle, b"persistent_bwd", int(persistent)))
    for p in m.parameters():
        p.grad = None
    m._eps_override = eps
    out = m(feats, None, None, toks, sent)
    (out["loss"].mean() + out["kld"].mean() / 750.0).backward()
    torch.cuda.synchronize()
    return {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}


def _compare(m, feats, toks, sent, eps, B, N):
    L = _lib()
    L.check(L.lib().sscvae_set_option(m._handle, b"persistent_bwd", 1))
    assert L.lib().sscvae_train_backward_is_persistent(m._handle, B, N) == 1, "the persistent BPTT kernel does not cover this shape"
    per_launch = _grads(m, feats, toks, sent, eps, False)
    # three times: eager, graph capture, graph replay
    for _ in range(3):
        persistent = _grads(m, feats, toks, sent, eps, True)
        assert set(persistent) == set(per_launch)
        bad = {k: rel_err(persistent[k], per_launch[k]) for k in per_launch}
        bad = {k: v for k, v in bad.items() if not v < TOL_BETWEEN_PATHS}
        assert not bad, bad


@pytest.mark.parametrize("name", ["train_tied_sv1", "train_tied300_sv0", "train_untied_sv1", "train_tied_simple"])
def test_persistent_bwd_matches_per_launch_at_golden_sizes(name):
    g = load_golden(name)
    m = module_from_cfg(g["cfg"], g["params"])
    m.train()
    B, N, _ = g["image_features"].shape
    _compare(m, g["image_features"].cuda(), g["caption_tokens"].cuda(), g["sentiment"].cuda(), g["eps"].cuda(), B, N)


@pytest.mark.parametrize("name,cfg,B", [("Y4", DIMS_Y, 4), ("Y256", DIMS_Y, 256), ("Y130", DIMS_Y, 130), ("D8", DIMS_D, 8)])
def test_persistent_bwd_matches_per_launch_at_full_dims(name, cfg, B):
    torch.manual_seed(0)
    m = module_from_cfg(cfg)
    m.train()
    N = 36
    feats, toks, sent, eps = _batch(cfg, B, N, 5)
    _compare(m, feats.cuda(), toks.cuda(), sent.cuda(), eps.cuda(), B, N)
