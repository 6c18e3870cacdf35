"""GPU: the training-step tail and batched CBS inputs (8(f)-2) on the CUDA path.

  * FusedClipSGD over several iterations against clip_grad_norm_ + torch.optim.SGD(momentum, weight_decay) + LambdaLR
    (var_updown/scripts/train.py:126-134, 173-176), including an interval in which the decoder LSTM is frozen by the
    reference's schedule (train.py:156-161), in both zero_grad semantics (torch >= 2: None; the reference's pinned
    torch 1.1: zero tensors);
  * pad_fsm_batch batches (S = 2 / 4 / 8 mixed) through the CUDA constrained beam search against one image at a time.
"""
import pytest
import torch

import sscvae
from helpers import module_from_cfg
from oracle import fsm_oracle as fo

pytestmark = pytest.mark.gpu

CFG = dict(vocab_size=300, image_feature_size=64, embedding_size=600, hidden_size=32, attention_projection_size=24,
           z_space=16, sentiment_vae=1, simple_vae=False, max_caption_length=20, prior_std=1.0, senti_prior_multip=0.5)


@pytest.mark.parametrize("legacy", [False, True])
def test_fused_clip_sgd_matches_torch_sgd_over_a_freeze_schedule(legacy):
    torch.manual_seed(0)
    m = module_from_cfg(CFG)
    m.train()
    named = [(k, p) for k, p in m.named_parameters() if p.requires_grad]
    ref = {k: p.detach().clone().requires_grad_(True) for k, p in named}
    LR, MOM, WD, CLIP, NIT = 0.05, 0.9, 1e-3, 0.5, 40            # small clip threshold: the clip is active
    ours = sscvae.FusedClipSGD([p for _, p in named], lr=LR, momentum=MOM, weight_decay=WD, max_norm=CLIP,
                               num_iterations=NIT, legacy_zero_grad=legacy)
    topt = torch.optim.SGD(list(ref.values()), lr=LR, momentum=MOM, weight_decay=WD)
    sched = torch.optim.lr_scheduler.LambdaLR(topt, lr_lambda=lambda it: 1 - it / NIT)      # train.py:132-134
    g = torch.Generator().manual_seed(1)
    dec = "_updown_cell._language_lstm_cell_decoder"
    frozen_its = {2, 3, 5}                                          # train.py:156-161: decoder LSTM frozen on these
    for it in range(7):
        frozen = it in frozen_its
        ours.zero_grad()
        topt.zero_grad(set_to_none=not legacy)
        grads = {k: torch.randn(p.shape, generator=g).to(p.device) * (3.0 if it % 2 else 0.01) for k, p in named}
        for k, p in named:
            if frozen and k.startswith(dec):
                continue                                            # requires_grad = False: backward leaves no gradient
            p.grad = grads[k].clone()
            if ref[k].grad is None:
                ref[k].grad = grads[k].clone()
            else:
                ref[k].grad.copy_(grads[k])
        torch.nn.utils.clip_grad_norm_(list(ref.values()), CLIP)    # train.py:173
        topt.step()
        sched.step()
        ours.step()
        assert abs(ours.lr - sched.get_last_lr()[0]) < 1e-9
        for k, p in named:
            assert torch.allclose(p.detach(), ref[k].detach(), rtol=2e-5, atol=2e-6), (it, k)
    # the two semantics differ exactly on the frozen tensors
    assert ours.iteration == 7


def test_padded_fsm_batch_through_the_cuda_search_equals_one_image_at_a_time():
    V = CFG["vocab_size"]
    torch.manual_seed(2)
    K = 5
    m = module_from_cfg(CFG, beam_size=K, use_cbs=True)
    m.eval()
    cons = [[[5, 6], [9]], [[11]], [[7], [8, 10], [20, 21, 22]], [[30]]]          # 2, 1, 3, 1 constraints -> S = 4, 2, 8, 2
    fsms = [torch.from_numpy(fo.single_word_fsm(c, V)) for c in cons]
    ncs = [len(c) for c in cons]
    fsm, nc = sscvae.pad_fsm_batch(fsms, ncs)
    B, S = fsm.shape[0], fsm.shape[1]
    assert (B, S) == (4, 8)
    gen = torch.Generator().manual_seed(3)
    feats = torch.rand(B, 7, CFG["image_feature_size"], generator=gen)
    feats[1, 4:] = 0
    sent = torch.tensor([[1.0], [-1.0], [0.0], [1.0]])
    Z = CFG["z_space"]
    eps = torch.randn(20, B * S * K, Z, generator=gen)
    m._eps_override = eps.cuda()
    full = m(feats.cuda(), None, None, fsm=fsm.cuda(), num_constraints=nc.cuda(), sentiment=sent.cuda())["predictions"].cpu()
    full_sc = m.last_search["log_probs"].cpu()
    for b in range(B):
        Sb = fsms[b].shape[0]
        # the image alone with its OWN (unpadded) FSM: rows (s, k) of the padded run with s < Sb must see the same noise
        e_b = eps[:, b * S * K:(b + 1) * S * K].reshape(20, S, K, Z)[:, :Sb].reshape(20, Sb * K, Z).contiguous()
        m._eps_override = e_b.cuda()
        one = m(feats[b:b + 1].cuda(), None, None, fsm=fsms[b][None].cuda(), num_constraints=nc[b:b + 1].cuda(),
                sentiment=sent[b:b + 1].cuda())["predictions"].cpu()
        one_sc = m.last_search["log_probs"].cpu()
        n = min(one.shape[1], full.shape[1])
        assert torch.equal(one[0, :n], full[b, :n]), b
        fin = one_sc[0] > -1e19
        assert torch.allclose(full_sc[b, :Sb][fin], one_sc[0][fin], rtol=0, atol=1e-4), b
        assert bool((full_sc[b, Sb:] < -1e19).all())               # padded states are never entered
