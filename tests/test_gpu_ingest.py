"""GPU: feature ingest: bf16 region features straight into the kernels, and reuse of the per-image
decode state across calls on the same feature tensor (the reference's lru_cache on the projected features,
updown-baseline/updown/modules/attention.py:99; its loop over latent samples, var_updown/scripts/inference.py:138-167)."""
import pytest
import torch

import sscvae
from helpers import module_from_cfg

pytestmark = pytest.mark.gpu

CFG = dict(vocab_size=300, image_feature_size=64, embedding_size=600, hidden_size=32, attention_projection_size=24,
           z_space=16, sentiment_vae=1, simple_vae=False, max_caption_length=20, prior_std=1.0, senti_prior_multip=0.5)


def _inputs(B=5, N=7):
    g = torch.Generator().manual_seed(9)
    feats = torch.rand(B, N, 64, generator=g)
    feats[1, 4:] = 0
    toks = torch.randint(2, 300, (B, 20), generator=g)
    for b in range(B):
        toks[b, 5 + b:] = 0
    sent = torch.randint(-1, 2, (B, 1), generator=g).float()
    eps = torch.randn(21, B, 16, generator=g)
    return feats, toks, sent, eps


def test_bf16_features_are_bit_identical_to_their_fp32_values():
    torch.manual_seed(0)
    m = module_from_cfg(CFG)
    m.train()
    feats, toks, sent, eps = _inputs()
    packed = sscvae.pack_features(feats)                       # bf16 host cache entry
    assert packed.dtype == torch.bfloat16 and packed.shape == feats.shape
    res = []
    for x in (packed.cuda(), packed.float().cuda(), feats.cuda()):
        for p in m.parameters():
            p.grad = None
        m._eps_override = eps.cuda()
        out = m(x, None, None, toks.cuda(), sent.cuda())
        (out["loss"].mean() + out["kld"].mean() / 750.0).backward()
        g = m._updown_cell._butd_attention._image_features_projection_layer.weight.grad.clone()
        res.append((out["loss"].clone(), out["kld"].clone(), g))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])
    # and the fp32 originals give the same result too: the kernels round the features to bf16 first
    assert torch.equal(res[0][0], res[2][0]) and torch.equal(res[0][2], res[2][2])
    # collate: ragged per-image arrays -> fixed-N bf16 batch with zero padding rows (datasets.py:623-632)
    c = sscvae.collate_features([feats[0, :3], feats[2]], num_boxes=9)
    assert c.shape == (2, 9, 64) and c.dtype == torch.bfloat16 and float(c[0, 3:].abs().sum()) == 0.0
    cache = sscvae.FeatureCache(7, 64, 8, pin=False)
    for i in range(5):
        cache.put(f"img{i}", feats[i])
    b = cache.batch(["img3", "img1"])
    assert torch.equal(b[0], packed[3]) and torch.equal(b[1], packed[1])


def test_decode_reuses_the_image_state_of_the_same_feature_tensor():
    torch.manual_seed(1)
    m = module_from_cfg(CFG, beam_size=1, use_cbs=False)
    m.eval()
    feats, _, sent, _ = _inputs(B=4)
    x, s = feats.cuda(), sent.cuda()
    eps = torch.randn(20, 4, 16, generator=torch.Generator().manual_seed(2))
    m._eps_override = eps.cuda()
    a = m(x, sentiment=s)["predictions"].clone()
    for _ in range(3):                                           # 2nd call: reuse (eager), 3rd: captured, 4th: replayed
        b = m(x, sentiment=s)["predictions"]
        assert torch.equal(a[:, :min(a.shape[1], b.shape[1])], b[:, :min(a.shape[1], b.shape[1])])
    # a changed tensor (version counter bumps) is NOT served from the cached state
    x2 = x.clone()
    ref = m(x2, sentiment=s)["predictions"].clone()
    x.copy_(torch.rand_like(x))
    c = m(x, sentiment=s)["predictions"]
    fresh = module_from_cfg(CFG, {k: v.detach().cpu() for k, v in m.state_dict().items()}, beam_size=1, use_cbs=False)
    fresh.eval()
    fresh._eps_override = eps.cuda()
    d = fresh(x, sentiment=s)["predictions"]
    n = min(c.shape[1], d.shape[1])
    assert torch.equal(c[:, :n], d[:, :n])
    assert ref.shape[0] == 4
    # new weights invalidate the state as well
    with torch.no_grad():
        m._updown_cell._butd_attention._image_features_projection_layer.weight.mul_(0.5)
    e = m(x, sentiment=s)["predictions"]
    fresh2 = module_from_cfg(CFG, {k: v.detach().cpu() for k, v in m.state_dict().items()}, beam_size=1, use_cbs=False)
    fresh2.eval()
    fresh2._eps_override = eps.cuda()
    f = fresh2(x, sentiment=s)["predictions"]
    n = min(e.shape[1], f.shape[1])
    assert torch.equal(e[:, :n], f[:, :n])
