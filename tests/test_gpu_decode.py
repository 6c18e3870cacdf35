"""GPU: the fused eval-mode decode (greedy / CBS) through UpDownCaptioner.forward against the oracle
and the reference's golden predictions."""
import pytest
import torch

from conftest import load_golden
from helpers import module_from_cfg
from oracle import updown_oracle as uo
from oracle import search_oracle as so

pytestmark = pytest.mark.gpu


def _eps(g, S, K, Z, steps=20):
    e = torch.zeros(steps, S * K, Z)
    e[0, 0] = g["eps0"][0]
    e[1:] = g["eps_rest"]
    return e


@pytest.mark.parametrize("name,K", [("decode_e2e_cbs_k5", 5), ("decode_e2e_greedy", 1)])
def test_decode_matches_oracle(name, K):
    g = load_golden(name)
    cfg = g["cfg"]
    S = g["fsm"].shape[1]
    m = module_from_cfg(cfg, g["params"], beam_size=K, use_cbs=True, min_sat=2)
    m.eval()
    m._eps_override = _eps(g, S, K, cfg["z_space"]).cuda()
    out = m(g["image_features"].cuda(), None, None, fsm=g["fsm"].cuda(), num_constraints=g["num_constraints"].cuda(),
            sentiment=g["sentiment"].cuda())
    pred = out["predictions"].cpu()
    # same-rounding oracle
    ocfg = uo.OracleConfig(**cfg)
    stepper = uo.DecodeStepper(g["params"], ocfg, g["image_features"], g["sentiment"], q=uo.Rounding("bf16"))
    ctr = {"t": 0}

    def step(last, state):
        t = ctr["t"]; ctr["t"] += 1
        return stepper(last, state, g["eps0"] if t == 0 else g["eps_rest"][t - 1])
    op, os_ = so.cbs_search(torch.ones(1, dtype=torch.long), step, g["fsm"], K, (K // 2) or None, 1, 20)
    ob, _ = so.select_best_beam_with_constraints(op, os_, g["num_constraints"], 2)
    assert pred.shape == ob.shape
    # scores of the surviving beams agree to bf16 accumulation noise
    fin = os_ > -1e19
    sc = m.last_search["log_probs"].cpu()
    assert torch.allclose(sc[fin], os_[fin], rtol=2e-2, atol=0.15), (sc[fin], os_[fin])
    agree_oracle = (pred == ob).float().mean().item()
    agree_ref = (pred == g["predictions"]).float().mean().item() if pred.shape == g["predictions"].shape else 0.0
    assert agree_oracle == 1.0 or agree_ref == 1.0, (pred, ob, g["predictions"])


def test_decode_batch_of_images_is_independent_per_image():
    """Images are independent: decoding a batch equals decoding each image alone."""
    g = load_golden("decode_e2e_cbs_k5")
    cfg = g["cfg"]
    S, K, Z = g["fsm"].shape[1], 5, cfg["z_space"]
    m = module_from_cfg(cfg, g["params"], beam_size=K, use_cbs=True)
    m.eval()
    gen = torch.Generator().manual_seed(0)
    B = 3
    feats = torch.rand(B, 7, cfg["image_feature_size"], generator=gen)
    feats[1, 4:] = 0
    fsm = g["fsm"].repeat(B, 1, 1, 1)
    sent = torch.tensor([[1.0], [-1.0], [0.0]])
    eps = torch.randn(20, B * S * K, Z, generator=gen)
    nc = g["num_constraints"].repeat(B)
    m._eps_override = eps.cuda()
    full = m(feats.cuda(), None, None, fsm=fsm.cuda(), num_constraints=nc.cuda(), sentiment=sent.cuda())["predictions"].cpu()
    full_all = m.last_search["predictions"].cpu()
    for b in range(B):
        m._eps_override = eps[:, b * S * K:(b + 1) * S * K].contiguous().cuda()
        one = m(feats[b:b + 1].cuda(), None, None, fsm=fsm[b:b + 1].cuda(), num_constraints=nc[b:b + 1].cuda(),
                sentiment=sent[b:b + 1].cuda())["predictions"].cpu()
        n = min(one.shape[1], full.shape[1])
        assert torch.equal(one[0, :n], full[b, :n]), b


def test_plain_beam_and_greedy_run_without_fsm():
    cfg = dict(vocab_size=300, image_feature_size=64, embedding_size=48, hidden_size=32, attention_projection_size=24,
               z_space=16, sentiment_vae=1, simple_vae=False, max_caption_length=20, prior_std=1.0, senti_prior_multip=0.5)
    torch.manual_seed(0)
    for K in (1, 5):
        m = module_from_cfg(cfg, beam_size=K, use_cbs=False)
        m.eval()
        out = m(torch.rand(4, 5, 64).cuda(), sentiment=torch.zeros(4, 1).cuda())["predictions"]
        assert out.dtype == torch.long and out.shape[0] == 4 and 1 <= out.shape[1] <= 20
        assert (out >= 0).all() and (out < 300).all()
