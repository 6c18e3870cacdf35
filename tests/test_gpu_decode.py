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


def _replay_path_check(m, g, cfg, S, K, eps, feats, sent, fsm, tol=6e-3, obj_means=None):
    """Search amplifies bf16-level logit noise into different (equally valid) beams, so beams are not
    compared token by token with an independent oracle search. Instead the oracle cell is replayed ALONG
    THE PATH THE CUDA SEARCH TOOK (its tokens and back-pointers): every finite beam score the device
    produced must equal parent score + oracle log-prob of the chosen token. This checks the eval cell,
    the row->image sharing, the state gather and the score bookkeeping end to end."""
    B, N = feats.shape[0], feats.shape[1]
    L, R, SK = cfg["max_caption_length"], B * S * K, S * K
    tok = m.decode_region(B, N, S, K, "tok_hist", torch.int32, (L, R)).cpu().long()
    bp = m.decode_region(B, N, S, K, "bp_hist", torch.int32, (L, R)).cpu().long()
    sc = m.decode_region(B, N, S, K, "score_hist", torch.float32, (L, R)).cpu()
    ocfg = uo.OracleConfig(**cfg)
    stepper = uo.DecodeStepper({k: v.detach().cpu() for k, v in m.state_dict().items()}, ocfg, feats, sent,
                               q=uo.Rounding("bf16"), obj_means=obj_means)
    base = (torch.arange(R) // SK) * SK
    logp0, state = stepper(torch.ones(B, dtype=torch.long), None, eps[0, ::SK])
    exp0 = logp0[torch.arange(R) // SK, tok[0]]
    fin = sc[0] > -1e19
    assert fin.any()
    # tolerance = the per-step logit tolerance of the training tests, relative to the logit range
    rng = (logp0.max() - logp0.min()).item()
    worst = (sc[0][fin] - exp0[fin]).abs().max().item() / rng
    # allowed-by-FSM check of the first step (cbs.py:130-136)
    s_of = (torch.arange(R) % SK) // K
    if fsm is not None:
        assert bool(fsm[torch.arange(R) // SK, 0, s_of, tok[0]][fin].bool().all())
    state = {k: v.repeat_interleave(SK, dim=0) for k, v in state.items()}
    n = m.last_search["n_steps"]
    for t in range(1, n):
        logp, new_state = stepper(tok[t - 1], state, eps[t])
        parent = base + bp[t]
        ended = tok[t - 1][parent] == 1
        step_lp = torch.where(ended, torch.zeros(R), logp[parent, tok[t]])
        expect = sc[t - 1][parent] + step_lp
        fin = sc[t] > -1e19
        rng = (logp.max() - logp.min()).item()
        worst = max(worst, (sc[t][fin] - expect[fin]).abs().max().item() / rng)
        assert bool((tok[t][ended & fin] == 1).all())              # boundary stays boundary (cbs.py:177-181)
        if fsm is not None:                                        # every finite beam made an allowed transition
            s_from = (parent % SK) // K
            assert bool(fsm[torch.arange(R) // SK, s_from, s_of, tok[t]][fin].bool().all())
        state = {k: v[parent] for k, v in new_state.items()}
    assert worst < tol, worst
    return worst


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
    eps = _eps(g, S, K, cfg["z_space"])
    m._eps_override = eps.cuda()
    out = m(g["image_features"].cuda(), None, None, fsm=g["fsm"].cuda(), num_constraints=g["num_constraints"].cuda(),
            sentiment=g["sentiment"].cuda())
    pred = out["predictions"].cpu()
    _replay_path_check(m, g, cfg, S, K, eps, g["image_features"], g["sentiment"], g["fsm"])
    # the returned caption is beam 0 of the best valid state (decoding.py:82-134)
    allp, sc = m.last_search["predictions"].cpu(), m.last_search["log_probs"].cpu()
    ob, _ = so.select_best_beam_with_constraints(allp, sc, g["num_constraints"], 2)
    assert torch.equal(pred, ob)
    if K == 1:   # greedy has no competing beams: the reference's own tokens must come out
        assert torch.equal(pred, g["predictions"])


def test_decode_large_batch_path_replay():
    """B=3 images with different sentiments and ragged boxes, S=4 (two constraints), beam 3."""
    from oracle import fsm_oracle as fo
    cfg = dict(vocab_size=400, image_feature_size=64, embedding_size=600, hidden_size=40, attention_projection_size=24,
               z_space=12, sentiment_vae=1, simple_vae=False, max_caption_length=20, prior_std=0.7, senti_prior_multip=0.5)
    torch.manual_seed(3)
    S, K, B = 4, 3, 3
    m = module_from_cfg(cfg, beam_size=K, use_cbs=True)
    m.eval()
    gen = torch.Generator().manual_seed(2)
    feats = torch.rand(B, 6, 64, generator=gen)
    feats[0, 3:] = 0
    sent = torch.tensor([[1.0], [-1.0], [0.0]])
    fsm = torch.from_numpy(fo.single_word_fsm([[5, 6], [9]], 400))[None].repeat(B, 1, 1, 1)
    eps = torch.randn(20, B * S * K, 12, generator=gen)
    m._eps_override = eps.cuda()
    m(feats.cuda(), None, None, fsm=fsm.cuda(), num_constraints=torch.tensor([2] * B).cuda(), sentiment=sent.cuda())
    _replay_path_check(m, None, cfg, S, K, eps, feats, sent, fsm)


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


@pytest.mark.parametrize("H", [32, 40])
def test_sample_equals_the_reference_loop_over_latent_draws(H):
    """sample(n) = the reference's `for k in range(N_Z_SAMPLES): model(image_features)` (inference.py:138-167): with
    the same normals, sample j of image b is exactly the caption the j-th greedy call returns for image b."""
    cfg = dict(vocab_size=300, image_feature_size=64, embedding_size=600, hidden_size=H, attention_projection_size=24,
               z_space=16, sentiment_vae=1, simple_vae=False, max_caption_length=20, prior_std=1.0, senti_prior_multip=0.5)
    torch.manual_seed(1)
    B, J, Z, L = 3, 5, 16, 20
    m = module_from_cfg(cfg, beam_size=1, use_cbs=False)
    m.eval()
    gen = torch.Generator().manual_seed(4)
    feats = torch.rand(B, 6, 64, generator=gen)
    feats[2, 2:] = 0
    sent = torch.tensor([[1.0], [-1.0], [0.0]])
    eps = torch.randn(L, B * J, Z, generator=gen)
    m._eps_override = eps.cuda()
    out = m.sample(feats.cuda(), sentiment=sent.cuda(), n_samples=J)
    pred = out["predictions"].cpu()
    assert pred.shape[:2] == (B, J)
    differ = 0
    for j in range(J):
        m._eps_override = eps[:, j::J].contiguous().cuda()          # row b*J + j of the batched call
        one = m(feats.cuda(), sentiment=sent.cuda())["predictions"].cpu()
        n = min(one.shape[1], pred.shape[2])
        assert torch.equal(one[:, :n], pred[:, j, :n]), j
        assert bool((pred[:, j, n:] == 1).all()) and bool((one[:, n:] == 1).all())   # beyond either early exit: boundary
        differ += int(j > 0 and not torch.equal(pred[:, j], pred[:, 0]))
    m._eps_override = None
    assert differ > 0                                               # different latent draws give different captions
    # device Philox draws: rows are independent sequences, deterministic under a fixed seed
    torch.manual_seed(5)
    a = m.sample(feats.cuda(), sentiment=sent.cuda(), n_samples=J)["predictions"].cpu()
    assert a.shape[:2] == (B, J) and (a >= 0).all() and (a < 300).all()
