"""SENTIMENT_VAE = 2, the attribute-grounded prior: per step the prior mean is the attention-weighted
sum of per-box attribute means, `prior_mean_t = sum_n alpha_n * obj_means_n` (var_updown/var_updown/modules/updown_cell.py:160-163),
it conditions the encoder and decoder LSTMs (`c`, :168-190, :211-224; Z columns with latent_embedding "glove", its
first column with "senti_word_net") and the KL is taken against it (updown_captioner.py:295-303); at eval time
z ~ N(prior_mean_t, prior_std^2) (:200-208). The gradient reaches the attention through all three uses.

CPU part: the oracle against fixtures produced by the reference's own forward / backward (oracle/gen_golden.py sv2; the
harness note in oracle/ref_harness.py explains how the reference is constructed), and the host-side translation of
attribute lists. GPU part: the CUDA path against the same fixtures and the same-rounding oracle.
"""
import json

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import search_oracle as so
from oracle import updown_oracle as uo

TRAIN_SV2 = ["train_tied_sv2_glove", "train_untied_sv2_swn"]
DECODE_SV2 = ["decode_greedy_sv2_swn", "decode_greedy_sv2_glove"]


def _params(g, cfg, grad=False):
    p = {k: v.clone().requires_grad_(grad) for k, v in g["params"].items()}
    if cfg.tied:
        p["_output_layer.weight"] = p["_embedding_layer.weight"]
    return p


# ---------------------------------------------------------------------------------------------- CPU: the oracle
@pytest.mark.parametrize("name", TRAIN_SV2)
def test_oracle_train_forward_backward_matches_reference(name):
    g = load_golden(name)
    cfg = uo.OracleConfig(**g["cfg"])
    assert {k: tuple(v.shape) for k, v in g["params"].items()} == uo.param_shapes(cfg)
    p = _params(g, cfg, grad=True)
    out = uo.train_forward(p, cfg, g["image_features"], g["caption_tokens"], g["sentiment"], g["eps"], record=True,
                           obj_means=g["obj_means"])
    torch.testing.assert_close(out["loss"], g["loss"], rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(out["kld"], g["kld"], rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(out["logits"], g["logits"], rtol=1e-4, atol=1e-4)
    uo.train_objective(out).backward()
    for k, gr in g["grads"].items():
        scale = gr.abs().max().item() + 1e-12
        assert (p[k].grad - gr).abs().max().item() / scale < 1e-4, k
    # the prior really follows the attention: the KL changes when the attribute means do
    out2 = uo.train_forward(_params(g, cfg), cfg, g["image_features"], g["caption_tokens"], g["sentiment"], g["eps"],
                            obj_means=g["obj_means"] * 0.5)
    assert not torch.allclose(out2["kld"], g["kld"], rtol=1e-3)


@pytest.mark.parametrize("name", DECODE_SV2)
def test_oracle_greedy_decode_matches_reference(name):
    g = load_golden(name)
    cfg = uo.OracleConfig(**g["cfg"])
    stepper = uo.DecodeStepper(g["params"], cfg, g["image_features"], None, obj_means=g["obj_means"])
    ctr = {"t": 0}

    def step(last, state):
        t = ctr["t"]
        ctr["t"] += 1
        return stepper(last, state, g["eps"][t])
    preds, scores = so.beam_search(torch.ones(1, dtype=torch.long), step, 1, None, 1, 20)
    n = g["predictions"].shape[1]
    assert torch.equal(preds[:, 0, :n], g["predictions"])


def test_attribute_lists_translate_like_the_reference():
    """translate_obj_atts2obj_means (updown_captioner.py:509-532) on the attribute lists of the fixture."""
    import sscvae
    from helpers import StubVocabulary
    import os
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, "train_tied_sv2_glove.npz"))
    g = load_golden("train_tied_sv2_glove")
    cfg = g["cfg"]
    obj_atts = json.loads(str(z["obj_atts"]))
    mean_choice = {k: np.asarray(v) for k, v in json.loads(str(z["mean_choice"])).items()}
    m = sscvae.UpDownCaptioner(
        StubVocabulary(cfg["vocab_size"]), cfg["image_feature_size"], cfg["embedding_size"], cfg["hidden_size"],
        cfg["attention_projection_size"], z_space=cfg["z_space"], latent_embedding="glove", sentiment_vae=2,
        latent_embedding_multip=float(z["latent_embedding_multip"]), mean_choice=mean_choice)
    got = m.translate_obj_atts2obj_means(obj_atts)
    torch.testing.assert_close(got, g["obj_means"], rtol=1e-6, atol=1e-7)
    assert (g["obj_means"].abs().sum(-1) == 0).any() and (g["obj_means"].abs().sum(-1) > 0).any()
    # parameter shapes of the drop-in module = the reference's state_dict
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in g["params"].items()}


# ---------------------------------------------------------------------------------------------- GPU: the CUDA path
TOL_LOGITS_VS_BF16_ORACLE = 6e-3
TOL_LOGITS_VS_FP32_REF = 3e-2
TOL_LOSS_VS_BF16_ORACLE = 3e-3
TOL_LOSS_VS_FP32_REF = 2e-2
TOL_GRAD_VS_FP32_REF = 6e-2
TOL_GRAD_VS_BF16_ORACLE = 2e-2


def _run_cuda(name):
    from helpers import module_from_cfg
    g = load_golden(name)
    m = module_from_cfg(g["cfg"], g["params"])
    m.train()
    m._eps_override = g["eps"].cuda()
    out = m(g["image_features"].cuda(), g["obj_means"].cuda(), None, g["caption_tokens"].cuda(), g["sentiment"].cuda())
    return g, m, out


@pytest.mark.gpu
@pytest.mark.parametrize("name", TRAIN_SV2)
def test_cuda_forward_matches_oracle_and_reference(name, monkeypatch):
    from helpers import oracle_params, rel_err
    monkeypatch.setenv("SSCVAE_DEBUG_LOGITS", "1")
    g, m, out = _run_cuda(name)
    cfg = g["cfg"]
    ocfg = uo.OracleConfig(**cfg)
    B, N, _ = g["image_features"].shape
    T, V, Z = cfg["max_caption_length"] + 1, cfg["vocab_size"], cfg["z_space"]
    ob = uo.train_forward(oracle_params(g["params"], ocfg), ocfg, g["image_features"], g["caption_tokens"], g["sentiment"],
                          g["eps"], q=uo.Rounding("bf16"), record=True, obj_means=g["obj_means"])
    torch.cuda.synchronize()
    logits = m.train_region(B, N, "logits", torch.float32, (T, B, V)).cpu().permute(1, 0, 2)
    alpha = m.train_region(B, N, "alpha", torch.float32, (T, B, N)).cpu()
    pm = m.train_region(B, N, "pm", torch.float32, (T, B, Z)).cpu()
    mean = m.train_region(B, N, "mean", torch.float32, (T, B, Z)).cpu()
    o_alpha = torch.stack([s["alpha"] for s in ob["steps"]])
    o_pm = torch.stack([s["prior_mean"] for s in ob["steps"]])
    o_mean = torch.stack([s["mean"] for s in ob["steps"]])
    assert (alpha - o_alpha).abs().max() < 5e-3
    assert (pm - o_pm).abs().max() < 5e-3 * max(1.0, g["obj_means"].abs().max().item())
    assert rel_err(mean, o_mean) < 1e-2
    assert rel_err(logits, ob["logits"]) < TOL_LOGITS_VS_BF16_ORACLE
    assert rel_err(logits, g["logits"]) < TOL_LOGITS_VS_FP32_REF
    loss, kld = out["loss"].cpu(), out["kld"].cpu()
    scale, kscale = g["loss"].abs().clamp(min=1.0), g["kld"].abs().clamp(min=1.0)
    assert ((loss - ob["loss"].detach()).abs() / scale).max() < TOL_LOSS_VS_BF16_ORACLE
    assert ((loss - g["loss"]).abs() / scale).max() < TOL_LOSS_VS_FP32_REF
    assert ((kld - ob["kld"].detach()).abs() / kscale).max() < TOL_LOSS_VS_BF16_ORACLE
    assert ((kld - g["kld"]).abs() / kscale).max() < TOL_LOSS_VS_FP32_REF


@pytest.mark.gpu
@pytest.mark.parametrize("name", TRAIN_SV2)
def test_cuda_backward_matches_reference_gradients(name):
    from helpers import oracle_params, rel_err
    g, m, out = _run_cuda(name)
    (out["loss"].mean() + out["kld"].mean() / 750.0).backward()
    torch.cuda.synchronize()
    named = dict(m.named_parameters())
    assert set(g["grads"]) == {k for k, p in named.items() if p.grad is not None}
    worst = {k: rel_err(named[k].grad, ref) for k, ref in g["grads"].items()}
    bad = {k: v for k, v in worst.items() if not v < TOL_GRAD_VS_FP32_REF}
    assert not bad, bad
    # and against autograd over the oracle with the same operand rounding
    ocfg = uo.OracleConfig(**g["cfg"])
    p = oracle_params(g["params"], ocfg, grad=True)
    o = uo.train_forward(p, ocfg, g["image_features"], g["caption_tokens"], g["sentiment"], g["eps"], q=uo.Rounding("bf16"),
                         obj_means=g["obj_means"])
    uo.train_objective(o).backward()
    worst = {k: rel_err(prm.grad, p[k].grad) for k, prm in named.items() if prm.grad is not None}
    bad = {k: v for k, v in worst.items() if not v < TOL_GRAD_VS_BF16_ORACLE}
    assert not bad, bad


@pytest.mark.gpu
def test_cuda_kl_weight_reaches_the_attention_through_the_prior():
    """With the CE weight at zero the only path from the objective to W_q / W_v / w_a that does not pass through the
    decoder is prior_mean_t = sum_n alpha_n obj_n: check those gradients against the oracle for a KL-only objective."""
    from helpers import module_from_cfg, oracle_params, rel_err
    g = load_golden("train_untied_sv2_swn")
    m = module_from_cfg(g["cfg"], g["params"])
    m.train()
    m._eps_override = g["eps"].cuda()
    out = m(g["image_features"].cuda(), g["obj_means"].cuda(), None, g["caption_tokens"].cuda(), g["sentiment"].cuda())
    out["kld"].mean().backward()
    ocfg = uo.OracleConfig(**g["cfg"])
    p = oracle_params(g["params"], ocfg, grad=True)
    o = uo.train_forward(p, ocfg, g["image_features"], g["caption_tokens"], g["sentiment"], g["eps"], q=uo.Rounding("bf16"),
                         obj_means=g["obj_means"])
    o["kld"].mean().backward()
    named = dict(m.named_parameters())
    for k in ("_updown_cell._butd_attention._query_vector_projection_layer.weight",
              "_updown_cell._butd_attention._image_features_projection_layer.weight",
              "_updown_cell._butd_attention._attention_layer.weight",
              "_updown_cell.fc_mean.weight"):
        assert p[k].grad.abs().max() > 0
        assert rel_err(named[k].grad, p[k].grad) < TOL_GRAD_VS_BF16_ORACLE, k


@pytest.mark.gpu
@pytest.mark.parametrize("name", DECODE_SV2)
def test_cuda_greedy_decode_matches_reference_tokens(name):
    from helpers import module_from_cfg
    from test_gpu_decode import _replay_path_check
    g = load_golden(name)
    cfg = g["cfg"]
    m = module_from_cfg(cfg, g["params"], beam_size=1, use_cbs=False)
    m.eval()
    m._eps_override = g["eps"].cuda()
    pred = m(g["image_features"].cuda(), g["obj_means"].cuda(), None)["predictions"].cpu()
    _replay_path_check(m, g, cfg, 1, 1, g["eps"], g["image_features"], None, None, obj_means=g["obj_means"])
    n = min(pred.shape[1], g["predictions"].shape[1])
    ref = g["predictions"]
    diff = (pred[0, :n] != ref[0, :n]).nonzero()
    if diff.numel():
        # greedy decoding turns a bf16-level logit difference between two near-tied words into another token: at the first
        # step where the device leaves the reference's caption, the reference's own (fp32) log-probs of the two words must be
        # within the logit tolerance of each other
        t0 = int(diff[0])
        assert t0 >= 8, (pred, ref)
        stepper = uo.DecodeStepper(g["params"], uo.OracleConfig(**cfg), g["image_features"], None, obj_means=g["obj_means"])
        last, state = torch.ones(1, dtype=torch.long), None
        for t in range(t0 + 1):
            logp, state = stepper(last, state, g["eps"][t])
            last = ref[:, t]
        gap = (logp[0, ref[0, t0]] - logp[0, pred[0, t0]]).item()
        assert 0 <= gap < 6e-3 * (logp.max() - logp.min()).item(), (t0, gap)


@pytest.mark.gpu
def test_cuda_beam_and_sampling_share_the_attribute_means_per_image():
    """Wider beams / several samples per image: the reference cannot run them (it does not replicate obj_atts per beam,
    updown_captioner.py:405-424); here the rows of an image read its attribute means through the row -> image map. The
    oracle cell (aligned replication) is replayed along the device's path."""
    from helpers import module_from_cfg
    from test_gpu_decode import _replay_path_check
    g = load_golden("train_untied_sv2_swn")
    cfg = dict(g["cfg"])
    B, K, N, Z = 3, 3, 7, cfg["z_space"]
    m = module_from_cfg(cfg, g["params"], beam_size=K, use_cbs=False)
    m.eval()
    gen = torch.Generator().manual_seed(4)
    feats = torch.rand(B, N, cfg["image_feature_size"], generator=gen)
    feats[1, 4:] = 0
    obj = torch.randn(B, N, Z, generator=gen) * 0.5
    eps = torch.randn(20, B * K, Z, generator=gen)
    m._eps_override = eps.cuda()
    m(feats.cuda(), obj.cuda(), None)
    _replay_path_check(m, None, cfg, 1, K, eps, feats, None, None, obj_means=obj)
    # diverse sampling, J rows per image
    J = 5
    m2 = module_from_cfg(cfg, g["params"], beam_size=1, use_cbs=False)
    m2.eval()
    eps = torch.randn(20, B * J, Z, generator=gen)
    m2._eps_override = eps.cuda()
    out = m2.sample(feats.cuda(), n_samples=J, obj_atts=obj.cuda())
    one = module_from_cfg(cfg, g["params"], beam_size=1, use_cbs=False)
    one.eval()
    for j in range(J):
        one._eps_override = eps[:, j::J].contiguous().cuda()
        p = one(feats.cuda(), obj.cuda(), None)["predictions"].cpu()
        n = min(p.shape[1], out["predictions"].shape[2])
        assert torch.equal(out["predictions"][:, j, :n].cpu(), p[:, :n])
