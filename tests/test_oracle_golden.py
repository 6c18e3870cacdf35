"""CPU: the oracle restatement (oracle/*.py) against the golden vectors produced by running the
unmodified reference (oracle/gen_golden.py). This is what pins the oracle (task ③)."""
import numpy as np
import pytest
import torch

from oracle import updown_oracle as uo
from oracle import search_oracle as so
from oracle import fsm_oracle as fo
from oracle.gen_golden import make_replay_step
from conftest import load_golden

TRAIN = ["train_tied_sv1", "train_tied300_sv0", "train_untied_sv1", "train_tied_simple"]


def _params(g, cfg, grad=False):
    p = {k: v.clone().requires_grad_(grad) for k, v in g["params"].items()}
    if cfg.tied:
        p["_output_layer.weight"] = p["_embedding_layer.weight"]
    return p


@pytest.mark.parametrize("name", TRAIN)
def test_train_forward_backward_matches_reference(name):
    g = load_golden(name)
    cfg = uo.OracleConfig(**g["cfg"])
    p = _params(g, cfg, grad=True)
    out = uo.train_forward(p, cfg, g["image_features"], g["caption_tokens"], g["sentiment"], g["eps"], record=True)
    torch.testing.assert_close(out["loss"], g["loss"], rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(out["kld"], g["kld"], rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(out["logits"], g["logits"], rtol=1e-4, atol=1e-4)
    uo.train_objective(out).backward()
    assert set(g["grads"]) <= set(p)
    for k, gr in g["grads"].items():
        scale = gr.abs().max().item() + 1e-12
        assert (p[k].grad - gr).abs().max().item() / scale < 1e-4, k
    # frozen tied embedding has no gradient in the reference
    if cfg.tied:
        assert "_embedding_layer.weight" not in g["grads"]


def test_param_shapes_match_reference_state_dict():
    for name in TRAIN:
        g = load_golden(name)
        cfg = uo.OracleConfig(**g["cfg"])
        assert {k: tuple(v.shape) for k, v in g["params"].items()} == uo.param_shapes(cfg)


def test_decode_step_matches_reference():
    g = load_golden("decode_step_tied")
    cfg = uo.OracleConfig(**g["cfg"])
    st = uo.DecodeStepper(g["params"], cfg, g["image_features"], g["sentiment"])
    logp, states = st(g["prev"], {k: v.clone() for k, v in g["state_in"].items()}, g["eps"])
    torch.testing.assert_close(logp, g["logp"], rtol=1e-5, atol=1e-5)
    for k, v in g["state_out"].items():
        torch.testing.assert_close(states[k], v, rtol=1e-5, atol=1e-6)


def test_bf16_rounding_mode_is_close_and_differentiable():
    g = load_golden("train_tied_sv1")
    cfg = uo.OracleConfig(**g["cfg"])
    p = _params(g, cfg, grad=True)
    out = uo.train_forward(p, cfg, g["image_features"], g["caption_tokens"], g["sentiment"], g["eps"],
                           q=uo.Rounding("bf16"))
    assert (out["loss"] - g["loss"]).abs().max() / g["loss"].abs().max() < 3e-2
    uo.train_objective(out).backward()
    assert p["_updown_cell.fc_mean.weight"].grad.abs().sum() > 0


CBS = ["cbs_s8_k5", "cbs_s8_k5_repeat", "cbs_s4_k3_b2", "cbs_s1_greedy", "cbs_s12_multiword",
       "cbs_s8_k5_ninf", "cbs_s2_k4_pfull"]


@pytest.mark.parametrize("name", CBS)
def test_cbs_search_matches_reference(name):
    g = load_golden(name)
    B = g["fsm"].shape[0]
    step = make_replay_step(g["tables"], g["ninf"].tolist(), False)
    preds, scores = so.cbs_search(torch.ones(B, dtype=torch.long), step, g["fsm"], g["K"], g["P"] or None,
                                  g["end_index"], g["max_steps"])
    assert preds.shape == g["predictions"].shape        # includes the early-exit step count
    fin = g["scores"] > -1e19
    assert torch.equal(fin, scores > -1e19)
    torch.testing.assert_close(scores[fin], g["scores"][fin], rtol=1e-6, atol=1e-5)
    assert torch.equal(preds[fin], g["predictions"][fin])
    best, _ = so.select_best_beam_with_constraints(preds, scores, g["num_constraints"],
                                                   g["min_constraints_to_satisfy"])
    assert torch.equal(best, g["best"])


@pytest.mark.parametrize("name", ["beam_k5_p2_b3", "beam_k3_pfull_early", "beam_k1_greedy"])
def test_plain_beam_matches_reference(name):
    g = load_golden(name)
    B = g["predictions"].shape[0]
    step = make_replay_step(g["tables"], [], False)
    preds, scores = so.beam_search(torch.ones(B, dtype=torch.long), step, g["K"], g["P"], g["end_index"],
                                   g["max_steps"])
    assert torch.equal(preds, g["predictions"])
    torch.testing.assert_close(scores, g["scores"], rtol=1e-6, atol=1e-5)
    assert torch.equal(so.select_best_beam(preds, scores), g["best"])


@pytest.mark.parametrize("name,K", [("decode_e2e_cbs_k5", 5), ("decode_e2e_greedy", 1)])
def test_full_decode_matches_reference(name, K):
    g = load_golden(name)
    cfg = uo.OracleConfig(**g["cfg"])
    stepper = uo.DecodeStepper(g["params"], cfg, g["image_features"], g["sentiment"])
    ctr = {"t": 0}

    def step(last, state):
        t = ctr["t"]
        ctr["t"] += 1
        return stepper(last, state, g["eps0"] if t == 0 else g["eps_rest"][t - 1])
    preds, scores = so.cbs_search(torch.ones(1, dtype=torch.long), step, g["fsm"], K, (K // 2) or None, 1, 20)
    best, _ = so.select_best_beam_with_constraints(preds, scores, g["num_constraints"], 2)
    assert torch.equal(best, g["predictions"])


def test_fsm_oracle_matches_reference_builder_fixtures():
    wf = {"pos": ["good", "nice", "great"], "neg": ["bad", "ugly"], "dog": ["dog", "dogs"],
          "fire": ["fire"], "hydrant": ["hydrant", "hydrants"], "cat": ["cat"]}
    words = ["@@UNKNOWN@@", "@@BOUNDARY@@"] + [f"w{i}" for i in range(38)]
    for v in wf.values():
        for w in v:
            if w not in words:
                words.append(w)
    idx = {w: i for i, w in enumerate(words)}
    for name, mg in [("cbs_s8_k5", 3), ("cbs_s12_multiword", 3), ("cbs_s4_k3_b2", 2), ("cbs_s2_k4_pfull", 1),
                     ("cbs_s1_greedy", 0)]:
        g = load_golden(name)
        for b, cons in enumerate(g["constraints"]):
            fsm, sub, _ = fo.build_fsm(cons, wf, lambda w: idx.get(w, 0), len(words), max_given_constraints=mg)
            assert np.array_equal(fo.trim_fsm(fsm, sub), g["fsm"][b].numpy()), (name, b)


def test_single_word_fsm_is_a_reference_shaped_fsm():
    fsm = fo.single_word_fsm([[5, 6], [9], [11, 12, 13]], 40)
    assert fsm.shape == (8, 8, 40)
    assert fsm[0, 0, 5] == 0 or fsm[0, 1, 5] == 1
    assert fsm[0, 1, 5] == 1 and fsm[0, 2, 9] == 1 and fsm[0, 4, 11] == 1 and fsm[3, 7, 12] == 1
    assert fsm[7, 7].all()
