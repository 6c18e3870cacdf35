"""Generates the golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

TEST INFRASTRUCTURE ONLY; runs only in the build container (needs /root/reference):

    python -m oracle.gen_golden            # rewrites tests/golden/*.npz

The reference has no tests or golden vectors of its own, so these files are what pins
the oracle (and through it the CUDA path): they hold seeded inputs, the reference's state_dict and
the reference's own outputs (loss, kld, per-step logits, parameter gradients, eval log-probs,
CBS / beam token ids and scores, FSM tensors from the reference's FiniteStateMachineBuilder).
While generating, every fixture is also checked against the oracle restatement so a mismatch
fails here rather than later.
"""
import csv
import json
import os
import tempfile

import numpy as np
import torch

from oracle import ref_harness as rh
from oracle import updown_oracle as uo
from oracle import search_oracle as so
from oracle import fsm_oracle as fo

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
EPS_SEED = 1234


def _np(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def synthetic_batch(B, N, F, V, L, seed, ragged=True):
    g = torch.Generator().manual_seed(seed)
    feats = torch.rand(B, N, F, generator=g)
    if ragged:
        for b in range(B):
            n = int(torch.randint(max(1, N // 3), N + 1, (1,), generator=g))
            feats[b, n:] = 0
    lengths = torch.randint(1, L + 1, (B,), generator=g)
    lengths[0] = L                      # one full-length caption
    if B > 2:
        lengths[2] = 0                  # one empty caption (only the two boundary tokens)
    toks = torch.randint(2, V, (B, L), generator=g)
    for b in range(B):
        toks[b, int(lengths[b]):] = 0
    sentiment = torch.randint(-1, 2, (B, 1), generator=g).float()
    return feats, toks, sentiment


def gen_train(name, *, E, sv, simple=False, V=120, F=64, H=32, A=24, Z=16, N=7, B=6, L=20,
              prior_std=1.0, seed=0):
    vocab = rh.make_vocabulary(V)
    m = rh.build_reference_model(vocab, image_feature_size=F, embedding_size=E, hidden_size=H,
                                 attention_projection_size=A, z_space=Z, sentiment_vae=sv,
                                 simple_vae=simple, max_caption_length=L, prior_std=prior_std, seed=seed)
    cfg = dict(vocab_size=V, image_feature_size=F, embedding_size=E, hidden_size=H,
               attention_projection_size=A, z_space=Z, sentiment_vae=sv, simple_vae=simple,
               max_caption_length=L, prior_std=prior_std, senti_prior_multip=0.5)
    feats, toks, sentiment = synthetic_batch(B, N, F, V, L, seed + 1)
    m.train()
    # record per-step logits through a forward hook on the output layer
    logits_rec = []
    hk = m._output_layer.register_forward_hook(lambda mod, i, o: logits_rec.append(o.detach().clone()))
    torch.manual_seed(EPS_SEED)
    out = m(feats.clone(), None, None, toks, sentiment)
    hk.remove()
    (out["loss"].mean() + out["kld"].mean() / 750.0).backward()
    torch.manual_seed(EPS_SEED)
    eps = torch.stack([torch.randn(B, Z) for _ in range(L + 1)])
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    grads = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}
    # cross-check the oracle right here
    ocfg = uo.OracleConfig(**cfg)
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    if ocfg.tied:
        p["_output_layer.weight"] = p["_embedding_layer.weight"]
    o = uo.train_forward(p, ocfg, feats, toks, sentiment, eps, record=True)
    assert torch.allclose(o["loss"], out["loss"], rtol=1e-5, atol=1e-4), name
    assert torch.allclose(o["kld"], out["kld"], rtol=1e-5, atol=1e-4), name
    assert torch.allclose(o["logits"], torch.stack(logits_rec, 1), rtol=1e-4, atol=1e-4), name
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        cfg=json.dumps(cfg), image_features=feats.numpy(), caption_tokens=toks.numpy(),
        sentiment=sentiment.numpy(), eps=eps.numpy(), loss=out["loss"].detach().numpy(),
        kld=out["kld"].detach().numpy(), logits=torch.stack(logits_rec, 1).numpy(),
        **{"param:" + k: v.numpy() for k, v in sd.items()},
        **{"grad:" + k: v.numpy() for k, v in grads.items()})
    print("wrote", name, "loss", out["loss"].detach().numpy().round(3))


ATT_WORDS = ["red", "old", "shiny", "happy", "broken", "wet"]


def make_mean_choice(Z, latent_embedding, seed):
    """{attribute word: (Z,) vector} like the tables the reference loads (updown_captioner.py:79-86): "senti_word_net"
    repeats one score over all Z components, "glove" has a free vector per word."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for w in ATT_WORDS:
        if latent_embedding == "senti_word_net":
            out[w] = np.repeat(float(torch.rand(1, generator=g)) * 2 - 1, Z)
        else:
            out[w] = (torch.randn(Z, generator=g) * 0.5).numpy().astype(np.float64)
    return out


def make_obj_atts(B, N, seed):
    """per image N `(object, [attribute strings])` entries in the reference's format (updown_captioner.py:509-520):
    the first word of an attribute string is looked up; unknown words and boxes without attributes occur."""
    g = torch.Generator().manual_seed(seed)
    words = ATT_WORDS + ["unknownword"]
    out = []
    for b in range(B):
        boxes = []
        for n in range(N):
            k = int(torch.randint(0, 4, (1,), generator=g))
            atts = [words[int(torch.randint(0, len(words), (1,), generator=g))] + " thing" for _ in range(k)]
            boxes.append(("obj%d" % n, atts))
        out.append(boxes)
    return out


def gen_train_sv2(name, *, E, le, Z, V=120, F=64, H=32, A=24, N=7, B=5, L=20, prior_std=1.0, multip=2.0, seed=0):
    """SENTIMENT_VAE = 2 training forward + backward of the reference (attribute-grounded prior, updown_cell.py:160-190)."""
    vocab = rh.make_vocabulary(V)
    mc = make_mean_choice(Z, le, seed + 5)
    m = rh.build_reference_model(vocab, image_feature_size=F, embedding_size=E, hidden_size=H,
                                 attention_projection_size=A, z_space=Z, sentiment_vae=2, latent_embedding=le,
                                 max_caption_length=L, prior_std=prior_std, seed=seed, latent_embedding_multip=multip,
                                 mean_choice=mc)
    cfg = dict(vocab_size=V, image_feature_size=F, embedding_size=E, hidden_size=H,
               attention_projection_size=A, z_space=Z, sentiment_vae=2, simple_vae=False, latent_embedding=le,
               max_caption_length=L, prior_std=prior_std, senti_prior_multip=0.5)
    feats, toks, sentiment = synthetic_batch(B, N, F, V, L, seed + 1)
    obj_atts = make_obj_atts(B, N, seed + 2)
    obj_means = m.translate_obj_atts2obj_means(obj_atts)
    m.train()
    logits_rec = []
    hk = m._output_layer.register_forward_hook(lambda mod, i, o: logits_rec.append(o.detach().clone()))
    torch.manual_seed(EPS_SEED)
    out = m(feats.clone(), obj_atts, None, toks, sentiment)
    hk.remove()
    (out["loss"].mean() + out["kld"].mean() / 750.0).backward()
    torch.manual_seed(EPS_SEED)
    eps = torch.stack([torch.randn(B, Z) for _ in range(L + 1)])
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    grads = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}
    ocfg = uo.OracleConfig(**cfg)
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    if ocfg.tied:
        p["_output_layer.weight"] = p["_embedding_layer.weight"]
    o = uo.train_forward(p, ocfg, feats, toks, sentiment, eps, record=True, obj_means=obj_means)
    assert torch.allclose(o["loss"], out["loss"], rtol=1e-5, atol=1e-4), name
    assert torch.allclose(o["kld"], out["kld"], rtol=1e-5, atol=1e-4), name
    assert torch.allclose(o["logits"], torch.stack(logits_rec, 1), rtol=1e-4, atol=1e-4), name
    uo.train_objective(o).backward()
    for k, gref in grads.items():
        assert torch.allclose(p[k].grad, gref, rtol=1e-3, atol=1e-5), (name, k)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        cfg=json.dumps(cfg), image_features=feats.numpy(), caption_tokens=toks.numpy(),
        sentiment=sentiment.numpy(), eps=eps.numpy(), loss=out["loss"].detach().numpy(),
        kld=out["kld"].detach().numpy(), logits=torch.stack(logits_rec, 1).numpy(),
        obj_means=obj_means.numpy(), obj_atts=json.dumps(obj_atts), latent_embedding_multip=multip,
        mean_choice=json.dumps({k: np.asarray(v).tolist() for k, v in mc.items()}),
        **{"param:" + k: v.numpy() for k, v in sd.items()},
        **{"grad:" + k: v.numpy() for k, v in grads.items()})
    print("wrote", name, "loss", out["loss"].detach().numpy().round(3), "kld", out["kld"].detach().numpy().round(3))


def gen_decode_sv2(name, *, le, Z, E=600, V0=40, F=64, H=32, A=24, N=7, seed=21):
    """SENTIMENT_VAE = 2 eval forward of the reference at B = 1, beam 1 (the reference does not replicate obj_atts per
    beam, updown_captioner.py:405-424, so wider beams do not run): z ~ N(sum_n alpha_n obj_n, prior_std^2) per step."""
    ref = rh.load_reference()
    with tempfile.TemporaryDirectory() as d:             # the unconstrained one-state FSM, built by the reference
        tsv = os.path.join(d, "wf.tsv")
        with open(tsv, "w") as f:
            f.write("dog\tdog,dogs\n")
        vocab = rh.make_vocabulary(V0)
        vocab = ref.add_constraint_words_to_vocabulary(vocab, tsv)
        V = vocab.get_vocab_size()
        builder = ref.FiniteStateMachineBuilder(vocab, tsv, None, max_given_constraints=0)
        fsm, nstates, _ = builder.build([])
        fsm = fsm[None, :nstates, :nstates]
    assert fsm.shape[1] == 1
    mc = make_mean_choice(Z, le, seed + 5)
    m = rh.build_reference_model(vocab, image_feature_size=F, embedding_size=E, hidden_size=H,
                                 attention_projection_size=A, z_space=Z, sentiment_vae=2, latent_embedding=le,
                                 beam_size=1, seed=seed, latent_embedding_multip=1.5, mean_choice=mc, prior_std=0.7)
    cfg = dict(vocab_size=V, image_feature_size=F, embedding_size=E, hidden_size=H,
               attention_projection_size=A, z_space=Z, sentiment_vae=2, simple_vae=False, latent_embedding=le,
               max_caption_length=20, prior_std=0.7, senti_prior_multip=0.5)
    m.eval()
    feats, _, _ = synthetic_batch(1, N, F, V, 20, seed + 1, ragged=False)
    obj_atts = make_obj_atts(1, N, seed + 2)
    obj_means = m.translate_obj_atts2obj_means(obj_atts)
    torch.manual_seed(EPS_SEED)
    with torch.no_grad():
        out = m(feats.clone(), obj_atts, None, fsm=fsm, num_constraints=torch.tensor([0]), constraints=None,
                constraint2states=None, sentiment=torch.zeros(1, 1))
    pred = out["predictions"]
    torch.manual_seed(EPS_SEED)
    eps = torch.stack([torch.randn(1, Z) for _ in range(20)])
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    ocfg = uo.OracleConfig(**cfg)
    stepper = uo.DecodeStepper(sd, ocfg, feats, None, obj_means=obj_means)
    ctr = {"t": 0}

    def step(last, state):
        t = ctr["t"]
        ctr["t"] += 1
        return stepper(last, state, eps[t])
    p2, s2 = so.cbs_search(torch.ones(1, dtype=torch.long), step, fsm, 1, None, 1, 20)
    b2, _ = so.select_best_beam_with_constraints(p2, s2, torch.tensor([0]), 2)
    assert torch.equal(b2, pred), (b2, pred)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), cfg=json.dumps(cfg), image_features=feats.numpy(),
                        obj_means=obj_means.numpy(), eps=eps.numpy(), predictions=pred.numpy(),
                        **{"param:" + k: v.numpy() for k, v in sd.items()})
    print("wrote", name, "pred", pred.tolist())


def gen_decode_step(name, *, E=600, V=120, F=64, H=32, A=24, Z=16, N=7, B=3, nb=4, seed=3):
    """Reference eval-mode `_decode_step` with replicated rows (updown_captioner.py:405-424),
    uniform sentiment so the reference's tiled repeat equals the aligned one."""
    vocab = rh.make_vocabulary(V)
    m = rh.build_reference_model(vocab, image_feature_size=F, embedding_size=E, hidden_size=H,
                                 attention_projection_size=A, z_space=Z, sentiment_vae=1, seed=seed)
    cfg = dict(vocab_size=V, image_feature_size=F, embedding_size=E, hidden_size=H,
               attention_projection_size=A, z_space=Z, sentiment_vae=1, simple_vae=False,
               max_caption_length=20, prior_std=1.0, senti_prior_multip=0.5)
    m.eval()
    feats, _, _ = synthetic_batch(B, N, F, V, 20, seed + 1)
    sentiment = torch.ones(B, 1)
    R = B * nb
    g = torch.Generator().manual_seed(seed + 2)
    prev = torch.randint(1, V, (R,), generator=g)
    states = {k: torch.randn(R, H, generator=g) * 0.3 for k in
              ("h1", "c1", "h_encoder", "c_encoder", "h_decoder", "c_decoder")}
    prior_mean = sentiment.repeat(1, Z) * 0.5
    prior_var = torch.ones(B, Z)
    torch.manual_seed(EPS_SEED)
    with torch.no_grad():
        logp, new_states, *_ = m._decode_step(feats.clone(), None, prev, {k: v.clone() for k, v in states.items()},
                                              sentiment=sentiment, prior_mean=prior_mean, prior_var=prior_var)
    torch.manual_seed(EPS_SEED)
    eps = torch.randn(R, Z)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    ocfg = uo.OracleConfig(**cfg)
    stepper = uo.DecodeStepper(sd, ocfg, feats, sentiment)
    lp2, st2 = stepper(prev, {k: v.clone() for k, v in states.items()}, eps)
    assert torch.allclose(lp2, logp, rtol=1e-5, atol=1e-5)
    for k in new_states:
        assert torch.allclose(st2[k], new_states[k], rtol=1e-5, atol=1e-6), k
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), cfg=json.dumps(cfg), image_features=feats.numpy(),
        sentiment=sentiment.numpy(), prev=prev.numpy(), eps=eps.numpy(), logp=logp.numpy(),
        **{"state_in:" + k: v.numpy() for k, v in states.items()},
        **{"state_out:" + k: v.numpy() for k, v in new_states.items()},
        **{"param:" + k: v.numpy() for k, v in sd.items()})
    print("wrote", name)


# ------------------------------------------------------------------------------------------
# search fixtures: replayed step function (a pure function of step index, previous token and a
# history-dependent state) so that token ids can be compared exactly.
# ------------------------------------------------------------------------------------------
def make_replay_step(tables: torch.Tensor, ninf_cols, five_tuple: bool):
    """tables (steps,V,V): logits row for (step, previous token). State `acc` (R,1) accumulates a
    function of the path so a wrong back-pointer gather changes later log-probs."""
    V = tables.shape[-1]
    ctr = {"t": 0}
    bias = torch.linspace(-1, 1, V).view(1, V)

    def step(last, state, *unused):
        t = ctr["t"]
        ctr["t"] += 1
        if state is None:
            state = {"acc": torch.zeros(last.shape[0], 1)}
        logits = tables[t][last] + state["acc"] * bias
        if len(ninf_cols):
            logits[:, ninf_cols] = float("-inf")
        logp = torch.log_softmax(logits, dim=1)
        new_state = {"acc": state["acc"] + ((last % 7).float().unsqueeze(1) - 3.0) * 0.05}
        return (logp, new_state, None, None, None) if five_tuple else (logp, new_state)
    return step


def search_tables(V, steps, seed, end_bias_from=None):
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(steps, V, V, generator=g) * 2.0
    if end_bias_from is not None:      # make the boundary token win from some step on -> early exit
        t[end_bias_from:, :, 1] += 30.0
    return t


def gen_cbs(name, constraints, *, mg, K, P_from_half=True, B=1, seed=0, end_bias_from=None,
            ninf=(), nc=None, min_sat=2, max_steps=20):
    ref = rh.load_reference()
    # small synthetic wordform table; ids are stable because the vocabulary is built in order
    wf = {"pos": ["good", "nice", "great"], "neg": ["bad", "ugly"], "dog": ["dog", "dogs"],
          "fire": ["fire"], "hydrant": ["hydrant", "hydrants"], "cat": ["cat"]}
    with tempfile.TemporaryDirectory() as d:
        tsv = os.path.join(d, "wf.tsv")
        with open(tsv, "w") as f:
            for k, v in wf.items():
                f.write(k + "\t" + ",".join(v) + "\n")
        vocab = rh.make_vocabulary(40)
        vocab = ref.add_constraint_words_to_vocabulary(vocab, tsv)
        V = vocab.get_vocab_size()
        builder = ref.FiniteStateMachineBuilder(vocab, tsv, None, max_given_constraints=mg)
        fsms = []
        for b in range(B):
            fsm, nstates, c2s = builder.build(constraints[b])
            fsms.append(fsm[:nstates, :nstates])
            f2, n2, c2 = fo.build_fsm(constraints[b], wf, vocab.get_token_index, V, max_given_constraints=mg)
            assert n2 == nstates and np.array_equal(fo.trim_fsm(f2, n2), fsms[-1].numpy()) and c2 == c2s
    S = fsms[0].shape[0]
    assert all(f.shape[0] == S for f in fsms)
    fsm = torch.stack(fsms)                      # (B,S,S,V) uint8
    P = (K // 2) if P_from_half else K           # captioner passes beam_size // 2 (updown_captioner.py:134)
    tables = search_tables(V, max_steps, seed, end_bias_from)
    cbs = ref.ConstrainedBeamSearch(1, max_steps=max_steps, beam_size=K, per_node_beam_size=P)
    start = torch.ones(B, dtype=torch.long)
    preds, scores = cbs.search(start, None, make_replay_step(tables, list(ninf), True), fsm)
    ncs = torch.tensor([len(c) for c in constraints] if nc is None else nc)
    best, _valid = ref.select_best_beam_with_constraints(preds, scores, ncs, None, None, min_sat, True)
    # oracle cross-check (finite-score beams only; tokens up to first boundary)
    p2, s2 = so.cbs_search(start, make_replay_step(tables, list(ninf), False), fsm, K, P or None, 1, max_steps)
    assert p2.shape == preds.shape, (p2.shape, preds.shape)
    fin = scores > -1e19
    assert torch.equal(fin, s2 > -1e19)
    assert torch.allclose(scores[fin], s2[fin], rtol=1e-6, atol=1e-5)
    assert torch.equal(preds[fin], p2[fin]), name
    b2, _ = so.select_best_beam_with_constraints(p2, s2, ncs, min_sat)
    assert torch.equal(b2, best)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), fsm=fsm.numpy(), tables=tables.numpy(),
                        K=K, P=P, end_index=1, max_steps=max_steps, ninf=np.asarray(list(ninf), dtype=np.int64),
                        predictions=preds.numpy(), scores=scores.numpy(), num_constraints=ncs.numpy(),
                        min_constraints_to_satisfy=min_sat, best=best.numpy(),
                        constraints=json.dumps(constraints))
    print("wrote", name, "S", S, "steps", preds.shape[-1], "finite beams", int(fin.sum()), "/", fin.numel())


def gen_beam(name, *, K, P, B, V=50, seed=0, end_bias_from=None, max_steps=20):
    ref = rh.load_reference()
    tables = search_tables(V, max_steps, seed, end_bias_from)
    bs = ref.BeamSearch(1, max_steps=max_steps, beam_size=K, per_node_beam_size=P)
    start = torch.ones(B, dtype=torch.long)
    preds, scores = bs.search(start, None, make_replay_step(tables, [], False))
    p2, s2 = so.beam_search(start, make_replay_step(tables, [], False), K, P, 1, max_steps)
    assert torch.equal(preds, p2) and torch.allclose(scores, s2, rtol=1e-6, atol=1e-5), name
    np.savez_compressed(os.path.join(OUT, name + ".npz"), tables=tables.numpy(), K=K, P=P or K,
                        end_index=1, max_steps=max_steps, predictions=preds.numpy(), scores=scores.numpy(),
                        best=ref.select_best_beam(preds, scores).numpy())
    print("wrote", name, "steps", preds.shape[-1])


def gen_decode_e2e(name, *, K, constraints, mg, E=600, V0=40, F=64, H=32, A=24, Z=16, N=7, seed=5):
    """Full reference eval forward: UpDownCaptioner.forward(..., fsm=...) at B=1 (the only batch
    size the reference's collate supports, datasets.py:604-620)."""
    ref = rh.load_reference()
    wf = {"pos": ["good", "nice", "great"], "neg": ["bad", "ugly"], "dog": ["dog", "dogs"]}
    with tempfile.TemporaryDirectory() as d:
        tsv = os.path.join(d, "wf.tsv")
        with open(tsv, "w") as f:
            for k, v in wf.items():
                f.write(k + "\t" + ",".join(v) + "\n")
        vocab = rh.make_vocabulary(V0)
        vocab = ref.add_constraint_words_to_vocabulary(vocab, tsv)
        V = vocab.get_vocab_size()
        builder = ref.FiniteStateMachineBuilder(vocab, tsv, None, max_given_constraints=mg)
        fsm, nstates, _ = builder.build(constraints)
        fsm = fsm[None, :nstates, :nstates]
    m = rh.build_reference_model(vocab, image_feature_size=F, embedding_size=E, hidden_size=H,
                                 attention_projection_size=A, z_space=Z, sentiment_vae=1, beam_size=K,
                                 min_constraints_to_satisfy=2, seed=seed)
    cfg = dict(vocab_size=V, image_feature_size=F, embedding_size=E, hidden_size=H,
               attention_projection_size=A, z_space=Z, sentiment_vae=1, simple_vae=False,
               max_caption_length=20, prior_std=1.0, senti_prior_multip=0.5)
    m.eval()
    feats, _, _ = synthetic_batch(1, N, F, V, 20, seed + 1)
    sentiment = torch.ones(1, 1)
    S = fsm.shape[1]
    torch.manual_seed(EPS_SEED)
    with torch.no_grad():
        out = m(feats.clone(), None, None, fsm=fsm, num_constraints=torch.tensor([len(constraints)]),
                constraints=None, constraint2states=None, sentiment=sentiment)
    pred = out["predictions"]
    # replay eps in the same draw order: step 0 draws (1,Z), later steps (S*K,Z)
    torch.manual_seed(EPS_SEED)
    eps0 = torch.randn(1, Z)
    eps_rest = torch.stack([torch.randn(S * K, Z) for _ in range(19)])
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    ocfg = uo.OracleConfig(**cfg)
    stepper = uo.DecodeStepper(sd, ocfg, feats, sentiment)
    ctr = {"t": 0}

    def step(last, state):
        t = ctr["t"]
        ctr["t"] += 1
        return stepper(last, state, eps0 if t == 0 else eps_rest[t - 1])
    p2, s2 = so.cbs_search(torch.ones(1, dtype=torch.long), step, fsm, K, (K // 2) or None, 1, 20)
    b2, _ = so.select_best_beam_with_constraints(p2, s2, torch.tensor([len(constraints)]), 2)
    assert torch.equal(b2, pred), (b2, pred)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), cfg=json.dumps(cfg), fsm=fsm.numpy(), K=K,
                        image_features=feats.numpy(), sentiment=sentiment.numpy(),
                        eps0=eps0.numpy(), eps_rest=eps_rest.numpy(), predictions=pred.numpy(),
                        num_constraints=np.asarray([len(constraints)]),
                        **{"param:" + k: v.numpy() for k, v in sd.items()})
    print("wrote", name, "pred", pred.tolist())


WF_TABLE = {"pos": ["good", "nice", "great"], "neg": ["bad", "ugly"], "dog": ["dog", "dogs"], "fire": ["fire"],
            "hydrant": ["hydrant", "hydrants"], "cat": ["cat"], "red": ["red", "reddish"], "old": ["old"]}


def _reference_builder(mg, mw=3):
    ref = rh.load_reference()
    d = tempfile.mkdtemp()
    tsv = os.path.join(d, "wf.tsv")
    with open(tsv, "w") as f:
        for k, v in WF_TABLE.items():
            f.write(k + "\t" + ",".join(v) + "\n")
    vocab = rh.make_vocabulary(40)
    vocab = ref.add_constraint_words_to_vocabulary(vocab, tsv)
    return vocab, ref.FiniteStateMachineBuilder(vocab, tsv, None, max_given_constraints=mg, max_words_per_constraint=mw)


def gen_fsm_build(name):
    """FiniteStateMachineBuilder.build of the reference (constraints.py:329-478) on single-word, multi-word, repeated and
    fewer-than-maximum constraint lists: trimmed dense tensors, state counts and constraint2states."""
    cases = [(3, ["pos", "neg", "dog"]), (3, ["dog", "pos", "fire hydrant"]), (3, ["pos", "pos", "pos"]),
             (3, ["fire hydrant", "fire hydrant"]), (3, ["cat"]), (3, []), (2, ["neg", "dog"]), (1, ["cat"]),
             (3, ["red", "dog", "old"])]
    vocab, _ = _reference_builder(3)
    word_ids = {w: vocab.get_token_index(w) for ws in WF_TABLE.values() for w in ws}
    out = {"wordforms": json.dumps(WF_TABLE), "word_ids": json.dumps(word_ids), "vocab_size": vocab.get_vocab_size(),
           "cases": json.dumps(cases)}
    for i, (mg, cons) in enumerate(cases):
        _, builder = _reference_builder(mg)
        fsm, nstates, c2s = builder.build(cons)
        out[f"fsm{i}"] = fsm[:nstates, :nstates].numpy()
        out[f"nstates{i}"] = nstates
        out[f"c2s{i}"] = json.dumps(c2s)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, [int(out[f"nstates{i}"]) for i in range(len(cases))])


def gen_select_attributes(name, seed=31):
    """select_best_beam_with_constraints with cbs_simple=False (decoding.py:87-138): object / attribute candidates as the
    evaluation dataset hands them over (datasets.py:528-580), constraint2states from the reference's FSM builder."""
    ref = rh.load_reference()
    _, builder = _reference_builder(3)
    cases = [
        ([["dog", ["red"]], ["cat", []]], ["dog", "red", "cat"]),
        ([["dog", ["red", "old"]]], ["dog", "red", "old"]),
        ([["dog", []], ["cat", []]], ["dog", "cat"]),
        ([["cat", ["old"]], ["dog", ["red"]]], ["cat", "old", "dog"]),      # "red" is not among the FSM constraints of this image
        ([["dog", []]], ["dog"]),
    ]
    g = torch.Generator().manual_seed(seed)
    out = {"cases": json.dumps([c for c, _ in cases]), "fsm_inputs": json.dumps([f for _, f in cases])}
    for i, (cands, fsm_input) in enumerate(cases):
        _, nstates, c2s = builder.build(fsm_input)
        for o in cands:                              # the selection indexes constraint2states by every listed name
            for a in [o[0]] + o[1]:
                c2s.setdefault(a, [])
        S, K, steps = nstates, 3, 9
        beams = torch.randint(2, 40, (1, S, K, steps), generator=g)
        logp = -torch.rand(1, S, K, generator=g) * 10
        logp, _ = logp.sort(dim=2, descending=True)
        for min_sat in (1, 2):
            try:
                best, _ = ref.select_best_beam_with_constraints(beams, logp, torch.tensor([len(fsm_input)]), [cands], [c2s],
                                                                min_sat, False)
                out[f"best{i}_{min_sat}"] = best.numpy()
            except (RuntimeError, IndexError, ValueError):       # no valid state: the reference fails on the empty arg max
                out[f"best{i}_{min_sat}"] = np.zeros((0,), dtype=np.int64)
        out[f"beams{i}"] = beams.numpy()
        out[f"logp{i}"] = logp.numpy()
        out[f"c2s{i}"] = json.dumps(c2s)
        out[f"nc{i}"] = len(fsm_input)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, {k: v.shape for k, v in out.items() if k.startswith("best")})


def main():
    os.makedirs(OUT, exist_ok=True)
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == "f2":        # only the FSM-builder / attribute-selection fixtures (round 2)
        gen_fsm_build("fsm_build")
        return gen_select_attributes("select_attributes")
    if len(sys.argv) > 1 and sys.argv[1] == "sv2":       # only the SENTIMENT_VAE = 2 fixtures (added in round 2)
        return main_sv2()
    main_sv2()
    gen_fsm_build("fsm_build")
    gen_select_attributes("select_attributes")
    gen_train("train_tied_sv1", E=600, sv=1)
    gen_train("train_tied300_sv0", E=300, sv=0, seed=7)
    gen_train("train_untied_sv1", E=40, sv=1, seed=11)
    gen_train("train_tied_simple", E=600, sv=1, simple=True, seed=13, prior_std=0.8)
    gen_decode_step("decode_step_tied")
    gen_cbs("cbs_s8_k5", [["pos", "neg", "dog"]], mg=3, K=5, seed=1)
    gen_cbs("cbs_s8_k5_repeat", [["pos", "pos", "pos"]], mg=3, K=5, seed=2, min_sat=2)
    gen_cbs("cbs_s4_k3_b2", [["neg", "neg"], ["pos", "dog"]], mg=2, K=3, B=2, seed=3)
    gen_cbs("cbs_s1_greedy", [[]], mg=0, K=1, seed=4, end_bias_from=6)
    gen_cbs("cbs_s12_multiword", [["dog", "pos", "fire hydrant"]], mg=3, K=5, seed=5, end_bias_from=12)
    gen_cbs("cbs_s8_k5_ninf", [["pos", "neg"]], mg=3, K=5, seed=6, ninf=(5, 9, 17), min_sat=1)
    gen_cbs("cbs_s2_k4_pfull", [["cat"]], mg=1, K=4, P_from_half=False, seed=8)
    gen_beam("beam_k5_p2_b3", K=5, P=2, B=3, seed=9)
    gen_beam("beam_k3_pfull_early", K=3, P=None, B=2, seed=10, end_bias_from=5)
    gen_beam("beam_k1_greedy", K=1, P=None, B=4, seed=12, end_bias_from=8)
    gen_decode_e2e("decode_e2e_cbs_k5", K=5, constraints=["pos", "dog"], mg=3)
    gen_decode_e2e("decode_e2e_greedy", K=1, constraints=[], mg=0, seed=6)


def main_sv2():
    gen_train_sv2("train_tied_sv2_glove", E=600, le="glove", Z=150, seed=17)
    gen_train_sv2("train_untied_sv2_swn", E=40, le="senti_word_net", Z=16, seed=19, prior_std=0.9)
    gen_decode_sv2("decode_greedy_sv2_swn", le="senti_word_net", Z=16)
    gen_decode_sv2("decode_greedy_sv2_glove", le="glove", Z=150, seed=29)


if __name__ == "__main__":
    main()
