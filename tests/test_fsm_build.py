"""constraint state machines built from word ids, and the object / attribute best-beam rule.

Fixtures (oracle/gen_golden.py f2) come from the reference's own FiniteStateMachineBuilder.build
(updown-baseline/updown/utils/constraints.py:329-478) and select_best_beam_with_constraints(cbs_simple=False)
(updown-baseline/updown/utils/decoding.py:87-138).
CPU: the product builder's connection list (replayed here in numpy), state counts and constraint2states, the oracle
builder, and the valid-state rule. GPU: `sscvae_fsm_build` bit-exact against `sscvae_fsm_pack` of the reference's dense
tensors for a ragged batch; decoding from the device-built table equals decoding from the dense tensor; the masked
selection kernel.
"""
import json
import os

import numpy as np
import pytest
import torch

import sscvae
from conftest import GOLDEN
from oracle import fsm_oracle as fo


def _fixture():
    z = np.load(os.path.join(GOLDEN, "fsm_build.npz"))
    return z, json.loads(str(z["cases"])), json.loads(str(z["wordforms"])), json.loads(str(z["word_ids"])), int(z["vocab_size"])


class _Vocab:
    def __init__(self, ids, size):
        self._ids, self._size = ids, size

    def get_token_index(self, w, namespace="tokens"):
        return self._ids[w]

    def get_vocab_size(self, namespace="tokens"):
        return self._size


def _replay(prog):
    """dense (S,S,V) tensor from a connection list, written straight from the semantics in include/sscvae.h"""
    S, V = prog.num_states, prog.vocab_size
    T = max([S] + [max(c[:3]) + 1 for c in prog.connections])      # built untrimmed, trimmed at the end
    fsm = np.zeros((T, T, V), dtype=np.uint8)
    for s in range(prog.num_main_states):
        fsm[s, s, :] = 1
    for frm, to, reset, ids in prog.connections:
        other = np.ones(V, dtype=bool)
        other[ids] = False
        fsm[frm, to, ids] = 1
        fsm[frm, frm, :] = 0
        fsm[frm, reset, other] = 1
        fsm[frm, reset, ids] = 0
    return fsm[:S, :S]


def test_builder_matches_the_reference_builder():
    z, cases, wf, ids, V = _fixture()
    for i, (mg, cons) in enumerate(cases):
        b = sscvae.FiniteStateMachineBuilder(_Vocab(ids, V), None, None, max_given_constraints=mg, wordforms=wf)
        prog, nstates, c2s = b.build(cons)
        assert nstates == int(z[f"nstates{i}"]) == prog.num_states, (i, cons)
        assert c2s == json.loads(str(z[f"c2s{i}"])), (i, cons)
        assert np.array_equal(_replay(prog), z[f"fsm{i}"]), (i, cons)
        # the oracle restatement, on the same fixture
        f2, n2, c2 = fo.build_fsm(cons, wf, lambda w: ids[w], V, max_given_constraints=mg)
        assert n2 == nstates and c2 == c2s and np.array_equal(fo.trim_fsm(f2, n2), z[f"fsm{i}"])
    # word forms can also come from the TSV files the reference reads
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "wf.tsv")
        with open(p, "w") as f:
            for k, v in wf.items():
                f.write(k + "\t" + ",".join(v) + "\n")
        b = sscvae.FiniteStateMachineBuilder(_Vocab(ids, V), p, None)
        prog, _, _ = b.build(cases[1][1])
        assert np.array_equal(_replay(prog), z["fsm1"])


def _select_fixture():
    z = np.load(os.path.join(GOLDEN, "select_attributes.npz"))
    return z, json.loads(str(z["cases"]))


def test_attribute_selection_matches_the_reference():
    z, cases = _select_fixture()
    seen_invalid = False
    for i, cands in enumerate(cases):
        c2s = json.loads(str(z[f"c2s{i}"]))
        beams, logp, nc = torch.from_numpy(z[f"beams{i}"]), torch.from_numpy(z[f"logp{i}"]), int(z[f"nc{i}"])
        for min_sat in (1, 2):
            want = z[f"best{i}_{min_sat}"]
            states = sscvae.valid_states_with_attributes(nc, cands, c2s, min_sat)
            if want.size == 0:                       # the reference fails on an empty arg max
                assert states == []
                seen_invalid = True
                continue
            best, _ = sscvae.select_best_beam_with_constraints(beams, logp, torch.tensor([nc]), [cands], [c2s], min_sat, False)
            assert np.array_equal(best.numpy(), want), (i, min_sat)
    assert seen_invalid


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_device_built_table_equals_the_packed_reference_tensors():
    from sscvae import _lib
    import ctypes as C
    z, cases, wf, ids, V = _fixture()
    progs, dense = [], []
    for i, (mg, cons) in enumerate(cases):
        if mg != 3:
            continue
        b = sscvae.FiniteStateMachineBuilder(_Vocab(ids, V), None, None, max_given_constraints=mg, wordforms=wf)
        progs.append(b.build(cons)[0])
        dense.append(torch.from_numpy(z[f"fsm{i}"]))
    assert len({p.num_states for p in progs}) > 2        # a ragged batch
    bits = sscvae.build_fsm_bits(progs, "cuda").bits
    padded, _ = sscvae.pad_fsm_batch(dense)
    B, S = padded.shape[:2]
    assert bits.shape == (B, S, V)
    want = torch.empty(B, S, V, dtype=torch.int32, device="cuda")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(_lib.lib().sscvae_fsm_pack(_lib.ptr(padded.cuda().contiguous()), B, S, V, _lib.ptr(want), stream))
    torch.cuda.synchronize()
    assert torch.equal(bits, want)
    # smaller main-state counts too
    for i, (mg, cons) in enumerate(cases):
        if mg == 3:
            continue
        b = sscvae.FiniteStateMachineBuilder(_Vocab(ids, V), None, None, max_given_constraints=mg, wordforms=wf)
        got = sscvae.build_fsm_bits([b.build(cons)[0]], "cuda").bits
        d = torch.from_numpy(z[f"fsm{i}"])[None].cuda().contiguous()
        S1 = d.shape[1]
        want = torch.empty(1, S1, V, dtype=torch.int32, device="cuda")
        _lib.check(_lib.lib().sscvae_fsm_pack(_lib.ptr(d), 1, S1, V, _lib.ptr(want), stream))
        assert torch.equal(got, want)


def _decode_setup(cbs_simple=True):
    from helpers import StubVocabulary
    cfg = dict(vocab_size=60, image_feature_size=64, embedding_size=600, hidden_size=40, attention_projection_size=24,
               z_space=12, sentiment_vae=1, simple_vae=False, max_caption_length=20, prior_std=0.7, senti_prior_multip=0.5)
    torch.manual_seed(3)
    m = sscvae.UpDownCaptioner(StubVocabulary(60), 64, 600, 40, 24, beam_size=3, use_cbs=True, z_space=12, prior_std=0.7,
                               latent_embedding="glove", sentiment_vae=1, senti_prior_multip=0.5, cbs_simple=cbs_simple,
                               min_constraints_to_satisfy=2).cuda()
    m.eval()
    wf = {"dog": ["w3", "w4"], "red": ["w7"], "cat": ["w9"], "fire": ["w11"], "hydrant": ["w12", "w13"]}
    b = sscvae.FiniteStateMachineBuilder(_Vocab({f"w{i}": i + 2 for i in range(58)}, 60), None, None, wordforms=wf)
    return cfg, m, b


@pytest.mark.gpu
def test_decode_from_the_device_built_table_equals_decode_from_the_dense_tensor():
    cfg, m, b = _decode_setup()
    cons = [["dog", "red", "cat"], ["fire hydrant", "dog"], ["cat"]]
    built = [b.build(c) for c in cons]
    progs = [p for p, _, _ in built]
    B, K, Z = 3, 3, 12
    gen = torch.Generator().manual_seed(5)
    feats = torch.rand(B, 6, 64, generator=gen)
    sent = torch.tensor([[1.0], [-1.0], [0.0]])
    nc = torch.tensor([len(c) for c in cons])
    bits = sscvae.build_fsm_bits(progs, "cuda")
    S = bits.shape[1]
    eps = torch.randn(20, B * S * K, Z, generator=gen)
    m._eps_override = eps.cuda()
    a = m(feats.cuda(), None, None, fsm=bits, num_constraints=nc, sentiment=sent.cuda())["predictions"].cpu()
    sa = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in m.last_search.items()}
    # the same machines as dense tensors (oracle builder = the reference's, pinned above), zero-padded to S states
    vocab_ids = {f"w{i}": i + 2 for i in range(58)}
    dense = []
    for c in cons:
        f, n, _ = fo.build_fsm(c, b._wordforms, lambda w: vocab_ids[w], 60)
        dense.append(torch.from_numpy(fo.trim_fsm(f, n)))
    padded, _ = sscvae.pad_fsm_batch(dense)
    assert padded.shape[1] == S
    bb = m(feats.cuda(), None, None, fsm=padded.cuda(), num_constraints=nc, sentiment=sent.cuda())["predictions"].cpu()
    assert torch.equal(a, bb)
    assert torch.equal(sa["predictions"], m.last_search["predictions"].cpu())
    assert torch.equal(sa["log_probs"], m.last_search["log_probs"].cpu())


@pytest.mark.gpu
def test_attribute_rule_selects_on_the_device():
    cfg, m, b = _decode_setup(cbs_simple=False)
    fsm_inputs = [["dog", "red", "cat"], ["dog", "cat"]]
    cands = [[["dog", ["red"]], ["cat", []]], [["dog", []], ["cat", []]]]
    built = [b.build(c) for c in fsm_inputs]
    c2s = [x[2] for x in built]
    B, K, Z = 2, 3, 12
    gen = torch.Generator().manual_seed(6)
    feats = torch.rand(B, 6, 64, generator=gen)
    sent = torch.tensor([[1.0], [0.0]])
    nc = torch.tensor([len(c) for c in fsm_inputs])
    bits = sscvae.build_fsm_bits([x[0] for x in built], "cuda")
    m._eps_override = torch.randn(20, B * bits.shape[1] * K, Z, generator=gen).cuda()
    for min_sat in (1, 2):
        m._min_constraints_to_satisfy = min_sat
        got = m(feats.cuda(), None, None, fsm=bits, num_constraints=nc, sentiment=sent.cuda(), constraints=cands,
                constraint2states=c2s)["predictions"].cpu()
        preds, logp = m.last_search["predictions"].cpu(), m.last_search["log_probs"].cpu()
        want, _ = sscvae.select_best_beam_with_constraints(preds, logp, nc, cands, c2s, min_sat, False)
        assert torch.equal(got, want)
        simple, _ = sscvae.select_best_beam_with_constraints(preds, logp, nc, None, None, min_sat, True)
        assert got.shape == simple.shape
