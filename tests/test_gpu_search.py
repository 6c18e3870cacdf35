"""GPU: beam / CBS selection kernels, token-exact against the oracle and the reference's golden outputs
when fed identical log-probabilities (replayed step function)."""
import pytest
import torch

import sscvae
from conftest import load_golden
from oracle import search_oracle as so
from oracle.gen_golden import make_replay_step

pytestmark = pytest.mark.gpu

CBS = ["cbs_s8_k5", "cbs_s8_k5_repeat", "cbs_s4_k3_b2", "cbs_s1_greedy", "cbs_s12_multiword", "cbs_s8_k5_ninf",
       "cbs_s2_k4_pfull"]


def cuda_replay_step(tables, ninf):
    """Same pure function as oracle.gen_golden.make_replay_step, evaluated on the CPU so the CUDA search
    is fed bit-identical log-probs, then moved to the device."""
    cpu_step = make_replay_step(tables, ninf, False)

    def step(last, state):
        st = None if state is None else {k: v.cpu() for k, v in state.items()}
        logp, new_state = cpu_step(last.cpu(), st)
        return logp.cuda(), {k: v.cuda() for k, v in new_state.items()}
    return step


@pytest.mark.parametrize("name", CBS)
def test_cbs_token_exact(name):
    g = load_golden(name)
    B = g["fsm"].shape[0]
    K, P = g["K"], g["P"] or None
    cbs = sscvae.ConstrainedBeamSearch(g["end_index"], max_steps=g["max_steps"], beam_size=K, per_node_beam_size=P)
    preds, scores = cbs.search(torch.ones(B, dtype=torch.long, device="cuda"), None,
                               cuda_replay_step(g["tables"], g["ninf"].tolist()), g["fsm"].cuda())
    preds, scores = preds.cpu(), scores.cpu()
    # (1) the oracle, including junk beams and their defined tie-break: everything bit-exact
    op, os_ = so.cbs_search(torch.ones(B, dtype=torch.long), make_replay_step(g["tables"], g["ninf"].tolist(), False),
                            g["fsm"], K, P, g["end_index"], g["max_steps"])
    assert preds.shape == op.shape
    assert torch.equal(preds, op)
    assert torch.equal(scores, os_)
    # (2) the reference's own output: finite-score beams
    fin = g["scores"] > -1e19
    assert torch.equal(preds[fin], g["predictions"][fin])
    assert torch.equal(scores[fin], g["scores"][fin])
    best, _ = sscvae.select_best_beam_with_constraints(preds, scores, g["num_constraints"], None, None,
                                                       g["min_constraints_to_satisfy"], True)
    assert torch.equal(best, g["best"])


@pytest.mark.parametrize("name", ["beam_k5_p2_b3", "beam_k3_pfull_early", "beam_k1_greedy"])
def test_plain_beam_token_exact(name):
    g = load_golden(name)
    B = g["predictions"].shape[0]
    bs = sscvae.BeamSearch(g["end_index"], max_steps=g["max_steps"], beam_size=g["K"], per_node_beam_size=g["P"])
    preds, scores = bs.search(torch.ones(B, dtype=torch.long, device="cuda"), None, cuda_replay_step(g["tables"], []))
    assert torch.equal(preds.cpu(), g["predictions"])
    assert torch.equal(scores.cpu(), g["scores"])
    assert torch.equal(sscvae.select_best_beam(preds, scores).cpu(), g["best"])


def test_large_vocab_random_fsm_exact():
    """V = 11 442 (the reference vocabulary with constraint words), S = 8, K = 5, B = 3,
    non-deterministic random FSM rows (a word may lead to several or no states)."""
    torch.manual_seed(0)
    B, S, K, P, V, steps = 3, 8, 5, 2, 11442, 6
    fsm = (torch.rand(B, S, S, V) < 0.2).to(torch.uint8)
    fsm[:, :, :, 1] = torch.eye(S, dtype=torch.uint8)       # boundary keeps the state
    tables = torch.randn(steps, 64, V)

    def mk(cuda):
        ctr = {"t": 0}

        def step(last, state):
            t = ctr["t"]; ctr["t"] += 1
            lp = torch.log_softmax(tables[t][last.cpu() % 64], dim=1)
            return (lp.cuda() if cuda else lp), {}
        return step
    cbs = sscvae.ConstrainedBeamSearch(1, max_steps=steps, beam_size=K, per_node_beam_size=P)
    preds, scores = cbs.search(torch.ones(B, dtype=torch.long, device="cuda"), None, mk(True), fsm.cuda())
    op, os_ = so.cbs_search(torch.ones(B, dtype=torch.long), mk(False), fsm, K, P, 1, steps)
    assert torch.equal(preds.cpu(), op) and torch.equal(scores.cpu(), os_)


@pytest.mark.parametrize("ninf", [False, True])
def test_block_kernel_random_fsm_exact(ninf):
    """V % 4 == 0 selects the CTA-per-row kernel (row staged in shared memory, 16-byte loads, disallowed to-states
    skipped once they provably cannot enter a list). Random non-deterministic FSM; with `ninf` a fifth of the
    log-probs is -inf, which is BELOW the -1e20 of disallowed words and forces the general path."""
    torch.manual_seed(3)
    B, S, K, P, V, steps = 3, 8, 5, 2, 11444, 6
    fsm = (torch.rand(B, S, S, V) < 0.2).to(torch.uint8)
    fsm[:, :, :, 1] = torch.eye(S, dtype=torch.uint8)
    tables = torch.randn(steps, 64, V)
    drop = torch.rand(steps, 64, V) < 0.2

    def mk(cuda):
        ctr = {"t": 0}

        def step(last, state):
            t = ctr["t"]; ctr["t"] += 1
            lp = torch.log_softmax(tables[t][last.cpu() % 64], dim=1)
            if ninf:
                lp = lp.masked_fill(drop[t][last.cpu() % 64], float("-inf"))
            return (lp.cuda() if cuda else lp), {}
        return step
    cbs = sscvae.ConstrainedBeamSearch(1, max_steps=steps, beam_size=K, per_node_beam_size=P)
    preds, scores = cbs.search(torch.ones(B, dtype=torch.long, device="cuda"), None, mk(True), fsm.cuda())
    op, os_ = so.cbs_search(torch.ones(B, dtype=torch.long), mk(False), fsm, K, P, 1, steps)
    assert torch.equal(preds.cpu(), op) and torch.equal(scores.cpu(), os_)


def test_fused_log_softmax_mode_matches_normalized_mode():
    """normalized=0 (raw logits in, log-softmax fused into the selection) picks the same tokens."""
    import ctypes as C
    from sscvae import _lib
    torch.manual_seed(1)
    B, S, K, V = 4, 1, 5, 10000
    logits = torch.randn(B, V, device="cuda") * 3
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    outs = []
    for normalized, x in ((1, torch.log_softmax(logits, 1)), (0, logits)):
        tok = torch.zeros(B, S, K, dtype=torch.int32, device="cuda")
        sc = torch.zeros(B, S, K, device="cuda")
        _lib.check(_lib.lib().sscvae_search_first_step(_lib.ptr(x.contiguous()), B, S, K, V, None, normalized,
                                                       _lib.ptr(tok), _lib.ptr(sc), s))
        outs.append((tok.cpu(), sc.cpu()))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.allclose(outs[0][1], outs[1][1], atol=1e-5)
    ref = torch.topk(torch.log_softmax(logits, 1), K).indices.cpu().int()
    assert torch.equal(outs[0][0][:, 0], ref)
