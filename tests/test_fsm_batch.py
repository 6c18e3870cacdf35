"""CPU: batching FSMs of different state counts by zero padding (sscvae.pad_fsm_batch) leaves the
reference's constrained beam search unchanged - checked with the CBS oracle (oracle/search_oracle.py, pinned to
updown-baseline/updown/modules/cbs.py by tests/golden) on a deterministic synthetic step function."""
import numpy as np
import pytest
import torch

from oracle import fsm_oracle as fo
from oracle import search_oracle as so


def _pad():
    import sscvae
    return sscvae.pad_fsm_batch


def _step_fn(V, seed):
    """log-probs that depend on the previous token and the step, the same for every row with the same history."""
    g = torch.Generator().manual_seed(seed)
    table = torch.log_softmax(torch.randn(6, V, V, generator=g) * 2.0, dim=-1)      # [step % 6][last token]

    def step(last, state):
        t = 0 if state is None else int(state["t"][0])
        logp = table[t % 6][last]
        return logp, {"t": torch.full((last.shape[0],), t + 1, dtype=torch.long)}
    return step


@pytest.mark.parametrize("K,P", [(5, 2), (3, 3)])
def test_padded_batch_equals_one_image_at_a_time(K, P):
    V, L = 60, 12
    cons = [[[5, 6], [9]], [[11]], [[7], [8, 10], [20, 21, 22]]]                 # 2, 1 and 3 single-word constraints
    fsms = [torch.from_numpy(fo.single_word_fsm(c, V)) for c in cons]             # 4, 2 and 8 states
    ncs = [len(c) for c in cons]
    fsm, nc = _pad()(fsms, ncs)
    assert fsm.shape == (3, 8, 8, V) and fsm.dtype == torch.uint8 and nc.tolist() == ncs
    step = _step_fn(V, 3)
    start = torch.ones(3, dtype=torch.long)
    allp, sc = so.cbs_search(start, step, fsm, K, P, end_index=1, max_steps=L)
    best, _ = so.select_best_beam_with_constraints(allp, sc, nc, 2)
    for b in range(3):
        one_p, one_s = so.cbs_search(start[:1], step, fsms[b][None], K, P, end_index=1, max_steps=L)
        one_best, _ = so.select_best_beam_with_constraints(one_p, one_s, nc[b:b + 1], 2)
        n = min(one_best.shape[1], best.shape[1])
        assert torch.equal(one_best[0, :n], best[b, :n]), b
        assert bool((best[b, n:] == 1).all()) and bool((one_best[0, n:] == 1).all())
        Sb = fsms[b].shape[0]
        fin = one_s[0] > -1e19
        assert torch.equal(sc[b, :Sb][fin], one_s[0][fin])                         # real states: same scores
        assert bool((sc[b, Sb:] < -1e19).all())                                    # padded states are never entered


def test_pad_fsm_batch_rejects_bad_input():
    pad = _pad()
    with pytest.raises(ValueError):
        pad([])
    with pytest.raises(ValueError):
        pad([torch.zeros(2, 3, 10, dtype=torch.uint8)])
    with pytest.raises(ValueError):
        pad([torch.zeros(2, 2, 10, dtype=torch.uint8)], [2])                        # 2 constraints need 4 main states
    f, nc = pad([torch.ones(2, 2, 10, dtype=torch.uint8), torch.ones(4, 4, 10, dtype=torch.bool)])
    assert nc is None and f.shape == (2, 4, 4, 10) and int(f[0].sum()) == 40 and int(f[1].sum()) == 160
