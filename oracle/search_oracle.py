"""CPU oracle for the caption search: constrained beam search (CBS), plain beam search and
best-beam selection, restated from the reference.

TEST INFRASTRUCTURE ONLY (see oracle/updown_oracle.py header for who may import this).

Follows, line by line:
  * updown-baseline/updown/modules/cbs.py:59-277  (ConstrainedBeamSearch.search), with the three
    torch>=1.2 fixes (:135/:205 `1 - uint8` -> logical not, :231 `/` -> `//`);
  * var_updown/var_updown/modules/beam_search.py:592-766 (BeamSearch._search with the
    DeterministicSampler :87-99) — the in-tree statement of allennlp's plain beam;
  * updown-baseline/updown/utils/decoding.py:10-138 (select_best_beam[_with_constraints]).

One deliberate definition: `torch.topk` leaves the order among EQUAL scores unspecified; this
oracle (and the CUDA kernel) break ties by LOWEST candidate index (stable descending sort).
Against the live reference only finite-score beams are therefore comparable.
"""
from typing import Callable, Dict, List, Optional, Tuple

import torch

NEG_MASK = -1e20          # cbs.py:205 masks disallowed transitions with -1e20 (not -inf)


def topk_stable(x: torch.Tensor, k: int):
    vals, idx = torch.sort(x, dim=-1, descending=True, stable=True)
    return vals[..., :k].contiguous(), idx[..., :k].contiguous()


def _enlarge(t: torch.Tensor, B: int, S: int, K: int) -> torch.Tensor:
    """cbs.py:10-17."""
    last = t.shape[1:]
    return t.view(B, 1, 1, *last).expand(B, S, K, *last).reshape(-1, *last)


def cbs_first_step(logp0: torch.Tensor, fsm: torch.Tensor, K: int):
    """cbs.py:130-145. logp0 (B,V), fsm (B,S,S,V) -> scores (B,S,K), tokens (B,S,K)."""
    B, S, _, V = fsm.shape
    cand = logp0.view(B, 1, V).expand(B, S, V).masked_fill(~fsm[:, 0].bool(), float("-inf"))
    return topk_stable(cand, K)


def cbs_step(logp: torch.Tensor, last_tokens: torch.Tensor, last_scores: torch.Tensor,
             fsm: torch.Tensor, K: int, P: int, end_index: int):
    """One inner iteration of cbs.py:161-234.
    logp (B*S*K,V); last_tokens (B,S*K); last_scores (B,S,K); fsm (B,S,S,V).
    Returns new_tokens (B,S,K) i64, backpointer (B,S,K) i64 (index into the image's S*K rows),
    new_scores (B,S,K)."""
    B, S, _, V = fsm.shape
    after_end = torch.full((1, V), float("-inf"))
    after_end[:, end_index] = 0.0                                           # :147-150
    is_end = (last_tokens.reshape(-1, 1) == end_index)
    cleaned = torch.where(is_end, after_end, logp).view(B, S, K, V)         # :177-184
    new_tokens = torch.zeros(B, S, K, dtype=torch.long)
    new_idx = torch.zeros(B, S, K, dtype=torch.long)
    new_scores = torch.zeros(B, S, K)
    mask = fsm.bool().view(B, S, S, 1, V)
    for i in range(S):                                                      # :200
        m = cleaned.masked_fill(~mask[:, :, i].expand(B, S, K, V), NEG_MASK)
        tp, tc = topk_stable(m, P)                                          # :207-209
        summed = (tp + last_scores.view(B, S, K, 1)).reshape(B, -1)         # :210-214
        bs, bi = topk_stable(summed, K)                                     # :220
        new_tokens[:, i] = tc.reshape(B, -1).gather(1, bi)                  # :222-226
        new_idx[:, i] = bi
        new_scores[:, i] = bs
    return new_tokens, new_idx // P, new_scores                             # :231


def backtrace(predictions: List[torch.Tensor], backpointers: List[torch.Tensor]) -> torch.Tensor:
    """cbs.py:252-274 / beam_search.py:488-514. predictions[t] (B,W); backpointers[t] (B,W)
    -> (B,W,steps)."""
    rec = [predictions[-1].unsqueeze(2)]
    if backpointers:
        cur = backpointers[-1]
        for t in range(len(predictions) - 2, 0, -1):
            rec.append(predictions[t].gather(1, cur).unsqueeze(2))
            cur = backpointers[t - 1].gather(1, cur)
        rec.append(predictions[0].gather(1, cur).unsqueeze(2))
    return torch.cat(list(reversed(rec)), 2)


StepFn = Callable[[torch.Tensor, Optional[Dict[str, torch.Tensor]]],
                  Tuple[torch.Tensor, Dict[str, torch.Tensor]]]


def cbs_search(start: torch.Tensor, step: StepFn, fsm: torch.Tensor, K: int, P: Optional[int],
               end_index: int, max_steps: int):
    """ConstrainedBeamSearch.search (cbs.py:59-277).
    step(last_predictions (R,), states|None) -> (logp (R,V), states). Returns
    predictions (B,S,K,steps) i64 and log_probs (B,S,K)."""
    P = P or K                                                              # :57
    B, S, _, V = fsm.shape
    logp0, state = step(start, None)                                        # :127
    scores, tok0 = cbs_first_step(logp0, fsm, K)
    predictions = [tok0.view(B, -1)]
    backpointers: List[torch.Tensor] = []
    state = {k: _enlarge(v, B, S, K) for k, v in state.items()}             # :152-155
    for _ in range(max_steps - 1):                                          # :161
        last = predictions[-1].reshape(B * S * K)
        if (last == end_index).all():                                       # :167
            break
        logp, state = step(last, state)                                     # :170
        new_tok, bp, scores = cbs_step(logp, predictions[-1], scores, fsm, K, P, end_index)
        predictions.append(new_tok.view(B, -1))
        bp = bp.view(B, -1)
        backpointers.append(bp)

        def track(t):                                                       # :236-248
            last_dims = t.shape[1:]
            idx = bp.view(B, S * K, *([1] * len(last_dims))).expand(B, S * K, *last_dims)
            return t.reshape(B, S * K, *last_dims).gather(1, idx).reshape(B * S * K, *last_dims)
        state = {k: track(v) for k, v in state.items()}                     # :250
    allp = backtrace(predictions, backpointers)
    return allp.view(B, S, K, -1), scores


def beam_search(start: torch.Tensor, step: StepFn, K: int, P: Optional[int], end_index: int,
                max_steps: int):
    """Plain beam search (vendored beam_search.py:592-766, DeterministicSampler).
    Returns predictions (B,K,steps), log_probs (B,K)."""
    P = P or K
    B = start.shape[0]
    logp0, state = step(start, None)                                        # :616
    V = logp0.shape[1]
    scores, tok0 = topk_stable(logp0, K)                                    # :634-638
    if K == 1 and (tok0 == end_index).all():                                # :640-646
        return tok0.unsqueeze(-1), scores
    predictions = [tok0]
    backpointers: List[torch.Tensor] = []
    after_end = torch.full((1, V), float("-inf"))
    after_end[:, end_index] = 0.0
    state = {k: v.unsqueeze(1).expand(B, K, *v.shape[1:]).reshape(B * K, *v.shape[1:])
             for k, v in state.items()}                                     # :775-799
    for _ in range(max_steps - 1):                                          # :665
        last = predictions[-1].reshape(B * K)
        if (last == end_index).all():                                       # :671
            break
        logp, state = step(last, state)                                     # :676
        cleaned = torch.where(last.unsqueeze(-1) == end_index, after_end, logp)   # :688-692
        tp, tc = topk_stable(cleaned, P)                                    # :695
        summed = (tp + scores.reshape(B * K, 1)).reshape(B, K * P)          # :703-715
        scores, bi = topk_stable(summed, K)                                 # :724-728
        predictions.append(tc.reshape(B, K * P).gather(1, bi))              # :732-736
        bp = bi // P                                                        # :746
        backpointers.append(bp)
        state = {k: v.reshape(B, K, *v.shape[1:])
                 .gather(1, bp.view(B, K, *([1] * (v.dim() - 1))).expand(B, K, *v.shape[1:]))
                 .reshape(B * K, *v.shape[1:]) for k, v in state.items()}   # :801-833
    return backtrace(predictions, backpointers), scores


def select_best_beam(beams: torch.Tensor, log_probs: torch.Tensor) -> torch.Tensor:
    """decoding.py:10-27."""
    return beams[:, 0, :]


def valid_states_simple(num_constraints: int, min_constraints_to_satisfy: int) -> List[int]:
    """decoding.py:82-86 (cbs_simple)."""
    need = min(num_constraints, min_constraints_to_satisfy)
    return [s for s in range(2 ** num_constraints) if bin(s).count("1") >= need]


def select_best_beam_with_constraints(beams, log_probs, num_constraints, min_constraints_to_satisfy=2):
    """decoding.py:30-138, cbs_simple=True branch. beams (B,S,K,steps), log_probs (B,S,K),
    num_constraints (B,) -> best (B,steps), list of valid beams per image."""
    best, valid_all = [], []
    for i in range(beams.shape[0]):
        vs = valid_states_simple(int(num_constraints[i]), min_constraints_to_satisfy)
        vb = beams[i, vs, 0, :]                                             # :128
        vl = log_probs[i, vs, 0]                                            # :130
        best.append(vb[torch.argmax(vl)])                                   # :132-133
        valid_all.append(vb)
    return torch.stack(best).long(), valid_all
