"""Beam / constrained beam search objects with the reference's interface, running the selection on
the sm_100a search kernels (libsscvae_b200.so).

Mirrors `ConstrainedBeamSearch` (updown-baseline/updown/modules/cbs.py:20-277), allennlp's
`BeamSearch` (in-tree text: var_updown/var_updown/modules/beam_search.py:434-766, deterministic
sampler) and `select_best_beam[_with_constraints]` (updown-baseline/updown/utils/decoding.py:10-138,
cbs_simple). `search(start_predictions, start_state, step, fsm)` takes any step function returning
log-probabilities — the captioner itself uses the fully fused `sscvae_decode` instead; these classes
exist so that callers (and the replay parity tests) can drive the search kernels with their own
log-probs. Ties are broken by lowest index (the oracle's definition; torch.topk leaves it open).
"""
import ctypes as C
from typing import Callable, Dict, Optional, Tuple

import torch

from . import _lib


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _run_search(start_predictions, start_state, step, fsm, end_index, max_steps, K, P, normalized=True):
    L = _lib.lib()
    dev = start_predictions.device
    if not start_predictions.is_cuda:
        raise RuntimeError("sscvae search runs only on CUDA tensors; there is no CPU fallback")
    B = start_predictions.shape[0]
    out = step(start_predictions, start_state)
    logp, state = out[0], out[1]
    logp = logp.contiguous().float()
    V = logp.shape[1]
    if fsm is not None:
        fsm = fsm.to(dev).to(torch.uint8).contiguous()
        S = fsm.shape[1]
        bits = torch.empty(B, S, V, dtype=torch.int32, device=dev)
        _lib.check(L.sscvae_fsm_pack(_lib.ptr(fsm), B, S, V, _lib.ptr(bits), _stream(dev)))
    else:
        S, bits = 1, None
    R = B * S * K
    tok_hist = torch.zeros(max_steps, R, dtype=torch.int32, device=dev)
    bp_hist = torch.zeros(max_steps, R, dtype=torch.int32, device=dev)
    sc_hist = torch.zeros(max_steps, R, dtype=torch.float32, device=dev)
    _lib.check(L.sscvae_search_first_step(_lib.ptr(logp), B, S, K, V, _lib.ptr(bits), int(normalized),
                                          _lib.ptr(tok_hist[0]), _lib.ptr(sc_hist[0]), _stream(dev)))
    # replicate every state tensor to (B*S*K, *) (cbs.py:10-17,152-155)
    rowmap = torch.arange(R, device=dev) // (S * K)
    state = {k: v.index_select(0, rowmap) for k, v in state.items()}
    nscratch = L.sscvae_search_scratch_bytes(B, S, K, P)
    scratch = torch.empty(nscratch, dtype=torch.uint8, device=dev)
    base = (torch.arange(R, device=dev) // (S * K)) * (S * K)
    steps_run = 1
    for t in range(1, max_steps):
        last = tok_hist[t - 1]
        if bool((last == end_index).all()):                          # cbs.py:167
            break
        out = step(last.long(), state)
        logp, state = out[0].contiguous().float(), out[1]
        _lib.check(L.sscvae_search_step(_lib.ptr(logp), B, S, K, P, V, _lib.ptr(bits), int(normalized), end_index,
                                        _lib.ptr(last), _lib.ptr(sc_hist[t - 1]), _lib.ptr(scratch), nscratch,
                                        _lib.ptr(tok_hist[t]), _lib.ptr(bp_hist[t]), _lib.ptr(sc_hist[t]), _stream(dev)))
        src = base + bp_hist[t].long()
        state = {k: v.index_select(0, src) for k, v in state.items()}   # cbs.py:236-250
        steps_run += 1
    preds = torch.empty(B, S, K, steps_run, dtype=torch.long, device=dev)
    scores = torch.empty(B, S, K, dtype=torch.float32, device=dev)
    best = torch.empty(B, steps_run, dtype=torch.long, device=dev)
    n_steps = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(L.sscvae_search_finish(_lib.ptr(tok_hist), _lib.ptr(bp_hist), _lib.ptr(sc_hist), steps_run, B, S, K,
                                      end_index, None, 0, _lib.ptr(preds), _lib.ptr(scores), _lib.ptr(best),
                                      _lib.ptr(n_steps), _stream(dev)))
    # tok_hist rows are (max_steps, R) with stride R, and only the first steps_run are used
    return preds, scores


class ConstrainedBeamSearch(object):
    def __init__(self, end_index: int, max_steps: int = 20, beam_size: int = 5, per_node_beam_size: Optional[int] = None):
        self._end_index = end_index
        self.max_steps = max_steps
        self.beam_size = beam_size
        self.per_node_beam_size = per_node_beam_size or self.beam_size

    def search(self, start_predictions, start_state, step: Callable, fsm: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> predictions (B,S,K,steps) int64, log_probabilities (B,S,K)."""
        return _run_search(start_predictions, start_state, step, fsm, self._end_index, self.max_steps,
                           self.beam_size, self.per_node_beam_size)


class BeamSearch(object):
    def __init__(self, end_index: int, max_steps: int = 50, beam_size: int = 10, per_node_beam_size: int = None):
        if not max_steps > 0:
            raise ValueError("max_steps must be positive")
        if not beam_size > 0:
            raise ValueError("beam_size must be positive")
        if per_node_beam_size is not None and not per_node_beam_size > 0:
            raise ValueError("per_node_beam_size must be positive")
        self._end_index = end_index
        self.max_steps = max_steps
        self.beam_size = beam_size
        self.per_node_beam_size = per_node_beam_size or beam_size

    def search(self, start_predictions, start_state, step: Callable) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> predictions (B,K,steps) int64, log_probabilities (B,K). Plain beam = the S=1, no-FSM case."""
        preds, scores = _run_search(start_predictions, start_state, step, None, self._end_index, self.max_steps,
                                    self.beam_size, self.per_node_beam_size)
        return preds[:, 0], scores[:, 0]


def select_best_beam(beams: torch.Tensor, beam_log_probabilities: torch.Tensor) -> torch.Tensor:
    return beams[:, 0, :]


def select_best_beam_with_constraints(beams, beam_log_probabilities, given_constraints, constraints=None,
                                      constraint2states=None, min_constraints_to_satisfy: int = 2, cbs_simple=True):
    """Best first-beam among the valid end states (decoding.py:82-138). cbs_simple: the states whose bit count
    satisfies min(#constraints, min_constraints_to_satisfy); otherwise the object / attribute rule
    (`fsm.valid_states_with_attributes`, decoding.py:87-123) over `constraints[i]` and `constraint2states[i]`."""
    from .fsm import valid_states_with_attributes
    B = beams.shape[0]
    best, valid = [], []
    for i in range(B):
        nc = int(given_constraints[i])
        if cbs_simple:
            need = min(nc, min_constraints_to_satisfy)
            states = [s for s in range(2 ** nc) if bin(s).count("1") >= need]
        else:
            states = valid_states_with_attributes(nc, constraints[i], constraint2states[i], min_constraints_to_satisfy)
        vb = beams[i, states, 0, :]
        vl = beam_log_probabilities[i, states, 0]
        best.append(vb[torch.argmax(vl)])
        valid.append(vb)
    # the reference stacks the per-image valid beams (it decodes one image at a time); images of a batch can have
    # different numbers of valid states, then the list is returned as it is
    same = all(v.shape == valid[0].shape for v in valid)
    return torch.stack(best).long(), (torch.stack(valid) if same else valid)


def pad_fsm_batch(fsms, num_constraints=None):
    """Batch per-image finite-state machines with DIFFERENT state counts.

    The reference builds one trimmed `(S_i, S_i, V)` uint8 FSM per image (constraints.py:329-478, trimmed at
    datasets.py:611-613) and its evaluation collate can therefore only stack a batch of ONE image
    (datasets.py:604-620). Padding with zeros up to the largest state count is exact: a padded state has no incoming
    transition, so its beams keep the masked score forever - the same thing that happens to the unreachable main
    states the reference itself carries when an image has fewer constraints than `max_given_constraints` - and the
    best-beam selection only looks at the `2**num_constraints` main states of each image (decoding.py:82-134).

    fsms: sequence of (S_i, S_i, V) uint8 / bool tensors; num_constraints: optional sequence of ints.
    Returns (fsm (B, S, S, V) uint8, num_constraints (B,) int64 or None)."""
    if len(fsms) == 0:
        raise ValueError("empty FSM batch")
    V = fsms[0].shape[-1]
    S = max(int(f.shape[0]) for f in fsms)
    out = torch.zeros(len(fsms), S, S, V, dtype=torch.uint8, device=fsms[0].device)
    for i, f in enumerate(fsms):
        if f.dim() != 3 or f.shape[0] != f.shape[1] or f.shape[2] != V:
            raise ValueError(f"FSM {i} has shape {tuple(f.shape)}, expected (S, S, {V})")
        si = f.shape[0]
        out[i, :si, :si] = f.to(torch.uint8)
    nc = None
    if num_constraints is not None:
        if len(num_constraints) != len(fsms):
            raise ValueError("one constraint count per FSM")
        nc = torch.as_tensor([int(n) for n in num_constraints], dtype=torch.long, device=fsms[0].device)
        for i, f in enumerate(fsms):
            if 2 ** int(nc[i]) > f.shape[0]:
                raise ValueError(f"FSM {i} has {f.shape[0]} states, fewer than the 2**{int(nc[i])} main states")
    return out, nc
