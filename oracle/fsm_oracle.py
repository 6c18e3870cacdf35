"""CPU oracle for the CBS finite-state machine the reference feeds to the search.

TEST INFRASTRUCTURE ONLY. Restates `FiniteStateMachineBuilder.build` / `_add_nth_constraint` /
`_connect` (updown-baseline/updown/utils/constraints.py:329-478) in numpy so that tests and the
benchmark can construct the same `(S,S,V)` uint8 adjacency tensors the reference's data pipeline
produces (trimmed to the used states as updown-baseline/updown/data/datasets.py:611-613 does),
without the reference tree. Pinned against the reference builder in tests/golden (fsm_*.npz).
"""
from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np


def build_fsm(constraints: Sequence[str], wordforms: Dict[str, List[str]],
              token_index: Callable[[str], int], vocab_size: int,
              max_given_constraints: int = 3, max_words_per_constraint: int = 3
              ) -> Tuple[np.ndarray, int, Dict[str, List[int]]]:
    """Returns (fsm (T,T,V) uint8 with T = 2**k * w, next free sub-state index, constraint2states).

    fsm[s1, s2, w] = 1  <=>  emitting word w moves state s1 -> s2 (constraints.py:227-236)."""
    n_main = 2 ** max_given_constraints
    n_total = n_main * max_words_per_constraint
    fsm = np.zeros((n_total, n_total, vocab_size), dtype=np.uint8)
    for s in range(n_main):                       # self loops on main states, every word (:344-349)
        fsm[s, s, :] = 1

    def connect(frm: int, to: int, word: str, reset: int):
        """constraints.py:427-478."""
        ids = [token_index(w) for w in wordforms[word]]
        for i in ids:
            fsm[frm, to, i] = 1
            fsm[frm, frm, i] = 0
        # reset_state is always passed by the caller (:403-412), so this block always runs
        fsm[frm, frm, :] = 0
        fsm[frm, reset, :] = 1
        for i in ids:
            fsm[frm, reset, i] = 0

    constraint2states: Dict[str, List[int]] = {}
    sub = n_main
    seen: Dict[str, List[int]] = {}
    max_valid = 2 ** len(constraints)
    for n0, constraint in enumerate(constraints):             # :354-357
        n = n0 + 1
        words = constraint.split()
        stride = 2 ** (n - 1)
        mains: List[int] = []
        if constraint in seen:                                # repeated constraint (:385-392)
            frm = seen[constraint][-1]
            frm_max = frm + 1
            seen[constraint].append(n)
        else:
            frm, frm_max = 0, n_main
            seen[constraint] = [n]
        while frm < frm_max:                                  # :394-415
            for _ in range(stride):
                wfrom = frm
                for i, word in enumerate(words):
                    if i != len(words) - 1:
                        connect(wfrom, sub, word, frm)
                        wfrom = sub
                        sub += 1
                    else:
                        if frm + stride < max_valid:
                            mains.append(frm + stride)
                        connect(wfrom, frm + stride, word, frm)
                frm += 1
            frm += stride
        constraint2states[constraint] = mains
    return fsm, sub, constraint2states


def trim_fsm(fsm: np.ndarray, next_substate: int) -> np.ndarray:
    """What the evaluation collate hands to the model: `fsm[:num_states, :num_states]` with
    num_states = the builder's next free sub-state index, i.e. all 2**max_given_constraints main
    states plus the used sub-states (updown-baseline/updown/data/datasets.py:611-613)."""
    return np.ascontiguousarray(fsm[:next_substate, :next_substate])


def single_word_fsm(constraint_token_ids: Sequence[Sequence[int]], vocab_size: int) -> np.ndarray:
    """Convenience for synthetic benchmarks: k single-word constraints, each a set of token ids
    (word-forms) -> trimmed (2**k, 2**k, V) FSM, identical to build_fsm+trim_fsm for distinct
    single-word constraints."""
    k = len(constraint_token_ids)
    wf = {f"c{i}": [f"t{j}" for j in ids] for i, ids in enumerate(constraint_token_ids)}
    fsm, sub, _ = build_fsm([f"c{i}" for i in range(k)], wf, lambda w: int(w[1:]), vocab_size,
                            max_given_constraints=k, max_words_per_constraint=1)
    return trim_fsm(fsm, sub)
