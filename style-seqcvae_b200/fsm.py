"""Constraint finite-state machines for constrained beam search, built on the device from word ids.

The reference's `FiniteStateMachineBuilder.build` (updown-baseline/updown/utils/constraints.py:329-478) materialises a
dense `(24, 24, V)` uint8 adjacency tensor per image on the host, the dataset trims it to the used states
(updown-baseline/updown/data/datasets.py:611-613) and the model receives `(B, S, S, V)` bytes (0.64 MB per image at
S = 8, V = 10 000). Here the builder keeps the machine in the form it is specified in - the ordered list of connections
`(from_state, to_state, reset_state, word-form ids)` that the reference's `_connect` (:427-478) would apply - and
`sscvae_fsm_build` expands a batch of such lists directly into the `(B, S, V)` uint32 bit table the search kernels read
(bit i of entry [b, s, w] = word w moves state s to state i). A few dozen ints per image cross the boundary.

`FiniteStateMachineBuilder` mirrors the reference's ctor arguments and `build()` return triple, with an `FsmProgram` in
place of the dense tensor; `build_fsm_bits` batches programs with different state counts.
"""
import csv
import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib


class FsmProgram:
    """One image's machine: `connections` = [(from, to, reset, [token ids]), ...] in application order, `num_main_states`
    states start with a self loop on every word, `num_states` = main states + the sub-states in use. Like the reference's
    builder, a repeated constraint can connect states beyond `num_states`; the table is trimmed to `num_states`
    (updown-baseline/updown/data/datasets.py:611-613), which drops those transitions."""

    def __init__(self, connections, num_main_states: int, num_states: int, vocab_size: int):
        self.connections = connections
        self.num_main_states = num_main_states
        self.num_states = num_states
        self.vocab_size = vocab_size


class FsmBits:
    """A batch of machines as the device bit table: `bits` (B, S, V) int32 on a CUDA device."""

    def __init__(self, bits: torch.Tensor):
        self.bits = bits

    @property
    def shape(self):
        return tuple(self.bits.shape)


class FiniteStateMachineBuilder:
    def __init__(self, vocabulary, wordforms_tsvpath: Optional[str], wordforms_attribs_tsvpath: Optional[str] = None,
                 max_given_constraints: int = 3, max_words_per_constraint: int = 3, use_coco_attributes=False,
                 wordforms: Optional[Dict[str, List[str]]] = None):
        if use_coco_attributes:
            raise NotImplementedError("use_coco_attributes filters the word forms by a module-level selection table of "
                                      "the reference's data pipeline; pass the filtered table as `wordforms`")
        self._vocabulary = vocabulary
        self._max_given_constraints = max_given_constraints
        self._max_words_per_constraint = max_words_per_constraint
        self._num_main_states = 2 ** max_given_constraints
        self._num_total_states = self._num_main_states * max_words_per_constraint
        self._wordforms: Dict[str, List[str]] = dict(wordforms or {})
        for path in (wordforms_tsvpath, wordforms_attribs_tsvpath):       # class name <TAB> comma separated word forms
            if path:
                with open(path, "r") as f:
                    for row in csv.reader(f, delimiter="\t"):
                        if len(row) >= 2:
                            self._wordforms[row[0]] = row[1].split(",")

    def _ids(self, word: str) -> List[int]:
        return [int(self._vocabulary.get_token_index(w)) for w in self._wordforms[word]]

    def build(self, constraints: Sequence[str]) -> Tuple[FsmProgram, int, Dict[str, List[int]]]:
        """-> (program, number of states in use, {constraint: main states it leads to}) like the reference's
        (fsm, substate_idx, constraint2states)."""
        n_main = self._num_main_states
        next_sub = n_main
        first_invalid = 2 ** len(constraints)
        added_at: Dict[str, List[int]] = {}
        connections = []
        constraint2states: Dict[str, List[int]] = {}
        for n, constraint in enumerate(constraints, start=1):
            words = constraint.split()
            stride = 1 << (n - 1)
            if constraint in added_at:           # a repeated constraint continues from where its previous copy was added
                origin = added_at[constraint][-1]
                stop = origin + 1
            else:
                origin, stop = 0, n_main
            added_at.setdefault(constraint, []).append(n)
            reached: List[int] = []
            while origin < stop:
                for _ in range(stride):
                    src = origin
                    for word in words[:-1]:      # inner words of a multi-word constraint walk through fresh sub-states
                        connections.append((src, next_sub, origin, self._ids(word)))
                        src = next_sub
                        next_sub += 1
                    if words:
                        target = origin + stride
                        if target < first_invalid:
                            reached.append(target)
                        connections.append((src, target, origin, self._ids(words[-1])))
                    origin += 1
                origin += stride
            constraint2states[constraint] = reached
        if next_sub > self._num_total_states:
            raise ValueError("constraints need %d states, more than max_given_constraints * max_words_per_constraint "
                             "allows (%d)" % (next_sub, self._num_total_states))
        prog = FsmProgram(connections, n_main, next_sub, int(self._vocabulary.get_vocab_size()))
        return prog, next_sub, constraint2states


def build_fsm_bits(programs: Sequence[FsmProgram], device) -> FsmBits:
    """Expands a batch of programs (possibly with different state counts) into the (B, S, V) bit table on `device`."""
    if len(programs) == 0:
        raise ValueError("empty FSM batch")
    V = programs[0].vocab_size
    if any(p.vocab_size != V for p in programs):
        raise ValueError("all programs of a batch must share the vocabulary")
    S = max(p.num_states for p in programs)
    rec, off, wf, n_main = [], [0], [], []
    for p in programs:
        for frm, to, reset, ids in p.connections:
            if min(frm, to, reset) < 0 or any(not 0 <= i < V for i in ids):
                raise ValueError("connection out of range")
            rec += [frm, to, reset, len(wf), len(ids)]     # states >= p.num_states are trimmed away on the device
            wf += ids
        off.append(len(rec) // 5)
        n_main += [p.num_main_states, p.num_states]
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("sscvae FSM tables are built on a CUDA device; there is no CPU fallback")
    host = torch.tensor(rec + off + wf + n_main + [0], dtype=torch.int32)
    buf = host.to(dev)
    a, b, c = len(rec), len(rec) + len(off), len(rec) + len(off) + len(wf)
    bits = torch.empty(len(programs), S, V, dtype=torch.int32, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(_lib.lib().sscvae_fsm_build(_lib.ptr(buf[:a]) if a else _lib.ptr(buf), _lib.ptr(buf[a:b]), _lib.ptr(buf[b:c + 1]),
                                           _lib.ptr(buf[c:]), len(programs), S, V, _lib.ptr(bits), stream))
    return FsmBits(bits)


def valid_states_with_attributes(num_constraints: int, constraints, constraint2states, min_constraints_to_satisfy: int = 2):
    """The non-`cbs_simple` set of acceptable end states of one image (updown-baseline/updown/utils/decoding.py:87-123).

    constraints: [(object, [attributes]), ...]; constraint2states: {constraint: main states}. A state counts an object
    when it has seen the object word and (if the object lists attributes) at least one of them; once any object with
    attributes is matched, only states that match such an object count. Returns the list of valid state indices."""
    n = 2 ** int(num_constraints)
    count = [0] * n
    attributed = [False] * n
    for obj, attrs in constraints:
        has_obj = set(constraint2states[obj])
        if attrs:
            has_attr = set()
            for a in attrs:
                has_attr.update(constraint2states[a])
        else:
            has_attr = set(range(n))
        both = {s for s in has_obj & has_attr if s < n}
        if len({s for s in has_attr if s < n}) < n:
            for s in both:
                attributed[s] = True
        for s in both:
            count[s] += 1
    if any(attributed):
        count = [c if attributed[s] else 0 for s, c in enumerate(count)]
    need = min(len(constraints), min_constraints_to_satisfy)
    return [s for s in range(n) if count[s] >= need]
