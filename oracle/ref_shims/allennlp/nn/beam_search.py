"""Constructor-only stub: the reference builds `allennlp.nn.beam_search.BeamSearch` when
use_cbs=False (updown_captioner.py:129-135) but its 0.8.4 `search` cannot consume the
var_updown 5-tuple step output, so only the ctor is needed."""


class BeamSearch:
    def __init__(self, end_index, max_steps=50, beam_size=10, per_node_beam_size=None):
        self._end_index = end_index
        self.max_steps = max_steps
        self.beam_size = beam_size
        self.per_node_beam_size = per_node_beam_size or beam_size
