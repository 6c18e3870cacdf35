"""style-seqcvae_b200: B200-native (sm_100a) implementation of the Style-SeqCVAE `var_updown`
sequential-decoder hot path, drop-in behind the reference's UpDownCaptioner module API.

The directory name carries a hyphen (it is the repo's package directory); import it
through the repo-root shim: `import sscvae` (sscvae.py), which registers this package as
`style_seqcvae_b200`.
"""
from . import _lib
from .captioner import UpDownCaptioner
from .search import (BeamSearch, ConstrainedBeamSearch, select_best_beam, select_best_beam_with_constraints, pad_fsm_batch)
from .dp import BucketedGradReducer, shard_batch, global_grad_norm
from .optim import FusedClipSGD
from .ingest import FeatureCache, collate_features, pack_features
from .fsm import FiniteStateMachineBuilder, FsmProgram, FsmBits, build_fsm_bits, valid_states_with_attributes

__all__ = ["FeatureCache", "collate_features", "pack_features", "UpDownCaptioner", "ConstrainedBeamSearch", "BeamSearch", "select_best_beam",
           "select_best_beam_with_constraints", "pad_fsm_batch", "BucketedGradReducer", "shard_batch", "global_grad_norm",
           "FusedClipSGD", "FiniteStateMachineBuilder", "FsmProgram", "FsmBits", "build_fsm_bits", "valid_states_with_attributes"]
