"""Harness that imports the UNMODIFIED reference (visinf/style-seqcvae) from /root/reference.

TEST INFRASTRUCTURE ONLY — never imported by the product path, and only usable in the build
container (the GPU box has no /root/reference; nothing under `-m gpu`, smoke() or bench.py
imports this file). It is used by `oracle/gen_golden.py` to produce the fixtures under
`tests/golden/` and by `tests/test_oracle_vs_reference.py` (skipped when the reference is absent)
to pin the restatement in `oracle/updown_oracle.py` / `oracle/search_oracle.py`.

Third-party packages the reference imports but this image lacks (allennlp 0.8.4, torchtext,
yacs, overrides, anytree) are replaced by the stubs in `oracle/ref_shims/`.
Nothing from /root/reference is copied into the repo: the modules are imported in place, and the
three torch>=1.2 incompatibilities of `updown-baseline/updown/modules/cbs.py` (:135, :205 `1 - uint8
mask`; :231 `/` on int64) are patched on the source text in memory.

SENTIMENT_VAE = 2: with a frozen-GloVe embedding size the reference's ctor opens hard-coded
`/path/to/*.pkl|json` files and reads an attribute that is never set (`self.senti_glove_5`,
updown_captioner.py:76-93), so it cannot be constructed as shipped. `build_reference_model` therefore constructs the
captioner with sentiment_vae=1 and then swaps in the reference's own `UpDownCell(..., sentiment_vae=2, ...)`, sets
`model.sentiment_vae = 2` and installs a caller-supplied `mean_choice` table: every line of the forward pass that runs
is the reference's.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SSCVAE_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shims")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "var_updown", "var_updown"))


_loaded = {}


def load_reference():
    """Returns a namespace with the reference's UpDownCaptioner, patched CBS, vendored BeamSearch,
    select_best_beam*, FiniteStateMachineBuilder and the stub Vocabulary."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (os.path.join(REFERENCE_ROOT, "var_updown"),
              os.path.join(REFERENCE_ROOT, "updown-baseline"), _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    # the reference `updown/modules/__init__.py` imports cbs.py as-is (importable; only its
    # execution fails on torch 2), so a plain import works.
    models = importlib.import_module("var_updown.models")
    decoding = importlib.import_module("updown.utils.decoding")
    constraints = importlib.import_module("updown.utils.constraints")
    vendored_beam = importlib.import_module("var_updown.modules.beam_search")
    from allennlp.data import Vocabulary

    # CBS with the 3 torch-2 compat edits, applied to the source text in memory.
    cbs_path = os.path.join(REFERENCE_ROOT, "updown-baseline", "updown", "modules", "cbs.py")
    src = open(cbs_path).read()
    edits = [
        ("1 - fsm[:, 0, :, :]", "~fsm[:, 0, :, :].bool()"),
        ("1 - step_state_mask[:, :, i, :, :]", "~step_state_mask[:, :, i, :, :].bool()"),
        ("restricted_beam_indices / self.per_node_beam_size",
         "restricted_beam_indices // self.per_node_beam_size"),
    ]
    for old, new in edits:
        assert src.count(old) == 1, f"reference cbs.py changed: {old!r}"
        src = src.replace(old, new)
    mod = types.ModuleType("sscvae_ref_cbs_patched")
    exec(compile(src, cbs_path + " (patched in memory)", "exec"), mod.__dict__)

    _loaded.update(
        UpDownCaptioner=models.UpDownCaptioner,
        ConstrainedBeamSearch=mod.ConstrainedBeamSearch,
        BeamSearch=vendored_beam.BeamSearch,
        select_best_beam=decoding.select_best_beam,
        select_best_beam_with_constraints=decoding.select_best_beam_with_constraints,
        FiniteStateMachineBuilder=constraints.FiniteStateMachineBuilder,
        add_constraint_words_to_vocabulary=constraints.add_constraint_words_to_vocabulary,
        Vocabulary=Vocabulary,
    )
    return types.SimpleNamespace(**_loaded)


def make_vocabulary(vocab_size: int):
    ref = load_reference()
    return ref.Vocabulary(["@@UNKNOWN@@", "@@BOUNDARY@@"] + [f"w{i}" for i in range(vocab_size - 2)])


def build_reference_model(vocab, *, image_feature_size, embedding_size, hidden_size,
                          attention_projection_size, max_caption_length=20, beam_size=1,
                          use_cbs=None, min_constraints_to_satisfy=2, z_space=150, prior_std=1.0,
                          simple_vae=False, latent_embedding="glove", sentiment_vae=1,
                          senti_prior_multip=0.5, cbs_simple=True, seed=0, latent_embedding_multip=1,
                          mean_choice=None):
    """Constructs the reference captioner on CPU with `torch.manual_seed(seed)` default init
    (SURVEY Appendix C)."""
    import torch
    ref = load_reference()
    if use_cbs is None:
        use_cbs = embedding_size in (300, 600)
    torch.manual_seed(seed)
    model = ref.UpDownCaptioner(
        vocab, image_feature_size, embedding_size, hidden_size, attention_projection_size,
        max_caption_length=max_caption_length, beam_size=beam_size, use_cbs=use_cbs,
        min_constraints_to_satisfy=min_constraints_to_satisfy, z_space=z_space,
        prior_std=prior_std, simple_vae=simple_vae, latent_embedding=latent_embedding,
        latent_embedding_multip=latent_embedding_multip,
        sentiment_vae=1 if sentiment_vae == 2 else sentiment_vae, senti_prior_multip=senti_prior_multip,
        cbs_simple=cbs_simple, device=torch.device("cpu"))
    if sentiment_vae == 2:                               # see the module docstring
        cell_mod = importlib.import_module("var_updown.modules.updown_cell")
        model._updown_cell = cell_mod.UpDownCell(
            image_feature_size, embedding_size, hidden_size, attention_projection_size, z_space, 2, simple_vae,
            torch.device("cpu"), latent_embedding)
        model.sentiment_vae = 2
        model.mean_choice = mean_choice
    if use_cbs:
        model._beam_search = ref.ConstrainedBeamSearch(
            model._boundary_index, max_steps=max_caption_length, beam_size=beam_size,
            per_node_beam_size=beam_size // 2)
    return model
