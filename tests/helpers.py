"""Shared helpers for the GPU parity tests: build the drop-in module from a golden fixture and run
the oracle on the same inputs."""
import torch

import sscvae
from oracle import updown_oracle as uo


class StubVocabulary:
    """The four methods the captioner needs."""

    def __init__(self, size):
        self._size = size

    def get_vocab_size(self, namespace="tokens"):
        return self._size

    def get_token_index(self, token, namespace="tokens"):
        return {"@@UNKNOWN@@": 0, "@@BOUNDARY@@": 1}.get(token, 0)

    def get_token_to_index_vocabulary(self, namespace="tokens"):
        d = {"@@UNKNOWN@@": 0, "@@BOUNDARY@@": 1}
        d.update({f"w{i}": i + 2 for i in range(self._size - 2)})
        return d


def module_from_cfg(cfg: dict, params=None, beam_size=1, use_cbs=None, device="cuda", min_sat=2, **extra):
    tied = cfg["embedding_size"] in (300, 600)
    if use_cbs is None:
        use_cbs = tied
    m = sscvae.UpDownCaptioner(
        StubVocabulary(cfg["vocab_size"]), cfg["image_feature_size"], cfg["embedding_size"], cfg["hidden_size"],
        cfg["attention_projection_size"], max_caption_length=cfg["max_caption_length"], beam_size=beam_size,
        use_cbs=use_cbs, min_constraints_to_satisfy=min_sat, z_space=cfg["z_space"], prior_std=cfg["prior_std"],
        simple_vae=cfg["simple_vae"], latent_embedding=cfg.get("latent_embedding", "glove"),
        sentiment_vae=cfg["sentiment_vae"], senti_prior_multip=cfg["senti_prior_multip"], cbs_simple=True,
        device=torch.device(device), **extra)
    if params is not None:
        missing, unexpected = m.load_state_dict(params, strict=True)
    return m.to(device)


def oracle_params(params, cfg: uo.OracleConfig, grad=False):
    p = {k: v.clone().requires_grad_(grad) for k, v in params.items()}
    if cfg.tied:
        p["_output_layer.weight"] = p["_embedding_layer.weight"]
    return p


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()
