"""CPU oracle: plain-torch fp32 restatement of the Style-SeqCVAE `var_updown` decoder hot path.

TEST INFRASTRUCTURE ONLY. Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import this file; the product path
(`style-seqcvae_b200/`) never does and fails loudly when its CUDA library is missing.

Parity status: the reference ships no tests or golden vectors, so this restatement
is pinned against OUTPUTS OF THE REFERENCE ITSELF: `oracle/gen_golden.py` runs the unmodified
reference module (imported in place from /root/reference through `oracle/ref_harness.py`) on seeded
inputs and commits the results under `tests/golden/`; `tests/test_oracle_golden.py` checks this file
against those fixtures on every CPU run, and `tests/test_oracle_vs_reference.py` re-runs the live
comparison whenever /root/reference is present.

Every function cites the reference lines it follows. Paths are relative to /root/reference.
The math is written out explicitly (no nn.LSTMCell / nn.Linear modules) so that it can also be
run with `Rounding("bf16")`, which rounds exactly the tensors the CUDA path stores in bf16
(GEMM operands, projected features) and lets the GPU tests separate "algorithm differs" from
"bf16 operands".  Autograd over this forward is the backward (BPTT) oracle.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch


# ----------------------------------------------------------------------------------------------
# configuration
# ----------------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    """Mirrors the ctor arguments of the reference captioner
    (var_updown/var_updown/models/updown_captioner.py:21-41)."""
    vocab_size: int
    image_feature_size: int = 2048
    embedding_size: int = 600
    hidden_size: int = 900
    attention_projection_size: int = 768
    z_space: int = 150
    max_caption_length: int = 20
    sentiment_vae: int = 1
    simple_vae: bool = False
    prior_std: float = 1.0
    senti_prior_multip: float = 0.5
    latent_embedding: str = "glove"
    pad_index: int = 0        # "@@UNKNOWN@@"  (updown_captioner.py:61)
    boundary_index: int = 1   # "@@BOUNDARY@@" (updown_captioner.py:62)

    @property
    def tied(self) -> bool:
        # frozen embedding tied to the output layer (updown_captioner.py:75,112-119)
        return self.embedding_size in (300, 600)

    @property
    def cond_size(self) -> int:
        # width of the conditioning column block of the enc/dec LSTM inputs
        # (var_updown/var_updown/modules/updown_cell.py:47-81)
        if self.simple_vae or self.sentiment_vae == 0:
            return 0
        if self.sentiment_vae == 1 or self.latent_embedding == "senti_word_net":
            return 1                                     # cell:55-61 (the elif order of the reference)
        if self.sentiment_vae == 2:
            return self.z_space                          # cell:63-70 writes the literal 150 = its Z_SPACE
        raise NotImplementedError


class Rounding:
    """Operand rounding used to emulate the CUDA path's storage precision.

    mode "fp32": identity (the reference's arithmetic).
    mode "bf16": tensors that the CUDA path feeds to tensor-core GEMMs (and the projected
    region features it keeps in bf16) are rounded to bf16 with a straight-through gradient;
    everything else (cell states, softmax, KL, CE, accumulators) stays fp32.
    """

    def __init__(self, mode: str = "fp32"):
        assert mode in ("fp32", "bf16")
        self.mode = mode

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if self.mode == "fp32":
            return x
        return x + (x.detach().to(torch.bfloat16).to(torch.float32) - x.detach())


FP32 = Rounding("fp32")


# ----------------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------------
PARAM_KEYS_COMMON = [
    "_embedding_layer.weight",
    "_updown_cell._attention_lstm_cell.weight_ih", "_updown_cell._attention_lstm_cell.weight_hh",
    "_updown_cell._attention_lstm_cell.bias_ih", "_updown_cell._attention_lstm_cell.bias_hh",
    "_updown_cell._butd_attention._query_vector_projection_layer.weight",
    "_updown_cell._butd_attention._image_features_projection_layer.weight",
    "_updown_cell._butd_attention._attention_layer.weight",
    "_updown_cell._language_lstm_cell_encoder.weight_ih", "_updown_cell._language_lstm_cell_encoder.weight_hh",
    "_updown_cell._language_lstm_cell_encoder.bias_ih", "_updown_cell._language_lstm_cell_encoder.bias_hh",
    "_updown_cell._language_lstm_cell_decoder.weight_ih", "_updown_cell._language_lstm_cell_decoder.weight_hh",
    "_updown_cell._language_lstm_cell_decoder.bias_ih", "_updown_cell._language_lstm_cell_decoder.bias_hh",
    "_updown_cell.fc_mean.weight", "_updown_cell.fc_mean.bias",
    "_updown_cell.fc_log_var.weight", "_updown_cell.fc_log_var.bias",
]


def param_shapes(cfg: OracleConfig) -> Dict[str, tuple]:
    """state_dict keys and shapes of the reference module."""
    E, F, H, A, Z, V, s = (cfg.embedding_size, cfg.image_feature_size, cfg.hidden_size,
                           cfg.attention_projection_size, cfg.z_space, cfg.vocab_size, cfg.cond_size)
    c = "_updown_cell."
    shapes = {
        "_embedding_layer.weight": (V, E),
        c + "_attention_lstm_cell.weight_ih": (4 * H, E + F + 2 * H),
        c + "_attention_lstm_cell.weight_hh": (4 * H, H),
        c + "_attention_lstm_cell.bias_ih": (4 * H,),
        c + "_attention_lstm_cell.bias_hh": (4 * H,),
        c + "_butd_attention._query_vector_projection_layer.weight": (A, H),
        c + "_butd_attention._image_features_projection_layer.weight": (A, F),
        c + "_butd_attention._attention_layer.weight": (1, A),
        c + "_language_lstm_cell_encoder.weight_ih": (4 * H, s + F + 2 * H),
        c + "_language_lstm_cell_encoder.weight_hh": (4 * H, H),
        c + "_language_lstm_cell_encoder.bias_ih": (4 * H,),
        c + "_language_lstm_cell_encoder.bias_hh": (4 * H,),
        c + "_language_lstm_cell_decoder.weight_ih": (4 * H, s + F + 2 * H + Z),
        c + "_language_lstm_cell_decoder.weight_hh": (4 * H, H),
        c + "_language_lstm_cell_decoder.bias_ih": (4 * H,),
        c + "_language_lstm_cell_decoder.bias_hh": (4 * H,),
        c + "fc_mean.weight": (Z, H), c + "fc_mean.bias": (Z,),
        c + "fc_log_var.weight": (Z, H), c + "fc_log_var.bias": (Z,),
    }
    if cfg.tied:
        shapes["_output_projection.0.weight"] = (E, H)
        shapes["_output_projection.0.bias"] = (E,)
        shapes["_output_layer.weight"] = (V, E)      # aliases _embedding_layer.weight
    else:
        shapes["_output_layer.weight"] = (V, H)
        shapes["_output_layer.bias"] = (V,)
    return shapes


def init_params(cfg: OracleConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random init with the reference's distributions: LSTMCell/Linear default U(-1/sqrt(fan),..),
    Embedding N(0,1) with a zero padding row, and for the tied case the reference's OOV rule
    `2*randn-1` per row (updown_captioner.py:197,209,215). Seeded; NOT bit-identical to the
    reference's init order (tests that need the reference's weights load its state_dict)."""
    g = torch.Generator().manual_seed(seed)
    H = cfg.hidden_size
    out = {}
    for k, shp in param_shapes(cfg).items():
        if k == "_output_layer.weight" and cfg.tied:
            continue
        if k == "_embedding_layer.weight":
            if cfg.tied:
                w = 2 * torch.randn(shp, generator=g) - 1
            else:
                w = torch.randn(shp, generator=g)
                w[cfg.pad_index] = 0
            out[k] = w
            continue
        if "lstm_cell" in k:
            bound = 1.0 / (H ** 0.5)
        else:
            fan_in = shp[1] if len(shp) == 2 else param_shapes(cfg)[k.replace("bias", "weight")][1]
            bound = 1.0 / (fan_in ** 0.5)
        out[k] = (torch.rand(shp, generator=g) * 2 - 1) * bound
    if cfg.tied:
        out["_output_layer.weight"] = out["_embedding_layer.weight"]
    return out


# ----------------------------------------------------------------------------------------------
# restated third-party helpers (allennlp 0.8.4 semantics)
# ----------------------------------------------------------------------------------------------
def masked_softmax(x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """allennlp masked_softmax, memory_efficient=False (attention.py:93):
    softmax(x*m) * m / (sum + 1e-13)."""
    m = mask.to(x.dtype)
    r = torch.softmax(x * m, dim=-1) * m
    return r / (r.sum(dim=-1, keepdim=True) + 1e-13)


def masked_mean_boxes(feats: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """allennlp masked_mean over the box axis (updown_cell.py:266): sum(x*m)/clamp(sum m, 1e-8)."""
    m = mask.to(feats.dtype).unsqueeze(-1)
    return (feats * m).sum(dim=1) / m.sum(dim=1).clamp(min=1e-8)


def add_boundary_tokens(tokens: torch.Tensor, pad: int, boundary: int) -> torch.Tensor:
    """allennlp add_sentence_boundary_token_ids as called at updown_captioner.py:265-270:
    (B,L) -> (B,L+2) = [boundary, w_1..w_n, boundary, 0...] with n = #(tokens != pad)."""
    B, L = tokens.shape
    lengths = (tokens != pad).sum(dim=1)
    out = tokens.new_zeros(B, L + 2)
    out[:, 1:-1] = tokens
    out[:, 0] = boundary
    out[torch.arange(B), lengths + 1] = boundary
    return out


def lstm_cell(x_gates: torch.Tensor, c_prev: torch.Tensor):
    """torch.nn.LSTMCell pointwise part: gate order i,f,g,o along 4H."""
    i, f, g, o = x_gates.chunk(4, dim=1)
    i, f, g, o = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(g), torch.sigmoid(o)
    c = f * c_prev + i * g
    h = o * torch.tanh(c)
    return h, c


# ----------------------------------------------------------------------------------------------
# per-image precompute and one decoder step
# ----------------------------------------------------------------------------------------------
def image_precompute(p, cfg: OracleConfig, image_features: torch.Tensor, q: Rounding = FP32):
    """mask, masked mean and W_v projection of the region features.
    updown_cell.py:233-270 (`_average_image_features`), attention.py:99-125."""
    feats = q(image_features)
    mask = image_features.abs().sum(dim=-1) > 0                      # cell:263 (on the fp32 input)
    avg = masked_mean_boxes(feats, mask)                             # cell:266
    Wv = p["_updown_cell._butd_attention._image_features_projection_layer.weight"]
    proj = q(torch.matmul(feats, q(Wv).t()))                         # attn:125
    return feats, mask, avg, proj


def prior(cfg: OracleConfig, sentiment: Optional[torch.Tensor], batch: int):
    """updown_captioner.py:250-261 and updown_cell.py:165-166."""
    Z = cfg.z_space
    if cfg.sentiment_vae == 0 or cfg.simple_vae:
        prior_mean = torch.zeros(batch, Z)
    elif cfg.sentiment_vae == 1:
        prior_mean = sentiment.repeat(1, Z) * cfg.senti_prior_multip
    elif cfg.sentiment_vae == 2:
        prior_mean = torch.zeros(batch, Z)               # capt:254-256; replaced per step by the cell (cell:160-163)
    else:
        raise NotImplementedError
    prior_var = (torch.ones(batch, Z) * cfg.prior_std).pow(2)
    return prior_mean, prior_var


def zero_states(rows: int, H: int) -> Dict[str, torch.Tensor]:
    """updown_cell.py:131-140."""
    return {k: torch.zeros(rows, H) for k in
            ("h1", "c1", "h_encoder", "c_encoder", "h_decoder", "c_decoder")}


def decoder_step(p, cfg: OracleConfig, feats, mask, avg, proj, tokens, states, sentiment,
                 prior_mean, prior_var, eps, training: bool, q: Rounding = FP32, obj_means=None):
    """One timestep: embedding -> UpDownCell -> output head.
    updown_captioner.py:430-450, updown_cell.py:143-231, attention.py:68-97.
    obj_means: (rows, N, Z) per-box attribute means, sentiment_vae == 2 only.
    Returns (logits, new_states, aux) with aux = {alpha, attended, mean, log_var, z, h_dec, prior_mean}."""
    c = "_updown_cell."
    E, F, H = cfg.embedding_size, cfg.image_feature_size, cfg.hidden_size
    emb = p["_embedding_layer.weight"][tokens]                       # capt:430
    # --- attention LSTM (cell:143-148)
    x_att = torch.cat([emb, avg, states["h1"], states["h_decoder"]], dim=1)
    gates = (torch.matmul(q(x_att), q(p[c + "_attention_lstm_cell.weight_ih"]).t())
             + p[c + "_attention_lstm_cell.bias_ih"]
             + torch.matmul(q(states["h1"]), q(p[c + "_attention_lstm_cell.weight_hh"]).t())
             + p[c + "_attention_lstm_cell.bias_hh"])
    h1, c1 = lstm_cell(gates, states["c1"])
    # --- bottom-up top-down attention (attn:69-93)
    qv = torch.matmul(q(h1), q(p[c + "_butd_attention._query_vector_projection_layer.weight"]).t())
    u = torch.matmul(torch.tanh(qv.unsqueeze(1) + proj),
                     p[c + "_butd_attention._attention_layer.weight"].t()).squeeze(-1)
    alpha = masked_softmax(u, mask)
    attended = (alpha.unsqueeze(-1) * feats).sum(dim=1)              # cell:156-158
    # --- attribute-grounded prior (cell:160-174): the prior mean of THIS step follows the attention
    if cfg.sentiment_vae == 2:
        prior_mean = (alpha.unsqueeze(-1) * obj_means).sum(dim=1)    # cell:160-163
    if cfg.simple_vae:
        prior_mean = torch.zeros_like(prior_mean)                    # cell:165-166
    # --- conditioning column (cell:168-190, 211-224)
    if cfg.cond_size == 0:
        cond = []
    elif cfg.sentiment_vae == 1:
        cond = [sentiment]
    elif cfg.latent_embedding == "glove":
        cond = [prior_mean]                                          # cell:168-169
    else:
        cond = [prior_mean[:, 0].unsqueeze(1)]                       # cell:170-171
    h_dec_prev = states["h_decoder"]
    if training:
        x_enc = torch.cat([attended, h1, h_dec_prev] + cond, dim=1)  # cell:178-190
        g_enc = (torch.matmul(q(x_enc), q(p[c + "_language_lstm_cell_encoder.weight_ih"]).t())
                 + p[c + "_language_lstm_cell_encoder.bias_ih"]
                 + torch.matmul(q(states["h_encoder"]), q(p[c + "_language_lstm_cell_encoder.weight_hh"]).t())
                 + p[c + "_language_lstm_cell_encoder.bias_hh"])
        h_enc, c_enc = lstm_cell(g_enc, states["c_encoder"])         # cell:192-194
        mean = torch.matmul(q(h_enc), q(p[c + "fc_mean.weight"]).t()) + p[c + "fc_mean.bias"]
        log_var = torch.matmul(q(h_enc), q(p[c + "fc_log_var.weight"]).t()) + p[c + "fc_log_var.bias"]
        var = log_var.exp()                                          # cell:196-198
    else:
        h_enc, c_enc = states["h_encoder"], states["c_encoder"]
        mean, var = prior_mean, prior_var                            # cell:201-203
        log_var = var.log()
    z = eps * var.sqrt() + mean                                      # cell:206-208
    x_dec = torch.cat([attended, h1, h_dec_prev] + cond + [z], dim=1)  # cell:211-224
    g_dec = (torch.matmul(q(x_dec), q(p[c + "_language_lstm_cell_decoder.weight_ih"]).t())
             + p[c + "_language_lstm_cell_decoder.bias_ih"]
             + torch.matmul(q(h_dec_prev), q(p[c + "_language_lstm_cell_decoder.weight_hh"]).t())
             + p[c + "_language_lstm_cell_decoder.bias_hh"])
    h_dec, c_dec = lstm_cell(g_dec, states["c_decoder"])             # cell:227-229
    # --- output head (capt:112-126, 444-445)
    if cfg.tied:
        o = torch.tanh(torch.matmul(q(h_dec), q(p["_output_projection.0.weight"]).t())
                       + p["_output_projection.0.bias"])
        logits = torch.matmul(q(o), q(p["_embedding_layer.weight"]).t())
    else:
        logits = torch.matmul(q(h_dec), q(p["_output_layer.weight"]).t()) + p["_output_layer.bias"]
    new_states = {"h1": h1, "c1": c1, "h_encoder": h_enc, "c_encoder": c_enc,
                  "h_decoder": h_dec, "c_decoder": c_dec}
    aux = {"alpha": alpha, "attended": attended, "mean": mean, "log_var": log_var, "z": z,
           "h_dec": h_dec, "h1": h1, "prior_mean": prior_mean}
    return logits, new_states, aux


def kl_step(cfg: OracleConfig, mean, log_var, prior_mean, prior_var):
    """updown_captioner.py:295-303."""
    if cfg.sentiment_vae == 0:
        return -0.5 * torch.sum(1 + log_var - mean.pow(2) - log_var.exp(), dim=1)
    prior_log_var = prior_var.log()
    kld = 1 + log_var - prior_log_var - ((mean - prior_mean).pow(2) + log_var.exp()) / (prior_var + 0.00001)
    return -0.5 * kld.sum(1)


# ----------------------------------------------------------------------------------------------
# training forward (teacher forced) — updown_captioner.py:263-323
# ----------------------------------------------------------------------------------------------
def train_forward(p, cfg: OracleConfig, image_features, caption_tokens, sentiment, eps,
                  q: Rounding = FP32, record: bool = False, obj_means=None):
    """image_features (B,N,F) f32; caption_tokens (B,L) i64 pad=0; sentiment (B,1) f32 or None;
    eps (T,B,Z) with T=L+1 (the reference draws one (B,Z) normal per step, cell:206).
    Returns {"loss": (B,), "kld": (B,)} (+ per-step records)."""
    B = image_features.shape[0]
    tokens = add_boundary_tokens(caption_tokens, cfg.pad_index, cfg.boundary_index)   # capt:265-270
    T = tokens.shape[1] - 1                                                           # capt:278
    tokens_mask = tokens != cfg.pad_index                                             # capt:274
    feats, mask, avg, proj = image_precompute(p, cfg, image_features, q)
    prior_mean, prior_var = prior(cfg, sentiment, B)
    states = zero_states(B, cfg.hidden_size)
    step_logits, step_klds, rec = [], [], []
    for t in range(T):                                                                # capt:282
        logits, states, aux = decoder_step(p, cfg, feats, mask, avg, proj, tokens[:, t], states,
                                           sentiment, prior_mean, prior_var, eps[t], True, q, obj_means)
        prior_mean = aux["prior_mean"]                   # capt:285 reassigns it from the step's return value
        step_klds.append(kl_step(cfg, aux["mean"], aux["log_var"], prior_mean, prior_var))
        step_logits.append(logits)
        if record:
            rec.append(aux)
    logits = torch.stack(step_logits, dim=1)                                          # (B,T,V) capt:312
    klds = torch.stack(step_klds, dim=1) * tokens_mask[:, 1:].float()                 # capt:315
    targets = tokens[:, 1:]
    tmask = tokens_mask[:, 1:].float()
    # _get_loss (capt:457-466): len * [ sum_t m*nll / (sum_t m + 1e-13) ]
    logp = torch.log_softmax(logits, dim=-1)
    nll = -logp.gather(2, targets.unsqueeze(-1)).squeeze(-1) * tmask
    lengths = tmask.sum(dim=-1)
    loss = lengths * (nll.sum(1) / (lengths + 1e-13))
    out = {"loss": loss, "kld": klds.sum(dim=1)}                                      # capt:318-323
    if record:
        out.update(logits=logits, klds=klds, steps=rec, tokens=tokens)
    return out


def train_objective(out, kld_weight: float = 750.0):
    """var_updown/scripts/train.py:168-171."""
    return out["loss"].mean() + out["kld"].mean() / kld_weight


# ----------------------------------------------------------------------------------------------
# decode step function (eval) — updown_captioner.py:371-455 with the ALIGNED replication of
# sentiment/prior rows (the reference tiles them, which is only self-consistent
# at B=1 / uniform sentiment; both agree there).
# ----------------------------------------------------------------------------------------------
class DecodeStepper:
    """Callable `(last_predictions (R,), states|None, eps (R,Z)) -> (logp (R,V), states)` closed
    over one batch of images; R = B * net_beam with rows of an image contiguous
    (row r belongs to image r // net_beam), as updown_captioner.py:405-416 lays them out."""

    def __init__(self, p, cfg: OracleConfig, image_features, sentiment, q: Rounding = FP32, obj_means=None):
        self.p, self.cfg, self.q = p, cfg, q
        self.B = image_features.shape[0]
        self.feats, self.mask, self.avg, self.proj = image_precompute(p, cfg, image_features, q)
        self.sentiment = sentiment
        self.obj_means = obj_means
        self.prior_mean, self.prior_var = prior(cfg, sentiment, self.B)

    def __call__(self, last_predictions, states, eps):
        R = last_predictions.shape[0]
        nb = R // self.B
        rep = lambda x: None if x is None else x.repeat_interleave(nb, dim=0)
        if states is None:
            states = zero_states(R, self.cfg.hidden_size)
        logits, states, aux = decoder_step(
            self.p, self.cfg, rep(self.feats), rep(self.mask), rep(self.avg), rep(self.proj),
            last_predictions, states, rep(self.sentiment), rep(self.prior_mean), rep(self.prior_var),
            eps, False, self.q, rep(self.obj_means))
        return torch.log_softmax(logits, dim=1), states                               # capt:450
