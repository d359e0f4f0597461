"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (same rules as oracle/seqdiff_oracle.py).

Plain-PyTorch (CPU, fp32) restatement of the reference's `structure_model` angle denoiser and its Gaussian
reverse step:

  * `struct_forward`          structure_model/model.py:155-215  (ConditionalBertForDiffusionBase.forward)
  * `cosine_beta_schedule`    structure_model/utils.py:8-18
  * `compute_alphas`          structure_model/utils.py:42-58
  * `modulo_with_wrapped_range` structure_model/utils.py:20-40
  * `p_sample` / `p_sample_loop` structure_model/sample.py:55-144

The blocks (SELayer, BertAttention with the 4.38.2 relative_key term, BertLayer, BertEmbeddings,
GaussianFourierProjection, the MLM-style head) are the ones already restated in seqdiff_oracle.py; only the wiring
differs.  Parity pin: oracle/make_golden.py::golden_structure_model runs the UNMODIFIED reference module imported in
place (oracle/ref_import.py::load_structure_reference) with and without the restored relative_key attention and
requires max |oracle - reference| == 0; the fixtures are tests/golden/struct_*.pt.  The relative_key term itself is
"parity unpinned" for the same reason as in the sequence model (transformers 4.38.2 is not vendored).

The reference draws the reverse-step noise with `torch.randn_like` on its device; "same noise" here = both sides are
handed the same explicit N(0,1) tensor per step.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import seqdiff_oracle as O

Tensor = torch.Tensor
OracleConfig = O.OracleConfig  # feature_size = number of angle features (8 with the reference's dataset)

_SE_BLOCKS = ("receptor_emb", "timestep_emb")


def struct_state_dict_schema(cfg: OracleConfig) -> Dict[str, tuple]:
    """name -> shape of structure_model ConditionalBertForDiffusionBase.state_dict() (model.py:163-178); encoder and decoder
    share one BertConfig apart from the cross-attention flags (sample.py:151-173)."""
    H, I, P, Fs = cfg.hidden_size, cfg.intermediate_size, cfg.max_position_embeddings, cfg.feature_size
    dh = cfg.head_dim
    s: Dict[str, tuple] = {"timestep_projector.W": (H // 2,)}
    for p, fin in (("receptor_seq_emb", 20), ("receptor_angle_emb", Fs), ("ligand_angle_emb", Fs)):
        s[f"{p}.linear.weight"] = (H, fin)
        s[f"{p}.linear.bias"] = (H,)
        s[f"{p}.LayerNorm.weight"] = (H,)
        s[f"{p}.LayerNorm.bias"] = (H,)

    def attn(prefix, rel):
        for n in ("query", "key", "value"):
            s[f"{prefix}.self.{n}.weight"] = (H, H)
            s[f"{prefix}.self.{n}.bias"] = (H,)
        if rel and cfg.relative_key:
            s[f"{prefix}.self.distance_embedding.weight"] = (2 * P - 1, dh)
        s[f"{prefix}.output.dense.weight"] = (H, H)
        s[f"{prefix}.output.dense.bias"] = (H,)
        s[f"{prefix}.output.LayerNorm.weight"] = (H,)
        s[f"{prefix}.output.LayerNorm.bias"] = (H,)

    for blk in _SE_BLOCKS:
        s[f"{blk}.adaLN_modulation.0.weight"] = (H, H)
        s[f"{blk}.adaLN_modulation.0.bias"] = (H,)
        s[f"{blk}.adaLN_modulation.2.weight"] = (6 * H, H)
        s[f"{blk}.adaLN_modulation.2.bias"] = (6 * H,)
        attn(f"{blk}.attn", True)
        s[f"{blk}.mlp.0.weight"] = (4 * H, H)
        s[f"{blk}.mlp.0.bias"] = (4 * H,)
        s[f"{blk}.mlp.3.weight"] = (H, 4 * H)
        s[f"{blk}.mlp.3.bias"] = (H,)
    for stack, cross in (("encoder", False), ("decoder", True)):
        for i in range(cfg.num_hidden_layers):
            p = f"{stack}.layer.{i}"
            attn(f"{p}.attention", True)
            if cross:
                attn(f"{p}.crossattention", False)
            s[f"{p}.intermediate.dense.weight"] = (I, H)
            s[f"{p}.intermediate.dense.bias"] = (I,)
            s[f"{p}.output.dense.weight"] = (H, I)
            s[f"{p}.output.dense.bias"] = (H,)
            s[f"{p}.output.LayerNorm.weight"] = (H,)
            s[f"{p}.output.LayerNorm.bias"] = (H,)
    s["angles_predictor.dense1.weight"] = (H, H)
    s["angles_predictor.dense1.bias"] = (H,)
    s["angles_predictor.layer_norm.weight"] = (H,)
    s["angles_predictor.layer_norm.bias"] = (H,)
    s["angles_predictor.dense2.weight"] = (Fs, H)
    s["angles_predictor.dense2.bias"] = (Fs,)
    return s


def init_struct_state_dict(cfg: OracleConfig, seed: int = 0) -> Dict[str, Tensor]:
    """Synthetic weights that exercise every path (the reference zero-initialises adaLN_modulation.0, model.py:49-50, which
    would make both SELayers ignore their conditioning): xavier Linear weights, biases N(0,0.02), LayerNorm 1+N(0,0.1) /
    N(0,0.1), distance_embedding N(0,0.5), W = randn * 2pi (model.py:81)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for name, shape in struct_state_dict_schema(cfg).items():
        if name == "timestep_projector.W":
            t = torch.randn(shape, generator=g) * (2 * math.pi)
        elif name.endswith("distance_embedding.weight"):
            t = torch.randn(shape, generator=g) * 0.5
        elif "LayerNorm" in name or "layer_norm" in name:
            t = (torch.ones(shape) if name.endswith("weight") else torch.zeros(shape)) + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        else:
            fan_out, fan_in = shape
            t = (torch.rand(shape, generator=g) * 2 - 1) * math.sqrt(6.0 / (fan_in + fan_out))
        sd[name] = t.float().contiguous()
    return sd


def encoder_layer(sd, cfg, prefix, h, add_mask):
    """HF BertLayer without cross-attention (the receptor encoder, model.py:177): self-attention -> FFN, post-LN."""
    h1 = O.bert_attention(sd, cfg, prefix + ".attention", h, add_mask)
    inter = F.gelu(O._linear(sd, prefix + ".intermediate.dense", h1))
    out = O._linear(sd, prefix + ".output.dense", inter)
    return O._ln(sd, prefix + ".output.LayerNorm", out + h1, cfg.layer_norm_eps)


def struct_encode(sd, cfg: OracleConfig, receptor_seq, receptor_angles, receptor_attention_masks):
    """Receptor branch, model.py:192-202.  Independent of the timestep and of the ligand: a sampler may compute it once."""
    eps = cfg.layer_norm_eps
    rm = O.extend_mask(receptor_attention_masks)
    r_ang = O.bert_embeddings(sd, "receptor_angle_emb", receptor_angles, eps)
    r_seq = O.bert_embeddings(sd, "receptor_seq_emb", receptor_seq, eps)
    h = O.se_layer(sd, cfg, "receptor_emb", r_ang, r_seq, rm)  # x = angles, c = sequence (model.py:195-198)
    for i in range(cfg.num_hidden_layers):
        h = encoder_layer(sd, cfg, f"encoder.layer.{i}", h, rm)
    return h


def struct_forward(sd: Dict[str, Tensor], cfg: OracleConfig, timestep, noised_ligand_angles, ligand_attention_masks, receptor_seq,
                   receptor_angles, receptor_attention_masks, encoder_outputs: Optional[Tensor] = None):
    """ConditionalBertForDiffusionBase.forward, structure_model/model.py:180-215 (eval mode).  timestep: [B] (long in the
    sampler, sample.py:136) or [B,1]."""
    eps = cfg.layer_norm_eps
    lm = O.extend_mask(ligand_attention_masks)
    rm = O.extend_mask(receptor_attention_masks)
    enc = struct_encode(sd, cfg, receptor_seq, receptor_angles, receptor_attention_masks) if encoder_outputs is None else encoder_outputs
    x = O.bert_embeddings(sd, "ligand_angle_emb", noised_ligand_angles, eps)
    te = O.timestep_embedding(sd, timestep.squeeze(dim=-1)).unsqueeze(1)  # [B,1,H]; int64 * f32 -> f32
    h = O.se_layer(sd, cfg, "timestep_emb", x, te, lm)
    for i in range(cfg.num_hidden_layers):
        h = O.bert_layer(sd, cfg, f"decoder.layer.{i}", h, lm, enc, rm)
    p = "angles_predictor"
    y = F.gelu(O._linear(sd, p + ".dense1", h))
    y = O._ln(sd, p + ".layer_norm", y, 1e-12)  # AnglesPredictor eps default, model.py:134
    return O._linear(sd, p + ".dense2", y)


# --------------------------------------------------------------------------------------------
# Gaussian schedule + reverse step
# --------------------------------------------------------------------------------------------
def cosine_beta_schedule(timesteps: int, s: float = 8e-3) -> Tensor:
    """structure_model/utils.py:8-18 (all fp32 torch ops, as in the reference)."""
    steps = timesteps + 1
    x = torch.linspace(0, timesteps, steps)
    ac = torch.cos(((x / timesteps) + s) / (1 + s) * torch.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = 1 - (ac[1:] / ac[:-1])
    return torch.clip(betas, 0.0001, 0.9999)


def compute_alphas(betas: Tensor) -> Dict[str, Tensor]:
    """structure_model/utils.py:42-58."""
    alphas = 1.0 - betas
    ac = torch.cumprod(alphas, dim=0)
    ac_prev = F.pad(ac[:-1], (1, 0), value=1.0)
    return {
        "betas": betas,
        "alphas": alphas,
        "alphas_cumprod": ac,
        "sqrt_alphas_cumprod": torch.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - ac),
        "posterior_variance": betas * (1.0 - ac_prev) / (1.0 - ac),
    }


def modulo_with_wrapped_range(vals, range_min: float = -math.pi, range_max: float = math.pi):
    """structure_model/utils.py:20-40: ((v - min) % (max - min)) + min, python-float bounds, torch `%` (floor mod)."""
    top_end = range_max - range_min
    return (vals - range_min) % top_end + range_min


def step_coefficients(betas: Tensor) -> Tensor:
    """[T,4] fp32 per step index: (sqrt_recip_alphas, betas, sqrt_one_minus_alphas_cumprod, sqrt(posterior_variance)) exactly as
    p_sample derives them (sample.py:72-85,97-101)."""
    ab = compute_alphas(betas)
    return torch.stack([1.0 / torch.sqrt(ab["alphas"]), betas, ab["sqrt_one_minus_alphas_cumprod"], torch.sqrt(ab["posterior_variance"])], dim=1)


def p_sample_update(x_t: Tensor, model_output: Tensor, coef: Tensor, t_index: int, noise: Optional[Tensor]) -> Tensor:
    """sample.py:92-101 on an already computed model output; the wrap of p_sample_loop (sample.py:139-141) is NOT applied."""
    a, b, c, sd_ = coef[t_index]
    mean = a * (x_t - b * model_output / c)
    if t_index == 0:
        return mean
    return mean + sd_ * noise


def p_sample_loop(sd, cfg, ligand_mask, x_T, receptor_seq, receptor_mask, receptor_angle, total_timesteps: int, betas: Tensor,
                  noise_fn, cache_encoder: bool = True):
    """sample.py:104-144 (STEP = 1): returns [T, B, L, F], entry k = the wrapped angles after the k-th reverse step.
    noise_fn(t_index) -> N(0,1) tensor like x (called for t_index > 0 only)."""
    coef = step_coefficients(betas)
    enc = struct_encode(sd, cfg, receptor_seq, receptor_angle, receptor_mask) if cache_encoder else None
    x = x_T
    B = x.shape[0]
    outs = []
    for i in reversed(range(total_timesteps)):
        t = torch.full((B,), i, dtype=torch.long)
        out = struct_forward(sd, cfg, t, x, ligand_mask, receptor_seq, receptor_angle, receptor_mask, encoder_outputs=enc)
        x = p_sample_update(x, out, coef, i, noise_fn(i) if i > 0 else None)
        x = modulo_with_wrapped_range(x, -math.pi, math.pi)
        outs.append(x)
    return torch.stack(outs)
