"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

A plain-PyTorch (CPU, fp32) restatement of the reference algorithm for the one hot path this
repository accelerates: the `sequence_model` denoiser forward and the discrete BLOSUM reverse
step.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module, and only as the checker / the CPU arm -- never as the product path.
The product (`e3-invaraint-diffusion-model_b200`) fails loudly if its CUDA library is missing.

It needs neither `/root/reference` nor `transformers`, so it travels to the GPU box.

Parity pin (see oracle/make_golden.py, tests/test_oracle_pin.py, DESIGN.md section "Oracle"):
  * everything except the relative_key term is pinned against the unmodified reference imported
    in place (oracle/ref_import.py) -- forward logits with distance_embedding == 0, the reverse
    step, schedules, transition tables: golden vectors in tests/golden/.
  * the `relative_key` arithmetic lives in HuggingFace transformers 4.38.2
    (environment.yml:229), which is NOT vendored under /root/reference and is not installable
    here (no network; 5.5.0 installed, which dropped that code path).  It is restated from the
    published 4.38.2 semantics and cross-checked against a naive einsum:
    for that ONE term the parity is "unpinned" by any reference artefact.

All citations are to files under /root/reference/.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
NUM_CLASSES = 20


# --------------------------------------------------------------------------------------------
# configuration + state-dict schema
# --------------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    """The BertConfig fields the path consumes (sequence_model/sample.py:69-92)."""

    hidden_size: int = 768
    num_attention_heads: int = 12
    intermediate_size: int = 1024
    num_hidden_layers: int = 6
    max_position_embeddings: int = 128
    layer_norm_eps: float = 1e-12
    feature_size: int = 20
    relative_key: bool = True  # position_embedding_type == "relative_key" (4.38.2 semantics)

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads


_SE_BLOCKS = ("ligand_feature_emb", "receptor_feature_emb", "decoder_normalize")


def state_dict_schema(cfg: OracleConfig) -> Dict[str, tuple]:
    """name -> shape for every tensor of `ConditionalBertForDiffusionBase.state_dict()`
    (sequence_model/model.py:156-181), with the nine 4.38.2 distance_embedding tensors when
    cfg.relative_key."""
    H, I, P = cfg.hidden_size, cfg.intermediate_size, cfg.max_position_embeddings
    dh = cfg.head_dim
    s: Dict[str, tuple] = {"timestep_projector.W": (H // 2,)}
    for side in ("ligand", "receptor"):
        for kind, fin in (("seq", 20), ("angle", 8)):
            p = f"{side}_{kind}_embedding"
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
    for i in range(cfg.num_hidden_layers):
        p = f"decoder.layer.{i}"
        attn(f"{p}.attention", True)
        attn(f"{p}.crossattention", False)
        s[f"{p}.intermediate.dense.weight"] = (I, H)
        s[f"{p}.intermediate.dense.bias"] = (I,)
        s[f"{p}.output.dense.weight"] = (H, I)
        s[f"{p}.output.dense.bias"] = (H,)
        s[f"{p}.output.LayerNorm.weight"] = (H,)
        s[f"{p}.output.LayerNorm.bias"] = (H,)
    s["amino_acid_predictor.dense1.weight"] = (H, H)
    s["amino_acid_predictor.dense1.bias"] = (H,)
    s["amino_acid_predictor.layer_norm.weight"] = (H,)
    s["amino_acid_predictor.layer_norm.bias"] = (H,)
    s["amino_acid_predictor.dense2.weight"] = (cfg.feature_size, H)
    s["amino_acid_predictor.dense2.bias"] = (cfg.feature_size,)
    return s


def init_state_dict(cfg: OracleConfig, seed: int = 0, variant: str = "A") -> Dict[str, Tensor]:
    """Synthetic weights (no checkpoint is reachable: README.md:5-6).

    variant "A": the reference's own init rule (model.py:183-198): xavier-uniform Linear weights,
        zero biases, LayerNorm (1,0), distance_embedding N(0,1), timestep W = randn*2pi
        (model.py:81), decoder_normalize.adaLN_modulation.0 all zero  => that block is the identity.
    variant "B": every path exercised (SURVEY.md section 7 "identity-at-init trap"): xavier weights
        everywhere (incl. decoder_normalize.adaLN_modulation.0), biases N(0,0.02),
        LayerNorm weight 1+N(0,0.1) / bias N(0,0.1), distance_embedding N(0,1)*0.5.
    Not bit-identical to the reference's RNG draw order -- parity tests always hand the SAME
    state_dict to both sides instead."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for name, shape in state_dict_schema(cfg).items():
        if name == "timestep_projector.W":
            t = torch.randn(shape, generator=g) * (2 * math.pi)
        elif name.endswith("distance_embedding.weight"):
            t = torch.randn(shape, generator=g) * (1.0 if variant == "A" else 0.5)
        elif "LayerNorm" in name or "layer_norm" in name:
            if name.endswith("weight"):
                t = torch.ones(shape) + (0.1 * torch.randn(shape, generator=g) if variant == "B" else 0)
            else:
                t = 0.1 * torch.randn(shape, generator=g) if variant == "B" else torch.zeros(shape)
        elif name.endswith("bias"):
            t = 0.02 * torch.randn(shape, generator=g) if variant == "B" else torch.zeros(shape)
        else:  # Linear weight [out, in]
            fan_out, fan_in = shape
            a = math.sqrt(6.0 / (fan_in + fan_out))
            t = (torch.rand(shape, generator=g) * 2 - 1) * a
            if variant == "A" and name.startswith("decoder_normalize.adaLN_modulation.0"):
                t = torch.zeros(shape)
        sd[name] = t.float().contiguous()
    return sd


# --------------------------------------------------------------------------------------------
# denoiser forward  (sequence_model/model.py:200-237)
# --------------------------------------------------------------------------------------------
def _linear(sd, prefix, x):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def _ln(sd, prefix, x, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def timestep_embedding(sd, t: Tensor) -> Tensor:
    """GaussianFourierProjection.forward, model.py:85-97.  Multiply order ((t*W)*2)*pi in fp32."""
    if t.ndim > 1:
        t = t.squeeze()
    if t.ndim < 1:
        t = t.unsqueeze(0)
    x_proj = t[:, None] * sd["timestep_projector.W"][None, :] * 2 * torch.pi
    return torch.cat([torch.sin(x_proj), torch.cos(x_proj)], dim=-1)


def extend_mask(mask: Tensor) -> Tensor:
    """_exetend_attention_mask, model.py:248-253: additive -10000 (not -inf), [B,1,1,L]."""
    return (1.0 - mask[:, None, None, :].type_as(mask)) * -10000.0


def bert_embeddings(sd, prefix, x, eps):
    """BertEmbeddings.forward, model.py:110-117 (dropout = identity in eval)."""
    return _ln(sd, prefix + ".LayerNorm", _linear(sd, prefix + ".linear", x), eps)


def attention_core(cfg: OracleConfig, q, k, v, add_mask, dist_emb: Optional[Tensor]):
    """BertSelfAttention of transformers 4.38.2 (Appendix A).  q:[B,Lq,H] k,v:[B,Lk,H]."""
    B, Lq, H = q.shape
    Lk = k.shape[1]
    nh, dh = cfg.num_attention_heads, cfg.head_dim
    qh = q.view(B, Lq, nh, dh).permute(0, 2, 1, 3)
    kh = k.view(B, Lk, nh, dh).permute(0, 2, 1, 3)
    vh = v.view(B, Lk, nh, dh).permute(0, 2, 1, 3)
    s = qh @ kh.transpose(-1, -2)
    if dist_emb is not None:
        P = cfg.max_position_embeddings
        dist = torch.arange(Lq).view(-1, 1) - torch.arange(Lk).view(1, -1)
        pe = dist_emb[dist + P - 1]  # [Lq, Lk, dh]
        s = s + torch.einsum("bhld,lrd->bhlr", qh, pe)
    s = s / math.sqrt(dh)  # scale applied AFTER adding Rel
    s = s + add_mask
    p = torch.softmax(s, dim=-1)
    ctx = (p @ vh).permute(0, 2, 1, 3).reshape(B, Lq, H)
    return ctx


def bert_attention(sd, cfg, prefix, x, add_mask, kv=None, kv_mask=None, rel=True):
    """HF BertAttention = BertSelfAttention + BertSelfOutput (dense, dropout, LN(out + x))."""
    src = x if kv is None else kv
    q = _linear(sd, prefix + ".self.query", x)
    k = _linear(sd, prefix + ".self.key", src)
    v = _linear(sd, prefix + ".self.value", src)
    de = sd.get(prefix + ".self.distance_embedding.weight") if (rel and cfg.relative_key) else None
    ctx = attention_core(cfg, q, k, v, add_mask if kv is None else kv_mask, de)
    out = _linear(sd, prefix + ".output.dense", ctx)
    return _ln(sd, prefix + ".output.LayerNorm", out + x, cfg.layer_norm_eps)


def se_layer(sd, cfg, prefix, x, c, add_mask):
    """SELayer.forward, model.py:52-66 (adaLN block; norm1/norm2 have no affine, eps 1e-5)."""
    mod = _linear(sd, prefix + ".adaLN_modulation.2", F.silu(_linear(sd, prefix + ".adaLN_modulation.0", c)))
    sh1, sc1, g1, sh2, sc2, g2 = mod.chunk(6, dim=-1)
    H = x.shape[-1]
    a = bert_attention(sd, cfg, prefix + ".attn", x, add_mask)
    x = x + g1 * (F.layer_norm(a, (H,)) * (1 + sc1) + sh1)
    m = _linear(sd, prefix + ".mlp.3", F.gelu(_linear(sd, prefix + ".mlp.0", x)))
    x = x + g2 * (F.layer_norm(m, (H,)) * (1 + sc2) + sh2)
    return x


def bert_layer(sd, cfg, prefix, h, add_mask, enc, enc_mask):
    """HF BertLayer with is_decoder + add_cross_attention (Appendix A): self (bidirectional --
    BertEncoder is called with the caller's mask, so no causal mask) -> cross (absolute => no Rel)
    -> FFN, post-LN."""
    h1 = bert_attention(sd, cfg, prefix + ".attention", h, add_mask)
    h2 = bert_attention(sd, cfg, prefix + ".crossattention", h1, None, kv=enc, kv_mask=enc_mask, rel=False)
    inter = F.gelu(_linear(sd, prefix + ".intermediate.dense", h2))
    out = _linear(sd, prefix + ".output.dense", inter)
    return _ln(sd, prefix + ".output.LayerNorm", out + h2, cfg.layer_norm_eps)


def denoiser_forward(sd: Dict[str, Tensor], cfg: OracleConfig, timestep, noised_ligand_seq, ligand_angle,
                     ligand_attention_masks, receptor_seq, receptor_angle, receptor_attention_masks,
                     return_intermediates: bool = False):
    """ConditionalBertForDiffusionBase.forward, model.py:200-237 (eval mode)."""
    eps = cfg.layer_norm_eps
    inter = {}
    lm = extend_mask(ligand_attention_masks)
    rm = extend_mask(receptor_attention_masks)
    te = timestep_embedding(sd, timestep.squeeze(dim=-1)).unsqueeze(1)  # [B,1,H]
    h_seq = bert_embeddings(sd, "ligand_seq_embedding", noised_ligand_seq, eps)
    c_lig = bert_embeddings(sd, "ligand_angle_embedding", ligand_angle, eps) + te
    lig = se_layer(sd, cfg, "ligand_feature_emb", h_seq, c_lig, lm)
    r_seq = bert_embeddings(sd, "receptor_seq_embedding", receptor_seq, eps)
    c_rec = bert_embeddings(sd, "receptor_angle_embedding", receptor_angle, eps) + te
    rec = se_layer(sd, cfg, "ligand_feature_emb", r_seq, c_rec, rm)  # quirk Q1: ligand block reused (model.py:221)
    inter.update(te=te, h_seq=h_seq, c_lig=c_lig, lig=lig, rec=rec)
    h = lig
    for i in range(cfg.num_hidden_layers):
        h = bert_layer(sd, cfg, f"decoder.layer.{i}", h, lm, rec, rm)
        inter[f"dec{i}"] = h
    h = se_layer(sd, cfg, "decoder_normalize", h, te, lm)
    inter["dec_norm"] = h
    p = "amino_acid_predictor"
    y = _linear(sd, p + ".dense1", h)
    y = F.gelu(y)
    y = _ln(sd, p + ".layer_norm", y, 1e-12)  # AminoAcidPredictor eps default, model.py:134,145
    logits = _linear(sd, p + ".dense2", y)
    if return_intermediates:
        return logits, inter
    return logits


# --------------------------------------------------------------------------------------------
# schedules and transitions (sequence_model/utils.py)
# --------------------------------------------------------------------------------------------
def cosine_beta_schedule_discrete(timesteps: int, s: float = 0.008) -> np.ndarray:
    """utils.py:99-108 (float64)."""
    steps = timesteps + 2
    x = np.linspace(0, steps, steps)
    ac = np.cos(0.5 * np.pi * ((x / steps) + s) / (1 + s)) ** 2
    ac = ac / ac[0]
    alphas = ac[1:] / ac[:-1]
    return (1 - alphas).squeeze()


class NoiseScheduleDiscrete:
    """PredefinedNoiseScheduleDiscrete, utils.py:206-233."""

    def __init__(self, noise_schedule: str, timesteps: int):
        self.timesteps = timesteps
        self.betas = torch.from_numpy(cosine_beta_schedule_discrete(timesteps)).float()
        self.alphas = 1 - torch.clamp(self.betas, min=0, max=0.9999)
        self.alphas_bar = torch.exp(torch.cumsum(torch.log(self.alphas), dim=0))

    def get_alpha_bar(self, t_normalized=None, t_int=None):
        assert int(t_normalized is None) + int(t_int is None) == 1
        if t_int is None:
            t_int = torch.round(t_normalized * self.timesteps)
        return self.alphas_bar[t_int.long()]


_BLOSUM_NPZ = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "blosum_tables.npz")


def load_blosum_tables(path: Optional[str] = None):
    """The three arrays of sequence_model/blosum_substitute.pt (fixture; values re-exported by
    oracle/make_golden.py -- sha256 of the source file recorded in the npz)."""
    z = np.load(path or _BLOSUM_NPZ)
    return (torch.from_numpy(z["original_score"]), torch.from_numpy(z["Qtb_temperature"]),
            torch.from_numpy(z["Qt_temperature"]))


class BlosumTransition:
    """utils.py:273-314.  Note the always-true shape test at :286 => both temperature tables are
    linearly re-interpolated 500 -> timestep+1 entries (align_corners=True)."""

    def __init__(self, tables=None, x_classes: int = 20, timestep: int = 500):
        score, temp, qt_temp = tables if tables is not None else load_blosum_tables()
        self.original_score = score.float()
        self.X_classes = x_classes
        self.timestep = timestep
        self.temperature_list = F.interpolate(temp.float()[None, None], size=timestep + 1, mode="linear",
                                              align_corners=True).squeeze()
        self.Qt_temperature = F.interpolate(qt_temp.float()[None, None], size=timestep + 1, mode="linear",
                                            align_corners=True).squeeze()

    def get_Qt_bar(self, t_normal, device=None):
        t_int = torch.round(t_normal * self.timestep)
        temp = self.temperature_list[t_int.long()]  # [B,1]
        q = self.original_score.unsqueeze(0) / temp.unsqueeze(2)
        q = torch.softmax(q, dim=2)
        q[q < 1e-6] = 1e-6
        return q

    def get_Qt(self, t_normal, device=None):
        t_int = torch.round(t_normal * self.timestep)
        temp = self.Qt_temperature[t_int.long()]
        return torch.softmax(self.original_score.unsqueeze(0) / temp.unsqueeze(2), dim=2)


class DiscreteUniformTransition:
    """utils.py:235-271."""

    def __init__(self, x_classes: int = 20):
        self.X_classes = x_classes
        self.u_x = torch.ones(1, x_classes, x_classes) / x_classes

    def get_Qt(self, beta_t, device=None):
        beta_t = beta_t.unsqueeze(1)
        return beta_t * self.u_x + (1 - beta_t) * torch.eye(self.X_classes).unsqueeze(0)

    def get_Qt_bar(self, alpha_bar_t, device=None):
        a = alpha_bar_t.unsqueeze(1)
        return a * torch.eye(self.X_classes).unsqueeze(0) + (1 - a) * self.u_x


# --------------------------------------------------------------------------------------------
# reverse step (sequence_model/sample.py:120-179)
# --------------------------------------------------------------------------------------------
def step_matrices(t, s, noise_schedule, transition):
    """sample.py:156-160.  quirk Q2: get_Qt_bar is fed alpha_bar, not t.  Returns Qt,Qsb,Qtb [B,20,20]."""
    a_t = noise_schedule.get_alpha_bar(t_normalized=t)
    a_s = noise_schedule.get_alpha_bar(t_normalized=s)
    Qtb = transition.get_Qt_bar(a_t, None)
    Qsb = transition.get_Qt_bar(a_s, None)
    Qt = (Qsb / Qtb) / (Qsb / Qtb).sum(dim=-1).unsqueeze(dim=2)
    return Qt, Qsb, Qtb


def posterior_over0(X_t, Q_t, Qsb, Qtb, batch):
    """compute_batched_over0_posterior_distribution, sample.py:120-139."""
    Qt_T = Q_t.transpose(-1, -2)
    left = X_t.unsqueeze(-2) @ Qt_T[batch]  # [N,1,d]
    num = left * Qsb[batch]  # [N,d0,d]
    den = Qtb[batch] @ X_t.unsqueeze(2)  # [N,d0,1]
    den[den == 0] = 1e-6
    return num / den


def reverse_step_probs(t, s, noised_data, pred_noise, noise_schedule, transition):
    """sample.py:149-168 up to the normalised posterior prob_X [N,20]."""
    B, L, C = noised_data.shape
    batch = torch.arange(B).repeat_interleave(L)
    x = noised_data.reshape(B * L, C)
    logits = pred_noise.reshape(B * L, C)
    Qt, Qsb, Qtb = step_matrices(t, s, noise_schedule, transition)
    pred = F.softmax(logits, dim=-1)
    post = posterior_over0(x, Qt, Qsb, Qtb, batch)
    un = (pred.unsqueeze(-1) * post).sum(dim=1)
    un[torch.sum(un, dim=-1) == 0] = 1e-5
    return un / torch.sum(un, dim=-1, keepdim=True)


def sample_indices(prob_X: Tensor, diverse: bool, noise_E: Optional[Tensor]) -> Tensor:
    """sample.py:169-177, vectorised.  `prob.multinomial(1)` on CPU is the exponential race
    argmax(prob / E), E ~ Exp(1) (probe) -- `noise_E` [N,20] is that E, handed
    to both sides of a parity test.  Rows with zero mass -> class 0."""
    if diverse:
        idx = (prob_X / noise_E).argmax(dim=-1)
    else:
        idx = prob_X.argmax(dim=-1)
    return torch.where(prob_X.sum(-1) != 0, idx, torch.zeros_like(idx))


def reverse_step(t, s, noised_data, pred_noise, noise_schedule, transition, diverse, is_last_step,
                 noise_E: Optional[Tensor] = None):
    """sample_p_zs_given_zt_discrete, sample.py:141-179, with the noise made explicit."""
    if is_last_step:
        return pred_noise  # quirk Q4: raw logits
    B, L, C = noised_data.shape
    prob = reverse_step_probs(t, s, noised_data, pred_noise, noise_schedule, transition)
    idx = sample_indices(prob, diverse, noise_E)
    return F.one_hot(idx.reshape(B, L), num_classes=C).float()


def reverse_step_python_loop(t, s, noised_data, pred_noise, noise_schedule, transition, diverse, is_last_step):
    """The reference's own control flow, Python row loop included (sample.py:150,169-178); draws from
    torch's global CPU generator like the reference.  Used as the CPU baseline ("port")."""
    if is_last_step:
        return pred_noise
    B, L, C = noised_data.shape
    _ = torch.tensor([[i] * L for i in range(B)]).reshape(-1)  # sample.py:150 cost kept
    prob_X = reverse_step_probs(t, s, noised_data, pred_noise, noise_schedule, transition)
    out = []
    for prob in prob_X:
        if prob.sum() != 0 and diverse:
            out.append(prob.multinomial(1)[0])
        elif prob.sum() != 0 and (not diverse):
            out.append(prob.argmax())
        else:
            out.append(0)
    idx = torch.Tensor(out).reshape(B, L).long()
    return F.one_hot(idx, num_classes=C).float()


def generate_discrete_noise(batch_size, length, num_classes=20, generator=None):
    """sample.py:112-116."""
    idx = torch.randint(0, num_classes, (batch_size, length), generator=generator)
    return F.one_hot(idx, num_classes).float()


def denoise_loop(sd, cfg, batch, noise_schedule, transition, diverse, timesteps, x_T, noise_fn=None,
                 python_loop=False):
    """denoise, sample.py:181-207 (the T-step loop; decode/recovery omitted).  quirk Q3: the raw
    integer step is fed to forward.  noise_fn(step) -> E [N,20]."""
    x = x_T
    B = x.shape[0]
    for s_int in reversed(range(timesteps)):
        s_array = s_int * torch.ones((B, 1))
        s_norm = s_array / timesteps
        t_norm = (s_array + 1) / timesteps
        logits = denoiser_forward(sd, cfg, s_array, x, batch["ligand_angles"], batch["ligand_attn_mask"],
                                  batch["receptor_seq"], batch["receptor_angles"], batch["receptor_attn_mask"])
        if python_loop:
            x = reverse_step_python_loop(t_norm, s_norm, x, logits, noise_schedule, transition, diverse, s_int == 0)
        else:
            E = noise_fn(s_int) if (diverse and noise_fn is not None and s_int != 0) else None
            x = reverse_step(t_norm, s_norm, x, logits, noise_schedule, transition, diverse, s_int == 0, E)
    return x


# --------------------------------------------------------------------------------------------
# training-side pieces (model.py:291-345, utils.py:132-161) -- "next" rows of SURVEY.md section 8f
# --------------------------------------------------------------------------------------------
def apply_aa_noise_probs(ligand_seq, t_int, timesteps, noise_schedule, transition):
    """model.py:291-301: prob[n,i] = Qtb[b,i,x0[n]] (a column of Qbar_t)."""
    B, L, C = ligand_seq.shape
    x = ligand_seq.reshape(B * L, C)
    batch = torch.arange(B).repeat_interleave(L)
    a = noise_schedule.get_alpha_bar(t_normalized=t_int / timesteps)
    Qtb = transition.get_Qt_bar(a, None)[batch]
    return (Qtb @ x.unsqueeze(2)).squeeze(-1)


def apply_aa_noise(ligand_seq, t_int, timesteps, noise_schedule, transition, noise_E):
    """model.py:291-311 with explicit exponential-race noise; zero rows (padding) -> class 0."""
    B, L, C = ligand_seq.shape
    prob = apply_aa_noise_probs(ligand_seq, t_int, timesteps, noise_schedule, transition)
    idx = sample_indices(prob, True, noise_E)
    return F.one_hot(idx.reshape(B, L), C).float()


def elbo_loss(logits1, logits2, eps=1e-6):
    """utils.py:132-161."""
    probs1 = F.softmax(logits1, dim=-1)
    probs2 = F.softmax(logits2, dim=-1)
    lp1 = F.log_softmax(logits1 + eps, dim=-1)
    kl = F.kl_div(lp1, probs2, reduction="batchmean")
    nll = -torch.mean(torch.sum(probs1 * lp1, dim=-1))
    return nll + kl


def get_loss(pred_aa, batch, noised_ligand_seq):
    """PeptideDiff.get_loss, model.py:313-345, on already computed logits `pred_aa` (loss_function = CrossEntropyLoss()).
    Returns (total_loss, elbo, aa_noised_loss, aa_all_loss, aa_recovery_rate, aa_noise_rate)."""
    ligand_mask = batch["ligand_attn_mask"].bool()
    x0_idx = batch["ligand_seq"].argmax(dim=-1)
    noised_mask = noised_ligand_seq.argmax(dim=-1) != x0_idx
    aa_noise_rate = (noised_ligand_seq.argmax(dim=-1)[ligand_mask] == batch["ligand_seq"][ligand_mask].argmax(dim=-1)).sum() / ligand_mask.sum()
    aa_recovery_rate = (pred_aa.argmax(dim=-1)[ligand_mask] == batch["ligand_seq"][ligand_mask].argmax(dim=-1)).sum() / ligand_mask.sum()
    aa_noised_loss = F.cross_entropy(pred_aa[noised_mask].view(-1, 20), batch["ligand_seq"][noised_mask].argmax(dim=-1).view(-1))
    sel = ligand_mask & (~noised_mask)
    aa_all_loss = F.cross_entropy(pred_aa[sel].view(-1, 20), batch["ligand_seq"][sel].argmax(dim=-1).view(-1))
    elbo = elbo_loss(pred_aa[noised_mask], batch["ligand_seq"][noised_mask])
    return aa_noised_loss + elbo, elbo, aa_noised_loss, aa_all_loss, aa_recovery_rate, aa_noise_rate


# --------------------------------------------------------------------------------------------
# training step (sequence_model/model.py:347-367 under autograd; train_model.py:30-33,95)
# --------------------------------------------------------------------------------------------
def training_grads(sd: Dict[str, Tensor], cfg: OracleConfig, batch, t_norm, noised_ligand_seq):
    """What `loss = training_step(batch); loss.backward()` leaves behind in the reference with dropout off (eval-mode arithmetic):
    (loss 6-tuple, {state_dict key: gradient or None}).  torch autograd on the oracle forward: `receptor_feature_emb.*` never
    enters the graph (model.py:221, quirk Q1) and `timestep_projector.W` is a buffer -> None."""
    leaf = {k: v.detach().clone().requires_grad_(k != "timestep_projector.W") for k, v in sd.items()}
    logits = denoiser_forward(leaf, cfg, t_norm, noised_ligand_seq, batch["ligand_angles"], batch["ligand_attn_mask"], batch["receptor_seq"],
                              batch["receptor_angles"], batch["receptor_attn_mask"])
    out = get_loss(logits, batch, noised_ligand_seq)
    out[0].backward()
    return tuple(o.detach() for o in out), {k: (v.grad.detach() if v.grad is not None else None) for k, v in leaf.items()}, logits.detach()


def adamw_reference_step(params: Dict[str, Tensor], grads: Dict[str, Optional[Tensor]], state: Dict, lr=5e-5, weight_decay=0.1, betas=(0.9, 0.999),
                         eps=1e-8, max_norm=1.0):
    """torch.nn.utils.clip_grad_norm_(max_norm) followed by ONE torch.optim.AdamW step (the real torch objects) on clones of
    `params`; tensors whose gradient is None are left untouched, as torch does.  `state` carries the optimizer between calls."""
    if "opt" not in state:
        state["p"] = {k: torch.nn.Parameter(v.detach().clone()) for k, v in params.items()}
        state["opt"] = torch.optim.AdamW(list(state["p"].values()), lr=lr, weight_decay=weight_decay, betas=betas, eps=eps)
    for k, prm in state["p"].items():
        prm.grad = None if grads.get(k) is None else grads[k].detach().clone()
    live = [prm for prm in state["p"].values() if prm.grad is not None]
    norm = torch.nn.utils.clip_grad_norm_(live, max_norm) if max_norm and max_norm > 0 else torch.linalg.vector_norm(torch.stack([p.grad.norm() for p in live]))
    for g_ in state["opt"].param_groups:
        g_["lr"] = lr
    state["opt"].step()
    return {k: v.detach().clone() for k, v in state["p"].items()}, float(norm)


def decode(final, batch):
    """The per-graph tail of denoise(), sample.py:208-224: (true_sequences, pred_sequences, recovery_rates)."""
    AA = "ACDEFGHIKLMNPQRSTVWY"
    rates, preds, trues = [], [], []
    for i in range(final.shape[0]):
        pred_seq = final[i].argmax(dim=1)
        true_seq = batch["ligand_seq"][i].argmax(dim=1)
        mask = batch["ligand_attn_mask"][i].bool()
        r = (pred_seq[mask] == true_seq[mask]).sum() / mask.sum()
        rates.append(r.item())
        preds.append("".join(AA[j] for j in pred_seq[mask]))
        trues.append("".join(AA[j] for j in true_seq[mask]))
    return trues, preds, rates


# --------------------------------------------------------------------------------------------
# dataset item construction (sequence_model/dataset.py:41-49, 97-129) -- the "graph construction" analogue
# --------------------------------------------------------------------------------------------
def dataset_item(data: Dict, max_len: int, pocket_ext: int) -> Dict:
    """LigandBindingSiteDataset.__getitem__, dataset.py:97-129."""
    def pad(x):
        if x.shape[0] > max_len:
            raise RuntimeError("Length exceed")
        return F.pad(x, (0, 0, 0, max_len - x.shape[0]), mode="constant", value=0)

    ligand_mask = data["ligand_mask"]
    left = torch.roll(data["pocket_mask"], pocket_ext)
    left[0] = False
    right = torch.roll(data["pocket_mask"], -pocket_ext)
    right[-1] = False
    pocket_mask = data["pocket_mask"] | left | right
    lig_attn = torch.zeros(max_len)
    lig_attn[:ligand_mask.sum()] = 1.0
    rec_attn = torch.zeros(max_len)
    rec_attn[:pocket_mask.sum()] = 1.0
    return {"ligand_angles": pad(data["angle_features"][ligand_mask]), "ligand_attn_mask": lig_attn,
            "ligand_seq": pad(data["amino_acid"][ligand_mask]), "receptor_angles": pad(data["angle_features"][pocket_mask]),
            "receptor_attn_mask": rec_attn, "receptor_seq": pad(data["amino_acid"][pocket_mask]),
            "ligand_length": ligand_mask.sum(), "receptor_length": pocket_mask.sum(), "structure_ids": data["structure_ids"]}


def synthetic_records(n_complex: int, seed: int, n_lo: int = 40, n_hi: int = 260):
    """records with the post-_load_file schema of the reference dataset (dataset.py:68-73): a contiguous ligand chain,
    a sparse pocket on the other chain (incl. first/last residues so the wrap-around quirk Q9 is exercised)."""
    g = torch.Generator().manual_seed(seed)
    recs = []
    for c in range(n_complex):
        n = int(torch.randint(n_lo, n_hi + 1, (1,), generator=g))
        nl = int(torch.randint(5, min(40, n // 2) + 1, (1,), generator=g))
        start = int(torch.randint(0, n - nl + 1, (1,), generator=g))
        lig = torch.zeros(n, dtype=torch.bool)
        lig[start:start + nl] = True
        poc = (torch.rand(n, generator=g) < 0.15) & ~lig
        if c % 2 == 0:
            poc[0] = poc[-1] = True  # wrap-around cases
            poc &= ~lig
        recs.append({"amino_acid": F.one_hot(torch.randint(0, 20, (n,), generator=g), 20).float(),
                     "angle_features": (torch.rand(n, 8, generator=g) * 2 - 1) * math.pi, "ligand_mask": lig, "pocket_mask": poc,
                     "structure_ids": {"pdb_id": f"c{c:03d}", "ligand_chain": "B"}})
    return recs


# --------------------------------------------------------------------------------------------
# synthetic workloads
# --------------------------------------------------------------------------------------------
def synthetic_batch(B: int, L: int, n_lig, n_rec, seed: int):
    """Padded batch dict with the key names of sample.py:183-190 / dataset.py:115-129: one-hots and
    angles zero on padded rows (dataset.py:41-49), masks prefix-ones (dataset.py:110-114), angles
    ~U(-pi,pi) (radians, clean_data/data_preprocessing.py:886).  n_lig/n_rec: int or (lo,hi)."""
    g = torch.Generator().manual_seed(seed)

    def lens(n):
        if isinstance(n, int):
            return torch.full((B,), n, dtype=torch.long)
        lo, hi = n
        return torch.randint(lo, hi + 1, (B,), generator=g)

    nl, nr = lens(n_lig), lens(n_rec)
    pos = torch.arange(L)[None, :]
    lm = (pos < nl[:, None]).float()
    rm = (pos < nr[:, None]).float()

    def side(mask):
        seq = F.one_hot(torch.randint(0, 20, (B, L), generator=g), 20).float() * mask[..., None]
        ang = ((torch.rand(B, L, 8, generator=g) * 2 - 1) * math.pi) * mask[..., None]
        return seq, ang

    lseq, lang = side(lm)
    rseq, rang = side(rm)
    return {
        "ligand_seq": lseq, "ligand_angles": lang, "ligand_attn_mask": lm,
        "receptor_seq": rseq, "receptor_angles": rang, "receptor_attn_mask": rm,
    }
