"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/* by running the UNMODIFIED reference
(imported in place from /root/reference, see oracle/ref_import.py) in this container.

    python -m oracle.make_golden            # rewrites tests/golden/

What each fixture pins:
  blosum_tables.npz        values of sequence_model/blosum_substitute.pt (+ sha256 of the source file)
  schedule_T{50,500}.pt    PredefinedNoiseScheduleDiscrete.alphas_bar, BlosumTransition tables,
                           get_Qt_bar at a few alpha_bar values, DiscreteUniformTransition
  forward_norel_*.pt       logits of the reference forward on installed transformers (no relative_key
                           term = what 5.5.0 computes) -> pins everything but the relative term
  forward_rel_*.pt         logits of the reference module tree with the restated 4.38.2 relative_key
                           self-attention injected (parity UNPINNED for that one term)
  reverse_step_*.pt        reference sample_p_zs_given_zt_discrete under torch.manual_seed(k):
                           inputs, the Exp(1) noise E drawn from the same generator state, outputs
  denoise_T4.pt            the reference denoise() loop end to end (tiny T), same-noise replay
  apply_aa_noise.pt        PeptideDiff.apply_aa_noise (training q-sample) under a fixed seed
  get_loss.pt              PeptideDiff.get_loss with its forward replaced by fixed logits (loss reductions + elbo_loss)
  struct_forward_*.pt      structure_model denoiser forward (norel = as installed, rel = restored relative_key term)
  struct_schedule_T*.pt    structure_model cosine_beta_schedule / compute_alphas tables
  struct_p_sample_loop_T4.pt  the reference's own p_sample_loop under torch.manual_seed, with the N(0,1) draws it made

Weights are never stored: both sides regenerate them with oracle.init_state_dict(cfg, seed, variant)
and the reference model gets them through load_state_dict(strict=True).
"""
from __future__ import annotations

import hashlib
import math
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

from oracle import ref_import as R
from oracle import seqdiff_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def export_blosum():
    src = os.path.join(R.SEQ_DIR, "blosum_substitute.pt")
    d = torch.load(src)
    sha = hashlib.sha256(open(src, "rb").read()).hexdigest()
    np.savez(os.path.join(GOLDEN, "blosum_tables.npz"),
             original_score=d["original_score"].numpy(), Qtb_temperature=d["Qtb_temperature"].numpy(),
             Qt_temperature=d["Qt_temperature"].numpy(), source_sha256=np.array(sha))
    print("blosum sha256", sha)


def golden_schedules():
    _, _, ref_utils = R.load_reference()
    for T in (50, 500):
        sched = ref_utils.PredefinedNoiseScheduleDiscrete("cosine", T)
        trans = R.make_blosum_transition()
        uni = ref_utils.DiscreteUniformTransition(20)
        ab = sched.alphas_bar.clone()
        probe_t = torch.tensor([[0.0], [1.0 / T], [0.5], [(T - 1.0) / T], [1.0]])
        a = sched.get_alpha_bar(t_normalized=probe_t)
        qtb = trans.get_Qt_bar(a, torch.device("cpu")).clone()
        qt = trans.get_Qt(a, torch.device("cpu")).clone()
        uq = uni.get_Qt_bar(a, torch.device("cpu")).clone()
        torch.save({"T": T, "betas": sched.betas.clone(), "alphas_bar": ab, "probe_t": probe_t, "alpha_bar_probe": a,
                    "temperature_list": trans.temperature_list.clone(), "Qt_temperature": trans.Qt_temperature.clone(),
                    "Qtb_probe": qtb, "Qt_probe": qt, "uniform_Qtb_probe": uq},
                   os.path.join(GOLDEN, f"schedule_T{T}.pt"))


FORWARD_CASES = [
    # name, L(max_seq_len), B, n_lig, n_rec, weight seed, variant, input seed, timestep
    ("cfg1_A", 128, 1, 30, 98, 0, "A", 2, 7.0),
    ("cfg1_B", 128, 1, 30, 98, 1, "B", 2, 7.0),
    ("ragged_B", 128, 3, (5, 64), (16, 128), 1, "B", 3, 33.0),
    ("L64_B", 64, 2, (5, 64), (16, 64), 1, "B", 4, 0.5),
]


def golden_forward():
    for rel in (False, True):
        for name, L, B, nl, nr, wseed, variant, iseed, tval in FORWARD_CASES:
            cfg = O.OracleConfig(max_position_embeddings=L, relative_key=rel)
            sd = O.init_state_dict(cfg, wseed, variant)
            m = R.build_reference_model(L, relative_key=rel)
            m.load_state_dict(sd, strict=True)
            batch = O.synthetic_batch(B, L, nl, nr, iseed)
            g = torch.Generator().manual_seed(iseed + 100)
            x_t = F.one_hot(torch.randint(0, 20, (B, L), generator=g), 20).float()
            t = torch.full((B, 1), tval)
            with torch.no_grad():
                y = m(t, x_t, batch["ligand_angles"], batch["ligand_attn_mask"], batch["receptor_seq"],
                      batch["receptor_angles"], batch["receptor_attn_mask"])
                y_o = O.denoiser_forward(sd, cfg, t, x_t, batch["ligand_angles"], batch["ligand_attn_mask"],
                                         batch["receptor_seq"], batch["receptor_angles"], batch["receptor_attn_mask"])
            err = (y - y_o).abs().max().item()
            print(f"forward {'rel' if rel else 'norel'} {name}: max|ref-oracle| = {err:.3e}  max|ref| = {y.abs().max():.3f}")
            assert err < 2e-5, err
            torch.save({"name": name, "L": L, "B": B, "n_lig": nl, "n_rec": nr, "weight_seed": wseed, "variant": variant,
                        "input_seed": iseed, "timestep": tval, "relative_key": rel, "x_t_idx": x_t.argmax(-1).to(torch.uint8),
                        "logits": y.clone()},
                       os.path.join(GOLDEN, f"forward_{'rel' if rel else 'norel'}_{name}.pt"))


def golden_reverse_step():
    _, ref_sample, ref_utils = R.load_reference()
    cases = [("T50_s30", 50, 30, True, "blosum"), ("T50_s1", 50, 1, True, "blosum"), ("T500_s499", 500, 499, True, "blosum"),
             ("T500_s250", 500, 250, True, "blosum"), ("T500_s3_argmax", 500, 3, False, "blosum"),
             ("T50_s20_uniform", 50, 20, True, "uniform"), ("T50_s0_last", 50, 0, True, "blosum")]
    for name, T, s_int, diverse, kind in cases:
        B, L = 4, 48
        sched = ref_utils.PredefinedNoiseScheduleDiscrete("cosine", T)
        trans = R.make_blosum_transition() if kind == "blosum" else ref_utils.DiscreteUniformTransition(20)
        g = torch.Generator().manual_seed(1000 + s_int)
        x_idx = torch.randint(0, 20, (B, L), generator=g)
        x_t = F.one_hot(x_idx, 20).float()
        logits = torch.randn(B, L, 20, generator=g) * 3.0
        s_arr = s_int * torch.ones((B, 1))
        s_norm, t_norm = s_arr / T, (s_arr + 1) / T
        torch.manual_seed(77 + s_int)
        out = ref_sample.sample_p_zs_given_zt_discrete(t_norm, s_norm, x_t.clone(), logits.clone(), sched, trans, diverse, s_int == 0)
        torch.manual_seed(77 + s_int)
        E = torch.empty(B * L, 20).exponential_(1)
        o_sched = O.NoiseScheduleDiscrete("cosine", T)
        o_trans = O.BlosumTransition() if kind == "blosum" else O.DiscreteUniformTransition(20)
        out_o = O.reverse_step(t_norm, s_norm, x_t.clone(), logits.clone(), o_sched, o_trans, diverse, s_int == 0, E)
        prob = O.reverse_step_probs(t_norm, s_norm, x_t.clone(), logits.clone(), o_sched, o_trans)
        same = torch.equal(out, out_o)
        print(f"reverse {name}: oracle == reference: {same}")
        assert same
        Qt, Qsb, Qtb = O.step_matrices(t_norm, s_norm, o_sched, o_trans)
        rec = {"name": name, "T": T, "s_int": s_int, "diverse": diverse, "kind": kind, "x_t_idx": x_idx.to(torch.uint8),
               "logits": logits, "E": E, "prob": prob, "Qt": Qt[0].clone(), "Qsb": Qsb[0].clone(), "Qtb": Qtb[0].clone()}
        rec["out"] = out.clone() if s_int == 0 else out.argmax(-1).to(torch.uint8)
        torch.save(rec, os.path.join(GOLDEN, f"reverse_step_{name}.pt"))


def golden_denoise():
    """reference denoise() (sample.py:181-229) end to end at tiny T, replayed by the oracle with the
    same noise: x_T from randint, then one Exp(1) block of [N,20] per non-final step, all from the
    global CPU generator in the reference's draw order."""
    ref_model, ref_sample, ref_utils = R.load_reference()
    T, B, L = 4, 2, 64
    cfg = O.OracleConfig(max_position_embeddings=L, relative_key=True)
    sd = O.init_state_dict(cfg, 1, "B")
    m = R.build_reference_model(L, relative_key=True)
    m.load_state_dict(sd, strict=True)
    batch = O.synthetic_batch(B, L, (5, 40), (16, 64), 9)
    batch["structure_ids"] = {"pdb_id": ["xxxx"] * B, "ligand_chain": ["A"] * B}
    old = ref_sample.CONFIG["timesteps"]
    ref_sample.CONFIG["timesteps"] = T
    sched = ref_utils.PredefinedNoiseScheduleDiscrete("cosine", T)
    trans = R.make_blosum_transition()
    try:
        torch.manual_seed(5)
        ids, true_seq, pred_seq, rates = ref_sample.denoise(batch, m, sched, trans, True)
    finally:
        ref_sample.CONFIG["timesteps"] = old
    torch.manual_seed(5)
    x_T = O.generate_discrete_noise(B, L, 20)
    noises = {}

    def noise_fn(s_int):
        noises[s_int] = torch.empty(B * L, 20).exponential_(1)
        return noises[s_int]

    with torch.no_grad():
        final = O.denoise_loop(sd, cfg, batch, O.NoiseScheduleDiscrete("cosine", T), O.BlosumTransition(), True, T, x_T, noise_fn)
    AA = "ACDEFGHIKLMNPQRSTVWY"
    pred_o = []
    for i in range(B):
        mask = batch["ligand_attn_mask"][i].bool()
        pred_o.append("".join(AA[j] for j in final[i].argmax(1)[mask]))
    print("denoise ref:", pred_seq, "\n     oracle:", pred_o)
    assert pred_o == pred_seq
    torch.save({"T": T, "B": B, "L": L, "weight_seed": 1, "variant": "B", "batch_seed": 9, "n_lig": (5, 40), "n_rec": (16, 64),
                "x_T_idx": x_T.argmax(-1).to(torch.uint8), "E": torch.stack([noises[s] for s in sorted(noises)]),
                "E_steps": sorted(noises), "pred_sequences": pred_seq, "true_sequences": true_seq,
                "recovery_rates": rates, "final_logits": final.clone()},
               os.path.join(GOLDEN, "denoise_T4.pt"))


def golden_apply_aa_noise():
    ref_model, _, ref_utils = R.load_reference()
    T, B, L = 50, 4, 32

    class _Shim:  # apply_aa_noise only touches these attributes (model.py:291-311)
        timesteps = T
        discrete_noise_schedule = ref_utils.PredefinedNoiseScheduleDiscrete("cosine", T)
        aa_transition_model = R.make_blosum_transition()

    batch = O.synthetic_batch(B, L, (5, 32), (16, 32), 21)
    t_int = torch.tensor([[0.0], [13.0], [37.0], [50.0]])
    torch.manual_seed(11)
    out = ref_model.PeptideDiff.apply_aa_noise(_Shim(), batch["ligand_seq"], t_int)
    # replay: the reference draws 20 exponentials per NON-padded row only (model.py:304-308)
    torch.manual_seed(11)
    x = batch["ligand_seq"].reshape(B * L, 20)
    E = torch.ones(B * L, 20)
    for n in range(B * L):
        if x[n].sum() != 0:
            E[n] = torch.empty(20).exponential_(1)
    out_o = O.apply_aa_noise(batch["ligand_seq"], t_int, T, O.NoiseScheduleDiscrete("cosine", T), O.BlosumTransition(), E)
    same = torch.equal(out, out_o)
    print("apply_aa_noise oracle == reference:", same)
    assert same
    torch.save({"T": T, "B": B, "L": L, "batch_seed": 21, "t_int": t_int, "E": E, "x0_idx": batch["ligand_seq"].argmax(-1).to(torch.uint8),
                "x0_valid": batch["ligand_attn_mask"].to(torch.uint8), "out_idx": out.argmax(-1).to(torch.uint8)},
               os.path.join(GOLDEN, "apply_aa_noise.pt"))


def golden_dataset():
    """reference LigandBindingSiteDataset.__getitem__ (dataset.py:81-129) on synthetic records, ext in {0,1,3}."""
    R.load_reference()
    ref_ds = sys.modules["_ref_seq_dataset"]
    recs = O.synthetic_records(6, 41)
    out = {}
    for ext, max_len in ((0, 128), (1, 128), (3, 160)):
        ds = ref_ds.LigandBindingSiteDataset.__new__(ref_ds.LigandBindingSiteDataset)
        ds.data, ds.max_len, ds.pocket_ext = recs, max_len, ext
        items = [ds[i] for i in range(len(recs))]
        for i, it in enumerate(items):
            o = O.dataset_item(recs[i], max_len, ext)
            for k in ("ligand_angles", "ligand_attn_mask", "ligand_seq", "receptor_angles", "receptor_attn_mask", "receptor_seq"):
                assert torch.equal(it[k], o[k]), (ext, i, k)
            assert int(it["ligand_length"]) == int(o["ligand_length"]) and int(it["receptor_length"]) == int(o["receptor_length"])
        out[(ext, max_len)] = {"ligand_length": torch.tensor([int(it["ligand_length"]) for it in items]),
                               "receptor_length": torch.tensor([int(it["receptor_length"]) for it in items]),
                               "receptor_seq_idx": torch.stack([it["receptor_seq"].argmax(-1).to(torch.uint8) for it in items]),
                               "receptor_angle_sum": torch.stack([it["receptor_angles"].double().sum(-1) for it in items]),
                               "ligand_seq_idx": torch.stack([it["ligand_seq"].argmax(-1).to(torch.uint8) for it in items])}
    print("dataset items: oracle == reference for ext 0/1/3")
    torch.save({"n_complex": 6, "seed": 41, "cases": out}, os.path.join(GOLDEN, "dataset_items.pt"))


def golden_structure_feed():
    """BASELINE configs[4]: the reference structure_model denoiser (12 + 12 layers, feature_size 8) on 128-residue synthetic
    torsion sequences; its output [B, 128, 8] f32 is what sample_by_generated_angles.py:202 hands the sequence model as
    `ligand_angle`.  Stored: the generated angles (wrapped to [-pi, pi) like structure_model/sample.py:139-141)."""
    from transformers.models.bert.modeling_bert import BertConfig
    SM = R.load_structure_reference()
    B, L = 8, 128
    torch.manual_seed(21)
    common = dict(max_position_embeddings=L, num_attention_heads=12, hidden_size=768, intermediate_size=1024, num_hidden_layers=12,
                  position_embedding_type="relative_key", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)
    enc = BertConfig(**common)
    dec = BertConfig(**common, is_decoder=True, add_cross_attention=True)
    m = SM.ConditionalBertForDiffusionBase(enc, dec, 8).eval()
    batch = O.synthetic_batch(B, L, (5, 64), (16, 128), 55)
    g = torch.Generator().manual_seed(56)
    noised = (torch.rand(B, L, 8, generator=g) * 2 - 1) * math.pi * batch["ligand_attn_mask"][..., None]
    t = torch.randint(0, 1000, (B,), generator=g)
    with torch.no_grad():
        out = m(t, noised, batch["ligand_attn_mask"], batch["receptor_seq"], batch["receptor_angles"], batch["receptor_attn_mask"])
    assert out.shape == (B, L, 8) and out.dtype == torch.float32 and torch.isfinite(out).all()
    wrapped = torch.remainder(out + math.pi, 2 * math.pi) - math.pi
    print(f"structure feed: out {tuple(out.shape)} {out.dtype}  max|out| = {out.abs().max():.3f}")
    torch.save({"B": B, "L": L, "batch_seed": 55, "n_lig": (5, 64), "n_rec": (16, 128), "timesteps": t, "angles": wrapped.clone()},
               os.path.join(GOLDEN, "structure_feed_cfg5.pt"))


def golden_get_loss():
    """PeptideDiff.get_loss (model.py:313-345) of the reference with its forward replaced by fixed logits: pins the loss
    reductions (masked rates, the two CrossEntropyLoss terms, elbo_loss) independently of the denoiser."""
    ref_model, _, _ = R.load_reference()
    B, L = 6, 48
    batch = O.synthetic_batch(B, L, (5, 48), (16, 48), 61)
    g = torch.Generator().manual_seed(62)
    logits = torch.randn(B, L, 20, generator=g) * 2.5
    x0_idx = batch["ligand_seq"].argmax(-1)
    flip = (torch.rand(B, L, generator=g) < 0.4) & batch["ligand_attn_mask"].bool()
    xt_idx = torch.where(flip, torch.randint(0, 20, (B, L), generator=g), x0_idx)
    x_t = F.one_hot(xt_idx, 20).float()

    class _Shim:
        loss_function = torch.nn.CrossEntropyLoss()

        def forward(self, *a, **k):
            return logits

    out = ref_model.PeptideDiff.get_loss(_Shim(), batch, torch.zeros(B, 1), x_t)
    out_o = O.get_loss(logits, batch, x_t)
    for a, b in zip(out, out_o):
        assert torch.equal(a, b), (a, b)
    print("get_loss oracle == reference:", [round(float(v), 6) for v in out])
    torch.save({"B": B, "L": L, "batch_seed": 61, "logits": logits, "x_t_idx": xt_idx.to(torch.uint8), "out": [v.clone() for v in out]},
               os.path.join(GOLDEN, "get_loss.pt"))


STRUCT_CASES = [
    # name, L, layers, B, n_lig, n_rec, weight_seed, input_seed
    ("L64_l2", 64, 2, 3, (5, 40), (16, 64), 31, 32),
    ("L128_l12", 128, 12, 2, (5, 64), (16, 128), 33, 34),
]


def _struct_reference_model(SM, L, layers, rel):
    from transformers.models.bert.modeling_bert import BertConfig
    common = dict(max_position_embeddings=L, num_attention_heads=12, hidden_size=768, intermediate_size=1024, num_hidden_layers=layers,
                  position_embedding_type="relative_key", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, use_cache=False)
    enc = BertConfig(**common)
    dec = BertConfig(**common, is_decoder=True, add_cross_attention=True)
    for c in (enc, dec):
        try:
            c._attn_implementation = "eager"
        except Exception:
            pass
    m = SM.ConditionalBertForDiffusionBase(enc, dec, 8)
    if rel:
        R.patch_relative_key_struct(m, L)
    return m.eval()


def golden_structure_model():
    """SURVEY.md section 8(f) row 3.  (i) structure_model forward (model.py:180-215): reference module vs oracle on the same
    weights, with and without the restored relative_key term; (ii) the reference's own p_sample_loop (sample.py:104-144) for
    T = 4 under torch.manual_seed, replayed by the oracle with the same N(0,1) draws; (iii) schedule tables."""
    from oracle import structdiff_oracle as S
    SM, SS, SU = R.load_structure_sample_reference()
    for rel in (False, True):
        for name, L, layers, B, nl, nr, wseed, iseed in STRUCT_CASES:
            cfg = S.OracleConfig(max_position_embeddings=L, num_hidden_layers=layers, feature_size=8, relative_key=rel)
            sd = S.init_struct_state_dict(cfg, wseed)
            m = _struct_reference_model(SM, L, layers, rel)
            m.load_state_dict(sd, strict=True)
            batch = O.synthetic_batch(B, L, nl, nr, iseed)
            g = torch.Generator().manual_seed(iseed + 100)
            noised = (torch.rand(B, L, 8, generator=g) * 2 - 1) * math.pi * batch["ligand_attn_mask"][..., None]
            t = torch.randint(0, 1000, (B,), generator=g)
            with torch.no_grad():
                y = m(t, noised, batch["ligand_attn_mask"], batch["receptor_seq"], batch["receptor_angles"], batch["receptor_attn_mask"])
                y_o = S.struct_forward(sd, cfg, t, noised, batch["ligand_attn_mask"], batch["receptor_seq"], batch["receptor_angles"],
                                       batch["receptor_attn_mask"])
            err = (y - y_o).abs().max().item()
            print(f"struct forward {'rel' if rel else 'norel'} {name}: max|ref-oracle| = {err:.3e}  max|ref| = {y.abs().max():.3f}")
            assert err < 2e-5, err
            torch.save({"name": name, "L": L, "layers": layers, "B": B, "n_lig": nl, "n_rec": nr, "weight_seed": wseed, "input_seed": iseed,
                        "relative_key": rel, "timestep": t, "noised": noised, "out": y.clone()},
                       os.path.join(GOLDEN, f"struct_forward_{'rel' if rel else 'norel'}_{name}.pt"))
    # schedule tables
    for T in (50, 1000):
        b_ref = SU.cosine_beta_schedule(T)
        ab = SU.compute_alphas(b_ref)
        b_o = S.cosine_beta_schedule(T)
        coef = S.step_coefficients(b_o)
        assert torch.equal(b_ref, b_o)
        assert torch.equal(coef[:, 0], 1.0 / torch.sqrt(ab["alphas"])) and torch.equal(coef[:, 2], ab["sqrt_one_minus_alphas_cumprod"])
        assert torch.equal(coef[:, 3], torch.sqrt(ab["posterior_variance"]))
        torch.save({"T": T, "betas": b_ref.clone(), "coef": coef.clone()}, os.path.join(GOLDEN, f"struct_schedule_T{T}.pt"))
    v = torch.linspace(-12, 12, 4001)
    assert torch.equal(SU.modulo_with_wrapped_range(v, -math.pi, math.pi), S.modulo_with_wrapped_range(v, -math.pi, math.pi))
    # the reference p_sample_loop end to end
    T, L, layers, B = 4, 64, 2, 2
    cfg = S.OracleConfig(max_position_embeddings=L, num_hidden_layers=layers, feature_size=8, relative_key=True)
    sd = S.init_struct_state_dict(cfg, 35)
    m = _struct_reference_model(SM, L, layers, True)
    m.load_state_dict(sd, strict=True)
    batch = O.synthetic_batch(B, L, (5, 40), (16, 64), 36)
    g = torch.Generator().manual_seed(37)
    x_T = S.modulo_with_wrapped_range(torch.randn(B, L, 8, generator=g) * batch["ligand_attn_mask"][..., None])
    betas = SU.cosine_beta_schedule(T)
    torch.manual_seed(38)
    ref = SS.p_sample_loop(model=m, ligand_mask=batch["ligand_attn_mask"], ligand_angle_noise=x_T, receptor_seq=batch["receptor_seq"],
                           receptor_mask=batch["receptor_attn_mask"], receptor_angle=batch["receptor_angles"], total_timesteps=T,
                           betas=betas, disable_pbar=True)
    torch.manual_seed(38)
    noises = {}

    def noise_fn(i):
        noises[i] = torch.randn(B, L, 8)
        return noises[i]

    with torch.no_grad():
        got = S.p_sample_loop(sd, cfg, batch["ligand_attn_mask"], x_T, batch["receptor_seq"], batch["receptor_attn_mask"],
                              batch["receptor_angles"], T, betas, noise_fn)
    err = (ref - got).abs().max().item()
    print(f"struct p_sample_loop T={T}: max|ref-oracle| = {err:.3e}")
    assert err < 1e-5, err
    noise = torch.zeros(T, B, L, 8)
    for i, n in noises.items():
        noise[i] = n
    torch.save({"T": T, "L": L, "layers": layers, "B": B, "weight_seed": 35, "batch_seed": 36, "n_lig": (5, 40), "n_rec": (16, 64),
                "x_T": x_T, "noise": noise, "steps": ref.clone()}, os.path.join(GOLDEN, "struct_p_sample_loop_T4.pt"))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(8)
    export_blosum()
    golden_schedules()
    golden_reverse_step()
    golden_apply_aa_noise()
    golden_dataset()
    golden_forward()
    golden_denoise()
    golden_structure_feed()
    golden_structure_model()
    golden_get_loss()
    print("golden fixtures written to", os.path.normpath(GOLDEN))


if __name__ == "__main__":
    sys.exit(main())
