"""Size-independent properties of the host logic and the oracles (CPU, hypothesis): what the GPU parity tests rely on."""
import math

import torch
from hypothesis import given, settings, strategies as st

from helpers import O
from oracle import structdiff_oracle as S


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 2000), st.integers(1, 16))
def test_shard_bounds_is_a_partition(n, world):
    import seqdiff_b200 as sd
    spans = [sd.shard_bounds(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


@settings(max_examples=40, deadline=None)
@given(st.lists(st.floats(-50, 50, allow_nan=False, width=32), min_size=1, max_size=64))
def test_wrap_is_idempotent_periodic_and_in_range(vals):
    """structure_model/utils.py:20-40: the wrap maps into [-pi, pi], fixes its own outputs, and does not care about whole turns."""
    v = torch.tensor(vals, dtype=torch.float32)
    w = S.modulo_with_wrapped_range(v)
    assert (w >= -math.pi - 1e-6).all() and (w <= math.pi + 1e-6).all()
    assert torch.allclose(S.modulo_with_wrapped_range(w), w, atol=1e-6)
    d = (S.modulo_with_wrapped_range(v + 2 * math.pi) - w).abs()
    assert (torch.minimum(d, (2 * math.pi - d).abs()) < 5e-5).all()
    assert torch.allclose(torch.cos(w), torch.cos(v), atol=2e-5) and torch.allclose(torch.sin(w), torch.sin(v), atol=2e-5)


@settings(max_examples=10, deadline=None)
@given(st.integers(2, 1000))
def test_gaussian_schedule_tables_are_sane(T):
    """structure_model/utils.py:8-58: betas clipped to [1e-4, 0.9999]; alphabar decreasing; step 0 adds no variance."""
    betas = S.cosine_beta_schedule(T)
    coef = S.step_coefficients(betas)
    assert betas.shape == (T,) and (betas >= 1e-4 - 1e-9).all() and (betas <= 0.9999 + 1e-7).all()
    ab = S.compute_alphas(betas)
    assert (ab["alphas_cumprod"][1:] <= ab["alphas_cumprod"][:-1] + 1e-7).all()
    assert coef.shape == (T, 4) and torch.isfinite(coef).all()
    assert coef[0, 3].item() == 0.0           # posterior_variance[0] = beta_0 * (1 - 1) / (1 - alphabar_0)
    assert (coef[:, 0] >= 1.0 - 1e-6).all()   # 1 / sqrt(alpha) >= 1


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 5), st.integers(1, 40), st.integers(0, 2 ** 31 - 1))
def test_decode_oracle_counts(B, L, seed):
    """sample.py:208-224: recovery rate = matches over the masked positions; strings have the masked length."""
    g = torch.Generator().manual_seed(seed)
    final = torch.randn(B, L, 20, generator=g)
    true = torch.nn.functional.one_hot(torch.randint(0, 20, (B, L), generator=g), 20).float()
    n = torch.randint(1, L + 1, (B,), generator=g)
    mask = (torch.arange(L)[None, :] < n[:, None]).float()
    trues, preds, rates = O.decode(final, {"ligand_seq": true, "ligand_attn_mask": mask})
    for i in range(B):
        assert len(preds[i]) == len(trues[i]) == int(n[i])
        hits = sum(a == b for a, b in zip(preds[i], trues[i]))
        assert abs(rates[i] - hits / int(n[i])) < 1e-6


@settings(max_examples=15, deadline=None)
@given(st.integers(0, 2 ** 31 - 1))
def test_loss_terms_identities(seed):
    """model.py:313-345: total = CE(noised) + elbo; rates are fractions of the masked positions; elbo >= entropy term >= 0."""
    g = torch.Generator().manual_seed(seed)
    B, L = 3, 24
    batch = O.synthetic_batch(B, L, (4, 24), (8, 24), seed % 1000)
    logits = torch.randn(B, L, 20, generator=g) * 2
    x0 = batch["ligand_seq"].argmax(-1)
    flip = (torch.rand(B, L, generator=g) < 0.5) & batch["ligand_attn_mask"].bool()
    xt = torch.where(flip, (x0 + torch.randint(1, 20, (B, L), generator=g)) % 20, x0)
    x_t = torch.nn.functional.one_hot(xt, 20).float()
    if not flip.any() or not (batch["ligand_attn_mask"].bool() & ~flip).any():
        return
    total, elbo, ce_n, ce_all, rec, nrate = O.get_loss(logits, batch, x_t)
    assert torch.allclose(total, ce_n + elbo)
    assert 0 <= rec.item() <= 1 and 0 <= nrate.item() <= 1
    n_mask = batch["ligand_attn_mask"].sum().item()
    assert abs(nrate.item() - (1 - flip.sum().item() / n_mask)) < 1e-6
    assert ce_n.item() > 0 and ce_all.item() > 0 and elbo.item() > 0


def test_padding_never_reaches_valid_rows():
    """The exactness argument behind packing ragged batches: with the reference's -10000 key mask a padded key's
    softmax weight underflows to exactly 0 and every other operator is row-local, so the logits of the VALID ligand rows of a padded
    batch equal those of the same complex run alone at its true lengths (relative positions are within-graph, so nothing shifts)."""
    L = 48
    cfg = O.OracleConfig(max_position_embeddings=L, num_hidden_layers=2)
    state = O.init_state_dict(cfg, 5, "B")
    batch = O.synthetic_batch(3, L, (5, 30), (9, 48), 91)
    x_t = O.generate_discrete_noise(3, L, generator=torch.Generator().manual_seed(92))
    t = torch.full((3, 1), 13.0)
    torch.set_num_threads(4)
    with torch.no_grad():
        padded = O.denoiser_forward(state, cfg, t, x_t, batch["ligand_angles"], batch["ligand_attn_mask"], batch["receptor_seq"],
                                    batch["receptor_angles"], batch["receptor_attn_mask"])
        for b in range(3):
            nl, nr = int(batch["ligand_attn_mask"][b].sum()), int(batch["receptor_attn_mask"][b].sum())
            alone = O.denoiser_forward(state, cfg, t[b:b + 1], x_t[b:b + 1, :nl], batch["ligand_angles"][b:b + 1, :nl], torch.ones(1, nl),
                                       batch["receptor_seq"][b:b + 1, :nr], batch["receptor_angles"][b:b + 1, :nr], torch.ones(1, nr))
            assert alone.shape == (1, nl, 20)
            assert (alone[0] - padded[b, :nl]).abs().max().item() < 2e-5 * padded[b, :nl].abs().max().item()
            # and garbage in the padded ligand rows of x_t changes nothing on the valid rows
            x_bad = x_t.clone()
            x_bad[b, nl:] = torch.roll(x_t[b, nl:], 7, dims=-1)
            again = O.denoiser_forward(state, cfg, t, x_bad, batch["ligand_angles"], batch["ligand_attn_mask"], batch["receptor_seq"],
                                       batch["receptor_angles"], batch["receptor_attn_mask"])
            assert torch.equal(again[b, :nl], padded[b, :nl])
