"""Output decode (sample.py:208-224) and loss reductions (model.py:313-345 + utils.py:132-161,
row a17): oracle against the reference golden on CPU, CUDA kernels against the oracle on the GPU."""
import os

import pytest
import torch
import torch.nn.functional as F

from helpers import GOLDEN, O

gpu = pytest.mark.gpu


def _loss_case():
    g = torch.load(os.path.join(GOLDEN, "get_loss.pt"))
    batch = O.synthetic_batch(g["B"], g["L"], (5, 48), (16, 48), g["batch_seed"])
    x_t = F.one_hot(g["x_t_idx"].long(), 20).float()
    return g, batch, x_t


def test_get_loss_oracle_matches_reference_golden():
    g, batch, x_t = _loss_case()
    got = O.get_loss(g["logits"], batch, x_t)
    for a, b in zip(got, g["out"]):
        assert torch.allclose(a, b, rtol=1e-6, atol=0), (a, b)


def test_decode_oracle_on_denoise_golden():
    """The reference's denoise() output strings (tests/golden/denoise_T4.pt) are reproduced by the decode restatement."""
    g = torch.load(os.path.join(GOLDEN, "denoise_T4.pt"))
    batch = O.synthetic_batch(g["B"], g["L"], tuple(g["n_lig"]), tuple(g["n_rec"]), g["batch_seed"])
    trues, preds, rates = O.decode(g["final_logits"], batch)
    assert preds == g["pred_sequences"] and trues == g["true_sequences"]
    assert rates == pytest.approx(g["recovery_rates"], abs=0)


@gpu
def test_loss_terms_kernel_matches_oracle_and_reference_golden():
    import seqdiff_b200 as sd
    g, batch, x_t = _loss_case()
    terms = sd.model.loss_terms(g["logits"].cuda(), batch["ligand_seq"], x_t, batch["ligand_attn_mask"]).cpu()
    mask = batch["ligand_attn_mask"].bool()
    x0i, xti = batch["ligand_seq"].argmax(-1), x_t.argmax(-1)
    noised = xti != x0i
    sel = mask & ~noised
    # integer terms are exact
    assert terms[0].item() == mask.sum().item() and terms[1].item() == noised.sum().item() and terms[2].item() == sel.sum().item()
    assert terms[3].item() == (mask & (xti == x0i)).sum().item()
    assert terms[4].item() == (mask & (g["logits"].argmax(-1) == x0i)).sum().item()
    total, elbo, ce_n, ce_all, rec, nrate = g["out"]
    n_noised, n_sel, n_mask = terms[1], terms[2], terms[0]
    assert (terms[5] / n_noised).item() == pytest.approx(ce_n.item(), rel=1e-5)
    assert (terms[6] / n_sel).item() == pytest.approx(ce_all.item(), rel=1e-5)
    assert ((-terms[7] + terms[8]) / n_noised).item() == pytest.approx(elbo.item(), rel=1e-5)
    assert (terms[4] / n_mask).item() == pytest.approx(rec.item(), rel=1e-6)
    assert (terms[3] / n_mask).item() == pytest.approx(nrate.item(), rel=1e-6)
    # run-to-run deterministic (fixed reduction order), also on a size that needs the CTA cap
    N = 300_000
    gen = torch.Generator().manual_seed(3)
    lg = torch.randn(N, 20, generator=gen).cuda()
    a = F.one_hot(torch.randint(0, 20, (N,), generator=gen), 20).float().cuda()
    b = F.one_hot(torch.randint(0, 20, (N,), generator=gen), 20).float().cuda()
    m = (torch.rand(N, generator=gen) < 0.7).float().cuda()
    t1 = sd.model.loss_terms(lg, a, b, m).cpu()
    t2 = sd.model.loss_terms(lg, a, b, m).cpu()
    assert torch.equal(t1, t2)
    nm = (b.argmax(-1) != a.argmax(-1))
    want = F.cross_entropy(lg[nm], a[nm].argmax(-1), reduction="sum").double().item()
    assert t1[5].item() == pytest.approx(want, rel=1e-5)
    assert ((-t1[7] + t1[8]) / t1[1]).item() == pytest.approx(O.elbo_loss(lg[nm].cpu(), a[nm].cpu()).item(), rel=1e-5)


@gpu
def test_get_loss_through_the_model_mirror():
    """PeptideDiff.get_loss (CUDA forward + CUDA reductions) == oracle get_loss on the CUDA logits."""
    import seqdiff_b200 as sd
    L, B = 64, 3
    common = dict(max_position_embeddings=L, intermediate_size=1024, num_hidden_layers=2, position_embedding_type="relative_key")
    p = sd.PeptideDiff(sd.BertConfig(**common), sd.BertConfig(**common, is_decoder=True, add_cross_attention=True), list(sd.AA_VOCAB),
                       torch.nn.CrossEntropyLoss(), "cosine", 50)
    cfg = O.OracleConfig(max_position_embeddings=L, num_hidden_layers=2)
    missing = p.load_state_dict(O.init_state_dict(cfg, 1, "B"), strict=False)
    assert missing.unexpected_keys == [] and all(k.startswith("discrete_noise_schedule.") for k in missing.missing_keys)
    p = p.eval().cuda()
    p.precision = "fp32"
    batch = {k: v.cuda() for k, v in O.synthetic_batch(B, L, (5, 40), (16, 64), 71).items()}
    t_int = torch.tensor([[10.0], [25.0], [40.0]], device="cuda")
    x_t = p.apply_aa_noise(batch["ligand_seq"], t_int)
    out = p.get_loss(batch, t_int / 50, x_t)
    logits = p(t_int / 50, x_t, batch["ligand_angles"], batch["ligand_attn_mask"], batch["receptor_seq"], batch["receptor_angles"],
               batch["receptor_attn_mask"]).cpu()
    want = O.get_loss(logits, {k: v.cpu() for k, v in batch.items()}, x_t.cpu())
    for a, b in zip(out, want):
        assert a.dtype == torch.float32 and a.ndim == 0
        assert a.item() == pytest.approx(b.item(), rel=2e-5), (a, b)
    assert isinstance(p.validation_step(batch, 0).item(), float)
    p.release()


@gpu
def test_decode_kernel_matches_oracle():
    import seqdiff_b200 as sd
    g = torch.load(os.path.join(GOLDEN, "denoise_T4.pt"))
    batch = O.synthetic_batch(g["B"], g["L"], tuple(g["n_lig"]), tuple(g["n_rec"]), g["batch_seed"])
    pred, true, counts = sd.decode_tensors(g["final_logits"].cuda(), batch["ligand_seq"], batch["ligand_attn_mask"])
    assert torch.equal(pred.cpu().long(), g["final_logits"].argmax(-1))
    assert torch.equal(true.cpu().long(), batch["ligand_seq"].argmax(-1))
    mask = batch["ligand_attn_mask"].bool()
    for i in range(g["B"]):
        assert counts[i, 1].item() == mask[i].sum().item()
        assert counts[i, 0].item() == (pred[i].cpu()[mask[i]] == true[i].cpu()[mask[i]]).sum().item()
    rates = (counts[:, 0] / counts[:, 1]).tolist()
    assert rates == g["recovery_rates"]  # bit-identical float32 ratios
    # ties and larger / odd shapes: first maximum wins, like torch.argmax
    B, L = 37, 301
    gen = torch.Generator().manual_seed(9)
    fin = torch.randint(0, 3, (B, L, 20), generator=gen).float()  # many ties
    tru = F.one_hot(torch.randint(0, 20, (B, L), generator=gen), 20).float()
    msk = (torch.rand(B, L, generator=gen) < 0.5).float()
    pred, true, counts = sd.decode_tensors(fin.cuda(), tru, msk)
    first_max = (fin == fin.max(-1, keepdim=True).values).float().argmax(-1)
    assert torch.equal(pred.cpu().long(), first_max)
    assert torch.equal(counts[:, 1].cpu(), msk.sum(1).long())
    assert torch.equal(counts[:, 0].cpu(), ((first_max == tru.argmax(-1)) & msk.bool()).sum(1))
