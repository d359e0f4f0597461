"""Shared helpers for the parity tests."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import seqdiff_oracle as O  # noqa: E402  (the checker)

GOLDEN = os.path.join(ROOT, "tests", "golden")
FP32, BF16, FP16 = 0, 1, 2


def sd_pkg():
    import seqdiff_b200 as sd
    return sd


def stream_ptr(dev="cuda:0"):
    return ctypes.c_void_p(torch.cuda.current_stream(torch.device(dev)).cuda_stream)


def make_model(sd, cfg: "O.OracleConfig", state, precision="bf16", device="cuda:0"):
    pos = "relative_key" if cfg.relative_key else "absolute"
    common = dict(max_position_embeddings=cfg.max_position_embeddings, num_attention_heads=cfg.num_attention_heads,
                  hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size, num_hidden_layers=cfg.num_hidden_layers,
                  position_embedding_type=pos)
    enc = sd.BertConfig(**common)
    dec = sd.BertConfig(**common, is_decoder=True, add_cross_attention=True)
    m = sd.ConditionalBertForDiffusionBase(enc, dec, cfg.feature_size)
    m.load_state_dict(state, strict=True)
    m = m.eval().to(device)
    m.precision = precision
    return m


def rel_err(a, b):
    """max|a-b| / max|b| -- the 'relative' of the parity tolerances (1e-5 fp32, 1e-2 bf16)."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def l2_rel(a, b):
    """||a-b||_2 / ||b||_2"""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def top2_margin(v):
    """relative gap between the best and the second best entry of each row."""
    t = torch.topk(v, 2, dim=-1).values
    return ((t[..., 0] - t[..., 1]).abs() / t[..., 0].abs().clamp_min(1e-30))


def assert_indices_match(idx_gpu, idx_ref, score_ref, what, tol=1e-5):
    """Sampled indices must be bit-exact, except where the reference's own top-2 race scores are within
    `tol` relative (CPU libm vs CUDA expf/div differ by an ulp there)."""
    idx_gpu, idx_ref = idx_gpu.reshape(-1).cpu().long(), idx_ref.reshape(-1).cpu().long()
    bad = (idx_gpu != idx_ref).nonzero().reshape(-1)
    if bad.numel() == 0:
        return 0
    margins = top2_margin(score_ref.reshape(idx_ref.numel(), -1)[bad])
    assert (margins < tol).all(), f"{what}: {bad.numel()} index mismatches, worst margin {margins.max().item():.3e}"
    assert bad.numel() <= max(1, idx_ref.numel() // 1000), f"{what}: too many near-tie flips ({bad.numel()})"
    return int(bad.numel())
