#!/usr/bin/env python
"""Per-tensor-class attribution of the bf16 logit error (CPU, oracle arithmetic + emulated roundings).

The CUDA path keeps accumulators, the residual stream, LayerNorm, softmax and logits in fp32 and rounds ONLY the operands of its
tensor-core contractions to 16 bits.  This script restates the forward (same structure as oracle/seqdiff_oracle.py) with an
explicit rounding hook at every place where the CUDA path stores a 16-bit tensor, so that each class of rounding can be switched
on alone ("only") or off alone ("all-but") and its share of the final logit error measured against the exact fp32 forward.

    python scripts/bf16_attribution.py [--fmt bf16|fp16] [--variant A|B] [--L 128] [--B 2]

Output: one table per weight variant; committed as profiles/bf16_attribution_r02.txt.  Test infrastructure (imports the oracle).
"""
from __future__ import annotations

import argparse
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import seqdiff_oracle as O  # noqa: E402

CLASSES = ["w", "a_in", "c", "u", "mod", "qkv", "E", "p", "ctx", "act", "ckv", "te16", "y"]
DESCR = {
    "w": "every Linear weight (and nothing else)",
    "a_in": "16-bit copy of the fp32 residual stream fed to a GEMM as A (post-LN h, x, x1)",
    "c": "conditioning c = LN(Linear(angles)) + te (A of adaLN_modulation.0)",
    "u": "SiLU(adaLN_modulation.0(c)) (A of adaLN_modulation.2)",
    "mod": "adaLN shift/scale/gate (output of adaLN_modulation.2)",
    "qkv": "Q, K, V of self- and cross-attention queries",
    "E": "distance_embedding table",
    "p": "softmax probabilities P before P.V",
    "ctx": "attention context (A of the output dense)",
    "act": "GELU outputs (mlp.0 / intermediate.dense)",
    "ckv": "cross-attention K, V of the receptor features",
    "te16": "16-bit timestep features fed to decoder_normalize's adaLN",
    "y": "GELU(dense1) of the predictor (read by the LN + 768->20 tail)",
}


class Emu:
    def __init__(self, on, fmt, head_fmt=None):
        self.on = set(on)
        self.dt = torch.bfloat16 if fmt == "bf16" else torch.float16
        # operand format of the output head (dense1 of the predictor + its GELU output): the product runs it in fp16 in bf16 mode
        self.head_dt = self.dt if head_fmt is None else (torch.bfloat16 if head_fmt == "bf16" else torch.float16)

    def r(self, cls, x):
        return x.to(self.dt).float() if cls in self.on else x

    def lin(self, sd, prefix, x):
        return F.linear(x, self.r("w", sd[prefix + ".weight"]), sd[prefix + ".bias"])


def attention_core(e: Emu, cfg, q, k, v, add_mask, dist_emb):
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
        pe = e.r("E", dist_emb)[dist + P - 1]
        s = s + torch.einsum("bhld,lrd->bhlr", qh, pe)
    s = s / math.sqrt(dh) + add_mask
    # the kernel normalises AFTER P.V: p~ = exp(s - max) is rounded, the fp32 row sum uses the unrounded values
    mx = s.max(dim=-1, keepdim=True).values
    pu = torch.exp(s - mx)
    ctx = (e.r("p", pu) @ vh) / pu.sum(-1, keepdim=True)
    return ctx.permute(0, 2, 1, 3).reshape(B, Lq, H)


def bert_attention(e, sd, cfg, prefix, x, add_mask, kv=None, kv_mask=None, rel=True):
    xa = e.r("a_in", x)
    q = e.r("qkv", e.lin(sd, prefix + ".self.query", xa))
    if kv is None:
        k = e.r("qkv", e.lin(sd, prefix + ".self.key", xa))
        v = e.r("qkv", e.lin(sd, prefix + ".self.value", xa))
    else:
        ka = e.r("a_in", kv)
        k = e.r("ckv", e.lin(sd, prefix + ".self.key", ka))
        v = e.r("ckv", e.lin(sd, prefix + ".self.value", ka))
    de = sd.get(prefix + ".self.distance_embedding.weight") if (rel and cfg.relative_key) else None
    ctx = e.r("ctx", attention_core(e, cfg, q, k, v, add_mask if kv is None else kv_mask, de))
    out = e.lin(sd, prefix + ".output.dense", ctx)
    return O._ln(sd, prefix + ".output.LayerNorm", out + x, cfg.layer_norm_eps)


def se_layer(e, sd, cfg, prefix, x, c16, add_mask):
    u = e.r("u", F.silu(e.lin(sd, prefix + ".adaLN_modulation.0", c16)))
    mod = e.r("mod", e.lin(sd, prefix + ".adaLN_modulation.2", u))
    sh1, sc1, g1, sh2, sc2, g2 = mod.chunk(6, dim=-1)
    H = x.shape[-1]
    a = bert_attention(e, sd, cfg, prefix + ".attn", x, add_mask)
    x = x + g1 * (F.layer_norm(a, (H,)) * (1 + sc1) + sh1)
    m1 = e.r("act", F.gelu(e.lin(sd, prefix + ".mlp.0", e.r("a_in", x))))
    m = e.lin(sd, prefix + ".mlp.3", m1)
    return x + g2 * (F.layer_norm(m, (H,)) * (1 + sc2) + sh2)


def bert_layer(e, sd, cfg, prefix, h, add_mask, enc, enc_mask):
    h1 = bert_attention(e, sd, cfg, prefix + ".attention", h, add_mask)
    h2 = bert_attention(e, sd, cfg, prefix + ".crossattention", h1, None, kv=enc, kv_mask=enc_mask, rel=False)
    inter = e.r("act", F.gelu(e.lin(sd, prefix + ".intermediate.dense", e.r("a_in", h2))))
    out = e.lin(sd, prefix + ".output.dense", inter)
    return O._ln(sd, prefix + ".output.LayerNorm", out + h2, cfg.layer_norm_eps)


def forward(e: Emu, sd, cfg, timestep, x_t, lig_ang, lig_mask, rec_seq, rec_ang, rec_mask):
    eps = cfg.layer_norm_eps
    lm, rm = O.extend_mask(lig_mask), O.extend_mask(rec_mask)
    te = O.timestep_embedding(sd, timestep.squeeze(dim=-1)).unsqueeze(1)
    # embeddings are fp32 SIMT in the kernel (weights fp32)
    h_seq = O.bert_embeddings(sd, "ligand_seq_embedding", x_t, eps)
    c_lig = e.r("c", O.bert_embeddings(sd, "ligand_angle_embedding", lig_ang, eps) + te)
    lig = se_layer(e, sd, cfg, "ligand_feature_emb", h_seq, c_lig, lm)
    r_seq = O.bert_embeddings(sd, "receptor_seq_embedding", rec_seq, eps)
    c_rec = e.r("c", O.bert_embeddings(sd, "receptor_angle_embedding", rec_ang, eps) + te)
    rec = se_layer(e, sd, cfg, "ligand_feature_emb", r_seq, c_rec, rm)
    h = lig
    for i in range(cfg.num_hidden_layers):
        h = bert_layer(e, sd, cfg, f"decoder.layer.{i}", h, lm, rec, rm)
    h = se_layer(e, sd, cfg, "decoder_normalize", h, e.r("te16", te), lm)
    p = "amino_acid_predictor"
    hd = (lambda cls, x: x.to(e.head_dt).float() if cls in e.on else x)
    y = hd("y", F.gelu(F.linear(hd("a_in", h), hd("w", sd[p + ".dense1.weight"]), sd[p + ".dense1.bias"])))
    y = O._ln(sd, p + ".layer_norm", y, 1e-12)
    return F.linear(y, sd[p + ".dense2.weight"], sd[p + ".dense2.bias"])  # the tail keeps W2 in fp32


def errs(a, b):
    return ((a - b).abs().max() / b.abs().max()).item(), ((a - b).norm() / b.norm()).item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fmt", default="bf16")
    ap.add_argument("--L", type=int, default=128)
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--variants", default="A,B")
    ap.add_argument("--seeds", default="0,1")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    for variant in args.variants.split(","):
        rows = {}
        for seed in [int(s) for s in args.seeds.split(",")]:
            cfg = O.OracleConfig(max_position_embeddings=args.L)
            sd = O.init_state_dict(cfg, seed, variant)
            batch = O.synthetic_batch(args.B, args.L, (20, 64), (60, 128), 10 + seed)
            x_t = O.generate_discrete_noise(args.B, args.L, generator=torch.Generator().manual_seed(20 + seed))
            t = torch.full((args.B, 1), 17.0)
            a = (t, x_t, batch["ligand_angles"], batch["ligand_attn_mask"], batch["receptor_seq"], batch["receptor_angles"], batch["receptor_attn_mask"])
            with torch.no_grad():
                want = O.denoiser_forward(sd, cfg, *a)
                exact = forward(Emu([], args.fmt), sd, cfg, *a)
                assert errs(exact, want)[0] < 1e-5, errs(exact, want)
                rows.setdefault("ALL", []).append(errs(forward(Emu(CLASSES, args.fmt), sd, cfg, *a), want))
                rows.setdefault("ALL, fp16 head", []).append(errs(forward(Emu(CLASSES, args.fmt, "fp16"), sd, cfg, *a), want))
                rows.setdefault("ALL but w", []).append(errs(forward(Emu([c for c in CLASSES if c != "w"], args.fmt), sd, cfg, *a), want))
                for c in CLASSES:
                    rows.setdefault("only " + c, []).append(errs(forward(Emu([c], args.fmt), sd, cfg, *a), want))
        print(f"\n== {args.fmt}, weights variant {variant}, B={args.B}, L={args.L}, seeds {args.seeds}: error of the logits vs exact fp32 (mean over seeds)")
        print(f"{'roundings switched on':<14} {'max-norm rel':>12} {'L2 rel':>10}   share of ALL variance (L2^2)")
        tot = sum(x[1] for x in rows["ALL"]) / len(rows["ALL"])
        for k, v in rows.items():
            mx = sum(x[0] for x in v) / len(v)
            l2 = sum(x[1] for x in v) / len(v)
            share = (l2 / tot) ** 2
            d = DESCR.get(k.replace("only ", ""), "")
            print(f"{k:<14} {mx:12.3e} {l2:10.3e}   {share:6.1%}  {d}")


if __name__ == "__main__":
    main()
