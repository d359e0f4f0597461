"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Imports the *unmodified* reference (`/root/reference/sequence_model/{model,sample,utils}.py`)
in place, in THIS container only (the GPU box has no /root/reference), so that

  * `oracle/seqdiff_oracle.py` (the CPU restatement that travels) can be pinned against it, and
  * `oracle/make_golden.py` can emit golden vectors into `tests/golden/`.

What has to be stubbed and why:
  - `pytorch_lightning`          : absent; model.py:1,256 only needs `LightningModule` as a base
                                   class and `utilities.rank_zero_info`.
  - `torch_geometric.loader`     : absent; sample.py:1 only aliases `DataLoader`.
  - cwd                          : `BlosumTransition` opens './blosum_substitute.pt' (utils.py:274-279).
  - `sample.DEVICE`              : module global `cuda:4` read at call time (sample.py:20,116,158,178).
  - transformers 5.5 (installed) silently drops the 4.38.2 `relative_key` term the authors trained
    with (environment.yml:229).  `patch_relative_key()` re-injects it, restated from the 4.38.2
    semantics in SURVEY.md Appendix A, into the reference's own module tree.
"""
from __future__ import annotations

import contextlib
import math
import os
import sys
import types

import torch
from torch import nn

REFERENCE_ROOT = "/root/reference"
SEQ_DIR = os.path.join(REFERENCE_ROOT, "sequence_model")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(SEQ_DIR, "model.py"))


def _install_stubs() -> None:
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(nn.Module):
            def log_dict(self, *a, **k):
                return None

            def log(self, *a, **k):
                return None

        util = types.ModuleType("pytorch_lightning.utilities")
        util.rank_zero_info = lambda *a, **k: None
        pl.LightningModule = LightningModule
        pl.utilities = util
        sys.modules["pytorch_lightning"] = pl
        sys.modules["pytorch_lightning.utilities"] = util
    if "torch_geometric" not in sys.modules:
        tg = types.ModuleType("torch_geometric")
        tgl = types.ModuleType("torch_geometric.loader")
        from torch.utils.data import DataLoader

        tgl.DataLoader = DataLoader
        tg.loader = tgl
        sys.modules["torch_geometric"] = tg
        sys.modules["torch_geometric.loader"] = tgl


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


_CACHE = {}


def load_reference():
    """Returns (model_module, sample_module, utils_module) of the reference, imported in place."""
    if "mods" in _CACHE:
        return _CACHE["mods"]
    if not reference_available():
        raise RuntimeError("reference tree not present (expected only in the build container)")
    _install_stubs()
    # the reference modules are called `model`, `sample`, `utils`, `dataset`: import them under
    # their own names from their own directory, then hide them again so they cannot shadow anything.
    saved = {k: sys.modules.pop(k) for k in ("model", "sample", "utils", "dataset") if k in sys.modules}
    sys.path.insert(0, SEQ_DIR)
    try:
        with _cwd(SEQ_DIR):
            import model as ref_model  # noqa
            import utils as ref_utils  # noqa
            import sample as ref_sample  # noqa
    finally:
        sys.path.remove(SEQ_DIR)
        for k in ("model", "sample", "utils", "dataset"):
            m = sys.modules.pop(k, None)
            if m is not None:
                sys.modules["_ref_seq_" + k] = m
        sys.modules.update(saved)
    ref_sample.DEVICE = torch.device("cpu")
    _CACHE["mods"] = (ref_model, ref_sample, ref_utils)
    return _CACHE["mods"]


STRUCT_DIR = os.path.join(REFERENCE_ROOT, "structure_model")


def load_structure_reference():
    """The reference's structure_model/model.py imported in place under a private name (its `model` / `utils` module names
    collide with the sequence model's).  BASELINE configs[4]: only its output SHAPE / dtype feeds the sequence model."""
    if "struct" in _CACHE:
        return _CACHE["struct"]
    if not os.path.isfile(os.path.join(STRUCT_DIR, "model.py")):
        raise RuntimeError("reference tree not present (expected only in the build container)")
    _install_stubs()
    saved = {k: sys.modules.pop(k) for k in ("model", "utils") if k in sys.modules}
    sys.path.insert(0, STRUCT_DIR)
    try:
        with _cwd(STRUCT_DIR):
            import model as struct_model  # noqa
    finally:
        sys.path.remove(STRUCT_DIR)
        for k in ("model", "utils"):
            m = sys.modules.pop(k, None)
            if m is not None:
                sys.modules["_ref_struct_" + k] = m
        sys.modules.update(saved)
    _CACHE["struct"] = struct_model
    return struct_model


def load_structure_sample_reference():
    """structure_model/{sample,utils}.py imported in place (p_sample, p_sample_loop, compute_alphas, ...).  sample.py selects
    `cuda:3` at import (sample.py:45-53): `torch.cuda.set_device` is a no-op for the duration of the import and the module's
    DEVICE global is pointed at the CPU afterwards."""
    if "struct_sample" in _CACHE:
        return _CACHE["struct_sample"]
    if not os.path.isfile(os.path.join(STRUCT_DIR, "sample.py")):
        raise RuntimeError("reference tree not present (expected only in the build container)")
    _install_stubs()
    names = ("model", "sample", "utils", "dataset")
    saved = {k: sys.modules.pop(k) for k in names if k in sys.modules}
    sys.path.insert(0, STRUCT_DIR)
    real_set_device = torch.cuda.set_device
    torch.cuda.set_device = lambda *a, **k: None
    try:
        with _cwd(STRUCT_DIR), contextlib.redirect_stdout(None):
            import sample as struct_sample  # noqa
            import utils as struct_utils  # noqa
            import model as struct_model  # noqa
    finally:
        torch.cuda.set_device = real_set_device
        sys.path.remove(STRUCT_DIR)
        for k in names:
            m = sys.modules.pop(k, None)
            if m is not None:
                sys.modules["_ref_struct_s_" + k] = m
        sys.modules.update(saved)
    struct_sample.DEVICE = "cpu"
    _CACHE["struct_sample"] = (struct_model, struct_sample, struct_utils)
    return _CACHE["struct_sample"]


def patch_relative_key_struct(model, max_pos: int):
    """4.38.2 relative_key self-attention restored in the structure model's tree: both SELayers, every encoder and decoder
    self-attention (NOT the decoder cross-attentions: 4.38.2 builds those with position_embedding_type="absolute")."""
    hidden = model.decoder_config.hidden_size
    heads = model.decoder_config.num_attention_heads
    for blk in (model.receptor_emb, model.timestep_emb):
        blk.attn.self = _RelKeySelfAttention(blk.attn.self, hidden, heads, max_pos)
    for stack in (model.encoder, model.decoder):
        for layer in stack.layer:
            layer.attention.self = _RelKeySelfAttention(layer.attention.self, hidden, heads, max_pos)
    return model


def make_blosum_transition(timestep: int = 500):
    _, _, ref_utils = load_reference()
    with _cwd(SEQ_DIR):
        return ref_utils.BlosumTransition(x_classes=20, timestep=timestep)


class _RelKeySelfAttention(nn.Module):
    """transformers 4.38.2 `BertSelfAttention(position_embedding_type="relative_key")`, restated
    (SURVEY.md Appendix A).  Returns a tuple like the HF module so `BertAttention` keeps working."""

    def __init__(self, old: nn.Module, hidden: int, heads: int, max_pos: int):
        super().__init__()
        self.query, self.key, self.value = old.query, old.key, old.value
        self.heads = heads
        self.dh = hidden // heads
        self.max_pos = max_pos
        self.distance_embedding = nn.Embedding(2 * max_pos - 1, self.dh)  # default init N(0,1)

    def _split(self, x):
        b, l, _ = x.shape
        return x.view(b, l, self.heads, self.dh).permute(0, 2, 1, 3)

    def forward(self, hidden_states, attention_mask=None, *args, **kwargs):
        q = self._split(self.query(hidden_states))
        k = self._split(self.key(hidden_states))
        v = self._split(self.value(hidden_states))
        s = q @ k.transpose(-1, -2)
        L = hidden_states.shape[1]
        pos_l = torch.arange(L).view(-1, 1)
        pos_r = torch.arange(L).view(1, -1)
        dist = pos_l - pos_r
        pe = self.distance_embedding(dist + self.max_pos - 1).to(q.dtype)  # [L, L, dh]
        s = s + torch.einsum("bhld,lrd->bhlr", q, pe)
        s = s / math.sqrt(self.dh)
        if attention_mask is not None:
            s = s + attention_mask
        p = torch.softmax(s, dim=-1)
        ctx = (p @ v).permute(0, 2, 1, 3).contiguous()
        ctx = ctx.view(ctx.shape[0], L, self.heads * self.dh)
        return (ctx, None)


def patch_relative_key(model, max_pos: int):
    """Injects the 4.38.2 relative_key self-attention into the reference module tree:
    SELayer.attn.self x3 and decoder.layer[i].attention.self x6 (NOT crossattention)."""
    hidden = model.decoder_config.hidden_size
    heads = model.decoder_config.num_attention_heads
    for blk in (model.ligand_feature_emb, model.receptor_feature_emb, model.decoder_normalize):
        blk.attn.self = _RelKeySelfAttention(blk.attn.self, hidden, heads, max_pos)
    for layer in model.decoder.layer:
        layer.attention.self = _RelKeySelfAttention(layer.attention.self, hidden, heads, max_pos)
    return model


def make_configs(max_seq_len: int, hidden=768, heads=12, inter=1024, layers=6):
    """The two BertConfig objects of sample.py:69-92."""
    from transformers import BertConfig

    common = dict(
        max_position_embeddings=max_seq_len,
        num_attention_heads=heads,
        hidden_size=hidden,
        intermediate_size=inter,
        num_hidden_layers=layers,
        position_embedding_type="relative_key",
        hidden_dropout_prob=0.1,
        attention_probs_dropout_prob=0.1,
        use_cache=False,
    )
    enc = BertConfig(**common)
    dec = BertConfig(**common, is_decoder=True, add_cross_attention=True)
    # transformers 5.x picks sdpa by default; the eager path is the 4.38.2-equivalent arithmetic.
    for c in (enc, dec):
        try:
            c._attn_implementation = "eager"
        except Exception:
            pass
    return enc, dec


def build_reference_model(max_seq_len: int, relative_key: bool = True, seed: int = 0, **kw):
    """`ConditionalBertForDiffusionBase` of the reference (model.py:156-253), eval mode, CPU."""
    ref_model, _, _ = load_reference()
    enc, dec = make_configs(max_seq_len, **kw)
    torch.manual_seed(seed)
    m = ref_model.ConditionalBertForDiffusionBase(enc, dec, 20)
    if relative_key:
        patch_relative_key(m, max_seq_len)
    return m.eval()
