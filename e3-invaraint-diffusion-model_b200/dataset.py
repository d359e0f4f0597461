"""Drop-in for `sequence_model/dataset.py` of the reference: same class name, constructor and item/batch keys, with
the per-item work of `__getitem__` (dataset.py:97-129: pocket-mask dilation, boolean-mask gathers, zero padding,
attention masks) done for a whole batch by one CUDA kernel (`seqdiff_collate`).  The reference does this on CPU
DataLoader workers, one complex at a time."""
from __future__ import annotations

import ctypes
import random
from typing import Dict, List, Sequence

import torch

from . import _cabi

RANDOM_SEED = 0
AA_VOCAB = "ACDEFGHIKLMNPQRSTVWY"
SS_VOCAB = "HBEGITS-"
_LUTS: Dict[str, Dict[str, int]] = {}


def collate_complexes(records: Sequence[Dict], max_len: int, pocket_ext: int, device) -> Dict:
    """records: dicts with `amino_acid` [n,20] one-hot f32, `angle_features` [n,8] f32, `ligand_mask`, `pocket_mask` [n] bool,
    `structure_ids` -- the schema written by clean_data/data_preprocessing.py:882-893 after `_load_file`.
    Returns the batch dict a DataLoader over the reference dataset yields (keys of dataset.py:115-129), on `device`."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("collate_complexes runs only on a CUDA device (no CPU fallback)")
    G = len(records)
    counts = [int(r["ligand_mask"].shape[0]) for r in records]
    offsets = torch.tensor([0] + list(torch.tensor(counts).cumsum(0).tolist()), dtype=torch.int32)
    cat = lambda k, dt: torch.cat([r[k].to(dt) for r in records], 0).contiguous()
    lig_m, poc_m = cat("ligand_mask", torch.uint8).to(dev), cat("pocket_mask", torch.uint8).to(dev)
    ang, aa = cat("angle_features", torch.float32).to(dev), cat("amino_acid", torch.float32).to(dev)
    off = offsets.to(dev)
    new = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
    out = {"ligand_angles": new(G, max_len, 8), "ligand_seq": new(G, max_len, 20), "ligand_attn_mask": new(G, max_len),
           "receptor_angles": new(G, max_len, 8), "receptor_seq": new(G, max_len, 20), "receptor_attn_mask": new(G, max_len)}
    lengths = torch.empty(G, 2, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        p = _cabi.ptr
        _cabi.check(_cabi.lib().seqdiff_collate(G, p(off), p(lig_m), p(poc_m), p(ang), p(aa), int(pocket_ext), int(max_len),
                                                p(out["ligand_angles"]), p(out["ligand_seq"]), p(out["ligand_attn_mask"]),
                                                p(out["receptor_angles"]), p(out["receptor_seq"]), p(out["receptor_attn_mask"]), p(lengths), stream))
    lengths_h = lengths.cpu()
    if int(lengths_h.max()) > max_len:
        raise RuntimeError("Length exceed")  # dataset.py:42-43
    out["ligand_length"], out["receptor_length"] = lengths_h[:, 0].clone(), lengths_h[:, 1].clone()
    out["ligand_pos_id"] = torch.zeros(G, dtype=torch.long)
    out["receptor_pos_id"] = torch.zeros(G, dtype=torch.long)
    ids = [r["structure_ids"] for r in records]
    out["structure_ids"] = {k: [d[k] for d in ids] for k in ids[0]} if ids and isinstance(ids[0], dict) else ids
    return out


class LigandBindingSiteDataset(torch.utils.data.Dataset):
    """reference dataset.py:12-129.  `__getitem__` keeps the reference contract (one padded item, CPU tensors are moved
    to `device`); `collate_batch(indices)` is the GPU path that builds a whole batch with one kernel."""
    feature_names = list(AA_VOCAB)

    def __init__(self, filepath: str, split: str, max_len: int = 64, pocket_ext: int = 1, device="cuda:0") -> None:
        super().__init__()
        self._load_file(filepath)
        self._split_data(split)
        self.max_len = max_len
        self.pocket_ext = pocket_ext
        self.device = device

    @staticmethod
    def _one_hot_encode(sequence, vocab):
        """letters -> [n, len(vocab)] one-hot f32 rows (a letter outside the vocabulary is a ValueError, as `str.index` raises)."""
        lut = _LUTS.setdefault(vocab, {ch: i for i, ch in enumerate(vocab)})
        try:
            idx = torch.tensor([lut[ch] for ch in sequence], dtype=torch.long)
        except KeyError as e:
            raise ValueError(f"letter {e.args[0]!r} is not in the vocabulary {vocab!r}") from None
        return torch.eye(len(vocab), dtype=torch.float32)[idx]

    def _split_data(self, split_name):
        """Deterministic 80 / 10 / 10 split (quirk Q12): the records are put in the order `random.seed(0); random.shuffle(data)`
        produces -- drawn here from a private generator as an index permutation, so the process-wide `random` state is left alone --
        and the named slice is kept (None or an unknown name keeps everything, like the reference)."""
        order = list(range(len(self.data)))
        random.Random(RANDOM_SEED).shuffle(order)
        n_train, n_val = int(len(order) * 0.8), int(len(order) * 0.1)
        bounds = {"train": (0, n_train), "validation": (n_train, n_train + n_val), "test": (n_train + n_val, len(order))}
        lo, hi = bounds.get(split_name, (0, len(order)))
        self.data = [self.data[i] for i in order[lo:hi]]

    def _load_file(self, filepath: str) -> None:
        """`biolip.pt` = list of per-complex dicts (clean_data/data_preprocessing.py:882-893); the two letter sequences are
        replaced by their one-hot encodings once, at load time."""
        print(f"Loading data from {filepath}")
        self.data = torch.load(filepath)
        for rec in self.data:
            for key, vocab in (("amino_acid", AA_VOCAB), ("secondary_structure", SS_VOCAB)):
                rec[key] = self._one_hot_encode("".join(rec[key]), vocab)

    def __len__(self) -> int:
        return len(self.data)

    def get_structure_id(self, index):
        return self.data[index]["structure_ids"]

    def collate_batch(self, indices: List[int]) -> Dict:
        for index in indices:
            if not 0 <= index < len(self):
                raise IndexError("Index out of range")
        return collate_complexes([self.data[i] for i in indices], self.max_len, self.pocket_ext, self.device)

    def __getitem__(self, index):
        b = self.collate_batch([index])
        item = {k: (v[0] if torch.is_tensor(v) else v) for k, v in b.items() if k != "structure_ids"}
        item["ligand_pos_id"] = 0
        item["receptor_pos_id"] = 0
        item["structure_ids"] = self.data[index]["structure_ids"]
        return item
