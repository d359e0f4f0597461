"""Drop-in for `sequence_model/dataset.py` of the reference: same class name, constructor and item/batch keys, with
the per-item work of `__getitem__` (dataset.py:97-129: pocket-mask dilation, boolean-mask gathers, zero padding,
attention masks) done for a whole batch by one CUDA kernel (`seqdiff_collate`).  The reference does this on CPU
DataLoader workers, one complex at a time."""
from __future__ import annotations

import ctypes
import random
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

from . import _cabi

RANDOM_SEED = 0
AA_VOCAB = "ACDEFGHIKLMNPQRSTVWY"
SS_VOCAB = "HBEGITS-"


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

    def _one_hot_encode(self, sequence, vocab):
        indices = [vocab.index(char) for char in sequence]
        return F.one_hot(torch.tensor(indices), num_classes=len(vocab)).float()

    def _split_data(self, split_name):
        random.seed(RANDOM_SEED)
        random.shuffle(self.data)
        if split_name is not None:
            split_idx = int(len(self.data) * 0.8)
            if split_name == "train":
                self.data = self.data[:split_idx]
            elif split_name == "validation":
                self.data = self.data[split_idx: split_idx + int(len(self.data) * 0.1)]
            elif split_name == "test":
                self.data = self.data[split_idx + int(len(self.data) * 0.1):]

    def _load_file(self, filepath: str) -> None:
        print(f"Loading data from {filepath}")
        self.data = torch.load(filepath)
        for d in self.data:
            d["amino_acid"] = self._one_hot_encode("".join(d["amino_acid"]), AA_VOCAB)
            d["secondary_structure"] = self._one_hot_encode("".join(d["secondary_structure"]), SS_VOCAB)

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
