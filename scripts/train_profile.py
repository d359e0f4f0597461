#!/usr/bin/env python
"""Per-kernel-tag device time of ONE training step (library event profiler: an event behind every launch), for the cfg-4 shapes.
    python scripts/train_profile.py [--batch 16] [--p 0.1] [--precision bf16]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench as Bn
import seqdiff_b200 as sd

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--p", type=float, default=0.1)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--L", type=int, default=128)
a = ap.parse_args()
dev = torch.device("cuda:0")
Bn.L = a.L
torch.manual_seed(0)
common = dict(max_position_embeddings=a.L, intermediate_size=1024, num_hidden_layers=6, position_embedding_type="relative_key",
              hidden_dropout_prob=a.p, attention_probs_dropout_prob=a.p)
model = sd.PeptideDiff(sd.BertConfig(**common), sd.BertConfig(**common, is_decoder=True, add_cross_attention=True), list(sd.AA_VOCAB),
                       torch.nn.CrossEntropyLoss(), "cosine", 50, l2_lambda=0.1, learning_rate=5e-5).to(dev).train()
model.precision = a.precision
batch, _ = Bn.synthetic_workload(a.batch)
dbatch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
opt = model.configure_optimizers()["optimizer"]
for i in range(3):
    model.training_step(dbatch, i); opt.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(5):
    model.training_step(dbatch, i); opt.step()
e1.record(); torch.cuda.synchronize()
print(f"batch {a.batch} L {a.L} p {a.p} {a.precision}: {e0.elapsed_time(e1)/5:.3f} ms per step (eager, incl. host gaps)")
if os.environ.get("SEQDIFF_PROFILER_RANGE") == "1":  # ncu --profile-from-start off: exactly one training step + optimizer step
    torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStart()
    model.training_step(dbatch, 0); opt.step()
    torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStop()
prof = sd._cabi.profile(lambda: (model.training_step(dbatch, 0), opt.step()))
tot = sum(v[0] for v in prof.values())
print(f"sum of kernel gaps {tot:.3f} ms, {sum(v[1] for v in prof.values())} launches")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:24s} {v[0]:8.3f} ms  {v[1]:5d} launches  {100*v[0]/tot:5.1f} %")
