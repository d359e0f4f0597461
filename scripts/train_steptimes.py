#!/usr/bin/env python
"""per-step wall times of the training loop (sync after each phase): looks for host-side stalls."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench as Bn
import seqdiff_b200 as sd
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=128); ap.add_argument("--steps", type=int, default=12)
a = ap.parse_args()
dev = torch.device("cuda:0"); Bn.L = 128; torch.manual_seed(0)
common = dict(max_position_embeddings=128, intermediate_size=1024, num_hidden_layers=6, position_embedding_type="relative_key",
              hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)
model = sd.PeptideDiff(sd.BertConfig(**common), sd.BertConfig(**common, is_decoder=True, add_cross_attention=True), list(sd.AA_VOCAB),
                       torch.nn.CrossEntropyLoss(), "cosine", 50, l2_lambda=0.1, learning_rate=5e-5).to(dev).train()
batch, _ = Bn.synthetic_workload(a.batch)
dbatch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
opt = model.configure_optimizers()["optimizer"]
sync = lambda: torch.cuda.synchronize()
for i in range(a.steps):
    sync(); t0 = time.perf_counter()
    loss = model.training_step(dbatch, i)
    t1 = time.perf_counter(); sync(); t2 = time.perf_counter()
    opt.step()
    t3 = time.perf_counter(); sync(); t4 = time.perf_counter()
    print(f"step {i}: training_step host {1e3*(t1-t0):7.2f} ms, +gpu drain {1e3*(t2-t1):7.2f}; opt.step host {1e3*(t3-t2):7.2f}, +drain {1e3*(t4-t3):7.2f}; total {1e3*(t4-t0):7.2f}")
