#!/usr/bin/env python
"""Data-parallel training on N GPUs: the bucketed all-reduce OVERLAPPED with the backward pass (events recorded by the library per
gradient bucket, NCCL on a communication stream) must give the same weights as the same all-reduce issued AFTER the backward pass
on the training stream, and every rank must end with identical weights (replicated model).  Under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P scripts/train_ddp_check.py
"""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench as Bn


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    import seqdiff_b200 as sd
    Bn.L = 128
    B, steps = 32, 4
    batch, _ = Bn.synthetic_workload(B * world)
    lo, hi = sd.shard_bounds(B * world, world, rank)
    dbatch = {k: (v[lo:hi].to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    common = dict(max_position_embeddings=128, intermediate_size=1024, num_hidden_layers=6, position_embedding_type="relative_key",
                  hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)

    def run(overlap):
        torch.manual_seed(0)
        m = sd.PeptideDiff(sd.BertConfig(**common), sd.BertConfig(**common, is_decoder=True, add_cross_attention=True), list(sd.AA_VOCAB),
                           torch.nn.CrossEntropyLoss(), "cosine", 50, l2_lambda=0.1, learning_rate=2e-4).to(dev).train()
        m.precision = "bf16"
        opt = m.configure_optimizers()["optimizer"]
        opt.overlap = overlap
        g = torch.Generator(device="cpu").manual_seed(1)
        for i in range(steps):
            t_int = torch.randint(0, 51, (hi - lo, 1), generator=g).float().to(dev)  # same draws in both runs
            m.training_step(dbatch, i, t_int=t_int)
            opt.step()
        w = {k: v.detach().float().cpu().clone() for k, v in m.state_dict().items()}
        norm = float(opt.grad_norm)
        m.release()
        return w, norm

    w_ov, n_ov = run(True)
    w_bl, n_bl = run(False)
    num = sum(float((w_ov[k].double() - w_bl[k].double()).pow(2).sum()) for k in w_ov)
    den = sum(float(w_bl[k].double().pow(2).sum()) for k in w_bl)
    rel = math.sqrt(num / den)
    # replicas agree: compare a checksum of the overlapped run's weights across ranks
    chk = torch.tensor([sum(float(v.double().sum()) for v in w_ov.values())], device=dev, dtype=torch.float64)
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    same = all(abs(float(c) - float(allc[0])) <= 1e-6 * abs(float(allc[0])) for c in allc)
    ok = rel < 2e-4 and same and math.isfinite(n_ov)
    if rank == 0:
        print(json.dumps({"world": world, "steps": steps, "graphs_per_gpu": B, "weights_overlap_vs_after_backward_rel_l2": rel,
                          "replicas_identical": same, "grad_norm_last": [n_ov, n_bl], "ok": ok}))
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
