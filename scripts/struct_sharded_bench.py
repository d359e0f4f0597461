#!/usr/bin/env python
"""structure-model sampling sharded over the GPUs of one box (VERDICT r1 item 3 / SURVEY.md section 8(e)): p_sample_loop_sharded under torchrun.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/struct_sharded_bench.py

Workload: the reference's structure_model/sample.py defaults (L = 64, 12 + 12 layers, 8 angle features, bf16, in-kernel Philox
noise keyed by the GLOBAL complex id), 64 complexes per GPU (weak scaling: the reference's batch size per process), T = 200.
Every rank samples its block; rank 0 also re-samples the FIRST block of rank 1's complexes on its own GPU with the same global
ids and checks that the result is bit-identical (the sampling does not depend on the sharding).  Device time = CUDA events
around the local p_sample_loop, max over ranks; the all-gather of the [T,B,L,F] histories is reported separately (it is the
result transfer the reference's single-process call returns, not a data-path collective)."""
import json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist


def main():
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import seqdiff_b200 as sd
    SM = sd.structure_model
    per_gpu, L, T, Fs = 64, 64, int(os.environ.get("T", "200")), 8
    B = per_gpu * world
    torch.manual_seed(0)
    common = dict(max_position_embeddings=L, intermediate_size=1024, num_hidden_layers=12, position_embedding_type="relative_key")
    m = SM.ConditionalBertForDiffusionBase(sd.BertConfig(**common), sd.BertConfig(**common, is_decoder=True, add_cross_attention=True), Fs)
    for blk in (m.receptor_emb, m.timestep_emb):
        torch.nn.init.xavier_uniform_(blk.adaLN_modulation[0].weight)
    m = m.eval().to(dev)
    m.precision = "bf16"
    g = torch.Generator().manual_seed(3)
    nl, nr = torch.randint(5, L + 1, (B,), generator=g), torch.randint(16, L + 1, (B,), generator=g)
    pos = torch.arange(L)[None, :]
    lm, rm = (pos < nl[:, None]).float(), (pos < nr[:, None]).float()
    rseq = torch.nn.functional.one_hot(torch.randint(0, 20, (B, L), generator=g), 20).float() * rm[..., None]
    rang = ((torch.rand(B, L, Fs, generator=g) * 2 - 1) * math.pi) * rm[..., None]
    x_T = SM.modulo_with_wrapped_range(torch.randn(B, L, Fs, generator=g))
    betas = SM.cosine_beta_schedule(T)
    d = lambda t: t.to(dev)  # noqa: E731
    lo, hi = sd.shard_bounds(B, world, rank)
    local_args = lambda a, b: dict(model=m, ligand_mask=d(lm[a:b]), ligand_angle_noise=d(x_T[a:b]), receptor_seq=d(rseq[a:b]), receptor_mask=d(rm[a:b]),  # noqa: E731
                                   receptor_angle=d(rang[a:b]), total_timesteps=T, betas=betas, seed=5, graph_id0=a)
    SM.p_sample_loop(**local_args(lo, hi), keep_history=False)  # warm-up (tuner, graph capture)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        mine = SM.p_sample_loop(**local_args(lo, hi), keep_history=False)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1) / 2], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # the public sharded call (all-gathers the histories to every rank, as the single-process reference call returns them)
    t0 = time.perf_counter()
    full = sd.distributed.p_sample_loop_sharded(m, d(lm), d(x_T), d(rseq), d(rm), d(rang), T, betas, seed=5, graph_id0=0, keep_history=False)
    torch.cuda.synchronize(dev)
    t_pub = time.perf_counter() - t0
    ok_inv = None
    if world > 1 and rank == 0:
        a, b = sd.shard_bounds(B, world, 1)
        other = SM.p_sample_loop(**local_args(a, min(b, a + 8)), keep_history=False)
        ok_inv = bool(torch.equal(other.cpu()[-1] if other.dim() == 4 else other.cpu(), (full[-1] if full.dim() == 4 else full)[a:min(b, a + 8)].cpu()))
    if rank == 0:
        print(json.dumps({"workload": f"structure_model p_sample_loop, {per_gpu} complexes per GPU x {world} GPUs, L={L}, 12+12 layers, T={T}, bf16",
                          "n_gpus": world, "scaling": "weak", "value": B * T / (ms.item() * 1e-3), "unit": "graph-steps/s", "ms_per_sampling_max_rank": ms.item(),
                          "public_sharded_call_s": t_pub, "result_shape": list(full.shape),
                          "sharding_invariant_bit_exact": ok_inv, "finite": bool(torch.isfinite(full).all())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
