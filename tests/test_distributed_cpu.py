"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: partitioning, shard-invariant graph ids,
result gathering in rank order, max-over-ranks timing.  The per-shard sampler is injected (no GPU here)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_denoise(batch, model, noise_schedule, transition, diverse, graph_id0=0, x_T=None, **kw):
    """stands in for the CUDA sampler: 'predicts' a string that encodes the global graph id and x_T row"""
    n = batch["ligand_seq"].shape[0]
    ids = [f'{batch["structure_ids"]["pdb_id"][i]}_{batch["structure_ids"]["ligand_chain"][i]}' for i in range(n)]
    pred = [f"g{graph_id0 + i}:{int(x_T[i].sum())}" for i in range(n)]
    true = [f"t{int(batch['ligand_seq'][i].sum())}" for i in range(n)]
    return ids, true, pred, [float(graph_id0 + i) for i in range(n)]


def _make_batch(n):
    return {"ligand_seq": torch.arange(n).float()[:, None, None].expand(n, 4, 20).contiguous(),
            "ligand_attn_mask": torch.ones(n, 4), "ligand_angles": torch.zeros(n, 4, 8), "receptor_seq": torch.zeros(n, 4, 20),
            "receptor_angles": torch.zeros(n, 4, 8), "receptor_attn_mask": torch.ones(n, 4),
            "structure_ids": {"pdb_id": [f"p{i}" for i in range(n)], "ligand_chain": ["A"] * n}}


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import seqdiff_b200 as sd
    batch = _make_batch(n)
    x_T = torch.arange(n).float()[:, None, None].expand(n, 4, 20).contiguous()
    out = sd.denoise_sharded(batch, None, None, None, True, denoise_fn=_fake_denoise, x_T=x_T)
    slow = sd.distributed.max_over_ranks(1.0 + rank)
    sub, gid0 = sd.shard_batch(batch, world, rank)
    q.put((rank, out, slow, sub["ligand_seq"].shape[0], gid0, sub["structure_ids"]["pdb_id"]))
    dist.barrier()
    dist.destroy_process_group()


def _run(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_shard_bounds_partition():
    import seqdiff_b200 as sd
    for n in (0, 1, 7, 64, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [sd.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_denoise_sharded_world2_matches_single_process():
    n = 7  # ragged split: 4 + 3
    import seqdiff_b200 as sd
    batch = _make_batch(n)
    x_T = torch.arange(n).float()[:, None, None].expand(n, 4, 20).contiguous()
    want = _fake_denoise(batch, None, None, None, True, graph_id0=0, x_T=x_T)
    res = _run(2, n)
    for rank, out, slow, nloc, gid0, ids in res:
        assert tuple(out) == tuple(want)           # every rank returns the whole-batch result, in order
        assert slow == 2.0                          # max over ranks
        assert nloc == (4 if rank == 0 else 3) and gid0 == (0 if rank == 0 else 4)
        assert ids == [f"p{i}" for i in range(gid0, gid0 + nloc)]


def _fake_p_sample_loop(model, ligand_mask, x, receptor_seq, receptor_mask, receptor_angle, T, betas, graph_id0=0, noise_steps=None, **kw):
    """stands in for the CUDA structure sampler: entry [k, i] encodes (global graph id, x row, noise slice)"""
    n = x.shape[0]
    gid = torch.arange(graph_id0, graph_id0 + n).float()[None, :, None, None]
    out = gid + x[None] * 0.001 + torch.arange(T).float()[:, None, None, None] * 100
    if noise_steps is not None:
        out = out + noise_steps
    return out


def _struct_worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import seqdiff_b200 as sd
    x = torch.arange(n).float()[:, None, None].expand(n, 4, 8).contiguous()
    z = torch.arange(3 * n).float().reshape(3, n, 1, 1).expand(3, n, 4, 8).contiguous() * 1e-6
    out = sd.p_sample_loop_sharded(None, torch.ones(n, 4), x, torch.zeros(n, 4, 20), torch.ones(n, 4), torch.zeros(n, 4, 8), 3, None,
                                   sample_fn=_fake_p_sample_loop, noise_steps=z)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_struct_p_sample_loop_sharded_world2_matches_single_process():
    n = 5  # ragged split: 3 + 2
    x = torch.arange(n).float()[:, None, None].expand(n, 4, 8).contiguous()
    z = torch.arange(3 * n).float().reshape(3, n, 1, 1).expand(3, n, 4, 8).contiguous() * 1e-6
    want = _fake_p_sample_loop(None, None, x, None, None, None, 3, None, graph_id0=0, noise_steps=z)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_struct_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out in res:
        assert out.shape == (3, n, 4, 8) and torch.equal(out, want)
