"""Training step (BASELINE configs[3]): gradient parity of the hand-written backward against torch autograd on the oracle, the fused
clip + AdamW update against torch.optim.AdamW, dropout consistency, and the data-parallel plumbing.

Tolerances as written below: fp32 mode -- every gradient tensor within 1e-5 of the oracle's (max-norm relative to the tensor's own
max, with an absolute floor for tensors whose gradient is ~0); 16-bit modes -- global relative L2 error of the flat gradient
vector (bf16 < 3e-2 measured ~1e-2, fp16 < 5e-3) and cosine similarity > 0.999 per large tensor."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

from helpers import O, make_model, sd_pkg

gpu = pytest.mark.gpu
DEV = "cuda:0"


def _case(L, B, NL, n_lig, n_rec, seed, T=50, Lr=None):
    cfg = O.OracleConfig(max_position_embeddings=max(L, Lr or L), num_hidden_layers=NL)
    state = O.init_state_dict(cfg, seed, "B")
    batch = O.synthetic_batch(B, L, n_lig, n_rec, seed + 10)
    if Lr is not None and Lr != L:
        rec = O.synthetic_batch(B, Lr, n_lig, n_rec, seed + 11)
        for k in ("receptor_seq", "receptor_angles", "receptor_attn_mask"):
            batch[k] = rec[k]
    g = torch.Generator().manual_seed(seed + 20)
    t_int = torch.linspace(0.15 * T, 0.85 * T, B).round().reshape(B, 1)  # both noised and un-noised residues in every case
    E = torch.empty(B * L, 20).exponential_(1, generator=g)
    x_t = O.apply_aa_noise(batch["ligand_seq"], t_int, T, O.NoiseScheduleDiscrete("cosine", T), O.BlosumTransition(), E)
    return cfg, state, batch, t_int / T, x_t


# ---------------------------------------------------------------------------------------------------
# CPU: the formulas the kernels implement, checked against autograd before any GPU is involved
# ---------------------------------------------------------------------------------------------------
def test_loss_gradient_formula_matches_autograd():
    """csrc/train_kernels.cu loss_bwd_kernel: dz = [(p - y) + (p - q) - p (log p + H)] / N on the noised rows."""
    g = torch.Generator().manual_seed(0)
    N = 64
    z = (torch.randn(N, 20, generator=g) * 2).requires_grad_(True)
    x0 = F.one_hot(torch.randint(0, 20, (N,), generator=g), 20).float()
    x_t = F.one_hot(torch.randint(0, 20, (N,), generator=g), 20).float()
    batch = {"ligand_attn_mask": torch.ones(1, N), "ligand_seq": x0[None]}
    loss = O.get_loss(z[None], batch, x_t[None])[0]
    loss.backward()
    noised = x_t.argmax(-1) != x0.argmax(-1)
    p = torch.softmax(z.detach(), -1)
    lp = torch.log_softmax(z.detach(), -1)
    ent = -(p * lp).sum(-1, keepdim=True)
    q = torch.softmax(x0, -1)
    want = ((p - x0) + (p - q) - p * (lp + ent)) / noised.sum() * noised[:, None]
    assert torch.allclose(z.grad, want, atol=2e-6, rtol=1e-4), (z.grad - want).abs().max()


def test_linear_warmup_factor_matches_transformers():
    from transformers import get_linear_schedule_with_warmup
    import seqdiff_b200 as sd
    prm = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([prm], lr=1.0)
    sch = get_linear_schedule_with_warmup(opt, num_warmup_steps=15, num_training_steps=150)
    for epoch in range(150):
        assert opt.param_groups[0]["lr"] == pytest.approx(sd.train.linear_warmup_factor(epoch, 15, 150), abs=1e-12)
        opt.step()
        sch.step()


def _ddp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import seqdiff_b200 as sd

    class Flat:  # stand-in for train.FlatParams on CPU: only the flat gradient buffer takes part in the collective
        grads = torch.arange(5000, dtype=torch.float32) * (rank + 1)
        device = torch.device("cpu")
        model = None
        handle = None
        bucket_bounds = [0, 1200, 3100, 5000]  # the library's gradient buckets (seqdiff_train_grad_buckets): reduced tail first
    opt = sd.train.FlatAdamW.__new__(sd.train.FlatAdamW)
    opt.flat, opt.group, opt.grad_comm, opt.last_allreduce_bytes = Flat, None, "fp32", 0
    opt.all_reduce_grads()
    q.put((rank, Flat.grads.clone(), opt.last_allreduce_bytes))
    dist.destroy_process_group()


def test_gradient_all_reduce_world2_gloo():
    """the one collective of the training path: bucketed sum over the data-parallel group (world 2, gloo, CPU)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    import socket
    with socket.socket() as sk:  # a port the OS reports free (a fixed offset from the pid collided with other jobs now and then)
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(60)
    for rank, gsum, nbytes in res:
        assert torch.equal(gsum, torch.arange(5000, dtype=torch.float32) * 3)
        assert nbytes == 5000 * 4


# ---------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------
def _run_train_step(sd, m, batch, t_norm, x_t, **kw):
    flat = sd.train.FlatParams(m)
    terms, logits = sd.train.train_step_tensors(m, flat, batch, t_norm, x_t, want_logits=True, **kw)
    torch.cuda.synchronize()
    return flat, terms, logits


def _compare_grads(flat, m, grads_ref, tol_rel, what):
    worst = (0.0, "")
    got_all, ref_all = [], []
    gmax = max(float(v.abs().max()) for v in grads_ref.values() if v is not None)
    for k, prm in m.named_parameters():
        ref = grads_ref.get(k)
        if k.startswith("receptor_feature_emb.") or k == "timestep_projector.W":
            assert ref is None and k not in flat.table, k  # dead weight: no gradient on either side (quirk Q1)
            continue
        assert ref is not None and k in flat.table, k
        got = flat.grad(k, tuple(prm.shape)).cpu()
        got_all.append(got.reshape(-1))
        ref_all.append(ref.reshape(-1))
        err = float((got - ref).abs().max()) / max(float(ref.abs().max()), 1e-3 * gmax)
        if err > worst[0]:
            worst = (err, k)
    got_all, ref_all = torch.cat(got_all), torch.cat(ref_all)
    l2 = float((got_all - ref_all).norm() / ref_all.norm())
    print(f"{what}: worst per-tensor max-norm rel err {worst[0]:.3e} ({worst[1]}); flat-vector L2 rel err {l2:.3e}; {got_all.numel()} gradient elements")
    if tol_rel is not None:
        assert worst[0] < tol_rel, worst
    return worst[0], l2


CASES = {"small-2layer": (32, 3, 2, (5, 32), (8, 32), 1, None), "Ll!=Lr": (24, 2, 1, (3, 24), (8, 40), 2, 40), "cfg4-shape": (128, 2, 6, (20, 64), (60, 128), 3, None)}


@gpu
@pytest.mark.parametrize("case", list(CASES), ids=list(CASES))
def test_gradient_parity_fp32(case):
    """every one of the live parameter gradients vs torch autograd on the oracle, fp32 mode, dropout off: < 1e-5."""
    sd = sd_pkg()
    L, B, NL, n_lig, n_rec, seed, Lr = CASES[case]
    cfg, state, batch, t_norm, x_t = _case(L, B, NL, n_lig, n_rec, seed, Lr=Lr)
    loss_ref, grads_ref, logits_ref = O.training_grads(state, cfg, batch, t_norm, x_t)
    m = make_model(sd, cfg, state, "fp32", DEV)
    flat, terms, logits = _run_train_step(sd, m, batch, t_norm, x_t)
    assert float((logits.cpu() - logits_ref).abs().max() / logits_ref.abs().max()) < 1e-5
    loss = sd.train.loss_from_terms(terms)
    for a, b in zip(loss, loss_ref):
        assert float(a) == pytest.approx(float(b), rel=2e-5, abs=1e-6, nan_ok=True)
    _compare_grads(flat, m, grads_ref, 1e-5, f"fp32 {case}")
    live = sum(v.numel() for v in grads_ref.values() if v is not None)
    assert flat.live_numel == live
    if case == "cfg4-shape":
        assert 61_000_000 < live < 61_200_000  # ~61.06 M live elements (72.29 M - 11.24 M dead - buffer)
    m.release()


@gpu
@pytest.mark.parametrize("precision,l2_tol", [("bf16", 3e-2), ("fp16", 5e-3)])
def test_gradient_parity_16bit(precision, l2_tol):
    sd = sd_pkg()
    L, B, NL, n_lig, n_rec, seed, Lr = CASES["cfg4-shape"]
    cfg, state, batch, t_norm, x_t = _case(L, B, NL, n_lig, n_rec, seed)
    _, grads_ref, _ = O.training_grads(state, cfg, batch, t_norm, x_t)
    m = make_model(sd, cfg, state, precision, DEV)
    flat, terms, _ = _run_train_step(sd, m, batch, t_norm, x_t)
    _, l2 = _compare_grads(flat, m, grads_ref, None, f"{precision} cfg4-shape")
    assert l2 < l2_tol, l2
    for k, prm in m.named_parameters():
        if k in flat.table and prm.numel() >= 768 * 768:
            a, b = flat.grad(k).cpu(), grads_ref[k].reshape(-1)
            cos = float(torch.dot(a, b) / (a.norm() * b.norm()))
            assert cos > 0.999, (k, cos)
    # the small tensors too (biases, LayerNorm affine, distance embeddings): they vanish in the flat L2 norm, and their gradients come out of
    # the fused rowwise kernels (LayerNorm backward that also emits the bias gradient of the Linear in front of it, embedding backward)
    worst = _worst_small_tensor_cosine(flat, m, grads_ref)
    print(f"{precision}: worst cosine over the small gradient tensors {worst[0]:.5f} ({worst[1]})")
    assert worst[0] > SMALL_COS[precision], worst
    m.release()


SMALL_COS = {"bf16": 0.995, "fp16": 0.9999}  # measured on B200: 0.9996 / 0.99999 (dropout off), 0.9998 / 0.99999 (dropout on)


def _worst_small_tensor_cosine(flat, m, grads_ref):
    gmax = max(float(v.norm()) for v in grads_ref.values() if v is not None)
    worst = (1.0, "")
    for k, prm in m.named_parameters():
        if k not in flat.table or prm.numel() >= 768 * 768:
            continue
        a, b = flat.grad(k).cpu().double(), grads_ref[k].reshape(-1).double()
        if float(b.norm()) < 1e-4 * gmax:  # gradients that are ~0 (e.g. decoder_normalize gates at the reference init) have no direction
            continue
        cos = float(torch.dot(a, b) / (a.norm() * b.norm()))
        if cos < worst[0]:
            worst = (cos, k)
    return worst


@gpu
@pytest.mark.parametrize("precision,l2_tol", [("bf16", 3e-2), ("fp16", 5e-3)])
def test_gradient_parity_16bit_with_dropout(precision, l2_tol):
    """p = 0.1 in the 16-bit modes against the fp32 mode UNDER THE SAME MASKS (masks are keyed by (seed, site, step, element), not by the
    precision): the 16-bit step takes the fused paths the fp32 parity mode does not -- dropout + residual folded into the LayerNorm pass
    with the mask handed to the backward pass as bits, LayerNorm backward that also writes the masked 16-bit operand and the bias
    gradient, the tcgen05 attention kernels with their keep bits.  A mask applied at the wrong element shows up as a cosine of ~0.9."""
    sd = sd_pkg()
    L, B, NL, n_lig, n_rec, seed, Lr = CASES["cfg4-shape"]
    cfg, state, batch, t_norm, x_t = _case(L, B, NL, n_lig, n_rec, seed)
    kw = dict(p_hidden=0.1, p_attn=0.1, seed=23, step=4)
    m32 = make_model(sd, cfg, state, "fp32", DEV)
    m32.train()
    flat32, terms32, _ = _run_train_step(sd, m32, batch, t_norm, x_t, **kw)
    ref = {k: flat32.grad(k, tuple(prm.shape)).cpu().clone() for k, prm in m32.named_parameters() if k in flat32.table}
    loss32 = float(sd.train.loss_from_terms(terms32)[0])
    m32.release()
    m = make_model(sd, cfg, state, precision, DEV)
    m.train()
    flat, terms, _ = _run_train_step(sd, m, batch, t_norm, x_t, **kw)
    assert float(sd.train.loss_from_terms(terms)[0]) == pytest.approx(loss32, rel=2e-2)
    got_all = torch.cat([flat.grad(k).cpu() for k in ref])
    ref_all = torch.cat([v.reshape(-1) for v in ref.values()])
    l2 = float((got_all - ref_all).norm() / ref_all.norm())
    worst = _worst_small_tensor_cosine(flat, m, ref)
    print(f"{precision}, dropout 0.1, vs fp32 mode with the same masks: flat L2 rel err {l2:.3e}; worst small-tensor cosine {worst[0]:.5f} ({worst[1]})")
    assert l2 < l2_tol, l2
    assert worst[0] > SMALL_COS[precision], worst
    for k, v in ref.items():
        if v.numel() >= 768 * 768:
            a, b = flat.grad(k).cpu(), v.reshape(-1)
            assert float(torch.dot(a, b) / (a.norm() * b.norm())) > 0.999, k
    m.release()


def _attn_ref(q, k, v, E, mask, P):
    """attention core of transformers 4.38.2 BertSelfAttention in torch, per head layout [B, L, heads*64]."""
    B, Lq, H = q.shape
    Lk, nh = k.shape[1], H // 64
    qh, kh, vh = (x.view(B, -1, nh, 64).permute(0, 2, 1, 3) for x in (q, k, v))
    s = qh @ kh.transpose(-1, -2)
    if E is not None:
        dist = torch.arange(Lq).view(-1, 1) - torch.arange(Lk).view(1, -1)
        s = s + torch.einsum("bhld,lrd->bhlr", qh, E[dist + P - 1])
    s = s / 8.0 + ((1.0 - mask) * -10000.0)[:, None, None, :]
    return (torch.softmax(s, -1) @ vh).permute(0, 2, 1, 3).reshape(B, Lq, H)


@gpu
@pytest.mark.parametrize("B,heads,Lq,Lk,P,rel", [(3, 12, 128, 128, 128, True), (2, 4, 100, 77, 128, True), (2, 12, 48, 128, 128, False), (5, 2, 17, 33, 64, True),
                                                 (4, 12, 128, 128, 128, False), (3, 4, 100, 76, 128, False), (37, 12, 64, 128, 128, False), (5, 2, 17, 33, 64, False),
                                                 (30, 12, 128, 128, 128, True), (3, 4, 100, 76, 128, True), (7, 12, 64, 64, 64, True)])
def test_attention_train_kernels(B, heads, Lq, Lk, P, rel):
    """training attention at operator level: (1) fp32 SIMT forward / backward == torch autograd on the attention core, 1e-5;
    (2) the tensor-core kernels (bf16 / fp16 operands) agree with the SIMT kernels run on the same 16-bit inputs -- with and
    without dropout (same Philox masks on both sides) -- to operand-rounding level."""
    sd = sd_pkg()
    lib = sd.lib()
    from helpers import BF16, FP16, FP32, stream_ptr
    p = sd._cabi.ptr
    g = torch.Generator().manual_seed(B * 1000 + Lq + Lk)
    Hh = heads * 64
    q, k, v, do = (torch.randn(B, L_, Hh, generator=g) * 0.7 for L_ in (Lq, Lk, Lk, Lq))
    E = torch.randn(2 * P - 1, 64, generator=g) * 0.5 if rel else None
    lens = torch.randint(1, Lk + 1, (B,), generator=g)
    mask = (torch.arange(Lk)[None, :] < lens[:, None]).float()

    def run(prec, impl, dt, pdrop, fwd_impl=None, bwd_impl=None):
        dev = lambda x: None if x is None else x.to(DEV).to(dt).contiguous()
        dq_, dk_, dv_, do_, dE_ = dev(q), dev(k), dev(v), dev(do), dev(E)
        m_ = mask.to(DEV)
        out = torch.empty(B, Lq, Hh, device=DEV, dtype=dt)
        gq, gk, gv = torch.empty_like(dq_), torch.empty_like(dk_), torch.empty_like(dv_)
        gE = torch.zeros(2 * P - 1, 64, device=DEV) if rel else None
        st = stream_ptr()
        assert lib.seqdiff_op_attention_train_fwd(prec, impl if fwd_impl is None else fwd_impl, B, heads, Lq, Lk, p(dq_), Hh, p(dk_), Hh, p(dv_), Hh, p(dE_), P, p(m_), pdrop, 9, 3, 1,
                                                  p(out), st) == 0, lib.seqdiff_last_error()
        assert lib.seqdiff_op_attention_train_bwd(prec, impl if bwd_impl is None else bwd_impl, B, heads, Lq, Lk, p(dq_), Hh, p(dk_), Hh, p(dv_), Hh, p(dE_), P, p(m_), pdrop, 9, 3, 1,
                                                  p(do_), p(gq), p(gk), p(gv), p(gE), st) == 0, lib.seqdiff_last_error()
        torch.cuda.synchronize()
        return [x.float().cpu() if x is not None else None for x in (out, gq, gk, gv, gE)]

    # (1) fp32 SIMT vs autograd
    leaves = [x.clone().requires_grad_(True) for x in (q, k, v)] + ([E.clone().requires_grad_(True)] if rel else [])
    ref_out = _attn_ref(leaves[0], leaves[1], leaves[2], leaves[3] if rel else None, mask, P)
    ref_out.backward(do)
    ref = [ref_out.detach()] + [x.grad for x in leaves[:3]] + [leaves[3].grad if rel else None]
    got = run(FP32, 1, torch.float32, 0.0)
    for name, a, b in zip(("out", "dq", "dk", "dv", "dE"), got, ref):
        if b is not None:
            err = float((a - b).abs().max() / b.abs().max())
            assert err < 1e-5, (name, err)
    # (2) tensor-core vs SIMT on the same 16-bit inputs
    for prec, dt, tol in ((BF16, torch.bfloat16, 2e-2), (FP16, torch.float16, 3e-3)):
        for pdrop in (0.0, 0.1):
            tc, simt = run(prec, 0, dt, pdrop), run(prec, 1, dt, pdrop)
            for name, a, b in zip(("out", "dq", "dk", "dv", "dE"), tc, simt):
                if b is not None:
                    err = float((a - b).norm() / b.norm())
                    assert err < tol, (name, prec, pdrop, err)
            if pdrop == 0 or Lk % 4 == 0:  # the tcgen05 backward (what the training step runs)
                pipe = run(prec, 0, dt, pdrop, bwd_impl=3)
                for name, a, b in zip(("out", "dq", "dk", "dv", "dE"), pipe, simt):
                    if b is not None:
                        err = float((a - b).norm() / b.norm())
                        assert err < tol, (name + " (tcgen05 backward)", prec, pdrop, err)
            if pdrop > 0:
                assert float((simt[0] - run(prec, 1, dt, 0.0)[0]).abs().max()) > 1e-3  # the mask really was applied
                if Lk % 4 == 0:  # the forward the training step runs: pipelined tcgen05 kernel, dropout inside, SAME Philox masks
                    pipe_out = run(prec, 0, dt, pdrop, fwd_impl=2)[0]
                    err = float((pipe_out - simt[0]).norm() / simt[0].norm())
                    assert err < tol, ("out (tcgen05 pipe + dropout)", prec, err)


@gpu
def test_adamw_and_clip_match_torch():
    """the fused 1/world + clip_grad_norm_(1.0) + AdamW kernel against the real torch objects, three steps, fed the SAME gradients
    (the oracle's, written into the flat buffer by name) so that only the optimizer arithmetic is compared -- Adam's update
    g / (|g| + eps) amplifies 1e-9-level gradient differences on near-zero elements, which would test the backward, not the update."""
    sd = sd_pkg()
    L, B, NL, n_lig, n_rec, seed, Lr = CASES["small-2layer"]
    cfg, state, batch, t_norm, x_t = _case(L, B, NL, n_lig, n_rec, seed)
    m = make_model(sd, cfg, state, "fp32", DEV)
    flat = sd.train.FlatParams(m)
    opt = sd.train.FlatAdamW(flat, lr=5e-5, weight_decay=0.1, gradient_clip=1.0)
    cur, ostate = dict(state), {}
    for step in range(3):
        _, grads_ref, _ = O.training_grads(cur, cfg, batch, t_norm, x_t)
        if step == 2:  # a step below the clip threshold as well
            grads_ref = {k: (None if v is None else v * (0.5 / 80.0)) for k, v in grads_ref.items()}
        cur, norm_ref = O.adamw_reference_step(cur, grads_ref, ostate, lr=5e-5, weight_decay=0.1, max_norm=1.0)
        sd.train.train_step_tensors(m, flat, batch, t_norm, x_t)  # (exercises the real path; its result is then overwritten)
        for k, v in grads_ref.items():
            if v is not None:
                flat.grad(k).copy_(v.reshape(-1))
        opt.step()
        # the kernel accumulates the squared norm in fp64; torch's clip_grad_norm_ (fp32 per-tensor norms) agrees to ~1e-4
        norm64 = math.sqrt(sum(float((v.double() ** 2).sum()) for v in grads_ref.values() if v is not None))
        assert float(opt.grad_norm) == pytest.approx(norm64, rel=2e-6)
        assert float(opt.grad_norm) == pytest.approx(norm_ref, rel=2e-4)
    m._flat = flat
    sd.train.pull_weights(m)
    worst = 0.0
    for k, v in m.state_dict().items():
        ref = cur[k]
        worst = max(worst, float((v.cpu() - ref).abs().max()) / max(float(ref.abs().max()), 1e-6))
        if k.startswith("receptor_feature_emb."):
            assert torch.equal(v.cpu(), state[k])  # no gradient -> untouched (not even weight decay), as torch skips grad=None
    print(f"weights after 3 AdamW steps: worst rel err {worst:.3e}")
    assert worst < 2e-6
    # the forward now runs on the UPDATED weights (packed copies were refreshed in place)
    args = (t_norm, x_t, batch["ligand_angles"], batch["ligand_attn_mask"], batch["receptor_seq"], batch["receptor_angles"], batch["receptor_attn_mask"])
    with torch.no_grad():
        want = O.denoiser_forward(cur, cfg, *args)
        got = m(*[a.to(DEV) for a in args]).cpu()
    assert float((got - want).abs().max() / want.abs().max()) < 1e-5
    m.release()


@gpu
def test_dropout_masks_are_consistent_between_forward_and_backward():
    """p = 0.1: same (seed, step) -> identical gradients; another step -> different masks; and the gradient is the derivative of
    the loss UNDER THOSE MASKS: a central finite difference along a random direction (fp32 mode) matches g . d."""
    sd = sd_pkg()
    L, B, NL, n_lig, n_rec, seed, Lr = CASES["small-2layer"]
    cfg, state, batch, t_norm, x_t = _case(L, B, NL, n_lig, n_rec, seed)
    m = make_model(sd, cfg, state, "fp32", DEV)
    m.train()
    kw = dict(p_hidden=0.1, p_attn=0.1, seed=11, step=5)
    flat, terms, _ = _run_train_step(sd, m, batch, t_norm, x_t, **kw)
    g0 = flat.grads.clone()
    loss0 = float(sd.train.loss_from_terms(terms)[0])
    flat2, terms2, _ = _run_train_step(sd, m, batch, t_norm, x_t, **kw)
    # same masks -> same gradients (up to the summation order of the atomics that fold bias / LayerNorm / dE gradients)
    assert float((flat2.grads - g0).abs().max()) < 1e-5 * float(g0.abs().max()) and float(sd.train.loss_from_terms(terms2)[0]) == loss0
    flat3, terms3, _ = _run_train_step(sd, m, batch, t_norm, x_t, **dict(kw, step=6))
    assert float((flat3.grads - g0).abs().max()) > 1e-3 * float(g0.abs().max()) and float(sd.train.loss_from_terms(terms3)[0]) != loss0
    m.eval()
    _, terms_eval, _ = _run_train_step(sd, m, batch, t_norm, x_t, **kw)  # eval(): dropout off whatever p says
    assert float(sd.train.loss_from_terms(terms_eval)[0]) != loss0
    m.train()
    # directional derivative
    gen = torch.Generator().manual_seed(3)
    names = [k for k in flat.table if k.endswith("weight") and "LayerNorm" not in k and "layer_norm" not in k]
    direction = {k: torch.randn(state[k].shape, generator=gen) * state[k].abs().mean() for k in names}
    gd = sum(float((flat.grad(k, tuple(state[k].shape)).cpu().double() * direction[k].double()).sum()) for k in names)
    eps = 2e-3
    losses = []
    for sgn in (+1, -1):
        st = {k: (v + sgn * eps * direction[k] if k in direction else v) for k, v in state.items()}
        m.load_state_dict(st, strict=True)
        m.to(DEV)
        _, t_, _ = _run_train_step(sd, m, batch, t_norm, x_t, **kw)
        losses.append(float(sd.train.loss_from_terms(t_)[0].double()))
    fd = (losses[0] - losses[1]) / (2 * eps)
    print(f"dropout on: directional derivative {gd:.6f} vs central finite difference {fd:.6f}")
    assert fd == pytest.approx(gd, rel=2e-2, abs=1e-3)
    m.release()


@gpu
def test_optimizer_checkpoint_resume_is_exact():
    """4 optimizer steps in one go == 2 steps, optimizer + model checkpoint (state_dict), a NEW model + optimizer restored from it,
    2 more steps: the same weights (dropout masks are keyed by the step counter, AdamW by its step count and moments)."""
    sd = sd_pkg()
    L, B = 64, 4
    sd.sample.DEVICE = torch.device(DEV)
    sd.sample.CONFIG.update(max_seq_len=L, timesteps=50, num_hidden_layers=2)
    batch = O.synthetic_batch(B, L, (20, 60), (30, 64), 5)
    t_int = torch.full((B, 1), 30.0)
    E = torch.empty(B * L, 20).exponential_(1, generator=torch.Generator().manual_seed(9))

    def fresh(state=None):
        torch.manual_seed(0)
        m = sd.sample.get_model()
        if state is not None:
            m.load_state_dict(state, strict=True)
        m.lr, m.lr_scheduler, m.precision = 2e-4, None, "fp16"
        return m.train()

    def steps(m, opt, first, n):
        for i in range(first, first + n):
            m._train_steps = i            # the dropout step counter of training_step (it increments before use)
            m._noise_calls = i
            m.training_step(batch, i, t_int=t_int, noise_E=E)
            opt.step()

    try:
        a = fresh()
        oa = a.configure_optimizers()["optimizer"]
        steps(a, oa, 0, 4)
        want = {k: v.detach().cpu().clone() for k, v in a.state_dict().items()}
        b = fresh()
        ob = b.configure_optimizers()["optimizer"]
        steps(b, ob, 0, 2)
        ck_model = {k: v.detach().cpu().clone() for k, v in b.state_dict().items()}
        ck_opt = {k: (v.detach().cpu().clone() if torch.is_tensor(v) else v) for k, v in ob.state_dict().items()}
        c = fresh(ck_model)
        oc = c.configure_optimizers()["optimizer"]
        oc.load_state_dict(ck_opt)
        assert oc.step_count == 2 and torch.equal(oc.exp_avg.cpu(), ck_opt["exp_avg"]) and torch.equal(oc.exp_avg_sq.cpu(), ck_opt["exp_avg_sq"])
        steps(c, oc, 2, 2)
        got = c.state_dict()
        # (the gradient buffers are filled with floating-point atomics -- split-K, column sums, dE -- so two runs agree to rounding, not
        #  bit for bit; a resume that lost the moments or the step count would be off by ~lr per element and step: 5e-3 relative)
        num = sum(float((want[k].double() - got[k].cpu().double()).pow(2).sum()) for k in want)
        den = sum(float(want[k].double().pow(2).sum()) for k in want)
        assert math.sqrt(num / den) < 2e-4, math.sqrt(num / den)
        with pytest.raises(ValueError):
            oc.load_state_dict({**ck_opt, "exp_avg": ck_opt["exp_avg"][:-4]})
        for m in (a, b, c):
            m.release()
    finally:
        sd.sample.CONFIG.update(num_hidden_layers=6)


@gpu
def test_training_step_api_and_fit_reduce_the_loss():
    """PeptideDiff.training_step / configure_optimizers / train.fit: a few steps on one fixed batch drive the loss down, the trained
    weights come back through state_dict(), and sampling still works on the updated handle."""
    sd = sd_pkg()
    torch.manual_seed(0)
    L, B = 64, 4
    sd.sample.DEVICE = torch.device(DEV)
    sd.sample.CONFIG.update(max_seq_len=L, timesteps=50, num_hidden_layers=2)
    model = sd.sample.get_model()
    sd.sample.CONFIG.update(num_hidden_layers=6)
    model.lr, model.lr_scheduler, model.precision = 2e-4, None, "bf16"
    before = {k: v.clone() for k, v in model.state_dict().items()}
    batch = O.synthetic_batch(B, L, (20, 60), (30, 64), 5)
    model.train()
    opt = model.configure_optimizers()["optimizer"]
    t_int = torch.full((B, 1), 40.0)
    g = torch.Generator().manual_seed(9)
    E = torch.empty(B * L, 20).exponential_(1, generator=g)
    losses = []
    for i in range(12):
        losses.append(float(model.training_step(batch, i, t_int=t_int, noise_E=E)))
        opt.step()
    print("losses:", [round(x, 3) for x in losses])
    assert all(math.isfinite(x) for x in losses) and losses[-1] < losses[0] - 0.2
    after = model.state_dict()
    changed = [k for k in before if not torch.equal(before[k].cpu(), after[k].cpu())]
    assert any(k.startswith("decoder.layer.0") for k in changed) and not any(k.startswith("receptor_feature_emb") for k in changed)
    hist = sd.train.fit(model, [batch], max_epochs=2, log=None)
    assert len(hist) == 2 and all(math.isfinite(h) for h in hist)
    model.eval()
    out = sd.denoise_tensors(batch, model, sd.PredefinedNoiseScheduleDiscrete("cosine", 5), sd.BlosumTransition(x_classes=20), True, timesteps=5)
    assert torch.isfinite(out).all()
    # the packed operand copies the optimizer refreshes in place (batched copy / convert / transpose kernels) must equal what a
    # fresh handle builds from the same trained weights: eval forward of both, every precision, bit for bit
    trained = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    sd.sample.CONFIG.update(num_hidden_layers=2)
    fresh = sd.sample.get_model()
    sd.sample.CONFIG.update(num_hidden_layers=6)
    fresh.load_state_dict(trained, strict=True)
    fresh = fresh.eval()
    x_t = O.generate_discrete_noise(B, L, generator=torch.Generator().manual_seed(3))
    args = [a.to(DEV) for a in (torch.full((B, 1), 17.0), x_t, batch["ligand_angles"], batch["ligand_attn_mask"], batch["receptor_seq"],
                                batch["receptor_angles"], batch["receptor_attn_mask"])]
    for prec in ("bf16", "fp16", "fp32"):
        model.precision = fresh.precision = prec
        with torch.no_grad():
            assert torch.equal(model(*args), fresh(*args)), prec
    fresh.release()
    model.release()
