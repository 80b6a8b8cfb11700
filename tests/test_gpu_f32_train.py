"""The fp32 training path (ctk.set_precision(model, "fp32") -> ctk.train_f32.TrainEngineF32): float32 arithmetic on the
CUDA cores, the reference's own precision (train_model.py:419-424).  Kernel by kernel against fp64 PyTorch-CPU references,
then end to end against the fp32 oracle: loss, every gradient, BatchNorm buffers, and bit-reproducibility.

Tolerances are fp32-level: 1e-6 relative L2 for single kernels; for the whole network the bound is what the reference's own
arithmetic allows -- train-mode BatchNorm amplifies a 1e-6 relative perturbation of the input to 6e-4 (double-branch) .. 5e-3
(single-branch) of the whole gradient (measured in pure fp32 on the CPU), and fp32 summation order alone is a
1e-7 perturbation of everything.
"""
import os
from ctypes import c_double, c_float, c_int, c_longlong

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import crosstalk_oracle as orc

pytestmark = pytest.mark.gpu

P_DROP = {"single": 0.1, "double": 0.5}


@pytest.fixture(scope="module")
def L():
    from ctk import _lib
    _lib.load()
    return _lib


def ll(v):
    return c_longlong(int(v))


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-300)).item()


def _build(kind):
    import ctk
    torch.manual_seed(0)
    if kind == "single":
        return ctk.AdvancedRegressionModel(initial_filters=128, num_conv_blocks=6)
    return ctk.SimplifiedTwoBranchRegressionModel(initial_filters_per_branch=64)


# ------------------------------------------------------------------------------------------------ kernels
@pytest.mark.parametrize("n,H,W,cin,cout,layout", [(2, 32, 32, 64, 128, "nhwc"), (1, 16, 32, 128, 64, "nhwc"),
                                                   (2, 32, 64, 1, 64, "plane1"), (2, 32, 32, 2, 128, "nchw"),
                                                   (4, 8, 8, 128, 128, "nhwc"), (3, 8, 8, 64, 64, "nhwc")])
def test_conv_forward_dgrad_and_wgrad_f32(L, n, H, W, cin, cout, layout):
    torch.manual_seed(0)
    w = (torch.randn(cout, cin, 3, 3) / (3 * cin ** 0.5)).requires_grad_(True)
    if layout == "nhwc":
        xs = torch.randn(n, H, W, cin)
        x_nchw = xs.permute(0, 3, 1, 2).contiguous()
        strides, off, x_dev = (H * W * cin, W * cin, cin, 1), 0, xs.cuda()
    else:
        xs = torch.rand(n, 2, H, W)                        # the reference's NCHW input, read in place
        coff = 1 if layout == "plane1" else 0
        x_nchw = xs[:, coff:coff + cin].contiguous()
        strides, off, x_dev = (2 * H * W, W, 1, H * W), coff * H * W, xs.cuda()
    xr = x_nchw.double().requires_grad_(True)
    wr = w.detach().double().requires_grad_(True)
    ref = F.conv2d(xr, wr, padding=1)
    dy = torch.randn(n, H, W, cout)
    ref.backward(dy.permute(0, 3, 1, 2).double())
    wd = w.detach().cuda()
    wk = torch.empty(9 * cin, cout, device="cuda")
    L.call("ctk_pack_conv_weight_f32", L.ptr(wd), c_int(cout), c_int(cin), c_int(0), L.ptr(wk), L.stream())
    y = torch.empty(n, H, W, cout, device="cuda")
    xin = x_dev.view(-1)[off:]
    L.call("ctk_conv3x3_f32", L.ptr(xin), ll(strides[0]), ll(strides[1]), ll(strides[2]), ll(strides[3]), c_int(n), c_int(H),
           c_int(W), c_int(cin), L.ptr(wk), c_int(cout), L.ptr(y), L.stream())
    assert rel_l2(y.cpu(), ref.detach().permute(0, 2, 3, 1)) < 1e-6
    dyd = dy.cuda()
    dw, dw2 = torch.empty(cout, cin, 3, 3, device="cuda"), torch.empty(cout, cin, 3, 3, device="cuda")
    ws = L.workspace("ctk_conv3x3_wgrad_f32_workspace_bytes", n, H, W, cin, cout)
    for dst in (dw, dw2):
        L.call("ctk_conv3x3_wgrad_f32", L.ptr(dyd), L.ptr(xin), ll(strides[0]), ll(strides[1]), ll(strides[2]), ll(strides[3]),
               c_int(n), c_int(H), c_int(W), c_int(cin), c_int(cout), L.ptr(dst), ws[1], ws[2], L.stream())
    assert torch.equal(dw, dw2)
    assert rel_l2(dw.cpu(), wr.grad) < 1e-6
    if layout == "nhwc" and cin % 64 == 0:
        wg = torch.empty(9 * cout, cin, device="cuda")
        L.call("ctk_pack_conv_weight_f32", L.ptr(wd), c_int(cout), c_int(cin), c_int(1), L.ptr(wg), L.stream())
        dx = torch.empty(n, H, W, cin, device="cuda")
        L.call("ctk_conv3x3_f32", L.ptr(dyd), ll(H * W * cout), ll(W * cout), ll(cout), ll(1), c_int(n), c_int(H), c_int(W),
               c_int(cout), L.ptr(wg), c_int(cin), L.ptr(dx), L.stream())
        assert rel_l2(dx.cpu(), xr.grad.permute(0, 2, 3, 1)) < 1e-6


@pytest.mark.parametrize("C,flatten", [(64, False), (256, False), (512, True)])
def test_bn_pool_forward_and_backward_f32(L, C, flatten):
    torch.manual_seed(1)
    n, H, W = 3, 16, 32
    Hp, Wp = H // 2, W // 2
    y = torch.randn(n, H, W, C) * 0.7 + 0.3
    gamma, beta, bias = torch.randn(C), 0.3 * torch.randn(C), 0.05 * torch.randn(C)
    dp = torch.randn(n, Hp, Wp, C)
    yr = y.double().permute(0, 3, 1, 2).clone().requires_grad_(True)
    g, b = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    rm, rv = torch.zeros(C, dtype=torch.float64), torch.ones(C, dtype=torch.float64)
    z = F.batch_norm(yr + bias.double()[None, :, None, None], rm, rv, g, b, training=True, momentum=0.1, eps=1e-5)
    pooled = F.max_pool2d(F.leaky_relu(z, 0.01), 2)
    pooled.backward(dp.permute(0, 3, 1, 2).double())
    yd, gd, bd, biasd = y.cuda(), gamma.cuda(), beta.cuda(), bias.cuda()
    sums = torch.empty(2 * C, device="cuda", dtype=torch.float64)
    ws = L.workspace("ctk_channel_sums_f64_workspace_bytes", C)
    L.call("ctk_channel_stats_f32", L.ptr(yd), ll(n * H * W), c_int(C), L.ptr(sums), ws[1], ws[2], L.stream())
    np.testing.assert_allclose(sums[:C].cpu().numpy(), y.double().sum((0, 1, 2)).numpy(), rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(sums[C:].cpu().numpy(), (y.double() ** 2).sum((0, 1, 2)).numpy(), rtol=1e-12)
    scale, shift, mean, invstd = (torch.empty(C, device="cuda") for _ in range(4))
    rmd, rvd = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    nbt = torch.zeros((), dtype=torch.int64, device="cuda")
    L.call("ctk_bn_finalize_f64", L.ptr(sums), c_double(float(n * H * W)), L.ptr(biasd), L.ptr(gd), L.ptr(bd), L.ptr(rmd),
           L.ptr(rvd), L.ptr(nbt), c_float(0.1), c_float(1e-5), c_int(C), L.ptr(scale), L.ptr(shift), L.ptr(mean),
           L.ptr(invstd), L.stream())
    assert int(nbt) == 1
    np.testing.assert_allclose(rmd.cpu().numpy(), rm.numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(rvd.cpu().numpy(), rv.numpy(), rtol=1e-6, atol=1e-7)
    if flatten:                                            # nn.Flatten order [n][c * HpWp + p], inside a wider row
        feat_cols = (C + 8) * Hp * Wp
        out = torch.zeros(n, feat_cols, device="cuda")
        ostr, ooff = (feat_cols, 1, Hp * Wp), 8 * Hp * Wp
        dpd = torch.zeros(n, feat_cols, device="cuda")
        dpd.view(n, C + 8, Hp * Wp)[:, 8:] = dp.permute(0, 3, 1, 2).reshape(n, C, Hp * Wp).cuda()
    else:
        out = torch.zeros(n, Hp, Wp, C, device="cuda")
        ostr, ooff = (Hp * Wp * C, C, 1), 0
        dpd = dp.cuda()
    L.call("ctk_bn_act_pool_fwd_f32", L.ptr(yd), c_int(n), c_int(H), c_int(W), c_int(C), L.ptr(mean), L.ptr(invstd), L.ptr(gd),
           L.ptr(bd), c_float(0.01), L.ptr(out.view(-1)[ooff:]), ll(ostr[0]), ll(ostr[1]), ll(ostr[2]), L.stream())
    got = out.view(n, C + 8, Hp, Wp)[:, 8:].permute(0, 2, 3, 1) if flatten else out
    assert rel_l2(got.cpu(), pooled.detach().permute(0, 2, 3, 1)) < 1e-6
    if flatten:
        assert out.view(n, C + 8, Hp * Wp)[:, :8].abs().max().item() == 0
    s64 = torch.empty(2 * C, device="cuda", dtype=torch.float64)
    s32 = torch.empty(2 * C, device="cuda")
    L.call("ctk_bn_bwd_reduce_f32", L.ptr(yd), L.ptr(dpd.view(-1)[ooff:]), ll(ostr[0]), ll(ostr[1]), ll(ostr[2]), c_int(n), c_int(H),
           c_int(W), c_int(C), L.ptr(mean), L.ptr(invstd), L.ptr(gd), L.ptr(bd), c_float(0.01), L.ptr(s64), L.ptr(s32), ws[1],
           ws[2], L.stream())
    np.testing.assert_allclose(s32[:C].cpu().numpy(), b.grad.numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(s32[C:].cpu().numpy(), g.grad.numpy(), rtol=1e-5, atol=1e-5)
    dy = torch.empty(n, H, W, C, device="cuda")
    L.call("ctk_bn_bwd_apply_f32", L.ptr(yd), L.ptr(dpd.view(-1)[ooff:]), ll(ostr[0]), ll(ostr[1]), ll(ostr[2]), c_int(n), c_int(H),
           c_int(W), c_int(C), L.ptr(mean), L.ptr(invstd), L.ptr(gd), L.ptr(bd), L.ptr(s64), c_double(float(n * H * W)),
           c_float(0.01), L.ptr(dy), L.stream())
    assert rel_l2(dy.cpu(), yr.grad.permute(0, 2, 3, 1)) < 2e-6


def test_gemm_f32_layouts(L):
    torch.manual_seed(2)
    n, f, K = 24, 96, 1000
    feat, w, dz = torch.randn(n, K), torch.randn(f, K), torch.randn(n, f)
    bias = torch.randn(f)
    fd, wd, dzd, bd = feat.cuda(), w.cuda(), dz.cuda(), bias.cuda()
    z = torch.empty(n, f, device="cuda")
    L.call("ctk_gemm_f32", L.ptr(fd), ll(K), ll(1), L.ptr(wd), ll(K), ll(1), L.ptr(bd), c_int(n), c_int(f), c_int(K), L.ptr(z),
           ll(f), L.stream())
    assert rel_l2(z.cpu(), feat.double() @ w.double().t() + bias.double()) < 2e-7
    dw = torch.empty(f, K, device="cuda")
    L.call("ctk_gemm_f32", L.ptr(dzd), ll(1), ll(f), L.ptr(fd), ll(1), ll(K), L.ptr(None), c_int(f), c_int(K), c_int(n),
           L.ptr(dw), ll(K), L.stream())
    assert rel_l2(dw.cpu(), dz.double().t() @ feat.double()) < 2e-7
    dx = torch.empty(n, K, device="cuda")
    L.call("ctk_gemm_f32", L.ptr(dzd), ll(f), ll(1), L.ptr(wd), ll(1), ll(K), L.ptr(None), c_int(n), c_int(K), c_int(f),
           L.ptr(dx), ll(K), L.stream())
    assert rel_l2(dx.cpu(), dz.double() @ w.double()) < 2e-7


# ------------------------------------------------------------------------------------------------ end to end
# whole-gradient relative L2 against the fp32 oracle: fp32 summation order (a ~1e-7 relative perturbation of every
# intermediate) times the x600 .. x5000 amplification measured for these networks
# measured on B200: 2.6e-4 (double) / 3.7e-4 (single), losses within 5e-7 / 3e-6, outputs within 2.7e-7 / 4.7e-6
WHOLE_GRAD_BOUND = {"double": 1e-3, "single": 1.5e-3}


@pytest.mark.parametrize("kind", ["double", "single"])
def test_fp32_training_step_matches_the_fp32_oracle(kind):
    import ctk
    n = 16
    x, y = orc.synthetic_batch(n, seed=4321)
    model = _build(kind)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    masks = orc.dropout_masks(n, P_DROP[kind], seed=5)
    torch.set_num_threads(os.cpu_count() or 1)
    loss_ref, out_ref, grads_ref = orc.loss_and_grads(kind, sd, x, y, dropout_masks=masks, update_stats=True)
    runs = []
    for _ in range(2):
        m = _build(kind).cuda().train()
        ctk.set_precision(m, "fp32")
        ctk.models.get_train_engine(m).forced_masks = tuple(t.cuda() for t in masks)
        out = m(x.cuda())
        loss = ctk.MSELoss()(out, y.cuda())
        loss.backward()
        torch.cuda.synchronize()
        runs.append((loss.item(), out.detach().cpu(), {k: p.grad.detach().cpu() for k, p in m.named_parameters()},
                     {k: v.detach().cpu() for k, v in m.state_dict().items()}))
    (l0, o0, g0, s0), (l1, o1, g1, s1) = runs
    assert l0 == l1 and all(torch.equal(g0[k], g1[k]) for k in g0)          # bit-reproducible
    rel_loss = abs(l0 - float(loss_ref)) / abs(float(loss_ref))
    print(kind, "fp32 path: loss gpu %.8f oracle %.8f rel %.2e; out max abs err %.2e" %
          (l0, float(loss_ref), rel_loss, (o0 - out_ref).abs().max().item()))
    assert rel_loss <= 1e-4
    num = den = 0.0
    worst, worst_name = 0.0, ""
    for name, r in grads_ref.items():
        g = g0[name]
        if r.norm().item() <= 1e-3 * max(1e-30, max(v.norm().item() for k, v in grads_ref.items() if k.endswith("weight"))) \
                and name.endswith("bias") and g.abs().max().item() == 0.0:
            continue                                           # conv bias in front of a train-mode BatchNorm: exactly zero
        if name.endswith("fc_layers.9.bias") or name.endswith("fc_layers.1.bias") or name.endswith("fc_layers.5.bias"):
            continue                # sums that cancel (last bias) or are exactly zero (Linear bias in front of BatchNorm1d)
        rel = rel_l2(g, r)
        num += ((g.double() - r.double()) ** 2).sum().item()
        den += (r.double() ** 2).sum().item()
        print(f"  {name:45s} rel L2 {rel:.3e}  |ref| {r.norm().item():.3e}")
        if rel > worst:
            worst, worst_name = rel, name
    whole = (num / den) ** 0.5
    print(kind, "fp32 path: whole-gradient rel L2 vs fp32 oracle %.3e; worst tensor %s %.3e" % (whole, worst_name, worst))
    assert whole <= WHOLE_GRAD_BOUND[kind], whole
    for k, v in sd.items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            np.testing.assert_allclose(s0[k].numpy(), v.numpy(), rtol=2e-5, atol=2e-6, err_msg=k)
        if k.endswith("num_batches_tracked"):
            assert int(s0[k]) == int(v), k


def test_fp32_short_adam_curve_matches_the_oracle():
    """Six reference steps (Adam lr 5e-4, wd 1e-4, fixed Dropout masks) on 16 tiles: per-step loss within 1 %."""
    import ctk
    n, steps = 16, 6
    x, y = orc.synthetic_batch(n, seed=99)
    model = _build("double")
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    masks = [orc.dropout_masks(n, 0.5, seed=100 + t) for t in range(steps)]
    torch.set_num_threads(os.cpu_count() or 1)
    tr = orc.OracleTrainer("double", sd, lr=5e-4, weight_decay=1e-4)
    ref = [tr.step(x, y, dropout_masks=masks[t])[0] for t in range(steps)]
    model = model.cuda().train()
    ctk.set_precision(model, "fp32")
    eng = ctk.models.get_train_engine(model)
    opt = ctk.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)
    crit = ctk.MSELoss()
    got = []
    for t in range(steps):
        eng.forced_masks = tuple(m.cuda() for m in masks[t])
        opt.zero_grad()
        loss = crit(model(x.cuda()), y.cuda())
        loss.backward()
        opt.step()
        got.append(loss.item())
    rel = [abs(a - b) / abs(b) for a, b in zip(got, ref)]
    print("fp32 path, 6 Adam steps: gpu", got, "oracle", ref, "rel", rel)
    assert max(rel) <= 1e-2


# ------------------------------------------------------------------------------------------------ 200-step loss curves
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _curve(name):
    import json
    path = os.path.join(GOLDEN, name)
    return json.load(open(path)) if os.path.exists(path) else None


def _run_curve(kind, g, precision, steps):
    """The reference loop of tests/golden/make_loss_curve.py (same tiles, order, initial weights, Dropout draws) on the GPU."""
    import ctk
    pool, batch = g["pool"], g["batch"]
    x, y = orc.synthetic_batch(pool, seed=g["data_seed"])
    model = _build(kind).cuda().train()
    ctk.set_precision(model, precision)
    eng = ctk.models.get_train_engine(model)
    opt = ctk.Adam(model.parameters(), lr=g["lr"], weight_decay=g["weight_decay"])
    crit = ctk.MSELoss()
    xd, yd = x.cuda(), y.cuda()
    p = P_DROP[kind]
    losses = []
    for t in range(steps):
        s = (t * batch) % pool
        torch.manual_seed(g["seed0"] + t)                      # the two nn.Dropout draws of the reference forward
        m1 = (F.dropout(torch.ones(batch, 512), p, True) != 0).float().cuda()
        m2 = (F.dropout(torch.ones(batch, 128), p, True) != 0).float().cuda()
        eng.forced_masks = (m1, m2)
        opt.zero_grad()
        loss = crit(model(xd[s:s + batch].contiguous()), yd[s:s + batch].contiguous())
        loss.backward()
        opt.step()
        losses.append(loss.item())
    return np.array(losses)


# 25-step geometric-mean windows, gpu / reference.  Measured on B200: double-branch 0.990 - 1.002 (fp32 path) and 0.975 - 0.987
# (bf16 path); the single-branch trajectory is chaotic from step 2 on (its loss jumps from 0.16 to 1.43 at step 1, and the
# reference's own second run differs from its first by 25 % at single steps), so its floor is wider.
WINDOW_BAND = {"double": {"fp32": 1.03, "bf16": 1.08}, "single": {"fp32": 1.10, "bf16": 1.25}}


@pytest.mark.parametrize("kind", ["double", "single"])
def test_200_step_loss_curve_at_batch_64(kind):
    """north_star: "the loss curve over the first 200 steps matching within 1 % relative".  The reference's own loop
    (unmodified model class + torch.optim.Adam + MSELoss on the CPU, tests/golden/make_loss_curve.py) at batch 64 from a
    256-tile pool is the golden; the same loop re-run with another thread count (only ATen's reduction order changes) tells
    for how many steps the reference reproduces ITSELF to 1 %.  Asserted:
      fp32 path : every step within 1 % until the reference's two runs first differ by 1e-4; over the compared steps (100) the
                  distance to the reference run at most 3 x the distance between the reference's own two runs (median and
                  maximum); every 25-step window's geometric-mean loss within WINDOW_BAND (3 % double-branch, 10 % single) or
                  3 x the largest distance the reference has kept to itself in any window so far, whichever is wider;
      bf16 path : step 0 within 1 %, every window within 8 % (double) / 25 % (single) or that same reference band --
                  bf16 operand rounding is a 2^-9 perturbation where a thread count is a 2^-24 one."""
    g = _curve(f"loss_curve_{kind}_b64.json")
    if g is None:
        pytest.skip("golden curve not generated")
    ref = np.array(g["reference_fp32"])
    steps = len(ref)
    alt = _curve(f"loss_curve_{kind}_b64_t4.json")
    other = np.array(alt["reference_fp32"]) if alt is not None else None
    agree = 2                       # without the second reference run only the first two steps are held to 1 %
    if other is not None:
        m = min(len(other), steps)
        bad = np.nonzero(np.abs(other[:m] - ref[:m]) / ref[:m] > 1e-2)[0]
        agree = max(2, int(bad[0]) if len(bad) else m)
    print(f"{kind}: golden has {steps} steps; the reference's two runs agree to 1 % for {agree} steps")
    gm = lambda v, a: float(np.exp(np.log(v[a:a + 25]).mean()))         # noqa: E731
    all_steps, all_ref = steps, ref
    for precision in ("fp32", "bf16"):
        # the fp32 path (CUDA cores, ~0.4 s per batch-64 step) is run for the first 100 steps, the bf16 path for all of them
        steps = min(all_steps, 100) if precision == "fp32" else all_steps
        ref = all_ref[:steps]
        gpu = _run_curve(kind, g, precision, steps)
        assert np.isfinite(gpu).all()
        rel = np.abs(gpu - ref) / ref
        first_bad = int(np.nonzero(rel > 1e-2)[0][0]) if (rel > 1e-2).any() else steps
        print(f"{kind} {precision}: step-0 loss gpu {gpu[0]:.6f} reference {ref[0]:.6f}; within 1 % for the first {first_bad} steps; "
              f"median rel {np.median(rel):.2e}, max {rel.max():.2e}")
        assert rel[0] <= 1e-2
        if other is not None:
            m = min(len(other), steps)
            own_rel = np.abs(other[:m] - ref[:m]) / ref[:m]
            print("   step: gpu-vs-reference / reference-vs-itself  " +
                  "  ".join(f"{t}: {rel[t]:.1e}/{own_rel[t]:.1e}" for t in range(min(m, 12))))
            print(f"   first {m} steps: median {np.median(rel[:m]):.2e} / {np.median(own_rel):.2e}, max {rel[:m].max():.2e} / {own_rel.max():.2e}")
            if precision == "fp32":
                # Both the fp32 path and the reference's second run are fp32 summation-order perturbations of the same
                # trajectory, of different size: another thread count reorders the threaded reductions only (step-1 loss
                # moves by 1e-6), the CUDA path reorders every sum and keeps fp64 partials (gradient 2.6e-4 away in relative
                # L2, which Adam's first sign-like update turns into 5e-5 on the step-1 loss) -- about two of the trajectory's
                # x10 - x30 per-step amplifications ahead.  So: while the reference reproduces itself to 1e-4 the GPU must be
                # within 1 %; once the trajectory has gone chaotic, its distance from the reference run may not exceed the
                # reference's own (2 x the median, 3 x the maximum over the compared steps).
                prefix = int(np.argmax(own_rel > 1e-4)) if (own_rel > 1e-4).any() else m      # steps before the reference first strays
                assert (rel[:prefix] <= 1e-2).all(), (prefix, rel[:prefix], own_rel[:prefix])
                assert np.median(rel[:m]) <= 3.0 * np.median(own_rel) + 1e-2, (np.median(rel[:m]), np.median(own_rel))
                assert rel[:m].max() <= 3.0 * own_rel.max() + 1e-2, (rel[:m].max(), own_rel.max())
        elif precision == "fp32":
            assert first_bad >= min(agree, steps), (first_bad, agree)
        own_max = 1.0
        for a in range(0, steps - 24, 25):
            r_, g_ = gm(ref, a), gm(gpu, a)
            own = max(gm(other, a) / r_, r_ / gm(other, a)) if other is not None and a + 25 <= len(other) else 1.0
            # The distance between the reference's two runs is ONE draw of the spread two fp32 realisations of this chaotic
            # trajectory have in that window (it grows with the step count: 0.3 / 1.0 / 2.6 / 6.8 % for the double-branch
            # model); the GPU run is another draw.  Two draws of the same spread differ by a factor of three or more one time
            # in five, so the band is three times the largest distance seen so far -- windows past the end of the second
            # reference run keep the last one.  (Measured over three kernel revisions of round 2: the fp32 path's window
            # 75 - 99 came out at 0.925, 0.925 and 0.864 of the reference's, the bf16 path's at 1.070.)
            own_max = max(own_max, own)
            band = max(WINDOW_BAND[kind][precision], 1.0 + 3.0 * (own_max - 1.0))
            print(f"   window {a:3d}: reference {r_:.5f} gpu {g_:.5f} ratio {g_ / r_:.4f} (reference vs itself {own:.4f}, band {band:.3f})")
            assert 1.0 / band <= g_ / r_ <= band, (precision, a, g_, r_, band)
        assert gpu[-10:].mean() < 0.5 * gpu[0]                  # and it trained
