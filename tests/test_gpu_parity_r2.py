"""Round-2 parity tests: the north_star's tolerances asserted on well-conditioned problems, and the properties that make
that possible (bit-reproducible training steps, caches that follow raw parameter writes).

* every cross-CTA reduction is a fixed-order two-stage sum, so two identical training steps give BIT-IDENTICAL losses,
  gradients and BatchNorm buffers;
* gradients at batch 64 (tiles from a 256-tile pool) are compared with the fp32 oracle DIRECTLY -- whole-gradient and
  per-tensor relative L2, no bf16-emulation yardstick;
* "CPU-warmed" weights (SURVEY 8c fallback (ii): seed-0 init + 20 oracle Adam steps at batch 16) give eval outputs with
  real spread; on them both models meet the north_star bounds: 1e-3 absolute for bf16 operands, 1e-5 for the fp32-class path;
* the exact bench configuration (256-tile batch, 64-tile HostScorer slices) is spot-checked against the oracle.
"""
import os

import numpy as np
import pytest
import torch

import crosstalk_oracle as orc

pytestmark = pytest.mark.gpu

P_DROP = {"single": 0.1, "double": 0.5}


def _build(kind):
    import ctk
    torch.manual_seed(0)
    if kind == "single":
        return ctk.AdvancedRegressionModel(initial_filters=128, num_conv_blocks=6)
    return ctk.SimplifiedTwoBranchRegressionModel(initial_filters_per_branch=64)


def _rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-300)).item()


# ------------------------------------------------------------------------------------------------ determinism
@pytest.mark.parametrize("kind", ["single", "double"])
def test_training_steps_are_bit_reproducible(kind):
    """Two fresh models, the same two steps: losses, every gradient, every parameter and buffer must be bit-identical.
    (Round 1 summed batch statistics, BN-backward sums and split-K weight gradients with floating-point atomics: two runs
    differed by 5-15 % in whole-gradient L2 on ill-conditioned problems.)"""
    import ctk
    x, y = orc.synthetic_batch(12, seed=31)
    masks = [tuple(m.cuda() for m in orc.dropout_masks(12, P_DROP[kind], seed=7 + t)) for t in range(2)]
    runs = []
    for _ in range(2):
        model = _build(kind).cuda().train()
        eng = ctk.models.get_train_engine(model)
        opt = ctk.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)
        crit = ctk.MSELoss()
        losses, grads = [], None
        for t in range(2):
            eng.forced_masks = masks[t]
            opt.zero_grad()
            loss = crit(model(x.cuda()), y.cuda())
            loss.backward()
            grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
            opt.step()
            losses.append(loss.item())
        torch.cuda.synchronize()
        runs.append((losses, grads, {k: v.detach().clone() for k, v in model.state_dict().items()}))
    (l0, g0, s0), (l1, g1, s1) = runs
    assert l0 == l1
    for k in g0:
        assert torch.equal(g0[k], g1[k]), k
    for k in s0:
        assert torch.equal(s0[k], s1[k]), k


def test_gradient_accumulation_two_forwards_before_backward():
    """Saved activations live on the autograd ctx: two train-mode forwards may be pending before one backward runs over
    both (micro-batch accumulation, two losses).  The accumulated gradient equals the sum of the separate ones exactly."""
    import ctk
    x, y = orc.synthetic_batch(8, seed=41)
    xa, ya, xb, yb = x[:4].cuda(), y[:4].cuda(), x[4:].cuda(), y[4:].cuda()
    ma = tuple(m.cuda() for m in orc.dropout_masks(4, 0.5, seed=1))
    mb = tuple(m.cuda() for m in orc.dropout_masks(4, 0.5, seed=2))

    def grads_of(model):
        return {k: p.grad.detach().clone() for k, p in model.named_parameters()}

    sep = []
    for xs, ys, ms in ((xa, ya, ma), (xb, yb, mb)):
        model = _build("double").cuda().train()
        ctk.models.get_train_engine(model).forced_masks = ms
        ctk.MSELoss()(model(xs), ys).backward()
        sep.append(grads_of(model))
    model = _build("double").cuda().train()
    eng = ctk.models.get_train_engine(model)
    eng.forced_masks = ma
    out_a = model(xa)
    eng.forced_masks = mb
    out_b = model(xb)                       # second forward before the first backward
    with torch.no_grad():
        model(xa)                           # and a train-mode forward under no_grad in between: must not disturb either
    (ctk.MSELoss()(out_a, ya) + ctk.MSELoss()(out_b, yb)).backward()
    acc = grads_of(model)
    for k in acc:
        assert torch.equal(acc[k], sep[0][k] + sep[1][k]) or torch.equal(acc[k], sep[1][k] + sep[0][k]), k
    with pytest.raises(Exception):
        # the graph of out_a has been consumed (like autograd without retain_graph)
        ctk.MSELoss()(out_a, ya).backward()


def test_eval_after_ctk_adam_steps_uses_the_new_weights():
    """train() / eval() alternate every epoch in the reference loop (train_model.py:408-446).  ctk.Adam and the train-mode
    BatchNorm update write through raw pointers; the eval engine's derived cache (packed weights, folded BN) must notice."""
    import ctk
    x, y = orc.synthetic_batch(8, seed=51)
    xd, yd = x.cuda(), y.cuda()
    model = _build("double")
    model.load_state_dict(orc.randomize_bn(model.state_dict(), seed=7))
    model = model.cuda()
    opt = ctk.Adam(model.parameters(), lr=5e-3, weight_decay=1e-4)
    crit = ctk.MSELoss()
    model.eval()
    with torch.no_grad():
        before = model(xd).clone()                      # builds the cache
    model.train()
    for _ in range(3):
        opt.zero_grad()
        crit(model(xd), yd).backward()
        opt.step()
    model.eval()
    with torch.no_grad():
        after = model(xd).clone()
        ref = orc.double_forward({k: v.detach().cpu() for k, v in model.state_dict().items()}, x)
    torch.cuda.synchronize()
    assert (after - before).abs().max().item() > 1e-3           # the weights did move
    assert (after.cpu() - ref).abs().max().item() <= 1e-3        # and eval scores follow the CURRENT state_dict


# ------------------------------------------------------------------------------------------------ small device-side pieces
def test_philox_dropout_masks_and_mse_module():
    import ctk
    from ctypes import c_float, c_longlong, c_ulonglong
    from ctk import _lib as L
    n1, n2 = 256 * 512, 256 * 128 + 3
    m1 = torch.full((n1,), -1.0, device="cuda")
    m2 = torch.full((n2,), -1.0, device="cuda")

    def draw(a, b, seed, off):
        L.call("ctk_dropout_masks", L.ptr(a), c_longlong(a.numel()), c_float(0.5), L.ptr(b), c_longlong(b.numel()),
               c_float(0.1), c_ulonglong(seed), c_ulonglong(off), L.stream())
    draw(m1, m2, 1234, 0)
    a1, a2 = m1.clone(), m2.clone()
    draw(m1, m2, 1234, 0)
    assert torch.equal(a1, m1) and torch.equal(a2, m2)                              # reproducible
    assert set(a1.unique().tolist()) == {0.0, 1.0} and set(a2.unique().tolist()) == {0.0, 1.0}
    assert abs(a1.mean().item() - 0.5) < 5e-3 and abs(a2.mean().item() - 0.9) < 5e-3  # keep fractions 1 - p
    draw(m1, m2, 1234, 1)
    assert (a1 != m1).float().mean().item() > 0.4                                   # another step, another draw
    draw(m1, m2, 99, 0)
    assert (a1 != m1).float().mean().item() > 0.4                                   # another seed
    # the model draws its own masks through this kernel when none are forced: same seed -> same step
    x, y = orc.synthetic_batch(4, seed=3)
    losses = []
    for _ in range(2):
        torch.manual_seed(77)
        model = _build("double").cuda().train()
        loss = ctk.MSELoss()(model(x.cuda()), y.cuda())
        losses.append(loss.item())
    assert losses[0] == losses[1]
    # ctk.MSELoss == torch.nn.MSELoss, value and gradient, also when the loss is scaled before backward
    o = torch.randn(37, 1, device="cuda", requires_grad=True)
    t = torch.randn(37, 1, device="cuda")
    (ctk.MSELoss()(o, t) * 3.0).backward()
    g = o.grad.clone()
    o.grad = None
    ref = torch.nn.MSELoss()(o, t)
    (ref * 3.0).backward()
    assert abs(ctk.MSELoss()(o, t).item() - ref.item()) <= 1e-6 * abs(ref.item())
    assert torch.allclose(g, o.grad, rtol=1e-6, atol=1e-8)


# ------------------------------------------------------------------------------------------------ CPU-warmed weights
@pytest.fixture(scope="module")
def warmed():
    """SURVEY 8c fallback weight set (ii): seed-0 init + 20 oracle Adam steps (lr 5e-4, wd 1e-4) at batch 16 on a 64-tile
    pool -- BatchNorm running statistics and the head have moved off their initial values and eval outputs have real
    spread.  Built once per test session on the host (about a minute per model on 16 cores)."""
    cache = {}

    def get(kind):
        if kind not in cache:
            torch.set_num_threads(os.cpu_count() or 1)
            x, y = orc.synthetic_batch(64, seed=4321)
            tr = orc.OracleTrainer(kind, orc.INIT[kind](0), lr=5e-4, weight_decay=1e-4)
            for t in range(20):
                s = (t * 16) % 64
                tr.step(x[s:s + 16], y[s:s + 16], dropout_masks=orc.dropout_masks(16, P_DROP[kind], 1000 + t))
            cache[kind] = {k: v.detach().clone() for k, v in tr.sd.items()}
        return cache[kind]
    return get


# ------------------------------------------------------------------------------------------------ gradients at batch 64
# Whole-gradient relative L2 distance to the fp32 oracle with bf16 operand storage.  Conditioning of this problem, measured
# on the CPU in pure fp32 (double-branch): a 1e-6 relative perturbation of the INPUT moves the whole gradient by 6e-4 on the
# warmed weights and 9e-4 at seed-0 init (x600-900 through train-mode BatchNorm); the oracle with bf16-rounded operands is
# 3.4e-2 from fp32 on the warmed weights and 1.7e-1 at init.  The bounds are ~1.7x the bf16 rounding model on the warmed set.
# Measured on B200: whole gradient 2.9e-2 (double) / 1.3e-2 (single); single tensors up to 19 % / 27 % (the first conv blocks,
# whose gradient norms are 1e-3 of the whole).
WHOLE_GRAD_BOUND = {"single": 3e-2, "double": 5e-2}
TENSOR_GRAD_BOUND = {"single": 4e-1, "double": 3e-1}


@pytest.mark.parametrize("kind", ["double", "single"])
def test_gradients_at_batch_64_against_fp32_oracle(warmed, kind):
    """One training step from the CPU-warmed weights on 64 tiles of a 256-tile pool: loss within 1 %, whole-gradient and
    per-tensor relative L2 against the fp32 oracle asserted directly (no bf16-emulation yardstick), BatchNorm running
    statistics within 2e-3 relative / 1e-3 absolute."""
    import ctk
    pool_x, pool_y = orc.synthetic_batch(256, seed=4321)
    idx = torch.arange(0, 256, 4)
    x, y = pool_x[idx].contiguous(), pool_y[idx].contiguous()
    n = x.shape[0]
    model = _build(kind)
    model.load_state_dict(warmed(kind))
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    masks = orc.dropout_masks(n, P_DROP[kind], seed=5)
    torch.set_num_threads(os.cpu_count() or 1)
    loss_ref, out_ref, grads_ref = orc.loss_and_grads(kind, sd, x, y, dropout_masks=masks, update_stats=True)
    model = model.cuda().train()
    ctk.models.get_train_engine(model).forced_masks = tuple(m.cuda() for m in masks)
    out = model(x.cuda())
    loss = ctk.MSELoss()(out, y.cuda())
    loss.backward()
    torch.cuda.synchronize()
    rel_loss = abs(loss.item() - float(loss_ref)) / abs(float(loss_ref))
    print(kind, "batch 64: loss gpu %.6f oracle %.6f rel %.2e; out max abs err %.2e" %
          (loss.item(), float(loss_ref), rel_loss, (out.detach().cpu() - out_ref).abs().max().item()))
    assert rel_loss <= 1e-2
    num = den = 0.0
    worst, worst_name = 0.0, ""
    for name, p in model.named_parameters():
        g, r = p.grad.detach().cpu(), grads_ref[name]
        if name.endswith("bias") and p.dim() == 1 and (".conv_blocks." in name or name.startswith("conv_layers.")) and \
                int(name.split(".")[-2]) % 4 == 0:
            # conv bias in front of a train-mode BatchNorm: the exact gradient is 0 (autograd leaves fp32 cancellation noise)
            assert g.abs().max().item() == 0.0, name           # (the oracle's value is what cancels to in fp32, not a signal)
            continue
        if name.endswith("fc_layers.9.bias"):
            scale = (2.0 * (out_ref - y).abs() / n).sum().item()      # a sum that cancels: bound by the size of its terms
            assert (g - r).abs().item() <= 2e-2 * scale, (name, g.item(), r.item(), scale)
            continue
        if name.endswith("fc_layers.1.bias") or name.endswith("fc_layers.5.bias"):
            # Linear bias in front of a train-mode BatchNorm1d: the exact gradient is 0, both sides hold cancellation noise
            assert g.abs().max().item() <= 1e-6 * max(1.0, grads_ref[name[:-4] + "weight"].norm().item()), name
            continue
        rel = _rel(g, r)
        num += ((g.double() - r.double()) ** 2).sum().item()
        den += (r.double() ** 2).sum().item()
        print(f"  {name:45s} rel L2 {rel:.3e}  |ref| {r.norm().item():.3e}")
        if rel > worst:
            worst, worst_name = rel, name
    whole = (num / den) ** 0.5
    print(kind, "batch 64: whole-gradient rel L2 vs fp32 oracle %.3e; worst tensor %s %.3e" % (whole, worst_name, worst))
    assert whole <= WHOLE_GRAD_BOUND[kind], whole
    assert worst <= TENSOR_GRAD_BOUND[kind], (worst_name, worst)
    msd = model.state_dict()
    for k, v in sd.items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            # bf16 storage of the block inputs moves a batch mean by a few 1e-4 absolute (measured: up to 3.1e-4)
            np.testing.assert_allclose(msd[k].cpu().numpy(), v.numpy(), rtol=2e-3, atol=1e-3, err_msg=k)
        if k.endswith("num_batches_tracked"):
            assert int(msd[k]) == int(v), k


@pytest.mark.parametrize("kind", ["single", "double"])
def test_cpu_warmed_weights_meet_the_north_star_tolerances(warmed, kind):
    """Predicted crosstalk score within 1e-3 absolute (bf16 operands) and 1e-5 (fp32-class path) of the fp32 CPU oracle."""
    import ctk
    sd = warmed(kind)
    x, _ = orc.synthetic_batch(16, seed=99)
    with torch.no_grad():
        ref = orc.FORWARD[kind](sd, x).flatten()
    model = _build(kind)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    with torch.no_grad():
        got_bf16 = model(x.cuda()).flatten().cpu()
        ctk.set_precision(model, "fp32")
        got_fp32 = model(x.cuda()).flatten().cpu()
    e16, e32 = (got_bf16 - ref).abs().max().item(), (got_fp32 - ref).abs().max().item()
    print(kind, "CPU-warmed weights: output spread %.3e (std %.3e); max abs err bf16 %.2e, fp32-class %.2e" %
          ((ref.max() - ref.min()).item(), ref.std().item(), e16, e32))
    assert (ref.max() - ref.min()).item() > 1e-3          # not the vacuous near-constant random-init output
    assert e16 <= 1e-3
    assert e32 <= 1e-5


# ------------------------------------------------------------------------------------------------ the bench configuration
def test_bench_configuration_spot_check():
    """The exact inference path bench.py times -- double-branch model, randomised BatchNorm, ONE 256-tile batch through
    engine.forward, and the same batch through HostScorer's 64-tile slices from pinned host memory -- with 16 sampled
    tiles against the oracle (1e-3 on the score, 1e-9 on Pearson r)."""
    import ctk
    from ctk import synthetic
    base, _ = synthetic.synthetic_batch(32, seed=1234)
    host = base.repeat(8, 1, 1, 1).contiguous().pin_memory()
    model = _build("double")
    sd = synthetic.randomize_bn(model.state_dict(), seed=7)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    xd = host.cuda()
    with torch.no_grad():
        scores = ctk.models.get_engine(model).forward(xd).flatten().cpu()
        r = ctk.pearson_per_image(xd).cpu()
        scorer = ctk.HostScorer(model, slice_tiles=64, device="cuda")
        s_host, r_host = scorer.score(host)
    idx = torch.arange(3, 256, 16)
    with torch.no_grad():
        ref = orc.double_forward(sd, host[idx]).flatten()
    r_ref = torch.from_numpy(orc.pearson_batch(host[idx]))
    assert (scores[idx] - ref).abs().max().item() <= 1e-3
    assert (r[idx] - r_ref).abs().max().item() <= 1e-9
    # 64-tile slices pad FC1's M to 128 and pick another split-K factor than the 256-tile batch: same scores up to fp32
    # summation order; Pearson is per tile and identical
    assert (s_host.flatten() - scores).abs().max().item() <= 1e-5
    assert torch.equal(r_host.flatten().double(), r.double())
    assert (s_host.flatten()[idx] - ref).abs().max().item() <= 1e-3


def test_32_reference_fixture_pairs_against_the_reference_outputs(golden32):
    """Config 1 of BASELINE.json widened to 32 of the reference's own Training_Data pairs: scores of both models in both
    precisions and the comparison metrics against the values the UNMODIFIED reference computed (tests/golden/golden32.json)."""
    import ctk
    x = golden32["x"].cuda()
    m = ctk.tile_metrics(x)
    np.testing.assert_allclose(m["pearson"].cpu().numpy(), np.array(golden32["pearson_f32"]), rtol=0, atol=1e-6)
    np.testing.assert_allclose(ctk.pearson_per_image(x).cpu().numpy(), np.array(golden32["pearson_f32"]), rtol=0, atol=1e-6)
    np.testing.assert_allclose(m["rmse"].cpu().numpy(), np.array(golden32["rmse"]), rtol=0, atol=1e-6)
    np.testing.assert_allclose(m["hist_corr"].cpu().numpy(), np.array(golden32["hist_corr"]), rtol=0, atol=1e-9)
    for kind in ("single", "double"):
        model = _build(kind)
        model.load_state_dict(orc.randomize_bn(model.state_dict(), seed=7))
        model = model.cuda().eval()
        ref = torch.tensor(golden32[f"{kind}_eval_randomized_bn"])
        with torch.no_grad():
            e16 = (model(x).flatten().cpu() - ref).abs().max().item()
            ctk.set_precision(model, "fp32")
            e32 = (model(x).flatten().cpu() - ref).abs().max().item()
        print(kind, "32 fixture pairs: max abs err vs the reference's scores: bf16 %.2e, fp32-class %.2e" % (e16, e32))
        assert e16 <= 1e-3 and e32 <= 1e-5


def test_edge_cases_empty_ragged_and_single_sample_batches():
    """Empty and ragged batches through the eval path (FC1 pads its M to 128, batches above 256 tiles run in slices), the
    smallest training batch, and the reference's own refusal of a 1-sample training batch (nn.BatchNorm1d raises ValueError)."""
    import ctk
    x, y = orc.synthetic_batch(5, seed=8)
    model = _build("double")
    sd = orc.randomize_bn(model.state_dict(), seed=7)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    with torch.no_grad():
        ref = orc.double_forward(sd, x).flatten()
        assert tuple(model(x[:0].cuda()).shape) == (0, 1)
        assert ctk.pearson_per_image(x[:0].cuda()).numel() == 0
        for n in (1, 3, 5):
            assert (model(x[:n].cuda()).flatten().cpu() - ref[:n]).abs().max().item() <= 1e-3
        big = x.repeat(52, 1, 1, 1)[:257].contiguous()                 # 256 + 1: two slices
        out = model(big.cuda()).flatten().cpu()
        assert (out - ref.repeat(52)[:257]).abs().max().item() <= 1e-3
    with pytest.raises(ctk.CtkError):
        model(x.cuda().permute(0, 1, 3, 2))                            # non-contiguous input
    with pytest.raises(ctk.CtkError):
        model(x[:, :1].contiguous().cuda())                            # one channel
    model.train()
    for precision in ("bf16", "fp32"):
        ctk.set_precision(model, precision)
        loss = ctk.MSELoss()(model(x[:2].cuda()), y[:2].cuda())        # the smallest batch BatchNorm accepts
        loss.backward()
        assert np.isfinite(loss.item()) and all(torch.isfinite(p.grad).all() for p in model.parameters())
        with pytest.raises(ValueError):
            model(x[:1].cuda())
    # an odd batch through the single-branch model (its deepest conv sees 3 x 8 x 8 = 192 pixels: a ragged last tile)
    single = _build("single").cuda().train()
    for precision in ("bf16", "fp32"):
        ctk.set_precision(single, precision)
        loss = ctk.MSELoss()(single(x[:3].cuda()), y[:3].cuda())
        loss.backward()
        assert np.isfinite(loss.item()) and all(torch.isfinite(p.grad).all() for p in single.parameters())
