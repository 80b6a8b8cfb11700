"""Pins oracle/crosstalk_oracle.py to what the unmodified reference computed (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

import crosstalk_oracle as orc


def _inputs(golden, n=4):
    tiles = golden["tiles"]
    xn = np.stack([np.stack([orc.normalize_image(t[0]), orc.normalize_image(t[1])]) for t in tiles])
    return torch.from_numpy(xn[:n]), torch.from_numpy(golden["labels_np"][:n])[:, None], xn


def _close_stats(sd, stats, rtol):
    for k, (s, a) in stats.items():
        v = sd[k].double()
        assert float(v.abs().sum()) == pytest.approx(a, rel=rtol, abs=1e-9), k
        assert float(v.sum()) == pytest.approx(s, rel=rtol, abs=max(1e-9, rtol * a)), k


@pytest.mark.parametrize("kind", ["single", "double"])
def test_init_matches_reference_constructor(golden, kind):
    sd = orc.INIT[kind](seed=0)
    assert set(sd.keys()) == set(golden[kind]["init_stats"].keys())
    assert len(sd) == (58 if kind == "single" else 72)
    _close_stats(sd, golden[kind]["init_stats"], 1e-6)


@pytest.mark.parametrize("kind", ["single", "double"])
def test_eval_forward_matches_reference(golden, kind):
    x, _, _ = _inputs(golden)
    sd = orc.INIT[kind](seed=0)
    with torch.no_grad():
        out = orc.FORWARD[kind](sd, x).flatten().numpy()
        out_r = orc.FORWARD[kind](orc.randomize_bn(sd, seed=7), x).flatten().numpy()
    np.testing.assert_allclose(out, golden[kind]["eval_out"], atol=2e-6, rtol=0)
    # non-vacuous check: randomised BN gives outputs with real spread
    np.testing.assert_allclose(out_r, golden[kind]["eval_out_randbn"], atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("kind", ["single", "double"])
def test_calibrated_eval_forward_matches_reference(golden, kind):
    # BN running stats := batch stats of the five fixture tiles -> eval outputs with real spread
    _, _, xn = _inputs(golden)
    x5 = torch.from_numpy(xn)
    sd = orc.calibrate_bn(kind, orc.INIT[kind](seed=0), x5)
    with torch.no_grad():
        out = orc.FORWARD[kind](sd, x5).flatten().numpy()
    ref = np.array(golden[kind]["eval_out_calibrated"])
    assert ref.max() - ref.min() > 0.02
    np.testing.assert_allclose(out, ref, atol=2e-5, rtol=0)


@pytest.mark.parametrize("kind", ["single", "double"])
def test_train_two_steps_match_reference(golden, kind):
    x, y, _ = _inputs(golden)
    sd = orc.INIT[kind](seed=0)
    tr = orc.OracleTrainer(kind, sd, lr=5e-4, weight_decay=1e-4)
    loss0, out0, grads = orc.loss_and_grads(kind, {k: v.clone() for k, v in sd.items()}, x, y, update_stats=False)
    np.testing.assert_allclose(out0.flatten().numpy(), golden[kind]["train_out"], atol=5e-6, rtol=0)
    for k, gn in golden[kind]["grad_norms"].items():
        assert float(grads[k].double().norm()) == pytest.approx(gn, rel=2e-3, abs=1e-9), k
    losses = [tr.step(x, y)[0] for _ in range(2)]
    np.testing.assert_allclose(losses, golden[kind]["train_losses"], rtol=1e-3)
    _close_stats(sd, golden[kind]["after2_stats"], 2e-4)


def test_pearson_matches_scipy(golden):
    _, _, xn = _inputs(golden, 5)
    for i in range(5):
        r32 = orc.pearson_f32(xn[i, 0], xn[i, 1])
        r64 = orc.pearson_f64(xn[i, 0], xn[i, 1])
        assert abs(r32 - golden["pearson_scipy_f32"][i]) <= 1e-6
        assert abs(r64 - golden["pearson_scipy_f64"][i]) <= 1e-9
        assert abs(r64 - golden["pearson_scipy_f32"][i]) <= 1e-6     # north_star tolerance


def test_pearson_constant_plane_is_nan():
    a = np.full((256, 256), 0.1, dtype=np.float32)
    b = np.random.default_rng(0).random((256, 256), dtype=np.float32)
    assert np.isnan(orc.pearson_f32(a, b)) and np.isnan(orc.pearson_f64(b, a))


def test_adam_step_matches_torch_optim():
    torch.manual_seed(3)
    p0 = torch.randn(1000)
    g = torch.randn(1000)
    p_ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([p_ref], lr=5e-4, weight_decay=1e-4)
    p, m, v = p0.clone(), torch.zeros(1000), torch.zeros(1000)
    for t in range(1, 4):
        p_ref.grad = g.clone() * t
        opt.step()
        orc.adam_step(p, g * t, m, v, t, 5e-4)
    np.testing.assert_allclose(p.numpy(), p_ref.detach().numpy(), rtol=0, atol=1e-7)


def test_metrics_restatement_matches_reference_calls(golden):
    """RMSE / 256-bin histogram / histogram correlation restatements against the reference's own numpy + scipy calls
    (tests/golden/metrics.json, written by make_metrics_golden.py with the expressions of test-cross-talk-model.py:65-79)
    and against np.histogram directly on awkward inputs."""
    import json
    import os
    m = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "metrics.json")))
    tiles = golden["tiles"].astype(np.float32)
    xn = np.stack([np.stack([orc.normalize_image(t[0]), orc.normalize_image(t[1])]) for t in tiles])
    for name, imgs in (("normalised", xn), ("raw", tiles)):
        g = m[name]
        for j in range(imgs.shape[0]):
            h0, h1 = orc.histogram256_f32(imgs[j, 0]), orc.histogram256_f32(imgs[j, 1])
            assert [int(v) for v in h0[:16]] == g["hist0_head"][j] and [int(v) for v in h1[:16]] == g["hist1_head"][j]
            assert int((h0 ** 2).sum()) == g["hist0_sum_sq"][j] and int((h1 ** 2).sum()) == g["hist1_sum_sq"][j]
            assert abs(orc.hist_correlation(imgs[j, 0], imgs[j, 1]) - g["hist_corr"][j]) <= 1e-12
            assert abs(orc.rmse_f32(imgs[j, 0], imgs[j, 1]) - g["rmse"][j]) <= 1e-7 * g["rmse"][j]
            assert abs(orc.nmi_digitized(imgs[j, 0], imgs[j, 1]) - g["nmi"][j]) <= 1e-13
    rng = np.random.default_rng(0)
    awkward = [np.full(4096, 0.3, np.float32), (rng.standard_normal(65536) * 1e-3 + 5).astype(np.float32),
               (rng.random(65536) ** 3 * 7 - 2).astype(np.float32), np.arange(257, dtype=np.float32) / 256,
               np.array([0, 1, 1, 1, 0.5, 0.00390625, 0.99609375] * 8, np.float32)]
    for a in awkward:
        assert np.array_equal(orc.histogram256_f32(a), np.histogram(a, bins=256)[0])
    assert np.isnan(orc.hist_correlation(np.arange(256, dtype=np.float32), np.arange(256, dtype=np.float32)))   # flat histograms
    for a in awkward[1:]:
        assert np.array_equal(orc.digitize256_f32(a), np.digitize(a, bins=np.linspace(a.min(), a.max(), 256)))
    const = np.full(1024, 0.5, np.float32)
    assert orc.nmi_digitized(const, const) == 1.0 and orc.nmi_digitized(const, awkward[1][:1024]) == 0.0


def _ssim_device_model(a, b):
    """NumPy model of the arithmetic ctk_tile_ssim_f32 performs (csrc/ssim.cu): direct 7-tap float64 sums, float32 rounding
    after each 1-D pass, the SSIM map in float32 in the library's operation order, float64 mean."""
    a = a.astype(np.float32)
    b = b.astype(np.float32)
    rng = np.float32(max(a.max(), b.max())) - np.float32(min(a.min(), b.min()))

    def pass1d(img, axis):
        v = np.lib.stride_tricks.sliding_window_view(img.astype(np.float64), 7, axis=axis)
        s = np.zeros(v.shape[:-1])
        for k in range(7):
            s = s + v[..., k]
        return (s / 7.0).astype(np.float32)

    def box(img):
        return pass1d(pass1d(img, 0), 1)

    ux, uy, uxx, uyy, uxy = box(a), box(b), box(a * a), box(b * b), box(a * b)
    cn = np.float32(49 / 48)
    vx, vy, vxy = cn * (uxx - ux * ux), cn * (uyy - uy * uy), cn * (uxy - ux * uy)
    t1, t2 = np.float32(0.01) * rng, np.float32(0.03) * rng
    c1, c2 = t1 * t1, t2 * t2
    with np.errstate(all="ignore"):
        smap = (((np.float32(2) * ux) * uy + c1) * (np.float32(2) * vxy + c2)) / (((ux * ux + uy * uy) + c1) * ((vx + vy) + c2))
    assert smap.dtype == np.float32
    return float(smap.astype(np.float64).sum() / smap.size)


def test_ssim_restatement(golden):
    """SSIM (test-cross-talk-model.py:80-82).  scikit-image is not installed here, so the restatement cannot be pinned to
    a reference output ("parity unpinned", oracle header); what can be checked on the CPU: known answers (identical
    planes -> exactly 1, symmetry), agreement with an independent float64 evaluation of the definition up to float32 noise,
    and that the operation-order model the CUDA kernel implements reproduces the scipy-based restatement."""
    tiles = golden["tiles"].astype(np.float32)
    cases = [(orc.normalize_image(t[0]), orc.normalize_image(t[1])) for t in tiles[:3]] + [(t[0], t[1]) for t in tiles[:3]]
    xs, _ = orc.synthetic_batch(2, seed=5)
    xs = xs.numpy()
    cases += [(xs[0, 0], xs[0, 1]), (xs[1, 0], np.full_like(xs[1, 1], 0.25)), (np.round(xs[1, 0] * 6) / 6, np.round(xs[1, 1] * 6) / 6)]
    r = np.random.default_rng(3)
    cases += [(r.random((40, 72), dtype=np.float32), r.random((40, 72), dtype=np.float32)),
              (r.random((7, 8), dtype=np.float32), r.random((7, 8), dtype=np.float32))]
    for a, b in cases:
        a, b = np.ascontiguousarray(a, dtype=np.float32), np.ascontiguousarray(b, dtype=np.float32)
        v = orc.ssim_f32(a, b)
        assert -1.0 <= v <= 1.0
        assert v == orc.ssim_f32(b, a)                                      # symmetric, bit for bit
        assert orc.ssim_f32(a, a) == 1.0                                    # identical planes
        scale = max(1.0, float(max(a.max(), b.max()) - min(a.min(), b.min())))
        assert abs(v - orc.ssim_f64(a, b)) <= (1e-6 if scale == 1.0 else 1e-3), (v, orc.ssim_f64(a, b))
        assert abs(v - _ssim_device_model(a, b)) <= 1e-12, (v, _ssim_device_model(a, b))
    with np.errstate(all="ignore"):
        both = np.full((16, 16), 0.5, dtype=np.float32)
        assert np.isnan(orc.ssim_f32(both, both.copy())) and np.isnan(_ssim_device_model(both, both.copy()))


def _adversarial_planes():
    """float32 planes that stress the bin rules: values exactly on edges, tiny and huge ranges, negative offsets, denormal
    steps, two-valued planes, heavy ties."""
    r = np.random.default_rng(11)
    out = [r.random(4096, dtype=np.float32),
           (r.random(4096, dtype=np.float32) * np.float32(1e-30)),
           (r.random(4096, dtype=np.float32) * np.float32(3e4) - np.float32(1.5e4)),
           np.float32(1000.0) + r.random(4096, dtype=np.float32) * np.float32(1e-3),
           (r.integers(0, 256, 4096) / np.float32(255)).astype(np.float32),          # every value an exact linspace edge
           (r.integers(0, 257, 4096) / np.float32(256)).astype(np.float32),          # every value an exact histogram edge
           (r.integers(0, 7, 4096) / np.float32(6)).astype(np.float32),
           np.where(r.random(4096) < 0.5, np.float32(-2.5), np.float32(7.25)).astype(np.float32),
           np.linspace(-1, 1, 4096, dtype=np.float32) ** 3,
           np.nextafter(np.float32(0.5), np.float32(1), dtype=np.float32) * np.ones(4096, np.float32) + (r.integers(0, 3, 4096) * np.float32(6e-8)).astype(np.float32)]
    for k in range(6):                                                                # random affine maps of random edges
        lo, span = np.float32(r.normal() * 10), np.float32(abs(r.normal()) * 10.0 ** int(r.integers(-3, 4)))
        out.append((lo + span * (r.integers(0, 256, 4096) / np.float32(255)).astype(np.float32)).astype(np.float32))
    return out


def test_bin_rules_match_numpy_on_adversarial_planes():
    """The histogram and digitisation restatements (which the CUDA kernels follow operation by operation) against NumPy
    itself on planes built to sit on bin edges (test-cross-talk-model.py:65-66, 71-74)."""
    for x in _adversarial_planes():
        x = np.ascontiguousarray(x, dtype=np.float32)
        try:
            want = np.histogram(x, bins=256)[0]
        except ValueError:               # range too narrow for 256 float32 bins: the reference's loop raises here too
            want = None
        if want is not None:
            assert np.array_equal(orc.histogram256_f32(x), want)
        assert np.array_equal(orc.digitize256_f32(x), np.digitize(x, bins=np.linspace(x.min(), x.max(), 256)))


def test_oracle_on_32_reference_fixture_pairs(golden32):
    """The wider fixture set (32 Training_Data pairs): Pearson r (float32, like scipy), RMSE, histogram correlation and the
    eval-mode scores of both models (randomised BatchNorm) against the unmodified reference's values."""
    x = golden32["x"]
    xn = x.numpy()
    r = orc.pearson_batch(x, f64=False)
    np.testing.assert_allclose(r, np.array(golden32["pearson_f32"]), rtol=0, atol=1e-6)
    np.testing.assert_allclose(orc.pearson_batch(x, f64=True), np.array(golden32["pearson_f32"]), rtol=0, atol=1e-6)
    for j in range(x.shape[0]):
        assert abs(orc.rmse_f32(xn[j, 0], xn[j, 1]) - golden32["rmse"][j]) <= 1e-7
        assert abs(orc.hist_correlation(xn[j, 0], xn[j, 1]) - golden32["hist_corr"][j]) <= 1e-9
    for kind in ("single", "double"):
        sd = orc.randomize_bn(orc.INIT[kind](0), seed=7)
        with torch.no_grad():
            out = orc.FORWARD[kind](sd, x).flatten().numpy()
        np.testing.assert_allclose(out, np.array(golden32[f"{kind}_eval_randomized_bn"]), rtol=0, atol=2e-6)
