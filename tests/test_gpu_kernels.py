"""GPU parity tests, kernel by kernel, through the C ABI (ctypes) against the CPU oracle.

Tolerances (written here, as the task asks):
  * Pearson r:               |r_gpu - r_oracle_f64| <= 1e-9, and <= 1e-6 against the float32 SciPy golden.
  * fp32 kernels (first conv -> bf16 store, head, MSE, Adam): a few float32 / bf16 ulps, stated per test.
  * bf16 tensor-core kernels: operands are rounded to bf16 by construction, accumulation is fp32, the
    output is rounded to bf16 once: |err| <= 2^-7 * |ref| + small absolute term.
"""
import math
from ctypes import c_float, c_int

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import crosstalk_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctk():
    import ctk as _ctk
    _ctk.load()
    return _ctk


def _bf16_close(out, ref, rel=2 ** -7, abs_=2e-2):
    err = (out.float().cpu() - ref).abs()
    bound = rel * ref.abs() + abs_
    bad = (err > bound).float().mean().item()
    assert bad == 0.0, f"max err {err.max().item():.4g}, frac over bound {bad:.4g}"


# ------------------------------------------------------------------ Pearson
def test_pearson_synthetic_and_golden(ctk, golden):
    x, _ = orc.synthetic_batch(16, seed=1234)
    r = ctk.pearson_per_image(x.cuda()).cpu().numpy()
    ref = orc.pearson_batch(x, f64=True)
    assert np.abs(r - ref).max() <= 1e-9
    tiles = golden["tiles"]
    xn = np.stack([np.stack([orc.normalize_image(t[0]), orc.normalize_image(t[1])]) for t in tiles])
    r = ctk.pearson_per_image(torch.from_numpy(xn).cuda()).cpu().numpy()
    assert np.abs(r - np.array(golden["pearson_scipy_f32"])).max() <= 1e-6      # north_star tolerance
    assert np.abs(r - np.array(golden["pearson_scipy_f64"])).max() <= 1e-9


def test_pearson_edge_cases(ctk):
    x = torch.rand(4, 2, 64, 64)
    x[1, 0] = 0.25                      # constant plane -> NaN (test-cross-talk-model.py:61-62)
    x[2, 1] = x[2, 0]                   # identical planes -> exactly 1 after clipping
    x[3, 1] = 1.0 - x[3, 0]             # anti-correlated -> -1
    r = ctk.pearson_per_image(x.cuda()).cpu().numpy()
    assert math.isnan(r[1])
    assert r[2] == pytest.approx(1.0, abs=1e-12) and r[2] <= 1.0
    assert r[3] == pytest.approx(-1.0, abs=1e-6) and r[3] >= -1.0
    assert abs(r[0] - orc.pearson_f64(x[0, 0].numpy(), x[0, 1].numpy())) <= 1e-9
    assert ctk.pearson_per_image(torch.empty(0, 2, 64, 64).cuda()).numel() == 0


def test_pearson_affine_invariance_full_size(ctk):
    # size-independent property at the bench size: r(a*x+b, c*y+d) == sign(a*c) * r(x, y)
    x, _ = orc.synthetic_batch(256, seed=5)
    xd = x.cuda()
    r0 = ctk.pearson_per_image(xd)
    y = xd.clone()
    y[:, 0] = 3.0 * y[:, 0] + 0.5
    y[:, 1] = -0.25 * y[:, 1] + 2.0
    r1 = ctk.pearson_per_image(y)
    assert (r0 + r1).abs().max().item() <= 5e-7


# ------------------------------------------------------------------ first conv block
def test_prepare_tiles_bit_exact(ctk, golden):
    """Device input pipeline (float64 / float32 payload -> cast -> per-plane min-max normalise -> flips) against the
    oracle restatement of train_model.py:166-167, 211-232: float32 arithmetic, so equality is exact.  Covers the
    256 x 256 cluster kernel, the generic kernel (ragged 40 x 24), constant planes, all four flip states, empty batches."""
    rng = np.random.default_rng(1)
    tiles = golden["tiles"].astype(np.float64)
    synth = rng.random((4, 2, 256, 256)) * 3.7 - 1.2
    synth[1, 0] = 0.625                                                    # constant plane: left unchanged
    raw = np.concatenate([tiles, synth], axis=0)
    flips = np.array([0, 1, 2, 3, 0, 3, 1, 2, 0], dtype=np.uint8)
    for dtype in (np.float64, np.float32):
        r = raw.astype(dtype)
        ref = orc.prepare_tiles(r, flips)
        got = ctk.prepare_tiles(torch.from_numpy(r).cuda(), torch.from_numpy(flips).cuda()).cpu().numpy()
        assert np.array_equal(got, ref)
        got0 = ctk.prepare_tiles(torch.from_numpy(r).cuda()).cpu().numpy()
        assert np.array_equal(got0, orc.prepare_tiles(r))
    small = rng.random((3, 2, 40, 24))
    fl = np.array([3, 0, 1], dtype=np.uint8)
    assert np.array_equal(ctk.prepare_tiles(torch.from_numpy(small).cuda(), torch.from_numpy(fl).cuda()).cpu().numpy(),
                          orc.prepare_tiles(small, fl))
    assert ctk.prepare_tiles(torch.empty(0, 2, 256, 256, dtype=torch.float64).cuda()).shape == (0, 2, 256, 256)
    with pytest.raises(ctk.CtkError):
        ctk.prepare_tiles(torch.zeros(1, 2, 8, 8, dtype=torch.float64))      # host tensor: no CPU fallback


def test_tile_ssim(ctk, golden):
    """Structural similarity (test-cross-talk-model.py:80-82) against the oracle's restatement of
    scikit-image's algorithm on scipy.ndimage.uniform_filter: the kernel follows the library's float32 / float64 operation
    order, so the agreement is at float64 summation order (1e-9 absolute asked here), not at float32 noise (~1e-7).
    Fixture tiles (normalised and raw intensities), synthetic tiles, a constant plane, both planes constant (0/0 -> NaN
    like the library), a ragged tile size, and the exact properties ssim(x, x) == 1 and symmetry on a full 256-tile batch."""
    tiles = torch.from_numpy(golden["tiles"].astype(np.float32))
    xn = torch.stack([torch.stack([torch.from_numpy(orc.normalize_image(t[0].numpy())),
                                   torch.from_numpy(orc.normalize_image(t[1].numpy()))]) for t in tiles])
    xs, _ = orc.synthetic_batch(3, seed=5)
    const = xs[:1].clone()
    const[0, 1] = 0.25
    quant = torch.round(xs[:1] * 6) / 6                      # large flat areas: variances cancel to exactly 0
    x = torch.cat([xn, tiles, xs, const, quant], dim=0).contiguous()
    got = ctk.ssim_per_image(x.cuda()).cpu().numpy()
    ref = np.array([orc.ssim_f32(t[0].numpy(), t[1].numpy()) for t in x])
    ref64 = np.array([orc.ssim_f64(t[0].numpy(), t[1].numpy()) for t in x])
    print("ssim gpu", got, "\n  max |gpu - oracle|", np.abs(got - ref).max(), " max |oracle f32 - f64 definition|", np.abs(ref - ref64).max())
    np.testing.assert_allclose(got, ref, atol=1e-9, rtol=0)
    np.testing.assert_allclose(got[:5], ref64[:5], atol=1e-6, rtol=0)         # normalised tiles: float32 noise of the definition
    with np.errstate(all="ignore"):
        both = torch.full((1, 2, 16, 16), 0.5)
        assert np.isnan(orc.ssim_f32(both[0, 0].numpy(), both[0, 1].numpy())) and np.isnan(ctk.ssim_per_image(both.cuda()).item())
    rag = torch.rand(3, 2, 40, 72)
    np.testing.assert_allclose(ctk.ssim_per_image(rag.cuda()).cpu().numpy(),
                               [orc.ssim_f32(t[0].numpy(), t[1].numpy()) for t in rag], atol=1e-9, rtol=0)
    small = torch.rand(2, 2, 7, 8)                            # a single window row
    np.testing.assert_allclose(ctk.ssim_per_image(small.cuda()).cpu().numpy(),
                               [orc.ssim_f32(t[0].numpy(), t[1].numpy()) for t in small], atol=1e-9, rtol=0)
    assert ctk.ssim_per_image(torch.empty(0, 2, 32, 32).cuda()).numel() == 0
    # full-size properties: identical planes give exactly 1, swapping the planes changes nothing
    big, _ = orc.synthetic_batch(256, seed=11)
    big = big.cuda()
    same = big.clone()
    same[:, 1] = same[:, 0]
    assert torch.equal(ctk.ssim_per_image(same), torch.ones(256, dtype=torch.float64, device="cuda"))
    a, b = ctk.ssim_per_image(big), ctk.ssim_per_image(big.flip(1).contiguous())
    assert torch.allclose(a, b, rtol=0, atol=1e-13)
    np.testing.assert_allclose(a[:4].cpu().numpy(), [orc.ssim_f32(t[0].numpy(), t[1].numpy()) for t in big[:4].cpu()], atol=1e-9, rtol=0)
    with pytest.raises(ctk.CtkError):
        ctk.ssim_per_image(torch.zeros(1, 2, 6, 32).cuda())   # smaller than the window
    with pytest.raises(ctk.CtkError):
        ctk.ssim_per_image(torch.zeros(1, 2, 32, 32))         # host tensor: no CPU fallback


def test_tile_metrics_fused(ctk, golden):
    """Pearson + RMSE + 256-bin histograms + histogram correlation in one fused pass against the oracle
    restatement, the reference-call goldens, and np.histogram bit for bit -- on fixture tiles (normalised and raw),
    synthetic tiles, a constant plane, flat histograms and a ragged plane size."""
    import json
    import os
    tiles = torch.from_numpy(golden["tiles"].astype(np.float32))
    xn = torch.stack([torch.stack([torch.from_numpy(orc.normalize_image(t[0].numpy())),
                                   torch.from_numpy(orc.normalize_image(t[1].numpy()))]) for t in tiles])
    xs, _ = orc.synthetic_batch(3, seed=5)
    const = xs[:1].clone()
    const[0, 1] = 0.25                                     # constant plane: Pearson NaN, histogram all in one bin
    x = torch.cat([xn, tiles, xs, const], dim=0).contiguous()
    got = ctk.tile_metrics(x.cuda())
    ref = orc.tile_metrics_batch(x)
    hist = got["hist"].cpu().numpy().astype(np.int64)
    assert np.array_equal(hist, ref["hist"])                                              # integer work: bit-exact
    for i in range(x.shape[0]):
        for c in range(2):
            assert np.array_equal(hist[i, c], np.histogram(x[i, c].numpy().flatten(), bins=256)[0])
    np.testing.assert_allclose(got["pearson"].cpu().numpy(), ref["pearson"], atol=1e-9, equal_nan=True)
    np.testing.assert_allclose(got["hist_corr"].cpu().numpy(), ref["hist_corr"], atol=1e-12, equal_nan=True)
    np.testing.assert_allclose(got["rmse"].cpu().numpy(), ref["rmse"], rtol=2e-6)          # fp32 pairwise mean vs fp64 sum
    assert np.isnan(got["pearson"][-1].item()) and not np.isnan(got["hist_corr"][-1].item())
    m = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "metrics.json")))
    np.testing.assert_allclose(got["hist_corr"][:5].cpu().numpy(), m["normalised"]["hist_corr"], atol=1e-12)
    np.testing.assert_allclose(got["hist_corr"][5:10].cpu().numpy(), m["raw"]["hist_corr"], atol=1e-12)
    np.testing.assert_allclose(got["rmse"][:5].cpu().numpy(), m["normalised"]["rmse"], rtol=2e-6)
    # flat histograms (every bin hit equally often) -> NaN like np.std(hist) == 0; ragged plane (48 x 40)
    ramp = (torch.arange(256, dtype=torch.float32) / 255).repeat(8).reshape(1, 1, 32, 64).repeat(1, 2, 1, 1).contiguous()
    assert np.isnan(ctk.tile_metrics(ramp.cuda())["hist_corr"].item()) and np.isnan(orc.hist_correlation(ramp[0, 0].numpy(), ramp[0, 1].numpy()))
    rag = torch.rand(2, 2, 48, 40)
    g2, r2 = ctk.tile_metrics(rag.cuda()), orc.tile_metrics_batch(rag)
    assert np.array_equal(g2["hist"].cpu().numpy().astype(np.int64), r2["hist"])
    np.testing.assert_allclose(g2["hist_corr"].cpu().numpy(), r2["hist_corr"], atol=1e-12)
    assert ctk.tile_metrics(torch.empty(0, 2, 32, 32).cuda())["pearson"].numel() == 0
    # NMI of the digitised planes (sklearn's formula on NumPy's digitisation): fixture goldens and the oracle restatement
    nmi = ctk.nmi_per_image(x.cuda()).cpu().numpy()
    np.testing.assert_allclose(nmi, ref["nmi"], atol=1e-12)
    np.testing.assert_allclose(nmi[:5], m["normalised"]["nmi"], atol=1e-12)
    np.testing.assert_allclose(nmi[5:10], m["raw"]["nmi"], atol=1e-12)
    assert nmi[-1] == 0.0                                                   # one constant plane
    both_const = torch.full((1, 2, 32, 32), 0.5)
    assert ctk.nmi_per_image(both_const.cuda()).item() == 1.0
    np.testing.assert_allclose(ctk.nmi_per_image(rag.cuda()).cpu().numpy(), r2["nmi"], atol=1e-12)
    # a range spanning fewer than 256 float32 values: runs of equal linspace edges (np.histogram refuses such planes,
    # np.digitize does not)
    narrow = (1000.0 + torch.rand(2, 2, 64, 64) * 1e-3).float().contiguous()
    np.testing.assert_allclose(ctk.nmi_per_image(narrow.cuda()).cpu().numpy(),
                               [orc.nmi_digitized(t[0].numpy(), t[1].numpy()) for t in narrow], atol=1e-12)
    # same Pearson as the stand-alone kernel
    np.testing.assert_allclose(got["pearson"].cpu().numpy(), ctk.pearson_per_image(x.cuda()).cpu().numpy(), atol=1e-12, equal_nan=True)


@pytest.mark.parametrize("cin,cout,c_off", [(1, 64, 0), (1, 64, 1), (2, 128, 0)])
def test_conv_first_eval(ctk, cin, cout, c_off):
    from ctk._lib import call, ptr, stream
    torch.manual_seed(2)
    n, H, W = 3, 64, 96
    x = torch.rand(n, 2, H, W)
    conv = torch.nn.Conv2d(cin, cout, 3, 1, 1)
    bn = torch.nn.BatchNorm2d(cout)
    with torch.no_grad():
        bn.weight.copy_(torch.randn(cout))
        bn.bias.copy_(0.1 * torch.randn(cout))
        bn.running_mean.copy_(0.1 * torch.randn(cout))
        bn.running_var.copy_(0.5 + torch.rand(cout))
    bn.eval()
    with torch.no_grad():
        ref = F.max_pool2d(F.leaky_relu(bn(conv(x[:, c_off:c_off + cin])), 0.01), 2).permute(0, 2, 3, 1).contiguous()
    conv, bn = conv.cuda(), bn.cuda()
    scale = torch.empty(cout, device="cuda")
    shift = torch.empty(cout, device="cuda")
    call("ctk_fold_bn_eval", ptr(conv.bias), ptr(bn.weight), ptr(bn.bias), ptr(bn.running_mean), ptr(bn.running_var),
         c_float(bn.eps), c_int(cout), ptr(scale), ptr(shift), stream())
    wf = torch.empty(cout, cin * 9, device="cuda")
    call("ctk_pack_first_weight", ptr(conv.weight), ptr(scale), c_int(cout), c_int(cin), ptr(wf), stream())
    cstride = cout + 64
    out = torch.zeros(n, H // 2, W // 2, cstride, device="cuda", dtype=torch.bfloat16)
    xd = x.cuda()                                              # keep every device operand alive until the sync
    call("ctk_conv_first_eval", ptr(xd), c_int(n), c_int(2), c_int(c_off), c_int(cin), c_int(H), c_int(W), ptr(wf),
         ptr(shift), c_int(cout), c_float(0.01), ptr(out), c_int(cstride), c_int(64), stream())
    torch.cuda.synchronize()
    assert out[..., :64].abs().max().item() == 0.0            # channel offset respected
    _bf16_close(out[..., 64:], ref, rel=2 ** -7, abs_=1e-4)   # fp32-class math, bf16 store (negative side rounded twice)


@pytest.mark.parametrize("cin,cout,c_off,H,W", [(1, 64, 1, 64, 96), (2, 128, 0, 32, 32), (1, 64, 0, 36, 20)])
def test_conv_first_pool_codes(ctk, cin, cout, c_off, H, W):
    """Pooled first block + 4-bit arg-max / sign codes (what the training backward consumes) against
    F.max_pool2d(return_indices=True) of the fp32 reference; windows whose top two entries (or whose maximum and 0)
    are closer than the bf16-operand-split accuracy are left out of the code comparison."""
    from ctk._lib import call, ptr, stream
    torch.manual_seed(3)
    n = 2
    x = torch.rand(n, 2, H, W)
    w = torch.randn(cout, cin, 3, 3) / 3
    scale = torch.randn(cout) * 0.5 + 1.0
    shift = 0.2 * torch.randn(cout)
    z = F.conv2d(x[:, c_off:c_off + cin], w * scale[:, None, None, None], padding=1) + shift[None, :, None, None]
    pooled, idx = F.max_pool2d(F.leaky_relu(z, 0.01), 2, return_indices=True)
    zmax = F.max_pool2d(z, 2)
    wd, sd, hd, xd = w.cuda(), scale.cuda(), shift.cuda(), x.cuda()
    wf = torch.empty(cout, cin * 9, device="cuda")
    call("ctk_pack_first_weight", ptr(wd), ptr(sd), c_int(cout), c_int(cin), ptr(wf), stream())
    out = torch.zeros(n, H // 2, W // 2, cout, device="cuda", dtype=torch.bfloat16)
    codes = torch.zeros(n, H // 2, W // 2, cout // 8, device="cuda", dtype=torch.int32)
    call("ctk_conv_first_pool_codes", ptr(xd), c_int(n), c_int(2), c_int(c_off), c_int(cin), c_int(H), c_int(W), ptr(wf),
         ptr(hd), c_int(cout), c_float(0.01), ptr(out), c_int(cout), c_int(0), ptr(codes), stream())
    torch.cuda.synchronize()
    _bf16_close(out, pooled.permute(0, 2, 3, 1).contiguous(), rel=2 ** -7, abs_=1e-4)
    cw = codes.cpu().to(torch.int64) & 0xFFFFFFFF
    nib = torch.stack([(cw >> (4 * i)) & 0xF for i in range(8)], dim=-1).reshape(n, H // 2, W // 2, cout)
    pos_gpu, neg_gpu = nib & 3, (nib >> 2) & 1
    iy, ix = idx // W, idx % W
    pos_ref = ((iy % 2) * 2 + (ix % 2)).permute(0, 2, 3, 1)
    neg_ref = (zmax < 0).long().permute(0, 2, 3, 1)
    win = z.unfold(2, 2, 2).unfold(3, 2, 2).reshape(n, cout, H // 2, W // 2, 4)
    top2 = win.topk(2, dim=-1).values
    clear_arg = ((top2[..., 0] - top2[..., 1]) > 1e-4).permute(0, 2, 3, 1)
    clear_sign = (top2[..., 0].abs() > 1e-4).permute(0, 2, 3, 1)
    assert clear_arg.float().mean() > 0.99
    assert torch.equal(pos_gpu[clear_arg], pos_ref[clear_arg])
    assert torch.equal(neg_gpu[clear_sign], neg_ref[clear_sign])


# ------------------------------------------------------------------ tensor-core conv block
def _conv_tc_case(ctk, n, H, W, cin, cout, flags=0, coff=0, extra=0):
    from ctk._lib import call, ptr, stream
    torch.manual_seed(3)
    x = torch.randn(n, H, W, cin).to(torch.bfloat16)
    w = (torch.randn(cout, cin, 3, 3) / (3.0 * cin ** 0.5)).to(torch.bfloat16).float()
    scale = 0.5 + torch.rand(cout)
    scale[::3] *= -1.0                  # negative BN scale: pooling must come after the affine + LeakyReLU
    shift = 0.1 * torch.randn(cout)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w, padding=1) * scale[None, :, None, None] + shift[None, :, None, None]
    ref = F.max_pool2d(F.leaky_relu(ref, 0.01), 2).permute(0, 2, 3, 1).contiguous()
    wp = torch.empty(9, cout, cin, device="cuda", dtype=torch.bfloat16)
    xd, wd, sc, sh = x.cuda(), w.cuda(), scale.cuda(), shift.cuda()   # keep device operands alive until the sync
    call("ctk_pack_conv_weight_bf16", ptr(wd), c_int(cout), c_int(cin), ptr(wp), stream())
    cstride = cout + extra
    out = torch.zeros(n, H // 2, W // 2, cstride, device="cuda", dtype=torch.bfloat16)
    call("ctk_conv3x3_tc_eval", ptr(xd), c_int(n), c_int(H), c_int(W), c_int(cin), ptr(wp), c_int(cout),
         ptr(sc), ptr(sh), c_float(0.01), ptr(out), c_int(cstride), c_int(coff), c_int(flags), stream())
    torch.cuda.synchronize()
    _bf16_close(out[..., coff:coff + cout], ref)
    if extra:
        rest = torch.cat([out[..., :coff], out[..., coff + cout:]], dim=-1)
        assert rest.abs().max().item() == 0.0


@pytest.mark.parametrize("n,H,W,cin,cout", [
    (2, 32, 32, 64, 128),     # two tiles per image row band, one K chunk
    (3, 16, 16, 128, 256),    # two K chunks, two N tiles
    (1, 8, 8, 512, 512),      # image smaller than the 16-row tile (single-branch block 6)
    (2, 48, 24, 64, 128),     # ragged: H not a multiple of the tile height
    (5, 16, 8, 256, 512),     # odd batch, odd number of spatial tiles (padded CTA pair)
    (3, 32, 32, 128, 128),    # cout = 128 with two K chunks
])
@pytest.mark.parametrize("flags", [0, 4], ids=["cta_pair", "single_cta"])
def test_conv_tc_eval_shapes(ctk, n, H, W, cin, cout, flags):
    _conv_tc_case(ctk, n, H, W, cin, cout, flags=flags)


def test_conv_tc_eval_channel_offset(ctk):
    _conv_tc_case(ctk, 2, 16, 16, 64, 128, coff=128, extra=256)


@pytest.mark.parametrize("flags", [0, 4], ids=["cta_pair", "single_cta"])
@pytest.mark.parametrize("cin,cout", [(64, 128), (128, 256)])
def test_conv_tc_many_tiles_persistent(ctk, cin, cout, flags):
    # more tiles than SMs so every CTA loops: exercises barrier phase wrap-around and TMEM double buffering
    _conv_tc_case(ctk, 8, 64, 64, cin, cout, flags=flags)


# ------------------------------------------------------------------ split-K GEMM + head
def test_gemm_splitk(ctk):
    from ctk._lib import call, ptr, stream
    torch.manual_seed(4)
    M, N, K, splits = 256, 512, 4096, 8
    a = torch.randn(M, K).to(torch.bfloat16)
    b = (torch.randn(N, K) / K ** 0.5).to(torch.bfloat16)
    ref = a.double() @ b.double().t()
    part = torch.empty(splits, M, N, device="cuda")
    ad, bd = a.cuda(), b.cuda()
    call("ctk_gemm_bf16_splitk", ptr(ad), ptr(bd), c_int(M), c_int(N), c_int(K), c_int(splits), ptr(part),
         stream())
    torch.cuda.synchronize()
    got = part.double().sum(0).cpu()
    assert (got - ref).abs().max().item() <= 1e-3
    # each split is the partial sum over its own K range
    ks = K // splits
    ref0 = a[:, :ks].double() @ b[:, :ks].double().t()
    assert (part[0].double().cpu() - ref0).abs().max().item() <= 1e-3


def test_gemm_splitk_ragged_splits(ctk):
    """The split count follows the SM count, not the divisors of K / 64: 64 K blocks over 18 splits = 10 splits of 4 blocks and
    8 of 3, each split the partial sum over its own contiguous K range."""
    from ctk._lib import call, ptr, stream
    torch.manual_seed(6)
    M, N, K, splits = 128, 256, 4096, 18
    a = torch.randn(M, K).to(torch.bfloat16)
    b = (torch.randn(N, K) / K ** 0.5).to(torch.bfloat16)
    part = torch.full((splits, M, N), float("nan"), device="cuda")
    ad, bd = a.cuda(), b.cuda()
    call("ctk_gemm_bf16_splitk", ptr(ad), ptr(bd), c_int(M), c_int(N), c_int(K), c_int(splits), ptr(part), stream())
    torch.cuda.synchronize()
    base, extra = divmod(K // 64, splits)
    k0 = 0
    for s_ in range(splits):
        k1 = k0 + 64 * (base + (1 if s_ < extra else 0))
        ref = a[:, k0:k1].double() @ b[:, k0:k1].double().t()
        assert (part[s_].double().cpu() - ref).abs().max().item() <= 1e-3, s_
        k0 = k1
    assert k0 == K
    ref = a.double() @ b.double().t()
    assert (part.double().sum(0).cpu() - ref).abs().max().item() <= 1e-3


def test_head_eval(ctk):
    from ctk._lib import call, ptr, stream
    torch.manual_seed(5)
    n, splits, mstride, f1, f2 = 37, 4, 128, 512, 128
    part = torch.randn(splits, mstride, f1)
    s1, h1 = 0.5 + torch.rand(f1), 0.1 * torch.randn(f1)
    w2 = torch.randn(f2, f1) / f1 ** 0.5
    s2, h2 = 0.5 + torch.rand(f2), 0.1 * torch.randn(f2)
    w3, b3 = torch.randn(f2) / f2 ** 0.5, torch.randn(1)
    a = F.leaky_relu(part.sum(0)[:n] * s1 + h1, 0.01)
    b = F.leaky_relu((a @ w2.t()) * s2 + h2, 0.01)
    z = b @ w3 + b3
    dv = [t.cuda() for t in (part, s1, h1, w2, s2, h2, w3, b3)]
    for sig in (0, 1):
        out = torch.empty(n, device="cuda")
        call("ctk_head_eval", ptr(dv[0]), c_int(splits), c_int(mstride), c_int(n), c_int(f1), c_int(f2),
             ptr(dv[1]), ptr(dv[2]), ptr(dv[3]), ptr(dv[4]), ptr(dv[5]), ptr(dv[6]), ptr(dv[7]), c_float(0.01),
             c_int(sig), ptr(out), stream())
        ref = 0.5 * torch.sigmoid(z) if sig else z
        assert (out.cpu() - ref).abs().max().item() <= 2e-5


# ------------------------------------------------------------------ MSE + Adam
def test_mse_loss(ctk):
    torch.manual_seed(6)
    o, t = torch.rand(300, 1), torch.rand(300, 1)
    loss, grad = ctk.mse_loss(o.cuda(), t.cuda(), want_grad=True)
    assert loss.item() == pytest.approx(orc.mse_loss(o, t).item(), rel=1e-6)
    np.testing.assert_allclose(grad.cpu().numpy(), (2 * (o - t) / 300).numpy(), rtol=1e-6, atol=1e-9)


def test_adam_matches_oracle_and_torch(ctk):
    torch.manual_seed(7)
    shapes = [(70000,), (513, 9), (3,), (128, 64, 3, 3)]
    ps = [torch.randn(s) for s in shapes]
    dev = [torch.nn.Parameter(p.clone().cuda()) for p in ps]
    opt = ctk.Adam(dev, lr=5e-4, weight_decay=1e-4)
    ref = [p.clone() for p in ps]
    m = [torch.zeros_like(p) for p in ps]
    v = [torch.zeros_like(p) for p in ps]
    for t in range(1, 4):
        gs = [torch.randn(s) * 0.1 for s in shapes]
        for d, g in zip(dev, gs):
            d.grad = g.clone().cuda()
        opt.step()
        for p, g, mm, vv in zip(ref, gs, m, v):
            orc.adam_step(p, g, mm, vv, t, 5e-4)
    for d, p in zip(dev, ref):
        np.testing.assert_allclose(d.detach().cpu().numpy(), p.numpy(), rtol=0, atol=2e-6)
    assert opt.param_groups[0]["lr"] == 5e-4 and int(opt.state[dev[0]]["step"]) == 3
