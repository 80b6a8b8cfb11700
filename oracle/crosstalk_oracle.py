"""CPU oracle for the CrosstalkPy hot path.  TEST INFRASTRUCTURE ONLY.

This file is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product path
(``torch-unet_b200/``) must never import anything from ``oracle/``.

It restates, as plain functional PyTorch-on-CPU / NumPy code over a
``state_dict``, the arithmetic the reference performs on its hot path:

* single-branch CNN forward   -- /root/reference/regression_model.py:5-61
* double-branch CNN forward   -- /root/reference/two_branch_regression.py:5-100
* MSE loss                    -- /root/reference/train_model.py:636,421
* Adam(weight_decay) update   -- /root/reference/train_model.py:637,424
* per-image Pearson r         -- /root/reference/test-cross-talk-model.py:59-64
* RMSE, 256-bin histogram correlation, NMI of the digitised planes
                              -- /root/reference/test-cross-talk-model.py:65-79,84
* structural similarity (scikit-image's algorithm; PARITY UNPINNED, see ssim_f32)
                              -- /root/reference/test-cross-talk-model.py:80-82
* per-plane min-max normalise, cast and flips of the input pipeline
                              -- /root/reference/train_model.py:166-167,211-232

The arithmetic itself lives in third-party libraries that are not vendored in
the reference (PyTorch and NumPy -- unpinned in requirements.txt:1-2 -- and SciPy /
scikit-learn, which are not listed at all).  The versions this oracle was pinned
against are torch 2.11.0, numpy 2.3.5, scipy 1.18.1 and scikit-learn 1.9.0.  The reference ships no tests and no golden
vectors, so the pins are created by ``tests/golden/make_golden.py``, which
imports the *unmodified* reference modules from /root/reference, runs them on
the reference's own ``Training_Data`` fixtures and records their outputs in
``tests/golden/golden.json`` (``make_loss_curve.py``: 200 steps of the reference's
training loop; ``make_metrics_golden.py``: the reference's metric expressions
verbatim); ``tests/test_oracle_golden.py`` holds the oracle to those numbers.  Parity status: pinned against reference outputs generated
in the build container (not against reference-owned tests, which do not exist).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

LEAKY_SLOPE = 0.01          # regression_model.py:16,25,38,43 / two_branch_regression.py:12,...
BN_EPS = 1e-5               # nn.BatchNorm defaults
BN_MOMENTUM = 0.1

SINGLE_CHANNELS = (2, 128, 256, 512, 512, 512, 512)   # train_model.py:537 (128 filters, 6 blocks, cap 512)
DOUBLE_CHANNELS = (1, 64, 128, 256, 512)               # train_model.py:535 (64 filters per branch)
SINGLE_CONV_IDX = (0, 4, 8, 12, 16, 20)
DOUBLE_CONV_IDX = (0, 4, 8, 12)


# --------------------------------------------------------------------------
# state_dict construction (same keys / shapes / init RNG stream as the reference)
# --------------------------------------------------------------------------
def _conv_bn(sd, prefix, idx, cin, cout):
    conv = torch.nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1)
    bn = torch.nn.BatchNorm2d(cout)
    sd[f"{prefix}.{idx}.weight"] = conv.weight.detach().clone()
    sd[f"{prefix}.{idx}.bias"] = conv.bias.detach().clone()
    for k, v in bn.state_dict().items():
        sd[f"{prefix}.{idx + 1}.{k}"] = v.detach().clone()


def _fc_head(sd, prefix, in_features):
    for idx, (fi, fo, has_bn) in zip((1, 5, 9), ((in_features, 512, True), (512, 128, True), (128, 1, False))):
        lin = torch.nn.Linear(fi, fo)
        sd[f"{prefix}.{idx}.weight"] = lin.weight.detach().clone()
        sd[f"{prefix}.{idx}.bias"] = lin.bias.detach().clone()
        if has_bn:
            bn = torch.nn.BatchNorm1d(fo)
            for k, v in bn.state_dict().items():
                sd[f"{prefix}.{idx + 1}.{k}"] = v.detach().clone()


def init_single_state_dict(seed: Optional[int] = 0) -> Dict[str, torch.Tensor]:
    """state_dict of AdvancedRegressionModel(initial_filters=128, num_conv_blocks=6).

    Consumes the RNG exactly like the reference constructor (regression_model.py:6-50:
    conv/bn pairs in order, then Linear/BatchNorm1d in order) and replays the
    train-mode dummy pass of ``_get_conv_output`` (regression_model.py:52-56), which
    leaves every conv-stack BN with num_batches_tracked=1.
    """
    if seed is not None:
        torch.manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    ch = SINGLE_CHANNELS
    for i, idx in enumerate(SINGLE_CONV_IDX):
        _conv_bn(sd, "conv_layers", idx, ch[i], ch[i + 1])
    # dummy pass in train mode on zeros(1, 2, 256, 256): updates running stats
    with torch.no_grad():
        x = torch.zeros(1, 2, 256, 256)
        for idx in SINGLE_CONV_IDX:
            x = _conv_block(sd, "conv_layers", idx, x, training=True, update_stats=True)
    _fc_head(sd, "fc_layers", 512 * 4 * 4)
    return sd


def init_double_state_dict(seed: Optional[int] = 0) -> Dict[str, torch.Tensor]:
    """state_dict of SimplifiedTwoBranchRegressionModel(initial_filters_per_branch=64).

    two_branch_regression.py:60-83: bleed branch, source branch, then the head; the
    dummy pass runs in eval mode so the BN buffers stay pristine.
    """
    if seed is not None:
        torch.manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    ch = DOUBLE_CHANNELS
    for br in ("bleed_branch", "source_branch"):
        for i, idx in enumerate(DOUBLE_CONV_IDX):
            _conv_bn(sd, f"{br}.conv_blocks", idx, ch[i], ch[i + 1])
    _fc_head(sd, "regression_head.fc_layers", 1024 * 16 * 16)
    return sd


def randomize_bn(sd: Dict[str, torch.Tensor], seed: int = 7) -> Dict[str, torch.Tensor]:
    """Give every BN layer non-trivial gamma (some negative), beta and running stats.

    Random-init eval outputs are nearly constant, which makes an
    absolute tolerance vacuous; this is fallback weight set (iii) of SURVEY 8c.
    """
    g = torch.Generator().manual_seed(seed)
    out = {k: v.clone() for k, v in sd.items()}
    for k in list(out.keys()):
        if k.endswith("running_mean"):
            base = k[: -len("running_mean")]
            n = out[k].numel()
            gamma = 0.5 + torch.rand(n, generator=g)
            sign = torch.where(torch.rand(n, generator=g) < 0.25, -1.0, 1.0)
            out[base + "weight"] = gamma * sign
            out[base + "bias"] = 0.2 * torch.randn(n, generator=g)
            out[base + "running_mean"] = out[base + "running_mean"] + 0.05 * torch.randn(n, generator=g)
            out[base + "running_var"] = out[base + "running_var"] * (0.5 + torch.rand(n, generator=g))
    return out


def calibrate_bn(kind: str, sd: Dict[str, torch.Tensor], x: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Weights whose eval outputs have real spread: one train-mode pass over ``x`` with BN momentum 1.0, so every
    running_mean / running_var becomes that batch's statistics (what nn.BatchNorm does with momentum=1.0).
    tests/golden/make_golden.py does the same to the reference modules (``bn.momentum = 1.0``)."""
    global BN_MOMENTUM
    out = {k: v.clone() for k, v in sd.items()}
    old, BN_MOMENTUM = BN_MOMENTUM, 1.0
    try:
        with torch.no_grad():
            FORWARD[kind](out, x, training=True, update_stats=True, dropout_masks=None)
    finally:
        BN_MOMENTUM = old
    return out


# --------------------------------------------------------------------------
# optional bf16-rounding restatement of the GPU path (precision model, not the reference)
# --------------------------------------------------------------------------
# The CUDA path keeps conv operands (weights of blocks 2+, block inputs/outputs, FC1 operands and the gradients that
# flow between blocks) in bf16 and accumulates in fp32.  With EMULATE_BF16 = True the same roundings are applied here,
# everything else staying fp32, so tests can separate "kernel is wrong" from "bf16 operands lose this much".
EMULATE_BF16 = False


class _RoundBoth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).float()


class _RoundFwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).float()


class emulate_bf16:
    """Context manager: ``with orc.emulate_bf16(): ...`` runs the forward/backward with the GPU path's roundings."""

    def __enter__(self):
        global EMULATE_BF16
        self.old, EMULATE_BF16 = EMULATE_BF16, True

    def __exit__(self, *a):
        global EMULATE_BF16
        EMULATE_BF16 = self.old


# --------------------------------------------------------------------------
# forward passes
# --------------------------------------------------------------------------
def _batch_norm(sd, key, x, training, update_stats):
    rm, rv = sd[f"{key}.running_mean"], sd[f"{key}.running_var"]
    if training and not update_stats:
        rm, rv = None, None          # batch statistics, buffers untouched
    y = F.batch_norm(x, rm, rv, sd[f"{key}.weight"], sd[f"{key}.bias"],
                     training=training, momentum=BN_MOMENTUM, eps=BN_EPS)
    if training and update_stats:
        sd[f"{key}.num_batches_tracked"] = sd[f"{key}.num_batches_tracked"] + 1
    return y


def _conv_block(sd, prefix, idx, x, training=False, update_stats=False, taps=None):
    """Conv3x3(p=1)+bias -> BatchNorm2d -> LeakyReLU(0.01) -> MaxPool2d(2,2).

    regression_model.py:14-17,23-26 / two_branch_regression.py:10-13,...
    """
    w = sd[f"{prefix}.{idx}.weight"]
    if EMULATE_BF16 and idx != 0:            # the first block runs split-bf16 (fp32-class) on the GPU
        w = _RoundFwd.apply(w)
    y = F.conv2d(x, w, sd[f"{prefix}.{idx}.bias"], stride=1, padding=1)
    if EMULATE_BF16 and training:            # train mode stores the raw conv output in bf16
        y = _RoundBoth.apply(y)
    if taps is not None:
        taps[f"{prefix}.{idx}.conv"] = y
    y = _batch_norm(sd, f"{prefix}.{idx + 1}", y, training, update_stats)
    y = F.leaky_relu(y, LEAKY_SLOPE)
    y = F.max_pool2d(y, kernel_size=2, stride=2)
    if EMULATE_BF16:
        y = _RoundBoth.apply(y)
    if taps is not None:
        taps[f"{prefix}.{idx}.pool"] = y
    return y


def _head(sd, prefix, x, training, update_stats, dropout_p, dropout_masks, taps):
    """Flatten -> Linear -> BN1d -> LeakyReLU -> Dropout -> Linear -> BN1d -> LeakyReLU -> Dropout -> Linear.

    regression_model.py:34-50 / two_branch_regression.py:40-54.  ``dropout_masks`` are
    explicit keep-masks ([N,512], [N,128], 0/1) so both paths can be fed the same
    Bernoulli draw; ``None`` in training mode means "no dropout" (p forced to 0).
    """
    x = torch.flatten(x, 1)                       # NCHW order: c*H*W + h*W + w
    for j, idx in enumerate((1, 5)):
        w = sd[f"{prefix}.{idx}.weight"]
        if EMULATE_BF16 and idx == 1:
            w = _RoundFwd.apply(w)
        x = F.linear(x, w, sd[f"{prefix}.{idx}.bias"])
        if EMULATE_BF16 and idx == 1:
            x = _RoundBwd.apply(x)           # dZ1 feeds the FC1 backward GEMMs as bf16
        if taps is not None:
            taps[f"{prefix}.{idx}.fc"] = x
        x = _batch_norm(sd, f"{prefix}.{idx + 1}", x, training, update_stats)
        x = F.leaky_relu(x, LEAKY_SLOPE)
        if training and dropout_masks is not None:
            x = x * dropout_masks[j] * (1.0 / (1.0 - dropout_p))
    x = F.linear(x, sd[f"{prefix}.9.weight"], sd[f"{prefix}.9.bias"])
    return x


def single_forward(sd, x, training=False, update_stats=False, dropout_masks=None, taps=None):
    """AdvancedRegressionModel.forward -- regression_model.py:58-61 (dropout p=0.1, :39,44)."""
    for idx in SINGLE_CONV_IDX:
        x = _conv_block(sd, "conv_layers", idx, x, training, update_stats, taps)
    return _head(sd, "fc_layers", x, training, update_stats, 0.1, dropout_masks, taps)


def double_forward(sd, x, training=False, update_stats=False, dropout_masks=None, taps=None):
    """SimplifiedTwoBranchRegressionModel.forward -- two_branch_regression.py:85-100 (dropout p=0.5, :45,50)."""
    feats = []
    for br, ch in (("bleed_branch", 0), ("source_branch", 1)):
        y = x[:, ch:ch + 1]                                   # :88-89
        for idx in DOUBLE_CONV_IDX:
            y = _conv_block(sd, f"{br}.conv_blocks", idx, y, training, update_stats, taps)
        feats.append(y)
    y = torch.cat(feats, dim=1)                                # :96
    z = _head(sd, "regression_head.fc_layers", y, training, update_stats, 0.5, dropout_masks, taps)
    return torch.sigmoid(z) * 0.5                              # :53,100


FORWARD = {"single": single_forward, "double": double_forward}
INIT = {"single": init_single_state_dict, "double": init_double_state_dict}


def mse_loss(out: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """torch.nn.MSELoss() (mean reduction) -- train_model.py:636,421."""
    return ((out - target) ** 2).mean()


# --------------------------------------------------------------------------
# training step: autograd over the functional forward + restated Adam
# --------------------------------------------------------------------------
def is_param(key: str) -> bool:
    return not (key.endswith("running_mean") or key.endswith("running_var") or key.endswith("num_batches_tracked"))


def loss_and_grads(kind, sd, x, y, dropout_masks=None, update_stats=True):
    """zero_grad -> forward -> MSELoss -> backward (train_model.py:419-422). Returns (loss, out, grads)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if is_param(k)}
    work = dict(sd)
    work.update(leaves)
    out = FORWARD[kind](work, x, training=True, update_stats=update_stats, dropout_masks=dropout_masks)
    loss = mse_loss(out, y)
    grads = torch.autograd.grad(loss, list(leaves.values()))
    if update_stats:
        for k in sd:
            if not is_param(k):
                sd[k] = work[k].detach()
    return loss.detach(), out.detach(), dict(zip(leaves.keys(), grads))


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-4):
    """One torch.optim.Adam update with coupled L2 on fp32 tensors, in place.

    train_model.py:637: optim.Adam(lr, weight_decay=1e-4); the update follows
    torch/optim/adam.py ``_single_tensor_adam``: g += wd*p; m.lerp_(g, 1-b1);
    v = b2*v + (1-b2) g*g; denom = sqrt(v)/sqrt(1-b2^t) + eps; p -= lr/(1-b1^t) * m/denom.
    """
    g = g + weight_decay * p
    m.lerp_(g, 1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = lr / bc1
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-step_size)
    return p, m, v


class OracleTrainer:
    """The reference's inner loop (train_model.py:415-426) over a state_dict."""

    def __init__(self, kind, sd, lr=5e-4, weight_decay=1e-4):
        self.kind, self.sd, self.lr, self.wd = kind, sd, lr, weight_decay
        self.m = {k: torch.zeros_like(v) for k, v in sd.items() if is_param(k)}
        self.v = {k: torch.zeros_like(v) for k, v in sd.items() if is_param(k)}
        self.t = 0

    def step(self, x, y, dropout_masks=None):
        loss, out, grads = loss_and_grads(self.kind, self.sd, x, y, dropout_masks)
        self.t += 1
        for k, g in grads.items():
            adam_step(self.sd[k], g, self.m[k], self.v[k], self.t, self.lr, weight_decay=self.wd)
        return float(loss), out


# --------------------------------------------------------------------------
# Pearson r (test-cross-talk-model.py:59-64)
# --------------------------------------------------------------------------
def pearson_f32(a: np.ndarray, b: np.ndarray) -> float:
    """scipy.stats.pearsonr restated for float32 input (dtype preserved).

    xm = x - mean(x); r = dot(xm/||xm||, ym/||ym||) clipped to [-1, 1]; NaN when
    either plane is constant (the np.std()==0 guard at :61-62 and scipy's own
    constant-input rule give the same answer).
    """
    x = np.asarray(a, dtype=np.float32).ravel()
    y = np.asarray(b, dtype=np.float32).ravel()
    if (x == x[0]).all() or (y == y[0]).all() or np.std(x) == 0 or np.std(y) == 0:
        return float("nan")
    xm = x - x.mean(dtype=np.float32)
    ym = y - y.mean(dtype=np.float32)
    nx = np.float32(np.linalg.norm(xm))
    ny = np.float32(np.linalg.norm(ym))
    r = np.dot(xm / nx, ym / ny)
    return float(max(min(np.float32(r), np.float32(1.0)), np.float32(-1.0)))


def pearson_f64(a: np.ndarray, b: np.ndarray) -> float:
    """Float64 restatement of the same quantity (the tie-breaker the f32 oracle is graded against)."""
    x = np.asarray(a, dtype=np.float64).ravel()
    y = np.asarray(b, dtype=np.float64).ravel()
    if x.min() == x.max() or y.min() == y.max():
        return float("nan")
    xm = x - x.mean()
    ym = y - y.mean()
    r = float(np.dot(xm, ym) / math.sqrt(np.dot(xm, xm) * np.dot(ym, ym)))
    return max(min(r, 1.0), -1.0)


def pearson_batch(x: torch.Tensor, f64: bool = True) -> np.ndarray:
    """Per-image r of channel 0 vs channel 1 of an [N,2,H,W] float32 batch."""
    xs = x.detach().cpu().numpy()
    fn = pearson_f64 if f64 else pearson_f32
    return np.array([fn(xs[i, 0], xs[i, 1]) for i in range(xs.shape[0])], dtype=np.float64)


# --------------------------------------------------------------------------
# RMSE and histogram correlation (test-cross-talk-model.py:65-70,79) -- SURVEY 8f row 1
# --------------------------------------------------------------------------
def rmse_f32(a: np.ndarray, b: np.ndarray) -> float:
    """``np.sqrt(np.mean((img0 - img1) ** 2))`` on float32 planes (test-cross-talk-model.py:79)."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    return float(np.sqrt(np.mean((a - b) ** 2)))


def histogram256_f32(a: np.ndarray) -> np.ndarray:
    """``np.histogram(plane.flatten(), bins=256)[0]`` restated for a float32 plane (test-cross-talk-model.py:65-66).

    NumPy >= 2 (numpy/lib/_histograms_impl.py, uniform-bin path) keeps everything in the array's dtype:
    edges = float32(arange(257) * float32((max-min)/256) + min) with the last edge set to max; a first guess
    ``int(((x - min) / (max - min)) * 256)`` (float32 ops, truncation), clamped from 256 to 255, is then corrected by at
    most one bin against the float32 edges (left-closed bins, the last bin closed on both sides).  A constant plane uses
    the range [v - 0.5, v + 0.5].  The CUDA kernel repeats exactly these float32 operations, so counts are bit-exact.
    """
    x = np.asarray(a, dtype=np.float32).ravel()
    first, last = x.min(), x.max()
    if first == last:
        first, last = np.float32(first - np.float32(0.5)), np.float32(last + np.float32(0.5))
    step = np.float32((last - first) / np.float32(256))
    edges = (np.arange(257, dtype=np.float32) * step + first).astype(np.float32)
    edges[-1] = last
    denom = np.float32(last - first)
    idx = (((x - first) / denom) * np.float32(256)).astype(np.int64)
    idx[idx == 256] -= 1
    idx[x < edges[idx]] -= 1
    inc = (x >= edges[idx + 1]) & (idx != 255)
    idx[inc] += 1
    return np.bincount(idx, minlength=256).astype(np.int64)


def hist_correlation(a: np.ndarray, b: np.ndarray) -> float:
    """Pearson r of the two 256-bin histograms, NaN if either histogram is flat (test-cross-talk-model.py:65-70)."""
    h1, h2 = histogram256_f32(a), histogram256_f32(b)
    if np.std(h1) == 0 or np.std(h2) == 0:
        return float("nan")
    return pearson_f64(h1.astype(np.float64), h2.astype(np.float64))


def digitize256_f32(a: np.ndarray) -> np.ndarray:
    """``np.digitize(plane.flatten(), bins=np.linspace(plane.min(), plane.max(), 256))`` restated for a float32 plane
    (test-cross-talk-model.py:71-74): label = number of float32 edges ``k * step + min`` (last edge = max) that are <= x,
    i.e. 1..256.  NumPy >= 2 keeps the edges in float32 (step = float32((max - min) / 255), one multiply and one add per
    edge).  Planes whose range spans fewer than 256 float32 values get runs of equal edges; counting (rather than
    guessing a bin from (x - min) / step) stays exact there.  Not restated: NumPy's denormal special case
    (max - min != 0 but step == 0)."""
    x = np.asarray(a, dtype=np.float32).ravel()
    lo, hi = x.min(), x.max()
    step = np.float32((hi - lo) / np.float32(255))
    edges = (np.arange(256, dtype=np.float32) * step + lo).astype(np.float32)
    edges[-1] = hi
    return np.searchsorted(edges, x, side="right").astype(np.int64)


def nmi_from_joint(joint: np.ndarray) -> float:
    """sklearn.metrics.normalized_mutual_info_score (average_method='arithmetic') from a contingency table:
    MI = sum_ij n_ij/N (log n_ij - log N - log a_i - log b_j + log A + log B), entropies -sum p (log n - log N), all in
    float64 with natural logs; 1.0 if both labelings are constant, 0.0 if |MI| < eps (sklearn/metrics/cluster/_supervised.py)."""
    joint = np.asarray(joint, dtype=np.int64)
    pi, pj = joint.sum(1), joint.sum(0)
    if (pi > 0).sum() == 1 and (pj > 0).sum() == 1:
        return 1.0
    if (pi > 0).sum() == 1 or (pj > 0).sum() == 1:
        return 0.0
    n = float(joint.sum())
    nzx, nzy = np.nonzero(joint)
    nz = joint[nzx, nzy].astype(np.float64)
    outer = pi[nzx].astype(np.int64) * pj[nzy].astype(np.int64)
    log_outer = -np.log(outer) + math.log(pi.sum()) + math.log(pj.sum())
    mi = (nz / n) * (np.log(nz) - math.log(n)) + (nz / n) * log_outer
    mi = np.where(np.abs(mi) < np.finfo(np.float64).eps, 0.0, mi)
    mi = float(np.clip(mi.sum(), 0.0, None))
    if abs(mi) < np.finfo(np.float64).eps:
        return 0.0

    def entropy(c):
        c = c[c > 0].astype(np.float64)
        if c.size == 1:
            return 0.0
        return float(-np.sum((c / c.sum()) * (np.log(c) - math.log(c.sum()))))
    return mi / (0.5 * (entropy(pi) + entropy(pj)))


def nmi_digitized(a: np.ndarray, b: np.ndarray) -> float:
    """normalized_mutual_info_score(digitize(img0), digitize(img1)) -- test-cross-talk-model.py:71-74,84."""
    la, lb = digitize256_f32(a), digitize256_f32(b)
    joint = np.zeros((256, 256), dtype=np.int64)
    np.add.at(joint, (la - 1, lb - 1), 1)
    return nmi_from_joint(joint)


def ssim_f32(a: np.ndarray, b: np.ndarray) -> float:
    """Mean structural similarity as the reference calls it -- test-cross-talk-model.py:80-82:
    ``ssim(img0, img1, data_range=max(img0.max(), img1.max()) - min(img0.min(), img1.min()))`` with scikit-image's
    defaults (7x7 uniform window, K1 = 0.01, K2 = 0.03, sample covariance, float32 arithmetic for float32 images,
    mean over the image cropped by 3 pixels, accumulated in float64).

    PARITY UNPINNED for this function: scikit-image is neither vendored in /root/reference nor installed in the build
    image, so no reference output could be recorded.  This is a restatement of the published algorithm of
    ``skimage.metrics.structural_similarity`` (scikit-image 0.19-0.25, metrics/_structural_similarity.py) on top of the
    one third-party routine it delegates to, ``scipy.ndimage.uniform_filter`` (scipy 1.18.1, installed here).
    """
    from scipy.ndimage import uniform_filter
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    data_range = np.max([a.max(), b.max()]) - np.min([a.min(), b.min()])       # np.float32, as at the call site
    win, k1, k2 = 7, 0.01, 0.03
    npix = win ** a.ndim
    cov_norm = npix / (npix - 1)                                              # use_sample_covariance=True
    ux = uniform_filter(a, size=win)
    uy = uniform_filter(b, size=win)
    uxx = uniform_filter(a * a, size=win)
    uyy = uniform_filter(b * b, size=win)
    uxy = uniform_filter(a * b, size=win)
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)
    c1 = (k1 * data_range) ** 2
    c2 = (k2 * data_range) ** 2
    a1, a2, b1, b2 = (2 * ux * uy + c1, 2 * vxy + c2, ux ** 2 + uy ** 2 + c1, vx + vy + c2)
    smap = (a1 * a2) / (b1 * b2)
    pad = (win - 1) // 2
    return float(smap[pad:-pad, pad:-pad].mean(dtype=np.float64))


def ssim_f64(a: np.ndarray, b: np.ndarray) -> float:
    """The same definition evaluated in float64 with direct window sums (no float32 rounding anywhere): the yardstick
    for how much of a difference is float32 noise of the definition itself."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    data_range = max(a.max(), b.max()) - min(a.min(), b.min())
    win = 7

    def box(img):
        v = np.lib.stride_tricks.sliding_window_view(img, (win, win))
        return v.mean(axis=(2, 3))

    ux, uy, uxx, uyy, uxy = box(a), box(b), box(a * a), box(b * b), box(a * b)
    cov_norm = 49.0 / 48.0
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    smap = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2))
    return float(smap.mean())


def tile_metrics_batch(x: torch.Tensor) -> Dict[str, np.ndarray]:
    """Per-image Pearson r, RMSE and histogram correlation of channel 0 vs channel 1 of an [N,2,H,W] float32 batch."""
    xs = x.detach().cpu().numpy()
    n = xs.shape[0]
    return {"pearson": np.array([pearson_f64(xs[i, 0], xs[i, 1]) for i in range(n)]),
            "rmse": np.array([rmse_f32(xs[i, 0], xs[i, 1]) for i in range(n)]),
            "hist_corr": np.array([hist_correlation(xs[i, 0], xs[i, 1]) for i in range(n)]),
            "nmi": np.array([nmi_digitized(xs[i, 0], xs[i, 1]) for i in range(n)]),
            "ssim": np.array([ssim_f32(xs[i, 0], xs[i, 1]) for i in range(n)]),
            "hist": np.stack([np.stack([histogram256_f32(xs[i, 0]), histogram256_f32(xs[i, 1])]) for i in range(n)])
            if n else np.zeros((0, 2, 256), dtype=np.int64)}


# --------------------------------------------------------------------------
# inputs
# --------------------------------------------------------------------------
def normalize_image(img: np.ndarray) -> np.ndarray:
    """train_model.py:211-216 (float32 arithmetic on a float32 plane)."""
    lo, hi = img.min(), img.max()
    if hi > lo:
        return (img - lo) / (hi - lo)
    return img


def prepare_tiles(raw: np.ndarray, flips: Optional[np.ndarray] = None) -> np.ndarray:
    """The reference's per-sample input path restated over a batch: raw [N,2,H,W] (float64 as stored in the TIFFs, or
    float32) -> ``.astype(np.float32)`` (train_model.py:166-167) -> normalize_image per plane (:211-216) -> hflip if
    ``flips[i] & 1``, vflip if ``flips[i] & 2`` on both planes (TF.hflip / TF.vflip, :225-232)."""
    x = np.asarray(raw).astype(np.float32)
    out = np.empty_like(x)
    for i in range(x.shape[0]):
        for c in range(2):
            p = normalize_image(x[i, c])
            if flips is not None and flips[i] & 1:
                p = p[:, ::-1]
            if flips is not None and flips[i] & 2:
                p = p[::-1, :]
            out[i, c] = p
    return out


def synthetic_batch(n: int, seed: int = 1234, size: int = 256) -> Tuple[torch.Tensor, torch.Tensor]:
    """SURVEY 8d generator: U[0,1) source plane, ch0 = normalise(alpha*src + (1-alpha)*noise), labels alpha."""
    g = torch.Generator().manual_seed(seed)
    src = torch.rand(n, size, size, generator=g)
    noise = torch.rand(n, size, size, generator=g)
    alpha = 0.01 + 0.49 * torch.rand(n, generator=g)
    mixed = alpha[:, None, None] * src + (1.0 - alpha[:, None, None]) * noise

    def norm(t):
        lo = t.amin(dim=(1, 2), keepdim=True)
        hi = t.amax(dim=(1, 2), keepdim=True)
        return (t - lo) / (hi - lo)

    x = torch.stack([norm(mixed), norm(src)], dim=1).contiguous()
    return x, alpha[:, None].contiguous()


def dropout_masks(n: int, p: float, seed: int) -> Tuple[torch.Tensor, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    return ((torch.rand(n, 512, generator=g) >= p).float(), (torch.rand(n, 128, generator=g) >= p).float())
