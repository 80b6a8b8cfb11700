"""Host-side driver of the CUDA hot path: owns the derived parameter cache (packed bf16 weights, folded BN)
and the activation buffers, and strings the libctk calls together for one forward pass.

The nn.Module (reference constructor, reference state_dict) stays the single source of truth for the
parameters; everything here is a cache keyed on the parameters' storage and version counters, rebuilt
when an optimizer step or load_state_dict changes them (checkpoint/resume row).
"""
from __future__ import annotations

from ctypes import c_float, c_int
from typing import Dict, List, Optional, Tuple

import os

import torch

from . import _lib
from ._lib import call, ptr, stream

LEAKY_SLOPE = 0.01     # nn.LeakyReLU(0.01): regression_model.py:16,25,38,43; two_branch_regression.py:12,18,24,30,44,49


_SM_COUNT = {}


def fc1_splits(tiles: int, K: int, dev, sms: Optional[int] = None) -> int:
    """Split-K factor of the FC1 GEMM: as many K ranges as there are SMs per output tile (18 x 8 tiles = 144 CTAs on a B200
    for a 256-tile batch), at least eight 64-element K blocks each.  The kernel deals the K blocks out raggedly, so the
    count need not divide K / 64."""
    if os.environ.get("CTK_FC1_SPLITS") == "pow2":      # round 1's rule (a power of two that divides K / 64), for A/B runs
        splits = 1
        while splits * 2 * tiles <= 160 and (K // 64) % (splits * 2) == 0 and K // (splits * 2) >= 512:
            splits *= 2
        return splits
    if sms is None:
        idx = dev.index if getattr(dev, "index", None) is not None else torch.cuda.current_device()
        sms = _SM_COUNT.get(idx)
        if sms is None:
            sms = _SM_COUNT[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return max(1, min(sms // tiles if tiles <= sms else 1, (K // 64) // 8))
MAX_SUB_BATCH = 256    # images per pass; larger batches are processed in slices (eval BN is batch-independent)


def _conv_bn_pairs(seq) -> List[Tuple[torch.nn.Conv2d, torch.nn.BatchNorm2d]]:
    mods = list(seq)
    pairs = []
    for i, m in enumerate(mods):
        if isinstance(m, torch.nn.Conv2d):
            bn = mods[i + 1]
            if not isinstance(bn, torch.nn.BatchNorm2d):
                raise _lib.CtkError("conv block layout differs from the reference (Conv2d must be followed by BatchNorm2d)")
            if m.kernel_size != (3, 3) or m.stride != (1, 1) or m.padding != (1, 1):
                raise _lib.CtkError("only 3x3 / stride 1 / pad 1 convolutions are on the hot path")
            pairs.append((m, bn))
    return pairs


def _head_layers(seq):
    lin = [m for m in seq if isinstance(m, torch.nn.Linear)]
    bns = [m for m in seq if isinstance(m, torch.nn.BatchNorm1d)]
    if len(lin) != 3 or len(bns) != 2:
        raise _lib.CtkError("head layout differs from the reference (3 Linear + 2 BatchNorm1d expected)")
    return lin, bns


class _Branch:
    """One conv stack: first block on the fp32 pipe (cin 1 or 2), the rest on tcgen05."""

    def __init__(self, pairs, c_offset: int, training: bool = False):
        self.pairs = pairs
        self.c_offset = c_offset
        self.cin = pairs[0][0].in_channels
        self.channels = [p[0].out_channels for p in pairs]
        self._validate(training)

    def _validate(self, training: bool) -> None:
        """The kernels cover the widths the reference instantiates (train_model.py:535,537: 64 filters per branch, or 128
        filters x 6 blocks) and anything built from the same pieces; say so at construction instead of failing inside
        a kernel launch with an opaque status."""
        supported = ("supported: first Conv2d 1->64 or 2->128 channels; later Conv2d layers with in_channels a multiple of 64 "
                     "and out_channels a multiple of 64 (128 for training) up to 512 -- e.g. "
                     "AdvancedRegressionModel(2, 128, 6) and SimplifiedTwoBranchRegressionModel(64)")
        for li, (conv, bn) in enumerate(self.pairs):
            cin, cout = conv.in_channels, conv.out_channels
            if li == 0:
                ok = (cin, cout) in ((1, 64), (2, 128))
            else:
                ok = cin % 64 == 0 and cout % (128 if training else 64) == 0 and cout <= 512
            if not ok:
                raise _lib.CtkError(f"conv block {li} ({cin}->{cout} channels) has no libctk kernel; {supported}")
            if bn.num_features != cout or not bn.affine or not bn.track_running_stats:
                raise _lib.CtkError(f"conv block {li}: BatchNorm2d must be affine, track running statistics and match the conv")
            if training and bn.momentum is None:
                raise _lib.CtkError("BatchNorm momentum=None (cumulative average) is not supported by the ctk training path; "
                                    "the reference uses the default momentum 0.1")


class InferenceEngine:
    """Eval-mode forward of either reference model through libctk."""

    def __init__(self, model: torch.nn.Module, conv_flags: int = 0, precision: str = "bf16"):
        if precision not in ("bf16", "fp32"):
            raise _lib.CtkError("precision must be 'bf16' (bf16 operands, fp32 accumulate) or 'fp32' (split-bf16, fp32-class)")
        self.model = model
        self.conv_flags = conv_flags
        self.precision = precision
        name = type(model).__name__
        if hasattr(model, "conv_layers") and hasattr(model, "fc_layers"):
            self.kind = "single"
            self.branches = [_Branch(_conv_bn_pairs(model.conv_layers), 0)]
            self.lin, self.bns = _head_layers(model.fc_layers)
            self.sigmoid_half = 0
        elif hasattr(model, "bleed_branch") and hasattr(model, "source_branch"):
            self.kind = "double"
            self.branches = [_Branch(_conv_bn_pairs(model.bleed_branch.conv_blocks), 0),
                             _Branch(_conv_bn_pairs(model.source_branch.conv_blocks), 1)]
            self.lin, self.bns = _head_layers(model.regression_head.fc_layers)
            self.sigmoid_half = 1
        else:
            raise _lib.CtkError(f"{name} is not one of the two crosstalk regression models")
        self.feat_channels = sum(b.channels[-1] for b in self.branches)
        self._cache_key = None
        self._packed: Dict[str, torch.Tensor] = {}
        self._bufs: Dict[Tuple, torch.Tensor] = {}

    # ------------------------------------------------------------------ parameter cache
    def _params_key(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.model.parameters()) + list(self.model.buffers()))

    def _fold(self, bias, bn) -> Tuple[torch.Tensor, torch.Tensor]:
        c = bn.num_features
        scale = torch.empty(c, device=bn.weight.device, dtype=torch.float32)
        shift = torch.empty_like(scale)
        call("ctk_fold_bn_eval", ptr(bias), ptr(bn.weight), ptr(bn.bias), ptr(bn.running_mean), ptr(bn.running_var),
             c_float(bn.eps), c_int(c), ptr(scale), ptr(shift), stream())
        return scale, shift

    def refresh(self, force: bool = False) -> None:
        key = self._params_key()
        if not force and key == self._cache_key:
            return
        dev = next(self.model.parameters()).device
        pk: Dict[str, torch.Tensor] = {}
        for bi, br in enumerate(self.branches):
            for li, (conv, bn) in enumerate(br.pairs):
                _lib.require_device(conv.weight, torch.float32, "conv weight")
                scale, shift = self._fold(conv.bias, bn)
                pk[f"b{bi}.l{li}.scale"], pk[f"b{bi}.l{li}.shift"] = scale, shift
                if li == 0:
                    w = torch.empty(conv.out_channels, conv.in_channels * 9, device=dev, dtype=torch.float32)
                    call("ctk_pack_first_weight", ptr(conv.weight), ptr(scale), c_int(conv.out_channels),
                         c_int(conv.in_channels), ptr(w), stream())
                elif self.precision == "fp32":
                    w = torch.empty(9, conv.out_channels, 3 * conv.in_channels, device=dev, dtype=torch.bfloat16)
                    call("ctk_pack_conv_weight_split_bf16", ptr(conv.weight), c_int(conv.out_channels),
                         c_int(conv.in_channels), ptr(w), stream())
                else:
                    w = torch.empty(9, conv.out_channels, conv.in_channels, device=dev, dtype=torch.bfloat16)
                    call("ctk_pack_conv_weight_bf16", ptr(conv.weight), c_int(conv.out_channels),
                         c_int(conv.in_channels), ptr(w), stream())
                pk[f"b{bi}.l{li}.w"] = w
        fc1, fc2, fc3 = self.lin
        hw = fc1.in_features // self.feat_channels
        w1 = torch.empty(fc1.out_features, fc1.in_features, device=dev, dtype=torch.bfloat16)
        if self.precision == "fp32":
            w1_lo = torch.empty_like(w1)
            call("ctk_pack_fc1_weight_split_bf16", ptr(fc1.weight), c_int(fc1.out_features), c_int(self.feat_channels),
                 c_int(hw), ptr(w1), ptr(w1_lo), stream())
            pk["fc1.w_lo"] = w1_lo
        else:
            call("ctk_pack_fc1_weight_bf16", ptr(fc1.weight), c_int(fc1.out_features), c_int(self.feat_channels), c_int(hw),
                 ptr(w1), stream())
        pk["fc1.w"] = w1
        pk["fc1.scale"], pk["fc1.shift"] = self._fold(fc1.bias, self.bns[0])
        pk["fc2.scale"], pk["fc2.shift"] = self._fold(fc2.bias, self.bns[1])
        self._packed = pk
        self._cache_key = key

    # ------------------------------------------------------------------ buffers
    def _buf(self, tag: str, shape, dtype, dev) -> torch.Tensor:
        key = (tag, tuple(shape), dtype, str(dev))
        b = self._bufs.get(key)
        if b is None:
            b = torch.empty(shape, device=dev, dtype=dtype)
            self._bufs[key] = b
        return b

    # ------------------------------------------------------------------ forward
    def _forward_slice_fp32(self, x: torch.Tensor, out: torch.Tensor, taps: Optional[dict]) -> None:
        """fp32-class pass: activations as bf16 (hi, lo) pairs, three MMAs per product (include/ctk.h, split entry points)."""
        n, c_total, H, W = x.shape
        dev = x.device
        pk = self._packed
        depth = len(self.branches[0].pairs)
        hf, wf = H >> depth, W >> depth
        m_pad = (n + 127) // 128 * 128
        feat = self._buf("feat_split", (2, m_pad, hf, wf, self.feat_channels), torch.bfloat16, dev)
        c_out_off = 0
        for bi, br in enumerate(self.branches):
            h, w = H, W
            cur = None
            for li, (conv, bn) in enumerate(br.pairs):
                cout = conv.out_channels
                last = li == len(br.pairs) - 1
                if last:
                    hi, lo, cstride, coff = feat[0], feat[1], self.feat_channels, c_out_off
                else:
                    dst = self._buf(f"act_split{li}", (n, h // 2, w // 2, 2 * cout), torch.bfloat16, dev)
                    hi, lo, cstride, coff = dst, dst[..., cout:], 2 * cout, 0
                if li == 0:
                    call("ctk_conv_first_eval_split", ptr(x), c_int(n), c_int(c_total), c_int(br.c_offset), c_int(br.cin),
                         c_int(h), c_int(w), ptr(pk[f"b{bi}.l0.w"]), ptr(pk[f"b{bi}.l0.shift"]), c_int(cout),
                         c_float(LEAKY_SLOPE), ptr(hi), ptr(lo), c_int(cstride), c_int(coff), stream())
                else:
                    call("ctk_conv3x3_tc_eval_split", ptr(cur), c_int(n), c_int(h), c_int(w), c_int(conv.in_channels),
                         ptr(pk[f"b{bi}.l{li}.w"]), c_int(cout), ptr(pk[f"b{bi}.l{li}.scale"]),
                         ptr(pk[f"b{bi}.l{li}.shift"]), c_float(LEAKY_SLOPE), ptr(hi), ptr(lo), c_int(cstride), c_int(coff),
                         stream(), meta={"flops": 3 * 2.0 * n * h * w * cout * 9 * conv.in_channels})
                if taps is not None and not last:
                    taps[f"b{bi}.l{li}"] = dst[..., :cout].float() + dst[..., cout:].float()
                cur = None if last else dst
                h, w = h // 2, w // 2
            c_out_off += br.channels[-1]
        if taps is not None:
            taps["feat"] = feat[0, :n].float() + feat[1, :n].float()
        fc1, fc2, fc3 = self.lin
        K = fc1.in_features
        tiles = (m_pad // 128) * (fc1.out_features // 128)
        splits = fc1_splits(tiles, K, dev)
        partial = self._buf("fc1p_split", (3 * splits, m_pad, fc1.out_features), torch.float32, dev)
        for j, (a, b) in enumerate(((feat[0], pk["fc1.w"]), (feat[1], pk["fc1.w"]), (feat[0], pk["fc1.w_lo"]))):
            call("ctk_gemm_bf16_splitk", ptr(a), ptr(b), c_int(m_pad), c_int(fc1.out_features), c_int(K), c_int(splits),
                 ptr(partial[j * splits:]), stream())
        call("ctk_head_eval", ptr(partial), c_int(3 * splits), c_int(m_pad), c_int(n), c_int(fc1.out_features),
             c_int(fc2.out_features), ptr(pk["fc1.scale"]), ptr(pk["fc1.shift"]), ptr(fc2.weight), ptr(pk["fc2.scale"]),
             ptr(pk["fc2.shift"]), ptr(fc3.weight), ptr(fc3.bias), c_float(LEAKY_SLOPE), c_int(self.sigmoid_half),
             ptr(out), stream())

    def _forward_slice(self, x: torch.Tensor, out: torch.Tensor, taps: Optional[dict]) -> None:
        if self.precision == "fp32":
            return self._forward_slice_fp32(x, out, taps)
        n, c_total, H, W = x.shape
        dev = x.device
        pk = self._packed
        depth = len(self.branches[0].pairs)
        hf, wf = H >> depth, W >> depth
        m_pad = (n + 127) // 128 * 128
        feat = self._buf("feat", (m_pad, hf, wf, self.feat_channels), torch.bfloat16, dev)
        c_out_off = 0
        for bi, br in enumerate(self.branches):
            h, w = H, W
            cur = None
            for li, (conv, bn) in enumerate(br.pairs):
                cout = conv.out_channels
                last = li == len(br.pairs) - 1
                if last:
                    dst, cstride, coff = feat, self.feat_channels, c_out_off
                else:
                    dst, cstride, coff = self._buf(f"act{li}", (n, h // 2, w // 2, cout), torch.bfloat16, dev), cout, 0
                if li == 0:
                    call("ctk_conv_first_eval", ptr(x), c_int(n), c_int(c_total), c_int(br.c_offset), c_int(br.cin),
                         c_int(h), c_int(w), ptr(pk[f"b{bi}.l0.w"]), ptr(pk[f"b{bi}.l0.shift"]), c_int(cout),
                         c_float(LEAKY_SLOPE), ptr(dst), c_int(cstride), c_int(coff), stream())
                else:
                    call("ctk_conv3x3_tc_eval", ptr(cur), c_int(n), c_int(h), c_int(w), c_int(conv.in_channels),
                         ptr(pk[f"b{bi}.l{li}.w"]), c_int(cout), ptr(pk[f"b{bi}.l{li}.scale"]),
                         ptr(pk[f"b{bi}.l{li}.shift"]), c_float(LEAKY_SLOPE), ptr(dst), c_int(cstride), c_int(coff),
                         c_int(self.conv_flags), stream(),
                         meta={"flops": 2.0 * n * h * w * cout * 9 * conv.in_channels})
                if taps is not None and not last:
                    taps[f"b{bi}.l{li}"] = dst.clone()
                cur = dst
                h, w = h // 2, w // 2
            c_out_off += br.channels[-1]
        if taps is not None:
            taps["feat"] = feat[:n].clone()
        fc1, fc2, fc3 = self.lin
        K = fc1.in_features
        tiles = (m_pad // 128) * (fc1.out_features // 128)
        splits = fc1_splits(tiles, K, dev)
        partial = self._buf("fc1p", (splits, m_pad, fc1.out_features), torch.float32, dev)
        call("ctk_gemm_bf16_splitk", ptr(feat), ptr(pk["fc1.w"]), c_int(m_pad), c_int(fc1.out_features), c_int(K),
             c_int(splits), ptr(partial), stream())
        call("ctk_head_eval", ptr(partial), c_int(splits), c_int(m_pad), c_int(n), c_int(fc1.out_features),
             c_int(fc2.out_features), ptr(pk["fc1.scale"]), ptr(pk["fc1.shift"]), ptr(fc2.weight), ptr(pk["fc2.scale"]),
             ptr(pk["fc2.shift"]), ptr(fc3.weight), ptr(fc3.bias), c_float(LEAKY_SLOPE), c_int(self.sigmoid_half),
             ptr(out), stream())
        if taps is not None:
            taps["fc1_partial"] = partial[:, :n].clone()

    @torch.no_grad()
    def forward(self, x: torch.Tensor, taps: Optional[dict] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        _lib.require_device(x, torch.float32, "input batch")
        if x.dim() != 4 or x.shape[1] != 2 or x.shape[2] % 32 or x.shape[3] % 32:
            raise _lib.CtkError(f"input must be [N,2,H,W] float32 with H, W multiples of 32, got {tuple(x.shape)}")
        call("ctk_device_check")
        self.refresh()
        n = x.shape[0]
        if out is None:
            out = torch.empty(n, 1, device=x.device, dtype=torch.float32)
        else:
            _lib.require_device(out, torch.float32, "out")
            if out.numel() != n:
                raise _lib.CtkError("out must hold one float32 per tile")
        for s in range(0, n, MAX_SUB_BATCH):
            e = min(n, s + MAX_SUB_BATCH)
            self._forward_slice(x[s:e], out[s:e], taps)
        return out
