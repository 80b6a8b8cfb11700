"""fp32 training path of both reference models: the reference's own arithmetic (float32 operands and tensors,
train_model.py:419-424) on the CUDA cores, selected with ``ctk.set_precision(model, "fp32")``.

Same contract as ``TrainEngine`` (forward returns (scores, saved state), backward returns every parameter's gradient;
``on_grad_ready`` / ``finalize_grads`` hooks for the data-parallel exchange), different kernels: ``ctk_conv3x3_f32`` /
``ctk_conv3x3_wgrad_f32`` (FFMA implicit GEMMs with fp64 running sums), fp64 batch statistics, fp32 BatchNorm / LeakyReLU /
MaxPool passes, ``ctk_gemm_f32`` for FC1 in the reference's own weight layout, and the fp32 head kernels the bf16 path
already uses.  Activations are NHWC float32; the input's NCHW planes are read in place; the last block writes its pooled
output in nn.Flatten's order, so FC1 needs no permuted weight copy.  Every reduction is a fixed-order two-stage sum: two
identical steps are bit-identical.  About 1/20 of the tensor-core path's throughput -- this is the mode parity runs use.
"""
from __future__ import annotations

from ctypes import c_double, c_float, c_int, c_longlong
from typing import Dict, Tuple

import torch

from . import _lib
from ._lib import call, ptr, stream, workspace
from .engine import LEAKY_SLOPE
from .train import TrainEngine


def _ll(v) -> c_longlong:
    return c_longlong(int(v))


class TrainEngineF32(TrainEngine):
    precision = "fp32"

    def __init__(self, model: torch.nn.Module):
        super().__init__(model)
        if any(bn.momentum is None for br in self.branches for _, bn in br.pairs):
            raise _lib.CtkError("BatchNorm momentum=None is not supported")

    # ------------------------------------------------------------------ helpers
    def _stats(self, y: torch.Tensor, pixels: int, c: int) -> torch.Tensor:
        sums = self._new((2 * c,), torch.float64, y.device)
        ws = workspace("ctk_channel_sums_f64_workspace_bytes", c, device=y.device)
        call("ctk_channel_stats_f32", ptr(y), _ll(pixels), c_int(c), ptr(sums), ws[1], ws[2], stream())
        return sums

    def _finalize64(self, sums, count, bias, bn, dev):
        c = bn.num_features
        scale, shift, mean, invstd = (self._new((c,), torch.float32, dev) for _ in range(4))
        call("ctk_bn_finalize_f64", ptr(sums), c_double(count), ptr(bias), ptr(bn.weight), ptr(bn.bias), ptr(bn.running_mean),
             ptr(bn.running_var), ptr(bn.num_batches_tracked), c_float(bn.momentum), c_float(bn.eps), c_int(c), ptr(scale),
             ptr(shift), ptr(mean), ptr(invstd), stream())
        return scale, shift, mean, invstd

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, dict]:
        _lib.require_device(x, torch.float32, "input batch")
        if x.dim() != 4 or x.shape[1] != 2 or x.shape[2] % 32 or x.shape[3] % 32:
            raise _lib.CtkError(f"input must be [N,2,H,W] float32 with H, W multiples of 32, got {tuple(x.shape)}")
        call("ctk_device_check")
        for p in self.params:
            _lib.require_device(p, torch.float32, "parameter")
        if self.stat_allreduce is not None:
            raise _lib.CtkError("sync_bn is not wired into the fp32 training path (single-process parity mode)")
        n, c_total, H, W = x.shape
        if n < 2:
            # nn.BatchNorm1d in train() refuses a single sample (torch.nn.functional._verify_batch_size); so does this path
            raise ValueError(f"Expected more than 1 value per channel when training, got input size torch.Size([{n}, "
                             f"{self.lin[0].out_features}])")
        dev = x.device
        depth = len(self.branches[0].pairs)
        hf, wf = H >> depth, W >> depth
        hw = hf * wf
        fc1, fc2, fc3 = self.lin
        K = fc1.in_features
        if K != hw * self.feat_channels:
            raise _lib.CtkError("input size does not match the model's first Linear layer")
        sv = {"x": x, "n": n, "H": H, "W": W, "branches": []}
        feat = self._new((n, K), torch.float32, dev)                 # nn.Flatten order: [n][c * hw + p]
        c_off = 0
        for br in self.branches:
            h, w = H, W
            # the block input as an [n][y][x][c] view: the reference's NCHW planes first, NHWC activations afterwards
            cur, strides = x, (c_total * H * W, W, 1, H * W)
            cur_off = br.c_offset * H * W
            blocks = []
            for li, (conv, bn) in enumerate(br.pairs):
                cout, cin = conv.out_channels, conv.in_channels
                last = li == len(br.pairs) - 1
                wk = self._new((9 * cin, cout), torch.float32, dev)
                call("ctk_pack_conv_weight_f32", ptr(conv.weight), c_int(cout), c_int(cin), c_int(0), ptr(wk), stream())
                y = self._new((n, h, w, cout), torch.float32, dev)
                xin = cur.view(-1)[cur_off:] if cur_off else cur
                call("ctk_conv3x3_f32", ptr(xin), _ll(strides[0]), _ll(strides[1]), _ll(strides[2]), _ll(strides[3]), c_int(n),
                     c_int(h), c_int(w), c_int(cin), ptr(wk), c_int(cout), ptr(y), stream(),
                     meta={"flops": 2.0 * n * h * w * cout * 9 * cin})
                sums = self._stats(y, n * h * w, cout)
                scale, shift, mean, invstd = self._finalize64(sums, float(n) * h * w, conv.bias, bn, dev)
                hp, wp = h // 2, w // 2
                if last:
                    dst, ostr, ooff = feat, (K, 1, hw), c_off * hw
                else:
                    dst, ostr, ooff = self._new((n, hp, wp, cout), torch.float32, dev), (hp * wp * cout, cout, 1), 0
                dview = dst.view(-1)[ooff:] if ooff else dst
                call("ctk_bn_act_pool_fwd_f32", ptr(y), c_int(n), c_int(h), c_int(w), c_int(cout), ptr(mean), ptr(invstd),
                     ptr(bn.weight), ptr(bn.bias), c_float(LEAKY_SLOPE), ptr(dview), _ll(ostr[0]), _ll(ostr[1]), _ll(ostr[2]),
                     stream())
                blocks.append({"y": y, "x_in": cur, "x_off": cur_off, "x_strides": strides, "pool_strides": ostr,
                               "pool_off": ooff, "mean": mean, "invstd": invstd, "h": h, "w": w, "conv": conv, "bn": bn})
                cur, strides, cur_off = dst, (hp * wp * cout, wp * cout, cout, 1), 0
                h, w = hp, wp
            sv["branches"].append({"branch": br, "blocks": blocks, "c_off": c_off})
            c_off += br.channels[-1]
        sv["feat"], sv["hw"] = feat, hw
        # ---- FC1 in the reference's own layout, then the fp32 head (the same kernels as the bf16 path)
        f1, f2 = fc1.out_features, fc2.out_features
        z1 = self._new((n, f1), torch.float32, dev)
        call("ctk_gemm_f32", ptr(feat), _ll(K), _ll(1), ptr(fc1.weight), _ll(K), _ll(1), ptr(fc1.bias), c_int(n), c_int(f1),
             c_int(K), ptr(z1), _ll(f1), stream(), meta={"flops": 2.0 * n * f1 * K})
        st1 = self._new((2 * f1,), torch.float32, dev)
        call("ctk_colstat", ptr(z1), c_int(1), c_longlong(0), c_int(f1), ptr(None), c_int(n), c_int(f1), ptr(None), ptr(st1),
             stream())
        bn1 = self._bn_finalize(st1, float(n), None, self.bns[0], dev)
        masks = self._masks(n, f1, f2, dev)
        a1 = self._new((n, f1), torch.float32, dev)
        call("ctk_bn1d_act_drop_fwd", ptr(z1), ptr(bn1[0]), ptr(bn1[1]), ptr(masks[0]), c_float(self.drop_p[0]),
             c_float(LEAKY_SLOPE), c_int(n), c_int(f1), ptr(a1), stream())
        z2 = self._new((n, f2), torch.float32, dev)
        call("ctk_gemm_f32", ptr(a1), _ll(f1), _ll(1), ptr(fc2.weight), _ll(f1), _ll(1), ptr(fc2.bias), c_int(n), c_int(f2),
             c_int(f1), ptr(z2), _ll(f2), stream())
        st2 = self._new((2 * f2,), torch.float32, dev)
        call("ctk_colstat", ptr(z2), c_int(1), c_longlong(0), c_int(f2), ptr(None), c_int(n), c_int(f2), ptr(None), ptr(st2),
             stream())
        bn2 = self._bn_finalize(st2, float(n), None, self.bns[1], dev)
        a2 = self._new((n, f2), torch.float32, dev)
        call("ctk_bn1d_act_drop_fwd", ptr(z2), ptr(bn2[0]), ptr(bn2[1]), ptr(masks[1]), c_float(self.drop_p[1]),
             c_float(LEAKY_SLOPE), c_int(n), c_int(f2), ptr(a2), stream())
        out = self._new((n, 1), torch.float32, dev)
        call("ctk_head_out_fwd", ptr(a2), ptr(fc3.weight), ptr(fc3.bias), c_int(n), c_int(f2), c_int(self.sigmoid_half),
             ptr(out), stream())
        sv.update(z1=z1, bn1=bn1, a1=a1, z2=z2, bn2=bn2, a2=a2, out=out.detach(), masks=masks)
        torch.autograd.graph.increment_version([b for b in self.model.buffers()])
        return out, sv

    # ------------------------------------------------------------------ backward
    def _backward(self, sv: dict, dout: torch.Tensor) -> Dict[torch.nn.Parameter, torch.Tensor]:
        grads: Dict[torch.nn.Parameter, torch.Tensor] = {}

        def done(p, g):
            grads[p] = g
            if self.on_grad_ready is not None:
                self.on_grad_ready(p, g)

        n, dev = sv["n"], dout.device
        fc1, fc2, fc3 = self.lin
        f1, f2 = fc1.out_features, fc2.out_features
        K = fc1.in_features
        masks = sv["masks"]
        # ---- head (fp32 kernels shared with the bf16 path)
        da2 = self._new((n, f2), torch.float32, dev)
        dw3 = self._new((1, f2), torch.float32, dev)
        db3 = self._new((1,), torch.float32, dev)
        call("ctk_head_out_bwd", ptr(dout), ptr(sv["out"]), ptr(sv["a2"]), ptr(fc3.weight), c_int(n), c_int(f2),
             c_int(self.sigmoid_half), ptr(da2), ptr(dw3), ptr(db3), stream())
        done(fc3.weight, dw3)
        done(fc3.bias, db3)
        sc2, sh2, mu2, is2 = sv["bn2"]
        dact2 = self._new((n, f2), torch.float32, dev)
        sums2 = self._new((2 * f2,), torch.float32, dev)
        call("ctk_bn1d_bwd_reduce", ptr(da2), ptr(masks[1]), c_float(self.drop_p[1]), ptr(sv["z2"]), ptr(sc2), ptr(sh2),
             ptr(mu2), ptr(is2), c_float(LEAKY_SLOPE), c_int(n), c_int(f2), ptr(dact2), ptr(sums2), stream())
        done(self.bns[1].bias, sums2[:f2])
        done(self.bns[1].weight, sums2[f2:])
        dz2 = self._new((n, f2), torch.float32, dev)
        call("ctk_bn1d_bwd_apply", ptr(dact2), ptr(sv["z2"]), ptr(sc2), ptr(mu2), ptr(is2), ptr(sums2), c_int(n), c_int(f2),
             ptr(dz2), ptr(None), ptr(None), c_int(0), stream())
        dw2 = self._new((f2, f1), torch.float32, dev)
        call("ctk_gemm_f32", ptr(dz2), _ll(1), _ll(f2), ptr(sv["a1"]), _ll(1), _ll(f1), ptr(None), c_int(f2), c_int(f1),
             c_int(n), ptr(dw2), _ll(f1), stream())
        done(fc2.weight, dw2)
        done(fc2.bias, self._colsum(dz2, n, f2))
        da1 = self._new((n, f1), torch.float32, dev)
        call("ctk_gemm_f32", ptr(dz2), _ll(f2), _ll(1), ptr(fc2.weight), _ll(1), _ll(f1), ptr(None), c_int(n), c_int(f1),
             c_int(f2), ptr(da1), _ll(f1), stream())
        sc1, sh1, mu1, is1 = sv["bn1"]
        dact1 = self._new((n, f1), torch.float32, dev)
        sums1 = self._new((2 * f1,), torch.float32, dev)
        call("ctk_bn1d_bwd_reduce", ptr(da1), ptr(masks[0]), c_float(self.drop_p[0]), ptr(sv["z1"]), ptr(sc1), ptr(sh1),
             ptr(mu1), ptr(is1), c_float(LEAKY_SLOPE), c_int(n), c_int(f1), ptr(dact1), ptr(sums1), stream())
        done(self.bns[0].bias, sums1[:f1])
        done(self.bns[0].weight, sums1[f1:])
        dz1 = self._new((n, f1), torch.float32, dev)
        call("ctk_bn1d_bwd_apply", ptr(dact1), ptr(sv["z1"]), ptr(sc1), ptr(mu1), ptr(is1), ptr(sums1), c_int(n), c_int(f1),
             ptr(dz1), ptr(None), ptr(None), c_int(0), stream())
        done(fc1.bias, self._colsum(dz1, n, f1))
        # ---- FC1: dW1[f][k] = sum_n dz1[n][f] feat[n][k];  dfeat[n][k] = sum_f dz1[n][f] W1[f][k]  (reference layouts)
        feat, hw = sv["feat"], sv["hw"]
        dw1 = self._new((f1, K), torch.float32, dev)
        call("ctk_gemm_f32", ptr(dz1), _ll(1), _ll(f1), ptr(feat), _ll(1), _ll(K), ptr(None), c_int(f1), c_int(K), c_int(n),
             ptr(dw1), _ll(K), stream(), meta={"flops": 2.0 * n * f1 * K})
        done(fc1.weight, dw1)
        dfeat = self._new((n, K), torch.float32, dev)
        call("ctk_gemm_f32", ptr(dz1), _ll(f1), _ll(1), ptr(fc1.weight), _ll(1), _ll(K), ptr(None), c_int(n), c_int(K),
             c_int(f1), ptr(dfeat), _ll(K), stream(), meta={"flops": 2.0 * n * f1 * K})
        # ---- conv stacks, last block first
        for entry in sv["branches"]:
            blocks = entry["blocks"]
            dp, dstr, doff = dfeat, (K, 1, hw), entry["c_off"] * hw
            for li in range(len(blocks) - 1, -1, -1):
                b = blocks[li]
                conv, bn, h, w = b["conv"], b["bn"], b["h"], b["w"]
                cout, cin = conv.out_channels, conv.in_channels
                dpv = dp.view(-1)[doff:] if doff else dp
                sums64 = self._new((2 * cout,), torch.float64, dev)
                sums32 = self._new((2 * cout,), torch.float32, dev)
                ws = workspace("ctk_channel_sums_f64_workspace_bytes", cout, device=dev)
                call("ctk_bn_bwd_reduce_f32", ptr(b["y"]), ptr(dpv), _ll(dstr[0]), _ll(dstr[1]), _ll(dstr[2]), c_int(n), c_int(h),
                     c_int(w), c_int(cout), ptr(b["mean"]), ptr(b["invstd"]), ptr(bn.weight), ptr(bn.bias),
                     c_float(LEAKY_SLOPE), ptr(sums64), ptr(sums32), ws[1], ws[2], stream())
                done(bn.bias, sums32[:cout])
                done(bn.weight, sums32[cout:])
                dy = self._new((n, h, w, cout), torch.float32, dev)
                call("ctk_bn_bwd_apply_f32", ptr(b["y"]), ptr(dpv), _ll(dstr[0]), _ll(dstr[1]), _ll(dstr[2]), c_int(n), c_int(h),
                     c_int(w), c_int(cout), ptr(b["mean"]), ptr(b["invstd"]), ptr(bn.weight), ptr(bn.bias), ptr(sums64),
                     c_double(float(n) * h * w), c_float(LEAKY_SLOPE), ptr(dy), stream())
                b["y"] = None
                xs = b["x_strides"]
                xin = b["x_in"].view(-1)[b["x_off"]:] if b["x_off"] else b["x_in"]
                dw = self._new(tuple(conv.weight.shape), torch.float32, dev)
                ws = workspace("ctk_conv3x3_wgrad_f32_workspace_bytes", n, h, w, cin, cout, device=dev)
                call("ctk_conv3x3_wgrad_f32", ptr(dy), ptr(xin), _ll(xs[0]), _ll(xs[1]), _ll(xs[2]), _ll(xs[3]), c_int(n), c_int(h),
                     c_int(w), c_int(cin), c_int(cout), ptr(dw), ws[1], ws[2], stream(),
                     meta={"flops": 2.0 * n * h * w * cout * 9 * cin})
                done(conv.weight, dw)
                # the conv bias feeds a train-mode BatchNorm, so its gradient is sum(dY) = 0 identically
                done(conv.bias, self._zero_grad_of(conv.bias))
                if li > 0:
                    wg = self._new((9 * cout, cin), torch.float32, dev)
                    call("ctk_pack_conv_weight_f32", ptr(conv.weight), c_int(cout), c_int(cin), c_int(1), ptr(wg), stream())
                    dx = self._new((n, h, w, cin), torch.float32, dev)
                    call("ctk_conv3x3_f32", ptr(dy), _ll(h * w * cout), _ll(w * cout), _ll(cout), _ll(1), c_int(n), c_int(h),
                         c_int(w), c_int(cout), ptr(wg), c_int(cin), ptr(dx), stream(),
                         meta={"flops": 2.0 * n * h * w * cout * 9 * cin})
                    dp, dstr, doff = dx, (h * w * cin, cin, 1), 0
                del dy
        if self.finalize_grads is not None:
            grads = self.finalize_grads(grads)
        return grads
