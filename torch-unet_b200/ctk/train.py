"""Train-mode forward and backward of both reference models through libctk.

``TrainEngine.forward`` runs conv (raw + batch statistics) -> BN finalise -> normalise/LeakyReLU/pool per block, the
FC1 split-K GEMM and the small fp32 head; ``TrainEngine.backward`` walks the same graph in reverse with the dedicated
kernels (BN backward reduce/apply, tcgen05 dgrad and wgrad, FC1 dX/dW GEMMs).  ``_CtkTrainFunction`` plugs the pair into
``torch.autograd`` so the reference loop (train_model.py:419-424: zero_grad / model(x) / criterion / backward / step)
runs unchanged; PyTorch only owns the tensors and the graph edge.
"""
from __future__ import annotations

import contextlib
import os
from ctypes import c_double, c_float, c_int, c_longlong, c_size_t, c_ulonglong, c_void_p
from typing import Callable, Dict, List, Optional, Tuple

import torch

from . import _lib
from ._lib import call, ptr, stream, workspace
from .engine import LEAKY_SLOPE, _Branch, _conv_bn_pairs, _head_layers, fc1_splits


def _pad(n: int, m: int) -> int:
    return (n + m - 1) // m * m


class TrainEngine:
    def __init__(self, model: torch.nn.Module):
        self.model = model
        if hasattr(model, "conv_layers") and hasattr(model, "fc_layers"):
            self.kind = "single"
            self.branches = [_Branch(_conv_bn_pairs(model.conv_layers), 0, training=True)]
            seq = model.fc_layers
            self.sigmoid_half = 0
        elif hasattr(model, "bleed_branch") and hasattr(model, "source_branch"):
            self.kind = "double"
            self.branches = [_Branch(_conv_bn_pairs(model.bleed_branch.conv_blocks), 0, training=True),
                             _Branch(_conv_bn_pairs(model.source_branch.conv_blocks), 1, training=True)]
            seq = model.regression_head.fc_layers
            self.sigmoid_half = 1
        else:
            raise _lib.CtkError(f"{type(model).__name__} is not one of the two crosstalk regression models")
        self.lin, self.bns = _head_layers(seq)
        if any(bn.momentum is None for bn in self.bns):
            raise _lib.CtkError("BatchNorm momentum=None (cumulative average) is not supported by the ctk training path")
        self.drop_p = [m.p for m in seq if isinstance(m, torch.nn.Dropout)]
        if len(self.drop_p) != 2:
            raise _lib.CtkError("head layout differs from the reference (2 Dropout layers expected)")
        self.feat_channels = sum(b.channels[-1] for b in self.branches)
        self.params: List[torch.nn.Parameter] = list(model.parameters())
        # "gram": first block's BN statistics and weight gradient from the input's patch Gram matrix (no full-resolution
        # activation is stored); "stored": the generic path (raw conv output kept, generic BN backward + wgrad)
        self.first_block_mode = "gram"
        self.forced_masks: Optional[Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]] = None
        self.on_grad_ready: Optional[Callable[[torch.nn.Parameter, torch.Tensor], None]] = None
        self.finalize_grads: Optional[Callable[[dict], dict]] = None      # e.g. parallel.GradSynchronizer.finalize
        # SyncBN (parallel.attach(..., sync_bn=True)): in-place SUM over ranks of a small statistics tensor, and the number
        # of ranks.  The reference is one process, so its BatchNorm statistics span the whole global batch; with this
        # hook the data-parallel path reproduces that instead of per-rank statistics.
        self.stat_allreduce: Optional[Callable[[torch.Tensor], None]] = None
        self.stat_world: int = 1
        # SyncBN forms global statistics as (sum over ranks) / (n * world): every rank must hold the same number of tiles.
        # parallel.attach installs a check (one 16-byte all-reduce per forward) that raises instead of silently normalising
        # with the wrong count when shards are ragged.
        self.stat_check_equal_batch: Optional[Callable[[int, torch.device], None]] = None
        # Experiment: every branch's conv stack on its own stream and every wgrad on a further side stream, so that the
        # HBM-bound BatchNorm passes of one branch / layer could overlap the tensor-bound conv kernels of the other.
        # Measured (gpurun_out/ab_overlap.txt, round 2): 15.08 ms per step against 14.94 ms -- one persistent CTA per SM with
        # 210 - 227 KB of shared memory leaves no room for a second kernel.  Off by default; see DESIGN.md section 4.8.
        self.overlap_streams: bool = False
        # Side-stream schedules, both OFF by default (CTK_OVERLAP_WGRAD = "0" (default) / "pack" / "wgrad" / "1" = both).
        # Measured on B200 in round 2 (tools/r2_run21.sh, r2_run22.sh, r2_run26.sh; same box, 20 - 30 steps each):
        #  * overlap_pack  -- the step's bf16 weight copies (conv forward / dgrad layouts, FC1: 0.35 ms, 805 MB of traffic)
        #    built on a side stream beside the first block: 15.00 / 15.24 ms per step against 14.93 / 15.05 ms single-stream.
        #    The Gram and first-conv kernels it runs beside hold the whole register file of an SM, so the packing kernels
        #    mostly wait for them anyway.
        #  * overlap_wgrad -- weight gradients on ONE high-priority side stream, each started behind the input gradient of
        #    its layer, so that wgrad_tc_kernel (192 threads x 48 registers, 166 KB of shared memory, one CTA per SM) could
        #    share its SMs with the HBM-bound BatchNorm-backward passes of the next layer down.  The two then take exactly
        #    the SUM of their solo times (wgrad 2.8 -> 5.0 ms of event time per step, BatchNorm passes 2.3 -> 4.1 ms; step
        #    15.2 - 15.3 ms against 15.3 - 15.4 ms), with or without stream priority, 128-thread BatchNorm CTAs or a maximum
        #    shared-memory carve-out on the streaming kernels: wgrad_tc_kernel itself keeps 50 - 65 % of the DRAM bandwidth
        #    busy and the pair is bound by it.
        # Results are bit-identical under every schedule (test_stream_overlap_gives_the_same_step).
        mode = os.environ.get("CTK_OVERLAP_WGRAD", "0")
        self.overlap_wgrad: bool = mode in ("1", "wgrad")
        self.overlap_pack: bool = mode in ("1", "pack")
        # data parallel: SMs the persistent tensor-core kernels of the BACKWARD pass leave free for the gradient all-reduce
        # that runs beside them (parallel.attach sets it; 0 = fill the GPU, the single-GPU setting)
        self.backward_sm_reserve: int = 0
        self._zero_grads: Dict[torch.nn.Parameter, torch.Tensor] = {}
        self.dropout_seed: int = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
        self._dropout_calls: int = 0
        self._streams: Dict[Tuple[str, int, int], torch.cuda.Stream] = {}

    # ------------------------------------------------------------------ helpers
    def _side_stream(self, kind: str, index: int, dev) -> torch.cuda.Stream:
        key = (kind, index, dev.index if dev.index is not None else torch.cuda.current_device())
        st = self._streams.get(key)
        if st is None:
            # the weight-gradient stream outranks the default stream: its one-CTA-per-SM kernels take their SMs first and the
            # streaming passes fill in beside them
            st = self._streams[key] = torch.cuda.Stream(device=dev, priority=-1 if kind == "wgrad" and os.environ.get("CTK_WGRAD_PRIORITY", "1") != "0" else 0)
        return st

    @staticmethod
    def _join(main: torch.cuda.Stream, side_streams) -> None:
        """``main`` waits for everything enqueued so far on every side stream."""
        for st in side_streams:
            ev = torch.cuda.Event()
            ev.record(st)
            main.wait_event(ev)

    @staticmethod
    def _new(shape, dtype, dev):
        return torch.empty(shape, device=dev, dtype=dtype)

    @staticmethod
    def _padded(shape, used, padded, dev, dtype=torch.bfloat16):
        """A GEMM operand whose leading / trailing extent is padded from ``used`` to ``padded``: zero-filled only when
        there is padding to zero (the kernels overwrite the used part)."""
        return torch.empty(shape, device=dev, dtype=dtype) if used == padded else torch.zeros(shape, device=dev, dtype=dtype)

    def _zero_grad_of(self, p: torch.nn.Parameter) -> torch.Tensor:
        """The (identically zero) gradient of a conv bias that feeds a train-mode BatchNorm: one cached tensor per bias
        instead of a fill kernel per layer and step."""
        z = self._zero_grads.get(p)
        if z is None or z.device != p.device:
            z = self._zero_grads[p] = torch.zeros_like(p)
        return z

    def _bn_finalize(self, sums, count, bias, bn, dev, moments=False):
        c = bn.num_features
        scale, shift, mean, invstd = (self._new((c,), torch.float32, dev) for _ in range(4))
        mom = bn.momentum
        track = bn.track_running_stats and bn.running_mean is not None
        call("ctk_bn_finalize_moments" if moments else "ctk_bn_finalize", ptr(sums), c_double(count), ptr(bias), ptr(bn.weight), ptr(bn.bias),
             ptr(bn.running_mean if track else None), ptr(bn.running_var if track else None),
             ptr(bn.num_batches_tracked if track else None), c_float(mom), c_float(bn.eps), c_int(c), ptr(scale), ptr(shift),
             ptr(mean), ptr(invstd), stream())
        return scale, shift, mean, invstd

    def _sync_stats(self, t: torch.Tensor) -> torch.Tensor:
        if self.stat_allreduce is not None:
            self.stat_allreduce(t)
        return t

    def _global_sums(self, sums: torch.Tensor) -> torch.Tensor:
        """BN-backward reductions for the apply pass.  The kernels divide by the LOCAL element count, so under SyncBN they
        are handed (sum over ranks) / world: that is the mean over the global batch.  The local sums stay the parameter
        gradients (the gradient exchange averages them like every other gradient)."""
        if self.stat_allreduce is None:
            return sums
        g = sums.clone()
        self.stat_allreduce(g)
        return g.div_(self.stat_world)

    def _colsum(self, t, n, f):
        st = self._new((2 * f,), torch.float32, t.device)
        call("ctk_colstat", ptr(t), c_int(1), c_longlong(0), c_int(f), ptr(None), c_int(n), c_int(f), ptr(None), ptr(st), stream())
        return st[:f]

    def _pack_weights(self, dev, main: torch.cuda.Stream, params_ready: Optional[torch.cuda.Event] = None) -> dict:
        """bf16 operand copies of the weights for this step -- every tensor-core conv's forward and input-gradient layouts
        and FC1's column-permuted matrix -- built on a side stream at the start of the forward pass, so that the 0.3 ms
        they take (FC1: 805 MB of traffic) run beside the first block instead of on the critical path.  The compute stream
        waits for ``conv_ready`` before its first tensor-core conv and for ``fc1_ready`` before the FC1 GEMM."""
        side = self._side_stream("pack", 0, dev) if self.overlap_pack else None
        packs = {"conv": {}, "conv_ready": None, "fc1_ready": None}
        if side is not None:
            if params_ready is None:
                params_ready = torch.cuda.Event()
                params_ready.record(main)
            side.wait_event(params_ready)       # the parameters are final behind this point (e.g. the optimizer's update)
        with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
            convs = [conv for br in self.branches for li, (conv, _) in enumerate(br.pairs) if li > 0]
            for conv in convs:
                cout, cin = conv.out_channels, conv.in_channels
                packs["conv"][conv] = (self._new((9, cout, cin), torch.bfloat16, dev), self._new((9, cin, cout), torch.bfloat16, dev))
            for g0 in range(0, len(convs), 8):                  # forward and dgrad layouts of eight layers per launch
                grp = convs[g0:g0 + 8]
                k = len(grp)
                call("ctk_pack_conv_weights_train", c_int(k),
                     (c_void_p * k)(*[conv.weight.data_ptr() for conv in grp]),
                     (c_int * k)(*[conv.out_channels for conv in grp]), (c_int * k)(*[conv.in_channels for conv in grp]),
                     (c_void_p * k)(*[packs["conv"][conv][0].data_ptr() for conv in grp]),
                     (c_void_p * k)(*[packs["conv"][conv][1].data_ptr() for conv in grp]), stream())
            if side is not None:
                packs["conv_ready"] = torch.cuda.Event()
                packs["conv_ready"].record(side)
            fc1 = self.lin[0]
            hw = fc1.in_features // self.feat_channels
            w1p = self._new((fc1.out_features, fc1.in_features), torch.bfloat16, dev)
            call("ctk_pack_fc1_weight_bf16", ptr(fc1.weight), c_int(fc1.out_features), c_int(self.feat_channels), c_int(hw),
                 ptr(w1p), stream())
            packs["w1p"] = w1p
            if side is not None:
                packs["fc1_ready"] = torch.cuda.Event()
                packs["fc1_ready"].record(side)
        if side is not None:
            # allocated on the side stream, read on the compute stream until the end of backward
            for wp, wg in packs["conv"].values():
                wp.record_stream(main)
                wg.record_stream(main)
            packs["w1p"].record_stream(main)
        return packs

    def _forward_branch(self, br, x: torch.Tensor, feat: torch.Tensor, c_off: int, get_packs) -> List[dict]:
        """Conv stack of one branch on the CURRENT stream; the last block writes its channel range of ``feat``.
        ``get_packs()`` returns the step's packed weights, building them at the first call -- which comes right after the
        first block's kernels have been enqueued, so the host launches the ~13 packing kernels while the GPU is busy (at the
        very start of forward the GPU has just been drained by the previous step's ``loss.item()`` and would sit idle for the
        0.3 ms those launches take on the host)."""
        n, c_total, H, W = x.shape
        dev = x.device
        h, w = H, W
        cur = None
        blocks = []
        for li, (conv, bn) in enumerate(br.pairs):
            cout, cin = conv.out_channels, conv.in_channels
            last = li == len(br.pairs) - 1
            if last:
                dst, cstride, coff = feat, self.feat_channels, c_off
            else:
                dst, cstride, coff = self._new((n, h // 2, w // 2, cout), torch.bfloat16, dev), cout, 0
            if li == 0 and self.first_block_mode == "gram":
                # first block: statistics from the input's patch Gram matrix, then the fused eval-style kernel;
                # the full-resolution conv output is never written
                T = 9 * cin
                gram = self._new((T + T * T,), torch.float64, dev)
                ws = workspace("ctk_first_patch_gram_workspace_bytes", cin, device=dev)
                call("ctk_first_patch_gram", ptr(x), c_int(n), c_int(c_total), c_int(br.c_offset), c_int(cin), c_int(h),
                     c_int(w), ptr(gram), ws[1], ws[2], stream())
                self._sync_stats(gram)
                count = float(n) * h * w * self.stat_world
                mom = self._new((2 * cout,), torch.float32, dev)
                call("ctk_first_moments", ptr(gram), ptr(conv.weight), c_int(cout), c_int(cin), c_double(count),
                     ptr(mom), stream())
                scale, shift, mean, invstd = self._bn_finalize(mom, count, conv.bias, bn, dev, moments=True)
                wfold = self._new((cout, T), torch.float32, dev)
                call("ctk_pack_first_weight", ptr(conv.weight), ptr(scale), c_int(cout), c_int(cin), ptr(wfold), stream())
                codes = self._new((n, h // 2, w // 2, cout // 8), torch.int32, dev)
                call("ctk_conv_first_pool_codes", ptr(x), c_int(n), c_int(c_total), c_int(br.c_offset), c_int(cin), c_int(h),
                     c_int(w), ptr(wfold), ptr(shift), c_int(cout), c_float(LEAKY_SLOPE), ptr(dst), c_int(cstride),
                     c_int(coff), ptr(codes), stream())
                blocks.append({"y": None, "x_in": None, "pooled": (dst, cstride, coff), "scale": scale, "shift": shift,
                               "mean": mean, "invstd": invstd, "h": h, "w": w, "conv": conv, "bn": bn, "gram": gram,
                               "codes": codes})
                cur = dst
                h, w = h // 2, w // 2
                continue
            y = self._new((n, h, w, cout), torch.bfloat16, dev)
            stats = self._new((2 * cout,), torch.float32, dev)
            if li == 0:
                ws = workspace("ctk_conv_first_raw_workspace_bytes", cout, device=dev)
                call("ctk_conv_first_raw", ptr(x), c_int(n), c_int(c_total), c_int(br.c_offset), c_int(cin), c_int(h),
                     c_int(w), ptr(conv.weight), c_int(cout), ptr(y), ptr(stats), ws[1], ws[2], stream())
            else:
                packs = get_packs()
                wp = packs["conv"][conv][0]
                if packs["conv_ready"] is not None:
                    torch.cuda.current_stream(dev).wait_event(packs["conv_ready"])
                ws = workspace("ctk_conv3x3_tc_raw_workspace_bytes", cout, device=dev)
                call("ctk_conv3x3_tc_raw", ptr(cur), c_int(n), c_int(h), c_int(w), c_int(cin), ptr(wp), c_int(cout),
                     ptr(y), ptr(stats), ws[1], ws[2], stream(),
                     meta={"flops": 2.0 * n * h * w * cout * 9 * cin, "kernels": 2, "role": "fwd"})
            self._sync_stats(stats)
            scale, shift, mean, invstd = self._bn_finalize(stats, float(n) * h * w * self.stat_world, conv.bias, bn, dev)
            call("ctk_bn_act_pool_fwd", ptr(y), c_int(n), c_int(h), c_int(w), c_int(cout), ptr(scale), ptr(shift),
                 c_float(LEAKY_SLOPE), ptr(dst), c_int(cstride), c_int(coff), stream())
            blocks.append({"y": y, "x_in": cur, "pooled": (dst, cstride, coff), "scale": scale, "shift": shift, "mean": mean, "invstd": invstd,
                           "h": h, "w": w, "conv": conv, "bn": bn, "w_dgrad": get_packs()["conv"][conv][1] if li > 0 else None})
            cur = dst
            h, w = h // 2, w // 2
        return blocks

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, dict]:
        """Returns (scores [n,1], saved state).  The saved state belongs to THIS forward pass (the autograd edge keeps it
        on its ctx), so several forwards may be in flight before their backwards run -- gradient accumulation over
        micro-batches, two losses, a train-mode forward under no_grad in between."""
        _lib.require_device(x, torch.float32, "input batch")
        if x.dim() != 4 or x.shape[1] != 2 or x.shape[2] % 32 or x.shape[3] % 32:
            raise _lib.CtkError(f"input must be [N,2,H,W] float32 with H, W multiples of 32, got {tuple(x.shape)}")
        call("ctk_device_check")
        for p in self.params:
            _lib.require_device(p, torch.float32, "parameter")
        n, c_total, H, W = x.shape
        if n < 2:
            # nn.BatchNorm1d in train() refuses a single sample (torch.nn.functional._verify_batch_size); so does this path
            raise ValueError(f"Expected more than 1 value per channel when training, got input size torch.Size([{n}, "
                             f"{self.lin[0].out_features}])")
        dev = x.device
        depth = len(self.branches[0].pairs)
        hf, wf = H >> depth, W >> depth
        fc1, fc2, fc3 = self.lin
        K = fc1.in_features
        if K != hf * wf * self.feat_channels:
            raise _lib.CtkError("input size does not match the model's first Linear layer")
        m_pad = _pad(n, 128)
        if self.stat_check_equal_batch is not None:
            self.stat_check_equal_batch(n, dev)
        sv = {"x": x, "n": n, "H": H, "W": W, "m_pad": m_pad, "blocks": []}
        feat = self._padded((m_pad, hf, wf, self.feat_channels), n, m_pad, dev)
        c_off = 0
        main = torch.cuda.current_stream(dev)
        pack_box: list = []
        params_ready = None
        if self.overlap_pack:
            params_ready = torch.cuda.Event()
            params_ready.record(main)           # recorded before this step's first kernel: the side stream need not wait for it

        def get_packs() -> dict:
            if not pack_box:
                pack_box.append(self._pack_weights(dev, main, params_ready))
            return pack_box[0]

        overlap = self.overlap_streams and len(self.branches) > 1
        if overlap:
            get_packs()                         # the branch streams fork below: build the packs on the main stream's timeline
        used_streams = []
        if overlap:
            fork = torch.cuda.Event()
            fork.record(main)                   # feat is zero-filled and x is ready behind this point
        for bi, br in enumerate(self.branches):
            if overlap:
                side = self._side_stream("branch", bi, dev)
                side.wait_event(fork)
                used_streams.append(side)
                with torch.cuda.stream(side):
                    blocks = self._forward_branch(br, x, feat, c_off, get_packs)
            else:
                blocks = self._forward_branch(br, x, feat, c_off, get_packs)
            sv["blocks"].append({"branch": br, "blocks": blocks, "c_off": c_off})
            c_off += br.channels[-1]
        if overlap:
            self._join(main, used_streams)
        sv["feat"], sv["hf"], sv["wf"] = feat, hf, wf

        # ---- FC1 (tcgen05 split-K) + fp32 head
        f1, f2 = fc1.out_features, fc2.out_features
        hw = hf * wf
        packs = get_packs()
        w1p = packs["w1p"]
        if packs["fc1_ready"] is not None:
            main.wait_event(packs["fc1_ready"])
        tiles = (m_pad // 128) * (f1 // 128)
        splits = fc1_splits(tiles, K, dev)
        partial = self._new((splits, m_pad, f1), torch.float32, dev)
        call("ctk_gemm_bf16_splitk", ptr(feat), ptr(w1p), c_int(m_pad), c_int(f1), c_int(K), c_int(splits), ptr(partial),
             stream(), meta={"flops": 2.0 * m_pad * f1 * K})
        z1 = self._new((n, f1), torch.float32, dev)
        st1 = self._new((2 * f1,), torch.float32, dev)
        call("ctk_colstat", ptr(partial), c_int(splits), c_longlong(m_pad * f1), c_int(f1), ptr(fc1.bias), c_int(n), c_int(f1),
             ptr(z1), ptr(st1), stream())
        self._sync_stats(st1)
        bn1 = self._bn_finalize(st1, float(n) * self.stat_world, None, self.bns[0], dev)
        masks = self._masks(n, f1, f2, dev)
        a1 = self._new((n, f1), torch.float32, dev)
        call("ctk_bn1d_act_drop_fwd", ptr(z1), ptr(bn1[0]), ptr(bn1[1]), ptr(masks[0]), c_float(self.drop_p[0]),
             c_float(LEAKY_SLOPE), c_int(n), c_int(f1), ptr(a1), stream())
        z2 = self._new((n, f2), torch.float32, dev)
        call("ctk_sgemm_strided", ptr(a1), c_longlong(f1), c_longlong(1), ptr(fc2.weight), c_longlong(f1), c_longlong(1),
             ptr(fc2.bias), c_int(n), c_int(f2), c_int(f1), ptr(z2), c_int(f2), stream())
        st2 = self._new((2 * f2,), torch.float32, dev)
        call("ctk_colstat", ptr(z2), c_int(1), c_longlong(0), c_int(f2), ptr(None), c_int(n), c_int(f2), ptr(None), ptr(st2),
             stream())
        self._sync_stats(st2)
        bn2 = self._bn_finalize(st2, float(n) * self.stat_world, None, self.bns[1], dev)
        a2 = self._new((n, f2), torch.float32, dev)
        call("ctk_bn1d_act_drop_fwd", ptr(z2), ptr(bn2[0]), ptr(bn2[1]), ptr(masks[1]), c_float(self.drop_p[1]),
             c_float(LEAKY_SLOPE), c_int(n), c_int(f2), ptr(a2), stream())
        out = self._new((n, 1), torch.float32, dev)
        call("ctk_head_out_fwd", ptr(a2), ptr(fc3.weight), ptr(fc3.bias), c_int(n), c_int(f2), c_int(self.sigmoid_half),
             ptr(out), stream())
        # (a detached alias of the output: the returned tensor will own the autograd node that owns this dict -- saving the
        #  tensor itself would make a reference cycle that keeps a whole step's activations alive until the garbage collector runs)
        sv.update(z1=z1, bn1=bn1, a1=a1, z2=z2, bn2=bn2, a2=a2, out=out.detach(), masks=masks, w1p=w1p)
        # ctk_bn_finalize wrote the running statistics through raw pointers: bump the tensors' version counters so that
        # anything keyed on them (the eval engine's derived-parameter cache) sees the change
        torch.autograd.graph.increment_version([b for b in self.model.buffers()])
        return out, sv

    def _masks(self, n, f1, f2, dev):
        if self.forced_masks is not None:
            ms = []
            for m, f in zip(self.forced_masks, (f1, f2)):
                if m is not None:
                    _lib.require_device(m, torch.float32, "dropout mask")
                    if tuple(m.shape) != (n, f):
                        raise _lib.CtkError("dropout mask has the wrong shape")
                ms.append(m)
            return tuple(ms)
        # keep-masks drawn by libctk's Philox kernel (one launch for both layers): seeded from torch's seed at engine
        # creation, one counter offset per forward pass -- reproducible under torch.manual_seed, no ATen kernel in the step
        p1, p2 = self.drop_p
        if p1 == 0.0 and p2 == 0.0:
            return (None, None)
        m1 = self._new((n, f1), torch.float32, dev) if p1 > 0.0 else None
        m2 = self._new((n, f2), torch.float32, dev) if p2 > 0.0 else None
        call("ctk_dropout_masks", ptr(m1), c_longlong(n * f1 if m1 is not None else 0), c_float(p1), ptr(m2),
             c_longlong(n * f2 if m2 is not None else 0), c_float(p2), c_ulonglong(self.dropout_seed),
             c_ulonglong(self._dropout_calls), stream())
        self._dropout_calls += 1
        return (m1, m2)

    def _backward_branch(self, entry: dict, sv: dict, dfeat: torch.Tensor, done, wgrad_stream=None,
                         defer_wgrad: bool = False) -> None:
        """Backward of one branch's conv stack (last block first) on the CURRENT stream.

        ``wgrad_stream``: weight gradients go to that stream.  ``defer_wgrad`` picks the order: False (the
        ``overlap_streams`` experiment) forks right behind the BatchNorm backward of the layer, True (``overlap_wgrad``)
        forks behind the layer's INPUT gradient, so that the weight gradient runs beside the next layer's BatchNorm passes
        instead of competing with the dgrad kernel for whole SMs."""
        x, n = sv["x"], sv["n"]
        dev = x.device
        br, blocks = entry["branch"], entry["blocks"]
        dp, dp_cstride, dp_coff = dfeat, self.feat_channels, entry["c_off"]

        def fork(*read_on_side):
            """Context of the weight-gradient stream, ordered behind everything enqueued so far on the current one; the
            tensors it reads were allocated on the current stream, so the allocator must not recycle them before the side
            stream is done with them."""
            if wgrad_stream is None:
                return contextlib.nullcontext()
            ready = torch.cuda.Event()
            ready.record()
            wgrad_stream.wait_event(ready)
            for t in read_on_side:
                if t is not None:
                    t.record_stream(wgrad_stream)
            return torch.cuda.stream(wgrad_stream)

        for li in range(len(blocks) - 1, -1, -1):
            b = blocks[li]
            conv, bn, h, w = b["conv"], b["bn"], b["h"], b["w"]
            cout, cin = conv.out_channels, conv.in_channels
            if b.get("gram") is not None:
                # first block: the forward pass stored arg-max / sign codes, so the data term of the weight gradient and
                # sum(dA) are one gather over the input; sum(dA * xhat) follows from them in the finalize kernel
                if dp_cstride != cout or dp_coff != 0:
                    raise _lib.CtkError("the first block's output gradient must be dense")
                T = 9 * cin
                # nothing downstream reads these gradients: with deferred weight gradients the whole first block goes to
                # the side stream (not under SyncBN, whose blocking statistics all-reduces belong on the compute stream)
                side = defer_wgrad and self.stat_allreduce is None
                with (fork(dp, b["codes"], b["gram"], b["scale"], b["mean"], b["invstd"]) if side
                      else contextlib.nullcontext()):
                    sums = self._new((2 * cout,), torch.float32, dev)
                    t1 = self._new((cout, T), torch.float32, dev)
                    ws = workspace("ctk_first_wgrad_codes_workspace_bytes", cin, cout, device=dev)
                    call("ctk_first_wgrad_codes", ptr(x), c_int(n), c_int(x.shape[1]), c_int(br.c_offset), c_int(cin),
                         c_int(h), c_int(w), ptr(b["codes"]), ptr(dp), c_int(cout), c_float(LEAKY_SLOPE), ptr(t1), ptr(sums),
                         ws[1], ws[2], stream())
                    dw = self._new(tuple(conv.weight.shape), torch.float32, dev)
                    # SyncBN: the saved Gram matrix is already the global one, so reduce t1 / sum(dA) too and form the
                    # global gradient on every rank; dividing by the world size makes the exchange's mean leave it as is
                    self._sync_stats(t1)
                    self._sync_stats(sums)
                    call("ctk_first_wgrad_finalize", ptr(t1), ptr(b["gram"]), ptr(conv.weight), ptr(b["scale"]),
                         ptr(b["mean"]), ptr(b["invstd"]), ptr(sums), c_double(float(n) * h * w * self.stat_world),
                         c_int(cout), c_int(cin), ptr(dw), stream())
                    if self.stat_world > 1:
                        dw.div_(self.stat_world)
                        sums.div_(self.stat_world)
                    done(bn.bias, sums[:cout])
                    done(bn.weight, sums[cout:])
                    done(conv.weight, dw)
                done(conv.bias, self._zero_grad_of(conv.bias))
                continue
            sums = self._new((2 * cout,), torch.float32, dev)
            pooled, p_cstride, p_coff = b["pooled"]
            # sums from the pooled tensors; channel groups whose BatchNorm parameters make that reconstruction lossy
            # (gamma == 0 or |beta| > 8 |gamma|) are reduced from the raw conv output instead -- decided on the device
            ws = workspace("ctk_bn_bwd_reduce_workspace_bytes", cout, device=dev)
            call("ctk_bn_bwd_reduce_guarded", ptr(b["y"]), c_int(n), c_int(h), c_int(w), ptr(b["scale"]), ptr(b["shift"]),
                 ptr(b["mean"]), ptr(b["invstd"]), ptr(pooled), c_int(p_cstride), c_int(p_coff), ptr(dp), c_int(dp_cstride),
                 c_int(dp_coff), c_int(cout), ptr(bn.weight), ptr(bn.bias), c_float(LEAKY_SLOPE), ptr(sums), ws[1], ws[2],
                 stream())
            done(bn.bias, sums[:cout])
            done(bn.weight, sums[cout:])
            dy = self._new((n, h, w, cout), torch.bfloat16, dev)
            call("ctk_bn_bwd_apply", ptr(b["y"]), ptr(dp), c_int(dp_cstride), c_int(dp_coff), c_int(n), c_int(h), c_int(w),
                 c_int(cout), ptr(b["scale"]), ptr(b["shift"]), ptr(b["mean"]), ptr(b["invstd"]), ptr(self._global_sums(sums)),
                 c_float(LEAKY_SLOPE), ptr(dy), stream())
            b["y"] = None

            def weight_gradient():
                # off the critical path (nothing in this backward reads it)
                with fork(dy, b["x_in"]):
                    dw = self._new(tuple(conv.weight.shape), torch.float32, dev)
                    if li == 0:
                        ws = workspace("ctk_conv_first_wgrad_workspace_bytes", cin, cout, device=dev)
                        call("ctk_conv_first_wgrad", ptr(dy), ptr(x), c_int(n), c_int(x.shape[1]), c_int(br.c_offset),
                             c_int(cin), c_int(h), c_int(w), c_int(cout), ptr(dw), ws[1], ws[2], stream())
                    else:
                        ws = workspace("ctk_conv3x3_wgrad_tc_workspace_bytes", cin, cout, device=dev)
                        call("ctk_conv3x3_wgrad_tc", ptr(dy), ptr(b["x_in"]), c_int(n), c_int(h), c_int(w), c_int(cin),
                             c_int(cout), ptr(dw), ws[1], ws[2], stream(), meta={"flops": 2.0 * n * h * w * cout * 9 * cin})
                    done(conv.weight, dw)

            if not defer_wgrad:
                weight_gradient()
            # the conv bias feeds a train-mode BatchNorm, so its gradient is sum(dY) = 0 identically
            done(conv.bias, self._zero_grad_of(conv.bias))
            if li > 0:
                wg = b["w_dgrad"]
                dx = self._new((n, h, w, cin), torch.bfloat16, dev)
                call("ctk_conv3x3_tc_raw", ptr(dy), c_int(n), c_int(h), c_int(w), c_int(cout), ptr(wg), c_int(cin), ptr(dx),
                     ptr(None), ptr(None), c_size_t(0), stream(),
                     meta={"flops": 2.0 * n * h * w * cout * 9 * cin, "role": "dgrad"})
                dp, dp_cstride, dp_coff = dx, cin, 0
            if defer_wgrad:
                weight_gradient()
            del dy

    # ------------------------------------------------------------------ backward
    def backward(self, sv: dict, dout: torch.Tensor) -> Dict[torch.nn.Parameter, torch.Tensor]:
        """Gradients of every parameter for the forward pass that produced ``sv`` (consumed: a second backward through the
        same pass raises, like autograd without retain_graph)."""
        if sv.get("consumed"):
            raise _lib.CtkError("backward called twice for the same train-mode forward (activations already released)")
        sv["consumed"] = True
        dout = dout.contiguous().float()
        _lib.require_device(dout, torch.float32, "output gradient")
        if self.backward_sm_reserve > 0:
            call("ctk_set_persistent_sm_reserve", c_int(self.backward_sm_reserve))
        try:
            return self._backward(sv, dout)
        finally:
            if self.backward_sm_reserve > 0:
                call("ctk_set_persistent_sm_reserve", c_int(0))

    def _backward(self, sv: dict, dout: torch.Tensor) -> Dict[torch.nn.Parameter, torch.Tensor]:
        grads: Dict[torch.nn.Parameter, torch.Tensor] = {}

        n, dev = sv["n"], dout.device
        overlap = self.overlap_streams                        # experiment: branches and weight gradients on side streams
        defer = self.overlap_wgrad and not overlap            # weight gradients beside the next layer's BatchNorm passes
        comm = self._side_stream("comm", 0, dev) if (overlap or defer) and self.on_grad_ready is not None else None

        def done(p, g):
            grads[p] = g
            if self.on_grad_ready is None:
                return
            if comm is None:
                self.on_grad_ready(p, g)
                return
            # gradients are produced on several streams: hand them to the exchange from ONE stream that has waited for
            # each producer, so that a bucket flushed under it never contains a tensor that is still being written
            ev = torch.cuda.Event()
            ev.record()
            comm.wait_event(ev)
            g.record_stream(comm)
            with torch.cuda.stream(comm):
                self.on_grad_ready(p, g)

        fc1, fc2, fc3 = self.lin
        f1, f2 = fc1.out_features, fc2.out_features
        K = fc1.in_features
        m_pad = sv["m_pad"]
        k_pad = _pad(n, 64)
        masks = sv["masks"]
        # ---- head
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
        call("ctk_bn1d_bwd_apply", ptr(dact2), ptr(sv["z2"]), ptr(sc2), ptr(mu2), ptr(is2), ptr(self._global_sums(sums2)),
             c_int(n), c_int(f2), ptr(dz2), ptr(None), ptr(None), c_int(0), stream())
        dw2 = self._new((f2, f1), torch.float32, dev)
        call("ctk_sgemm_strided", ptr(dz2), c_longlong(1), c_longlong(f2), ptr(sv["a1"]), c_longlong(1), c_longlong(f1),
             ptr(None), c_int(f2), c_int(f1), c_int(n), ptr(dw2), c_int(f1), stream())
        done(fc2.weight, dw2)
        done(fc2.bias, self._colsum(dz2, n, f2))
        da1 = self._new((n, f1), torch.float32, dev)
        call("ctk_sgemm_strided", ptr(dz2), c_longlong(f2), c_longlong(1), ptr(fc2.weight), c_longlong(1), c_longlong(f1),
             ptr(None), c_int(n), c_int(f1), c_int(f2), ptr(da1), c_int(f1), stream())
        sc1, sh1, mu1, is1 = sv["bn1"]
        dact1 = self._new((n, f1), torch.float32, dev)
        sums1 = self._new((2 * f1,), torch.float32, dev)
        call("ctk_bn1d_bwd_reduce", ptr(da1), ptr(masks[0]), c_float(self.drop_p[0]), ptr(sv["z1"]), ptr(sc1), ptr(sh1),
             ptr(mu1), ptr(is1), c_float(LEAKY_SLOPE), c_int(n), c_int(f1), ptr(dact1), ptr(sums1), stream())
        done(self.bns[0].bias, sums1[:f1])
        done(self.bns[0].weight, sums1[f1:])
        dz1 = self._new((n, f1), torch.float32, dev)
        dz1_bf = self._padded((m_pad, f1), n, m_pad, dev)
        dz1t_bf = self._padded((f1, k_pad), n, k_pad, dev)
        call("ctk_bn1d_bwd_apply", ptr(dact1), ptr(sv["z1"]), ptr(sc1), ptr(mu1), ptr(is1), ptr(self._global_sums(sums1)),
             c_int(n), c_int(f1), ptr(dz1), ptr(dz1_bf), ptr(dz1t_bf), c_int(k_pad), stream())
        done(fc1.bias, self._colsum(dz1, n, f1))
        # ---- FC1: dW1 (reference column order) and dfeat (NHWC order)
        hf, wf = sv["hf"], sv["wf"]
        hw = hf * wf
        featT = self._new((K, k_pad), torch.bfloat16, dev)
        call("ctk_feat_transpose_bf16", ptr(sv["feat"]), c_int(n), c_int(hw), c_int(self.feat_channels), ptr(featT),
             c_int(k_pad), stream())
        dw1 = self._new((f1, K), torch.float32, dev)
        call("ctk_gemm_bf16_splitk", ptr(dz1t_bf), ptr(featT), c_int(f1), c_int(K), c_int(k_pad), c_int(1), ptr(dw1), stream(),
             meta={"flops": 2.0 * f1 * K * k_pad})
        done(fc1.weight, dw1)
        del featT
        # dfeat = dZ1 * W1 with the forward pass's packed weight [f1][HW*C] as an MN-major operand (no transposed copy)
        dfeat = self._new((m_pad, hf, wf, self.feat_channels), torch.bfloat16, dev)
        call("ctk_gemm_bf16_bt_out_bf16", ptr(dz1_bf), ptr(sv["w1p"]), c_int(m_pad), c_int(K), c_int(f1), ptr(dfeat), stream(),
             meta={"flops": 2.0 * m_pad * f1 * K})
        # ---- conv stacks, last block first
        main = torch.cuda.current_stream(dev)
        used_streams = [comm] if comm is not None else []
        if overlap:
            fork = torch.cuda.Event()
            fork.record(main)                   # dfeat is complete behind this point
        for bi, entry in enumerate(sv["blocks"]):
            if overlap:
                wg = self._side_stream("wgrad", bi, dev)
                used_streams.append(wg)
                if len(sv["blocks"]) > 1:
                    side = self._side_stream("branch", bi, dev)
                    side.wait_event(fork)
                    used_streams.append(side)
                    with torch.cuda.stream(side):
                        self._backward_branch(entry, sv, dfeat, done, wg)
                else:
                    self._backward_branch(entry, sv, dfeat, done, wg)
            elif defer:
                wg = self._side_stream("wgrad", 0, dev)
                if wg not in used_streams:
                    used_streams.append(wg)
                self._backward_branch(entry, sv, dfeat, done, wg, defer_wgrad=True)
            else:
                self._backward_branch(entry, sv, dfeat, done)
        if overlap or defer:
            self._join(main, used_streams)
            for g in grads.values():            # produced on side streams, consumed (optimizer, NCCL, user) on this one
                g.record_stream(main)
        if self.finalize_grads is not None:
            grads = self.finalize_grads(grads)
        return grads


class _CtkTrainFunction(torch.autograd.Function):
    """Graph edge for the whole network: forward = TrainEngine.forward, backward hands every parameter its gradient."""

    @staticmethod
    def forward(ctx, engine, x, *params):
        ctx.engine = engine
        out, ctx.sv = engine.forward(x)
        return out

    @staticmethod
    def backward(ctx, dout):
        engine, sv = ctx.engine, ctx.sv
        ctx.sv = None
        grads = engine.backward(sv, dout)
        return (None, None) + tuple(grads.get(p) for p in engine.params)


def train_forward(engine: TrainEngine, x: torch.Tensor) -> torch.Tensor:
    return _CtkTrainFunction.apply(engine, x, *engine.params)
