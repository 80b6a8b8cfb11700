"""Host logic of ctk.CosineWarmupLR -- the scheduler train_model.py:356-365 configures ("cosine_warmup") but :376-387 never
creates.  Pure param-group arithmetic: runs on the CPU with a torch optimizer."""
import math

import pytest
import torch

pytestmark = pytest.mark.filterwarnings("ignore:Detected call of `lr_scheduler.step")

import ctk

PARAMS = {"warmup_epochs": 5, "max_lr": 1e-4, "final_lr": 1e-7, "total_epochs": 40}     # the reference's dict, num_epochs = 40


def _make(lr0=5e-4, **kw):
    p = [torch.nn.Parameter(torch.zeros(3))]
    opt = torch.optim.Adam(p, lr=lr0, weight_decay=1e-4)                                 # train_model.py:637
    return p, opt, ctk.CosineWarmupLR(opt, **{**PARAMS, **kw})


def test_cosine_warmup_schedule_shape():
    p, opt, sched = _make()
    lrs = [opt.param_groups[0]["lr"]]                                                    # what train_model.py:412,465 log
    for _ in range(PARAMS["total_epochs"] + 3):
        p[0].grad = torch.ones(3)
        opt.step()
        sched.step()                                                                     # once per epoch, :451-452
        lrs.append(opt.param_groups[0]["lr"])
    w, total, mx, fin = PARAMS["warmup_epochs"], PARAMS["total_epochs"], PARAMS["max_lr"], PARAMS["final_lr"]
    assert lrs[0] == pytest.approx(mx / w) and lrs[w - 1] == pytest.approx(mx)           # linear ramp ends at max_lr
    assert all(b > a for a, b in zip(lrs[:w - 1], lrs[1:w]))
    assert all(b < a for a, b in zip(lrs[w - 1:total - 1], lrs[w:total]))                # then strictly decreasing
    assert lrs[total - 1] == pytest.approx(fin, rel=1e-9) and lrs[total + 2] == pytest.approx(fin, rel=1e-9)
    e = 20                                                                               # closed form at an interior epoch
    t = (e - w + 1) / (total - w)
    assert lrs[e] == pytest.approx(fin + (mx - fin) * 0.5 * (1 + math.cos(math.pi * t)), rel=1e-12)
    assert sched.get_last_lr() == [lrs[-1]]


def test_cosine_warmup_checkpoint_resume_and_validation():
    p, opt, sched = _make()
    for _ in range(7):
        sched.step()
    osd, ssd = opt.state_dict(), sched.state_dict()
    want = []
    for _ in range(6):
        sched.step()
        want.append(opt.param_groups[0]["lr"])
    p2, opt2, sched2 = _make()
    opt2.load_state_dict(osd)
    sched2.load_state_dict(ssd)
    got = []
    for _ in range(6):
        sched2.step()
        got.append(opt2.param_groups[0]["lr"])
    assert got == want
    # no warm-up: starts the cosine at once; every param group follows
    q = [torch.nn.Parameter(torch.zeros(2)), torch.nn.Parameter(torch.zeros(2))]
    opt3 = torch.optim.SGD([{"params": [q[0]]}, {"params": [q[1]], "lr": 1.0}], lr=0.1)
    s3 = ctk.CosineWarmupLR(opt3, warmup_epochs=0, max_lr=1e-3, final_lr=0.0, total_epochs=4)
    seq = [opt3.param_groups[1]["lr"]]
    for _ in range(4):
        s3.step()
        seq.append(opt3.param_groups[1]["lr"])
    assert seq[0] == pytest.approx(1e-3 * 0.5 * (1 + math.cos(math.pi / 4))) and seq[3] == pytest.approx(0.0, abs=1e-18)
    assert opt3.param_groups[0]["lr"] == opt3.param_groups[1]["lr"]
    for bad in ({"warmup_epochs": 40}, {"warmup_epochs": -1}, {"max_lr": 0.0}, {"final_lr": 1.0}):
        with pytest.raises(ValueError):
            _make(**bad)


def test_fc1_split_factor_follows_the_sm_count():
    """Host logic of the FC1 split-K GEMM (ctk.engine.fc1_splits): as many K ranges as SMs per output tile, at least eight
    64-element K blocks each; the kernel deals the blocks out raggedly (tests/test_gpu_kernels.py checks the ranges)."""
    from ctk.engine import fc1_splits
    assert fc1_splits(8, 262144, None, sms=148) == 18          # double-branch, 256-tile batch: 8 tiles x 18 = 144 CTAs
    assert fc1_splits(4, 262144, None, sms=148) == 37          # HostScorer's 64-tile slices (m_pad 128): 148 CTAs
    assert fc1_splits(8, 8192, None, sms=148) == 16            # single-branch: K / 64 = 128 blocks, eight per split
    assert fc1_splits(200, 262144, None, sms=148) == 1         # more tiles than SMs: no split
    assert fc1_splits(1, 64, None, sms=148) == 1
    for tiles, K in ((8, 262144), (4, 262144), (8, 8192), (3, 4096)):
        s = fc1_splits(tiles, K, None, sms=148)
        q, r = divmod(K // 64, s)
        assert q >= 8 or s == 1
        assert r * (q + 1) + (s - r) * q == K // 64             # the ragged ranges cover every K block exactly once
