"""Generate tests/golden/loss_curve.json: the reference's own training loop run for 200 steps on the CPU.

Run in the build container only (needs /root/reference, ~25 min on 8 cores):

    python tests/golden/make_loss_curve.py [single|double] [steps] [pool] [batch] [threads] [emulate 0|1]

(defaults: 200 steps, 64-tile pool, batches of 16, all cores, with the bf16 emulation -> loss_curve_<kind>.json; any other
pool / batch / thread count is written to loss_curve_<kind>_b<batch>[_t<threads>].json.  The batch-64 curves from a 256-tile
pool are the well-conditioned problem the 1 % loss-curve bound is asserted on; the same loop re-run with a different thread
count changes only ATen's reduction order and measures how closely the reference reproduces ITSELF.)

The UNMODIFIED reference model class + torch.optim.Adam(lr=5e-4, weight_decay=1e-4) + nn.MSELoss run the inner loop of
train_model.py:415-426 (zero_grad / forward / loss / backward / step / loss.item()) on synthetic tiles
(oracle.synthetic_batch, 64 tiles cycled in batches of 16).  ``torch.manual_seed(SEED0 + t)`` before every forward makes the
two nn.Dropout draws reproducible: the GPU test rebuilds the same keep-masks with F.dropout on ones in the same order.
Beside the fp32 curve the oracle's bf16-rounding emulation (same masks) is recorded: it is the yardstick for how far a
bf16-operand trajectory may drift from the fp32 one on this problem.
"""
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))

from regression_model import AdvancedRegressionModel            # noqa: E402
from two_branch_regression import SimplifiedTwoBranchRegressionModel  # noqa: E402
import crosstalk_oracle as orc                                   # noqa: E402

SEED0 = 1000
POOL, BATCH = 64, 16
DATA_SEED = 4321


def masks_for(step, n, p):
    """The keep-masks nn.Dropout draws inside the reference forward after torch.manual_seed(SEED0 + step)."""
    torch.manual_seed(SEED0 + step)
    m1 = (F.dropout(torch.ones(n, 512), p, True) != 0).float()
    m2 = (F.dropout(torch.ones(n, 128), p, True) != 0).float()
    return m1, m2


def main():
    global POOL, BATCH
    kind = sys.argv[1] if len(sys.argv) > 1 else "single"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    default_shape = len(sys.argv) <= 3
    if len(sys.argv) > 3:
        POOL, BATCH = int(sys.argv[3]), int(sys.argv[4])
    threads = int(sys.argv[5]) if len(sys.argv) > 5 else (os.cpu_count() or 1)
    emulate = (sys.argv[6] != "0") if len(sys.argv) > 6 else True
    torch.set_num_threads(threads)
    x, y = orc.synthetic_batch(POOL, seed=DATA_SEED)
    torch.manual_seed(0)
    model = (AdvancedRegressionModel(initial_filters=128, num_conv_blocks=6) if kind == "single"
             else SimplifiedTwoBranchRegressionModel(initial_filters_per_branch=64))
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    p = 0.1 if kind == "single" else 0.5
    opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)        # train_model.py:637
    crit = torch.nn.MSELoss()                                                      # train_model.py:636
    model.train()
    ref, emu = [], []
    tre = orc.OracleTrainer(kind, {k: v.clone() for k, v in sd0.items()}, lr=5e-4, weight_decay=1e-4)
    t0 = time.time()
    for t in range(steps):
        s = (t * BATCH) % POOL
        xb, yb = x[s:s + BATCH], y[s:s + BATCH]
        torch.manual_seed(SEED0 + t)
        opt.zero_grad()
        loss = crit(model(xb), yb)
        loss.backward()
        opt.step()
        ref.append(loss.item())
        if emulate:
            with orc.emulate_bf16():
                emu.append(tre.step(xb, yb, dropout_masks=masks_for(t, BATCH, p))[0])
        print(f"step {t}: reference {ref[-1]:.6f}  bf16-emulation {emu[-1] if emu else float('nan'):.6f}  ({time.time() - t0:.0f}s)",
              flush=True)
        out = {"kind": kind, "steps": t + 1, "pool": POOL, "batch": BATCH, "data_seed": DATA_SEED, "seed0": SEED0, "threads": threads,
               "lr": 5e-4, "weight_decay": 1e-4, "torch": torch.__version__, "reference_fp32": ref, "oracle_bf16_emulation": emu}
        name = f"loss_curve_{kind}.json" if default_shape else f"loss_curve_{kind}_b{BATCH}" + (f"_t{threads}" if len(sys.argv) > 5 and threads != 6 else "") + ".json"
        path = os.path.join(HERE, name)
        json.dump(out, open(path + ".tmp", "w"))
        os.replace(path + ".tmp", path)


if __name__ == "__main__":
    main()
