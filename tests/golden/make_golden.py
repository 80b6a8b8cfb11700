"""Generate tests/golden/{tiles.npz,golden.json} from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports regression_model.py / two_branch_regression.py straight from
/root/reference (nothing is copied), reads five of the reference's own
Training_Data fixture pairs, and records what the reference computes for them:
eval / train-mode outputs, losses, per-tensor gradient norms, parameters after
two Adam steps, and scipy Pearson r.  tests/test_oracle_golden.py then holds
oracle/crosstalk_oracle.py to these numbers; the GPU parity tests hold the CUDA
path to the oracle.
"""
import json
import os
import re
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))

from regression_model import AdvancedRegressionModel            # noqa: E402
from two_branch_regression import SimplifiedTwoBranchRegressionModel  # noqa: E402
from scipy.stats import pearsonr                                 # noqa: E402
import crosstalk_oracle as orc                                   # noqa: E402

IDS = ["14144", "14162", "14470", "14692", "15017"]


def read_tiff_f64(path):
    """The fixtures are classic little-endian TIFFs with one uncompressed 256x256 float64 strip at offset 272."""
    b = open(path, "rb").read()
    assert b[:4] == b"II*\x00" and len(b) == 524560
    return np.frombuffer(b, dtype="<f8", count=65536, offset=272).reshape(256, 256).copy()


def load_tiles():
    mixed_dir, src_dir = os.path.join(REF, "Training_Data/Mixed"), os.path.join(REF, "Training_Data/Source")
    pat = re.compile(r"image_(\d+)_alpha_(\d+\.?\d*)_(mixed|source)\.tif")
    tiles, labels = [], []
    for iid in IDS:
        fm = [f for f in os.listdir(mixed_dir) if pat.search(f) and pat.search(f).group(1) == iid][0]
        alpha = pat.search(fm).group(2)
        fs = f"image_{iid}_alpha_{alpha}_source.tif"
        # train_model.py:166-167: imread(...).astype(np.float32)
        m = read_tiff_f64(os.path.join(mixed_dir, fm)).astype(np.float32)
        s = read_tiff_f64(os.path.join(src_dir, fs)).astype(np.float32)
        tiles.append(np.stack([m, s]))
        labels.append(float(alpha))
    return np.stack(tiles), np.array(labels, dtype=np.float32)


def normalised(tiles):
    out = np.empty_like(tiles)
    for i in range(tiles.shape[0]):
        for c in range(2):
            out[i, c] = orc.normalize_image(tiles[i, c])     # train_model.py:211-216
    return out


def tensor_stats(sd):
    return {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in sd.items()}


def run_model(kind, x, y, x5):
    torch.manual_seed(0)
    model = AdvancedRegressionModel(initial_filters=128, num_conv_blocks=6) if kind == "single" \
        else SimplifiedTwoBranchRegressionModel(initial_filters_per_branch=64)
    res = {"init_stats": tensor_stats(model.state_dict())}
    model.eval()
    with torch.no_grad():
        res["eval_out"] = model(x).flatten().tolist()
    # randomised BN (non-vacuous eval outputs): SURVEY 8c fallback (iii)
    rsd = orc.randomize_bn(model.state_dict(), seed=7)
    saved = {k: v.clone() for k, v in model.state_dict().items()}
    model.load_state_dict(rsd)
    with torch.no_grad():
        res["eval_out_randbn"] = model(x).flatten().tolist()
    model.load_state_dict(saved)
    # calibrated BN: one train-mode pass with momentum 1.0 (running stats := batch stats), then eval on the same tiles
    model.train()
    for mod in model.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
        if isinstance(mod, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            mod.momentum = 1.0
    with torch.no_grad():
        model(x5)
    model.eval()
    with torch.no_grad():
        res["eval_out_calibrated"] = model(x5).flatten().tolist()
    for mod in model.modules():
        if isinstance(mod, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            mod.momentum = 0.1
    model.load_state_dict(saved)
    # train mode, dropout p forced to 0
    model.train()
    for mod in model.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)   # train_model.py:637
    crit = torch.nn.MSELoss()
    losses = []
    for step in range(2):
        opt.zero_grad()
        out = model(x)
        loss = crit(out, y)
        loss.backward()
        if step == 0:
            res["train_out"] = out.detach().flatten().tolist()
            res["grad_norms"] = {k: float(p.grad.double().norm()) for k, p in model.named_parameters()}
        opt.step()
        losses.append(float(loss))
    res["train_losses"] = losses
    res["after2_stats"] = tensor_stats(model.state_dict())
    return res


def main():
    torch.set_num_threads(8)
    tiles, labels = load_tiles()
    np.savez_compressed(os.path.join(HERE, "tiles.npz"), tiles=tiles, labels=labels, ids=np.array(IDS))
    xn = normalised(tiles)
    x4 = torch.from_numpy(xn[:4])
    y4 = torch.from_numpy(labels[:4])[:, None]
    gold = {"ids": IDS, "labels": labels.tolist(), "torch": torch.__version__}
    import scipy
    gold["scipy"] = scipy.__version__
    # test-cross-talk-model.py:59-64
    gold["pearson_scipy_f32"] = [float(pearsonr(xn[i, 0].flatten(), xn[i, 1].flatten())[0]) for i in range(len(IDS))]
    gold["pearson_scipy_f64"] = [float(pearsonr(xn[i, 0].flatten().astype(np.float64),
                                                xn[i, 1].flatten().astype(np.float64))[0]) for i in range(len(IDS))]
    for kind in ("single", "double"):
        gold[kind] = run_model(kind, x4, y4, torch.from_numpy(xn))
        print(kind, "eval", gold[kind]["eval_out"], "train", gold[kind]["train_out"], gold[kind]["train_losses"])
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(gold, f, indent=1)
    print("pearson", gold["pearson_scipy_f32"])


if __name__ == "__main__":
    main()
