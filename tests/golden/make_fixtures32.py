"""Generate tests/golden/tiles32.npz + golden32.json: 32 of the reference's own Training_Data fixture pairs and what the
UNMODIFIED reference computes for them.

Run in the build container only (needs /root/reference):

    python tests/golden/make_fixtures32.py

Tiles: the first 32 pairs in the dataset's sort order (train_model.py:150), float64 TIFF payload -> float32 exactly as
train_model.py:166-167 does.  Recorded per tile: the label alpha from the file name (train_model.py:105), the metrics of
test-cross-talk-model.py:59-79 evaluated with the reference's expressions on the normalised tile (scipy pearsonr in
float32, histogram correlation, RMSE), and the eval-mode outputs of both reference model classes (seed-0 init with the
randomised BatchNorm set, oracle.randomize_bn(seed 7), so that the outputs have spread).
"""
import json
import os
import re
import sys

import numpy as np
import torch
from scipy.stats import pearsonr

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))

from regression_model import AdvancedRegressionModel            # noqa: E402
from two_branch_regression import SimplifiedTwoBranchRegressionModel  # noqa: E402
import crosstalk_oracle as orc                                   # noqa: E402

N = 32


def read_tiff_f64(path):
    b = open(path, "rb").read()
    assert b[:4] == b"II*\x00" and len(b) == 524560
    return np.frombuffer(b, dtype="<f8", count=65536, offset=272).reshape(256, 256).copy()


def main():
    pat = re.compile(r"image_(\d+)_alpha_(\d+\.?\d*)_(mixed|source)\.tif")
    mixed_dir, src_dir = os.path.join(REF, "Training_Data/Mixed"), os.path.join(REF, "Training_Data/Source")
    names = sorted(os.listdir(mixed_dir))[:N]
    tiles, labels, ids = [], [], []
    for fm in names:
        m = pat.search(fm)
        iid, alpha = m.group(1), m.group(2)
        a = read_tiff_f64(os.path.join(mixed_dir, fm)).astype(np.float32)
        b = read_tiff_f64(os.path.join(src_dir, f"image_{iid}_alpha_{alpha}_source.tif")).astype(np.float32)
        tiles.append(np.stack([a, b]))
        labels.append(float(alpha))
        ids.append(iid)
    tiles = np.stack(tiles)
    xn = np.stack([np.stack([orc.normalize_image(t[0]), orc.normalize_image(t[1])]) for t in tiles])   # train_model.py:211-216
    gold = {"ids": ids, "labels": labels, "torch": torch.__version__, "pearson_f32": [], "rmse": [], "hist_corr": []}
    for j in range(N):
        a, b = xn[j][0].flatten(), xn[j][1].flatten()
        gold["pearson_f32"].append(float(pearsonr(a, b)[0]))                                  # test-cross-talk-model.py:64
        gold["rmse"].append(float(np.sqrt(np.mean((xn[j][0] - xn[j][1]) ** 2))))              # :79
        h1, h2 = np.histogram(a, bins=256)[0], np.histogram(b, bins=256)[0]                   # :65-66
        gold["hist_corr"].append(float(pearsonr(h1, h2)[0]))                                  # :70
    x = torch.from_numpy(xn)
    for kind, ctor in (("single", lambda: AdvancedRegressionModel(initial_filters=128, num_conv_blocks=6)),
                       ("double", lambda: SimplifiedTwoBranchRegressionModel(initial_filters_per_branch=64))):
        torch.manual_seed(0)
        model = ctor()
        model.load_state_dict(orc.randomize_bn(model.state_dict(), seed=7))
        model.eval()
        with torch.no_grad():
            gold[f"{kind}_eval_randomized_bn"] = model(x).flatten().tolist()
    np.savez_compressed(os.path.join(HERE, "tiles32.npz"), tiles=tiles)
    json.dump(gold, open(os.path.join(HERE, "golden32.json"), "w"), indent=1)
    print({k: (v[:3] if isinstance(v, list) else v) for k, v in gold.items()})


if __name__ == "__main__":
    main()
