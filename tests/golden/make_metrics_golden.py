"""Generate tests/golden/metrics.json: the comparison metrics of test-cross-talk-model.py:59-79 on the fixture tiles.

Run in the build container (needs numpy + scipy, the libraries the reference calls; the tiles come from tiles.npz, which
make_golden.py cut from the reference's Training_Data):

    python tests/golden/make_metrics_golden.py

Every value is computed with the reference's own expressions, verbatim: np.histogram(plane.flatten(), bins=256)[0],
scipy.stats.pearsonr on the histograms (NaN guard on np.std == 0), np.sqrt(np.mean((img0 - img1) ** 2)),
sklearn's normalized_mutual_info_score on np.digitize'd planes (test-cross-talk-model.py:71-74,84) -- on the
per-plane min-max normalised tiles the reference's dataset yields (train_model.py:211-216) and, as a second case with
a non-trivial value range, on the raw tiles.
"""
import json
import os
import sys

import numpy as np
from scipy.stats import pearsonr
from sklearn.metrics import normalized_mutual_info_score

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import crosstalk_oracle as orc   # noqa: E402  (normalize_image only)


def metrics(images):
    out = {"rmse": [], "hist_corr": [], "nmi": [], "hist0_head": [], "hist1_head": [], "hist0_sum_sq": [], "hist1_sum_sq": []}
    for j in range(images.shape[0]):
        hist1 = np.histogram(images[j][0].flatten(), bins=256)[0]
        hist2 = np.histogram(images[j][1].flatten(), bins=256)[0]
        if np.std(hist1) == 0 or np.std(hist2) == 0:
            hist_p = np.nan
        else:
            hist_p, p1 = pearsonr(hist1, hist2)
        img1_binned = np.digitize(images[j][0].flatten(),
                                  bins=np.linspace(images[j][0].min(), images[j][0].max(), 256))
        img2_binned = np.digitize(images[j][1].flatten(),
                                  bins=np.linspace(images[j][1].min(), images[j][1].max(), 256))
        out["nmi"].append(float(normalized_mutual_info_score(img1_binned, img2_binned)))
        out["rmse"].append(float(np.sqrt(np.mean((images[j][0] - images[j][1]) ** 2))))
        out["hist_corr"].append(float(hist_p))
        out["hist0_head"].append([int(v) for v in hist1[:16]])
        out["hist1_head"].append([int(v) for v in hist2[:16]])
        out["hist0_sum_sq"].append(int((hist1.astype(np.int64) ** 2).sum()))     # checksum over all 256 bins
        out["hist1_sum_sq"].append(int((hist2.astype(np.int64) ** 2).sum()))
    return out


def main():
    tiles = np.load(os.path.join(HERE, "tiles.npz"))["tiles"].astype(np.float32)
    xn = np.stack([np.stack([orc.normalize_image(t[0]), orc.normalize_image(t[1])]) for t in tiles])
    import scipy
    import sklearn
    gold = {"numpy": np.__version__, "scipy": scipy.__version__, "sklearn": sklearn.__version__, "normalised": metrics(xn), "raw": metrics(tiles)}
    json.dump(gold, open(os.path.join(HERE, "metrics.json"), "w"), indent=1)
    print(json.dumps(gold)[:600])


if __name__ == "__main__":
    main()
