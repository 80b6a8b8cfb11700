"""Does the reference reproduce ITSELF to 1 %?  Same loop as tests/golden/make_loss_curve.py, run with a different
number of CPU threads (different reduction order inside ATen), compared against the stored curve.  Golden generator: run
in the build container only (needs /root/reference), from the repo root:

    python tests/golden/make_loss_curve_threads.py <threads> <steps> [single|double]   ->  tests/golden/loss_curve_<kind>_threads<threads>.json
"""
import json, os, sys, time, torch
sys.path.insert(0, "/root/reference"); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "oracle"))
from regression_model import AdvancedRegressionModel
from two_branch_regression import SimplifiedTwoBranchRegressionModel
import crosstalk_oracle as orc
threads = int(sys.argv[1]); steps = int(sys.argv[2]); kind = sys.argv[3] if len(sys.argv) > 3 else "single"
torch.set_num_threads(threads)
g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), f"loss_curve_{kind}.json")))
x, y = orc.synthetic_batch(g["pool"], seed=g["data_seed"])
torch.manual_seed(0)
model = (AdvancedRegressionModel(initial_filters=128, num_conv_blocks=6) if kind == "single"
         else SimplifiedTwoBranchRegressionModel(initial_filters_per_branch=64))
opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)
crit = torch.nn.MSELoss(); model.train()
out = []
t0 = time.time()
for t in range(steps):
    s = (t * g["batch"]) % g["pool"]
    torch.manual_seed(g["seed0"] + t)
    opt.zero_grad(); loss = crit(model(x[s:s + g["batch"]]), y[s:s + g["batch"]]); loss.backward(); opt.step()
    out.append(loss.item())
    ref = g["reference_fp32"][t]
    print(f"step {t}: threads={threads} {out[-1]:.6f}  stored(8 threads) {ref:.6f}  rel {abs(out[-1]-ref)/ref:.2e}  ({time.time()-t0:.0f}s)", flush=True)
json.dump({"what": f"the reference's own {kind}-branch loop (make_loss_curve.py) run with torch.set_num_threads({threads}) instead of 8",
           "threads": threads, "steps": steps, "reference_fp32": out},
          open(os.path.join(os.path.dirname(os.path.abspath(__file__)), f"loss_curve_{kind}_threads{threads}.json"), "w"))
