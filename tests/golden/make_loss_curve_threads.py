"""Does the reference reproduce ITSELF to 1 %?  Same loop as tests/golden/make_loss_curve.py, run with a different
number of CPU threads (different reduction order inside ATen), compared against the stored curve.  Golden generator: run
in the build container only (needs /root/reference), from the repo root:

    python tests/golden/make_loss_curve_threads.py <threads> <steps>      ->  tests/golden/loss_curve_single_threads<threads>.json
"""
import json, os, sys, time, torch
sys.path.insert(0, "/root/reference"); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "oracle"))
from regression_model import AdvancedRegressionModel
import crosstalk_oracle as orc
threads = int(sys.argv[1]); steps = int(sys.argv[2])
torch.set_num_threads(threads)
g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "loss_curve_single.json")))
x, y = orc.synthetic_batch(g["pool"], seed=g["data_seed"])
torch.manual_seed(0)
model = AdvancedRegressionModel(initial_filters=128, num_conv_blocks=6)
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
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), f"loss_curve_single_threads{threads}.json"), "w"))
