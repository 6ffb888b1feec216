"""GPU helper: does the training path LEARN?  EM-EVRP-N from random initialisation, REINFORCE with the exponential baseline
(trainer.py:100-143 protocol through evrp_b200.train.reinforce_step), fresh instances every step; prints the mean sampled
cost per block of steps and the greedy cost on a fixed validation set before / after.
usage: python tools/train_curve.py [N] [batch] [steps] [out.md]"""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import evrp_b200
from evrp_b200 import policy_math as PM, train as T
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
STEPS = int(sys.argv[3]) if len(sys.argv) > 3 else 400
a = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=N, charging_num=5)
torch.manual_seed(1234)
ds = evrp_b200.VehicleRoutingDataset(2, N, 10., 80., 50., 4, 4, 5, 1, a)
m = evrp_b200.AttentionModel(128, 128, a, n_encode_layers=3, update_dynamic=ds.update_dynamic, update_mask=ds.update_mask).cuda()
opt = torch.optim.Adam(m.parameters(), lr=1e-4)
bl = T.ExponentialBaseline(0.8)


def batch(seed, n):
    s, d, e = evrp_b200.synthetic_instances(n, N, seed=seed)
    st, dy = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
    D, G = PM.pairwise_matrices(st, torch.from_numpy(e).cuda())
    return st, dy, D, G


def greedy_cost(val):
    m.eval(); m.set_decode_type("greedy")
    with torch.no_grad():
        _, _, R = m(val)
    return float(R.sum(1).mean())


val = batch(999_999, 2048)
lines = [f"EM-EVRP-{N}, batch {B}, Adam lr 1e-4, exponential baseline (beta 0.8), max grad norm 2, fresh instances per step, one B200", ""]
c0 = greedy_cost(val)
lines.append(f"greedy cost on 2,048 fixed validation instances before training: {c0:.2f}")
lines += ["", "| steps | mean sampled cost | mean loss | steps/s |", "|---|---|---|---|"]
m.train(); m.set_decode_type("sample")
costs, losses, t0 = [], [], time.time()
for step in range(1, STEPS + 1):
    out = T.reinforce_step(m, bl, opt, batch(step, B), max_grad_norm=2.0)
    costs.append(out["reward"]); losses.append(out["loss"])
    if step % 50 == 0:
        torch.cuda.synchronize()
        c = float(torch.stack(costs).mean()); l = float(torch.stack(losses).mean())
        lines.append(f"| {step - 49}-{step} | {c:.2f} | {l:.3f} | {50 / (time.time() - t0):.1f} |")
        costs, losses, t0 = [], [], time.time()
        m.train(); m.set_decode_type("sample")
c1 = greedy_cost(val)
lines += ["", f"greedy cost on the same validation instances after {STEPS} steps: {c1:.2f} ({100 * (c1 - c0) / c0:+.1f} %)"]
text = "\n".join(lines)
print(text)
if len(sys.argv) > 4:
    open(sys.argv[4], "w").write(text + "\n")
