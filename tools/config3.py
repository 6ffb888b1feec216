"""GPU helper: BASELINE configs[2] — EM-EVRP-50, sampled rollout, 1280 samples per instance, batch 1024 (sample_many)."""
import os, sys, types, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import evrp_b200
from evrp_b200 import policy_math as PM
from oracle import evrp_oracle as O
N, B, R = 50, int(sys.argv[1]) if len(sys.argv) > 1 else 1024, int(sys.argv[2]) if len(sys.argv) > 2 else 1280
W = dict(np.load("tests/golden/weights_seed0.npz"))
a = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=N, charging_num=5)
m = evrp_b200.AttentionModel(128, 128, a, n_encode_layers=3)
m.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}); m = m.cuda().eval(); m.set_decode_type("sample")
m.max_steps = 400
s, d, e = O.generate_instances(B, N, seed=9)
st, dy = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
D, G = PM.pairwise_matrices(st, torch.from_numpy(e).cuda())
chunk = 128                                     # instances per sample_many call (163,840 rollouts each)
def run():
    costs = []
    for i in range(0, B, chunk):
        sl = slice(i, i + chunk)
        c, p = m.sample_many((st[sl], dy[sl], D[sl], G[sl]), batch_rep=R, iter_rep=1)
        costs.append(c)
    return torch.cat(costs)
run(); torch.cuda.synchronize(); t = time.perf_counter(); c = run(); torch.cuda.synchronize(); dt = time.perf_counter() - t
m.set_decode_type("greedy")
with torch.no_grad(): _, _, Rg = m((st, dy, D, G))
print(f"config3: {B} instances x {R} samples = {B*R} rollouts in {dt:.2f} s -> {B*R/dt:.0f} rollouts/s, {B/dt:.1f} instances/s; "
      f"mean best-of-{R} cost {c.mean():.2f} vs greedy {Rg.sum(1).mean():.2f}; peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
