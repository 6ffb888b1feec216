"""GPU helper: bytes the rollout really needs -- records of the FEASIBLE nodes only (the kernel skips masked nodes)."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import evrp_b200
from evrp_b200 import policy_math as PM
N, B = 100, 8192
a = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=N, charging_num=5)
torch.manual_seed(0)
m = evrp_b200.AttentionModel(128, 128, a, n_encode_layers=3).cuda().eval(); m.set_decode_type("greedy")
s, d, e = evrp_b200.synthetic_instances(B, N, seed=1000)
st, dy = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
D, G = PM.pairwise_matrices(st, torch.from_numpy(e).cuda())
with torch.no_grad():
    emb, kvl, fixed = m.encode(st, dy)
    o = m.rollout(emb, kvl, fixed, dy, D, G, save_trace=True)
S = N + 6
T = o["pi"].shape[1]
live = torch.arange(T, device="cuda")[None] < o["steps"][:, None]
nf = PM.bits_to_mask(o["mask_bits"], S).sum(-1) * live
steps = int(o["steps"].sum())
rec = int(nf.sum()) * 1536
other = steps * (6 * S * 4 + 2 * 512 + 512 + 2 * 512 + 2 * 512)      # D/G rows + rule-7 rows, emb[now], emb[c], fixed, demand + running-max rows r/w
print(f"instance-steps {steps}, mean feasible nodes per step {float(nf.sum()) / steps:.2f} of {S}; records {rec/1e9:.1f} GB + "
      f"state/rows {other/1e9:.1f} GB = {(rec+other)/1e9:.1f} GB needed per launch (all-node model: {steps*(1568*S+2048)/1e9:.1f} GB)")
