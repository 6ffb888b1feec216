"""GPU helper: top CUDA kernels of one REINFORCE step (torch.profiler)."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import evrp_b200
from evrp_b200 import policy_math as PM, train as T
from oracle import evrp_oracle as O
N, B = 100, 1024
W = dict(np.load("tests/golden/weights_seed0.npz"))
a = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=N, charging_num=5)
m = evrp_b200.AttentionModel(128, 128, a, n_encode_layers=3)
m.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}); m = m.cuda().train(); m.set_decode_type("sample")
s, d, e = O.generate_instances(B, N, seed=5)
st, dy = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
D, G = PM.pairwise_matrices(st, torch.from_numpy(e).cuda())
opt = torch.optim.Adam(m.parameters(), lr=1e-4)
bl = T.ExponentialBaseline(0.8)
for _ in range(2): T.reinforce_step(m, bl, opt, (st, dy, D, G))
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    T.reinforce_step(m, bl, opt, (st, dy, D, G)); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
