"""GPU helper: CUDA kernels of one REINFORCE step grouped by kernel name (torch.profiler, device time only)."""
import os, sys, types, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import evrp_b200
from evrp_b200 import policy_math as PM, train as T
N, B = 100, int(sys.argv[1]) if len(sys.argv) > 1 else 1024
a = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=N, charging_num=5)
torch.manual_seed(0)
m = evrp_b200.AttentionModel(128, 128, a, n_encode_layers=3).cuda().train(); m.set_decode_type("sample")
s, d, e = evrp_b200.synthetic_instances(B, N, seed=5)
st, dy = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
D, G = PM.pairwise_matrices(st, torch.from_numpy(e).cuda())
opt = torch.optim.Adam(m.parameters(), lr=1e-4)
bl = T.ExponentialBaseline(0.8)
for _ in range(2): T.reinforce_step(m, bl, opt, (st, dy, D, G))
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    T.reinforce_step(m, bl, opt, (st, dy, D, G)); torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        k = ev.name[:90]
        a_ = agg.setdefault(k, [0, 0.0]); a_[0] += 1; a_[1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
tot = sum(v[1] for v in agg.values())
print(f"total device time {tot/1e3:.2f} ms over {sum(v[0] for v in agg.values())} kernels")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:32]:
    print(f"{t/1e3:7.2f} ms {100*t/tot:5.1f}% x{n:<4d} {k}")
