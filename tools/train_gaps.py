"""GPU helper: idle gaps of the GPU inside one REINFORCE step (torch.profiler device events): total busy / idle time and the
largest gaps with the kernels on either side -- shows where the host (launch latency, syncs) leaves the GPU waiting."""
import os, sys, types
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
for _ in range(3): T.reinforce_step(m, bl, opt, (st, dy, D, G))
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): T.reinforce_step(m, bl, opt, (st, dy, D, G))
    torch.cuda.synchronize()
ev = sorted([(x.time_range.start, x.time_range.end, x.name) for x in prof.events() if x.device_type == torch.autograd.DeviceType.CUDA])
t0, t1 = ev[0][0], max(x[1] for x in ev)
busy, gaps, cur_end, last = 0.0, [], ev[0][0], ""
for s_, e_, n in ev:
    if s_ > cur_end:
        gaps.append((s_ - cur_end, last[:50], n[:50]))
    busy += max(0.0, e_ - max(s_, cur_end))
    if e_ > cur_end:
        cur_end, last = e_, n
print(f"3 steps: span {(t1 - t0) / 1e3:.2f} ms, busy {busy / 1e3:.2f} ms, idle {(t1 - t0 - busy) / 1e3:.2f} ms in {len(gaps)} gaps")
agg = {}
for g, a_, b_ in gaps:
    k = (a_, b_); v = agg.setdefault(k, [0, 0.0]); v[0] += 1; v[1] += g
for (a_, b_), (n, g) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
    print(f"{g / 1e3:7.3f} ms x{n:<3d} after [{a_}] before [{b_}]")
