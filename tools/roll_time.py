"""GPU helper: rollout-kernel time at EM-EVRP-N, batch B (env knobs are read by the library)."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import evrp_b200
from evrp_b200 import policy_math as PM
from oracle import evrp_oracle as O
N, B = int(sys.argv[1]) if len(sys.argv) > 1 else 100, int(sys.argv[2]) if len(sys.argv) > 2 else 8192
W = dict(np.load("tests/golden/weights_seed0.npz"))
a = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=N, charging_num=5)
m = evrp_b200.AttentionModel(128, 128, a, n_encode_layers=3)
m.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}); m = m.cuda().eval(); m.set_decode_type("greedy")
m.ff_mode = int(os.environ.get("FF_MODE", "0"))
s, d, e = O.generate_instances(B, N, seed=1000)
st, dy = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
D, G = PM.pairwise_matrices(st, torch.from_numpy(e).cuda())
with torch.no_grad():
    emb, kvl, fixed = m.encode(st, dy)
    for _ in range(2): o = m.rollout(emb, kvl, fixed, dy, D, G)
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); o = m.rollout(emb, kvl, fixed, dy, D, G); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
print(f"N={N} B={B} ff_mode={m.ff_mode} PF_ROWS={os.environ.get('EVRP_PF_ROWS')} PF_LINES={os.environ.get('EVRP_PF_LINES')}: "
      f"rollout {min(ts):.2f} ms (median {sorted(ts)[2]:.2f}) cost {o['R'].sum(1).mean():.4f}")
