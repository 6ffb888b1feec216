"""GPU helper: rollout kernel time with [B,S,S] matrices (4-tuple input) vs in-kernel geometry (3-tuple input)."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import evrp_b200
from evrp_b200 import policy_math as PM
from evrp_b200.model import _Elevations
N, B = int(sys.argv[1]) if len(sys.argv) > 1 else 100, int(sys.argv[2]) if len(sys.argv) > 2 else 8192
a = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=N, charging_num=5)
torch.manual_seed(0)
m = evrp_b200.AttentionModel(128, 128, a, n_encode_layers=3).cuda().eval(); m.set_decode_type("greedy")
s, d, e = evrp_b200.synthetic_instances(B, N, seed=1000)
st, dy, el = (torch.from_numpy(x).cuda() for x in (s, d, e))
with torch.no_grad():
    emb, kvl, fixed = m.encode(st, dy)
    for label in ("matrices", "coords"):
        ts = []
        for it in range(6):
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            if label == "matrices":
                D, G = PM.pairwise_matrices(st, el)
                e1.record()
                o = m.rollout(emb, kvl, fixed, dy, D, G)
            else:
                e1.record()
                o = m.rollout(emb, kvl, fixed, dy, None, _Elevations(el), static=st)
            e2.record(); torch.cuda.synchronize()
            ts.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
        ts = ts[2:]
        print(f"N={N} B={B} {label:9s}: geometry build {min(t[0] for t in ts):.2f} ms, rollout {min(t[1] for t in ts):.2f} ms, "
              f"T={o['pi'].shape[1]}, mean cost {float(o['R'].sum(1).mean()):.3f}")
