"""GPU helper: hash of the rollout outputs (pi, ll, R) at EM-EVRP-N, batch B, greedy and sampled; run under EVRP_TEAM=1 and
EVRP_TEAM=2: the digests must be identical (one warp or two warps per row decode bit-identically)."""
import hashlib, os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import evrp_b200
from evrp_b200 import policy_math as PM
from oracle import evrp_oracle as O
N, B = int(sys.argv[1]) if len(sys.argv) > 1 else 100, int(sys.argv[2]) if len(sys.argv) > 2 else 1024
W = dict(np.load("tests/golden/weights_seed0.npz"))
a = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=N, charging_num=5)
m = evrp_b200.AttentionModel(128, 128, a, n_encode_layers=3)
m.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}); m = m.cuda().eval()
m.sample_seed = 1234
s, d, e = O.generate_instances(B, N, seed=1000)
st, dy = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
D, G = PM.pairwise_matrices(st, torch.from_numpy(e).cuda())
with torch.no_grad():
    emb, kvl, fixed = m.encode(st, dy)
    for dt in ("greedy", "sample"):
        m._calls = 0
        o = m.rollout(emb, kvl, fixed, dy, D, G, decode_type=dt, save_trace=True)
        h = hashlib.sha256()
        for k in ("pi", "ll", "R", "mask_bits", "veh"):
            h.update(o[k].contiguous().cpu().numpy().tobytes())
        print(f"N={N} B={B} {dt}: T={o['pi'].shape[1]} cost {float(o['R'].sum(1).mean()):.4f} digest {h.hexdigest()[:24]}")
