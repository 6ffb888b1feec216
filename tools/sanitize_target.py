"""compute-sanitizer target: one small pass of every tcgen05 kernel (fused encoder at 1 / 2 / 4 instances per tile, the
per-sub-layer tcgen05 GEMM + attention path, the persistent rollout kernel greedy and sampled) at batch 64."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import evrp_b200
from evrp_b200 import policy_math as PM
for N in (100, 50, 20):
    a = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=N, charging_num=5)
    torch.manual_seed(0)
    m = evrp_b200.AttentionModel(128, 128, a, n_encode_layers=3).cuda().eval()
    s, d, e = evrp_b200.synthetic_instances(64, N, seed=3)
    st, dy = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
    D, G = PM.pairwise_matrices(st, torch.from_numpy(e).cuda())
    with torch.no_grad():
        for mode in (2, 0):
            m.gemm_mode = mode
            emb, kvl, fixed = m.encode(st, dy)
        for dec in ("greedy", "sample"):
            m.set_decode_type(dec)
            o = m.rollout(emb, kvl, fixed, dy, D, G)
    torch.cuda.synchronize()
    print(f"N={N}: T={o['pi'].shape[1]} mean cost {float(o['R'].sum(1).mean()):.2f} status {o['status']}")
print("done")
