"""GPU helper: greedy-tour agreement between a tensor-core encoder (gemm_mode 2 = the fused tcgen05 kernel, the default; 0 =
per-sub-layer tcgen05 3xTF32) and the fp32 FFMA encoder (gemm_mode 1), with a near-tie audit of every first divergence
(log-prob gap under the FFMA model).  usage: tour_agreement.py [N] [B] [mode]"""
import os, sys, types, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import evrp_b200
from oracle import evrp_oracle as O

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
MODE = int(sys.argv[3]) if len(sys.argv) > 3 else 2
W = dict(np.load("tests/golden/weights_seed0.npz"))
a = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=N, charging_num=5)
m = evrp_b200.AttentionModel(128, 128, a, n_encode_layers=3)
m.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}); m = m.cuda().eval(); m.set_decode_type("greedy")
s, d, e = O.generate_instances(B, N, seed=77)
st, dy = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
D, G = evrp_b200.policy_math.pairwise_matrices(st, torch.from_numpy(e).cuda())
res = {}
with torch.no_grad():
    for mode in (1, MODE):
        m.gemm_mode = mode
        res[mode] = m((st, dy, D, G))
    pi1, ll1, R1 = res[1]; pi0, ll0, R0 = res[MODE]
    T = max(pi0.shape[1], pi1.shape[1])
    pad = lambda x: torch.nn.functional.pad(x, (0, T - x.shape[1]))
    p0, p1 = pad(pi0), pad(pi1)
    same = (p0 == p1).all(1)
    print(f"N={N} B={B} gemm_mode {MODE} vs 1: identical tours {int(same.sum())}/{B}; mean cost ffma {R1.sum(1).mean():.4f} tc {R0.sum(1).mean():.4f}")
    m.gemm_mode = 1
    idx = torch.nonzero(~same)[:, 0]
    gaps = []
    if len(idx):
        first = ((p0 != p1)[idx]).float().argmax(1)                    # first divergent step
        # log p_ffma(tc's choice) at that step: teacher-force tc's tour prefix (+ its choice) under the ffma model;
        # after the divergent step continue with tc's own tour (feasible under identical dynamics)
        sub = (st[idx], dy[idx], D[idx], G[idx])
        o = m.forward_forced(sub, p0[idx][:, :pi0.shape[1]].contiguous())
        ll_forced = pad(o["ll"])
        own = pad(ll1)[idx]
        gaps = (own.gather(1, first[:, None]) - ll_forced.gather(1, first[:, None]))[:, 0].cpu().numpy()
        print("first-divergence log-prob gaps under the ffma model: max %.3e median %.3e; >1e-4: %d" %
              (gaps.max(), np.median(gaps), int((gaps > 1e-4).sum())))
    print(json.dumps({"N": N, "B": B, "gemm_mode": MODE, "identical": int(same.sum()), "max_gap": float(max(gaps) if len(gaps) else 0)}))
