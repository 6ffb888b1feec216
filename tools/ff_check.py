"""GPU helper: tcgen05 FF_tour rollout kernel (ff_mode 0) against the fp32 FFMA rollout kernel (ff_mode 1)."""
import os, sys, types, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import evrp_b200
from evrp_b200 import policy_math as PM
from oracle import evrp_oracle as O

W = dict(np.load("tests/golden/weights_seed0.npz"))
for N, B in ((20, 32), (20, 300), (50, 777), (100, 2048), (100, 8192)):
    a = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=N, charging_num=5)
    m = evrp_b200.AttentionModel(128, 128, a, n_encode_layers=3)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}); m = m.cuda().eval(); m.set_decode_type("greedy")
    s, d, e = O.generate_instances(B, N, seed=77 + N)
    st, dy = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
    D, G = PM.pairwise_matrices(st, torch.from_numpy(e).cuda())
    with torch.no_grad():
        emb, kvl, fixed = m.encode(st, dy)
        m.ff_mode = 1
        ref = m.rollout(emb, kvl, fixed, dy, D, G)
        torch.cuda.synchronize()
        m.ff_mode = 0
        t0 = time.perf_counter()
        forced = m.rollout(emb, kvl, fixed, dy, D, G, forced=ref["pi"].contiguous())
        torch.cuda.synchronize()
        print(f"N={N} B={B}: forced TC rollout done in {time.perf_counter() - t0:.3f}s status={forced['status']}", flush=True)
        dll = (forced["ll"] - ref["ll"]).abs().max().item()
        same_R = torch.equal(forced["R"], ref["R"])
        free = m.rollout(emb, kvl, fixed, dy, D, G)
        T = min(free["pi"].shape[1], ref["pi"].shape[1])
        same = (free["pi"][:, :T] == ref["pi"][:, :T]).all(1) & (free["steps"] == ref["steps"])
        ts = []
        for mode in (1, 0):
            m.ff_mode = mode
            m.rollout(emb, kvl, fixed, dy, D, G); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); m.rollout(emb, kvl, fixed, dy, D, G); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"   max |ll_tc - ll_ffma| (teacher-forced) {dll:.3e}  rewards bit-equal {same_R}  identical free tours "
              f"{int(same.sum())}/{B}  cost tc {free['R'].sum(1).mean():.4f} ffma {ref['R'].sum(1).mean():.4f}  "
              f"rollout ms: ffma {ts[0]:.2f} tc {ts[1]:.2f}", flush=True)
