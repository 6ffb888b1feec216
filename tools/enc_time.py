"""GPU helper: encoder (init-embed + 3 layers + precompute) time at EM-EVRP-N, batch B, and agreement with the fp32 FFMA path."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import evrp_b200
from oracle import evrp_oracle as O
N, B = int(sys.argv[1]) if len(sys.argv) > 1 else 100, int(sys.argv[2]) if len(sys.argv) > 2 else 8192
W = dict(np.load("tests/golden/weights_seed0.npz"))
a = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=N, charging_num=5)
m = evrp_b200.AttentionModel(128, 128, a, n_encode_layers=3)
m.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}); m = m.cuda().eval()
s, d, e = O.generate_instances(B, N, seed=1000)
st, dy = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
with torch.no_grad():
    for mode, am in ((0, 0), (0, 1), (1, 1)):
        m.gemm_mode = mode; m.attn_mode = am
        for _ in range(2): out = m.encode(st, dy)
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = m.encode(st, dy); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"N={N} B={B} gemm_mode={mode} attn_mode={am}: encoder {min(ts):.2f} ms (median {sorted(ts)[2]:.2f})")
        if (mode, am) == (0, 0): ref = [t.clone() for t in out]
        else: print("   max |emb_tc - emb_ffma| %.3e   max |kvl| diff %.3e" % (float((out[0] - ref[0]).abs().max()), float((out[1] - ref[1]).abs().max())))
    # against the numpy oracle on a few instances
    emb_o = O.encode(W, s[:8], d[:8]) if hasattr(O, "encode") else None
    if emb_o is not None:
        eo = emb_o[0] if isinstance(emb_o, (tuple, list)) else emb_o
        print("   max |emb - oracle| %.3e" % float(np.abs(out[0][:8].cpu().numpy() - eo).max()))
