import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import evrp_b200
from evrp_b200 import policy_math as PM
from oracle import evrp_oracle as O
a = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=100, charging_num=5)
torch.manual_seed(0)
m = evrp_b200.AttentionModel(128, 128, a, n_encode_layers=0).cuda().eval()
for B in (8, 64, 512, 2048, 8192):
    s, d, e = O.generate_instances(B, 100, seed=1)
    st, dy = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
    for mode in (0,):
        m.gemm_mode = mode
        emb, kvl, fixed = m.encode(st, dy)
        ref = emb.double() @ PM.folded_node_projection(m).double().t()
        err = (kvl.double() - ref).abs()
        rows = err.view(-1, 384).max(1)[0]
        bad = (rows > 1e-3).nonzero()[:, 0]
        print(B, "rows", B * 106, "maxerr", err.max().item(), "bad rows", len(bad), (bad[:6].tolist(), bad[-3:].tolist()) if len(bad) else "")
        if len(bad):
            cols = (err.view(-1, 384)[bad[0]] > 1e-3).nonzero()[:, 0]
            print("  first bad row cols", cols[:8].tolist(), "n", len(cols), "tile", int(bad[0]) // 128, "row in tile", int(bad[0]) % 128)
            tiles = torch.unique(bad // 128)
            print("  bad tiles", len(tiles), tiles[:10].tolist())
