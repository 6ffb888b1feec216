"""GPU debug helper: tcgen05 3xTF32 encoder path vs the fp32 FFMA path vs the oracle."""
import os, sys, types, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import evrp_b200
from oracle import evrp_oracle as O

name = sys.argv[1] if len(sys.argv) > 1 else "greedy_n20.npz"
g = np.load(f"tests/golden/{name}"); W = dict(np.load("tests/golden/weights_seed0.npz"))
N = g["static"].shape[2] - 6
a = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=N, charging_num=5)
m = evrp_b200.AttentionModel(128, 128, a, n_encode_layers=3)
m.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}); m = m.cuda().eval()
st, dy = torch.from_numpy(g["static"]).cuda(), torch.from_numpy(g["dynamic"]).cuda()
ref = O.encoder(W, O.init_embed(W, g["static"], g["dynamic"], 5))
out = {}
for mode in (1, 0):
    m.gemm_mode = mode
    emb, kvl, fixed = m.encode(st, dy)
    torch.cuda.synchronize()
    out[mode] = (emb.cpu().numpy(), kvl.cpu().numpy())
    print("mode", mode, "emb vs oracle", np.abs(out[mode][0] - ref).max(), "nan", np.isnan(out[mode][0]).sum(), flush=True)
print("tc vs ffma emb", np.abs(out[0][0] - out[1][0]).max(), "kvl", np.abs(out[0][1] - out[1][1]).max())
if len(sys.argv) > 2:
    B = int(sys.argv[2])
    s, d, e = O.generate_instances(B, 100, seed=1)
    st, dy = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
    a = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=100, charging_num=5)
    for mode in (1, 0):
        m.gemm_mode = mode
        for _ in range(2): m.encode(st, dy)
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(3): r = m.encode(st, dy)
        torch.cuda.synchronize(); print("mode", mode, "encode ms", (time.perf_counter() - t) / 3 * 1e3)
        out[mode] = r[0]
    print("B=%d tc vs ffma" % B, (out[0] - out[1]).abs().max().item())
