"""GPU helper: where the tcgen05 self-attention kernel's softmax warps spend their cycles (debug build)."""
import ctypes as C, os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import evrp_b200
from oracle import evrp_oracle as O
N, B = int(sys.argv[1]) if len(sys.argv) > 1 else 100, int(sys.argv[2]) if len(sys.argv) > 2 else 8192
L = evrp_b200.lib()
L.evrp_debug_attn_cycles.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
W = dict(np.load("tests/golden/weights_seed0.npz"))
a = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=N, charging_num=5)
m = evrp_b200.AttentionModel(128, 128, a, n_encode_layers=3)
m.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}); m = m.cuda().eval()
s, d, e = O.generate_instances(B, N, seed=1000)
st, dy = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
with torch.no_grad():
    m.encode(st, dy); torch.cuda.synchronize()
    buf = (C.c_ulonglong * 8)(); L.evrp_debug_attn_cycles(buf, 1)
    m.encode(st, dy); torch.cuda.synchronize(); L.evrp_debug_attn_cycles(buf, 0)
v = np.array(list(buf), dtype=np.float64)
names = ["wait S (QK MMA + operand planes)", "pass 1 (row max)", "wait P free (previous P V)", "pass 2 (exp, split, TMEM store)",
         "wait O (last P V)", "output scaling + store", "-", "-"]
for n, x in zip(names, v):
    print(f"  {n:40s} {100 * x / v.sum():5.1f} %   {x / (3 * B * 8 / 148 * 148):9.0f} cycles/head")
