"""GPU helper (ncu target): a few launches of the encoder at EM-EVRP-N, batch B.  usage: enc_once.py [N] [B] [mode] [launches]"""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import evrp_b200
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 2
n = int(sys.argv[4]) if len(sys.argv) > 4 else 3
a = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=N, charging_num=5)
torch.manual_seed(0)
m = evrp_b200.AttentionModel(128, 128, a, n_encode_layers=3).cuda().eval()
m.gemm_mode = mode
s, d, e = evrp_b200.synthetic_instances(B, N, seed=9)
st, dy = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
with torch.no_grad():
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = m.encode(st, dy); e1.record(); torch.cuda.synchronize()
        print(f"encoder N={N} B={B} mode={mode}: {e0.elapsed_time(e1):.3f} ms")
