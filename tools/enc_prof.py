"""GPU helper: wait-time profile of the fused encoder kernel (CTA 0), from the -DEVRP_FENC_PROF build.
usage: EVRP_B200_LIB=tools/_build/libevrp_prof.so python tools/enc_prof.py [N] [B]"""
import ctypes as C
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import evrp_b200
from evrp_b200 import _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
a = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=N, charging_num=5)
torch.manual_seed(0)
m = evrp_b200.AttentionModel(128, 128, a, n_encode_layers=3).cuda().eval()
s, d, e = evrp_b200.synthetic_instances(B, N, seed=9)
st, dy = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
L = _lib.lib()
L.evrp_encoder_fused_prof.restype = C.c_int
L.evrp_encoder_fused_prof.argtypes = [C.POINTER(C.c_ulonglong)]
with torch.no_grad():
    for _ in range(3):
        m.encode(st, dy)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m.encode(st, dy); e1.record(); torch.cuda.synchronize()
print(f"N={N} B={B}: encoder {e0.elapsed_time(e1):.3f} ms")
buf = (C.c_ulonglong * 300)()
assert L.evrp_encoder_fused_prof(buf) == 0
S = N + 6
ipt = 1 if S > 64 else 128 // S
tiles = -(-((B + ipt - 1) // ipt) // 148)      # tiles of CTA 0 (ceil)
names = {0: "producer", 1: "mma", 2: "worker0"}
for role in range(3):
    v = [int(buf[role * 100 + i]) for i in range(100)]
    tot = v[99]
    print(f"-- {names[role]}: total {tot} cycles = {tot / max(tiles,1):.0f} / tile; waits (cycles per tile, % of total):")
    for i in range(99):
        if v[i]:
            print(f"   site {i:2d}: {v[i] / max(tiles,1):9.0f}  {100.0 * v[i] / tot:5.1f}%")
    print(f"   not waiting: {(tot - sum(v[:99])) / max(tiles,1):9.0f}  {100.0 * (tot - sum(v[:99])) / tot:5.1f}%")
