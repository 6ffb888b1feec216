"""GPU helper: cycle split of the persistent rollout kernel by phase (debug build with -DEVRP_PHASE_TIMING).

  bash tools/build_timing.sh && EVRP_B200_LIB=tools/_build/libevrp_b200_timing.so python tools/phase_timing.py [N] [B]
"""
import ctypes as C, os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import evrp_b200
from evrp_b200 import policy_math as PM
from oracle import evrp_oracle as O

N, B = int(sys.argv[1]) if len(sys.argv) > 1 else 100, int(sys.argv[2]) if len(sys.argv) > 2 else 8192
L = evrp_b200.lib()
FFM = int(os.environ.get("FF_MODE", "0"))
dbg = L.evrp_debug_phase_cycles if FFM else L.evrp_debug_phase_cycles_tc
dbg.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
W = dict(np.load("tests/golden/weights_seed0.npz"))
a = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=N, charging_num=5)
m = evrp_b200.AttentionModel(128, 128, a, n_encode_layers=3)
m.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}); m = m.cuda().eval(); m.set_decode_type("greedy"); m.ff_mode = FFM
s, d, e = O.generate_instances(B, N, seed=1000)
st, dy = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
D, G = PM.pairwise_matrices(st, torch.from_numpy(e).cuda())
with torch.no_grad():
    emb, kvl, fixed = m.encode(st, dy)
    for _ in range(2):
        o = m.rollout(emb, kvl, fixed, dy, D, G)
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 4)()
    dbg(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); o = m.rollout(emb, kvl, fixed, dy, D, G); e1.record(); torch.cuda.synchronize()
    dbg(buf, 0)
v = np.array(list(buf), dtype=np.float64)
names = ["FF1", "FF2", "row phase (warp 0 busy)", "barrier wait + veh"] if FFM else ["FF: hidden tiles (wait + convert)", "FF: ctx wait + store + sync", "row phase (warp 0 busy)", "barrier wait (other warps' rows)"]
print(f"N={N} B={B} T={o['pi'].shape[1]} kernel {e0.elapsed_time(e1):.2f} ms")
for n, x in zip(names, v):
    print(f"  {n:28s} {100 * x / v.sum():5.1f} %")
if not FFM:
    rb = (C.c_ulonglong * 8)()
    L.evrp_debug_row_cycles_tc.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    L.evrp_debug_row_cycles_tc(rb, 0)
    r = np.array(list(rb), dtype=np.float64)
    rn = ["list + K prefetch + query", "compat pass (K)", "softmax (+V prefetch)", "heads pass (V, + logit-K prefetch)",
          "logits pass (logit-K)", "log-softmax + pick", "transition loads + state + running max", "look-ahead mask + bookkeeping"]
    fb = (C.c_ulonglong * 12)()
    L.evrp_debug_ff_cycles_tc.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    L.evrp_debug_ff_cycles_tc(fb, 0)
    f = np.array(list(fb), dtype=np.float64)
    T = o['pi'].shape[1]
    ncta = min(148, (B + 63) // 64 * 148 // 148) if False else None
    fn = ["wait tile 0", "convert 0", "wait tile 1", "convert 1", "wait tile 2", "convert 2", "wait tile 3", "convert 3",
          "ctx wait", "ctx store", "sync", "-"]
    print("  FF phase of warp 0 (share of the FF phase; cycles per CTA-step over all CTAs):")
    for n, x in zip(fn[:11], f[:11]):
        print(f"    {n:14s} {100 * x / f[:11].sum():5.1f} %")
    print("  row phase of warp 0 by pass:")
    for n, x in zip(rn, r):
        print(f"    {n:40s} {100 * x / r.sum():5.1f} %")
