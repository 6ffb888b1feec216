"""GPU helper: where the end-to-end (host buffers in / out) rollout call spends its time beyond the kernels."""
import os, sys, types, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import evrp_b200
from evrp_b200 import policy_math as PM
from oracle import evrp_oracle as O
N, B = 100, 8192
W = dict(np.load("tests/golden/weights_seed0.npz"))
a = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=N, charging_num=5)
m = evrp_b200.AttentionModel(128, 128, a, n_encode_layers=3)
m.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}); m = m.cuda().eval(); m.set_decode_type("greedy")
s, d, e = O.generate_instances(B, N, seed=1000)
st_h, dy_h = torch.from_numpy(s).pin_memory(), torch.from_numpy(d).pin_memory()
st, dy = st_h.cuda(), dy_h.cuda()
D, G = PM.pairwise_matrices(st, torch.from_numpy(e).cuda())
host = {}
def sync(): torch.cuda.synchronize(); return time.perf_counter()
def step(timing=None):
    t0 = sync()
    a_, b_ = st_h.to("cuda", non_blocking=True), dy_h.to("cuda", non_blocking=True)
    t1 = sync()
    with torch.no_grad():
        emb, kvl, fixed = m.encode(a_, b_)
        t2 = sync()
        o = m.rollout(emb, kvl, fixed, b_, D, G)
        t3 = sync()
        pi, cost = o["pi"], o["R"].sum(1)
        if "pi" not in host:
            host["pi"] = torch.empty(pi.shape, dtype=pi.dtype).pin_memory(); host["c"] = torch.empty(cost.shape).pin_memory()
        host["pi"].copy_(pi, non_blocking=True); host["c"].copy_(cost, non_blocking=True)
    t4 = sync()
    if timing is not None: timing.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
for _ in range(3): step()
T = []
for _ in range(5): step(T)
T = np.array(T) * 1e3
print("per-phase ms (H2D, encoder, rollout incl. alloc + t_max sync, D2H):", T.mean(0).round(2), "sum", T.mean(0).sum().round(2))
t0 = sync()
for _ in range(5):
    with torch.no_grad():
        pi, ll, R = m((st_h, dy_h, D, G))
        cost = R.sum(1)
        host["pi"].copy_(pi, non_blocking=True); host["c"].copy_(cost, non_blocking=True)
    torch.cuda.synchronize()
print("public-API e2e step: %.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
