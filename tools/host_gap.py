"""GPU helper: where does a bench step's time go between the kernels?  Host timestamps + CUDA events around each phase
of one resident step (encode -> output allocation -> rollout launch -> status read-back)."""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import evrp_b200
from evrp_b200 import _lib, policy_math as PM

N, B = 100, 8192
a = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=N, charging_num=5)
ds = evrp_b200.VehicleRoutingDataset(2, N, 10., 80., 50., 4, 4, 5, 1, a)
torch.manual_seed(0)
m = evrp_b200.AttentionModel(128, 128, a, n_encode_layers=3, update_dynamic=ds.update_dynamic, update_mask=ds.update_mask).cuda().eval()
m.set_decode_type("greedy")
s, d, e = evrp_b200.synthetic_instances(B, N, seed=1000)
st, dy = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
D, G = PM.pairwise_matrices(st, torch.from_numpy(e).cuda())

def step(stamps):
    with torch.no_grad():
        stamps.append(("start", time.perf_counter()))
        emb, kvl, fixed = m.encode(st, dy)
        stamps.append(("encode launched", time.perf_counter()))
        o = m.rollout(emb, kvl, fixed, dy, D, G)
        stamps.append(("rollout returned (synced)", time.perf_counter()))
    return o

for _ in range(4):
    step([])
torch.cuda.synchronize()
K = 8
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
m.profile_events = []
all_stamps = []
t0 = time.perf_counter(); e0.record()
for _ in range(K):
    st_ = []; step(st_); all_stamps.append(st_)
e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"wall {1e3*(t1-t0)/K:.2f} ms/step, events {e0.elapsed_time(e1)/K:.2f} ms/step, rollout kernel "
      f"{np.mean([s.elapsed_time(e) for s, e in m.profile_events]):.2f} ms")
enc_host = np.mean([s[1][1] - s[0][1] for s in all_stamps]) * 1e3
roll_host = np.mean([s[2][1] - s[1][1] for s in all_stamps]) * 1e3
print(f"host: encode() returns after {enc_host:.2f} ms, rollout() (alloc + launch + sync) {roll_host:.2f} ms")
# finer: time the pieces of rollout() on the host with the GPU idle
torch.cuda.synchronize()
import cProfile, pstats, io
pr = cProfile.Profile(); pr.enable()
for _ in range(4):
    step([])
pr.disable()
sio = io.StringIO(); pstats.Stats(pr, stream=sio).sort_stats("cumulative").print_stats(18); print(sio.getvalue()[:3500])
