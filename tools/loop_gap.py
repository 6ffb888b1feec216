"""GPU helper: which part of bench.py's device-timed loop costs time beyond the kernels?"""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import evrp_b200
from evrp_b200 import policy_math as PM
N, B = 100, 8192
a = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=N, charging_num=5)
torch.manual_seed(0)
m = evrp_b200.AttentionModel(128, 128, a, n_encode_layers=3).cuda().eval(); m.set_decode_type("greedy")
s, d, e = evrp_b200.synthetic_instances(B, N, seed=1000)
st, dy = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
D, G = PM.pairwise_matrices(st, torch.from_numpy(e).cuda())
def run(events, acc, K=8):
    enc_events = []
    m.profile_events = [] if events else None
    tot = torch.zeros((), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        with torch.no_grad():
            if events:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)); ev[0].record()
            emb, kvl, fixed = m.encode(st, dy)
            if events:
                ev[1].record(); enc_events.append(ev)
            o = m.rollout(emb, kvl, fixed, dy, D, G)
        if acc:
            tot += o["steps"].sum()
    e1.record(); torch.cuda.synchronize()
    k = np.mean([a_.elapsed_time(b_) for a_, b_ in m.profile_events]) if events else float("nan")
    en = np.mean([a_.elapsed_time(b_) for a_, b_ in enc_events]) if events else float("nan")
    m.profile_events = None
    return e0.elapsed_time(e1) / K, en, k
for _ in range(5):
    run(False, False, 1)
for events, acc in ((True, True), (False, False), (True, False), (False, True), (True, True), (False, False)):
    ms, en, k = run(events, acc)
    print(f"events={events} acc={acc}: {ms:.2f} ms/step (encoder {en:.2f} + rollout kernel {k:.2f})")
