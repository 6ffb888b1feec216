"""GPU helper: where does a REINFORCE step spend its time? (EM-EVRP-100, batch 1024)"""
import os, sys, types, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import evrp_b200
from evrp_b200 import policy_math as PM, train as T
from oracle import evrp_oracle as O

N, B = int(sys.argv[1]) if len(sys.argv) > 1 else 100, int(sys.argv[2]) if len(sys.argv) > 2 else 1024
W = dict(np.load("tests/golden/weights_seed0.npz"))
a = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=N, charging_num=5)
m = evrp_b200.AttentionModel(128, 128, a, n_encode_layers=3)
m.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}); m = m.cuda().train(); m.set_decode_type("sample")
s, d, e = O.generate_instances(B, N, seed=5)
st, dy = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
D, G = PM.pairwise_matrices(st, torch.from_numpy(e).cuda())
opt = torch.optim.Adam(m.parameters(), lr=1e-4)

def timed(fn, n=3):
    fn(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3, r

t_enc, (emb, kvl, fixed) = timed(lambda: PM.encode_torch(m, st, dy))
with torch.no_grad():
    t_roll, o = timed(lambda: m.rollout(emb.detach().contiguous(), kvl.detach().contiguous(), fixed.detach().contiguous(), dy, D, G, save_trace=True))
t_ll, ll = timed(lambda: PM.teacher_forced_ll(m, emb, kvl, fixed, o["pi"], o["mask_bits"], o["veh"], o["steps"]))
def fb():
    emb, kvl, fixed = PM.encode_torch(m, st, dy)
    ll = PM.teacher_forced_ll(m, emb, kvl, fixed, o["pi"], o["mask_bits"], o["veh"], o["steps"])
    loss, _, _ = T.ReinforceLoss.apply(ll, o["R"], torch.tensor(4000.0).cuda(), 1.0 / B)
    opt.zero_grad(); loss.backward(); return loss
t_fb, _ = timed(fb)
t_step, _ = timed(lambda: T.reinforce_step(m, T.ExponentialBaseline(0.8), opt, (st, dy, D, G)))
print(f"N={N} B={B} T={o['pi'].shape[1]}: encode_torch fwd {t_enc:.1f} ms | rollout kernel {t_roll:.1f} ms | ll fwd {t_ll:.1f} ms | "
      f"enc+ll fwd+bwd {t_fb:.1f} ms | full step {t_step:.1f} ms | peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
