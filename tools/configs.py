"""GPU helper: the BASELINE.json configs that bench.py's headline line does not cover, one markdown row each.

  configs[1]  DM-EVRP-20 REINFORCE training epoch with greedy rollout baseline, batch 512 (a bounded slice of the
              epoch: wrap_dataset greedy pass over the slice + the REINFORCE steps; extrapolated to 1.28 M instances)
  configs[2]  EM-EVRP-50 sampled rollout, 1280 samples per instance, batch 1024 (sample_many)
  extra       nets/AM.py variant, EM-EVRP-100 greedy, batch 8192 (the bench.py workload on the AM kernel)

usage: python tools/configs.py [out.md]
"""
import os
import sys
import time
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import evrp_b200
from evrp_b200 import policy_math as PM
from evrp_b200 import train as T

rows = []


def args_for(N, obj="EM"):
    return types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=N, charging_num=5,
                                 velocity=50., obj=obj)


def build(cls, N, obj="EM"):
    a = args_for(N, obj)
    ds = evrp_b200.VehicleRoutingDataset(2, N, 10., 80., 50., 4, 4, 5, 1, a, objective=obj)
    torch.manual_seed(0)
    m = cls(128, 128, a, n_encode_layers=3, update_dynamic=ds.update_dynamic, update_mask=ds.update_mask)
    return m.cuda(), a


def instances(B, N, seed):
    s, d, e = evrp_b200.synthetic_instances(B, N, seed=seed)
    st, dy = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
    D, G = PM.pairwise_matrices(st, torch.from_numpy(e).cuda())
    return st, dy, D, G


def config2(n_steps=64, batch=512):
    N = 20
    m, a = build(evrp_b200.AttentionModel, N, "DM")
    train = evrp_b200.VehicleRoutingDataset(n_steps * batch, N, 10., 80., 50., 4, 4, 5, 12346, a, objective="DM")
    valid = evrp_b200.VehicleRoutingDataset(2048, N, 10., 80., 50., 4, 4, 5, 12348, a, objective="DM")
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    bargs = types.SimpleNamespace(batch_size=batch, bl_alpha=0.05)
    bl = T.RolloutBaseline(m, valid, bargs)
    T.train_epoch(m, bl, opt, torch.utils.data.Subset(train, range(4 * batch)), batch)          # warm-up
    torch.cuda.synchronize(); t = time.perf_counter()
    r, l, steps = T.train_epoch(m, bl, opt, train, batch)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    rows.append(f"| configs[1] DM-EVRP-20 REINFORCE epoch slice, rollout baseline, batch {batch} | {steps} steps + greedy "
                f"baseline pass over {steps * batch} instances in {dt:.2f} s | **{steps / dt:.1f} train step/s** "
                f"(baseline pass included) | full 1.28 M-instance epoch = {1.28e6 / (steps * batch) * dt:.0f} s; mean reward {r:.2f} |")


def config3(B=1024, R=1280, chunk=128):
    N = 50
    m, a = build(evrp_b200.AttentionModel, N)
    m.eval(); m.set_decode_type("sample"); m.max_steps = 400
    st, dy, D, G = instances(B, N, 9)

    def run():
        return torch.cat([m.sample_many((st[i:i + chunk], dy[i:i + chunk], D[i:i + chunk], G[i:i + chunk]),
                                        batch_rep=R, iter_rep=1)[0] for i in range(0, B, chunk)])
    run(); torch.cuda.synchronize(); t = time.perf_counter()
    c = run(); torch.cuda.synchronize(); dt = time.perf_counter() - t
    m.set_decode_type("greedy")
    with torch.no_grad():
        _, _, Rg = m((st, dy, D, G))
    rows.append(f"| configs[2] EM-EVRP-50 sampled, {R} samples x {B} instances | {B * R} rollouts in {dt:.2f} s | "
                f"**{B * R / dt / 1e3:.0f} k rollouts/s** = {B / dt:.0f} instances/s | best-of-{R} cost {float(c.mean()):.1f} "
                f"vs greedy {float(Rg.sum(1).mean()):.1f}; peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB |")


def am_greedy(B=8192, N=100, steps=5):
    m, a = build(evrp_b200.AttentionModelAM, N)
    m.eval(); m.set_decode_type("greedy")
    st, dy, D, G = instances(B, N, 1000)
    with torch.no_grad():
        for _ in range(3):
            m((st, dy, D, G))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m.profile_events = []
        e0.record()
        for _ in range(steps):
            pi, ll, R = m((st, dy, D, G))
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    k = sum(s.elapsed_time(e) for s, e in m.profile_events) / steps
    rows.append(f"| nets/AM.py variant, EM-EVRP-100 greedy, batch {B} | {ms:.1f} ms per batch (rollout kernel {k:.1f} ms) | "
                f"**{B / ms:.0f} k instances/s** | T = {pi.shape[1]}, mean cost {float(R.sum(1).mean()):.1f} |")


if __name__ == "__main__":
    which = os.environ.get("CONFIGS", "2,3,am").split(",")
    if "2" in which:
        config2()
    if "3" in which:
        config3()
    if "am" in which:
        am_greedy()
    out = "| workload | measured | rate | notes |\n|---|---|---|---|\n" + "\n".join(rows) + "\n"
    print(out)
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            f.write(f"one B200, {torch.cuda.get_device_name()}, wall clock around the public API calls after one warm-up pass "
                    f"(tools/configs.py)\n\n" + out)
