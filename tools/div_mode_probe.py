"""GPU helper: run the patched reference on CUDA (oracle/_ref) and teacher-force the stand-alone step kernels on its tours
with both scalar-division conventions; prints the number of steps whose state / reward / mask differ."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.utils.data import DataLoader
import evrp_b200
from oracle import ref_shim, make_golden as MG

for N, B, seed in ((20, 64, 12345), (50, 32, 777), (100, 16, 4242)):
    ds, args = ref_shim.make_dataset(B, N, seed)
    m = ref_shim.make_model(ds, args, 0)
    import nets.DRLModel as RM
    m = m.to(RM.device).eval(); m.set_decode_type("greedy")
    bat = next(iter(DataLoader(ds, B, False)))
    tr = MG.Tracer(m)
    # Tracer.ud indexes with a CPU arange: patch for CUDA tensors
    def ud(dynamic, distances, slope, now, chosen, tr=tr):
        nd, r = tr._ud(dynamic, distances, slope, now, chosen)
        tr.dyns.append(nd.clone()); return nd, r
    m.update_dynamic = ud
    with torch.no_grad():
        pi, ll, R = m(bat)
    tr.close()
    print(f"N={N} B={B}: reference on {RM.device}, T={pi.shape[1]}, mean cost {float(R.sum(1).mean()):.4f}")
    t_dyn = torch.stack(tr.dyns, 1); t_mask = torch.stack(tr.masks, 1)
    ours = evrp_b200.VehicleRoutingDataset(2, N, 10., 80., 50., 4, 4, 5, 1, args)
    static, dynamic, D, G = [t.cuda().float() for t in bat]
    for mode in (0, 1):
        ours.div_mode = mode
        dyn = dynamic.clone(); now = torch.zeros(B, dtype=torch.int64, device="cuda")
        bad_s = bad_r = bad_m = 0
        for t in range(pi.shape[1]):
            nd, r = ours.update_dynamic(dyn, D, G, now, pi[:, t])
            bad_s += int(not torch.equal(nd, t_dyn[:, t])); bad_r += int(not torch.equal(r[:, 0], R[:, t]))
            if t + 1 < pi.shape[1]:
                mk = ours.update_mask(t_dyn[:, t], D, G, pi[:, t])
                bad_m += int(not torch.equal(mk, t_mask[:, t + 1]))
            dyn, now = t_dyn[:, t], pi[:, t]
        print(f"   div_mode {mode}: steps with state mismatch {bad_s}, reward mismatch {bad_r}, mask mismatch {bad_m} of {pi.shape[1]}")
