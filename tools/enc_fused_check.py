"""GPU helper: the fused encoder kernel (gemm_mode 2) against the fp32 FFMA path (gemm_mode 1) for 0..3 layers, three
graph sizes (one / two / four instances per 128-row tile) and ragged batch sizes; then its time at B = 8192.
usage: python tools/enc_fused_check.py [quick]"""
import ctypes as C
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import evrp_b200
from evrp_b200 import _lib

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"


def dbg():
    out = (C.c_int * 8)()
    ab = _lib.lib().evrp_encoder_fused_debug(out, 1)
    return ab, list(out)


def model(N, L, seed=0):
    a = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=N, charging_num=5)
    torch.manual_seed(seed)
    m = evrp_b200.AttentionModel(128, 128, a, n_encode_layers=L).cuda().eval()
    # BatchNorm away from the identity so that the folded affine is exercised
    g = torch.Generator().manual_seed(1)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.running_mean.copy_(torch.randn(128, generator=g) * 0.3)
            mod.running_var.copy_(torch.rand(128, generator=g) + 0.5)
            mod.weight.data.copy_(torch.rand(128, generator=g) + 0.5)
            mod.bias.data.copy_(torch.randn(128, generator=g) * 0.2)
    return m


bad = 0
with torch.no_grad():
    for N, B in ((100, 300), (50, 301), (20, 1027), (100, 1), (122, 40), (3, 77)):
        s, d, e = evrp_b200.synthetic_instances(B, N, seed=5 + N)
        st, dy = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
        for L in ((0, 1, 3) if quick else (0, 1, 2, 3)):
            m = model(N, L)
            m.gemm_mode = 1
            ref = [t.clone() for t in m.encode(st, dy)]
            m.gemm_mode = 2
            out = m.encode(st, dy)
            torch.cuda.synchronize()
            ab, info = dbg()
            errs = [float((o - r).abs().max()) for o, r in zip(out, ref)]
            scale = [float(r.abs().max()) for r in ref]
            nan = any(bool(torch.isnan(o).any()) for o in out)
            ok = (not ab) and (not nan) and errs[0] < 5e-5 * max(1.0, scale[0]) and errs[1] < 5e-5 * max(1.0, scale[1])
            bad += not ok
            print(f"N={N:3d} S={N + 6:3d} B={B:4d} layers={L}: max|d emb| {errs[0]:.2e} (|emb| {scale[0]:.1f})  max|d kvl| {errs[1]:.2e} "
                  f"(|kvl| {scale[1]:.1f})  max|d fixed| {errs[2]:.2e}  nan={nan} abort={ab} {info[:4] if ab else ''}  {'ok' if ok else 'FAIL'}")
            if ab:
                sys.exit(2)
    # timing at the headline size
    for N, B in ((100, 8192), (50, 8192), (20, 8192)):
        s, d, e = evrp_b200.synthetic_instances(B, N, seed=9)
        st, dy = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
        m = model(N, 3)
        for mode in (2, 0):
            m.gemm_mode = mode
            for _ in range(3):
                m.encode(st, dy)
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); m.encode(st, dy); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            print(f"N={N} B={B} gemm_mode={mode}: encoder {min(ts):.3f} ms (median {sorted(ts)[2]:.3f})")
print("FAILURES:", bad)
sys.exit(1 if bad else 0)
