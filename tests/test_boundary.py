"""CPU tests of the drop-in boundary and host logic (no compute kernel is launched here)."""
import ctypes
import os
import re
import types

import numpy as np
import pytest
import torch

import evrp_b200
from evrp_b200 import _lib, policy_math as PM
from oracle import evrp_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def args_for(N, F=5):
    return types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=N, charging_num=F)


def bn_perturb(sd, seed=7):
    rs = np.random.RandomState(seed)
    out = {}
    for k, v in sd.items():
        v = v.clone()
        if "normalizer.running_mean" in k:
            v = torch.from_numpy((rs.randn(*v.shape) * 0.2).astype(np.float32))
        elif "normalizer.running_var" in k:
            v = torch.from_numpy((rs.rand(*v.shape) * 1.5 + 0.5).astype(np.float32))
        elif "normalizer.weight" in k:
            v = torch.from_numpy((rs.rand(*v.shape) + 0.5).astype(np.float32))
        elif "normalizer.bias" in k:
            v = torch.from_numpy((rs.randn(*v.shape) * 0.1).astype(np.float32))
        out[k] = v
    return out


def test_library_loads_and_exports_every_declared_symbol():
    L = evrp_b200.lib()
    header = open(os.path.join(ROOT, "include", "evrp_b200.h")).read()
    declared = set(re.findall(r"\b(evrp_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/evrp_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes binding and header disagree"
    assert L.evrp_abi_version() == _lib.ABI_VERSION


def test_phys_constants_match_python_expressions():
    L = evrp_b200.lib()
    c = _lib.Phys()
    L.evrp_phys_default(ctypes.byref(c), 50., 80., 10., 4.)
    p = _lib.make_phys()
    for name, _ in _lib.Phys._fields_:
        assert getattr(c, name) == getattr(p, name), name
    oc = O.EVRPConsts()
    assert np.float32(p.aero) == oc.aero and np.float32(p.kd) == oc.kd and np.float32(p.kr) == oc.kr


def test_constructor_reproduces_reference_initialisation():
    """torch.manual_seed(0) + construction == the reference's AttentionModel weights, bit for bit"""
    W = np.load(os.path.join(GOLDEN, "weights_seed0.npz"))
    torch.manual_seed(0)
    m = evrp_b200.AttentionModel(128, 128, args_for(20), n_encode_layers=3)
    sd = m.state_dict()
    assert set(sd.keys()) == set(W.files) and len(list(m.parameters())) == 50
    assert sum(p.numel() for p in m.parameters()) == 874112
    for k in W.files:
        assert np.array_equal(sd[k].numpy(), W[k]), k


def test_am_constructor_reproduces_reference_initialisation():
    """nets/AM.py variant: torch.manual_seed(0) + construction == the reference's weights, bit for bit"""
    W = np.load(os.path.join(GOLDEN, "weights_am_seed0.npz"))
    torch.manual_seed(0)
    m = evrp_b200.AttentionModelAM(128, 128, args_for(20), n_encode_layers=3)
    sd = m.state_dict()
    assert set(sd.keys()) == set(W.files) and len(list(m.parameters())) == 46
    assert not hasattr(m, "FF_tour") and m.project_step_context.weight.shape == (128, 129)
    for k in W.files:
        assert np.array_equal(sd[k].numpy(), W[k]), k


def test_am_training_math_matches_reference_gradients():
    """nets/AM.py variant: encode_torch + teacher_forced_ll against the reference's REINFORCE gradients (CPU, fp32)"""
    g = np.load(os.path.join(GOLDEN, "am_n20.npz"))
    W = dict(np.load(os.path.join(GOLDEN, "weights_am_seed0.npz")))
    m = evrp_b200.AttentionModelAM(128, 128, args_for(20), n_encode_layers=3)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()})
    m.eval()
    o = O.rollout(W, g["static"], g["dynamic"], g["distances"], g["slope"], forced=g["s_pi"], trace=True)
    assert np.array_equal(o["R"], g["s_R"])
    B, T = g["s_pi"].shape
    mask = o["mask"].astype(bool)
    bits = np.zeros((B, T, 4), np.int64)
    for j in range(26):
        bits[..., j >> 5] |= mask[..., j].astype(np.int64) << (j & 31)
    bits = torch.from_numpy(bits.astype(np.uint32).view(np.int32).reshape(B, T, 4))
    before = np.concatenate([g["dynamic"][:, None], o["dynamic"][:, :-1]], 1)
    veh = torch.from_numpy(np.stack([before[:, :, 0, 0], before[:, :, 2, 0] / np.float32(80),
                                     before[:, :, 3, 0] / np.float32(10)], -1).astype(np.float32))
    emb, kvl, fixed = PM.encode_torch(m, torch.from_numpy(g["static"]), torch.from_numpy(g["dynamic"]))
    ll = PM.teacher_forced_ll(m, emb, kvl, fixed, torch.from_numpy(g["s_pi"]), bits, veh,
                              torch.full((B,), T, dtype=torch.int32))
    assert float((ll.detach() - torch.from_numpy(g["s_ll"])).abs().max()) < 2e-5
    (torch.from_numpy(g["s_adv"]) * ll.sum(1)).mean().backward()
    num = den = 0.0
    for k, p in m.named_parameters():
        ref = torch.from_numpy(g["grad." + k])
        num += float((p.grad - ref).norm() ** 2); den += float(ref.norm() ** 2)
    assert (num / den) ** 0.5 < 1e-4


def test_dataset_reproduces_reference_rng_stream():
    g = np.load(os.path.join(GOLDEN, "greedy_n20.npz"))
    ds = evrp_b200.VehicleRoutingDataset(32, 20, 10., 80., 50., 4, 4, 5, 12345, args_for(20), device="cpu")
    assert np.array_equal(ds.static.numpy(), g["static"])
    assert np.array_equal(ds.dynamic.numpy(), g["dynamic"])
    assert np.array_equal(ds.Elevation.numpy(), g["elevation"])
    assert np.array_equal(ds.distances.numpy(), g["distances"]) and np.array_equal(ds.slope.numpy(), g["slope"])
    item = ds[3]
    assert len(item) == 4 and item[2].shape == (26, 26)
    big = evrp_b200.VehicleRoutingDataset(12801, 21, 10., 80., 50., 4, 4, 5, 1, args_for(21), device="cpu")
    assert len(big[0]) == 3                                   # problems/EVRP.py:99-102 arity rule


def test_no_cpu_fallback():
    m = evrp_b200.AttentionModel(128, 128, args_for(20), n_encode_layers=3)
    m.set_decode_type("greedy")
    g = np.load(os.path.join(GOLDEN, "greedy_n20.npz"))
    batch = tuple(torch.from_numpy(g[k]) for k in ("static", "dynamic", "distances", "slope"))
    with pytest.raises(evrp_b200.EvrpLibraryError):
        m(batch)
    ds = evrp_b200.VehicleRoutingDataset(4, 20, 10., 80., 50., 4, 4, 5, 1, args_for(20), device="cpu")
    with pytest.raises(evrp_b200.EvrpLibraryError):
        ds.update_mask(batch[1], batch[2], batch[3], torch.zeros(32, dtype=torch.int64))


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libevrp_b200.so")
    with pytest.raises(evrp_b200.EvrpLibraryError):
        _lib.lib()


def test_foreign_callbacks_are_rejected():
    m = evrp_b200.AttentionModel(128, 128, args_for(20), n_encode_layers=3,
                                 update_dynamic=lambda *a: None, update_mask=lambda *a: None)
    assert not m._fused_path_ok()


def test_weight_packing_and_folds():
    torch.manual_seed(0)
    m = evrp_b200.AttentionModel(128, 128, args_for(20), n_encode_layers=3)
    m.load_state_dict(bn_perturb(m.state_dict()))
    m.eval()
    pk = m.packed_weights()
    assert pk is m.packed_weights()                           # cached until a parameter changes
    assert pk["l0.Wqkv"].shape == (128, 384) and pk["l0.ff1w"].shape == (128, 512) and pk["node_proj"].shape == (128, 384)
    x = torch.randn(7, 128)
    att = m.embedder.layers[0][0].module
    q_ref = torch.matmul(x, att.W_query).permute(1, 0, 2).reshape(7, 128)
    assert torch.allclose(x @ pk["l0.Wqkv"][:, :128], q_ref, atol=1e-5)
    bn = m.embedder.layers[0][1].normalizer
    assert torch.allclose(bn(x), x * pk["l0.bn1s"] + pk["l0.bn1b"], atol=1e-5)
    heads = torch.randn(7, 128)
    lk_ref = m.project_node_embeddings(x)[:, 256:]
    logits_ref = (m.project_out(heads) * lk_ref).sum(1)
    assert torch.allclose((heads * (x @ pk["node_proj"][:, 256:])).sum(1), logits_ref, atol=1e-4)
    veh = torch.randn(7, 3)
    h_ref = m.FF_tour[0](torch.cat([m.tour(veh), x], 1))
    h_fold = veh @ pk["w1_veh"] + x @ pk["w1_max"] + pk["b1"]
    assert torch.allclose(h_fold, h_ref, atol=1e-4)
    with torch.no_grad():
        m.tour.weight.add_(1.0)
    assert m.packed_weights() is not pk                       # version counter invalidates the cache


def test_ff_stream_layout_is_the_swizzled_stage_order():
    """policy_math.pack_ff_stream (host statement of include/evrp_b200.h ``ff_stream``): stage order, (hi | lo) halves and the
    128-byte swizzle, checked element by element against the definition and against a decoded copy"""
    g = torch.Generator().manual_seed(3)
    w1h, w1l = torch.randn(512, 128, generator=g).half(), torch.randn(512, 128, generator=g).half()
    w2h, w2l = torch.randn(128, 512, generator=g).half(), torch.randn(128, 512, generator=g).half()
    st = PM.pack_ff_stream(w1h, w1l, w2h, w2l)
    assert st.dtype == torch.float16 and st.numel() == 16 * 16384                      # 16 stages x 32 KB
    def at(stage, plane, r, k):                    # k in [0, 64): element of the [128 rows][64 halfs] tile
        c = (k >> 3) ^ (r & 7)                     # 16-byte chunk position after the swizzle
        return float(st[stage * 16384 + plane * 8192 + r * 64 + c * 8 + (k & 7)])
    rs = np.random.RandomState(0)
    for _ in range(400):
        f, k = int(rs.randint(512)), int(rs.randint(128))
        stage = (f >> 7) * 2 + (k >> 6)
        assert at(stage, 0, f & 127, k & 63) == float(w1h[f, k]) and at(stage, 1, f & 127, k & 63) == float(w1l[f, k])
        o, k2 = int(rs.randint(128)), int(rs.randint(512))
        assert at(8 + (k2 >> 6), 0, o, k2 & 63) == float(w2h[o, k2]) and at(8 + (k2 >> 6), 1, o, k2 & 63) == float(w2l[o, k2])


def test_bench_arms_print_the_same_config():
    """bench.py: the repo arm and --impl reference emit the same static `config` object (nothing measured inside it)"""
    import bench
    a = bench.static_config(100, 5, 8192, 1)
    assert a == bench.static_config(100, 5, 8192, 1) and set(a) == {"workload", "per_gpu_batch", "global_batch", "nodes", "stations",
                                                                     "l2", "parallelism"}
    assert bench.static_config(100, 5, 8192, 8)["global_batch"] == 65536


def test_tf32_split_is_exact_to_fp32_rounding():
    w = torch.randn(512, 128) * 3
    hi, lo = PM.split_tf32(w)
    assert int((hi.view(torch.int32) & 0x1FFF).abs().max()) == 0 and int((lo.view(torch.int32) & 0x1FFF).abs().max()) == 0
    assert float(((w - (hi + lo)).abs() / w.abs().clamp_min(1e-30)).max()) <= 2 ** -22


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_training_math_matches_reference_gradients(mode):
    """policy_math.encode_torch + teacher_forced_ll against the reference's REINFORCE gradients (CPU, fp32)"""
    g = np.load(os.path.join(GOLDEN, "sample_n20.npz"))
    torch.manual_seed(0)
    m = evrp_b200.AttentionModel(128, 128, args_for(20), n_encode_layers=3)
    m.load_state_dict(bn_perturb(m.state_dict()))
    W = {k: v.detach().numpy() for k, v in m.state_dict().items()}
    o = O.rollout(W, g["static"], g["dynamic"], g["distances"], g["slope"], forced=g["pi"], trace=True)
    B, T = g["pi"].shape
    S = 26
    mask = o["mask"].astype(bool)
    bits = np.zeros((B, T, 4), np.int64)
    for j in range(S):
        bits[..., j >> 5] |= mask[..., j].astype(np.int64) << (j & 31)
    bits = torch.from_numpy(bits.astype(np.uint32).view(np.int32).reshape(B, T, 4))
    before = np.concatenate([g["dynamic"][:, None], o["dynamic"][:, :-1]], 1)
    veh = torch.from_numpy(np.stack([before[:, :, 0, 0], before[:, :, 2, 0] / np.float32(80),
                                     before[:, :, 3, 0] / np.float32(10)], -1).astype(np.float32))
    m.train(mode == "train")
    emb, kvl, fixed = PM.encode_torch(m, torch.from_numpy(g["static"]), torch.from_numpy(g["dynamic"]))
    ll = PM.teacher_forced_ll(m, emb, kvl, fixed, torch.from_numpy(g["pi"]), bits, veh,
                              torch.full((B,), T, dtype=torch.int32))
    assert float((ll - torch.from_numpy(g[f"ll_{mode}"])).abs().max()) < 2e-5
    (torch.from_numpy(g["adv"]) * ll.sum(1)).mean().backward()
    num = den = 0.0
    for k, p in m.named_parameters():
        ref = torch.from_numpy(g[f"g_{mode}/{k}"])
        num += float((p.grad - ref).norm() ** 2); den += float(ref.norm() ** 2)
    assert (num / den) ** 0.5 < 1e-4


def test_cvrplib_parser(tmp_path):
    txt = "NAME : toy\nTYPE : CVRP\nDIMENSION : 4\nEDGE_WEIGHT_TYPE : EUC_2D\nCAPACITY : 100\nNODE_COORD_SECTION\n" \
          "1 30 40\n2 37 52\n3 49 43\n4 52 64\nDEMAND_SECTION\n1 0\n2 19\n3 30\n4 16\nDEPOT_SECTION\n1\n-1\nEOF\n"
    f = tmp_path / "toy.txt"
    f.write_text(txt)
    coords, dem, cap = evrp_b200.parse_cvrplib(str(f))
    assert coords == [(30, 40), (37, 52), (49, 43), (52, 64)] and dem == [0, 19, 30, 16] and cap == 100
    a = args_for(3); a.CVRP_lib_test = True; a.CVRP_lib_path = str(f)
    ds = evrp_b200.VehicleRoutingDataset(5, 3, 10., 80., 50., 4, 4, 5, 1, a, device="cpu")
    assert len(ds) == 1 and ds.static.shape == (1, 2, 9)
    assert torch.allclose(ds.dynamic[0, 1, 6:], torch.tensor([0.19, 0.30, 0.16]))
    assert torch.equal(ds.static[0, :, 0], torch.tensor([30., 40.]))


@pytest.mark.skipif(not os.path.isdir("/root/reference/EM-EVRP/ExperimentalData/CVRPlib"),
                    reason="reference data only exists in the build container")
def test_cvrplib_parser_matches_reference_reader_on_its_own_files():
    """utils/functions.py:213-231 read_file / parse_file of the (patched-import) reference against
    evrp_b200.parse_cvrplib on the 8 Augerat P-set instances the reference ships"""
    import glob
    from oracle import ref_shim
    ref_shim.load()
    from utils.functions import read_file                     # the reference's own reader
    files = sorted(glob.glob("/root/reference/EM-EVRP/ExperimentalData/CVRPlib/*.txt"))
    assert len(files) == 8
    for f in files:
        coords, dem, cap = evrp_b200.parse_cvrplib(f)
        rc, rd, rcap = read_file(f)
        assert [list(c) for c in coords] == [list(c) for c in rc] and list(dem) == list(rd) and cap == rcap, f


@pytest.mark.skipif(not os.path.isdir("/root/reference/EM-EVRP/ExperimentalData/CVRPlib"),
                    reason="reference data only exists in the build container")
@pytest.mark.parametrize("name,n", [("P-n21-k2.txt", 20), ("P-n51-k10.txt", 50), ("P-n101-k4.txt", 100)])
def test_cvrplib_dataset_branch_matches_reference(name, n):
    """problems/EVRP.py:58-66 (CVRP_lib_test): the instance built from a CVRPLIB file -- depot, the 5 random stations
    (same RNG stream), customers, demands / capacity, distance and slope matrices -- equals the reference's, bit for bit"""
    import types
    from oracle import ref_shim
    _, RefDS = ref_shim.load()
    path = "/root/reference/EM-EVRP/ExperimentalData/CVRPlib/" + name
    a = types.SimpleNamespace(CVRP_lib_test=True, CVRP_lib_path=path, Start_SOC=80., t_limit=10., num_nodes=n, charging_num=5)
    ref = RefDS(1, n, 10., 80., 50., 4, 4, 5, 12345, a)
    ours = evrp_b200.VehicleRoutingDataset(1, n, 10., 80., 50., 4, 4, 5, 12345, a, device="cpu")
    assert np.array_equal(ours.static.numpy(), ref.static.cpu().numpy())
    assert np.array_equal(ours.dynamic.numpy(), ref.dynamic.cpu().numpy())
    assert np.abs(ours.distances.numpy() - ref.distances.cpu().numpy()).max() <= 1e-5       # ATen's vectorised sqrt: 1 ulp
    assert np.abs(ours.slope.numpy() - ref.slope.cpu().numpy()).max() <= 1e-6


@pytest.mark.skipif(not os.path.isdir("/root/reference/EM-EVRP"), reason="the live reference only exists in the build container")
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_training_math_against_the_live_reference_on_fresh_inputs(mode):
    """policy_math (the differentiable half of the training path, CPU statement) against the patched reference run LIVE
    on a seed / graph size no fixture holds: sampled rollout of the reference, its REINFORCE gradients with the actions
    forced, vs encode_torch + teacher_forced_ll on the same weights -- log-probs 2e-5 (6e-5 with batch statistics),
    gradients of all 50 parameters 1e-4 relative, BatchNorm running statistics updated identically"""
    from oracle import make_golden as MG, ref_shim
    N, B = 30, 6
    ds, rargs = ref_shim.make_dataset(B, N, 999)
    rm = ref_shim.make_model(ds, rargs, 0)
    rm.load_state_dict(MG.perturb_bn(rm.state_dict()))
    bat = MG.batch_of(ds, B)
    rm.eval(); rm.set_decode_type("sample")
    torch.manual_seed(5)
    with torch.no_grad():
        pi, _, R = rm(bat)
    adv = (R.sum(1) - R.sum(1).mean()).detach()
    m = evrp_b200.AttentionModel(128, 128, args_for(N), n_encode_layers=3)
    m.load_state_dict(rm.state_dict())
    rm.train(mode == "train"); m.train(mode == "train")
    ll_ref, _, _, g_ref = MG.teacher_forced_grads(rm, bat, pi, adv)
    W = {k: v.detach().numpy() for k, v in m.state_dict().items()}
    static, dynamic, D, G = (x.numpy().astype(np.float32) for x in bat)
    o = O.rollout(W, static, dynamic, D, G, forced=pi.numpy(), trace=True)
    T = pi.shape[1]
    bits = np.zeros((B, T, 4), np.int64)
    for j in range(N + 6):
        bits[..., j >> 5] |= o["mask"][..., j].astype(np.int64) << (j & 31)
    bits = torch.from_numpy(bits.astype(np.uint32).view(np.int32).reshape(B, T, 4))
    before = np.concatenate([dynamic[:, None], o["dynamic"][:, :-1]], 1)
    veh = torch.from_numpy(np.stack([before[:, :, 0, 0], before[:, :, 2, 0] / np.float32(80),
                                     before[:, :, 3, 0] / np.float32(10)], -1).astype(np.float32))
    emb, kvl, fixed = PM.encode_torch(m, torch.from_numpy(static), torch.from_numpy(dynamic))
    ll = PM.teacher_forced_ll(m, emb, kvl, fixed, pi, bits, veh, torch.full((B,), T, dtype=torch.int32))
    assert float((ll.detach() - ll_ref).abs().max()) < (2e-5 if mode == "eval" else 6e-5)
    (adv * ll.sum(1)).mean().backward()
    num = den = 0.0
    for k, p in m.named_parameters():
        num += float((p.grad - g_ref[k]).norm() ** 2); den += float(g_ref[k].norm() ** 2)
    assert (num / den) ** 0.5 < 1e-4
    if mode == "train":
        for k, v in m.state_dict().items():
            if "running" in k:
                assert torch.allclose(v, rm.state_dict()[k], rtol=1e-4, atol=1e-5), k


def test_test_set_pickle_round_trip(tmp_path):
    """the reference's test-set file (trainer.py:361-381 writer, :469-482 reader): nested-list 4-tuples"""
    import pickle
    import evrp_b200
    a = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=10, charging_num=5)
    ds = evrp_b200.VehicleRoutingDataset(6, 10, 10., 80., 50., 4, 4, 5, 77, a, device="cpu")
    path = evrp_b200.save_dataset(ds, str(tmp_path / "test_data" / "10" / "6_seed77"))
    assert path.endswith(".pkl")
    raw = pickle.load(open(path, "rb"))
    assert isinstance(raw, list) and len(raw) == 6 and len(raw[0]) == 4 and isinstance(raw[0][0], list)
    assert np.asarray(raw[0][0]).shape == (2, 16) and np.asarray(raw[0][2]).shape == (16, 16)
    back = evrp_b200.EVRPDataset(path, num_samples=4, offset=1, device="cpu")
    assert len(back) == 4
    for i, item in enumerate(back):
        for got, want in zip(item, ds[i + 1]):
            assert torch.equal(got, torch.as_tensor(want).float().cpu())


def test_eval_dataset_derives_replicas_like_the_reference(tmp_path):
    """trainer.py:421-466: (batch_rep, iter_rep) from width / eval_batch_size / max_calc_batch_size, the greedy and
    beam-search guards, the CSV -- host logic only, with a stub model"""
    import types
    from evrp_b200 import train as T
    calls = []

    class Stub:
        def eval(self): pass
        def set_decode_type(self, t, temp=None): calls.append(("type", t, temp))
        def sample_many(self, batch, batch_rep=1, iter_rep=1):
            calls.append(("sample_many", batch[0].shape[0], batch_rep, iter_rep))
            return torch.full((batch[0].shape[0],), float(batch_rep * iter_rep)), None

    data = [(torch.zeros(2, 5), torch.zeros(4, 5), torch.zeros(5, 5), torch.zeros(5, 5)) for _ in range(6)]
    a = types.SimpleNamespace(decode_strategy="greedy", eval_batch_size=4, max_calc_batch_size=100, width=None,
                              softmax_temperature=2.0)
    mean, dur = T.eval_dataset(data, 0, 2.0, a, Stub(), csv_path=str(tmp_path / "e.csv"))
    assert mean == 1.0 and len(dur) == 2 and calls[0] == ("type", "greedy", None)          # temperature NOT applied (as the reference)
    assert [c for c in calls if c[0] == "sample_many"] == [("sample_many", 4, 1, 1), ("sample_many", 2, 1, 1)]
    assert len(open(tmp_path / "e.csv").read().strip().splitlines()) == 7
    with pytest.raises(AssertionError):
        T.eval_dataset(data, 8, 1.0, a, Stub())                                           # "Do not set width when using greedy"
    del calls[:]
    a.decode_strategy, a.width = "sample", [8, 16]
    assert T.eval_widths(data, a, Stub()) == [8.0, 16.0]
    a.eval_batch_size, a.max_calc_batch_size = 1, 32                                      # width split into passes
    mean, _ = T.eval_dataset(data[:2], 128, 1.0, a, Stub())
    assert mean == 128.0 and calls[-1] == ("sample_many", 1, 32, 4)
    with pytest.raises(AssertionError):
        T.eval_dataset(data[:2], 100, 1.0, a, Stub())                                     # width % max_calc_batch_size != 0
    a.apply_softmax_temperature = True
    del calls[:]
    T.eval_dataset(data[:1], 32, 0.5, a, Stub())
    assert calls[0] == ("type", "sample", 0.5)
    a.decode_strategy = "bs"
    with pytest.raises(NotImplementedError):
        T.eval_dataset(data[:1], 2, 1.0, a, Stub())
