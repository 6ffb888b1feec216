#!/usr/bin/env python
"""bench.py — greedy rollout throughput of EM-EVRP-100 (BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--nodes N]

A step = one pass of the hot path (init-embed + encoder + precompute kernels, then ONE persistent rollout
kernel) over one batch of `--batch` synthetic EM-EVRP-100 instances per GPU (config 4 of BASELINE.json:
batch 8192; with N GPUs every rank owns its own 8192 instances -> weak scaling, no data-path collective).

value      instances/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
e2e        the same through the public API call `model((static, dynamic, distances, slope))` with static/dynamic
           in pinned HOST memory (distances/slope live on the device exactly as the reference's dataset keeps
           them, problems/EVRP.py:87-92) and pi / cost read back to the host, copies inside the timed region
roofline   rollout kernel, HBM-bound.  `achieved` = the bytes the launch NEEDS (counted live from the recorded
           feasibility masks of the same rollout: 1,536 B of K / V / logit-K per FEASIBLE node per step + the distance /
           slope / state rows; masked nodes contribute exactly 0 to the glimpse and the pointer, the kernel never reads
           them) / CUDA-event duration of the kernel launch; `frac` = achieved / MEASURED_PEAKS.json hbm_gbs.  SURVEY
           §8d's all-node model ((1568*S+2048) B per instance-step) is printed beside it as `allnode_*` (it counts the
           masked nodes too and therefore exceeds the peak).  `traffic` = ncu dram bytes of one launch, reported only
           when profiles/rollout_traffic.json was captured with THIS kernel source (hash), batch and graph size.
cpu_baseline  the PATCHED REFERENCE itself (oracle/_ref, torch on the host cores; kind "reference"), or the numpy
           oracle port where the copy is absent (kind "port"); bounded sample, child process without CUDA
reference_gpu  the patched reference as eager PyTorch on the same B200, batch 512
strong     (N > 1) BASELINE configs[3] as stated: ONE global batch of 8192 sharded B/N per rank
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "greedy rollout instances/sec (EVRP-100)"
UNIT = "instances/s"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def synth(B, N, seed):
    import evrp_b200                              # the package's own generator (distribution of problems/EVRP.py:43-92)
    return evrp_b200.synthetic_instances(B, N, seed=seed)


def golden_weights():
    """cpu_baseline / reference legs only: the torch.manual_seed(0) weights as numpy arrays for the oracle"""
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "weights_seed0.npz")))


def seeded_model(evrp_b200, args, ds, dev, seed=0):
    """random-init policy exactly as the reference builds it: torch.manual_seed(seed), then construct (parameter creation
    order and initialisers are the reference's, tests/test_boundary.py pins this against tests/golden/weights_seed0.npz)"""
    import torch
    torch.manual_seed(seed)
    m = evrp_b200.AttentionModel(128, 128, args, n_encode_layers=3, update_dynamic=ds.update_dynamic,
                                 update_mask=ds.update_mask)
    return m.to(dev)


def cpu_port_rate(N, sample, threads=None):
    """numpy oracle (kind 'port') greedy rollout on the host cores: instances/s on a bounded sample."""
    from oracle import evrp_oracle as O
    from threadpoolctl import threadpool_info
    W = golden_weights()
    s, d, e = O.generate_instances(sample, N, seed=4242)
    D, G = O.pairwise_matrices(s, e)
    O.rollout(W, s[:8], d[:8], D[:8], G[:8])      # warm-up
    t = time.perf_counter()
    o = O.rollout(W, s, d, D, G)
    dt = time.perf_counter() - t
    cores = max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
    return sample / dt, cores, dt, int(o["pi"].shape[1])


ROLLOUT_SOURCES = ("evrp_decode_tc.cu", "evrp_tc.cuh", "evrp_common.cuh")


def rollout_source_sha256():
    """hash of the rollout kernel's SOURCE files (the ncu traffic figure under profiles/ is tied to this: it stays valid when
    other kernels of the library change and goes stale when this kernel does)"""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "drl-energy-optimal-routing-for-electric-vehicles_b200", "csrc")
    for name in ROLLOUT_SOURCES:
        with open(os.path.join(d, name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def static_config(N, F, B, world):
    """the `config` object of both arms (repo and --impl reference): the workload as BASELINE.json states it, nothing measured"""
    S = N + F + 1
    return {"workload": f"EM-EVRP-{N} greedy rollout, random-init attention policy (BASELINE configs[3])",
            "per_gpu_batch": B, "global_batch": B * world, "nodes": N, "stations": F,
            "l2": "inputs_exceed_l2 (distances+slope+kvl = %.0f MB per step)" % ((2 * S * S + 384 * S) * 4 * B / 1e6),
            "parallelism": f"dp{world} (instances sharded, no data-path collective)"}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def reference_child(kind, N, B, steps, cuda, timeout=1500):
    """the patched reference (oracle/ref_runner.py) in a child process: all host threads, CUDA hidden unless `cuda`"""
    env = dict(os.environ)
    n = host_threads()
    env["OMP_NUM_THREADS"] = env["MKL_NUM_THREADS"] = str(n)          # torchrun exports OMP_NUM_THREADS=1
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    if not cuda:
        env["CUDA_VISIBLE_DEVICES"] = ""
    out = subprocess.run([sys.executable, "-m", "oracle.ref_runner", kind, str(N), str(B), str(steps)], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=timeout)
    if out.returncode != 0:
        raise RuntimeError(out.stderr.strip().splitlines()[-1] if out.stderr.strip() else "reference child failed")
    return json.loads(out.stdout.strip().splitlines()[-1])


def needed_bytes(o, S):
    """bytes one rollout launch needs (tools/needed_bytes.py): per live step 1,536 B per FEASIBLE node (glimpse K, glimpse V,
    logit K records) + D[c,:], G[c,:], D[0,:], G[0,:], rule-7 rows (6 S floats) + emb[now], emb[c], fixed context, demand
    row r/w, running-max row r/w (9 x 512 B).  `o` = a rollout with save_trace=True."""
    import torch
    from evrp_b200 import policy_math as PM
    T = o["pi"].shape[1]
    live = torch.arange(T, device=o["pi"].device)[None] < o["steps"][:, None]
    feas = int((PM.bits_to_mask(o["mask_bits"], S).sum(-1) * live).sum())
    steps = int(o["steps"].sum())
    return feas * 1536 + steps * (6 * S * 4 + 9 * 512), feas / max(steps, 1), steps


def multi_rank_gradient_check(evrp_b200, model, inst, rank, world, dev):
    """(N > 1) the data-parallel training path against a single-rank recompute: a 256-instance batch is sampled once
    (rank 0, actions broadcast), every rank takes the REINFORCE gradient of ITS shard teacher-forced on those actions
    (global-batch BatchNorm statistics + flat gradient all-reduce), rank 0 repeats the step on the whole batch alone.
    Also: the BatchNorm running statistics and a RolloutBaseline.epoch_callback (CPU statistics through NCCL) agree
    on every rank."""
    import torch
    import torch.distributed as dist
    from evrp_b200 import policy_math as PM, train as Tr
    s, d, e = inst
    st, dy = torch.from_numpy(s).to(dev), torch.from_numpy(d).to(dev)
    D, G = PM.pairwise_matrices(st, torch.from_numpy(e).to(dev))
    n = st.shape[0]
    state = {k: v.clone() for k, v in model.state_dict().items()}
    model.train(); model.set_decode_type("sample")
    with torch.no_grad(), PM.local_batch_norm():
        pi, _, _ = model((st, dy, D, G))
    pi = pi.contiguous()
    model.load_state_dict(state)
    tw = torch.tensor([pi.shape[1]], device=dev); dist.broadcast(tw, 0)
    if rank != 0:
        pi = torch.zeros(n, int(tw.item()), dtype=torch.int64, device=dev)
    dist.broadcast(pi, 0)
    params = [p for p in model.parameters()]

    def grads(lo, hi, local):
        model.load_state_dict(state)
        model.zero_grad()
        if local:
            with PM.local_batch_norm():
                _, ll, R = model((st[lo:hi], dy[lo:hi], D[lo:hi], G[lo:hi]), forced_actions=pi[lo:hi])
        else:
            _, ll, R = model((st[lo:hi], dy[lo:hi], D[lo:hi], G[lo:hi]), forced_actions=pi[lo:hi])
        ((R.sum(1) - 4000.0).detach() * ll.sum(1)).sum().div(n).backward()
        if not local:
            Tr.allreduce_gradients(params)
        return torch.cat([p.grad.reshape(-1) for p in params]).clone(), \
            torch.cat([b.reshape(-1).float() for k, b in model.state_dict().items() if "running" in k]).clone()
    lo, hi = rank * n // world, (rank + 1) * n // world
    g_multi, bn_multi = grads(lo, hi, False)
    out = None
    bn_all = [torch.zeros_like(bn_multi) for _ in range(world)]
    dist.all_gather(bn_all, bn_multi)
    if rank == 0:
        g_single, bn_single = grads(0, n, True)
        g_again, _ = grads(0, n, True)                      # run-to-run floor of the single-rank step (fp32 atomics)
        worst, off, gmax = ("", 0.0), 0, float(g_single.abs().max())
        for name, p_ in model.named_parameters():
            k = p_.numel()
            e_ = float((g_multi[off:off + k] - g_single[off:off + k]).abs().max()) / gmax
            if e_ > worst[1]:
                worst = (name, e_)
            off += k
        out = {"batch": n, "grad_rel_err": float((g_multi - g_single).abs().max() / g_single.abs().max()),
               "single_rank_repeat_rel_err": float((g_again - g_single).abs().max() / g_single.abs().max()),
               "worst_parameter": {"name": worst[0], "max_abs_err_over_global_max": worst[1]},
               "grad_l2_rel_err": float((g_multi - g_single).norm() / g_single.norm()),
               "bn_running_stats_max_diff_across_ranks": float(max((b - bn_all[0]).abs().max() for b in bn_all)),
               "bn_running_stats_vs_single_rank": float((bn_multi - bn_single).abs().max())}
    model.load_state_dict(state)
    # the rollout baseline's statistics pass through numpy (CPU tensors) before the NCCL reductions
    model.eval()
    class _DS(torch.utils.data.Dataset):
        def __len__(self): return hi - lo
        def __getitem__(self, i): return (st[lo + i], dy[lo + i], D[lo + i], G[lo + i])
    import types as _t
    bl = Tr.RolloutBaseline(model, _DS(), _t.SimpleNamespace(batch_size=64, bl_alpha=0.05))
    bl.epoch_callback(model, 1)
    m_all = torch.tensor([bl.mean], dtype=torch.float64, device=dev)
    ms = [torch.zeros_like(m_all) for _ in range(world)]
    dist.all_gather(ms, m_all)
    if rank == 0:
        out["rollout_baseline_mean_agrees"] = bool(all(float(x) == float(ms[0]) for x in ms))
        out["tolerance"] = "1e-3 of the largest gradient entry (north-star gradient bar); the two paths differ by GEMM tile / reduction order, the single-rank step repeats to single_rank_repeat_rel_err"
        out["ok"] = bool(out["grad_rel_err"] < 1e-3 and out["bn_running_stats_max_diff_across_ranks"] == 0.0
                         and out["rollout_baseline_mean_agrees"])
    model.train(); model.set_decode_type("sample")
    return out


class ClockSampler:
    """ONE nvidia-smi process (started by rank 0) samples every GPU of the job: 8 ranks each polling NVML at 50 ms
    measurably stall each other's launches on an 8-GPU box."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,index"

    def __init__(self, indices, period_ms=50):
        self.rows, self.p = [], None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--id=" + ",".join(str(i) for i in indices), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", str(period_ms)],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        rows = [r for ts, r in self.rows if t0 <= ts <= t1 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
               "samples": len(rows), "power_w_max": max(float(r[2]) for r in rows)}
        gpus = sorted({r[7] for r in rows if len(r) >= 8})
        if len(gpus) > 1:                                   # per-GPU median SM clock / max power (multi-GPU runs)
            out["sm_mhz_min_gpu"] = min(sorted(float(r[0]) for r in rows if r[7] == g)[len([r for r in rows if r[7] == g]) // 2] for g in gpus)
            out["per_gpu"] = {g: {"sm_mhz": sorted(float(r[0]) for r in rows if r[7] == g)[len([r for r in rows if r[7] == g]) // 2],
                                  "power_w_max": max(float(r[2]) for r in rows if r[7] == g)} for g in gpus}
        return out


def run_reference(a):
    """The reference's own CPU implementation of the path on the host cores: the patched reference (oracle/_ref, torch CPU,
    all host threads, child process without CUDA) when the copy is present, else the numpy oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = a.ref_sample
    from oracle import ref_shim
    kind, note = "port", "numpy oracle port of the reference path (oracle/_ref absent)"
    if ref_shim.available():
        try:
            r = reference_child("greedy", a.nodes, sample, a.steps, cuda=False)
            v, T, cores, t_all, kind = r["value"], r["decode_steps"], r["threads"], r["seconds"], "reference"
            note = "patched reference (oracle/_ref: EM-EVRP nets/DRLModel.py + problems/EVRP.py, torch CPU), actor.forward(batch)"
        except Exception as ex:
            note = f"numpy oracle port (reference child failed: {type(ex).__name__}: {ex})"
    if kind == "port":
        for _ in range(a.warmup if a.warmup < 1 else 1):
            cpu_port_rate(a.nodes, 16)
        rates, T, cores, t_all = [], 0, 1, 0.0
        for _ in range(a.steps):
            r, cores, dt, T = cpu_port_rate(a.nodes, sample)
            rates.append(r); t_all += dt
        v = float(np.mean(rates))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": min(a.warmup, 1), "ms_per_step": 1e3 * t_all / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": static_config(a.nodes, 5, a.batch, a.gpus),      # the repo arm's config, key for key
            "workload_stats": {"decode_steps": T, "per_step_instances": sample},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{sample} instances per step x {a.steps} steps; {note}"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


def _emit(line):
    """the ONE JSON line, written to the real stdout (everything else - e.g. NCCL's version banner - goes to stderr)"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)


def main():
    os.dup2(2, 1)                                     # libraries that print to fd 1 must not pollute the JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=8192, help="instances per GPU per step (BASELINE configs[3])")
    ap.add_argument("--nodes", type=int, default=100)
    ap.add_argument("--cpu-sample", type=int, default=256)
    ap.add_argument("--ref-sample", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--train-steps", type=int, default=3, help="timed REINFORCE steps (0 = skip)")
    ap.add_argument("--train-batch", type=int, default=1024)
    a = ap.parse_args()
    if a.impl == "reference":
        return run_reference(a)

    import torch
    import torch.distributed as dist
    import evrp_b200
    evrp_b200.lib()                                   # fail loudly without the CUDA library
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    N, B, F = a.nodes, a.batch, 5
    S = N + F + 1
    args = types.SimpleNamespace(CVRP_lib_test=False, Start_SOC=80., t_limit=10., num_nodes=N, charging_num=F)
    ds = evrp_b200.VehicleRoutingDataset(2, N, 10., 80., 50., 4, 4, F, 1, args)
    model = seeded_model(evrp_b200, args, ds, dev).eval()                                  # torch.manual_seed(0) init
    model.set_decode_type("greedy")
    static, dynamic, elev = synth(B, N, seed=1000 + rank)
    st_h = torch.from_numpy(static).pin_memory(); dy_h = torch.from_numpy(dynamic).pin_memory()
    st, dy = st_h.to(dev), dy_h.to(dev)
    D, G = evrp_b200.policy_math.pairwise_matrices(st, torch.from_numpy(elev).to(dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    enc_events = []

    def step_resident():
        with torch.no_grad():
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
            emb, kvl, fixed = model.encode(st, dy)
            ev[1].record()
            enc_events.append(ev)
            return model.rollout(emb, kvl, fixed, dy, D, G)

    host_out = {}

    def step_e2e():
        """public API call with host inputs (pinned) and the results (tour, cost) read back into pinned host memory"""
        with torch.no_grad():
            pi, ll, R = model((st_h, dy_h, D, G))
            cost = R.sum(1)
            if "pi" not in host_out or host_out["pi"].shape != pi.shape:
                host_out["pi"] = torch.empty(pi.shape, dtype=pi.dtype).pin_memory()
                host_out["cost"] = torch.empty(cost.shape, dtype=cost.dtype).pin_memory()
            host_out["pi"].copy_(pi, non_blocking=True)
            host_out["cost"].copy_(cost, non_blocking=True)
            torch.cuda.synchronize()
            return host_out["cost"], host_out["pi"]

    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    phys_ids = [int(x) for x in vis.split(",")][:world] if vis and all(x.strip().isdigit() for x in vis.split(",")) \
        else list(range(world))
    # period: the profiling recipe's 200 ms -- polling NVML every 50 ms measurably slows the GPUs it queries (the device-
    # timed loop ran 1.6 ms / step slower than the e2e loop that follows it, 3.5 ms with 4 GPUs polled)
    period = int(os.environ.get("EVRP_BENCH_SAMPLER_MS", "200"))
    sampler = ClockSampler(phys_ids, period) if rank == 0 and period > 0 else None        # started before the warm-up
    # warm-up steps are EXACTLY the timed steps (same event records, same step-count accumulation): the first launch of
    # any kernel -- including torch's int32 sum / int64 add behind `inst_steps_dev +=` -- loads its module lazily
    # and stalls the stream for milliseconds (tools/loop_gap.py)
    model.profile_events = []
    inst_steps_dev = torch.zeros((), dtype=torch.int64, device=dev)
    for _ in range(max(a.warmup, 3)):
        o = step_resident()
        inst_steps_dev += o["steps"].sum()
    # ---------------- device-resident timing ----------------
    model.profile_events = []
    del enc_events[:]
    barrier()
    t0 = time.perf_counter()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    inst_steps_dev = torch.zeros((), dtype=torch.int64, device=dev)
    for _ in range(a.steps):
        o = step_resident()
        inst_steps_dev += o["steps"].sum()
    e_end.record()
    barrier()
    inst_steps = int(inst_steps_dev)
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1) if sampler is not None else {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampler off"]}
    ms_total = e_start.elapsed_time(e_end)
    roll_ms = [s.elapsed_time(e) for s, e in model.profile_events]
    enc_ms = [s.elapsed_time(e) for s, e in enc_events]
    model.profile_events = None
    T = int(o["pi"].shape[1])
    # ---------------- end-to-end through the public API, host buffers ----------------
    for _ in range(2):
        step_e2e()
    barrier()
    te0 = time.perf_counter()
    for _ in range(a.steps):
        cost, pi_host = step_e2e()
    barrier()
    e2e_s = time.perf_counter() - te0
    h2d = st_h.numel() * 4 + dy_h.numel() * 4
    d2h = cost.numel() * 4 + pi_host.numel() * 8
    # ---------------- bytes the rollout launch needs: one untimed traced rollout of the same (deterministic) batch ------
    with torch.no_grad():
        emb, kvl, fixed = model.encode(st, dy)
        o_tr = model.rollout(emb, kvl, fixed, dy, D, G, save_trace=True)
    need_b, mean_feasible, steps_tr = needed_bytes(o_tr, S)
    assert steps_tr * a.steps == inst_steps, "traced rollout differs from the timed ones"
    del o_tr, emb, kvl, fixed
    # ---------------- strong scaling (BASELINE configs[3] as stated): ONE global batch of `B` sharded B/N per rank ------
    strong = None
    if world > 1:
        gs, gd, ge = synth(B, N, seed=1000)                          # the same global batch on every rank
        lo, hi = rank * B // world, (rank + 1) * B // world
        sst, sdy = torch.from_numpy(gs[lo:hi]).to(dev), torch.from_numpy(gd[lo:hi]).to(dev)
        sD, sG = evrp_b200.policy_math.pairwise_matrices(sst, torch.from_numpy(ge[lo:hi]).to(dev))

        def step_strong():
            with torch.no_grad():
                e_, k_, f_ = model.encode(sst, sdy)
                return model.rollout(e_, k_, f_, sdy, sD, sG)
        for _ in range(3):
            step_strong()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(a.steps):
            step_strong()
        s1.record()
        barrier()
        sms = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        dist.all_reduce(sms, op=dist.ReduceOp.MAX)
        strong = {"global_batch": B, "per_gpu_batch": hi - lo, "ms_per_step": sms.item() / a.steps,
                  "value": B * a.steps / (sms.item() * 1e-3), "unit": UNIT, "scaling": "strong",
                  "note": "the same instances/s metric on ONE global batch sharded over the ranks (device-timed, max over ranks)"}
        del sst, sdy, sD, sG
    tm = torch.tensor([ms_total, e2e_s * 1e3, float(np.mean(roll_ms)), float(inst_steps)], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        mine = torch.tensor([ms_total / a.steps, float(np.mean(enc_ms)), float(np.mean(roll_ms)), e2e_s * 1e3 / a.steps],
                            dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"columns": ["ms_per_step", "encoder_ms", "rollout_kernel_ms", "e2e_ms_per_step"],
                    "rows": [[round(float(v), 3) for v in t.tolist()] for t in allr]}
        mx = tm.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tm.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, e2e_ms, roll_avg_ms = mx[0].item(), mx[1].item(), mx[2].item()
        inst_steps_all = sm[3].item()
    else:
        ms_total, e2e_ms, roll_avg_ms, inst_steps_all = tm[0].item(), tm[1].item(), tm[2].item(), tm[3].item()
    # ---------------- REINFORCE training step (BASELINE configs[4] shape: EM-EVRP-100, batch 1024 per GPU) ----------
    train = None
    if a.train_steps > 0:
        from evrp_b200 import train as Tr
        tb = a.train_batch
        tm_model = seeded_model(evrp_b200, args, ds, dev).train()
        tm_model.set_decode_type("sample")
        opt = torch.optim.Adam(tm_model.parameters(), lr=1e-4)
        bl = Tr.ExponentialBaseline(0.8)
        tbatch = (st[:tb].contiguous(), dy[:tb].contiguous(), D[:tb].contiguous(), G[:tb].contiguous())
        for _ in range(2):
            Tr.reinforce_step(tm_model, bl, opt, tbatch, 2.0)
        barrier()
        tt0 = time.perf_counter()
        for _ in range(a.train_steps):
            out = Tr.reinforce_step(tm_model, bl, opt, tbatch, 2.0)
        barrier()
        tsec = torch.tensor([time.perf_counter() - tt0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tsec, op=dist.ReduceOp.MAX)
        # the rollout kernel inside the step (sampling, trace recorded): needed bytes / its CUDA-event time
        tm_model.profile_events = []
        with torch.no_grad():
            te, tk, tf = tm_model.encode(tbatch[0], tbatch[1])
            to = tm_model.rollout(te, tk, tf, tbatch[1], tbatch[2], tbatch[3], save_trace=True)
        torch.cuda.synchronize()
        t_roll_ms = tm_model.profile_events[-1][0].elapsed_time(tm_model.profile_events[-1][1])
        tm_model.profile_events = None
        t_need, _, _ = needed_bytes(to, S)
        del te, tk, tf, to
        grad_check = None
        if world > 1:
            grad_check = multi_rank_gradient_check(evrp_b200, tm_model, synth(256, N, seed=77), rank, world, dev)
        train = {"metric": "REINFORCE train step/sec (EM-EVRP-100)", "value": a.train_steps / tsec.item(),
                 "grad_check": grad_check,
                 "roofline": {"bound": "hbm", "kernel": "evrp::dtc::rollout_tc_kernel (sampling, trace recorded)",
                              "achieved": t_need / (t_roll_ms * 1e-3) / 1e9, "unit": "GB/s", "kernel_ms": t_roll_ms,
                              "kernel_share_of_step": t_roll_ms / (1e3 * tsec.item() / a.train_steps),
                              "needed_bytes_per_launch": t_need},
                 "unit": "step/s", "per_gpu_batch": tb, "global_batch": tb * world, "decode_steps": int(out["T"]),
                 "includes": "train-mode encoder (projections / FF: tcgen05 3xTF32 GEMM forward + input gradient, tcgen05 split-K "
                             "weight-gradient kernel evrp_wgrad.cu; attention core evrp_attn_train.cu; two-phase BatchNorm "
                             "kernels evrp_train.cu) + sampled rollout kernel + step-query assembly kernels + decoder "
                             "log-prob forward / backward kernels (evrp_logp.cu) + fused loss kernel + flat NCCL gradient "
                             "all-reduce + clip + Adam"}
    if rank == 0:
        pk, pk_src = peaks()
        kern_s = roll_avg_ms * 1e-3
        achieved = need_b / kern_s / 1e9                               # needed bytes of this rank's launch
        allnode = (1568 * S + 2048) * (inst_steps / a.steps)           # SURVEY §8d all-node model, same launch
        traffic, traffic_src = None, "no ncu capture of this build / batch / graph size under profiles/rollout_traffic.json"
        try:
            with open(os.path.join(ROOT, "profiles", "rollout_traffic.json")) as f:
                tj = json.load(f)
            if tj.get("source_sha256") == rollout_source_sha256() and tj.get("batch") == B and tj.get("nodes") == N:
                traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
        except Exception:
            pass
        line = {"metric": METRIC, "value": world * B * a.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_total / a.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": static_config(N, F, B, world),
                "workload_stats": {"decode_steps": T, "mean_steps_per_instance": inst_steps / (a.steps * B),
                                   "mean_feasible_nodes_per_step": mean_feasible},
                "e2e": {"value": world * B * a.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h,
                        "note": "static / dynamic from pinned host memory, tours + costs back to pinned host memory; distances / "
                                "slope stay device-resident exactly as the reference's dataset keeps them (problems/EVRP.py:87-92)"},
                "gpu_launches": a.steps * 3,                        # fused encoder kernel, fixed-context kernel, rollout kernel
                "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": "evrp::dtc::rollout_tc_kernel", "achieved": achieved,
                             "peak": pk["hbm_gbs"], "peak_source": pk_src, "unit": "GB/s",
                             "frac": achieved / pk["hbm_gbs"], "bytes": "needed (feasible nodes only), counted live",
                             "needed_bytes_per_launch": need_b, "traffic": traffic, "traffic_source": traffic_src,
                             "dram_achieved": (traffic / kern_s / 1e9) if traffic else None,
                             "kernel_ms": roll_avg_ms, "kernel_share_of_step": roll_avg_ms / (ms_total / a.steps),
                             "allnode_bytes_per_launch": allnode, "allnode_over_peak": allnode / kern_s / 1e9 / pk["hbm_gbs"],
                             "allnode_note": "SURVEY §8d's streaming model counts all S nodes per step; the kernel reads the "
                                             "feasible ones only, so this ratio is NOT a fraction of peak"}}
        f_enc = (1278720 * S + 1536 * S * S + 32768) * B               # SURVEY §8d algorithmic encoder FLOPs per launch
        enc_s = float(np.mean(enc_ms)) * 1e-3
        f16_peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])  # kind::f16 = the bf16 rate; timed inside a long step
        line["roofline_encoder"] = {"bound": "tensor", "kernels": "evrp::fenc::encoder_fused_kernel (init embedding, 3 layers, node "
                                    "projection; one launch) + fixed_context_kernel", "achieved": f_enc / enc_s / 1e12,
                                    "executed": 3 * f_enc / enc_s / 1e12, "peak": f16_peak,
                                    "peak_source": pk_src + (" bf16 sustained" if "bf16_tflops_sustained" in pk else " bf16"),
                                    "unit": "TFLOP/s", "frac": f_enc / enc_s / 1e12 / f16_peak,
                                    "frac_executed": 3 * f_enc / enc_s / 1e12 / f16_peak, "ms": enc_s * 1e3,
                                    "note": "every product is a 3-MMA fp16-pair split (fp32-faithful): executed = 3 x algorithmic "
                                            "(padding of the 128-row tiles not counted); tensor-pipe activity by ncu: profiles/"}
        line["breakdown"] = {"encoder_ms": float(np.mean(enc_ms)), "rollout_kernel_ms": float(np.mean(roll_ms)),
                             "note": "rank 0, CUDA events inside the timed region"}
        if per_rank is not None:
            line["per_rank"] = per_rank
        if strong is not None:
            strong["efficiency_vs_weak"] = strong["value"] / line["value"]
            line["strong"] = strong
        if train is not None:
            train["roofline"]["peak"] = pk["hbm_gbs"]
            train["roofline"]["frac"] = train["roofline"]["achieved"] / pk["hbm_gbs"]
            line["train"] = train
        if world == 1 and not a.no_cpu_baseline:
            # clean child processes (no CUDA context, no torch thread pools competing with the BLAS threads)
            try:
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2",
                                      "--warmup", "1", "--nodes", str(N), "--ref-sample", str(a.cpu_sample)],
                                     capture_output=True, text=True, timeout=1500).stdout.strip().splitlines()[-1]
                ref = json.loads(out)
                line["cpu_baseline"] = {"value": ref["value"], "unit": UNIT, "cores": ref["cpu_baseline"]["cores"],
                                        "kind": ref["cpu_baseline"]["kind"],
                                        "sample": f"2 x {a.cpu_sample} EM-EVRP-{N} instances, greedy, "
                                        f"{2 * ref['ms_per_step'] / 1e3:.1f} s (child process); " + ref["cpu_baseline"]["sample"]}
            except Exception as ex:                                   # fall back to timing the numpy port in this process
                r, cores, dt, Tc = cpu_port_rate(N, a.cpu_sample)
                line["cpu_baseline"] = {"value": r, "unit": UNIT, "cores": cores, "kind": "port",
                                        "sample": f"{a.cpu_sample} EM-EVRP-{N} instances, greedy, numpy oracle port, "
                                                  f"{dt:.1f} s (in-process: {type(ex).__name__})"}
            # config 1 of BASELINE.json, the reference's own CPU-runnable case: EM-EVRP-20, batch 256, greedy
            from oracle import ref_shim
            if ref_shim.available():
                try:
                    r = reference_child("greedy", 20, 256, 2, cuda=False)
                    line["cpu_config1"] = {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "reference",
                                           "sample": "EM-EVRP-20, batch 256, greedy, 2 forwards (BASELINE configs[0])"}
                except Exception as ex:
                    line["cpu_config1"] = {"unavailable": f"{type(ex).__name__}: {ex}"}
                # the reference as eager PyTorch on this B200: GPU-vs-GPU context for the headline
                try:
                    r = reference_child("greedy", N, 512, 1, cuda=True)
                    line["reference_gpu"] = {"value": r["value"], "unit": UNIT, "batch": 512, "seconds": r["seconds"],
                                             "decode_steps": r["decode_steps"], "device": r["device"],
                                             "what": "patched reference, eager PyTorch CUDA, actor.forward(batch), time.time() "
                                                     "around the call as trainer.py:430-453"}
                except Exception as ex:
                    line["reference_gpu"] = {"unavailable": f"{type(ex).__name__}: {ex}"}
                if train is not None:
                    try:
                        r = reference_child("train", N, 128, 1, cuda=False)
                        train["cpu_baseline"] = {"value": r["value"], "unit": "step/s", "batch": 128, "cores": r["threads"],
                                                 "kind": "reference", "equiv_step_per_s_at_batch_%d" % a.train_batch:
                                                 r["value"] * 128 / a.train_batch,
                                                 "sample": "patched reference REINFORCE step (trainer.py:112-127), EM-EVRP-%d, batch "
                                                           "128, 1 step after 1 warm-up, torch CPU" % N}
                        r = reference_child("train", N, a.train_batch, 2, cuda=True)
                        train["reference_gpu"] = {"value": r["value"], "unit": "step/s", "batch": a.train_batch,
                                                  "what": "patched reference REINFORCE step, eager PyTorch CUDA"}
                    except Exception as ex:
                        train["cpu_baseline"] = train.get("cpu_baseline") or {"unavailable": f"{type(ex).__name__}: {ex}"}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
