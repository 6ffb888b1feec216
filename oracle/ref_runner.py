"""Timing harness for the PATCHED REFERENCE itself (test / bench infrastructure; never on the product path).

Runs /root/reference/EM-EVRP through its own public API -- ``VehicleRoutingDataset`` + ``AttentionModel`` built
exactly as ``trainer.py:238-272`` does, ``actor.forward(batch)`` as ``trainer.py:40`` (greedy validation) and the
REINFORCE step of ``trainer.py:100-127`` -- from the patched scratch copy ``oracle/_ref`` (oracle/ref_shim.py: P1 + P2,
stubbed matplotlib / xlwt).  The copy is a built artefact that travels to the GPU box, so both of bench.py's
reference legs can time the REAL reference there: on the host cores (``cpu_baseline.kind = "reference"``) and, for an
honest GPU-vs-GPU figure, as eager PyTorch on the B200 (``reference_gpu``).

The reference chooses its device at import time (``cuda if available``, nets/DRLModel.py:10, problems/EVRP.py:6): a CPU
run therefore needs CUDA hidden from the process (bench.py starts a child with CUDA_VISIBLE_DEVICES="").

CLI (prints one JSON object):
  python -m oracle.ref_runner greedy N B steps     greedy eval-mode rollout, instances/s
  python -m oracle.ref_runner train  N B steps     sampled rollout + loss + backward + clip + Adam, steps/s
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _sync(torch):
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def _setup(N, B, seed=12345):
    import torch
    from torch.utils.data import DataLoader
    from oracle import ref_shim
    ds, args = ref_shim.make_dataset(B, N, seed)
    m = ref_shim.make_model(ds, args, 0)
    import nets.DRLModel as RM
    m = m.to(RM.device)
    batch = next(iter(DataLoader(ds, B, False, num_workers=0)))
    return torch, m, batch, str(RM.device)


def greedy_rate(N, B, steps, warmup=1):
    """instances/s of ``actor.forward(batch)`` in eval mode, greedy decoding, wall clock as trainer.py:430-453"""
    torch, m, batch, dev = _setup(N, B)
    m.eval(); m.set_decode_type("greedy")
    with torch.no_grad():
        for _ in range(warmup):
            m(batch)
        _sync(torch)
        t = time.time()
        for _ in range(steps):
            pi, ll, R = m(batch)
        _sync(torch)
        dt = time.time() - t
    return {"value": B * steps / dt, "unit": "instances/s", "seconds": dt, "steps": steps, "batch": B, "nodes": N,
            "decode_steps": int(pi.shape[1]), "device": dev, "threads": torch.get_num_threads(),
            "mean_cost": float(R.sum(1).mean())}


def train_rate(N, B, steps, warmup=1):
    """steps/s of the reference's REINFORCE step (trainer.py:112-127) with the exponential baseline's arithmetic inlined"""
    torch, m, batch, dev = _setup(N, B, seed=12346)
    import torch.optim as optim
    m.train(); m.set_decode_type("sample")
    opt = optim.Adam(m.parameters(), lr=1e-4)
    bl = None

    def step():
        nonlocal bl
        pi, ll, R = m(batch)
        reward = R.sum(1)
        bl = reward.mean().detach() if bl is None else 0.8 * bl + 0.2 * reward.mean().detach()
        loss = ((reward - bl).detach() * ll.sum(1)).mean()
        opt.zero_grad()
        loss.backward()
        for g in opt.param_groups:
            torch.nn.utils.clip_grad_norm_(g["params"], 2.0, norm_type=2)
        opt.step()
        return pi

    for _ in range(warmup):
        step()
    _sync(torch)
    t = time.time()
    for _ in range(steps):
        pi = step()
    _sync(torch)
    dt = time.time() - t
    return {"value": steps / dt, "unit": "step/s", "seconds": dt, "steps": steps, "batch": B, "nodes": N,
            "decode_steps": int(pi.shape[1]), "device": dev, "threads": torch.get_num_threads()}


if __name__ == "__main__":
    kind, N, B, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    out = greedy_rate(N, B, steps) if kind == "greedy" else train_rate(N, B, steps)
    print(json.dumps(out))
