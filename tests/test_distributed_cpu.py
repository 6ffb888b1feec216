"""world_size-2 gloo tests of the data-parallel plumbing (evrp_b200/train.py) — no GPU needed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from evrp_b200 import train as T
    out = {}
    n = 1001
    lo, hi = T.shard_bounds(n)
    out["bounds"] = (lo, hi)
    rs = np.random.RandomState(0)
    cand, base = rs.randn(n) * 3 + 10, rs.randn(n) * 3 + 10.4
    # global mean of a sharded vector
    out["mean"] = float(T.global_mean(torch.from_numpy(cand[lo:hi]).float()))
    # paired one-sided t-test from sharded sufficient statistics
    out["ttest"] = T.paired_ttest_one_sided(cand[lo:hi], base[lo:hi])
    # flat-bucket gradient all-reduce
    torch.manual_seed(0)
    lin = torch.nn.Linear(8, 4)
    x = torch.randn(10, 8)
    xs = x[5 * rank:5 * rank + 5]
    (lin(xs).pow(2).sum() / 10).backward()                     # this rank's share of the GLOBAL mean loss
    T.allreduce_gradients(list(lin.parameters()))
    out["grad"] = lin.weight.grad.clone().numpy()
    bl = T.ExponentialBaseline(0.8)
    v, _ = bl.eval(None, torch.from_numpy(cand[lo:hi]).float())
    out["ema"] = float(v)
    # train-mode BatchNorm over the global batch (nets/GraphEncoder.py:144-145) with rows sharded over ranks
    from evrp_b200 import policy_math as PM
    torch.manual_seed(1)
    bn = torch.nn.BatchNorm1d(6)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.5, 0.5)
    xb = torch.randn(14, 6) * 2 + 1
    gy = torch.randn(14, 6)
    a, b = (0, 9) if rank == 0 else (9, 14)                    # ragged shards
    xs = xb[a:b].clone().requires_grad_(True)
    y = PM._batch_norm(bn, xs, True)
    (y * gy[a:b]).sum().backward()
    T.allreduce_gradients([bn.weight, bn.bias])
    out["bn"] = dict(y=y.detach().numpy(), gx=xs.grad.numpy(), gw=bn.weight.grad.numpy(), gb=bn.bias.grad.numpy(),
                     rm=bn.running_mean.numpy().copy(), rv=bn.running_var.numpy().copy(),
                     nbt=int(bn.num_batches_tracked))
    q.put((rank, out))
    dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0]["bounds"] == (0, 501) and res[1]["bounds"] == (501, 1001)
    rs = np.random.RandomState(0)
    cand, base = rs.randn(1001) * 3 + 10, rs.randn(1001) * 3 + 10.4
    for r in (0, 1):
        assert abs(res[r]["mean"] - cand.astype(np.float32).mean()) < 1e-4
        assert abs(res[r]["ema"] - cand.astype(np.float32).mean()) < 1e-4
    from scipy.stats import ttest_rel
    t, p = ttest_rel(cand, base)
    for r in (0, 1):
        assert abs(res[r]["ttest"][0] - t) < 1e-8 * abs(t) + 1e-9 and abs(res[r]["ttest"][1] - p / 2) < 1e-10
    torch.manual_seed(0)
    lin = torch.nn.Linear(8, 4)
    x = torch.randn(10, 8)
    (lin(x).pow(2).sum() / 10).backward()
    assert np.allclose(res[0]["grad"], lin.weight.grad.numpy(), atol=1e-6)
    assert np.allclose(res[1]["grad"], res[0]["grad"])
    # sharded BatchNorm == nn.BatchNorm1d on the concatenated batch (outputs, input / affine gradients, running stats)
    torch.manual_seed(1)
    bn = torch.nn.BatchNorm1d(6)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.5, 0.5)
    xb = (torch.randn(14, 6) * 2 + 1).requires_grad_(True)
    gy = torch.randn(14, 6)
    y = bn(xb)
    (y * gy).sum().backward()
    for r, (a, b) in ((0, (0, 9)), (1, (9, 14))):
        o = res[r]["bn"]
        assert np.allclose(o["y"], y.detach().numpy()[a:b], atol=1e-5)
        assert np.allclose(o["gx"], xb.grad.numpy()[a:b], atol=1e-5)
        assert np.allclose(o["gw"], bn.weight.grad.numpy(), atol=1e-5) and np.allclose(o["gb"], bn.bias.grad.numpy(), atol=1e-5)
        assert np.allclose(o["rm"], bn.running_mean.numpy(), atol=1e-6) and np.allclose(o["rv"], bn.running_var.numpy(), atol=1e-5)
        assert o["nbt"] == 1


def test_shard_bounds_cover_everything():
    from evrp_b200 import train as T
    for n in (0, 1, 7, 8192, 10240):
        for w in (1, 2, 3, 8):
            spans = [T.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
