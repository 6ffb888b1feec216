"""Mint the committed golden fixtures under tests/golden/ from the PATCHED REFERENCE ITSELF.

Run in the build container only (needs /root/reference):   python -m oracle.make_golden

The reference has no tests or golden vectors of its own, so these files are the
parity pin: inputs and outputs of /root/reference/EM-EVRP (patches P1+P2, oracle/ref_shim.py)
executed on CPU with torch fp32.  Files:

  weights_seed0.npz        state_dict of AttentionModel built after torch.manual_seed(0)
                           (trainer.py:261-272 kwargs), plus a BatchNorm perturbation recipe flag
  greedy_n{20,50,100}.npz  instances (VehicleRoutingDataset seed 12345) + greedy eval-mode rollout:
                           pi, ll, R and the per-step trace (mask fed to the step, log_p, dynamic
                           after the step)
  greedy_bn_n20.npz        same with non-trivial BatchNorm running stats / affine (perturb_bn)
  sample_n20.npz           sampled rollout (actions recorded) + teacher-forced REINFORCE gradients of
                           all 50 parameters, eval-mode BN and train-mode BN
  dm_n20.npz               DM objective: reward := distance (DM-EVRP/problems/EVRP.py:231,280)
"""
import os
import sys

import numpy as np
import torch
from torch.utils.data import DataLoader

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def perturb_bn(sd, seed=7):
    """Deterministic (numpy) non-trivial BatchNorm statistics/affine; works on any state_dict with
    the reference's keys.  Returns a new dict of torch tensors."""
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


class Tracer:
    """Records per-step tensors by wrapping the model's injected callbacks (plain attributes,
    nets/DRLModel.py:67-68) and _get_log_p."""

    def __init__(self, m):
        self.m = m
        self.masks, self.logps, self.dyns, self.dist = [], [], [], []
        self._ud, self._glp = m.update_dynamic, m._get_log_p
        m.update_dynamic = self.ud
        m._get_log_p = self.glp

    def ud(self, dynamic, distances, slope, now, chosen):
        ar = torch.arange(distances.size(0))
        self.dist.append(distances[ar, now, chosen].clone())
        nd, r = self._ud(dynamic, distances, slope, now, chosen)
        self.dyns.append(nd.clone())
        return nd, r

    def glp(self, fixed, now_idx, mask, route_feature, embeddings):
        self.masks.append(mask.clone())
        lp = self._glp(fixed, now_idx, mask, route_feature, embeddings)
        self.logps.append(lp[:, 0, :].detach().clone())
        return lp

    def close(self):
        self.m.update_dynamic, self.m._get_log_p = self._ud, self._glp

    def arrays(self):
        return dict(t_mask=torch.stack(self.masks, 1).numpy().astype(np.uint8),
                    t_log_p=torch.stack(self.logps, 1).numpy(),
                    t_dynamic=torch.stack(self.dyns, 1).numpy(),
                    t_dist=torch.stack(self.dist, 1).numpy())


def batch_of(ds, B):
    return next(iter(DataLoader(ds, B, False)))


def inputs_dict(ds, bat, B):
    static, dynamic, D, G = bat
    return dict(static=static.numpy().astype(np.float32), dynamic=dynamic.numpy().astype(np.float32),
                elevation=ds.Elevation[:B].numpy().astype(np.float32),
                distances=D.numpy(), slope=G.numpy())


def greedy_fixture(name, N, B, seed=12345, bn=False):
    ds, args = ref_shim.make_dataset(B, N, seed)
    m = ref_shim.make_model(ds, args, 0)
    if bn:
        m.load_state_dict(perturb_bn(m.state_dict()))
    m.eval(); m.set_decode_type("greedy")
    bat = batch_of(ds, B)
    tr = Tracer(m)
    with torch.no_grad():
        pi, ll, R = m(bat)
    tr.close()
    d = inputs_dict(ds, bat, B)
    d.update(pi=pi.numpy(), ll=ll.numpy(), R=R.numpy(), bn_perturbed=np.array(bn), **tr.arrays())
    np.savez_compressed(os.path.join(OUT, name), **d)
    print(name, "T=", pi.shape[1], "mean cost", float(R.sum(1).mean()))


def teacher_forced_grads(m, bat, pi, adv):
    """REINFORCE gradient with the reference's own forward, actions forced to ``pi``
    (trainer.py:115-124 with _select_node patched)."""
    it = iter(pi.t())
    orig = m._select_node
    m._select_node = lambda probs, mask: next(it)
    m.zero_grad()
    pi2, ll, R = m(bat)
    m._select_node = orig
    assert torch.equal(pi2, pi)
    loss = (adv * ll.sum(1)).mean()
    loss.backward()
    return ll.detach(), R, float(loss), {k: p.grad.detach().clone() for k, p in m.named_parameters()}


def sample_fixture(name, N, B, seed=12346):
    ds, args = ref_shim.make_dataset(B, N, seed)
    m = ref_shim.make_model(ds, args, 0)
    m.load_state_dict(perturb_bn(m.state_dict()))
    bat = batch_of(ds, B)
    m.eval(); m.set_decode_type("sample")
    torch.manual_seed(99)
    with torch.no_grad():
        pi, ll0, R = m(bat)
    cost = R.sum(1)
    adv = (cost - cost.mean()).detach()
    d = inputs_dict(ds, bat, B)
    d.update(pi=pi.numpy(), R=R.numpy(), adv=adv.numpy())
    # eval-mode BN gradients
    ll, R2, loss, g = teacher_forced_grads(m, bat, pi, adv)
    assert torch.equal(R2, R)
    d.update(ll_eval=ll.numpy(), loss_eval=np.array(loss))
    d.update({"g_eval/" + k: v.numpy() for k, v in g.items()})
    # train-mode BN gradients (batch statistics; running stats get updated as a side effect)
    m.train()
    ll, R3, loss, g = teacher_forced_grads(m, bat, pi, adv)
    d.update(ll_train=ll.numpy(), loss_train=np.array(loss))
    d.update({"g_train/" + k: v.numpy() for k, v in g.items()})
    d.update({"bn_after/" + k: v.numpy() for k, v in m.state_dict().items() if "running" in k or "tracked" in k})
    np.savez_compressed(os.path.join(OUT, name), **d)
    print(name, "T=", pi.shape[1], "loss eval/train", float(d["loss_eval"]), float(d["loss_train"]))


def dm_fixture(name, N, B, seed=12345):
    """DM objective = EM dynamics + reward := distance."""
    ds, args = ref_shim.make_dataset(B, N, seed)
    m = ref_shim.make_model(ds, args, 0)
    m.eval(); m.set_decode_type("greedy")
    bat = batch_of(ds, B)
    tr = Tracer(m)
    with torch.no_grad():
        pi, ll, R = m(bat)
    tr.close()
    d = inputs_dict(ds, bat, B)
    a = tr.arrays()
    d.update(pi=pi.numpy(), ll=ll.numpy(), E=R.numpy(), R=a["t_dist"])
    np.savez_compressed(os.path.join(OUT, name), **d)
    print(name, "T=", pi.shape[1], "mean distance", float(a["t_dist"].sum(1).mean()))


def am_fixture(name, N, B, seed=12345):
    """nets/AM.py variant: greedy tours, per-step log-probs, and a sampled rollout with its REINFORCE
    gradients; weights torch.manual_seed(0) -> weights_am_seed0.npz"""
    ds, args = ref_shim.make_dataset(B, N, seed)
    m = ref_shim.make_model_am(ds, args, 0)
    np.savez_compressed(os.path.join(OUT, "weights_am_seed0.npz"), **{k: v.numpy() for k, v in m.state_dict().items()})
    m.eval(); m.set_decode_type("greedy")
    bat = batch_of(ds, B)
    cap = {}
    orig = m._calc_log_likelihood
    def spy(_log_p, a):
        cap["log_p"] = _log_p.detach().numpy().copy()
        return orig(_log_p, a)
    m._calc_log_likelihood = spy
    with torch.no_grad():
        pi, ll, R = m(bat)
    d = inputs_dict(ds, bat, B)
    d.update(pi=pi.numpy(), ll=ll.numpy(), R=R.numpy(), t_log_p=cap["log_p"])
    # sampled rollout + gradients of mean((R.sum - mean) * ll.sum) (eval-mode BatchNorm)
    m.set_decode_type("sample")
    torch.manual_seed(seed + 1)
    m.zero_grad()
    spi, sll, sR = m(bat)
    adv = (sR.sum(1) - sR.sum(1).mean()).detach()
    (adv * sll.sum(1)).mean().backward()
    d.update(s_pi=spi.numpy(), s_ll=sll.detach().numpy(), s_R=sR.numpy(), s_adv=adv.numpy(),
             **{"grad." + k: p.grad.numpy().copy() for k, p in m.named_parameters()})
    np.savez_compressed(os.path.join(OUT, name), **d)
    print(name, "T=", pi.shape[1], "mean cost", float(R.sum(1).mean()), "sampled T=", spi.shape[1])


def greedy_lite_fixture(name, N, B, seed):
    """many instances at one size, without the per-step state / log-prob trace: inputs, tours, log-likelihoods, rewards
    and the per-step masks (bit-packed)"""
    ds, args = ref_shim.make_dataset(B, N, seed)
    m = ref_shim.make_model(ds, args, 0)
    m.eval(); m.set_decode_type("greedy")
    bat = batch_of(ds, B)
    tr = Tracer(m)
    with torch.no_grad():
        pi, ll, R = m(bat)
    tr.close()
    d = inputs_dict(ds, bat, B)
    d.update(pi=pi.numpy(), ll=ll.numpy(), R=R.numpy(), bn_perturbed=np.array(False),
             t_mask_packed=np.packbits(tr.arrays()["t_mask"].astype(bool), axis=-1), n_nodes=np.array(N + 6))
    np.savez_compressed(os.path.join(OUT, name), **d)
    print(name, "T=", pi.shape[1], "mean cost", float(R.sum(1).mean()))


def sample_trace_fixture(name, N, B, seed, torch_seed):
    """SAMPLED rollout of the reference (torch.multinomial stream of `torch_seed`) with the full per-step trace"""
    ds, args = ref_shim.make_dataset(B, N, seed)
    m = ref_shim.make_model(ds, args, 0)
    m.eval(); m.set_decode_type("sample")
    bat = batch_of(ds, B)
    tr = Tracer(m)
    torch.manual_seed(torch_seed)
    with torch.no_grad():
        pi, ll, R = m(bat)
    tr.close()
    d = inputs_dict(ds, bat, B)
    d.update(pi=pi.numpy(), ll=ll.numpy(), R=R.numpy(), bn_perturbed=np.array(False), **tr.arrays())
    np.savez_compressed(os.path.join(OUT, name), **d)
    print(name, "T=", pi.shape[1], "mean cost", float(R.sum(1).mean()))


GRAD_STRIDE = 4


def grad_fixture(name, N, B, seed, torch_seed):
    """REINFORCE gradients at a large graph size (eval- and train-mode BatchNorm) teacher-forced on a sampled rollout.
    To keep the file small every gradient is stored as its flattened entries [::GRAD_STRIDE] plus its L2 norm."""
    ds, args = ref_shim.make_dataset(B, N, seed)
    m = ref_shim.make_model(ds, args, 0)
    m.load_state_dict(perturb_bn(m.state_dict()))
    bat = batch_of(ds, B)
    m.eval(); m.set_decode_type("sample")
    torch.manual_seed(torch_seed)
    with torch.no_grad():
        pi, ll0, R = m(bat)
    cost = R.sum(1)
    adv = (cost - cost.mean()).detach()
    d = inputs_dict(ds, bat, B)
    d.update(pi=pi.numpy(), R=R.numpy(), adv=adv.numpy(), stride=np.array(GRAD_STRIDE))
    for mode in ("eval", "train"):
        m.train(mode == "train")
        ll, R2, loss, g = teacher_forced_grads(m, bat, pi, adv)
        assert torch.equal(R2, R)
        d["ll_" + mode], d["loss_" + mode] = ll.numpy(), np.array(loss)
        for k, v in g.items():
            d[f"g_{mode}/" + k] = v.reshape(-1)[::GRAD_STRIDE].numpy().copy()
            d[f"n_{mode}/" + k] = np.array(float(v.norm()))
    d.update({"bn_after/" + k: v.numpy() for k, v in m.state_dict().items() if "running" in k or "tracked" in k})
    np.savez_compressed(os.path.join(OUT, name), **d)
    print(name, "T=", pi.shape[1], "loss eval/train", float(d["loss_eval"]), float(d["loss_train"]))


def main_round2():
    """fixtures added in round 2 (the round-1 files are left byte-identical)"""
    greedy_lite_fixture("greedy_n100_64.npz", 100, 64, seed=54321)
    sample_trace_fixture("sample_n50.npz", 50, 8, seed=4711, torch_seed=5)
    grad_fixture("sample_grad_n100.npz", 100, 12, seed=2025, torch_seed=7)


def main():
    os.makedirs(OUT, exist_ok=True)
    ds, args = ref_shim.make_dataset(2, 20, 1)
    m = ref_shim.make_model(ds, args, 0)
    np.savez_compressed(os.path.join(OUT, "weights_seed0.npz"),
                        **{k: v.numpy() for k, v in m.state_dict().items()})
    greedy_fixture("greedy_n20.npz", 20, 32)
    greedy_fixture("greedy_bn_n20.npz", 20, 16, bn=True)
    greedy_fixture("greedy_n50.npz", 50, 8)
    greedy_fixture("greedy_n100.npz", 100, 6)
    greedy_fixture("greedy_n35.npz", 35, 8, seed=31337)               # a size / seed off the reference's 20 / 50 / 100 grid
    greedy_fixture("greedy_bn_n50.npz", 50, 6, seed=2024, bn=True)
    sample_fixture("sample_n20.npz", 20, 16)
    dm_fixture("dm_n20.npz", 20, 8)
    am_fixture("am_n20.npz", 20, 16)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "round2":
        main_round2()
    else:
        main()
