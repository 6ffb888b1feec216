"""GPU helper: REINFORCE gradient error against the N = 100 reference fixture with parts of the training path swapped for
their plain-PyTorch statements (isolates which kernel carries the error)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import torch.nn.functional as Fn
import test_gpu_parity as T
import evrp_b200
from evrp_b200 import policy_math as PM
weights = dict(T.load("weights_seed0.npz"))
g = T.load("sample_grad_n100.npz"); stride = int(g["stride"])

class TorchLinear:
    @staticmethod
    def apply(x, W, bias, resid, relu):
        y = Fn.linear(x, W, bias)
        if relu: y = torch.relu(y)
        return y + resid if resid is not None else y
class TorchAttn:
    @staticmethod
    def apply(qkv): return PM.self_attention_torch(qkv, 8)
class TorchLogp:
    @staticmethod
    def apply(q, kvl, mask_bits, pi, steps, clip, temp): return PM.pointer_log_prob_torch(q, kvl, mask_bits, pi, steps.long(), clip, temp, 8)

orig = (PM.TCLinear, PM.SelfAttention, PM.PointerLogProb)
for label, swap in (("all kernels", ()), ("torch linear", (0,)), ("torch attention", (1,)), ("torch logp", (2,)), ("all torch", (0, 1, 2))):
    PM.TCLinear, PM.SelfAttention, PM.PointerLogProb = orig
    if 0 in swap: PM.TCLinear = TorchLinear
    if 1 in swap: PM.SelfAttention = TorchAttn
    if 2 in swap: PM.PointerLogProb = TorchLogp
    for mode in ("eval",):
        m, _ = T.make_model(evrp_b200, 100, T.bn_perturb_numpy(weights)); m.train(mode == "train")
        pi = torch.from_numpy(g["pi"]).cuda(); m.zero_grad()
        pi2, ll, R = m(T.batch_of(g), forced_actions=pi)
        dll = np.abs(ll.detach().cpu().numpy() - g[f"ll_{mode}"]).max()
        (torch.from_numpy(g["adv"]).cuda() * ll.sum(1)).mean().backward()
        num = den = 0; worst = ("", 0)
        for k, p in m.named_parameters():
            ref = torch.from_numpy(g[f"g_{mode}/{k}"]).cuda(); got = p.grad.reshape(-1)[::stride]
            num += float((got-ref).norm()**2); den += float(ref.norm()**2)
            rel = float((got - ref).norm() / (ref.norm()+1e-30))
            if rel > worst[1] and float(g[f"n_{mode}/{k}"]) > 1: worst = (k, rel)
        print(f"{label:16s} {mode}: max dll {dll:.2e}  global grad rel err {(num/den)**0.5:.2e}  worst {worst[0]} {worst[1]:.2e}")
