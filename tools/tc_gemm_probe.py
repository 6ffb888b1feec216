"""GPU helper: evrp_gemm_3xtf32 against an fp64 product for the shapes of the training path (incl. ragged M, wide-range data)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from evrp_b200 import policy_math as PM
torch.manual_seed(0)
for M in (1272, 1280, 416, 27136):
    for (N, K) in ((384, 128), (128, 128), (512, 128), (128, 512), (128, 384)):
        for wide in (0, 1):
            x = torch.randn(M, K, device="cuda")
            if wide: x = x * torch.exp(3 * torch.randn(M, 1, device="cuda"))
            W = torch.randn(N, K, device="cuda") * 0.1
            y = PM.tc_gemm(x, W)
            ref = (x.double() @ W.double().t())
            err = (y.double() - ref).abs()
            rowscale = ref.abs().amax(1, keepdim=True) + 1e-30
            print(f"M={M:6d} N={N:3d} K={K:3d} wide={wide}: max rel-to-row-max err {(err / rowscale).max().item():.2e}  global rel {err.norm().item() / ref.norm().item():.2e}")
