"""GPU helper: time of the tcgen05 weight-gradient kernel (csrc/evrp_wgrad.cu) per training-path shape, against cuBLAS fp32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import evrp_b200
from evrp_b200 import policy_math as PM
torch.backends.cuda.matmul.allow_tf32 = False
shapes = [("qkv", 108544, 384, 128), ("wo", 108544, 128, 128), ("ff1", 108544, 512, 128), ("ff2", 108544, 128, 512),
          ("fft1", 160768, 512, 128), ("fft2", 160768, 128, 512)]
for name, M, N, K in shapes:
    gy, x = torch.randn(M, N, device="cuda"), torch.randn(M, K, device="cuda")
    def t(fn, n=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    a = t(lambda: PM.tc_wgrad(gy, x, True))
    b = t(lambda: gy.t() @ x)
    gb = (M * (N + K) * 4) / 1e9
    print(f"{name:5s} M={M} N={N} K={K}: tcgen05 {a*1e3:7.1f} us ({gb/a*1e3:6.0f} GB/s of operand bytes, {2*M*N*K/a/1e9:6.1f} TFLOP/s) | cuBLAS fp32 {b*1e3:7.1f} us")
