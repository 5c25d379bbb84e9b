"""Runs the two fused-activation GEMMs of the FFN block at config-2 shape (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalsum_b200 import ops
dev = "cuda"; T, D, F = 18432, 1024, 4096
x = torch.randn(T, D, device=dev).to(torch.bfloat16)
w1 = (torch.randn(F, D, device=dev) * 0.02).to(torch.bfloat16)
w2 = (torch.randn(D, F, device=dev) * 0.02).to(torch.bfloat16)
b1 = torch.randn(F, device=dev) * 0.02
h = torch.empty(T, F, device=dev, dtype=torch.bfloat16)
a = torch.empty(T, F, device=dev, dtype=torch.bfloat16)
dy = torch.randn(T, D, device=dev).to(torch.bfloat16)
da = torch.empty(T, F, device=dev, dtype=torch.bfloat16)
plain = torch.empty(T, F, device=dev, dtype=torch.bfloat16)
n = int(os.environ.get("N_ITER", "3"))
def t(fn):
    for _ in range(2): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1000
print("fc1 fwd  (GELU + save preact)  %.1f us" % t(lambda: ops.gemm(x, w1, h, bias=b1, act=ops.ACT_GELU, aux=a, aux_mode=ops.AUX_STORE_PREACT)))
print("fc2 dgrad (x GELU'(preact))    %.1f us" % t(lambda: ops.gemm(dy, w2, da, b_t=True, act=ops.ACT_GELU, aux=a, aux_mode=ops.AUX_MUL_DACT)))
print("plain same shape (bias)        %.1f us" % t(lambda: ops.gemm(x, w1, plain, bias=b1)))
print("plain + GELU (no save)         %.1f us" % t(lambda: ops.gemm(x, w1, plain, bias=b1, act=ops.ACT_GELU)))
