"""A few training-step GEMM shapes for ncu / CUDA-event A/B of the CTA-pair multicast kernel (MMSUM_GEMM_CLUSTER=0 switches it off):
   ncu --set full --clock-control none -k regex:gemm_tcgen05 -o gpurun_out/gemm_cl python tools/profile_gemm_cluster.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodalsum_b200 import ops

dev = "cuda"
torch.manual_seed(0)
bf = lambda *s: (torch.randn(*s, device=dev) * 0.05).to(torch.bfloat16)
T = 18432
x, x4 = bf(T, 1024), bf(T, 4096)
w11, w41, w14 = bf(1024, 1024), bf(4096, 1024), bf(1024, 4096)
mem, wkv = bf(50544, 1024), bf(2048, 1024)
g11 = torch.zeros(1024, 1024, device=dev)
flush = torch.empty(96 << 20, device=dev, dtype=torch.float32)
CASES = [
    ("fprop 18432x1024x1024", 2.0 * T * 1024 * 1024, lambda: ops.gemm(x, w11)),
    ("fprop 18432x4096x1024 (fc1)", 2.0 * T * 4096 * 1024, lambda: ops.gemm(x, w41)),
    ("fprop 18432x1024x4096 (fc2)", 2.0 * T * 1024 * 4096, lambda: ops.gemm(x4, w14)),
    ("fprop 50544x2048x1024 (K|V)", 2.0 * 50544 * 2048 * 1024, lambda: ops.gemm(mem, wkv)),
    ("dgrad 18432x1024x1024 (B MN-major)", 2.0 * T * 1024 * 1024, lambda: ops.gemm(x, w11, b_t=True)),
    ("wgrad 1024x1024x18432 (split-K, fp32 +=)", 2.0 * T * 1024 * 1024, lambda: ops.gemm(x, x, g11, a_t=True, b_t=True, accumulate=True)),
]
ncu = os.environ.get("MMSUM_NCU") == "1"
for name, flops, fn in CASES:
    for _ in range(0 if ncu else 3):
        fn()
    tot, n = 0.0, (1 if ncu else 20)
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    ms = tot / n
    print("%-44s %8.1f us  %7.1f TFLOP/s  (cluster %s)" % (name, ms * 1e3, flops / ms / 1e9, os.environ.get("MMSUM_GEMM_CLUSTER", "1")), flush=True)
