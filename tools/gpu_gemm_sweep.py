"""Sweep split-K / tile-N choices for the wgrad shapes of the step (fp32 accumulate output)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalsum_b200 import ops
dev = "cuda"
def timeit(fn, n=20):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1000
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (M, N, K) in [(1024, 1024, 18432), (3072, 1024, 18432), (1024, 4096, 18432), (4096, 1024, 18432), (2048, 1024, 50544), (1024, 1024, 55296)]:
    A = torch.randn(K, M, device=dev).to(torch.bfloat16)
    B = torch.randn(K, N, device=dev).to(torch.bfloat16)
    out = torch.zeros(M, N, device=dev)
    print("wgrad M=%d N=%d K=%d" % (M, N, K))
    for bn in (128, 256):
        row = []
        for sp in (0, 1, 2, 3, 4, 5, 6, 8, 9, 10, 12):
            try:
                t = timeit(lambda: ops.gemm(A, B, out, a_t=True, b_t=True, accumulate=True, splits=sp, block_n=bn))
                row.append("s%d:%.1f" % (sp, t))
            except Exception as ex:
                row.append("s%d:err" % sp)
        print("  bn=%d  " % bn + "  ".join(row) + "   (us; best possible at 1.4 PF: %.1f)" % (2.0 * M * N * K / 1.4e9))
