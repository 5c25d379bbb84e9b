"""Micro-benchmark of the HBM-bound row kernels at config-2 sizes (T = 18432 rows)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalsum_b200 import ops
dev = "cuda"; D = 1024; T = 18432
torch.manual_seed(0)
bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)
res, y, d1, d2 = bf(T, D), bf(T, D), bf(T, D), bf(T, D)
out, dres, dy = torch.empty_like(res), torch.empty_like(res), torch.empty_like(res)
gamma, beta = torch.ones(D, device=dev), torch.zeros(D, device=dev)
mean, rstd = torch.empty(T, device=dev), torch.empty(T, device=dev)
dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
big = bf(T, 4096); bsum = torch.zeros(4096, device=dev)
flush = torch.empty(64 << 20, device=dev, dtype=torch.float32)
def timeit(fn, n=20):
    for _ in range(3): fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / n * 1000
MB = T * D * 2 / 1e6
for p in (0.0, 0.1):
    t = timeit(lambda: ops.add_ln_fwd(res, y, gamma, beta, out, mean, rstd, p, 1, 1)); print("add_ln_fwd p=%.1f %.1f us  %.0f GB/s" % (p, t, 3 * MB / t * 1e3))
    t = timeit(lambda: ops.add_ln_bwd(d1, d2, res, y, gamma, mean, rstd, dres, dy if p > 0 else dres, dg, db, p, 1, 1)); print("add_ln_bwd p=%.1f %.1f us  %.0f GB/s" % (p, t, (5 + (1 if p > 0 else 0)) * MB / t * 1e3))
t = timeit(lambda: ops.colsum(big, bsum)); print("colsum [T,4096] %.1f us  %.0f GB/s" % (t, 4 * MB / t * 1e3))
t = timeit(lambda: ops.colsum(res, dg)); print("colsum [T,1024] %.1f us  %.0f GB/s" % (t, MB / t * 1e3))
