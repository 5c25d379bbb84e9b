"""One launch of every HBM-bound row kernel at the config-2 sizes (T = 18432 rows, V = 50265), for
   ncu --set full --clock-control none -k regex:'add_ln|embed_ln|ce_fwd|colsum|gate_|cast_f32|adamw|sumsq' -o gpurun_out/rows python tools/profile_rows.py
Also prints CUDA-event timings (L2 flushed between launches) and the algorithmic GB/s of each kernel."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodalsum_b200 import _lib, ops

dev = "cuda"
D, T, V, S = 1024, 18432, 50265, 128
torch.manual_seed(0)
bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)
f32 = lambda *s: torch.randn(*s, device=dev)
res, y, d1, d2 = bf(T, D), bf(T, D), bf(T, D), bf(T, D)
out, dres, dy = torch.empty_like(res), torch.empty_like(res), torch.empty_like(res)
gamma, beta = torch.ones(D, device=dev), torch.zeros(D, device=dev)
mean, rstd = torch.empty(T, device=dev), torch.empty(T, device=dev)
dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
big = bf(T, 4096)
bsum = torch.zeros(4096, device=dev)
E, P = f32(V, D) * 0.02, f32(1026, D) * 0.02
dE, dP = torch.zeros_like(E), torch.zeros_like(P)
ids = torch.randint(3, V, (T,), device=dev, dtype=torch.int32)
rd, remb, dremb = f32(T // S), f32(D) * 0.02, torch.zeros(D, device=dev)
dz = torch.empty(T, D, device=dev)
ldv = (V + 7) // 8 * 8
logits = (torch.randn(T, ldv, device=dev) * 0.5).to(torch.bfloat16)
labels = torch.randint(0, V, (T,), device=dev, dtype=torch.int32)
loss_rows, loss = torch.empty(T, device=dev), torch.empty(1, device=dev)
o3, u, ab, yc = bf(3, T, D), bf(2, T, D), bf(2, T, D), bf(T, D)
du, do3, dca, dcb = bf(2, T, D), bf(3, T, D), bf(T, 2 * D), bf(T, 2 * D)
pres = torch.ones(T // (9 * S), 2, device=dev, dtype=torch.uint8)
NP = 460_852_224
w32, g32 = torch.zeros(NP, device=dev), torch.full((NP,), 1e-3, device=dev)
w16, m_, v_ = torch.empty(NP, device=dev, dtype=torch.bfloat16), torch.zeros(NP, device=dev), torch.zeros(NP, device=dev)
flags = torch.full((NP // 64,), 3, device=dev, dtype=torch.uint8)
partial, sumsq = torch.zeros(148 * 8, device=dev), torch.zeros(1, device=dev)
flush = torch.empty(96 << 20, device=dev, dtype=torch.float32)
lib = _lib.lib()
MB = T * D * 2 / 1e6


def adamw():
    s = ops._stream()
    ops.check(lib.mmsum_grad_sumsq(ops._ptr(g32), C.c_int64(NP), ops._ptr(partial), partial.numel(), ops._ptr(sumsq), s), "sumsq", 2)
    ops.check(lib.mmsum_adamw_step(ops._ptr(w32), ops._ptr(w16), ops._ptr(g32), ops._ptr(m_), ops._ptr(v_), ops._ptr(flags),
                                   C.c_int64(NP), C.c_float(1e-5), C.c_float(0.9), C.c_float(0.999), C.c_float(1e-6),
                                   C.c_float(0.01), C.c_float(1e-5), ops._ptr(sumsq), C.c_float(1.0), s), "adamw")


CASES = [
    # name, fn, algorithmic MB per launch
    ("add_ln_fwd p=0.1", lambda: ops.add_ln_fwd(res, y, gamma, beta, out, mean, rstd, 0.1, 1, 1), 3 * MB),
    ("add_ln_bwd p=0.1", lambda: ops.add_ln_bwd(d1, d2, res, y, gamma, mean, rstd, dres, dy, dg, db, 0.1, 1, 1), 6 * MB),
    ("embed_ln_fwd", lambda: ops.embed_ln_fwd(ids, E, P, rd, remb, gamma, beta, out, mean, rstd, T, S, 0.1, 1, 1), T * D * 4 / 1e6 + MB),
    ("embed_ln_bwd", lambda: ops.embed_ln_bwd(d1, d2, ids, E, P, rd, remb, gamma, mean, rstd, dE, dP, dremb, dg, db, dz, T, S, 1, 0.1, 1, 1),
     2 * MB + 3 * T * D * 4 / 1e6),
    ("ce fwd (loss)", lambda: ops.ce_fwd_bwd(logits, V, labels, 0.1, 0.0, None, loss_rows, loss, 1.0 / T, False), T * V * 2 / 1e6),
    ("ce bwd (grad in place)", lambda: ops.ce_fwd_bwd(logits, V, labels, 0.1, 1.0 / T, None, loss_rows, None, 0.0, True), 2 * T * V * 2 / 1e6),
    ("colsum [T,4096]", lambda: ops.colsum(big, bsum), 4 * MB),
    ("colsum [T,1024]", lambda: ops.colsum(res, dg), MB),
    ("gate_fwd", lambda: ops.gate_fwd(o3, u, pres, yc, ab, T, 9 * S, D), 8 * MB),
    ("gate_bwd_u", lambda: ops.gate_bwd_u(d1, o3, ab, du, T, D), 8 * MB),
    ("gate_bwd_o", lambda: ops.gate_bwd_o(d1, ab, dca, dcb, do3, T, D), 10 * MB),
    ("cast 460M", lambda: ops.cast_bf16(w32, w16), NP * 6 / 1e6),
    ("sumsq+adamw 460M", adamw, NP * 34 / 1e6),
]

peak = 6551.0
try:
    import json
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
NCU = os.environ.get("MMSUM_NCU") == "1"      # under ncu: exactly one launch per case
for name, fn, mb in CASES:
    for _ in range(0 if NCU else 2):
        fn()
    tot, n = 0.0, (1 if NCU else 10)
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    us = tot / n * 1000
    print("%-26s %9.1f us  %8.1f MB  %7.0f GB/s  frac %.3f of %.0f" % (name, us, mb, mb / us * 1e3, mb / us * 1e3 / peak, peak), flush=True)
