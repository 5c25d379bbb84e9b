"""Debug: per-role timeline of CTA 0 of the dK/dV backward kernel (needs a MMSUM_TRACE=1 build)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
exec(open(os.path.join(os.path.dirname(__file__), "gpu_bench_attn.py")).read().split("def timeit")[0])
from multimodalsum_b200 import _lib
ops.attn_fwd(a); ops.attn_bwd(a); torch.cuda.synchronize()
ops.attn_bwd(a); torch.cuda.synchronize()
buf = (C.c_longlong * (8 * 512))()
_lib.lib().mmsum_debug_read_trace(buf)
t = [[buf[r * 512 + i] for i in range(512)] for r in range(8)]
t0 = min(x for x in t[5][:40] + t[6][:60] if x > 0)
f = lambda x: (x - t0) if x > 0 else -1
print("step | MMA: qd_full ok, sdp_empty ok, pds_full ok, MMA2 issued | softmax: step start, bar done, sdp_full ok, tmem ld done, pds_free ok, compute done")
for s in range(8):
    print(s, [f(t[5][4*s+k]) for k in range(4)], "|", [f(t[6][6*s+k]) for k in range(6)])
