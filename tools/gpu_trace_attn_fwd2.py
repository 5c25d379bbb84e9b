"""clock64 timeline of CTA 0 of the v2 forward attention kernel (build with MMSUM_TRACE=1)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalsum_b200 import ops, _lib
exec(open(os.path.join(os.path.dirname(__file__), "gpu_bench_attn.py")).read().split("def timeit")[0])
ops.attn_fwd(a); torch.cuda.synchronize()
ops.attn_fwd(a); torch.cuda.synchronize()
buf = (C.c_longlong * (8 * 512))()
_lib.lib().mmsum_debug_read_trace(buf)
t = [[buf[r * 512 + i] for i in range(512)] for r in range(8)]
t0 = min(x for x in t[3][:6] if x > 0)
print("item | MMA: p_full ok, o_free ok, PV issued | group: start, s_full ok, max done, P stored, o_full ok, O read")
for i in range(19):
    g, k = i & 1, i >> 1
    mm = [t[1][3 * i + j] - t0 for j in range(3)]
    gg = [t[3 + g][6 * k + j] - t0 for j in range(6)]
    print(i, "AB"[g], mm, gg)
