"""Debug: per-role timeline of CTA 0 of the cross-attention forward kernel (needs a MMSUM_TRACE=1 build)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
exec(open(os.path.join(os.path.dirname(__file__), "gpu_bench_attn.py")).read().split("def timeit")[0])
from multimodalsum_b200 import _lib
ops.attn_fwd(a); torch.cuda.synchronize()
ops.attn_fwd(a); torch.cuda.synchronize()
buf = (C.c_longlong * (8 * 512))()
rc = _lib.lib().mmsum_debug_read_trace(buf)
t = [[buf[r * 512 + i] for i in range(512)] for r in range(8)]
t0 = min(x for x in t[0][:19] + t[1][:38] + t[3][:114] if x > 0)
print("item | K load issue | MMA1: k_full ok, s_empty ok | MMA2: v_full ok, p_full ok | softmax: start, s_full ok, max done, bar done, mma2_done(i-1) ok, exp done")
for i in range(19):
    f = lambda x: (x - t0) if x > 0 else -1
    print(i, f(t[0][i]), "|", f(t[1][2*i]), f(t[1][2*i+1]), "|", f(t[2][2*i]), f(t[2][2*i+1]), "|", [f(t[3][6*i+k]) for k in range(6)], "| MMA2 issued, done:", f(t[4][2*i]), f(t[4][2*i+1]))
