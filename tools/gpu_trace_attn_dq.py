"""clock64 timeline of CTA 0 of the dQ kernel (build with MMSUM_TRACE=1)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalsum_b200 import ops, _lib
os.environ["MMSUM_ATTN_BWD_PART"] = "1"
exec(open(os.path.join(os.path.dirname(__file__), "gpu_bench_attn.py")).read().split("def timeit")[0])
ops.attn_fwd(a); ops.attn_bwd(a); torch.cuda.synchronize()
ops.attn_bwd(a); torch.cuda.synchronize()
buf = (C.c_longlong * (8 * 512))()
_lib.lib().mmsum_debug_read_trace(buf)
t = [[buf[r * 512 + i] for i in range(512)] for r in range(8)]
t0 = t[3][0]
print("item | MMA: S/dP(i+1) issued, ds_full ok, dQ issued | softmax: start, sdp_full ok, pass1 done, bar done, pass2 done")
for i in range(19):
    print(i, [t[1][3 * i + j] - t0 for j in range(3)], [t[3][5 * i + j] - t0 for j in range(5)])
