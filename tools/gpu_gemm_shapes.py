"""Per-shape GEMM times inside a real config-2 training step (CUDA events around every mmsum_gemm_bf16 call)."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalsum_b200 import ops
from multimodalsum_b200.modules import MultimodalSum, YelpTableEncoder
from multimodalsum_b200.synth import ModelConfig, make_batch

dev = torch.device("cuda", 0)
cfg = ModelConfig(dataset="yelp", dropout=0.1)
torch.manual_seed(0)
model = MultimodalSum(TableEncoder=YelpTableEncoder, config=cfg, label_smoothing=0.1).to(dev).train()
b = make_batch(cfg, 16, seed=1, fixed_len=100, n_valid_imgs=10).to(dev)

def step():
    loss = model(b.reviews, b.reviews_mask, b.reviews_rating, b.field, b.field_value, b.img, b.img_mask)[0]
    model.zero_grad(set_to_none=True)
    loss.backward()

for _ in range(3):
    step()
torch.cuda.synchronize()
events = []
orig_gemm, orig_cat = ops.gemm, ops.gemm_cat

def wrap_gemm(A, B, out=None, **kw):
    a_t, b_t = kw.get("a_t", False), kw.get("b_t", False)
    K, M = A.shape if a_t else A.shape[::-1]
    N = B.shape[1] if b_t else B.shape[0]
    key = (M, N, K, "T" if a_t else "N", "T" if b_t else "N", "f32" if (out is not None and out.dtype == torch.float32) or kw.get("out_dtype") == torch.float32 else "bf16",
           "acc" if kw.get("accumulate") else "", "act%d" % kw.get("act", 0), "aux%d" % kw.get("aux_mode", 0), "bias" if kw.get("bias") is not None else "")
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); r = orig_gemm(A, B, out, **kw); e.record()
    events.append((key, 2.0 * M * N * K, s, e))
    return r

def wrap_cat(A, A2, B, out=None, **kw):
    M, K = A.shape[0], A.shape[1] + A2.shape[1]
    key = (M, B.shape[0], K, "N", "N", "bf16", "", "cat", "", "bias" if kw.get("bias") is not None else "")
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); r = orig_cat(A, A2, B, out, **kw); e.record()
    events.append((key, 2.0 * M * B.shape[0] * K, s, e))
    return r

import multimodalsum_b200.engine as engine
ops.gemm, ops.gemm_cat = wrap_gemm, wrap_cat
for _ in range(3):
    step()
torch.cuda.synchronize()
agg = collections.OrderedDict()
for key, fl, s, e in events:
    a = agg.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1; a[1] += fl; a[2] += s.elapsed_time(e)
tot_ms = sum(a[2] for a in agg.values()) / 3
tot_fl = sum(a[1] for a in agg.values()) / 3
print("total %.2f ms/step, %.1f TFLOP/step, %.0f TF/s" % (tot_ms, tot_fl / 1e12, tot_fl / tot_ms / 1e9))
print("%-62s %5s %9s %8s %7s" % ("shape (M,N,K,opA,opB,out,...)", "n", "ms/step", "us/call", "TF/s"))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    print("%-62s %5d %9.3f %8.1f %7.0f" % (" ".join(str(k) for k in key if k != ""), a[0] // 3, a[2] / 3, a[2] / a[0] * 1000, a[1] / a[2] / 1e9))
