"""Micro-benchmark of the attention kernels at BASELINE config-2 shapes (B=16 businesses)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalsum_b200 import ops
dev = "cuda"; D = 1024; H = 16; S = 128
B, R, F_, n_img, ik = 16, 9, 47, 10, 196
N, T = B * R, B * R * S
Tm = T + B * F_ + B * n_img * ik
Et = R + 1 + n_img
torch.manual_seed(0)
qc = torch.randn(T, D, device=dev).to(torch.bfloat16)
kv = torch.randn(Tm, 2 * D, device=dev).to(torch.bfloat16)
lens = torch.full((B, R), 100, device=dev)
tvalid = torch.arange(S, device=dev)[None, None, :] < lens[:, :, None]
mem_valid = torch.cat([tvalid.reshape(-1), torch.ones(B * F_ + B * n_img * ik, device=dev, dtype=torch.bool)]).to(torch.uint8)
ent_valid = torch.ones(B, Et, device=dev, dtype=torch.uint8)
inv_n = torch.tensor([1 / 8, 1.0, 1 / 10], device=dev).repeat(N, 1).contiguous()
A3 = torch.empty(3, T, D, device=dev, dtype=torch.bfloat16)
lse = torch.empty(N, H, Et, S, device=dev)
delta = torch.empty(N, H, Et, S, device=dev)
dqc = torch.empty(T, D, device=dev, dtype=torch.bfloat16)
dkv = torch.empty(Tm, 2 * D, device=dev, dtype=torch.bfloat16)
mods = [(0, 0, R, S, 1, 0), (T, T * D, 1, F_, 0, R), (T + B * F_, 2 * T * D, n_img, ik, 0, R + 1)]
kw = dict(Q=qc, ldq=D, q_col=0, KV=kv, ldkv=2 * D, k_col=0, v_col=D, O=A3, ldo=D, LSE=lse, key_valid=mem_valid,
          ent_valid=ent_valid, inv_n=inv_n, n_qseq=N, H=H, R=R, causal=0, E_total=Et, scale=0.125, mods=mods,
          DELTA=delta, dQ=dqc, lddq=D, dq_col=0, dKV=dkv, lddkv=2 * D, dk_col=0, dv_col=D)
a = ops.attn_args(**kw)
qkv = torch.randn(T, 3 * D, device=dev).to(torch.bfloat16)
ctx = torch.empty(T, D, device=dev, dtype=torch.bfloat16)
dqkv = torch.empty(T, 3 * D, device=dev, dtype=torch.bfloat16)
lse_s = torch.empty(N, H, 1, S, device=dev)
ks = dict(Q=qkv, ldq=3 * D, q_col=0, KV=qkv, ldkv=3 * D, k_col=D, v_col=2 * D, O=ctx, ldo=D, LSE=lse_s, key_valid=tvalid.reshape(-1).to(torch.uint8),
          ent_valid=None, inv_n=None, n_qseq=N, H=H, R=1, causal=1, E_total=1, scale=0.125, mods=[(0, 0, 1, S, 0, 0)],
          DELTA=delta, dQ=dqkv, lddq=3 * D, dq_col=0, dKV=dqkv, lddkv=3 * D, dk_col=D, dv_col=2 * D)
b = ops.attn_args(**ks)

def timeit(fn, n=10):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1000
keys = 8 * 128 + 47 + 10 * 196
unit = 2.0 * N * H * S * keys * 64  # one QK^T-sized product
t = timeit(lambda: ops.attn_fwd(a)); print("cross fwd  %.0f us  (%.0f TF/s algorithmic, 2 products)" % (t, 2 * unit / t / 1e6))
t = timeit(lambda: ops.attn_bwd(a)); print("cross bwd  %.0f us  (dq+dkv; %.0f TF/s algorithmic, 5 products)" % (t, 5 * unit / t / 1e6))
t = timeit(lambda: ops.attn_fwd(b)); print("self  fwd  %.0f us" % t)
t = timeit(lambda: ops.attn_bwd(b)); print("self  bwd  %.0f us" % t)
