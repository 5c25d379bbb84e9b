"""GPU bring-up check for the tcgen05 GEMM: prints error vs torch.matmul (checker only) per variant."""
import sys, time
import torch
sys.path.insert(0, ".")
from multimodalsum_b200 import ops

torch.manual_seed(0)
dev = "cuda"

def ref(A, B, a_t, b_t):
    Af = A.float().t() if a_t else A.float()
    Bf = B.float().t() if b_t else B.float()
    return Af @ Bf.t()

def run(M, N, K, a_t, b_t, out_f32=False, block_n=0, bias=False, act=0, accumulate=False, splits=0, aux_mode=0, tag=""):
    A = torch.randn((K, M) if a_t else (M, K), device=dev).to(torch.bfloat16)
    B = torch.randn((K, N) if b_t else (N, K), device=dev).to(torch.bfloat16)
    bias_t = torch.randn(N, device=dev) if bias else None
    out = None
    base = None
    if accumulate:
        base = torch.randn(M, N, device=dev)
        out = base.clone()
    aux = None
    if aux_mode == 1:
        aux = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    if aux_mode == 2:
        aux = torch.randn(M, N, device=dev).to(torch.bfloat16)
    try:
        D = ops.gemm(A, B, out, a_t=a_t, b_t=b_t, out_dtype=torch.float32 if out_f32 else torch.bfloat16,
                     bias=bias_t, act=act, accumulate=accumulate, splits=splits, block_n=block_n, aux=aux, aux_mode=aux_mode,
                     alpha=0.5)
        torch.cuda.synchronize()
    except Exception as e:
        print("FAIL(exc)", tag, M, N, K, a_t, b_t, repr(e)); return False
    R = 0.5 * ref(A, B, a_t, b_t)
    if bias: R = R + bias_t
    pre = R.clone()
    if aux_mode == 2:
        h = aux.float()
        if act == 1:
            g = 0.5 * (1 + torch.erf(h / 2**0.5)) + h * torch.exp(-0.5 * h * h) / (2 * 3.141592653589793) ** 0.5
        else:
            g = (h > 0).float()
        R = R * g
    else:
        if act == 1: R = torch.nn.functional.gelu(R)
        if act == 2: R = torch.relu(R)
    if accumulate: R = R + base
    err = (D.float() - R).abs().max().item()
    scale = R.abs().max().item()
    ok = err <= (2e-2 if not out_f32 else 2e-3) * max(scale, 1.0)
    extra = ""
    if aux_mode == 1:
        e2 = (aux.float() - pre).abs().max().item()
        ok = ok and e2 <= 2e-2 * max(pre.abs().max().item(), 1)
        extra = " aux_err=%.3g" % e2
    print("%s %s M=%d N=%d K=%d a_t=%d b_t=%d f32=%d bn=%d bias=%d act=%d acc=%d splits=%d aux=%d max_err=%.4g scale=%.3g%s" % (
        "OK  " if ok else "FAIL", tag, M, N, K, a_t, b_t, out_f32, block_n, bias, act, accumulate, splits, aux_mode, err, scale, extra))
    return ok

def bench(M, N, K, a_t, b_t, out_f32=False, accumulate=False, block_n=0, iters=20, raster=False):
    A = torch.randn((K, M) if a_t else (M, K), device=dev).to(torch.bfloat16)
    B = torch.randn((K, N) if b_t else (N, K), device=dev).to(torch.bfloat16)
    out = torch.zeros(M, N, device=dev, dtype=torch.float32 if out_f32 else torch.bfloat16)
    for _ in range(3):
        ops.gemm(A, B, out, a_t=a_t, b_t=b_t, accumulate=accumulate, block_n=block_n, raster_m_fast=raster)
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(iters):
        ops.gemm(A, B, out, a_t=a_t, b_t=b_t, accumulate=accumulate, block_n=block_n, raster_m_fast=raster)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # cuBLAS comparator (reference point only)
    Af = A.t() if a_t else A; Bf = B.t() if b_t else B
    for _ in range(3): torch.matmul(Af, Bf.t())
    torch.cuda.synchronize(); s.record()
    for _ in range(iters): torch.matmul(Af, Bf.t())
    e.record(); torch.cuda.synchronize()
    ms2 = s.elapsed_time(e) / iters
    print("BENCH M=%d N=%d K=%d a_t=%d b_t=%d f32=%d acc=%d bn=%d raster=%d: %.3f ms %.1f TF/s | cublas %.3f ms %.1f TF/s" % (
        M, N, K, a_t, b_t, out_f32, accumulate, block_n, raster, ms, tf, ms2, 2.0 * M * N * K / ms2 / 1e9))

which = sys.argv[1] if len(sys.argv) > 1 else "all"
allok = True
if which in ("all", "basic"):
    allok &= run(128, 256, 64, 0, 0, tag="1tile")
    allok &= run(128, 256, 256, 0, 0, tag="1tile-k4")
    allok &= run(128, 128, 512, 0, 0, block_n=128, tag="bn128")
    allok &= run(256, 512, 1024, 0, 0, tag="multi")
    allok &= run(1024, 1024, 1024, 0, 0, out_f32=True, tag="f32out")
    allok &= run(1000, 777 // 8 * 8, 520, 0, 0, tag="ragged")
    allok &= run(752, 1024, 2048, 0, 0, bias=True, act=2, tag="bias-relu")
    allok &= run(512, 4096, 1024, 0, 0, bias=True, act=1, aux_mode=1, tag="gelu+preact")
    allok &= run(512, 1024, 4096, 0, 0, act=1, aux_mode=2, tag="dgelu")
    allok &= run(4736, 4096, 1024, 0, 0, bias=True, tag="persist")
if which in ("all", "mn"):
    allok &= run(128, 256, 64, 0, 1, tag="B-mn 1tile")
    allok &= run(512, 1024, 1024, 0, 1, tag="dgrad")
    allok &= run(128, 256, 64, 1, 0, tag="A-mn 1tile")
    allok &= run(128, 256, 64, 1, 1, out_f32=True, tag="AB-mn 1tile")
    allok &= run(1024, 1024, 4096, 1, 1, out_f32=True, accumulate=True, tag="wgrad-auto-split")
    allok &= run(1024, 2048, 1000, 1, 1, out_f32=True, accumulate=True, splits=3, tag="wgrad-split3-ragK")
    allok &= run(4096, 1024, 2048, 1, 1, out_f32=True, accumulate=True, block_n=128, tag="wgrad-bn128")
    allok &= run(50264, 1024, 512, 1, 1, out_f32=True, accumulate=True, tag="wgrad-vocab")
    allok &= run(512, 50265 // 8 * 8, 1024, 0, 0, tag="lmhead-ish")
print("ALL_OK" if allok else "SOME_FAILED")
if which in ("all", "bench"):
    bench(18432, 1024, 1024, 0, 0)
    bench(18432, 3072, 1024, 0, 0)
    bench(18432, 4096, 1024, 0, 0)
    bench(18432, 1024, 4096, 0, 0)
    bench(18432, 1024, 4096, 0, 0, block_n=128)
    bench(18432, 4096, 1024, 0, 1)
    bench(4096, 1024, 18432, 1, 1, out_f32=True, accumulate=True)
    bench(1024, 1024, 18432, 1, 1, out_f32=True, accumulate=True)
    bench(18432, 50264, 1024, 0, 0, raster=True)
    bench(18432, 50264, 1024, 0, 0, raster=False)
    bench(8192, 8192, 8192, 0, 0)
