"""Thin torch-tensor -> C-ABI wrappers.  torch is used for device memory and streams only."""
import ctypes as C

import torch

from . import _lib
from ._lib import GemmArgs, check

ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
AUX_NONE, AUX_STORE_PREACT, AUX_MUL_DACT = 0, 1, 2


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _check2d(t, name, dtype=None):
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError("%s must be 2-D with unit inner stride, got %s / %s" % (name, tuple(t.shape), t.stride()))
    if dtype is not None and t.dtype != dtype:
        raise ValueError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (no CPU fallback)" % name)


def gemm(A, B, out=None, *, a_t=False, b_t=False, out_dtype=torch.bfloat16, bias=None, act=ACT_NONE,
         aux=None, aux_mode=AUX_NONE, alpha=1.0, accumulate=False, splits=0, block_n=0, raster_m_fast=False):
    """out[M,N] (+)= epi(alpha * A·Bᵀ).

    A: [M,K] (a_t=False) or [K,M] (a_t=True, read MN-major, no copy); B: [N,K] or [K,N] (b_t=True).
    """
    _check2d(A, "A", torch.bfloat16)
    _check2d(B, "B", torch.bfloat16)
    if a_t:
        K, M = A.shape
    else:
        M, K = A.shape
    if b_t:
        Kb, N = B.shape
    else:
        N, Kb = B.shape
    if K != Kb:
        raise ValueError("contraction mismatch: %d vs %d" % (K, Kb))
    if out is None:
        out = torch.empty((M, N), device=A.device, dtype=out_dtype)
        if accumulate:
            out.zero_()
    _check2d(out, "out")
    if tuple(out.shape) != (M, N):
        raise ValueError("out shape %s != (%d,%d)" % (tuple(out.shape), M, N))
    if out.dtype not in (torch.bfloat16, torch.float32):
        raise ValueError("out must be bf16 or fp32")
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != N or not bias.is_contiguous()):
        raise ValueError("bias must be contiguous fp32 [N]")
    if aux_mode != AUX_NONE:
        _check2d(aux, "aux", torch.bfloat16)
    a = GemmArgs()
    a.A, a.B, a.D = A.data_ptr(), B.data_ptr(), out.data_ptr()
    a.lda, a.ldb, a.ldd = A.stride(0), B.stride(0), out.stride(0)
    a.M, a.N, a.K = M, N, K
    a.a_mn_major, a.b_mn_major = int(a_t), int(b_t)
    a.out_f32 = int(out.dtype == torch.float32)
    a.accumulate = int(accumulate)
    a.splits, a.block_n, a.raster_m_fast = splits, block_n, int(raster_m_fast)
    a.alpha = alpha
    a.bias = bias.data_ptr() if bias is not None else None
    a.act, a.aux_mode = act, aux_mode
    a.aux = aux.data_ptr() if aux is not None else None
    a.ld_aux = aux.stride(0) if aux is not None else 0
    check(_lib.lib().mmsum_gemm_bf16(C.byref(a), _stream()), "mmsum_gemm_bf16")
    return out
