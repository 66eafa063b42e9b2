"""oracle/ref_hist.py -- TEST / BENCH INFRASTRUCTURE: ctypes front end of oracle/_ref/libref_hist.so, the reference's own
CUDA vote kernel (hist_cuda_core.cuh:23-99) built by oracle/Makefile from the sources under /root/reference.  Used by the
GPU tests (bit comparison with icpf_hist_votes_f32) and by bench.py's secondary C3 row (kernel-to-beat timing)."""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libref_hist.so")
_LIB = None


def available() -> bool:
    return os.path.exists(PATH)


def _lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(PATH)
        L.icpf_ref_hist_f32.restype = ctypes.c_int
        L.icpf_ref_hist_f32.argtypes = ([ctypes.c_void_p] * 2 + [ctypes.c_int] * 3 + [ctypes.c_float] * 6 + [ctypes.c_int] * 4
                                        + [ctypes.c_void_p] * 2)
        _LIB = L
    return _LIB


def hist(X: torch.Tensor, Y: torch.Tensor, mins, maxs, lens, mini_batch: int = 8, out: torch.Tensor = None) -> torch.Tensor:
    """`HIST.hist(X, Y, min..., max..., len..., mini_batch)` (hist_cuda/hist.py:39-51) on CUDA fp32 [B,N,4] tensors."""
    assert X.is_cuda and Y.is_cuda and X.dtype == torch.float32 and X.is_contiguous() and Y.is_contiguous()
    B = X.shape[0]
    if out is None:
        out = torch.empty(B, int(lens[0]), int(lens[1]), int(lens[2]), device=X.device, dtype=torch.float32)
    code = _lib().icpf_ref_hist_f32(X.data_ptr(), Y.data_ptr(), B, X.shape[1], Y.shape[1], *[float(v) for v in mins],
                                    *[float(v) for v in maxs], *[int(v) for v in lens], int(mini_batch), out.data_ptr(),
                                    torch.cuda.current_stream().cuda_stream)
    if code != 0:
        raise RuntimeError(f"reference hist kernel failed: cudaError {code}")
    return out
