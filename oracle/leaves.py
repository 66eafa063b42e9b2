"""oracle/leaves.py -- TEST INFRASTRUCTURE (CPU checker), never imported by the product path.

Thin ctypes front-end to the C leaves in ``oracle/knn_cpu.c`` (built by ``oracle/Makefile`` or
``__graft_entry__.build()``):

* ``knn1``      -- K=1 brute-force nearest neighbour with pytorch3d 0.7.4 ``knn_points`` semantics
                   (reference call sites: utils_icp_pytorch3d.py:154-156, utils_helper.py:27).
* ``hist_votes`` -- all-pairs difference histogram (hist_cuda/cpp/hist_cuda_core.cuh:35-62).

Both take/return torch CPU tensors so that the restatement in ``icp_oracle.py`` and the stubbed reference
run share exactly the same leaf arithmetic.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle_knn.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _HERE, "liboracle_knn.so"])
        lib = ctypes.CDLL(path)
        lib.icpf_oracle_knn1.restype = None
        lib.icpf_oracle_knn1.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int64] * 3 + [ctypes.c_void_p] * 2
        lib.icpf_oracle_hist.restype = None
        lib.icpf_oracle_hist.argtypes = (
            [ctypes.c_void_p] * 2 + [ctypes.c_int64] * 3 + [ctypes.c_float] * 6 + [ctypes.c_int64] * 3 + [ctypes.c_void_p]
        )
        _LIB = lib
    return _LIB


def knn1(p1: torch.Tensor, p2: torch.Tensor, lengths1=None, lengths2=None):
    """Squared distance and index of the nearest row of ``p2[b, :len2]`` for every ``p1[b, :len1]``.

    p1 [B,P1,3], p2 [B,P2,3] (any float dtype; computed in fp32 unless both are fp64, in which case a
    torch broadcast fallback is used).  Returns (d2 [B,P1], idx [B,P1] int64).
    """
    assert p1.dim() == 3 and p2.dim() == 3 and p1.shape[0] == p2.shape[0]
    assert p1.shape[2] == 3 and p2.shape[2] == 3
    B, P1, _ = p1.shape
    P2 = p2.shape[1]
    if p1.dtype == torch.float64:
        return _knn1_torch(p1, p2, lengths1, lengths2)
    a = p1.detach().to(torch.float32).contiguous()
    c = p2.detach().to(torch.float32).contiguous()
    d2 = torch.empty(B, P1, dtype=torch.float32)
    idx = torch.empty(B, P1, dtype=torch.int64)
    l1 = None if lengths1 is None else lengths1.detach().to(torch.int64).contiguous()
    l2 = None if lengths2 is None else lengths2.detach().to(torch.int64).contiguous()
    _lib().icpf_oracle_knn1(
        a.data_ptr(), c.data_ptr(),
        None if l1 is None else l1.data_ptr(), None if l2 is None else l2.data_ptr(),
        B, P1, P2, d2.data_ptr(), idx.data_ptr(),
    )
    return d2, idx


def _knn1_torch(p1, p2, lengths1, lengths2):
    """Broadcast fallback (used for the fp64 oracle mode and to cross-check the C leaf)."""
    B, P1, _ = p1.shape
    P2 = p2.shape[1]
    diff = p1[:, :, None, :] - p2[:, None, :, :]
    d = diff[..., 0] * diff[..., 0]
    d = d + diff[..., 1] * diff[..., 1]
    d = d + diff[..., 2] * diff[..., 2]
    if lengths2 is not None:
        bad = torch.arange(P2)[None, None, :] >= lengths2[:, None, None]
        d = d.masked_fill(bad, float("inf"))
    best, idx = d.min(dim=2)  # torch CPU min returns the first minimal index
    if lengths1 is not None:
        dead = torch.arange(P1)[None, :] >= lengths1[:, None]
        best = best.masked_fill(dead, 0.0)
        idx = idx.masked_fill(dead, 0)
    return best, idx


def hist_votes(X: torch.Tensor, Y: torch.Tensor, mins, maxs, lens) -> torch.Tensor:
    """[B,lx,ly,lz] fp32 counts of ``X[b,i,:3]-Y[b,j,:3]`` over rows whose flags are both > 0."""
    assert X.dim() == 3 and Y.dim() == 3 and X.shape[2] == 4 and Y.shape[2] == 4
    assert X.shape[0] == Y.shape[0]
    B, NX, _ = X.shape
    NY = Y.shape[1]
    x = X.detach().to(torch.float32).contiguous()
    y = Y.detach().to(torch.float32).contiguous()
    lx, ly, lz = (int(v) for v in lens)
    out = torch.empty(B, lx, ly, lz, dtype=torch.float32)
    _lib().icpf_oracle_hist(
        x.data_ptr(), y.data_ptr(), B, NX, NY,
        float(mins[0]), float(mins[1]), float(mins[2]),
        float(maxs[0]), float(maxs[1]), float(maxs[2]),
        lx, ly, lz, out.data_ptr(),
    )
    return out
