"""TEST INFRASTRUCTURE: run the host-side mirror (icp_flow_b200.ops / .scan) against the SIMT-on-CPU build of the
kernel sources (tests/simt/build.py), so that the kernels' logic can be checked against the oracle without a GPU.

Nothing in the package knows about this: the fixture swaps the ctypes handle the shim uses for the emulator library,
wraps the test inputs in a Tensor subclass that answers ``is_cuda`` with True (the shim refuses CPU tensors -- the
product has no CPU path), and neutralises the two CUDA context calls the shim makes (current stream, device guard).
"""
from __future__ import annotations

import contextlib
import ctypes
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

import build as simt_build  # noqa: E402  (tests/simt/build.py)


class SimtTensor(torch.Tensor):
    """A CPU tensor the shim takes for device memory: in the emulator build "device" pointers are host pointers."""

    @property
    def is_cuda(self):  # noqa: D401
        return True

    def cpu(self, *args, **kwargs):
        """Back "on the host": a plain tensor over the same memory (what `.cpu()` of a CUDA tensor is to the tests)."""
        return self.as_subclass(torch.Tensor)


def dev_tensor(x) -> torch.Tensor:
    t = torch.as_tensor(x)
    return t.contiguous().as_subclass(SimtTensor)


def plain(t: torch.Tensor) -> torch.Tensor:
    return t.as_subclass(torch.Tensor) if isinstance(t, SimtTensor) else t


@contextlib.contextmanager
def emulated(extra_flags=(), out=None):
    """Inside the block icp_flow_b200.ops / .scan call the emulator build of the kernels."""
    from icp_flow_b200 import _lib, ops, scan

    path = simt_build.build() if out is None else simt_build.build(force=True, extra_flags=extra_flags, out=out)
    saved = (_lib._LIB, _lib.LIB_PATH, ops._stream_ptr, scan._stream_ptr, torch.cuda.device, ops._require_cuda_f32,
             scan._require_cuda_f32)

    def require_f32(t, name):
        # outputs of one emulated call (plain CPU tensors the shim allocated "on the device") feed the next one
        if not torch.is_tensor(t):
            raise TypeError(f"{name} must be a torch.Tensor")
        if t.dtype != torch.float32:
            raise TypeError(f"{name} must be float32 (got {t.dtype})")
        return t.contiguous()

    try:
        _lib._LIB, _lib.LIB_PATH = None, path
        _lib.lib()                                       # binds the argtypes of the same ABI on the emulator library
        null_stream = lambda: ctypes.c_void_p(0)         # noqa: E731
        ops._stream_ptr = null_stream
        scan._stream_ptr = null_stream
        torch.cuda.device = lambda *_a, **_k: contextlib.nullcontext()
        ops._require_cuda_f32 = require_f32
        scan._require_cuda_f32 = require_f32
        yield _lib._LIB
    finally:
        (_lib._LIB, _lib.LIB_PATH, ops._stream_ptr, scan._stream_ptr, torch.cuda.device, ops._require_cuda_f32,
         scan._require_cuda_f32) = saved
