"""Where the parity tests execute the kernels.

  cuda  the product: libicpflow_b200.so on cuda:0 (tests marked ``gpu``; the parity tests proper)
  simt  TEST INFRASTRUCTURE: the same kernel sources compiled for the SIMT-on-CPU emulator (tests/simt/), so that the
        kernels' logic is checked against the oracle in the build container too (``-m "not gpu"``).  Approximate device
        functions (rsqrtf, atan2f ...) are the host's there, so bit-exact claims between kernel variants still hold
        (both sides run the same arithmetic) while comparisons with the oracle keep the stated tolerances.

Test modules call ``put(x)`` instead of ``x.to("cuda:0")`` and ``sync()`` instead of ``torch.cuda.synchronize()``;
the ``engine`` fixture (conftest.py) selects what they mean.
"""
from __future__ import annotations

import contextlib
import os
import sys

import torch

_ACTIVE = None


class CudaEngine:
    name = "cuda"

    def __init__(self):
        assert torch.cuda.is_available(), "the gpu-marked tests need a CUDA device"
        self.device = torch.device("cuda:0")

    def put(self, x):
        return torch.as_tensor(x).to(self.device)

    def sync(self):
        torch.cuda.synchronize()


class SimtEngine:
    name = "simt"

    def __init__(self):
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt"))
        import harness  # tests/simt/harness.py

        self._harness = harness
        self.device = torch.device("cpu")

    def put(self, x):
        return self._harness.dev_tensor(x)

    def sync(self):
        pass


@contextlib.contextmanager
def running(name: str):
    """Activate the engine ``name`` ("cuda" | "simt") for the duration of a test."""
    if name == "cuda":
        e = CudaEngine()
        activate(e)
        try:
            yield e
        finally:
            activate(None)
    else:
        e = SimtEngine()
        with e._harness.emulated():
            activate(e)
            try:
                yield e
            finally:
                activate(None)


def activate(engine):
    global _ACTIVE
    _ACTIVE = engine


def active():
    assert _ACTIVE is not None, "use the `engine` fixture"
    return _ACTIVE


def put(x):
    return active().put(x)


def sync():
    active().sync()


def device():
    return active().device


def is_simt() -> bool:
    return active().name == "simt"
