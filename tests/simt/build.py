"""TEST INFRASTRUCTURE: compile the kernel sources of icp_flow_b200/csrc for the SIMT-on-CPU emulator (simt.h).

    python tests/simt/build.py [--force]      ->  tests/simt/_build/libicpflow_simt.so

The result exports the same C ABI as the product library, but "device" pointers are host pointers and every kernel is
executed by fibers on one host thread.  Only the test-suite loads it (tests/simt/harness.py); the package never does.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "icp_flow_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
OUT = os.path.join(BUILD, "libicpflow_simt.so")

CXX = "/usr/bin/g++"
# IEEE fp32 semantics as in the CUDA intrinsics the kernels spell out: no contraction, no fast-math, SSE2 arithmetic
CXXFLAGS = ["-std=c++17", "-O2", "-g1", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-fno-strict-aliasing",
            "-Wno-unknown-pragmas", "-Wno-attributes",
            "-I", os.path.join(HERE, "shim"), "-include", os.path.join(HERE, "simt.h")]


def _deps():
    return (glob.glob(os.path.join(CSRC, "*")) + glob.glob(os.path.join(HERE, "*.h")) + glob.glob(os.path.join(HERE, "*.cpp"))
            + glob.glob(os.path.join(HERE, "shim", "*")) + [os.path.join(ROOT, "include", "icpflow_b200.h"), __file__])


def stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force: bool = False, extra_flags=(), out: str = OUT) -> str:
    if not force and out == OUT and not stale():
        return OUT
    os.makedirs(BUILD, exist_ok=True)
    tag = os.path.splitext(os.path.basename(out))[0]
    units = [(s, ["-x", "c++"]) for s in sorted(glob.glob(os.path.join(CSRC, "*.cu")))]
    units += [(os.path.join(HERE, "simt.cpp"), []), (os.path.join(HERE, "dynshared.cpp"), [])]
    objs = []

    def compile_one(unit):
        src, lang = unit
        obj = os.path.join(BUILD, f"{tag}.{os.path.basename(src)}.o")
        subprocess.check_call([CXX] + CXXFLAGS + list(extra_flags) + ["-c"] + lang + [src, "-o", obj])
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        objs = list(ex.map(compile_one, units))
    subprocess.check_call([CXX, "-shared", "-o", out] + objs)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
