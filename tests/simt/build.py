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


def stale_against(path: str) -> bool:
    if not os.path.exists(path):
        return True
    t = os.path.getmtime(path)
    return any(os.path.getmtime(d) > t for d in _deps())


def stale() -> bool:
    return stale_against(OUT)


ASAN_OUT = os.path.join(BUILD, "libicpflow_simt_asan.so")
# AddressSanitizer variant (tests/simt/memcheck.sh): every "global memory" access of a kernel is checked against the
# bounds of the host allocation behind it, and the dynamic shared-memory arrays get red zones.  Stack instrumentation is
# off because the fibers switch stacks behind ASan's back.
ASAN_FLAGS = ["-fsanitize=address", "--param", "asan-stack=0", "-fno-omit-frame-pointer"]


def build(force: bool = False, extra_flags=(), out: str = OUT) -> str:
    """Build (or reuse) the emulator library.  Serialised across processes with a file lock: the gloo tests start two
    ranks that both ask for it."""
    import fcntl

    os.makedirs(BUILD, exist_ok=True)
    with open(os.path.join(BUILD, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(force, extra_flags, out)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, extra_flags, out: str) -> str:
    if os.environ.get("ICPF_SIMT_ASAN") == "1" and out == OUT:
        out, extra_flags, force = ASAN_OUT, tuple(extra_flags) + tuple(ASAN_FLAGS), force or not os.path.exists(ASAN_OUT) or stale_against(ASAN_OUT)
    if not force and out == OUT and not stale():
        return OUT
    if not force and out != OUT and os.path.exists(out) and not stale_against(out):
        return out
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
    subprocess.check_call([CXX, "-shared", "-o", out] + [f for f in extra_flags if f.startswith("-fsanitize")] + objs)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
