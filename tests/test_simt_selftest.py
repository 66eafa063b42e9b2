"""Known-answer checks of the SIMT-on-CPU emulator itself (tests/simt/selftest.cpp): the parity tests that run through
it are only as good as its warp collectives, barriers and shared-memory model."""
import ctypes
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SIMT = os.path.join(HERE, "simt")
sys.path.insert(0, SIMT)
import build as simt_build  # noqa: E402


def test_emulator_known_answers():
    os.makedirs(simt_build.BUILD, exist_ok=True)
    out = os.path.join(simt_build.BUILD, "libsimt_selftest.so")
    flags = [f for f in simt_build.CXXFLAGS if f != "-include" and not f.endswith("simt.h")]
    subprocess.check_call([simt_build.CXX] + flags + ["-shared", "-o", out, os.path.join(SIMT, "selftest.cpp"),
                                                     os.path.join(SIMT, "simt.cpp"), os.path.join(SIMT, "dynshared.cpp")])
    # both hand-over orders, each in a fresh process (the order is read when the library is loaded)
    for order in ("forward", "reverse"):
        code = f"import ctypes; L = ctypes.CDLL({out!r}); L.simt_selftest.restype = ctypes.c_int; raise SystemExit(L.simt_selftest())"
        env = dict(os.environ, SIMT_ORDER=order)
        assert subprocess.run([sys.executable, "-c", code], env=env).returncode == 0, order
