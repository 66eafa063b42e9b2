#!/usr/bin/env python
"""Per-iteration cache statistics of the ICP loop on a C2 sample, counted by the kernel sources under the SIMT emulator
(build flag -DICPF_ITER_STATS; CPU only, no GPU needed)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "simt"))
import numpy as np, torch
import harness, build as simt_build
from icp_flow_b200 import ops, synth
P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
out = os.path.join(simt_build.BUILD, "libicpflow_simt_stats.so")
with harness.emulated(extra_flags=("-DICPF_ITER_STATS",), out=out) as L:
    s, d, _ = synth.make_pairs(P, 512, seed=1234, ragged=False, residual_only=True)
    prm = ops.make_params(thres=0.1, max_iterations=20, relative_rmse_thr=-1.0, early_exit=False, batch_stop=True)
    ops.icp_batch(harness.dev_tensor(s), harness.dev_tensor(d), prm)
    st = np.ctypeslib.as_array((ctypes.c_longlong * 512).in_dll(L, "icpf_dbg_iter_stats")).reshape(128, 4)
    print("iter  reval/pair  search/pair  warps_summed/4  (pairs)")
    for it in range(20):
        n = max(st[it, 3], 1)
        print(f"{it:3d}  {st[it,0]/n:9.1f}  {st[it,1]/n:9.1f}  {st[it,2]/n:9.2f}   {st[it,3]}")
    print("mean", st[:20, 0].sum() / st[:20, 3].sum(), st[:20, 1].sum() / st[:20, 3].sum(), st[:20, 2].sum() / st[:20, 3].sum())
