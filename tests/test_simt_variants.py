"""Kernel variants that must agree bit for bit, compared under the SIMT-on-CPU emulator (tests/simt/): the product
build of a kernel against a build of the same sources with the simpler algorithm it replaces switched back on.

  ICPF_COOP_LEVELS=0      every query of the unbounded NN (csrc/icpf_gridnn.cuh) scans every row, as the reference does,
                          instead of searching blocks of grid cells first and scanning only the far queries.  The NN
                          distance is a minimum, so hist_score's scores, apply_icp's errors and match_eval's metrics
                          must not move by a bit."""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt"))
import build as simt_build  # noqa: E402
import harness  # noqa: E402
from icp_flow_b200 import ops, synth  # noqa: E402


def _path_outputs(src, dst, args):
    put = harness.dev_tensor
    pose, dbg = ops.estimate_init_pose(args, put(src), put(dst), return_debug=True)
    out, adbg = ops.apply_icp(args, put(src), put(dst), put(pose), return_debug=True, auto_swap=True)
    ev = ops.match_eval(args, put(src), put(dst), put(out))
    res = {"pose": pose, "out": out}
    res.update({"init_" + k: v for k, v in dbg.items() if torch.is_tensor(v)})
    res.update({"apply_" + k: v for k, v in adbg.items() if torch.is_tensor(v)})
    res.update({f"eval_{i}": v for i, v in enumerate(ev)})
    return {k: harness.plain(v).clone() for k, v in res.items()}


def test_grid_search_equals_the_full_scan():
    args = types.SimpleNamespace(thres_dist=0.1, translation_frame=6.666, chunk_size=50)
    batches = []
    # unrelated clusters (every query far), large motions (most candidate translations are wrong), ragged sizes
    s, d, _ = synth.make_pairs(12, 256, seed=3, ragged=True, residual_only=False, wrong_frac=0.5)
    batches.append((s, d))
    s, d, _ = synth.make_pairs(6, 700, seed=8, ragged=False, residual_only=False, wrong_frac=0.3)
    d = d.copy()
    d[0, :, :3] = np.where(d[0, :, 3:4] > 0, d[0, :, :3] + np.float32(40.0), d[0, :, :3])       # tens of metres apart
    d[1, :, 1] = np.where(d[1, :, 3] > 0, d[1, :, 1] - np.float32(7.5), d[1, :, 1])             # far along y only
    batches.append((s, d))
    variant = os.path.join(simt_build.BUILD, "libicpflow_simt_fullscan.so")
    got, want = [], []
    with harness.emulated():
        for s, d in batches:
            got.append(_path_outputs(s, d, args))
    with harness.emulated(extra_flags=("-DICPF_COOP_LEVELS=0",), out=variant):
        for s, d in batches:
            want.append(_path_outputs(s, d, args))
    for g, w in zip(got, want):
        assert g.keys() == w.keys()
        for k in g:
            a, b = g[k], w[k]
            same = torch.equal(a, b) if not a.is_floating_point() else torch.equal(a.nan_to_num(-7.0), b.nan_to_num(-7.0))
            assert same, k


def test_continued_full_pass_equals_the_restarted_one():
    """ICPF_NO_RESUME: the full pass starts the pairs still moving at the cap over (the behaviour before the loop state was
    kept).  Continuing them must not move a bit: transforms, rmse, iterations, batch stop, convergence masks."""
    src, dst, _ = synth.make_pairs(96, 1024, seed=5, ragged=False, residual_only=True, wrong_frac=0.0)
    keep = [0, 1, 2, 3, 37, 73, 87, 45, 60]            # stop at iteration 48: beyond the cap of the first pass
    src, dst = src[keep], dst[keep]

    def run():
        out = []
        for mode in (3, 2, 1):
            r = ops.icp_batch(harness.dev_tensor(src), harness.dev_tensor(dst), ops.make_params(nn_mode=mode))
            out += [harness.plain(x).clone() for x in (r.R, r.T, r.rmse, r.iterations, r.batch, r.conv_mask, r.pose)]
        return out

    with harness.emulated():
        got = run()
    with harness.emulated(extra_flags=("-DICPF_NO_RESUME",), out=os.path.join(simt_build.BUILD, "libicpflow_simt_noresume.so")):
        want = run()
    assert got[4].tolist()[0] > 32
    assert all(torch.equal(a, b) for a, b in zip(got, want))
