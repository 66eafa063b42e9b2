"""Randomized parity of the whole per-pair path against the oracle, beyond the committed goldens: seeded batches over
cluster sizes, histogram windows (translation_frame 2.0 / 3.333 / 6.666), ragged and full clouds, related and unrelated
pairs.  EVERY pair must move its points within north_star's 1e-4 m of the reference's, or be adjudicated
(oracle/adjudicate.py): the engine's transform is one of the outcomes the reference itself admits on that pair (its
fp64 run, an fp32 run on inputs moved by a few ulps, the other top-k tie orders) or the pair's Kabsch system is rank
deficient.  An unexplained pair fails the test; the explained fraction is bounded by what was observed."""
import types

import numpy as np
import pytest
import torch

from engines import put
from icp_flow_b200 import ops, synth
from oracle import icp_oracle as O

pytestmark = [pytest.mark.usefixtures("engine"), pytest.mark.order_last]

TOL = 1e-4


def test_hist_icp_on_random_batches_vs_oracle():
    from parity import assert_path_parity
    rng = np.random.default_rng(7)
    total = explained = 0
    worst = 0.0
    kinds = []
    for i in range(8):
        P, N = 6, int(rng.choice([64, 160, 300, 512]))
        F = float(rng.choice([2.0, 3.333, 6.666]))
        src, dst, _ = synth.make_pairs(P, N, seed=500 + i, ragged=bool(i % 2), residual_only=(F == 2.0), wrong_frac=0.15)
        args = types.SimpleNamespace(thres_dist=0.1, translation_frame=F, chunk_size=50)
        p = O.PathParams(thres_dist=0.1, translation_frame=F)
        s_t, d_t = torch.from_numpy(src), torch.from_numpy(dst)
        want, odbg = O.hist_icp(s_t, d_t, p, return_debug=True)
        got, dbg = ops.hist_icp(args, put(src), put(dst), return_debug=True)
        sw = odbg["swapped"]
        a_, c_ = s_t.clone(), d_t.clone()
        a_[sw] = d_t[sw]
        c_[sw] = s_t[sw]
        trace = O.icp_loop(O.transform_points_batch(a_, odbg["init"]), c_, 0.1, 100, 1e-6, diagnostics=True)
        v = assert_path_parity(src, dst, got.cpu(), want, p, trace.iterations, dbg["batch"].tolist()[0], max_explained=0.5,
                               what=f"random batch {i} (N={N}, F={F})", trace=trace)
        total += P
        explained += int(v.explained.sum())
        kinds += [x for x in v.verdict if x != "ok"]
        ok = np.array([x == "ok" for x in v.verdict])
        worst = max(worst, float(v.err[ok].max()) if ok.any() else 0.0)
    assert explained <= 0.1 * total, (explained, total, kinds)
    print(f"random parity: {total} pairs, {explained} adjudicated {kinds}, worst error of the others {worst:.2e} m")
