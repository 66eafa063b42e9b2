"""Randomized parity of the whole per-pair path against the oracle, beyond the committed goldens: seeded batches over
cluster sizes, histogram windows (translation_frame 2.0 / 3.333 / 6.666), ragged and full clouds, related and unrelated
pairs.  Every pair the oracle marks numerically determined (oracle.undetermined_pairs: top-k ties, roll-back ties,
correspondences at the gate, rank-deficient Kabsch systems) must move its points within north_star's 1e-4 m of the
reference's; the flagged fraction is bounded so the test cannot pass vacuously.  A pair outside the tolerance that only
the wide early-flip criterion explains (oracle.unstable_pairs, `early_ulps`) is held to the quality of its registration
instead, and such pairs are bounded too."""
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
    rng = np.random.default_rng(7)
    total = flagged = early_flips = 0
    worst = 0.0
    for i in range(8):
        P, N = 6, int(rng.choice([64, 160, 300, 512]))
        F = float(rng.choice([2.0, 3.333, 6.666]))
        src, dst, _ = synth.make_pairs(P, N, seed=500 + i, ragged=bool(i % 2), residual_only=(F == 2.0), wrong_frac=0.15)
        args = types.SimpleNamespace(thres_dist=0.1, translation_frame=F, chunk_size=50)
        p = O.PathParams(thres_dist=0.1, translation_frame=F)
        want = O.hist_icp(torch.from_numpy(src), torch.from_numpy(dst), p)
        got = ops.hist_icp(args, put(src), put(dst)).cpu()
        s_t, d_t = torch.from_numpy(src), torch.from_numpy(dst)
        skip = O.ambiguous_topk_rows(s_t, d_t, p).numpy() | O.undetermined_pairs(s_t, d_t, p).numpy()
        flip_prone = None
        n_s = (src[:, :, 3] > 0).sum(1)
        for k in range(P):
            total += 1
            if skip[k]:
                flagged += 1
                continue
            pts = torch.from_numpy(src[k, : n_s[k], :3]).double()
            a = pts @ got[k, :3, :3].double().T + got[k, :3, 3].double()
            b = pts @ want[k, :3, :3].double().T + want[k, :3, 3].double()
            err = float((a - b).abs().max())
            if err > TOL:
                # Outside the tolerance on a pair the strict diagnosis does not flag: admissible only as an EARLY flip --
                # some correspondence sat within 2.5 fp32 ulps of the gate in some iteration (which side it falls on is a
                # matter of the last bits of R and T) -- and only if the engine's registration is as good as the
                # reference's: its mean NN error must not exceed the reference's by more than fp32 noise.
                if flip_prone is None:
                    flip_prone = O.undetermined_pairs(s_t, d_t, p, early_ulps=2.5).numpy()
                ev_g = O.match_eval(s_t[k:k + 1], d_t[k:k + 1], got[k:k + 1], p)[0][0, 0]
                ev_w = O.match_eval(s_t[k:k + 1], d_t[k:k + 1], want[k:k + 1], p)[0][0, 0]
                assert flip_prone[k] and float(ev_g) <= float(ev_w) * 1.02 + 1e-5, (i, k, N, F, err, float(ev_g), float(ev_w))
                early_flips += 1
                continue
            worst = max(worst, err)
    assert flagged <= 0.5 * total, (flagged, total)
    assert early_flips <= 0.1 * total, early_flips
    print(f"random parity: {total} pairs, {flagged} flagged undetermined, {early_flips} early flips held to the "
          f"registration quality, worst determined error {worst:.2e} m")
