"""Parity assertions shared by the test modules: EVERY pair is compared with the fp32 oracle at the stated tolerance;
a pair outside it has to be adjudicated (oracle/adjudicate.py: the engine's transform must be one of the outcomes the
reference admits -- its fp64 run or an fp32 run on inputs moved by a few ulps); an unexplained pair fails the test and
the fraction of explained pairs is bounded by what was observed on the B200 / under the emulator plus a small margin."""
import numpy as np
import torch

from oracle import adjudicate as A

TOL = 1e-4


def _rigid(R):
    Rn = np.asarray(R, dtype=np.float64)
    assert np.isfinite(Rn).all()
    assert np.abs(Rn @ Rn.transpose(0, 2, 1) - np.eye(3)).max() < 1e-5
    assert np.linalg.det(Rn).min() > 0.999


def assert_icp_parity(src, dst, R, T, its_eng, R_ref, T_ref, its_ref, thres=0.1, max_explained=0.1, what="icp",
                      trace=None):
    """ICP loop outputs (row convention R, T) against the oracle's.  `trace`: the oracle run with diagnostics (lets a pair
    with a rank <= 1 Kabsch system be recognised as such).  Returns the Verdicts."""
    _rigid(R)
    assert np.isfinite(np.asarray(T)).all()
    rd = A.rank_deficient_pairs(trace) if trace is not None and trace.min_inliers is not None else None
    v = A.adjudicate_icp(torch.as_tensor(src), torch.as_tensor(dst), torch.as_tensor(R), torch.as_tensor(T),
                         torch.as_tensor(R_ref), torch.as_tensor(T_ref), its_ref, its_eng, thres=thres, tol=TOL,
                         rank_deficient=rd)
    print(f"parity[{what}]: {v.summary()}")
    assert not v.unexplained.any(), (np.nonzero(v.unexplained)[0], v.err[v.unexplained])
    assert v.explained.mean() <= max_explained, (v.explained.mean(), [x for x in v.verdict if x != "ok"])
    return v


def assert_path_parity(src, dst, T_eng, T_ref, p, its_ref, its_eng=None, stage="hist_icp", init=None,
                       max_explained=0.1, what="hist_icp", trace=None):
    """hist_icp / apply_icp outputs (column-convention 4x4) against the oracle's.  Returns the Verdicts."""
    T_eng = torch.as_tensor(T_eng)
    _rigid(T_eng[:, :3, :3].transpose(1, 2))
    rd = A.rank_deficient_pairs(trace) if trace is not None and trace.min_inliers is not None else None
    v = A.adjudicate_path(torch.as_tensor(src), torch.as_tensor(dst), T_eng, torch.as_tensor(T_ref), p, its_ref, its_eng,
                          stage=stage, init=None if init is None else torch.as_tensor(init), tol=TOL, rank_deficient=rd)
    print(f"parity[{what}]: {v.summary()}")
    assert not v.unexplained.any(), (np.nonzero(v.unexplained)[0], v.err[v.unexplained])
    assert v.explained.mean() <= max_explained, (v.explained.mean(), [x for x in v.verdict if x != "ok"])
    return v


def selected_pairs_parity(stage_batches, p, ref_rows, T_eng, T_ref, max_explained=0.1, what="frame", engine_its=None):
    """Frame level: `ref_rows[:, :2]` are the selected (src label, dst label) pairs, `T_eng` / `T_ref` their transforms;
    `stage_batches` = [(segs_src [K,N,4], segs_dst, pairs [K,2]), ...] the padded candidate batches hist_icp saw (one per
    match_pairs call: the batch stop couples the pairs of a call).  Every selected pair is held to the tolerance or
    adjudicated on its own padded clouds.  `engine_its(segs_src, segs_dst) -> int` (optional): the batch iterations the
    ENGINE executes on such a candidate batch -- the batch stop is itself a threshold decision, so the admitted outcomes
    are taken at the oracle's and at the engine's count.  Returns the mask of the pairs that needed adjudication."""
    from oracle import icp_oracle as O
    flagged = np.zeros(len(ref_rows), dtype=bool)
    seen = np.zeros(len(ref_rows), dtype=bool)
    for segs_src, segs_dst, pairs in stage_batches:
        key = {(float(a), float(b)): k for k, (a, b) in enumerate(np.asarray(pairs))}
        rows = [r for r, (a, b) in enumerate(ref_rows[:, :2]) if (float(a), float(b)) in key and not seen[r]]
        if not rows:
            continue
        sel = np.array([key[(float(ref_rows[r, 0]), float(ref_rows[r, 1]))] for r in rows])
        _, odbg = O.hist_icp(segs_src, segs_dst, p, return_debug=True)
        sw = odbg["swapped"]
        a_, c_ = segs_src.clone(), segs_dst.clone()
        a_[sw] = segs_dst[sw]
        c_[sw] = segs_src[sw]
        trace = O.icp_loop(O.transform_points_batch(a_, odbg["init"]), c_, p.thres_dist, p.max_iterations,
                           p.relative_rmse_thr, diagnostics=True)
        sub = O.IcpTrace(*[None] * 10)._replace(min_inliers=trace.min_inliers[sel], min_sigma_ratio=trace.min_sigma_ratio[sel])
        its_eng = int(engine_its(segs_src, segs_dst)) if engine_its is not None else None
        v = assert_path_parity(segs_src[sel], segs_dst[sel], torch.as_tensor(T_eng[rows]), torch.as_tensor(T_ref[rows]), p,
                               trace.iterations, its_eng, max_explained=1.0, what=f"{what}, {len(rows)} selected pairs",
                               trace=sub)
        flagged[rows] = v.explained
        seen[rows] = True
    assert seen.all()
    assert flagged.mean() <= max_explained, flagged.mean()
    return flagged
