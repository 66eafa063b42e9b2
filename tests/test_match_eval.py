"""match_eval (utils_match.py:159-213): oracle pinned bitwise to the reference goldens on CPU;
engine (one fused launch through icpf_match_eval_f32) against goldens, oracle and analytic answers on the GPU.

Tolerances: inlier counts are integer work -> exact on every row whose NN distance is not within 1e-6 m of the gate
(the distance itself differs from torch CPU in the last ulp: different sqrt rounding); errors / translations are fp32 sums
in a different order -> 1e-5 m (well inside north_star's 1e-4); angles 1e-4 degrees.
"""
import types

import numpy as np
import pytest
import torch

from engines import put, sync
from oracle import icp_oracle as O

NAMES = ("errors", "inliers", "ratios", "ious", "translations", "rotations")


def _params(g):
    return O.PathParams(thres_dist=float(g["thres_dist"]), translation_frame=float(g["translation_frame"]))


@pytest.mark.parametrize("name", ["c1_demo.npz", "synth_hist.npz", pytest.param("synth_hist_default.npz", marks=pytest.mark.order_last)])
def test_oracle_match_eval_bitwise(golden, name):
    g = golden(name)
    ev = O.match_eval(torch.from_numpy(g["src"]), torch.from_numpy(g["dst"]), torch.from_numpy(g["T_hist_icp"]), _params(g))
    for n, v in zip(NAMES, ev):
        assert np.array_equal(v.numpy(), g["eval_" + n], equal_nan=True), n


def test_oracle_euler_known_answers():
    """Rz(30 deg) Ry(-10 deg) Rx(5 deg) decodes to (30, -10, 5) in the reference's ZYX convention."""
    def rot(axis, deg):
        a = np.deg2rad(deg)
        c, s = np.cos(a), np.sin(a)
        return {"x": np.array([[1, 0, 0], [0, c, -s], [0, s, c]]), "y": np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]]),
                "z": np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])}[axis]
    R = rot("z", 30.0) @ rot("y", -10.0) @ rot("x", 5.0)
    ang = O.euler_zyx_degrees(torch.from_numpy(R[None].astype(np.float32)))
    assert np.abs(ang.numpy()[0] - np.array([30.0, -10.0, 5.0])).max() < 1e-4


def _near_gate_pairs(src, dst, T, p, margin=1e-6):
    """Pairs with a valid row whose NN distance is within `margin` of the inlier gate (count not determined)."""
    moved = O.transform_points_batch(src, T)
    _, e12 = O.nearest_neighbor_batch(moved, dst)
    _, e21 = O.nearest_neighbor_batch(dst, moved)
    m1, m2 = src[:, :, 3] > 0, dst[:, :, 3] > 0
    near = (((e12 - p.thres_dist).abs() < margin) & m1).any(1) | (((e21 - p.thres_dist).abs() < margin) & m2).any(1)
    return near.numpy()


def _engine(src, dst, T, thres):
    from icp_flow_b200 import ops
    args = types.SimpleNamespace(thres_dist=thres)
    out = ops.match_eval(args, put(src), put(dst), put(T))
    sync()
    return [o.cpu().numpy() for o in out]


def _compare(ev, ref, det):
    errors, inliers, ratios, ious, trans, rots = ev
    assert np.array_equal(inliers[det], ref["inliers"][det])
    np.testing.assert_allclose(errors, ref["errors"], atol=1e-5, rtol=1e-5, equal_nan=True)
    np.testing.assert_allclose(ratios[det], ref["ratios"][det], rtol=1e-6, atol=0, equal_nan=True)
    np.testing.assert_allclose(ious[det], ref["ious"][det], rtol=1e-6, atol=0, equal_nan=True)
    np.testing.assert_allclose(trans, ref["translations"], atol=1e-5, equal_nan=True)
    np.testing.assert_allclose(rots, ref["rotations"], atol=1e-4, equal_nan=True)


@pytest.mark.usefixtures("engine")
@pytest.mark.parametrize("name", ["c1_demo.npz", "synth_hist.npz", pytest.param("synth_hist_default.npz", marks=pytest.mark.order_last)])
def test_engine_match_eval_vs_reference_golden(golden, name):
    g = golden(name)
    src, dst, T = torch.from_numpy(g["src"]), torch.from_numpy(g["dst"]), torch.from_numpy(g["T_hist_icp"])
    det = ~_near_gate_pairs(src, dst, T, _params(g))
    assert det.mean() > 0.9
    _compare(_engine(src, dst, T, float(g["thres_dist"])), {n: g["eval_" + n] for n in NAMES}, det)


@pytest.mark.usefixtures("engine")
@pytest.mark.parametrize("P,N", [(64, 512), (3, 9000)])
def test_engine_match_eval_vs_oracle_ragged_and_large(P, N):
    """Ragged batch incl. an EMPTY src cloud (0/0 -> NaN like the reference) and clusters beyond one staged tile."""
    from icp_flow_b200 import synth
    src, dst, meta = synth.make_pairs(P, N, seed=5 + N, ragged=True, residual_only=True)
    src[1, :, :3], src[1, :, 3] = 1e8, 0.0
    src_t, dst_t = torch.from_numpy(src), torch.from_numpy(dst)
    ang = torch.linspace(-0.05, 0.05, P)
    T = torch.eye(4).repeat(P, 1, 1)
    T[:, 0, 0], T[:, 0, 1], T[:, 1, 0], T[:, 1, 1] = ang.cos(), -ang.sin(), ang.sin(), ang.cos()
    T[:, :3, 3] = torch.from_numpy(meta["translation"]).float() * 0.0 + torch.tensor([0.02, -0.01, 0.0])
    p = O.PathParams(thres_dist=0.1)
    ref = dict(zip(NAMES, [v.numpy() for v in O.match_eval(src_t, dst_t, T, p)]))
    det = ~_near_gate_pairs(src_t, dst_t, T, p)
    ev = _engine(src_t, dst_t, T, 0.1)
    assert np.isnan(ev[0][1, 0]) and np.isnan(ref["errors"][1, 0]) and ev[1][1, 0] == 0
    _compare(ev, ref, det)


@pytest.mark.usefixtures("engine")
def test_engine_match_eval_known_answers():
    """dst = src + shift and T = that translation: zero error, every valid row an inlier, ratio 1, IoU n/(2n-n) = 1,
    translation = shift, angles 0; with T = identity and the shift > gate: no inliers."""
    rng = np.random.default_rng(0)
    n, N = 300, 512
    pts = (rng.uniform(-1, 1, size=(n, 3)) * np.array([2.0, 1.0, 0.8]) + np.array([15.0, -7.0, 0.5])).astype(np.float32)
    shift = np.array([0.5, -0.25, 0.125], np.float32)        # exactly representable: moved rows == dst rows bit for bit
    src = np.full((2, N, 4), 1e8, np.float32)
    src[:, :, 3] = 0
    dst = src.copy()
    for k in range(2):
        src[k, :n, :3], src[k, :n, 3] = pts, 1
        dst[k, :n, :3], dst[k, :n, 3] = pts + shift, 1
    T = torch.eye(4).repeat(2, 1, 1)
    T[0, :3, 3] = torch.from_numpy(shift)
    errors, inliers, ratios, ious, trans, rots = _engine(torch.from_numpy(src), torch.from_numpy(dst), T, 0.1)
    assert np.array_equal(errors[0], [0, 0]) and np.array_equal(inliers[0], [n, n])
    assert np.array_equal(ratios[0], [1, 1]) and np.array_equal(ious[0], [1, 1])
    assert np.abs(trans[0] - shift).max() < 1e-5 and np.array_equal(rots, np.zeros((2, 3), np.float32))
    assert inliers[1].max() < n * 0.2 and np.abs(trans[1]).max() == 0.0


@pytest.mark.usefixtures("engine")
def test_engine_match_eval_argument_errors():
    from icp_flow_b200 import _lib, ops
    a = put(torch.zeros(2, 16, 4))
    with pytest.raises(ValueError):
        ops.match_eval(types.SimpleNamespace(thres_dist=0.1), a, a, put(torch.eye(4)).repeat(3, 1, 1))
    with pytest.raises(_lib.IcpfError):
        ops.match_eval(types.SimpleNamespace(thres_dist=0.0), a, a, put(torch.eye(4)).repeat(2, 1, 1))


# ------------------------------------------------------------------------------------------ match_pairs (gates + selection)
def _gates(g):
    return O.MatchGates(max_points=int(g["max_points"]), thres_error=float(g["thres_error"]),
                        thres_iou=float(g["thres_iou"]), thres_rot=float(g["thres_rot"]))


def _mp_inputs(golden, name):
    g = golden(name)
    if "src_points" in g:
        return g, g["src_points"], g["src_labels"], g["dst_points"], g["dst_labels"], g["pairs"]
    from oracle.gen_golden import clouds_from_batches
    pairs = g["pair_labels"].astype(np.float32)
    return (g, *clouds_from_batches(g["src"], g["dst"], pairs), pairs)


@pytest.mark.parametrize("name", ["c1_demo.npz", "synth_match_dyn.npz"])
def test_oracle_match_pairs_bitwise(golden, name):
    g, sp, sl, dp, dl, pairs = _mp_inputs(golden, name)
    rows, T = O.match_pairs(*(torch.from_numpy(x) for x in (sp, dp, sl, dl, pairs)), _params(g), _gates(g))
    assert np.array_equal(rows.numpy(), g["mp_rows"]) and np.array_equal(T.numpy(), g["mp_T"])


@pytest.mark.usefixtures("engine")
@pytest.mark.parametrize("name", ["c1_demo.npz", "synth_match_dyn.npz"])
def test_match_select_equals_oracle_loop(golden, name):
    """icpf_match_select_f32 (one launch: 64-bit atomic arg-min per src cluster + ordered compaction) reproduces the
    reference's per-pair scatter loop and match_segments_descend bit for bit when fed the same metrics, and its unmatched
    label lists are what match_pcds derives with isin (utils_match.py:43-47)."""
    from icp_flow_b200 import ops
    g, sp, sl, dp, dl, pairs = _mp_inputs(golden, name)
    p, gates = _params(g), _gates(g)
    tens = [torch.from_numpy(x) for x in (sp, dp, sl, dl, pairs)]
    rows, T, dbg = O.match_pairs(*tens, p, gates, return_debug=True)
    ev = dbg["evals"]
    accept = torch.tensor([O.check_transformation(ev[4][k], ev[5][k], min(ev[3][k]), p, gates) for k in range(len(pairs))],
                          dtype=torch.int32)
    assert 0 < int(accept.sum()) < len(pairs) or name == "c1_demo.npz"
    args = types.SimpleNamespace(thres_error=gates.thres_error)
    su, du = torch.unique(tens[2]).long(), torch.unique(tens[3]).long()
    dev_ev = [put(x.contiguous()) for x in ev[0:4]]
    rows2, T2, s_left, d_left = ops.match_select(args, put(tens[4]), put(su), put(du), dev_ev, put(accept), put(dbg["T"]),
                                                 return_left=True)
    assert np.array_equal(rows2.cpu().numpy(), rows.numpy()) and np.array_equal(T2.cpu().numpy(), T.numpy())
    assert torch.equal(s_left.cpu(), su[~torch.isin(su, rows[:, 0].long())])
    assert torch.equal(d_left.cpu(), du[~torch.isin(du, rows[:, 1].long())])
    # nothing accepted -> empty [0,10] / [0,4,4] like the reference's else-branch, every label left
    rows0, T0, s0, d0 = ops.match_select(args, put(tens[4]), put(su), put(du), dev_ev, put(accept * 0), put(dbg["T"]),
                                         return_left=True)
    assert rows0.shape == (0, 10) and T0.shape == (0, 4, 4) and torch.equal(s0.cpu(), su) and torch.equal(d0.cpu(), du)


@pytest.mark.usefixtures("engine")
def test_match_select_ties_nan_and_threshold():
    """Arg-min semantics of the reference's matrices: equal errors -> the lowest dst position; a NaN error wins the arg-min
    (torch) and then fails the threshold, dropping its src cluster; rejected pairs and errors at the threshold are out."""
    from icp_flow_b200 import ops
    su, du = torch.tensor([3, 5, 9, 12]), torch.tensor([1, 4, 7])
    pairs = torch.tensor([[3, 7], [3, 4], [5, 1], [5, 4], [9, 1], [9, 7], [12, 4], [12, 1]])
    err = torch.tensor([[0.05, 0.09], [0.07, 0.05], [float("nan"), 0.01], [0.02, 0.03], [0.2, 0.5], [0.01, 0.3], [0.04, 0.04],
                        [0.03, 0.5]])
    accept = torch.tensor([1, 1, 1, 1, 1, 0, 1, 1], dtype=torch.int32)
    other = [torch.arange(16, dtype=torch.float32).reshape(8, 2) + k for k in (100, 200, 300)]
    T = torch.arange(8 * 16, dtype=torch.float32).reshape(8, 4, 4)
    args = types.SimpleNamespace(thres_error=0.2)
    rows, Tm, s_left, d_left = ops.match_select(args, put(pairs), put(su), put(du), [put(err)] + [put(x) for x in other],
                                                put(accept), put(T), return_left=True)
    rows, Tm = rows.cpu(), Tm.cpu()
    # 3: both pairs have min 0.05 -> dst 4 (position 1) before dst 7 (position 2);  5: NaN row dropped;
    # 9: its accepted pair has min error 0.2 (not < 0.2), the better one is rejected;  12: dst 1 (0.03) beats dst 4 (0.04)
    assert rows[:, 0].tolist() == [3.0, 12.0] and rows[:, 1].tolist() == [4.0, 1.0]
    assert torch.equal(rows[0, 2:4], err[1]) and torch.equal(rows[1, 2:4], err[7])
    assert torch.equal(rows[0, 4:6], other[0][1]) and torch.equal(rows[1, 8:10], other[2][7])
    assert torch.equal(Tm[0], T[1]) and torch.equal(Tm[1], T[7])
    assert s_left.cpu().tolist() == [5, 9] and d_left.cpu().tolist() == [7]
    # no pairs at all / no labels on one side: nothing selected, every label left
    none = [put(torch.zeros(0, 2)) for _ in range(4)]
    r0, T0, s0, d0 = ops.match_select(args, put(torch.zeros(0, 2, dtype=torch.int64)), put(su), put(du), none,
                                      put(torch.zeros(0, dtype=torch.int32)), put(torch.zeros(0, 4, 4)), return_left=True)
    assert r0.shape == (0, 10) and T0.shape == (0, 4, 4) and s0.cpu().tolist() == su.tolist() and d0.cpu().tolist() == du.tolist()
    r1, T1, s1, d1 = ops.match_select(args, put(pairs), put(su[:0]), put(du), [put(err)] + [put(x) for x in other],
                                      put(accept), put(T), return_left=True)
    assert r1.shape == (0, 10) and len(s1) == 0 and d1.cpu().tolist() == du.tolist()
    # a pair whose label is not in the lists is ignored (the reference's torch.nonzero finds no row for it)
    r2, _ = ops.match_select(args, put(torch.tensor([[3, 7], [4, 4]])), put(su), put(du), [put(err[:2])] + [put(x[:2]) for x in other],
                             put(accept[:2]), put(T[:2]))
    assert r2.cpu()[:, 0:2].tolist() == [[3.0, 7.0]]


@pytest.mark.usefixtures("engine")
@pytest.mark.parametrize("name", ["c1_demo.npz", "synth_match_dyn.npz"])
def test_engine_check_transformation_flags(golden, name):
    """Fused accept flag == check_transformation on the oracle's metrics (pairs near a gate excluded)."""
    from icp_flow_b200 import ops
    g, sp, sl, dp, dl, pairs = _mp_inputs(golden, name)
    p, gates = _params(g), _gates(g)
    _, _, dbg = O.match_pairs(*(torch.from_numpy(x) for x in (sp, dp, sl, dl, pairs)), p, gates, return_debug=True)
    ev = dbg["evals"]
    want = np.array([O.check_transformation(ev[4][k], ev[5][k], min(ev[3][k]), p, gates) for k in range(len(pairs))])
    args = types.SimpleNamespace(thres_dist=p.thres_dist, translation_frame=p.translation_frame,
                                 thres_iou=gates.thres_iou, thres_rot=gates.thres_rot)
    out = ops.match_eval(args, put(dbg["segs_src"]), put(dbg["segs_dst"]), put(dbg["T"]), return_accept=True)
    got = out[6].cpu().numpy().astype(bool)
    tn = torch.linalg.norm(ev[4], dim=1).numpy()
    near = (np.abs(tn - p.translation_frame) < 1e-4) | (np.abs(ev[3].min(1)[0].numpy() - gates.thres_iou) < 1e-3) | \
           (np.abs(ev[5][:, 1:3].abs().max(1)[0].numpy() - gates.thres_rot * 90) < 1e-3)
    det = ~near & ~_near_gate_pairs(dbg["segs_src"], dbg["segs_dst"], dbg["T"], p)
    assert det.mean() > 0.8 and np.array_equal(got[det], want[det])


@pytest.mark.usefixtures("engine")
@pytest.mark.parametrize("name", ["c1_demo.npz", "synth_match_dyn.npz"])
def test_engine_match_pairs_vs_reference_golden(golden, name):
    """Whole match_pairs on the engine: same selected (src, dst) label pairs as the reference, metrics within tolerance
    and transforms moving the cluster points within 1e-4 m on the numerically determined pairs."""
    from icp_flow_b200 import ops
    g, sp, sl, dp, dl, pairs = _mp_inputs(golden, name)
    p, gates = _params(g), _gates(g)
    args = types.SimpleNamespace(thres_dist=p.thres_dist, translation_frame=p.translation_frame, chunk_size=p.chunk_size,
                                 max_points=gates.max_points, thres_error=gates.thres_error, thres_iou=gates.thres_iou,
                                 thres_rot=gates.thres_rot)
    import icp_flow_b200
    rows, T = icp_flow_b200.match_pairs(args, *(put(torch.from_numpy(x)) for x in (sp, dp, sl, dl, pairs)))
    rows, T = rows.cpu().numpy(), T.cpu().numpy()
    ref_rows, ref_T = g["mp_rows"], g["mp_T"]
    assert rows.shape[1] == 10
    assert np.array_equal(rows[:, :2], ref_rows[:, :2]), (rows[:, :2], ref_rows[:, :2])
    # every selected pair within 1e-4 m of the reference's transform, or adjudicated (tests/parity.py, oracle/adjudicate.py)
    from parity import selected_pairs_parity
    _, _, dbg = O.match_pairs(*(torch.from_numpy(x) for x in (sp, dp, sl, dl, pairs)), p, gates, return_debug=True)
    flagged = selected_pairs_parity([(dbg["segs_src"], dbg["segs_dst"], pairs)], p, ref_rows, T, ref_T, max_explained=0.1,
                                    what=name)
    det = ~flagged
    np.testing.assert_allclose(rows[det, 2:4], ref_rows[det, 2:4], atol=2e-5)        # mean NN errors (m)
    exact = det & (np.abs(rows[:, 4:6] - ref_rows[:, 4:6]).max(1) == 0)
    assert exact.sum() >= 0.8 * det.sum()                                             # inlier counts: integer work
    np.testing.assert_allclose(rows[exact, 6:10], ref_rows[exact, 6:10], rtol=1e-6)
