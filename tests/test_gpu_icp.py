"""GPU parity: the CUDA ICP path (through the C ABI) against the CPU oracle and the committed reference goldens.

Tolerances (BASELINE.json north_star: transforms within 1e-4 of the reference, fp32):
  * moved points  max_i |(x_i R + T) - (x_i R_ref + T_ref)|_inf <= 1e-4 m over valid rows  (the flow difference)
  * R entries within 1e-4; T entries reported (|T| grows with the distance of the cluster from the origin)
"""
import numpy as np
import pytest
import torch

from engines import device, is_simt, put, sync
from icp_flow_b200 import ops, synth
from oracle import icp_oracle as O
from parity import assert_icp_parity

# every test runs on the GPU (marked gpu) and through the SIMT-on-CPU emulator build of the kernels (tests/engines.py)
pytestmark = pytest.mark.usefixtures("engine")

TOL = 1e-4


def _moved_err(src, R, T, R_ref, T_ref):
    pts = torch.from_numpy(src[:, :, :3]).double()
    valid = torch.from_numpy(src[:, :, 3] > 0)
    a = torch.bmm(pts, torch.as_tensor(R).double()) + torch.as_tensor(T).double()[:, None, :]
    b = torch.bmm(pts, torch.as_tensor(R_ref).double()) + torch.as_tensor(T_ref).double()[:, None, :]
    return ((a - b).abs().amax(dim=2) * valid).amax(dim=1).numpy()


def _assert_parity(src, dst, r, R_ref, T_ref, its_ref, max_explained=0.1, what="icp", trace=None):
    """Every pair within TOL of the oracle, or adjudicated (tests/parity.py, oracle/adjudicate.py): the engine's
    transform is one of the outcomes the reference admits on that pair.  Returns (err, explained mask)."""
    v = assert_icp_parity(src, dst, r.R.cpu(), r.T.cpu(), int(r.batch.tolist()[0]), R_ref, T_ref, its_ref,
                          max_explained=max_explained, what=what, trace=trace)
    return v.err, v.explained


def _run(src, dst, **kw):
    p = ops.make_params(**kw)
    r = ops.icp_batch(put(torch.from_numpy(src)), put(torch.from_numpy(dst)), p)
    sync()
    return r


@pytest.mark.parametrize("tag", ["full", "ragged"])
def test_fixed_20_iterations_vs_reference_golden(golden, tag):
    """BASELINE config C2 call shape: max_iterations=20, relative_rmse_thr=-1 (no early stop)."""
    g = golden("synth_icp20.npz")
    src, dst = g[f"{tag}_src"], g[f"{tag}_dst"]
    r = _run(src, dst, thres=0.1, max_iterations=20, relative_rmse_thr=-1.0, early_exit=False, batch_stop=True)
    assert r.batch.tolist() == [20, 0]
    assert (r.iterations.cpu().numpy() == 20).all()
    trace = O.icp_loop(torch.from_numpy(src), torch.from_numpy(dst), 0.1, 20, -1.0, diagnostics=True)
    assert np.array_equal(trace.R.numpy(), g[f"{tag}_fixed20_R"])        # the oracle IS the golden (bitwise)
    err, flagged = _assert_parity(src, dst, r, g[f"{tag}_fixed20_R"], g[f"{tag}_fixed20_T"], 20, what=f"fixed20/{tag}")
    ok = ~flagged
    assert np.abs(r.R.cpu().numpy() - g[f"{tag}_fixed20_R"])[ok].max() <= TOL
    assert np.allclose(r.rmse.cpu().numpy()[ok], g[f"{tag}_fixed20_rmse"][ok], rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("tag", ["full", "ragged"])
def test_reference_stopping_rule_vs_golden(golden, tag):
    g = golden("synth_icp20.npz")
    src, dst = g[f"{tag}_src"], g[f"{tag}_dst"]
    r = _run(src, dst, thres=0.1, max_iterations=100, relative_rmse_thr=1e-6, early_exit=True, batch_stop=True)
    its, conv = r.batch.tolist()
    assert conv == int(g[f"{tag}_stop_converged"])
    # the batch stop iteration may move by one when a pair's relative rmse sits at the 1e-6 threshold
    assert abs(its - int(g[f"{tag}_stop_iterations"])) <= 2
    trace = O.icp_loop(torch.from_numpy(src), torch.from_numpy(dst), 0.1, 100, 1e-6, diagnostics=True)
    _assert_parity(src, dst, r, g[f"{tag}_stop_R"], g[f"{tag}_stop_T"], trace.iterations, what=f"stop/{tag}", trace=trace)


def test_early_exit_is_result_identical():
    """SURVEY finding 1: stopping each pair at its bitwise fixed point does not change the result."""
    src, dst, _ = synth.make_pairs(48, 192, seed=11, ragged=True, residual_only=True)
    a = _run(src, dst, max_iterations=60, relative_rmse_thr=-1.0, early_exit=False, batch_stop=False)
    b = _run(src, dst, max_iterations=60, relative_rmse_thr=-1.0, early_exit=True, batch_stop=False)
    fixed = b.iterations.cpu().numpy() < 60
    assert fixed.sum() > 0
    assert torch.equal(a.R[torch.from_numpy(fixed)], b.R[torch.from_numpy(fixed)])
    assert torch.equal(a.T[torch.from_numpy(fixed)], b.T[torch.from_numpy(fixed)])


def test_oracle_parity_on_seeded_batch():
    src, dst, _ = synth.make_pairs(40, 160, seed=21, ragged=True, residual_only=True, wrong_frac=0.15)
    ref = O.icp_loop(torch.from_numpy(src), torch.from_numpy(dst), thres=0.1, max_iterations=100,
                     relative_rmse_thr=1e-6, diagnostics=True)
    r = _run(src, dst, max_iterations=100, relative_rmse_thr=1e-6)
    _assert_parity(src, dst, r, ref.R, ref.T, ref.iterations, max_explained=0.15, what="seeded batch", trace=ref)


def test_mirror_api_returns_reference_types():
    src, dst, _ = synth.make_pairs(8, 64, seed=3, residual_only=True)
    s, d = put(torch.from_numpy(src)), put(torch.from_numpy(dst))
    sol = ops.iterative_closest_point(s, d, thres=0.1, max_iterations=100, relative_rmse_thr=1e-6)
    assert isinstance(sol.converged, bool) and sol.RTs.R.shape == (8, 3, 3) and sol.RTs.T.shape == (8, 3)
    assert sol.Xt.shape == (8, 64, 3) and sol.rmse.shape == (8,) and len(sol.t_history) >= 1
    assert torch.equal(sol.RTs.s, torch.ones(8, device=device()))
    ref = O.icp_loop(torch.from_numpy(src), torch.from_numpy(dst), diagnostics=True)
    assert abs(len(sol.t_history) - ref.iterations) <= 2
    v = assert_icp_parity(src, dst, sol.RTs.R.cpu(), sol.RTs.T.cpu(), len(sol.t_history), ref.R, ref.T, ref.iterations,
                          max_explained=0.25, what="mirror api")      # 8 pairs: one flip is 12.5 %
    ok = torch.from_numpy(~v.explained)
    want = torch.bmm(torch.from_numpy(src[:, :, :3]), ref.R) + ref.T[:, None]
    assert (sol.Xt.cpu() - want)[ok].abs().max() <= TOL
    assert (sol.Xt.cpu() - (torch.bmm(torch.from_numpy(src[:, :, :3]), sol.RTs.R.cpu()) + sol.RTs.T.cpu()[:, None])
            ).abs().max() <= 2e-5


def test_c1_demo_icp_stage_vs_reference_golden(golden):
    """BASELINE config C1 (demo.npz clusters): ICP from the reference's own histogram initialisation."""
    g = golden("c1_demo.npz")
    src, dst = torch.from_numpy(g["src"]), torch.from_numpy(g["dst"])
    swap = torch.from_numpy(g["swapped"])
    a, c = src.clone(), dst.clone()
    a[swap] = dst[swap]; c[swap] = src[swap]
    moved = O.transform_points_batch(a, torch.from_numpy(g["init_pose"]))
    trace = O.icp_loop(moved, c, float(g["thres_dist"]), 100, 1e-6, diagnostics=True)
    assert np.array_equal(trace.R.numpy(), g["icp_R"])
    r = ops.icp_batch(put(moved), put(c), ops.make_params(thres=float(g["thres_dist"])))
    assert abs(r.batch.tolist()[0] - int(g["icp_iterations"])) <= 2
    _assert_parity(moved.numpy(), c.numpy(), r, g["icp_R"], g["icp_T"], trace.iterations, max_explained=0.15, what="c1 icp stage", trace=trace)


def test_degenerate_pairs():
    """identity pair (rmse 0 -> NaN relative rmse in the reference), zero-inlier pair, tiny clusters."""
    rng = np.random.default_rng(0)
    N = 64
    src = np.full((4, N, 4), 1e8, np.float32); src[..., 3] = 0
    dst = src.copy()
    pts = rng.uniform(-1, 1, size=(40, 3)).astype(np.float32) + np.array([20, -10, 1], np.float32)
    src[0, :40, :3] = pts; src[0, :40, 3] = 1; dst[0, :40, :3] = pts; dst[0, :40, 3] = 1        # identical clouds
    src[1, :40, :3] = pts; src[1, :40, 3] = 1; dst[1, :40, :3] = pts + 5; dst[1, :40, 3] = 1    # no inliers
    src[2, :3, :3] = pts[:3]; src[2, :3, 3] = 1; dst[2, :3, :3] = pts[:3] + 0.01; dst[2, :3, 3] = 1  # 3 points
    src[3, :40, :3] = pts; src[3, :40, 3] = 1; dst[3, :1, :3] = pts[:1]; dst[3, :1, 3] = 1       # single dst point
    r = ops.icp_batch(put(torch.from_numpy(src)), put(torch.from_numpy(dst)),
                      ops.make_params(max_iterations=100, relative_rmse_thr=1e-6, batch_stop=False))
    R, T = r.R.cpu().numpy(), r.T.cpu().numpy()
    assert np.isfinite(R).all() and np.isfinite(T).all()
    eye = np.eye(3, dtype=np.float32)
    assert np.abs(R[0] - eye).max() < 1e-5 and np.abs(T[0]).max() < 1e-3
    assert np.array_equal(R[1], eye) and np.array_equal(T[1], np.zeros(3, np.float32))
    # the well-posed pairs stop at their bitwise fixed point; rank-deficient ones (3 points, a single dst point) may
    # keep flipping last bits of an undetermined rotation and simply run all iterations
    assert r.iterations.cpu().numpy()[:2].max() < 100
    ref = O.icp_loop(torch.from_numpy(src[:2]), torch.from_numpy(dst[:2]), max_iterations=30)
    assert _moved_err(src[:2], R[:2], T[:2], ref.R, ref.T).max() <= TOL


def test_nn_and_transform_seams(golden):
    g = golden("c1_demo.npz")
    src, dst = torch.from_numpy(g["src"]), torch.from_numpy(g["dst"])
    swap = torch.from_numpy(g["swapped"])
    a, c = src.clone(), dst.clone()
    a[swap] = dst[swap]; c[swap] = src[swap]
    idx, dist = ops.nearest_neighbor_batch(put(a), put(c))
    assert np.array_equal(idx.cpu().numpy(), g["nn_idx"])            # index work: bit exact
    # squared distances use the reference leaf's fp32 op order; the engine's sqrt is IEEE-rounded while torch's
    # CPU sqrt (the golden) is 1 ulp off on ~0.2 % of the values, hence 1 ulp instead of array_equal
    got, want = dist.cpu().numpy(), g["nn_dist"]
    assert np.allclose(got, want, rtol=1.2e-7, atol=0)
    assert (got != want).mean() < 0.01
    idx3, dist3 = ops.nearest_neighbor_batch(put(a[:, :, :3]), put(c[:, :, :3]))
    assert torch.equal(idx3, idx) and torch.equal(dist3, dist)
    pose = torch.from_numpy(g["init_pose"])
    moved = ops.transform_points_batch(put(a), put(pose)).cpu()
    assert torch.equal(moved, O.transform_points_batch(a, pose))     # translation-only poses: exact
    torch.manual_seed(0)
    q = torch.linalg.qr(torch.randn(len(a), 3, 3))[0]
    pose2 = torch.eye(4).repeat(len(a), 1, 1); pose2[:, :3, :3] = q; pose2[:, :3, 3] = torch.randn(len(a), 3)
    got = ops.transform_points_batch(put(a), put(pose2)).cpu()
    want = O.transform_points_batch(a, pose2)
    valid = a[:, :, 3] > 0
    assert (got - want)[valid].abs().max() <= 2e-5
    assert torch.equal(got[:, :, 3], a[:, :, 3])


@pytest.mark.parametrize("ragged", [False, True])
def test_grid_search_is_bitwise_identical_to_brute_force(ragged):
    """SURVEY finding 2: a radius-bounded NN inside the ICP loop is result-identical to brute force.  Both modes rank
    candidates by (squared distance, original row index), so every output must agree bit for bit."""
    src, dst, _ = synth.make_pairs(64, 512 if not ragged else 384, seed=31, ragged=ragged, residual_only=True,
                                   wrong_frac=0.1)
    # duplicate a few dst rows so that exact distance ties occur
    dst[:, 7, :3] = dst[:, 3, :3]
    for kw in (dict(max_iterations=20, relative_rmse_thr=-1.0, early_exit=False),
               dict(max_iterations=100, relative_rmse_thr=1e-6, early_exit=True)):
        a = _run(src, dst, nn_mode=1, **kw)
        for mode in (2, 3):      # grid, grid + correspondence cache
            b = _run(src, dst, nn_mode=mode, **kw)
            assert torch.equal(a.R, b.R) and torch.equal(a.T, b.T) and torch.equal(a.rmse, b.rmse), mode
            assert torch.equal(a.iterations, b.iterations) and torch.equal(a.conv_mask, b.conv_mask), mode
            assert torch.equal(a.batch, b.batch) and torch.equal(a.pose, b.pose), mode


def test_cache_is_bitwise_identical_under_large_motion():
    """The correspondence cache must fall back to a full search whenever its triangle-inequality bound cannot prove
    the cached neighbour: large initial motion (many refreshes), far-away origin (1 km: fp32 ulp 6e-5 m)."""
    src, dst, _ = synth.make_pairs(48, 256, seed=41, ragged=True, residual_only=False, wrong_frac=0.1)
    src2, dst2 = src.copy(), dst.copy()
    shift = np.array([1000.0, -800.0, 0.0], np.float32)
    src2[:, :, :3] = np.where(src[:, :, 3:4] > 0, src[:, :, :3] + shift, src[:, :, :3])
    dst2[:, :, :3] = np.where(dst[:, :, 3:4] > 0, dst[:, :, :3] + shift, dst[:, :, :3])
    # small translations so that ICP from identity has inliers, plus the original large-motion set
    dst3 = dst.copy()
    dst3[:, :, :3] = np.where(dst[:, :, 3:4] > 0, src[:, :, :3].mean(1, keepdims=True) * 0 + dst[:, :, :3], dst[:, :, :3])
    for s_, d_ in ((src, dst), (src2, dst2)):
        for kw in (dict(max_iterations=40, relative_rmse_thr=-1.0, early_exit=False, batch_stop=False),
                   dict(max_iterations=100, relative_rmse_thr=1e-6, early_exit=True, batch_stop=True)):
            a = _run(s_, d_, nn_mode=1, **kw)
            b = _run(s_, d_, nn_mode=3, **kw)
            assert torch.equal(a.R, b.R) and torch.equal(a.T, b.T) and torch.equal(a.rmse, b.rmse)
            assert torch.equal(a.iterations, b.iterations) and torch.equal(a.conv_mask, b.conv_mask)


def test_grid_handles_large_and_flat_clusters():
    """Grid cell size adapts when the bbox would need more than kGridMaxCells cells (long wall, planar patch)."""
    rng = np.random.default_rng(5)
    N = 512
    src = np.zeros((3, N, 4), np.float32); src[..., 3] = 1
    wall = np.stack([rng.uniform(0, 30, N), rng.uniform(0, 0.05, N), rng.uniform(0, 4, N)], 1)      # 30 m wall
    flat = np.stack([rng.uniform(0, 6, N), rng.uniform(0, 6, N), np.zeros(N)], 1)                   # exactly planar
    line = np.stack([rng.uniform(0, 8, N), np.zeros(N), np.zeros(N)], 1)                            # collinear
    for k, pts in enumerate((wall, flat, line)):
        src[k, :, :3] = pts + np.array([40.0, -25.0, 1.0])
    dst = src.copy()
    dst[:, :, 0] += 0.03; dst[:, :, 1] -= 0.02
    dst[:, :, :3] += rng.normal(0, 0.002, size=dst[:, :, :3].shape).astype(np.float32)
    a = _run(src, dst, nn_mode=1, max_iterations=30, relative_rmse_thr=-1.0, early_exit=False, batch_stop=False)
    for mode in (2, 3):
        b = _run(src, dst, nn_mode=mode, max_iterations=30, relative_rmse_thr=-1.0, early_exit=False, batch_stop=False)
        assert torch.equal(a.R, b.R) and torch.equal(a.T, b.T) and torch.equal(a.rmse, b.rmse), mode
    T = b.T.cpu().numpy()
    moved = src[:2, :, :3] @ b.R.cpu().numpy()[:2] + T[:2, None]
    assert np.abs(moved - dst[:2, :, :3]).max() < 0.02


def test_batch_stop_beyond_the_first_pass_cap():
    """The first pass is capped at 32 iterations; when the reference's batch stop lies beyond (here: never, because a
    pair without inliers has rmse == 0 and a NaN relative rmse, SURVEY finding 1) a full pass must take over -- same results as
    the oracle, which runs all 100 iterations like the reference."""
    src, dst, _ = synth.make_pairs(12, 128, seed=77, ragged=True, residual_only=True, wrong_frac=0.0)
    dst[0, :, :3] = np.where(dst[0, :, 3:4] > 0, dst[0, :, :3] + 5.0, dst[0, :, :3])   # no inliers: rmse == 0 -> NaN
    ref = O.icp_loop(torch.from_numpy(src), torch.from_numpy(dst), 0.1, 100, 1e-6, diagnostics=True)
    assert ref.iterations == 100 and not ref.converged
    r = _run(src, dst, max_iterations=100, relative_rmse_thr=1e-6, early_exit=True, batch_stop=True)
    assert r.batch.tolist() == [100, 0]
    _assert_parity(src, dst, r, ref.R, ref.T, ref.iterations, max_explained=0.2, what="stop beyond the cap", trace=ref)
    # and a batch whose stop iteration is found below the cap still agrees with the uncapped logic
    ref2 = O.icp_loop(torch.from_numpy(src[1:]), torch.from_numpy(dst[1:]), 0.1, 100, 1e-6, diagnostics=True)
    r2 = _run(src[1:], dst[1:], max_iterations=100, relative_rmse_thr=1e-6, early_exit=True, batch_stop=True)
    assert r2.batch.tolist()[1] == int(ref2.converged) and abs(r2.batch.tolist()[0] - ref2.iterations) <= 2
    if r2.batch.tolist()[0] != ref2.iterations:
        # A flip-prone pair whose relative rmse sits at the 1e-6 threshold moves the BATCH stop by an iteration or two,
        # and with it the state of every pair that is still moving (utils_icp_pytorch3d.py:209 couples the pairs):
        # compare with the oracle stopped at the same iteration.
        ref2 = O.icp_loop(torch.from_numpy(src[1:]), torch.from_numpy(dst[1:]), 0.1, r2.batch.tolist()[0], -1.0,
                          diagnostics=True)
    _assert_parity(src[1:], dst[1:], r2, ref2.R, ref2.T, ref2.iterations, max_explained=0.2, what="stop below the cap", trace=ref2)


@pytest.mark.order_last
@pytest.mark.parametrize("case", ["stop_inside_the_record", "stop_beyond_the_first_pass"])
def test_state_at_the_batch_stop_equals_a_forced_run(case):
    """Whatever way a pair obtains its state at the batch stop k* -- it stopped at its fixed point before, it reads the
    per-iteration record of the first pass (k* < 32), or it is re-run (k* beyond the capped first pass) -- the result
    must be, bit for bit, what running every pair for exactly k*+1 iterations gives (utils_icp_pytorch3d.py:209: the
    reference stops ALL pairs at that iteration)."""
    if case == "stop_inside_the_record":
        src, dst, _ = synth.make_pairs(24, 192, seed=5, ragged=True, residual_only=True, wrong_frac=0.0)
    else:
        src, dst, _ = synth.make_pairs(96, 1024, seed=5, ragged=False, residual_only=True, wrong_frac=0.0)
        keep = [0, 1, 2, 3, 37, 73, 87]            # three slow pairs (41-48 iterations) among fast ones
        src, dst = src[keep], dst[keep]
    r = _run(src, dst, max_iterations=100, relative_rmse_thr=1e-6, early_exit=True, batch_stop=True)
    its, conv = r.batch.tolist()
    assert conv == 1
    if is_simt():
        assert (its < 32) if case == "stop_inside_the_record" else (its > 32), its      # the path this case is about
    assert int(r.iterations.max()) == its and int(r.iterations.min()) < its
    f = _run(src, dst, max_iterations=its, relative_rmse_thr=-1.0, early_exit=False, batch_stop=False)
    assert torch.equal(r.R, f.R) and torch.equal(r.T, f.T) and torch.equal(r.rmse, f.rmse) and torch.equal(r.pose, f.pose)


@pytest.mark.order_last
def test_paused_and_continued_pairs_equal_the_uninterrupted_run():
    """With early exit the first pass is capped at 32 iterations and the pairs still moving there are CONTINUED by the
    full pass (loop state + per-iteration record).  Without early exit nothing is capped or paused: both runs must give
    the same batch stop, the same transforms and the same convergence history up to the stop, bit for bit."""
    src, dst, _ = synth.make_pairs(96, 1024, seed=5, ragged=False, residual_only=True, wrong_frac=0.0)
    keep = [0, 1, 2, 3, 37, 73, 87, 45, 60]
    src, dst = src[keep], dst[keep]
    for mode in (3, 1):
        a = _run(src, dst, max_iterations=100, relative_rmse_thr=1e-6, early_exit=True, batch_stop=True, nn_mode=mode)
        b = _run(src, dst, max_iterations=100, relative_rmse_thr=1e-6, early_exit=False, batch_stop=True, nn_mode=mode)
        its = a.batch.tolist()[0]
        assert a.batch.tolist() == b.batch.tolist()
        if is_simt():
            assert its > 32
        assert torch.equal(a.R, b.R) and torch.equal(a.T, b.T) and torch.equal(a.rmse, b.rmse), mode
        ca, cb = a.conv_mask.cpu().numpy().astype(np.uint32), b.conv_mask.cpu().numpy().astype(np.uint32)
        for k in range(min(its, 128)):
            assert np.array_equal((ca[:, k >> 5] >> (k & 31)) & 1, (cb[:, k >> 5] >> (k & 31)) & 1), (mode, k)
