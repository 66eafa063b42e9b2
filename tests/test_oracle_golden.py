"""CPU: the restatement in oracle/icp_oracle.py must reproduce the REAL reference's outputs
(tests/golden/*.npz, written by oracle/gen_golden.py from /root/reference run verbatim) bit for bit."""
import numpy as np
import pytest
import torch

from oracle import icp_oracle as O
from oracle import leaves


def _params(g, **kw):
    return O.PathParams(thres_dist=float(g["thres_dist"]), translation_frame=float(g["translation_frame"]),
                        chunk_size=int(g["chunk_size"]), **kw)


def _swapped_inputs(g):
    src, dst = torch.from_numpy(g["src"]), torch.from_numpy(g["dst"])
    swap = torch.from_numpy(g["swapped"])
    a, c = src.clone(), dst.clone()
    a[swap] = dst[swap]
    c[swap] = src[swap]
    return src, dst, a, c


def test_hist_known_answer(golden):
    """hist_cuda/test.py:19-56 -- the peak of the difference histogram sits at the analytic bin (50,130,7)."""
    g = golden("hist_test_vector.npz")
    h = leaves.hist_votes(torch.from_numpy(g["X"]), torch.from_numpy(g["Y"]), g["mins"], g["maxs"], g["lens"])
    flat = h.reshape(3, -1).argmax(dim=1)
    _, hh, ww, dd = h.shape
    arg = torch.stack([flat // dd // ww % hh, flat // dd % ww, flat % dd], dim=1).numpy()
    assert (arg == np.array([[50, 130, 7]] * 3)).all()
    assert (arg == g["argmax"]).all()
    assert np.array_equal(h.reshape(3, -1).max(dim=1)[0].numpy(), g["peak"])
    assert np.array_equal(h.sum(dim=(1, 2, 3)).numpy(), g["total"])
    # every vote is a pair of flagged rows: total <= n_valid_x * n_valid_y
    nx = (g["X"][:, :, 3] > 0).sum(1)
    ny = (g["Y"][:, :, 3] > 0).sum(1)
    assert (g["total"] <= nx * ny).all()


@pytest.mark.parametrize("name", ["c1_demo.npz", "synth_hist.npz", "synth_hist_default.npz"])
def test_full_path_bitwise(golden, name):
    g = golden(name)
    p = _params(g)
    src, dst, a, c = _swapped_inputs(g)
    T, dbg = O.hist_icp(src, dst, p, return_debug=True)
    assert np.array_equal(dbg["swapped"].numpy(), g["swapped"])
    assert np.array_equal(dbg["init"].numpy(), g["init_pose"])
    assert np.array_equal(dbg["poses_before_unswap"].numpy(), g["T_apply_icp"])
    assert np.array_equal(T.numpy(), g["T_hist_icp"])
    assert int(dbg["iterations"]) == int(g["icp_iterations"])
    idx, dist = O.nearest_neighbor_batch(a, c)
    assert np.array_equal(idx.numpy(), g["nn_idx"])
    assert np.array_equal(dist.numpy(), g["nn_dist"])
    moved = O.transform_points_batch(a, torch.from_numpy(g["init_pose"]))
    tr = O.icp_loop(moved, c, thres=p.thres_dist, max_iterations=100, relative_rmse_thr=1e-6)
    assert np.array_equal(tr.R.numpy(), g["icp_R"])
    assert np.array_equal(tr.T.numpy(), g["icp_T"])
    assert np.array_equal(tr.rmse.numpy(), g["icp_rmse"], equal_nan=True)
    assert tr.converged == bool(g["icp_converged"])


@pytest.mark.parametrize("tag", ["full", "ragged"])
def test_icp_only_bitwise(golden, tag):
    g = golden("synth_icp20.npz")
    a, c = torch.from_numpy(g[f"{tag}_src"]), torch.from_numpy(g[f"{tag}_dst"])
    fixed = O.icp_loop(a, c, thres=0.1, max_iterations=20, relative_rmse_thr=-1.0)
    assert fixed.iterations == 20 == int(g[f"{tag}_fixed20_iterations"])
    assert np.array_equal(fixed.R.numpy(), g[f"{tag}_fixed20_R"])
    assert np.array_equal(fixed.T.numpy(), g[f"{tag}_fixed20_T"])
    assert np.array_equal(fixed.rmse.numpy(), g[f"{tag}_fixed20_rmse"], equal_nan=True)
    stop = O.icp_loop(a, c, thres=0.1, max_iterations=100, relative_rmse_thr=1e-6)
    assert stop.iterations == int(g[f"{tag}_stop_iterations"])
    assert stop.converged == bool(g[f"{tag}_stop_converged"])
    assert np.array_equal(stop.R.numpy(), g[f"{tag}_stop_R"])
    assert np.array_equal(stop.T.numpy(), g[f"{tag}_stop_T"])


def test_c1_flow_bitwise(golden):
    g = golden("c1_demo.npz")
    flow = O.flow_from_transforms(torch.from_numpy(g["flow_points"]), torch.from_numpy(g["flow_labels"]),
                                  torch.from_numpy(g["pair_labels"][:, 0]), torch.from_numpy(g["T_hist_icp"]))
    assert np.array_equal(flow.numpy(), g["flow"])


def test_knn_leaf_matches_torch_broadcast():
    """The C leaf and a torch broadcast restatement of pytorch3d knn_points agree exactly (ties -> lowest index)."""
    torch.manual_seed(3)
    p1 = torch.randn(5, 70, 3)
    p2 = torch.randn(5, 90, 3)
    p2[:, 40] = p2[:, 10]           # exact duplicate -> tie
    l1 = torch.tensor([70, 1, 33, 70, 0])
    l2 = torch.tensor([90, 90, 17, 1, 90])
    d_c, i_c = leaves.knn1(p1, p2, l1, l2)
    d_t, i_t = leaves._knn1_torch(p1, p2, l1, l2)
    assert torch.equal(i_c, i_t)
    assert torch.equal(d_c, d_t)
    assert not (i_c == 40).any()


def test_reference_still_matches_when_present(golden):
    """In the build container re-run the real reference on one fixture (guards against stale goldens)."""
    from oracle import ref_loader
    if not ref_loader.reference_available():
        pytest.skip("/root/reference not present (GPU box)")
    import types
    ref = ref_loader.load_reference()
    g = golden("synth_hist.npz")
    args = types.SimpleNamespace(thres_dist=float(g["thres_dist"]), translation_frame=float(g["translation_frame"]),
                                 chunk_size=int(g["chunk_size"]))
    T = ref.utils_match.hist_icp(args, torch.from_numpy(g["src"]), torch.from_numpy(g["dst"]))
    assert np.array_equal(T.numpy(), g["T_hist_icp"])
