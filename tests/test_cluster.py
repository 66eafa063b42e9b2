"""DBSCAN clustering of a scan (utils_cluster.cluster_dbscan, utils_cluster.py:32-48) on the engine
against the CPU oracle (scikit-learn's DBSCAN: oracle/cluster_oracle.py).  Labels are integer work: EXACT equality."""
import types

import numpy as np
import pytest
import torch

from engines import is_simt, put
from icp_flow_b200 import cluster, synth
from oracle import cluster_oracle as CO

pytestmark = pytest.mark.usefixtures("engine")


def _blobs(rng, n_blobs, n_noise, spread=0.12, size=(20, 120)):
    pts = []
    for _ in range(n_blobs):
        c = rng.uniform(-8, 8, 3) * [1, 1, 0.2]
        n = int(rng.integers(*size))
        pts.append(c + rng.normal(0, spread, (n, 3)) * [1.5, 1.0, 0.6])
    pts.append(rng.uniform(-9, 9, (n_noise, 3)) * [1, 1, 0.2])
    pts = np.concatenate(pts).astype(np.float32)
    return pts[rng.permutation(len(pts))]


@pytest.mark.parametrize("seed,eps,min_points", [(0, 0.25, 30), (1, 0.25, 10), (2, 0.4, 5), (3, 0.15, 8)])
def test_dbscan_labels_equal_sklearn(seed, eps, min_points):
    rng = np.random.default_rng(seed)
    pts = _blobs(rng, 12, 300)
    got = cluster.dbscan_labels(put(torch.from_numpy(pts)), eps, min_points).cpu().numpy()
    want = CO.dbscan_labels(pts, eps, min_points)
    assert want.max() >= 2                       # the fixture really has clusters, border points and noise
    assert (want == -1).any()
    assert np.array_equal(got, want), (int((got != want).sum()), len(pts))


def test_dbscan_edge_cases():
    eps, mp = 0.25, 4
    # empty scan, a single point, all noise, one dense blob, duplicated points, a non-finite row, a strided [n,4] input
    assert cluster.dbscan_labels(put(torch.zeros(0, 3)), eps, mp).numel() == 0
    assert cluster.dbscan_labels(put(torch.zeros(1, 3)), eps, mp).cpu().tolist() == [-1]
    far = np.arange(30, dtype=np.float32)[:, None] * np.array([[3.0, 0, 0]], np.float32)
    assert (cluster.dbscan_labels(put(torch.from_numpy(far)), eps, mp).cpu().numpy() == -1).all()
    rng = np.random.default_rng(5)
    blob = rng.normal(0, 0.05, (200, 3)).astype(np.float32)
    assert (cluster.dbscan_labels(put(torch.from_numpy(blob)), eps, mp).cpu().numpy() == 0).all()
    dup = np.repeat(rng.uniform(-3, 3, (7, 3)).astype(np.float32), 5, axis=0)
    assert np.array_equal(cluster.dbscan_labels(put(torch.from_numpy(dup)), eps, mp).cpu().numpy(), CO.dbscan_labels(dup, eps, mp))
    bad = np.concatenate([blob, [[np.nan, 0, 0]], [[np.inf, 1, 1]]]).astype(np.float32)
    got = cluster.dbscan_labels(put(torch.from_numpy(bad)), eps, mp).cpu().numpy()
    assert (got[:200] == 0).all() and (got[200:] == -1).all()
    four = np.concatenate([_blobs(rng, 5, 50), np.ones((0, 0), np.float32).reshape(-1, 0)], axis=1) if False else _blobs(rng, 5, 50)
    padded = np.concatenate([four, np.full((len(four), 1), 7.0, np.float32)], axis=1)
    assert np.array_equal(cluster.dbscan_labels(put(torch.from_numpy(padded)), eps, 10).cpu().numpy(), CO.dbscan_labels(four, eps, 10))


def test_cluster_dbscan_mirror_keeps_the_largest_clusters():
    """utils_cluster.cluster_dbscan semantics: the num_clusters largest clusters keep their DBSCAN labels, the rest -> -1."""
    if is_simt():
        pytest.skip("the numpy front end allocates on the current CUDA device")
    rng = np.random.default_rng(9)
    pts = _blobs(rng, 15, 200)
    args = types.SimpleNamespace(epsilon=0.25, min_cluster_size=10, num_clusters=6, if_hdbscan=False)
    got = cluster.cluster_dbscan(args, pts)
    want = CO.cluster_dbscan(pts, 0.25, 10, 6)
    assert np.array_equal(got, want)
    assert len(np.unique(got[got >= 0])) == 6
    nonground = np.nonzero(pts[:, 2] > -0.5)[0]
    lab = cluster.cluster_pcd(args, pts, nonground)
    assert lab.shape == (len(pts),) and (lab[np.setdiff1d(np.arange(len(pts)), nonground)] == -1e8).all()
    assert np.array_equal(lab[nonground], CO.cluster_dbscan(pts[nonground], 0.25, 10, 6).astype(np.float64))


@pytest.mark.gpu
def test_dbscan_on_a_waymo_shape_scan_equals_sklearn():
    """BASELINE config C4 shape: ~150 k points, 200 objects + ground + clutter (synth.make_scene), the reference's default
    eps = 0.25, min_points = 30: labels identical to sklearn's on every point."""
    if is_simt():
        pytest.skip("GPU-sized scan: the cuda engine only")
    sp, sl, _, _, _ = synth.make_scene()
    nonground = sp[sl > -1e7]
    got = cluster.dbscan_labels(torch.from_numpy(nonground).cuda(), 0.25, 30).cpu().numpy()
    want = CO.dbscan_labels(nonground, 0.25, 30)
    assert want.max() >= 50
    assert np.array_equal(got, want), int((got != want).sum())
