"""clustering of a scan -- DBSCAN (utils_cluster.cluster_dbscan, utils_cluster.py:32-48) and
HDBSCAN (utils_cluster.cluster_hdbscan, utils_cluster.py:10-29) -- on the engine against the CPU oracle (scikit-learn:
oracle/cluster_oracle.py).  Labels are integer work: EXACT equality (HDBSCAN: equal partitions, the numbering of the
clusters is the engine's own)."""
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


# ------------------------------------------------------------------------------------------------ HDBSCAN
@pytest.mark.parametrize("seed,mcs", [(0, 10), (1, 20), (2, 30), (3, 15)])
def test_hdbscan_partition_equals_sklearn(seed, mcs):
    """Core distances + Prim's spanning tree on the engine, condensed tree / excess of mass on the host: the partition of
    sklearn.cluster.HDBSCAN(min_cluster_size, min_samples=None) on blobs + clutter and on a slice of a synthetic scan."""
    rng = np.random.default_rng(seed)
    n = 500 if is_simt() else 1500
    for pts in (_blobs(rng, 8, 150, size=(25, 90))[:n], synth.make_scene(num_clusters=8, num_points=2400, seed=seed, max_size=300)[0][:, :3][-n:]):
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        got = cluster.hdbscan_labels(put(torch.from_numpy(pts)), mcs)
        want = CO.hdbscan_labels(pts, mcs)
        assert want.max() >= 2 and (want == -1).any()
        assert CO.same_partition(got, want), (int((got < 0).sum()), int((want < 0).sum()), got.max() + 1, want.max() + 1)
        # clusters are numbered by their lowest point
        first = [int(np.nonzero(got == c)[0][0]) for c in range(got.max() + 1)]
        assert first == sorted(first)


@pytest.mark.parametrize("seed,mcs", [(0, 10), (2, 10), (3, 20)])
def test_hdbscan_any_order_tree(seed, mcs):
    """exact_order=False: the spanning tree that is unique under the edge order (weight, min, max), by Boruvka rounds on the
    engine.  Its labels equal those of the same tree built by Kruskal on the host (oracle) bit for bit; against sklearn
    they may differ where sklearn's own result depends on the order of equal-weight edges -- bounded by the ARI."""
    import ctypes
    from sklearn.metrics import adjusted_rand_score
    from icp_flow_b200 import _lib
    n = 400 if is_simt() else 1200
    pts = np.ascontiguousarray(synth.make_scene(num_clusters=8, num_points=2400, seed=seed, max_size=300)[0][:, :3][-n:], dtype=np.float32)
    got = cluster.hdbscan_labels(put(torch.from_numpy(pts)), mcs, exact_order=False)
    ea, eb, ew = CO.mst_total_order(pts, mcs)
    want = np.empty(n, np.int32)
    rc = _lib.lib().icpf_hdbscan_labels_host(ea.ctypes.data_as(ctypes.c_void_p), eb.ctypes.data_as(ctypes.c_void_p),
                                             ew.ctypes.data_as(ctypes.c_void_p), n, mcs, 0, want.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0 and np.array_equal(got, want)
    assert adjusted_rand_score(CO.hdbscan_labels(pts, mcs), got) > 0.95          # (a few hundred points: each flip weighs)


def test_hdbscan_edge_cases():
    assert len(cluster.hdbscan_labels(put(torch.zeros(0, 3)), 5)) == 0
    assert cluster.hdbscan_labels(put(torch.zeros(1, 3)), 5).tolist() == [-1]
    rng = np.random.default_rng(3)
    few = rng.uniform(-1, 1, (4, 3)).astype(np.float32)                       # fewer points than min_cluster_size: all noise
    assert (cluster.hdbscan_labels(put(torch.from_numpy(few)), 5) == -1).all()
    two = np.concatenate([rng.normal(0, 0.05, (60, 3)), rng.normal(3, 0.05, (70, 3))]).astype(np.float32)
    got = cluster.hdbscan_labels(put(torch.from_numpy(two)), 10)
    assert CO.same_partition(got, CO.hdbscan_labels(two, 10)) and got.max() == 1
    dup = np.repeat(rng.uniform(-3, 3, (9, 3)).astype(np.float32), 12, axis=0)  # duplicated points: zero distances, lambda = inf
    assert CO.same_partition(cluster.hdbscan_labels(put(torch.from_numpy(dup)), 5), CO.hdbscan_labels(dup, 5))
    wide = np.concatenate([two, np.full((len(two), 2), 9.0, np.float32)], axis=1)   # [n,5] rows: only x, y, z are read
    assert np.array_equal(cluster.hdbscan_labels(put(torch.from_numpy(wide)), 10), got)
    with pytest.raises(ValueError, match="finite"):
        cluster.hdbscan_labels(put(torch.tensor([[0.0, 0, 0], [float("nan"), 0, 0]])), 2)
    with pytest.raises(RuntimeError):
        cluster.hdbscan_labels(put(torch.from_numpy(two)), 100)              # min_samples beyond the shared-memory heaps


def test_hdbscan_tree_stage_on_a_given_spanning_tree():
    """icpf_hdbscan_labels_host alone (host arrays in, host arrays out) on a path graph: an outlier that leaves the root
    is noise, two dense runs are two clusters, the sparse bridge falls out of the cluster it hangs on and keeps its
    label; unsorted edges are sorted by (weight, endpoints)."""
    import ctypes
    from icp_flow_b200 import _lib
    x = np.concatenate([[-100.0], np.arange(12) * 0.1, 5.0 + np.arange(3) * 1.0, 10.0 + np.arange(15) * 0.1])
    n = len(x)
    a, b = np.arange(n - 1, dtype=np.int32), np.arange(1, n, dtype=np.int32)
    w = np.diff(x).astype(np.float64)
    perm = np.random.default_rng(0).permutation(n - 1)
    lab = np.empty(n, np.int32)
    L = _lib.lib()
    for aa, bb, ww, presorted in ((a[perm], b[perm], w[perm], 0), (a[np.argsort(w, kind="stable")], b[np.argsort(w, kind="stable")], np.sort(w), 1)):
        aa, bb, ww = (np.ascontiguousarray(v) for v in (aa, bb, ww))
        rc = L.icpf_hdbscan_labels_host(aa.ctypes.data_as(ctypes.c_void_p), bb.ctypes.data_as(ctypes.c_void_p),
                                        ww.ctypes.data_as(ctypes.c_void_p), n, 5, presorted, lab.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0
        assert lab[0] == -1 and (lab[1:13] == 0).all() and (lab[13:16] == 1).all() and (lab[16:] == 1).all()
    assert L.icpf_hdbscan_labels_host(a.ctypes.data_as(ctypes.c_void_p), a.ctypes.data_as(ctypes.c_void_p),
                                      w.ctypes.data_as(ctypes.c_void_p), n, 5, 0, lab.ctypes.data_as(ctypes.c_void_p)) != 0   # self loops: no tree


def test_cluster_hdbscan_mirror_keeps_the_largest_clusters():
    if is_simt():
        pytest.skip("the numpy front end allocates on the current CUDA device")
    rng = np.random.default_rng(11)
    pts = _blobs(rng, 12, 200, size=(30, 100))
    args = types.SimpleNamespace(epsilon=0.25, min_cluster_size=15, num_clusters=5, if_hdbscan=True)
    got = cluster.cluster_hdbscan(args, pts)
    want = CO.cluster_hdbscan(pts, 15, 5)
    assert CO.same_partition(got, want) and len(np.unique(got[got >= 0])) == 5
    nonground = np.nonzero(pts[:, 2] > -0.5)[0]
    lab = cluster.cluster_pcd(args, pts, nonground)
    assert (lab[np.setdiff1d(np.arange(len(pts)), nonground)] == -1e8).all()
    assert CO.same_partition(lab[nonground].astype(np.int64), CO.cluster_hdbscan(pts[nonground], 15, 5))


@pytest.mark.gpu
def test_hdbscan_on_a_scan_equals_sklearn():
    """20 000 non-ground points of the C4-shape scene, the reference's min_cluster_size = 30 (main.sh): the same partition
    as sklearn's HDBSCAN on every point."""
    if is_simt():
        pytest.skip("GPU-sized scan: the cuda engine only")
    sp, sl, _, _, _ = synth.make_scene()
    pts = np.ascontiguousarray(sp[sl > -1e7][:20000, :3], dtype=np.float32)
    got = cluster.hdbscan_labels(torch.from_numpy(pts).cuda(), 30)
    want = CO.hdbscan_labels(pts, 30)
    assert want.max() >= 50 and CO.same_partition(got, want)
    from sklearn.metrics import adjusted_rand_score
    fast = cluster.hdbscan_labels(torch.from_numpy(pts).cuda(), 30, exact_order=False)
    assert adjusted_rand_score(want, fast) > 0.99
