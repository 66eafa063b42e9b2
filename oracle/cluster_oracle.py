"""oracle/cluster_oracle.py -- TEST INFRASTRUCTURE (CPU checker), never imported by the product path.

Row f4 of SURVEY.md section 8: the reference clusters with Open3D's `cluster_dbscan` (utils_cluster.py:32-48; third
party, not installable here).  The oracle is scikit-learn's DBSCAN -- the same sequential algorithm (core = at least
`min_samples` points within eps, the point itself included; clusters grown one after the other in index order; a border
point stays with the first cluster that reaches it) on float64 copies of the points through a kd-tree radius query --
followed by the reference's own "keep the num_clusters largest" lines restated.  The two libraries differ only for a
pair of points EXACTLY eps apart (sklearn `<=`, FLANN `<`): measure zero on real scans, and absent from the fixtures.
Parity is therefore pinned to sklearn 1.9's labels, stated as such in DESIGN.md.

HDBSCAN (utils_cluster.py:10-29, the clusterer every script of the reference selects): the reference calls the `hdbscan`
package (not installable here); the oracle is scikit-learn's port of that package, `sklearn.cluster.HDBSCAN` with the
reference's arguments (min_samples=None, alpha=1, euclidean, eom) on its kd-tree / Prim path.  The reference itself asks
for an APPROXIMATE spanning tree (approx_min_span_tree=True) and the two libraries need not order equal-weight edges
alike, so parity with the reference's own dependency is UNPINNED; what is pinned is equality with sklearn 1.9's
partition.
"""
from __future__ import annotations

import numpy as np


def dbscan_labels(points: np.ndarray, eps: float, min_points: int) -> np.ndarray:
    from sklearn.cluster import DBSCAN

    pts = np.asarray(points)[:, :3].astype(np.float64)
    if len(pts) == 0:
        return np.zeros(0, dtype=np.int64)
    return DBSCAN(eps=float(eps), min_samples=int(min_points), algorithm="kd_tree").fit(pts).labels_.astype(np.int64)


def cluster_dbscan(points: np.ndarray, eps: float, min_points: int, num_clusters: int) -> np.ndarray:
    """utils_cluster.py:32-48 with the Open3D call replaced by sklearn's."""
    labels = np.array(dbscan_labels(points, eps, min_points))
    lbls, counts = np.unique(labels, return_counts=True)
    cluster_info = np.array(list(zip(lbls[1:], counts[1:])))
    cluster_info = cluster_info[cluster_info[:, 1].argsort()]
    clusters_labels = cluster_info[::-1][:num_clusters, 0]
    labels[np.isin(labels, clusters_labels, invert=True)] = -1
    return labels


def hdbscan_labels(points: np.ndarray, min_cluster_size: int, min_samples=None) -> np.ndarray:
    import warnings
    from sklearn.cluster import HDBSCAN

    pts = np.asarray(points)[:, :3].astype(np.float64)
    if len(pts) == 0:
        return np.zeros(0, dtype=np.int64)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return HDBSCAN(min_cluster_size=int(min_cluster_size), min_samples=min_samples, alpha=1.0, algorithm="kd_tree",
                       leaf_size=100, cluster_selection_method="eom").fit(pts).labels_.astype(np.int64)


def same_partition(a: np.ndarray, b: np.ndarray) -> bool:
    """Equal clusterings up to the numbering of the clusters (noise = noise)."""
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape or not np.array_equal(a < 0, b < 0):
        return False
    pairs = set(zip(a[a >= 0].tolist(), b[b >= 0].tolist()))
    return len(pairs) == len({p[0] for p in pairs}) == len({p[1] for p in pairs})


def cluster_hdbscan(points: np.ndarray, min_cluster_size: int, num_clusters: int) -> np.ndarray:
    """utils_cluster.py:10-29 with the hdbscan call replaced by sklearn's."""
    labels = np.array(hdbscan_labels(points, min_cluster_size))
    lbls, counts = np.unique(labels, return_counts=True)
    cluster_info = np.array(list(zip(lbls[1:], counts[1:])))
    cluster_info = cluster_info[cluster_info[:, 1].argsort()]
    clusters_labels = cluster_info[::-1][:num_clusters, 0]
    labels[np.isin(labels, clusters_labels, invert=True)] = -1
    return labels


def mst_total_order(points: np.ndarray, min_samples: int):
    """The minimum spanning tree of the mutual-reachability graph that is UNIQUE under the strict edge order (weight,
    min(a,b), max(a,b)) -- Kruskal on the dense fp64 graph, for small n: (edge_a, edge_b, edge_w) in that order.  The
    definition the engine's any-order (Boruvka) mode implements."""
    X = np.asarray(points)[:, :3].astype(np.float64)
    n = len(X)
    d2 = np.zeros((n, n))
    for c in range(3):
        diff = X[:, None, c] - X[None, :, c]
        d2 += diff * diff
    D = np.sqrt(d2)
    core = np.sort(D, axis=1)[:, min(min_samples, n) - 1]
    M = np.maximum(np.maximum(core[:, None], core[None, :]), D)
    iu = np.triu_indices(n, 1)
    order = np.lexsort((iu[1], iu[0], M[iu]))
    parent = list(range(n))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    ea, eb, ew = [], [], []
    for e in order:
        a, b = int(iu[0][e]), int(iu[1][e])
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[rb] = ra
            ea.append(a), eb.append(b), ew.append(M[a, b])
            if len(ea) == n - 1:
                break
    return np.array(ea, np.int32), np.array(eb, np.int32), np.array(ew, np.float64)
