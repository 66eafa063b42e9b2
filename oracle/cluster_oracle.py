"""oracle/cluster_oracle.py -- TEST INFRASTRUCTURE (CPU checker), never imported by the product path.

Row f4 of SURVEY.md section 8: the reference clusters with Open3D's `cluster_dbscan` (utils_cluster.py:32-48; third
party, not installable here).  The oracle is scikit-learn's DBSCAN -- the same sequential algorithm (core = at least
`min_samples` points within eps, the point itself included; clusters grown one after the other in index order; a border
point stays with the first cluster that reaches it) on float64 copies of the points through a kd-tree radius query --
followed by the reference's own "keep the num_clusters largest" lines restated.  The two libraries differ only for a
pair of points EXACTLY eps apart (sklearn `<=`, FLANN `<`): measure zero on real scans, and absent from the fixtures.
Parity is therefore pinned to sklearn 1.9's labels, stated as such in DESIGN.md.
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
