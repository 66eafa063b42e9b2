"""``install()`` swaps the reference's per-cluster-pair registration callables for the B200 engine.

The reference pipeline (``utils_track.track`` -> ``utils_match.match_pcds`` -> ``match_pairs``) stays untouched;
only the hot-path entry points are rebound, exactly at the seams SURVEY.md section 8b lists:

    utils_match.hist_icp              <- icp_flow_b200.hist_icp            (sole caller: utils_match.py:92)
    utils_match.match_eval            <- icp_flow_b200.match_eval          (sole caller: utils_match.py:93)
    utils_match.match_pairs           <- icp_flow_b200.match_pairs         (callers: utils_match.py:38,55)
    utils_match.match_pcds            <- icp_flow_b200.match_pcds          (caller: utils_track.py:32; also rebound there)
    utils_check.sanity_check          <- icp_flow_b200.sanity_check        (callers: utils_match.py:35,51)
    utils_flow.flow_estimation_torch  <- icp_flow_b200.flow_estimation_torch (caller: demo.py:222)
    utils_hist.estimate_init_pose     <- icp_flow_b200.estimate_init_pose
    utils_icp.apply_icp               <- icp_flow_b200.apply_icp
    utils_icp.pytorch3d_icp           <- icp_flow_b200.pytorch3d_icp
    utils_icp_pytorch3d.iterative_closest_point <- icp_flow_b200.iterative_closest_point
    utils_helper.nearest_neighbor_batch / transform_points_batch (and the names re-imported by utils_match / utils_icp /
    utils_hist) <- the engine's, used by match_eval on CUDA tensors

    utils_cluster.cluster_hdbscan / cluster_dbscan / cluster_pcd <- icp_flow_b200.cluster_hdbscan / cluster_dbscan / cluster_pcd
                                                              (opt-in: patch_clustering=True)

``uninstall()`` restores the originals.  The reference modules must already be importable (``sys.path``).
"""
from __future__ import annotations

import importlib
import sys

from . import cluster, ops, scan

_SAVED = {}

_BINDINGS = (
    ("utils_match", "hist_icp", ops.hist_icp),
    ("utils_match", "match_eval", ops.match_eval),
    ("utils_match", "match_pairs", scan.match_pairs),
    ("utils_match", "match_pcds", scan.match_pcds),
    ("utils_match", "sanity_check", scan.sanity_check),
    ("utils_check", "sanity_check", scan.sanity_check),
    ("utils_flow", "flow_estimation_torch", scan.flow_estimation_torch),
    ("utils_match", "estimate_init_pose", ops.estimate_init_pose),
    ("utils_match", "apply_icp", ops.apply_icp),
    ("utils_hist", "estimate_init_pose", ops.estimate_init_pose),
    ("utils_icp", "apply_icp", ops.apply_icp),
    ("utils_icp", "pytorch3d_icp", ops.pytorch3d_icp),
    ("utils_icp", "iterative_closest_point", ops.iterative_closest_point),
    ("utils_icp_pytorch3d", "iterative_closest_point", ops.iterative_closest_point),
)

# names other reference modules copied with `from ... import ...`: rebound only where the module is already loaded
_LOADED_ONLY = (
    ("utils_track", "match_pcds", scan.match_pcds),
    ("demo", "flow_estimation_torch", scan.flow_estimation_torch),
)

# the reference's clusterers (hdbscan with --if_hdbscan, else Open3D DBSCAN) -- opt-in,
# `install(patch_clustering=True)`
_CLUSTERING = (
    ("utils_cluster", "cluster_hdbscan", cluster.cluster_hdbscan),
    ("utils_cluster", "cluster_dbscan", cluster.cluster_dbscan),
    ("utils_cluster", "cluster_pcd", cluster.cluster_pcd),
)
_CLUSTERING_LOADED_ONLY = (
    ("demo", "cluster_pcd", cluster.cluster_pcd),
    ("dataset_pca", "cluster_pcd", cluster.cluster_pcd),
    ("dataset_argo", "cluster_pcd", cluster.cluster_pcd),
)

_HELPERS = (
    ("utils_helper", "nearest_neighbor_batch", ops.nearest_neighbor_batch),
    ("utils_helper", "transform_points_batch", ops.transform_points_batch),
    ("utils_match", "nearest_neighbor_batch", ops.nearest_neighbor_batch),
    ("utils_match", "transform_points_batch", ops.transform_points_batch),
)


def install(patch_helpers: bool = False, patch_clustering: bool = False):
    """Rebind the reference's hot-path callables.  ``patch_helpers`` also routes the NN / transform helpers that
    ``match_eval`` calls (they require CUDA tensors); ``patch_clustering`` routes ``utils_cluster.cluster_dbscan`` /
    ``cluster_pcd`` (the reference's default, non-HDBSCAN clusterer) to the GPU DBSCAN."""
    todo = _BINDINGS + (_HELPERS if patch_helpers else ()) + (_CLUSTERING if patch_clustering else ())
    for mod_name, attr, fn in todo:
        mod = sys.modules.get(mod_name) or importlib.import_module(mod_name)
        if hasattr(mod, attr):
            _SAVED.setdefault((mod_name, attr), getattr(mod, attr))
            setattr(mod, attr, fn)
    for mod_name, attr, fn in _LOADED_ONLY + (_CLUSTERING_LOADED_ONLY if patch_clustering else ()):
        mod = sys.modules.get(mod_name)
        if mod is not None and hasattr(mod, attr):
            _SAVED.setdefault((mod_name, attr), getattr(mod, attr))
            setattr(mod, attr, fn)
    return sorted({m for m, _, _ in todo})


def uninstall():
    for (mod_name, attr), fn in list(_SAVED.items()):
        mod = sys.modules.get(mod_name)
        if mod is not None:
            setattr(mod, attr, fn)
        del _SAVED[(mod_name, attr)]
