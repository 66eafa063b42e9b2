"""icp_flow_b200 -- B200-native batched ICP registration engine behind ICP-Flow's per-cluster-pair alignment path.

Public surface = the reference's own callables for that path (same names / signatures), backed by hand-written sm_100a
kernels reached through the C ABI in ``include/icpflow_b200.h`` (``libicpflow_b200.so``, built in-tree).  There is no
CPU or PyTorch fallback: importing works anywhere (so the CPU test-suite can check the ABI), calling needs a GPU.
"""
from .ops import (ICPSolution, SimilarityTransform, IcpBatchResult, apply_icp, estimate_init_pose, hist, hist_icp,
                  icp_batch, iterative_closest_point, make_params, match_eval, match_select, nearest_neighbor_batch,
                  pytorch3d_icp, transform_points_batch)
from .scan import (ScanIndex, flow_estimation, flow_estimation_torch, match_pairs, match_pcds, pad_pairs,
                   sanity_check, scan_index)
from .cluster import cluster_dbscan, cluster_hdbscan, cluster_pcd, dbscan_labels, hdbscan_labels, segment_ground_thres
from .install import install, uninstall

__all__ = [
    "ICPSolution", "SimilarityTransform", "IcpBatchResult", "apply_icp", "estimate_init_pose", "hist", "hist_icp",
    "icp_batch", "iterative_closest_point", "make_params", "match_eval", "match_pairs", "nearest_neighbor_batch", "pytorch3d_icp",
    "transform_points_batch", "match_select", "ScanIndex", "scan_index", "sanity_check", "pad_pairs", "match_pcds",
    "flow_estimation_torch", "flow_estimation", "cluster_dbscan", "cluster_hdbscan", "cluster_pcd", "dbscan_labels", "hdbscan_labels", "segment_ground_thres",
    "install", "uninstall",
]
