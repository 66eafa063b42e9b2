"""Clustering of a scan on the GPU -- SURVEY.md section 8, row f4: the labels the per-pair path consumes.

Mirror of the reference's ``utils_cluster`` (``/root/reference/utils_cluster.py``): ``cluster_hdbscan(args, points)`` (the
clusterer of every script of the reference, ``--if_hdbscan``) = HDBSCAN(min_cluster_size=args.min_cluster_size,
min_samples=None) and ``cluster_dbscan(args, points)`` = Open3D ``cluster_dbscan(eps=args.epsilon,
min_points=args.min_cluster_size)``, each followed by "keep the ``args.num_clusters`` largest clusters";
``cluster_pcd(args, points, idxs_nonground)``; plus the z-threshold ground removal of ``utils_ground.segment_ground_thres``
(``utils_ground.py:27-34``).  DBSCAN runs in ``icpf_dbscan_f32`` (``csrc/icpf_cluster.cu``), HDBSCAN's quadratic stages in
``icpf_hdbscan_mst_f32`` (``csrc/icpf_hdbscan.cu``); both give the partition of the sequential CPU algorithm (oracle:
scikit-learn).  Patchwork++ ground removal is a third-party CPU library in the reference and stays there.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib, ops
from .ops import _ptr


def dbscan_labels(points: torch.Tensor, eps: float, min_points: int, return_count: bool = False):
    """Raw DBSCAN labels of a CUDA fp32 ``[n, >=3]`` scan: ``[n]`` int32 on the device, -1 = noise, clusters numbered by
    their lowest core-point index.  Stream-ordered, no host synchronisation."""
    if not torch.is_tensor(points) or not points.is_cuda:
        raise RuntimeError("points must be a CUDA tensor: icp_flow_b200 has no CPU implementation")
    if points.dim() != 2 or points.shape[1] < 3:
        raise ValueError("points must be [n, >=3] (x, y, z, ...)")
    pts = points.float()
    if pts.stride(1) != 1:
        pts = pts.contiguous()
    n = pts.shape[0]
    labels = torch.empty(n, device=pts.device, dtype=torch.int32)
    count = torch.zeros(1, device=pts.device, dtype=torch.int32)
    L = _lib.lib()
    ws = torch.empty(max(int(L.icpf_dbscan_workspace_bytes(n)), 1), device=pts.device, dtype=torch.uint8)
    with torch.cuda.device(pts.device):
        code = L.icpf_dbscan_f32(_ptr(pts), pts.stride(0) if n > 0 else 3, n, float(eps), int(min_points), _ptr(labels),
                                 _ptr(count), _ptr(ws), ws.numel(), ops._stream_ptr())
    _lib.check(code, "icpf_dbscan_f32")
    return (labels, count) if return_count else labels


def keep_largest(labels: np.ndarray, num_clusters: int) -> np.ndarray:
    """utils_cluster.py:40-46 restated on the label array: clusters outside the ``num_clusters`` largest become -1.
    (Host logic on a few hundred counts; the order among equally large clusters is numpy's argsort, as in the reference.)"""
    labels = np.array(labels)
    ids, counts = np.unique(labels, return_counts=True)
    ids, counts = ids[1:], counts[1:]       # the reference drops the first unique label, taking it for the noise label -1
    if ids.size == 0:                       # no cluster at all (the reference would raise on the empty array): all noise
        labels[:] = -1
        return labels
    # ascending argsort of the sizes, read backwards: among equally large clusters the order is numpy's, as in the reference
    keep = ids[np.argsort(counts)[::-1][:num_clusters]]
    labels[~np.isin(labels, keep)] = -1
    return labels


def cluster_dbscan(args, points) -> np.ndarray:
    """Drop-in for ``utils_cluster.cluster_dbscan(args, points)`` (reads ``args.epsilon``, ``args.min_cluster_size``,
    ``args.num_clusters``); ``points`` numpy or tensor ``[n, >=3]``; returns numpy int labels like the reference."""
    dev = points.device if torch.is_tensor(points) and points.is_cuda else torch.device("cuda", torch.cuda.current_device())
    pts = torch.as_tensor(np.ascontiguousarray(points[:, :3], dtype=np.float32)) if not torch.is_tensor(points) else points[:, :3]
    raw = dbscan_labels(pts.to(dev).float().contiguous(), args.epsilon, args.min_cluster_size)
    return keep_largest(raw.cpu().numpy().astype(np.int64), args.num_clusters)


def hdbscan_labels(points: torch.Tensor, min_cluster_size: int, min_samples: int | None = None,
                   exact_order: bool = True) -> np.ndarray:
    """HDBSCAN labels of a CUDA fp32 ``[n, >=3]`` scan (all rows finite): numpy ``[n]`` int64, -1 = noise, clusters
    numbered by their lowest point.  The two quadratic stages (core distances, Prim's minimum spanning tree of the mutual-
    reachability graph) run on the GPU (``icpf_hdbscan_mst_f32``); the n - 1 tree edges come back to the host, are sorted
    with ``np.argsort`` -- the call scikit-learn's implementation makes, so that edges of equal weight merge in the same
    order -- and condensed / selected by ``icpf_hdbscan_labels_host``.  Same partition as
    ``sklearn.cluster.HDBSCAN(min_cluster_size, min_samples).fit(points).labels_``.

    ``exact_order=False``: the spanning tree that is unique under the edge order (weight, min(a,b), max(a,b)), built by
    Boruvka rounds (~10x faster: a dozen rounds instead of n - 1 dependent steps) and merged in that order.  Same weights,
    but among EQUAL weights not the oracle's order, so a few labels per scan may differ where the oracle's own result
    depends on it (adjusted Rand index against sklearn >= 0.99 on the fixtures; the reference itself asks its library for
    an approximate tree)."""
    if not torch.is_tensor(points) or not points.is_cuda:
        raise RuntimeError("points must be a CUDA tensor: icp_flow_b200 has no CPU implementation")
    if points.dim() != 2 or points.shape[1] < 3:
        raise ValueError("points must be [n, >=3] (x, y, z, ...)")
    pts = points.float().contiguous()
    n = int(pts.shape[0])
    k = int(min_cluster_size if min_samples is None else min_samples)
    if n == 0:
        return np.zeros(0, np.int64)
    if not bool(torch.isfinite(pts[:, :3]).all()):
        raise ValueError("HDBSCAN needs finite coordinates")
    dev = pts.device
    L = _lib.lib()
    core = torch.empty(n, device=dev, dtype=torch.float64)
    e_src = torch.empty(max(n - 1, 1), device=dev, dtype=torch.int32)
    e_dst = torch.empty(max(n - 1, 1), device=dev, dtype=torch.int32)
    e_w = torch.empty(max(n - 1, 1), device=dev, dtype=torch.float64)
    ws = torch.empty(int(L.icpf_hdbscan_workspace_bytes(n)) + 256, device=dev, dtype=torch.uint8)
    off = (-ws.data_ptr()) % 256
    with torch.cuda.device(dev):
        code = L.icpf_hdbscan_mst_f32(_ptr(pts), int(pts.shape[1]), n, k, 1 if exact_order else 0, _ptr(core), _ptr(e_src), _ptr(e_dst), _ptr(e_w),
                                      ctypes.c_void_p(ws.data_ptr() + off), ws.numel() - off, ops._stream_ptr())
    _lib.check(code, "icpf_hdbscan_mst_f32")
    labels = np.full(n, -1, np.int32)
    if n > 1:
        w = e_w[:n - 1].cpu().numpy()
        a, b = e_src[:n - 1].cpu().numpy(), e_dst[:n - 1].cpu().numpy()
        if exact_order:
            order = np.argsort(w)                           # sklearn/cluster/_hdbscan/hdbscan.py: _process_mst
            a, b, w = a[order], b[order], w[order]
        a, b, w = np.ascontiguousarray(a), np.ascontiguousarray(b), np.ascontiguousarray(w)
        code = L.icpf_hdbscan_labels_host(a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p),
                                          w.ctypes.data_as(ctypes.c_void_p), n, int(min_cluster_size), 1 if exact_order else 0,
                                          labels.ctypes.data_as(ctypes.c_void_p))
        _lib.check(code, "icpf_hdbscan_labels_host")
    return labels.astype(np.int64)


def cluster_hdbscan(args, points) -> np.ndarray:
    """Drop-in for ``utils_cluster.cluster_hdbscan(args, points)`` (utils_cluster.py:10-29; reads ``args.min_cluster_size``,
    ``args.num_clusters``): HDBSCAN with ``min_samples = None``, then "keep the ``args.num_clusters`` largest clusters"."""
    dev = points.device if torch.is_tensor(points) and points.is_cuda else torch.device("cuda", torch.cuda.current_device())
    pts = torch.as_tensor(np.ascontiguousarray(points[:, :3], dtype=np.float32)) if not torch.is_tensor(points) else points[:, :3]
    raw = hdbscan_labels(pts.to(dev).float().contiguous(), args.min_cluster_size,
                         exact_order=not getattr(args, "hdbscan_any_order", False))
    return keep_largest(raw, args.num_clusters)


def cluster_pcd(args, points, idxs_nonground) -> np.ndarray:
    """Drop-in for ``utils_cluster.cluster_pcd`` (utils_cluster.py:50-63): ground points -1e8, the rest HDBSCAN
    (``args.if_hdbscan``) or DBSCAN labels."""
    pts = points.cpu().numpy() if torch.is_tensor(points) else np.asarray(points)
    idx = idxs_nonground.cpu().numpy() if torch.is_tensor(idxs_nonground) else np.asarray(idxs_nonground)
    labels_nonground = cluster_hdbscan(args, pts[idx]) if getattr(args, "if_hdbscan", False) else cluster_dbscan(args, pts[idx])
    labels = np.zeros((len(pts))) - 1e8
    labels[idx] = labels_nonground
    return labels


def segment_ground_thres(args, points) -> np.ndarray:
    """utils_ground.segment_ground_thres (utils_ground.py:27-34): True = non-ground, z above range_z + ground_slack."""
    z = points[:, 2].cpu().numpy() if torch.is_tensor(points) else np.asarray(points)[:, 2]
    labels = np.ones((len(z))).astype(bool)
    labels[z <= args.range_z + args.ground_slack] = False
    return labels
