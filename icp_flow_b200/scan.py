"""Scan-level callers of the per-cluster-pair path, same names and signatures
as the reference callables they replace:

    sanity_check            /root/reference/utils_check.py:21-49
    match_pairs             /root/reference/utils_match.py:69-135   (gather + pad_segment loop: utils_helper.py:185-201)
    match_pcds              /root/reference/utils_match.py:26-66    (pair enumeration, static then dynamic stage)
    flow_estimation_torch   /root/reference/utils_flow.py:57-69
    flow_estimation         /root/reference/utils_flow.py:23-50     (numpy front end of the same kernel)

The reference addresses a cluster as ``points[labels == l]`` -- a boolean mask over the whole scan for every use, inside
Python loops with host syncs.  Here each scan gets ONE cluster index (``ScanIndex``: stable counting sort + per-cluster
statistics, ``icpf_cluster_index_f32``) and every consumer is one launch on it.  No CPU / PyTorch fallback: CPU tensors
raise.  PyTorch is used for device memory, streams and the few [K]-sized index manipulations of the host logic.
"""
from __future__ import annotations

import weakref
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from .ops import _ptr, _require_cuda_f32, _stream_ptr, hist_icp, match_eval, match_select

MAX_LABELS = 1 << 20


class ScanIndex:
    """Cluster index of one scan: ``order[offsets[l]:offsets[l+1]]`` are the rows of ``points[labels == l]`` in scan
    order; ``stats[l] = (mean xyz, sorted bbox extents, 0, 0)``.  Labels that are not integers in ``[0, n_labels)``
    (ground -1e8, unclustered -1) are not indexed."""

    def __init__(self, points: torch.Tensor, labels: torch.Tensor, n_labels: Optional[int] = None):
        if not torch.is_tensor(points) or not torch.is_tensor(labels):
            raise TypeError("points and labels must be torch tensors")
        if not points.is_cuda or not labels.is_cuda:
            raise RuntimeError("points / labels must be CUDA tensors: icp_flow_b200 has no CPU implementation")
        if points.dim() != 2 or points.shape[1] < 3:
            raise ValueError("points must be [n, >=3]")
        if labels.dim() != 1 or len(labels) != len(points):
            raise ValueError("labels must be [n], one per point")
        pts = points if points.dtype == torch.float32 else points.float()
        if pts.stride(1) != 1 or (len(pts) > 1 and pts.stride(0) < pts.shape[1]):
            pts = pts.contiguous()
        self.points = pts
        self.stride = int(pts.stride(0)) if len(pts) > 1 else int(pts.shape[1])
        self.labels = labels.to(torch.float32).contiguous()
        self.n = int(len(pts))
        dev = pts.device
        if n_labels is None:
            n_labels = int(self.labels.max().item()) + 1 if self.n > 0 else 1
        n_labels = max(1, int(n_labels))
        if n_labels > MAX_LABELS:
            raise ValueError(f"cluster labels must be below {MAX_LABELS} (got {n_labels - 1})")
        self.n_labels = n_labels
        self.order = torch.empty(max(self.n, 1), device=dev, dtype=torch.int32)
        self.offsets = torch.empty(n_labels + 1, device=dev, dtype=torch.int32)
        self.stats = torch.empty(n_labels, 8, device=dev, dtype=torch.float32)
        L = _lib.lib()
        ws = torch.empty(max(1, L.icpf_cluster_index_workspace_bytes(self.n, n_labels)), device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            code = L.icpf_cluster_index_f32(_ptr(self.points), self.stride, _ptr(self.labels), self.n, n_labels,
                                            _ptr(self.order), _ptr(self.offsets), _ptr(self.stats), _ptr(ws),
                                            ws.numel(), _stream_ptr())
        _lib.check(code, "icpf_cluster_index_f32")
        self._counts_host = None
        self._present_mask = None
        self._present = None

    @property
    def counts(self) -> torch.Tensor:
        """[n_labels] int32 rows per label (device)."""
        return self.offsets[1:] - self.offsets[:-1]

    @property
    def counts_host(self) -> np.ndarray:
        if self._counts_host is None:
            self._counts_host = self.counts.cpu().numpy()
        return self._counts_host

    @property
    def present_mask(self) -> torch.Tensor:
        """[n_labels] bool: the label owns at least one point (cached)."""
        if self._present_mask is None:
            self._present_mask = self.counts > 0
        return self._present_mask

    def present_labels(self) -> torch.Tensor:
        """Sorted int64 labels >= 0 that own at least one point (``torch.unique(labels)`` without the negative ones)."""
        if self._present is None:
            self._present = torch.nonzero(self.present_mask).flatten()
        return self._present


# (points, labels) -> ScanIndex, so that the reference's own call sequence (sanity_check, match_pairs, again for the
# dynamic stage) builds each scan's index once.  Keyed on tensor identity + version counters.
_CACHE = []
_CACHE_SIZE = 4


def scan_index(points: torch.Tensor, labels: torch.Tensor) -> ScanIndex:
    """The cluster index of a scan, cached per (points tensor, labels tensor) OBJECT and their autograd versions: the same
    scan is addressed several times per frame pair (sanity_check, both stages of match_pcds, flow).  Cache contract: an
    in-place edit that bumps ``_version`` (every torch in-place op) invalidates the entry; writes that do NOT bump it --
    through ``.data``, a numpy view, another tensor aliasing the same storage, or a kernel writing the raw pointer -- are
    invisible here: call ``clear_cache()`` after such a write (or pass fresh tensors)."""
    for ent in _CACHE:
        rp, rl, vp, vl, idx = ent
        if rp() is points and rl() is labels and vp == points._version and vl == labels._version:
            return idx
    idx = ScanIndex(points, labels)
    _CACHE.append((weakref.ref(points), weakref.ref(labels), points._version, labels._version, idx))
    del _CACHE[:-_CACHE_SIZE]
    return idx


def clear_cache() -> None:
    del _CACHE[:]


def _pairs_i64(pairs, dev) -> torch.Tensor:
    if not torch.is_tensor(pairs):
        pairs = torch.as_tensor(np.asarray(pairs))
    if pairs.dim() != 2 or pairs.shape[1] != 2:
        raise ValueError("pairs must be [P, 2] (src label, dst label)")
    return pairs.to(device=dev, dtype=torch.int64).contiguous()


def sanity_check_indexed(args, si: ScanIndex, di: ScanIndex, pairs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """One launch over the candidate pairs: (kept pairs [K,2] int64 in input order, keep flags [P] int32)."""
    dev = si.points.device
    p64 = _pairs_i64(pairs, dev)
    P = len(p64)
    keep = torch.empty(max(P, 1), device=dev, dtype=torch.int32)
    out = torch.empty(max(P, 1), 2, device=dev, dtype=torch.int64)
    count = torch.zeros(1, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        code = _lib.lib().icpf_sanity_check_f32(_ptr(si.offsets), _ptr(si.stats), si.n_labels, _ptr(di.offsets),
                                                _ptr(di.stats), di.n_labels, _ptr(p64), P,
                                                int(args.min_cluster_size), float(args.translation_frame),
                                                float(args.thres_box), _ptr(keep), _ptr(out), _ptr(count),
                                                _stream_ptr())
    _lib.check(code, "icpf_sanity_check_f32")
    k = int(count.item())                      # data-dependent output length: the one host sync
    return out[:k], keep[:P]


def sanity_check(args, src_points, dst_points, src_labels, dst_labels, pairs):
    """Drop-in for ``utils_check.sanity_check`` (utils_check.py:21-49): the candidate pairs that are worth registering,
    in input order, as a ``[K,2]`` tensor of the dtype of ``pairs`` (``torch.zeros((0,2))`` when none is left)."""
    si, di = scan_index(src_points, src_labels), scan_index(dst_points, dst_labels)
    kept, _ = sanity_check_indexed(args, si, di, pairs)
    if len(kept) == 0:
        return torch.zeros((0, 2), device=si.points.device)
    return kept.to(pairs.dtype) if torch.is_tensor(pairs) else kept


def pad_pairs(si: ScanIndex, di: ScanIndex, pairs: torch.Tensor, max_points: int):
    """The gather / ``pad_segment`` loop of match_pairs (utils_match.py:81-91) as one launch: ``(segs_src, segs_dst)``,
    each ``[P, max_points, 4]``.  A cluster with more than ``max_points`` rows keeps ``torch.randperm(len)[:max_points]``
    like the reference (utils_helper.py:187-189,198-201); the permutations are drawn here with the same call in the same
    order (pair by pair, src before dst), i.e. from the same RNG stream."""
    dev = si.points.device
    p64 = _pairs_i64(pairs, dev)
    P = len(p64)
    max_points = int(max_points)
    segs_src = torch.empty(P, max_points, 4, device=dev, dtype=torch.float32)
    segs_dst = torch.empty(P, max_points, 4, device=dev, dtype=torch.float32)
    if P == 0:
        return segs_src, segs_dst
    sample_rows = sample_offsets = None
    if max(int(si.counts_host.max(initial=0)), int(di.counts_host.max(initial=0))) > max_points:
        ph = p64.cpu().numpy()
        cnt = np.zeros((P, 2), np.int64)
        for col, idx in ((0, si), (1, di)):
            lab = ph[:, col]
            ok = (lab >= 0) & (lab < idx.n_labels)
            cnt[ok, col] = idx.counts_host[lab[ok]]
        over = np.flatnonzero(cnt.reshape(-1) > max_points)          # row-major: pair by pair, src before dst
        if len(over) > 0:
            offs = np.full(P * 2, -1, np.int64)
            perms = []
            for j, flat in enumerate(over):
                perm = torch.randperm(int(cnt.reshape(-1)[flat]))[0:max_points]
                perms.append(perm.to(device=dev, dtype=torch.int32))
                offs[flat] = j * max_points
            sample_rows = torch.cat(perms).contiguous()
            sample_offsets = torch.from_numpy(offs).to(dev)
    with torch.cuda.device(dev):
        code = _lib.lib().icpf_gather_pairs_f32(_ptr(si.points), si.stride, _ptr(si.order), _ptr(si.offsets),
                                                si.n_labels, _ptr(di.points), di.stride, _ptr(di.order),
                                                _ptr(di.offsets), di.n_labels, _ptr(p64), P, max_points,
                                                _ptr(sample_rows), _ptr(sample_offsets), _ptr(segs_src),
                                                _ptr(segs_dst), _stream_ptr())
    _lib.check(code, "icpf_gather_pairs_f32")
    return segs_src, segs_dst


# The reference pads every cluster of a frame to args.max_points rows (default 10 000).  The path's results do not depend
# on the padding (rows beyond the valid prefix are never a nearest neighbour and never counted; padding-invariance is a
# tested, bit-exact property), so a batch is padded only up to its own largest cluster, rounded up to PAD_QUANTUM rows:
# the clusters of a real frame then run the shared-memory kernels instead of the large-cluster variants.
ADAPTIVE_PAD = True
PAD_QUANTUM = 256


def batch_rows(si: ScanIndex, di: ScanIndex, pairs: torch.Tensor, max_points: int) -> int:
    """Rows to pad this batch to: min(max_points, largest cluster of the batch rounded up to PAD_QUANTUM)."""
    max_points = int(max_points)
    if not ADAPTIVE_PAD or len(pairs) == 0:
        return max_points
    ph = _pairs_i64(pairs, si.points.device).cpu().numpy()
    top = 1
    for col, idx in ((0, si), (1, di)):
        lab = ph[:, col]
        ok = (lab >= 0) & (lab < idx.n_labels)
        if ok.any():
            top = max(top, int(idx.counts_host[lab[ok]].max()))
    return min(max_points, (top + PAD_QUANTUM - 1) // PAD_QUANTUM * PAD_QUANTUM)


def sanity_check_cross(args, si: ScanIndex, di: ScanIndex, src_list: torch.Tensor, dst_list: torch.Tensor) -> torch.Tensor:
    """``sanity_check`` over every pair of two label lists (the dynamic stage, utils_match.py:43-51), src-major like the
    reference's repeat_interleave / repeat enumeration, without materialising the candidates: kept pairs ``[K,2]`` int64."""
    dev = si.points.device
    n_s, n_d = int(len(src_list)), int(len(dst_list))
    out = torch.empty(max(n_s * n_d, 1), 2, device=dev, dtype=torch.int64)
    count = torch.zeros(1, device=dev, dtype=torch.int32)
    lists = torch.cat([src_list.to(device=dev, dtype=torch.int64), dst_list.to(device=dev, dtype=torch.int64)]).contiguous()
    with torch.cuda.device(dev):
        code = _lib.lib().icpf_sanity_check_cross_f32(_ptr(si.offsets), _ptr(si.stats), si.n_labels, _ptr(di.offsets),
                                                      _ptr(di.stats), di.n_labels, _ptr(lists), n_s, n_d,
                                                      int(args.min_cluster_size), float(args.translation_frame),
                                                      float(args.thres_box), _ptr(out), _ptr(count), _stream_ptr())
    _lib.check(code, "icpf_sanity_check_cross_f32")
    return out[:int(count.item())]


def _match_pairs_indexed(args, si: ScanIndex, di: ScanIndex, pairs, src_unq, dst_unq, return_left: bool = False):
    segs_src, segs_dst = pad_pairs(si, di, pairs, batch_rows(si, di, pairs, args.max_points))
    transformations = hist_icp(args, segs_src, segs_dst)
    *evals, accept = match_eval(args, segs_src, segs_dst, transformations, return_accept=True)
    return match_select(args, _pairs_i64(pairs, si.points.device), src_unq, dst_unq, evals, accept, transformations,
                        return_left=return_left)


def match_pairs(args, src_points, dst_points, src_labels, dst_labels, pairs):
    """Drop-in for ``utils_match.match_pairs`` (utils_match.py:69-135): gather + pad the candidate cluster pairs (one
    launch on the scan indices), register them (``hist_icp``), score + gate them (``match_eval`` with the fused
    ``check_transformation``) and select one dst cluster per src cluster.  Returns ``(rows [K,10], transformations
    [K,4,4])`` on the device of the inputs."""
    assert len(pairs) > 0
    si, di = scan_index(src_points, src_labels), scan_index(dst_points, dst_labels)
    return _match_pairs_indexed(args, si, di, pairs, torch.unique(src_labels), torch.unique(dst_labels))


def match_pcds(args, src_points, dst_points, src_labels, dst_labels):
    """Drop-in for ``utils_match.match_pcds`` (utils_match.py:26-66): static stage (every label against itself), then
    the dynamic stage (every unmatched src cluster against every unmatched dst cluster), each filtered by
    ``sanity_check`` and resolved by ``match_pairs``.  Returns ``(pairs_matched [K,10], transformations [K,4,4])``.

    Candidate pairs with a negative label are never enumerated (the reference enumerates them and ``sanity_check`` drops
    them, utils_check.py:32); the order of the surviving pairs is the reference's."""
    si, di = scan_index(src_points, src_labels), scan_index(dst_points, dst_labels)
    dev = si.points.device
    src_unq, dst_unq = si.present_labels(), di.present_labels()                   # sorted, >= 0
    empty = (torch.zeros(0, 10, device=dev), torch.zeros(0, 4, 4, device=dev))

    # stage 1 (utils_match.py:32-41): overlapped clusters keep their label
    # (torch.unique(torch.cat([src_unq, dst_unq])) through the two presence tables: no device sort)
    n_both = max(si.n_labels, di.n_labels)
    present = torch.zeros(n_both, device=dev, dtype=torch.bool)
    present[:si.n_labels] |= si.present_mask
    present[:di.n_labels] |= di.present_mask
    both = torch.nonzero(present).flatten()
    pairs_true, _ = sanity_check_indexed(args, si, di, torch.stack([both, both], dim=1))
    if len(pairs_true) > 0:
        rows_sta, T_sta, src_left, dst_left = _match_pairs_indexed(args, si, di, pairs_true, src_unq, dst_unq, return_left=True)
    else:
        (rows_sta, T_sta), src_left, dst_left = empty, src_unq, dst_unq

    # stage 2 (utils_match.py:43-57): what is left, all against all
    rows_dyn, T_dyn = empty
    if len(src_left) > 0 and len(dst_left) > 0:
        pairs_true = sanity_check_cross(args, si, di, src_left, dst_left)
        if len(pairs_true) > 0:
            rows_dyn, T_dyn = _match_pairs_indexed(args, si, di, pairs_true, src_left, dst_left)
    return torch.cat([rows_sta, rows_dyn], dim=0), torch.cat([T_sta, T_dyn], dim=0)


def flow_estimation_torch(args, src_points, dst_points, src_labels, dst_labels, pairs, transformations, pose):
    """Drop-in for ``utils_flow.flow_estimation_torch`` (utils_flow.py:57-69): per src point
    ``(T_cluster(label) @ pose) p - p`` with the identity for labels without a matched pair; ``[n,3]`` fp32.
    ``pairs`` are the ``[K,10]`` rows of ``match_pcds`` (column 0 = src label)."""
    pts = _require_cuda_f32(src_points, "src_points")
    assert len(src_points) == len(src_labels)
    if pts.dim() != 2 or pts.shape[1] < 3:
        raise ValueError("src_points must be [n, >=3]")
    dev = pts.device
    labels = src_labels.to(device=dev, dtype=torch.float32).contiguous()
    K = int(len(pairs))
    prow = pairs.to(device=dev, dtype=torch.float32).contiguous() if K > 0 else None
    T = transformations.to(device=dev, dtype=torch.float32).contiguous() if K > 0 else None
    if K > 0 and T.shape != (K, 4, 4):
        raise ValueError("transformations must be [K,4,4], one per row of pairs")
    pose_d = None if pose is None else torch.as_tensor(pose).to(device=dev, dtype=torch.float32).contiguous()
    flow = torch.empty(len(pts), 3, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        code = _lib.lib().icpf_flow_f32(_ptr(pts), int(pts.shape[1]), _ptr(labels), len(pts), _ptr(prow),
                                        int(prow.shape[1]) if K > 0 else 1, _ptr(T), K, _ptr(pose_d), _ptr(flow),
                                        _stream_ptr())
    _lib.check(code, "icpf_flow_f32")
    return flow


def flow_estimation(args, src_points, dst_points, src_labels, dst_labels, pairs, transformations, pose,
                    device: Optional[torch.device] = None):
    """numpy front end with the signature of ``utils_flow.flow_estimation`` (utils_flow.py:23-50; main.py:230): the
    arrays are staged to the GPU, the flow is computed by the same fp32 kernel and returned as a float64 array (the
    reference evaluates this variant in fp64; the difference is fp32 rounding of the coordinates, ~4e-6 m at 50 m)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    f = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a), dtype=np.float32)).to(dev)
    flow = flow_estimation_torch(args, f(src_points)[:, 0:3].contiguous(), None, f(src_labels), None,
                                 f(np.asarray(pairs).reshape(-1, np.asarray(pairs).shape[-1] if np.ndim(pairs) == 2 else 10)),
                                 f(np.asarray(transformations).reshape(-1, 4, 4)), f(pose))
    return flow.cpu().numpy().astype(np.float64)
