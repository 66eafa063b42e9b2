"""Host-side mirror of the reference's operator interface for the per-cluster-pair registration path.

Same names, argument meaning and error behaviour as the reference callables they replace; every function hands
raw device pointers to the C ABI (include/icpflow_b200.h) on torch's current CUDA stream.  PyTorch is used for
device memory and streams only.

    iterative_closest_point   /root/reference/utils_icp_pytorch3d.py:37-225
    nearest_neighbor_batch    /root/reference/utils_helper.py:20-30
    transform_points_batch    /root/reference/utils_helper.py:76-87
"""
from __future__ import annotations

import ctypes
from typing import List, NamedTuple, Optional, Union

import torch

from . import _lib


class SimilarityTransform(NamedTuple):
    """utils_icp_pytorch3d.py:23-26"""
    R: torch.Tensor
    T: torch.Tensor
    s: torch.Tensor


class ICPSolution(NamedTuple):
    """utils_icp_pytorch3d.py:29-34"""
    converged: bool
    rmse: Union[torch.Tensor, None]
    Xt: torch.Tensor
    RTs: SimilarityTransform
    t_history: List[SimilarityTransform]


class IcpBatchResult(NamedTuple):
    """Raw outputs of one ``icpf_icp_f32`` call (device tensors, no host sync)."""
    R: torch.Tensor           # [P,3,3] row-vector convention  X R + T
    T: torch.Tensor           # [P,3]
    rmse: torch.Tensor        # [P]
    pose: torch.Tensor        # [P,4,4] column-convention [[R^T, T],[0,1]] (utils_icp.py:60-65)
    iterations: torch.Tensor  # [P] int32 iterations each pair executed
    conv_mask: torch.Tensor   # [P,4] int32 bit k <=> relative rmse <= thr at iteration k
    batch: torch.Tensor       # [2] int32 {batch iterations of the reference loop, converged}


def _stream_ptr() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _require_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not torch.is_tensor(t):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: icp_flow_b200 has no CPU implementation")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (got {t.dtype})")
    return t.contiguous()


def make_params(thres: float = 0.1, max_iterations: int = 100, relative_rmse_thr: float = 1e-6,
                early_exit: bool = True, batch_stop: bool = True, nn_mode: int = 0) -> _lib.IcpfParams:
    p = _lib.default_params()
    p.thres_dist = float(thres)
    p.max_iterations = int(max_iterations)
    p.relative_rmse_thr = float(relative_rmse_thr)
    p.early_exit = int(bool(early_exit))
    p.batch_stop = int(bool(batch_stop))
    p.nn_mode = int(nn_mode)
    return p


def icp_ext(peer_pose_dev: int = 0, peer_world: int = 0, peer_row0: int = 0, start_event: int = 0,
            stop_event: int = 0) -> _lib.IcpfIcpExt:
    """Per-call extensions of ``icp_batch`` (``struct icpf_icp_ext``): the fused peer all-gather of the transforms and /
    or a pair of CUDA events recorded around the dominant kernel.  Plain values: the library keeps no state."""
    e = _lib.IcpfIcpExt()
    e.peer_pose_dev, e.peer_world, e.peer_row0 = peer_pose_dev or None, int(peer_world), int(peer_row0)
    e.start_event, e.stop_event = start_event or None, stop_event or None
    return e


def icp_batch(src: torch.Tensor, dst: torch.Tensor, params: _lib.IcpfParams,
              init_R: Optional[torch.Tensor] = None, init_T: Optional[torch.Tensor] = None,
              out: Optional[IcpBatchResult] = None, workspace: Optional[torch.Tensor] = None,
              ext: Optional[_lib.IcpfIcpExt] = None) -> IcpBatchResult:
    """Stream-ordered batched ICP on padded ``[P,N,4]`` CUDA tensors; no host synchronisation."""
    src = _require_cuda_f32(src, "src")
    dst = _require_cuda_f32(dst, "dst")
    if src.dim() != 3 or dst.dim() != 3 or src.shape[2] != 4 or dst.shape[2] != 4:
        raise ValueError("src and dst must be [P, N, 4] (x, y, z, flag)")
    if src.shape[0] != dst.shape[0]:
        raise ValueError("Point sets X and Y have to have the same number of batches and data dimensions.")
    if src.shape[1] != dst.shape[1]:
        raise ValueError("src and dst must be padded to the same number of rows")
    P, N, _ = src.shape
    dev = src.device
    if out is None:
        out = IcpBatchResult(
            R=torch.empty(P, 3, 3, device=dev, dtype=torch.float32),
            T=torch.empty(P, 3, device=dev, dtype=torch.float32),
            rmse=torch.empty(P, device=dev, dtype=torch.float32),
            pose=torch.empty(P, 4, 4, device=dev, dtype=torch.float32),
            iterations=torch.empty(P, device=dev, dtype=torch.int32),
            conv_mask=torch.empty(P, 4, device=dev, dtype=torch.int32),
            batch=torch.empty(2, device=dev, dtype=torch.int32),
        )
    L = _lib.lib()
    need = L.icpf_workspace_bytes(P, N, 0, 0, 0)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(max(need, 1), device=dev, dtype=torch.uint8)
    if init_R is not None:
        init_R = _require_cuda_f32(init_R, "init_R")
        init_T = _require_cuda_f32(init_T, "init_T")
    with torch.cuda.device(dev):
        code = L.icpf_icp_ex_f32(_ptr(src), _ptr(dst), _ptr(init_R), _ptr(init_T), P, N, ctypes.byref(params),
                                 _ptr(out.R), _ptr(out.T), _ptr(out.rmse), _ptr(out.pose), _ptr(out.iterations),
                                 _ptr(out.conv_mask),
                                 _ptr(out.batch), _ptr(workspace), workspace.numel(), _stream_ptr(),
                                 ctypes.byref(ext) if ext is not None else None)
    _lib.check(code, "icpf_icp_ex_f32")
    return out


def icp_stats(workspace: torch.Tensor, P: int) -> torch.Tensor:
    """Diagnostics the last ``icp_batch`` call left in its workspace: ``[P,2]`` int32
    {full grid searches executed, cache-refresh iterations} per pair (layout: csrc/icpf_internal.h)."""
    up = lambda v: (v + 255) // 256 * 256
    off = up(P * 4) + up(P * 16) + 256
    return workspace[off:off + P * 8].view(torch.int32).view(P, 2)


def iterative_closest_point(X, Y, init_transform: Optional[SimilarityTransform] = None, thres: float = 0.1,
                            max_iterations: int = 100, relative_rmse_thr: float = 1e-6,
                            estimate_scale: bool = False, allow_reflection: bool = False,
                            verbose: bool = False) -> ICPSolution:
    """Drop-in for ``utils_icp_pytorch3d.iterative_closest_point`` on padded ``[P,N,4]`` CUDA tensors.

    The reference's batch-coupled stopping rule is reproduced on the device.  ``t_history`` has one entry per batch
    iteration like the reference's, but only the last entry (the returned transform) is materialised; earlier
    entries are ``None``.  ``estimate_scale`` / ``allow_reflection`` are not used anywhere on the reference path
    (utils_icp.py:56-57 passes False) and raise ``NotImplementedError`` when set.
    """
    if estimate_scale or allow_reflection:
        raise NotImplementedError("the ICP-Flow path calls ICP with estimate_scale=False, allow_reflection=False")
    if not torch.is_tensor(X) or not torch.is_tensor(Y):
        raise ValueError("The inputs X, Y should be padded [P,N,4] tensors.")
    if (X.shape[2] != Y.shape[2]) or (X.shape[0] != Y.shape[0]):
        raise ValueError("Point sets X and Y have to have the same number of batches and data dimensions.")
    b = X.shape[0]
    init_R = init_T = None
    if init_transform is not None:
        try:
            R0, T0, s0 = init_transform
            assert R0.shape == torch.Size((b, 3, 3)) and T0.shape == torch.Size((b, 3)) and s0.shape == torch.Size((b,))
        except Exception:
            raise ValueError(
                "The initial transformation init_transform has to be a named tuple SimilarityTransform with "
                "elements (R, T, s). R are dim x dim orthonormal matrices of shape (minibatch, dim, dim), T is a "
                "batch of dim-dimensional translations of shape (minibatch, dim) and s is a batch of scalars of "
                "shape (minibatch,).") from None
        init_R = (s0[:, None, None] * R0).float()
        init_T = T0.float()
    params = make_params(thres, max_iterations, relative_rmse_thr, early_exit=True, batch_stop=True)
    r = icp_batch(X, Y, params, init_R, init_T)
    batch = r.batch.tolist()           # the one host sync of this wrapper (the reference syncs every iteration)
    iterations, converged = int(batch[0]), bool(batch[1])
    if verbose:
        print(f"ICP has converged in {iterations} iterations." if converged
              else f"ICP has not converged in {max_iterations} iterations.")
    s = r.T.new_ones(b)
    Xt = (X[:, :, 0:3, None] * r.R[:, None, :, :]).sum(dim=2) + r.T[:, None, :]
    final = SimilarityTransform(r.R, r.T, s)
    history: List[Optional[SimilarityTransform]] = [None] * (iterations - 1) + [final]
    return ICPSolution(converged, r.rmse, Xt, final, history)


def nearest_neighbor_batch(src: torch.Tensor, dst: torch.Tensor):
    """Drop-in for ``utils_helper.nearest_neighbor_batch``: unbounded K=1 NN over ALL rows -> (idx int64, dist)."""
    assert src.dim() == 3
    assert dst.dim() == 3
    assert len(src) == len(dst)
    assert src.shape[2] >= 3
    assert dst.shape[2] >= 3
    b, num, _ = src.shape

    def rows(t, name):
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor: icp_flow_b200 has no CPU implementation")
        t = t.float()
        # accept [B,N,3] / [B,N,4] contiguous storage directly, anything else is compacted to xyz
        if t.is_contiguous() and t.shape[2] in (3, 4):
            return t, t.shape[2]
        return t[:, :, 0:3].contiguous(), 3

    s, ss = rows(src, "src")
    d, ds = rows(dst, "dst")
    idx = torch.empty(b, num, device=s.device, dtype=torch.int64)
    dist = torch.empty(b, num, device=s.device, dtype=torch.float32)
    with torch.cuda.device(s.device):
        code = _lib.lib().icpf_nn_f32(_ptr(s), _ptr(d), b, num, d.shape[1], ss, ds, _ptr(idx), _ptr(dist), _stream_ptr())
    _lib.check(code, "icpf_nn_f32")
    return idx.view(b, num), dist.view(b, num)


def transform_points_batch(xyz: torch.Tensor, pose: torch.Tensor) -> torch.Tensor:
    """Drop-in for ``utils_helper.transform_points_batch``: ``[x y z 1] @ pose^T`` keeping the flag column."""
    assert xyz.dim() == 3
    assert pose.dim() == 3
    assert xyz.shape[2] == 4
    assert pose.shape[1] == 4
    assert pose.shape[2] == 4
    assert len(xyz) == len(pose)
    x = _require_cuda_f32(xyz, "xyz")
    m = _require_cuda_f32(pose, "pose")
    out = torch.empty_like(x)
    b, n, _ = x.shape
    with torch.cuda.device(x.device):
        code = _lib.lib().icpf_transform_points_f32(_ptr(x), _ptr(m), b, n, _ptr(out), _stream_ptr())
    _lib.check(code, "icpf_transform_points_f32")
    return out


def host_kabsch(H: torch.Tensor, sequence: bool = False) -> torch.Tensor:
    """CPU test hook: the kernels' closed-form Kabsch rotation for ``[n,3,3]`` cross-covariances (row convention).
    ``sequence=True`` solves them in order with the warm start an ICP run uses between iterations."""
    h = H.detach().to(torch.float32).contiguous().cpu()
    out = torch.empty_like(h)
    fn = _lib.lib().icpf_host_kabsch_sequence if sequence else _lib.lib().icpf_host_kabsch
    fn(ctypes.c_void_p(h.data_ptr()), h.shape[0], ctypes.c_void_p(out.data_ptr()))
    return out


# ----------------------------------------------------------------------------------------------------------------
# Histogram initialisation, apply_icp and hist_icp: the remaining seams of the path (same signatures as the reference)
# ----------------------------------------------------------------------------------------------------------------
class _HistBins:
    """Bin starts built exactly like utils_hist.py:60-65 (torch.arange in the default dtype) + the C struct."""

    def __init__(self, thres_dist: float, translation_frame: float, device):
        eps = 1e-8
        f, tau = translation_frame, thres_dist
        self.x = torch.arange(-f, f + tau - eps, tau, dtype=torch.float32).to(device)
        self.y = torch.arange(-f, f + tau - eps, tau, dtype=torch.float32).to(device)
        self.z = torch.arange(-tau, tau + tau - eps, tau, dtype=torch.float32).to(device)
        lo = [float(self.x.min()), float(self.y.min()), float(self.z.min())]
        hi = [float(self.x.max()), float(self.y.max()), float(self.z.max())]
        self.lens = (len(self.x), len(self.y), len(self.z))
        c = _lib.IcpfHistBins()
        c.bins_x, c.bins_y, c.bins_z = self.x.data_ptr(), self.y.data_ptr(), self.z.data_ptr()
        for k in range(3):
            c.len[k] = self.lens[k]
            c.min[k] = lo[k]
            c.max[k] = hi[k]
        c.half_bin = float(thres_dist // 2)        # utils_hist.py:78: python floor division
        self.c = c


_BINS_CACHE = {}


def _hist_bins(thres_dist, translation_frame, device) -> _HistBins:
    key = (float(thres_dist), float(translation_frame), str(device))
    hb = _BINS_CACHE.get(key)
    if hb is None:
        if len(_BINS_CACHE) > 64:
            _BINS_CACHE.clear()
        hb = _BINS_CACHE[key] = _HistBins(float(thres_dist), float(translation_frame), device)
    return hb


def _workspace(P, N, lens, device):
    need = _lib.lib().icpf_workspace_bytes(P, N, lens[0], lens[1], lens[2])
    return torch.empty(max(int(need), 1), device=device, dtype=torch.uint8)


def _check_pair_batch(src, dst):
    src = _require_cuda_f32(src, "src")
    dst = _require_cuda_f32(dst, "dst")
    assert len(src) == len(dst)
    if src.dim() != 3 or dst.dim() != 3 or src.shape[2] != 4 or dst.shape[2] != 4 or src.shape[1] != dst.shape[1]:
        raise ValueError("src and dst must be [P, N, 4] (x, y, z, flag) padded to the same N")
    return src, dst


def hist(X, Y, min_x, min_y, min_z, max_x, max_y, max_z, len_x, len_y, len_z, mini_batch=8):
    """Drop-in for ``hist_cuda.hist.hist`` (hist_cuda/hist.py:39-51): ``[B,len_x,len_y,len_z]`` fp32 vote counts of
    ``X[b,i,:3] - Y[b,j,:3]`` over rows whose flags are both > 0.  ``mini_batch`` is accepted and ignored (the
    reference needed it to keep its flat thread index inside an int)."""
    if not X.is_contiguous() or not Y.is_contiguous():
        raise RuntimeError("input tensor has to be contiguous")
    if not X.is_cuda or not Y.is_cuda:
        raise RuntimeError("input must be a CUDA tensor")
    if X.size(0) != Y.size(0):
        raise RuntimeError(f"batch_X ({X.size(0)}) != batch_Y ({Y.size(0)}).")
    if X.size(2) != Y.size(2):
        raise RuntimeError(f"dim_X ({X.size(2)}) != dim_Y ({Y.size(2)}).")
    if X.size(2) != 4:
        raise RuntimeError(f"dim ({X.size(2)}) != 4; 3 for (x, y, z); 1 for indicator,padded or not.")
    X, Y = X.float(), Y.float()
    B = X.size(0)
    out = torch.empty(B, int(len_x), int(len_y), int(len_z), device=X.device, dtype=torch.float32)
    mins = (ctypes.c_float * 3)(float(min_x), float(min_y), float(min_z))
    maxs = (ctypes.c_float * 3)(float(max_x), float(max_y), float(max_z))
    lens = (ctypes.c_int32 * 3)(int(len_x), int(len_y), int(len_z))
    with torch.cuda.device(X.device):
        code = _lib.lib().icpf_hist_votes_f32(_ptr(X), _ptr(Y), B, X.size(1), Y.size(1), mins, maxs, lens, _ptr(out),
                                              _stream_ptr())
    _lib.check(code, "icpf_hist_votes_f32")
    return out


def estimate_init_pose(args, src, dst, return_debug: bool = False, auto_swap: bool = False):
    """Drop-in for ``utils_hist.estimate_init_pose(args, src, dst) -> [P,4,4]`` (reads ``args.thres_dist`` and
    ``args.translation_frame``; ``args.chunk_size`` only bounded the reference's memory and is not needed)."""
    src, dst = _check_pair_batch(src, dst)
    P, N, _ = src.shape
    dev = src.device
    hb = _hist_bins(args.thres_dist, args.translation_frame, dev)
    pose = torch.empty(P, 4, 4, device=dev, dtype=torch.float32)
    dbg = None
    if return_debug:
        dbg = {"flat_idx": torch.empty(P, 5, device=dev, dtype=torch.int32),
               "votes": torch.empty(P, 5, device=dev, dtype=torch.float32),
               "scores": torch.empty(P, 6, device=dev, dtype=torch.float32),
               "which": torch.empty(P, device=dev, dtype=torch.int32)}
    ws = _workspace(P, N, hb.lens, dev)
    with torch.cuda.device(dev):
        code = _lib.lib().icpf_hist_init_f32(
            _ptr(src), _ptr(dst), P, N, ctypes.byref(hb.c), int(auto_swap), _ptr(pose),
            _ptr(dbg["flat_idx"]) if dbg else None, _ptr(dbg["votes"]) if dbg else None,
            _ptr(dbg["scores"]) if dbg else None, _ptr(dbg["which"]) if dbg else None,
            _ptr(ws), ws.numel(), _stream_ptr())
    _lib.check(code, "icpf_hist_init_f32")
    return (pose, dbg) if return_debug else pose


def _path_params(args) -> _lib.IcpfParams:
    # utils_icp.py:52-58: thres=args.thres_dist, max_iterations=100, relative_rmse_thr=1e-6
    return make_params(thres=args.thres_dist, max_iterations=100, relative_rmse_thr=1e-6, early_exit=True,
                       batch_stop=True)


def apply_icp(args, src, dst, init_poses, return_debug: bool = False, auto_swap: bool = False):
    """Drop-in for ``utils_icp.apply_icp(args, src, dst, init_poses) -> [P,4,4]``."""
    src, dst = _check_pair_batch(src, dst)
    init = _require_cuda_f32(init_poses, "init_poses")
    P, N, _ = src.shape
    dev = src.device
    assert init.shape == (P, 4, 4)
    out = torch.empty(P, 4, 4, device=dev, dtype=torch.float32)
    err = torch.empty(P, 2, device=dev, dtype=torch.float32) if return_debug else None
    flags = torch.empty(P, device=dev, dtype=torch.int32) if return_debug else None
    batch = torch.empty(2, device=dev, dtype=torch.int32) if return_debug else None
    ws = _workspace(P, N, (0, 0, 0), dev)
    params = _path_params(args)
    with torch.cuda.device(dev):
        code = _lib.lib().icpf_apply_icp_f32(_ptr(src), _ptr(dst), _ptr(init), P, N, ctypes.byref(params),
                                             int(auto_swap), _ptr(out), _ptr(err), _ptr(flags), _ptr(batch), _ptr(ws),
                                             ws.numel(), _stream_ptr())
    _lib.check(code, "icpf_apply_icp_f32")
    if return_debug:
        return out, {"errors": err, "flags": flags, "batch": batch}
    return out


class ApplyIcpPhases:
    """``apply_icp`` split at the two points where the reference's batch stop couples the pairs of a call
    (utils_icp_pytorch3d.py:209), for a batch that is spread over several devices (``shard.hist_icp_sharded``):

        ph = ApplyIcpPhases(args, src, dst, init_poses, auto_swap=True)
        words = ph.first_pass()              # [4] int32: AND of this shard's convergence masks
        ... AND the words of all shards; lowest set bit below ph.cap = the batch stop ...
        words = ph.full_pass()               # only if no stop below the cap and ph.cap < ph.max_iterations
        T = ph.finish(iterations, converged)

    With one shard the result is exactly ``apply_icp``'s.  A shard without pairs makes no native call and reports
    all-ones words."""

    def __init__(self, args, src, dst, init_poses, auto_swap: bool = False):
        self.src, self.dst = _check_pair_batch(src, dst)
        self.init = _require_cuda_f32(init_poses, "init_poses")
        self.P, self.N, _ = self.src.shape
        assert self.init.shape == (self.P, 4, 4)
        self.dev = self.src.device
        self.params = _path_params(args)
        self.auto_swap = int(auto_swap)
        self.max_iterations = int(self.params.max_iterations)
        self.cap = min(32, self.max_iterations) if self.params.early_exit else self.max_iterations
        self.ws = _workspace(self.P, self.N, (0, 0, 0), self.dev)
        self.words = torch.full((4,), -1, device=self.dev, dtype=torch.int32)
        self.out = torch.empty(self.P, 4, 4, device=self.dev, dtype=torch.float32)
        self.batch = torch.empty(2, device=self.dev, dtype=torch.int32)

    def _call(self, phase: int, iterations: int = 0, converged: int = 0):
        if self.P == 0:
            return
        with torch.cuda.device(self.dev):
            code = _lib.lib().icpf_apply_icp_phase_f32(
                _ptr(self.src), _ptr(self.dst), _ptr(self.init), self.P, self.N, ctypes.byref(self.params),
                self.auto_swap, phase, int(iterations), int(converged), _ptr(self.words), _ptr(self.out), None, None,
                _ptr(self.batch), _ptr(self.ws), self.ws.numel(), _stream_ptr())
        _lib.check(code, "icpf_apply_icp_phase_f32")

    def first_pass(self) -> torch.Tensor:
        self._call(0)
        return self.words

    def full_pass(self) -> torch.Tensor:
        self._call(1)
        return self.words

    def finish(self, iterations: int, converged: bool) -> torch.Tensor:
        self._call(2, iterations, int(bool(converged)))
        return self.out


def pytorch3d_icp(args, src, dst):
    """Drop-in for ``utils_icp.pytorch3d_icp(args, src, dst) -> [P,4,4]`` (utils_icp.py:50-73)."""
    src, dst = _check_pair_batch(src, dst)
    return icp_batch(src, dst, _path_params(args)).pose


def hist_icp(args, src, dst, return_debug: bool = False):
    """Drop-in for ``utils_match.hist_icp(args, src, dst) -> [P,4,4]``: the whole per-cluster-pair path (swap so the
    smaller cloud moves, histogram init, ICP with roll-back, un-swap) in one stream-ordered native call."""
    src, dst = _check_pair_batch(src, dst)
    P, N, _ = src.shape
    dev = src.device
    hb = _hist_bins(args.thres_dist, args.translation_frame, dev)
    out = torch.empty(P, 4, 4, device=dev, dtype=torch.float32)
    init = torch.empty(P, 4, 4, device=dev, dtype=torch.float32) if return_debug else None
    batch = torch.empty(2, device=dev, dtype=torch.int32) if return_debug else None
    ws = _workspace(P, N, hb.lens, dev)
    params = _path_params(args)
    with torch.cuda.device(dev):
        code = _lib.lib().icpf_hist_icp_f32(_ptr(src), _ptr(dst), P, N, ctypes.byref(hb.c), ctypes.byref(params),
                                            _ptr(out), _ptr(init), _ptr(batch), _ptr(ws), ws.numel(), _stream_ptr())
    _lib.check(code, "icpf_hist_icp_f32")
    if return_debug:
        return out, {"init": init, "batch": batch}
    return out


def match_eval(args, pcd1, pcd2, transformations, return_accept: bool = False):
    """Drop-in for ``utils_match.match_eval(args, pcd1, pcd2, transformations)`` (utils_match.py:159-213): returns
    ``(errors [P,2], inliers [P,2], ratios [P,2], ious [P,2], translations [P,3], rotations [P,3])`` of the registrations
    ``transformations`` (``[P,4,4]``, pcd1 -> pcd2) from one fused launch.  ``return_accept`` appends the ``[P]`` int32
    verdict of ``utils_check.check_transformation`` (needs args.translation_frame / thres_iou / thres_rot)."""
    pcd1, pcd2 = _check_pair_batch(pcd1, pcd2)
    P, N, _ = pcd1.shape
    dev = pcd1.device
    if transformations.shape != (P, 4, 4) or transformations.device != dev:
        raise ValueError("transformations must be [P,4,4] on the device of the clouds")
    pose = transformations.to(torch.float32).contiguous()
    errors, inliers, ratios, ious = (torch.empty(P, 2, device=dev, dtype=torch.float32) for _ in range(4))
    translations, rotations = (torch.empty(P, 3, device=dev, dtype=torch.float32) for _ in range(2))
    gates, accept = None, None
    if return_accept:
        gates = ctypes.byref(_lib.IcpfMatchGates(float(args.translation_frame), float(args.thres_iou),
                                                 float(args.thres_rot)))
        accept = torch.empty(P, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        code = _lib.lib().icpf_match_eval_f32(_ptr(pcd1), _ptr(pcd2), _ptr(pose), P, N, float(args.thres_dist),
                                              _ptr(errors), _ptr(inliers), _ptr(ratios), _ptr(ious),
                                              _ptr(translations), _ptr(rotations), gates, _ptr(accept), _stream_ptr())
    _lib.check(code, "icpf_match_eval_f32")
    if return_accept:
        return errors, inliers, ratios, ious, translations, rotations, accept
    return errors, inliers, ratios, ious, translations, rotations


def match_select(args, pairs, src_labels_unq, dst_labels_unq, evals, accept, transformations, return_left: bool = False):
    """The rejection loop and selection of match_pairs (utils_match.py:70-75, 94-135) as ONE launch
    (``icpf_match_select_f32``): per src cluster, among its accepted registrations, the dst cluster of least min(error)
    (``match_segments_descend`` on the reference's ``[n_src, n_dst]`` matrix), kept if that error < ``thres_error``.
    Returns ``(rows [K,10], transformations [K,4,4])`` in src-label order; with ``return_left`` also the sorted src / dst
    labels that stay unmatched (the candidates of match_pcds' dynamic stage).  Device tensors in, device tensors out; the
    one host sync is the read of the three data-dependent output lengths."""
    errors, inliers, ratios, ious = (_require_cuda_f32(x, "evals").contiguous() for x in evals[0:4])
    dev = errors.device
    P = int(errors.shape[0])
    ns, nd = int(len(src_labels_unq)), int(len(dst_labels_unq))
    p64 = pairs.to(device=dev, dtype=torch.int64).contiguous()
    su = src_labels_unq.to(device=dev, dtype=torch.int64).contiguous()
    du = dst_labels_unq.to(device=dev, dtype=torch.int64).contiguous()
    acc = accept.to(device=dev, dtype=torch.int32).contiguous()
    T = _require_cuda_f32(transformations, "transformations").contiguous()
    rows = torch.empty(max(ns, 1), 10, device=dev, dtype=torch.float32)
    T_out = torch.empty(max(ns, 1), 4, 4, device=dev, dtype=torch.float32)
    src_left = torch.empty(max(ns, 1), device=dev, dtype=torch.int64)
    dst_left = torch.empty(max(nd, 1), device=dev, dtype=torch.int64)
    counts = torch.empty(3, device=dev, dtype=torch.int32)
    L = _lib.lib()
    ws = torch.empty(int(L.icpf_match_select_workspace_bytes(ns, nd)) // 8 + 1, device=dev, dtype=torch.int64)
    with torch.cuda.device(dev):
        code = L.icpf_match_select_f32(_ptr(p64), P, _ptr(su), ns, _ptr(du), nd, _ptr(errors), _ptr(inliers), _ptr(ratios),
                                       _ptr(ious), _ptr(acc), _ptr(T), float(args.thres_error), _ptr(rows), _ptr(T_out),
                                       _ptr(src_left), _ptr(dst_left), _ptr(counts), _ptr(ws), ws.numel() * 8, _stream_ptr())
    _lib.check(code, "icpf_match_select_f32")
    k, n_sl, n_dl = (int(v) for v in counts.cpu().tolist())        # data-dependent output lengths: the one host sync
    if return_left:
        return rows[:k], T_out[:k], src_left[:n_sl], dst_left[:n_dl]
    return rows[:k], T_out[:k]


def expand_rows(rows: torch.Tensor, offsets: torch.Tensor, max_points: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Compact clusters -> the reference's padded batch (``pad_segment``, utils_helper.py:185-196) on the device.

    ``rows`` ``[total,3]`` fp32 xyz of the valid rows, cluster after cluster; ``offsets`` ``[B+1]`` int32 (CSR).  Returns
    ``[B, max_points, 4]``: ``(x,y,z,1)`` rows then ``(1e8,1e8,1e8,0)`` padding."""
    rows = _require_cuda_f32(rows, "rows")
    if offsets.dtype != torch.int32 or not offsets.is_cuda:
        raise TypeError("offsets must be a CUDA int32 tensor [B+1]")
    B = offsets.numel() - 1
    if out is None:
        out = torch.empty(B, int(max_points), 4, device=rows.device, dtype=torch.float32)
    with torch.cuda.device(rows.device):
        code = _lib.lib().icpf_expand_rows_f32(_ptr(rows), _ptr(offsets.contiguous()), B, int(max_points), _ptr(out), _stream_ptr())
    _lib.check(code, "icpf_expand_rows_f32")
    return out


def compact_rows(padded: torch.Tensor):
    """Host-side inverse of ``expand_rows`` for a padded ``[B,N,4]`` CPU tensor: (rows ``[total,3]``, offsets ``[B+1]`` int32).
    What a caller that builds its clusters on the host ships instead of the padded batch (12 B per valid row)."""
    flags = padded[:, :, 3] > 0
    counts = flags.sum(dim=1).to(torch.int32)
    offsets = torch.zeros(padded.shape[0] + 1, dtype=torch.int32)
    offsets[1:] = torch.cumsum(counts, dim=0)
    return padded[:, :, :3][flags].contiguous(), offsets


class PackedCompactBatch:
    """``compact_rows`` of both clouds in one pinned host buffer: [src rows | dst rows | src offsets | dst offsets], every
    block 256-byte aligned.  ``views(buf)`` carves the same blocks out of a device copy of the buffer."""

    def __init__(self, src_rows, src_offsets, dst_rows, dst_offsets):
        parts = [src_rows.contiguous().view(-1).view(torch.uint8), dst_rows.contiguous().view(-1).view(torch.uint8),
                 src_offsets.contiguous().view(torch.uint8), dst_offsets.contiguous().view(torch.uint8)]
        self.shapes = [(src_rows.shape[0], 3), (dst_rows.shape[0], 3), (src_offsets.numel(),), (dst_offsets.numel(),)]
        self.offsets, total = [], 0
        for p in parts:
            self.offsets.append(total)
            total += (p.numel() + 255) // 256 * 256
        self.buffer = torch.empty(total, dtype=torch.uint8).pin_memory()
        for p, o in zip(parts, self.offsets):
            self.buffer[o:o + p.numel()].copy_(p)
        self.payload_bytes = sum(p.numel() for p in parts)

    def views(self, buf: torch.Tensor):
        o, sh = self.offsets, self.shapes
        return (buf[o[0]:o[0] + sh[0][0] * 12].view(torch.float32).view(sh[0][0], 3),
                buf[o[2]:o[2] + sh[2][0] * 4].view(torch.int32),
                buf[o[1]:o[1] + sh[1][0] * 12].view(torch.float32).view(sh[1][0], 3),
                buf[o[3]:o[3] + sh[3][0] * 4].view(torch.int32))


def pack_compact(src_padded: torch.Tensor, dst_padded: torch.Tensor) -> PackedCompactBatch:
    """Host side: padded ``[P,N,4]`` CPU batches -> the single-buffer compact format (``IcpHostPipeline.submit_packed``)."""
    return PackedCompactBatch(*compact_rows(src_padded), *compact_rows(dst_padded))


class IcpHostPipeline:
    """Host-buffer front end of the ICP stage: pinned host inputs -> H2D -> kernels -> D2H of the 4x4 transforms.

    Steps are double-buffered: the H2D copy of step k+1 (copy stream) overlaps the kernels of step k (compute stream),
    so a sequence of batches runs at max(PCIe, kernel) instead of their sum.  ``submit`` never blocks the host; call
    ``synchronize`` (or wait on the returned event) before reading the output buffer.

    Two host formats: the reference's padded ``[P,N,4]`` batches (``submit``), or the compact one (``submit_compact``:
    xyz of the valid rows + CSR offsets, 12 bytes per valid row instead of 16 per padded row -- the path is PCIe-bound,
    so the bytes are the time), expanded to the padded layout by one small kernel on the device.
    """

    def __init__(self, num_pairs: int, max_points: int, params: _lib.IcpfParams, device=None):
        self.P, self.N, self.params = int(num_pairs), int(max_points), params
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        with torch.cuda.device(self.dev):
            self.copy_stream = torch.cuda.Stream()
            self.compute_stream = torch.cuda.Stream()
            self.d_src = [torch.empty(self.P, self.N, 4, device=self.dev) for _ in range(2)]
            self.d_dst = [torch.empty(self.P, self.N, 4, device=self.dev) for _ in range(2)]
            self.c_rows = [[None, None], [None, None]]       # compact staging (allocated on first use)
            self.c_offs = [[torch.empty(self.P + 1, device=self.dev, dtype=torch.int32) for _ in range(2)] for _ in range(2)]
            self.copied = [torch.cuda.Event() for _ in range(2)]
            self.consumed = [torch.cuda.Event() for _ in range(2)]
            self.outs = [None, None]
            self.ws = torch.empty(_lib.lib().icpf_workspace_bytes(self.P, self.N, 0, 0, 0), device=self.dev,
                                  dtype=torch.uint8)
        self.k = 0

    def submit(self, host_src: torch.Tensor, host_dst: torch.Tensor, host_pose_out: torch.Tensor, ext=None):
        """host_src / host_dst: pinned ``[P,N,4]`` fp32; host_pose_out: pinned ``[P,4,4]`` fp32 (written async)."""
        s = self.k & 1
        with torch.cuda.device(self.dev):
            with torch.cuda.stream(self.copy_stream):
                if self.k >= 2:
                    self.copy_stream.wait_event(self.consumed[s])       # the kernels of step k-2 released this slot
                self.d_src[s].copy_(host_src, non_blocking=True)
                self.d_dst[s].copy_(host_dst, non_blocking=True)
                self.copied[s].record(self.copy_stream)
            with torch.cuda.stream(self.compute_stream):
                self.compute_stream.wait_event(self.copied[s])
                self.outs[s] = icp_batch(self.d_src[s], self.d_dst[s], self.params, out=self.outs[s], workspace=self.ws, ext=ext)
                self.consumed[s].record(self.compute_stream)
                host_pose_out.copy_(self.outs[s].pose, non_blocking=True)
        self.k += 1
        return self.outs[s]

    def submit_compact(self, host_src_rows: torch.Tensor, host_src_offsets: torch.Tensor, host_dst_rows: torch.Tensor,
                       host_dst_offsets: torch.Tensor, host_pose_out: torch.Tensor, ext=None):
        """Compact host format (``compact_rows``): pinned ``[total,3]`` fp32 rows + ``[P+1]`` int32 offsets per cloud."""
        s = self.k & 1
        with torch.cuda.device(self.dev):
            with torch.cuda.stream(self.copy_stream):
                if self.k >= 2:
                    self.copy_stream.wait_event(self.consumed[s])
                for side, (rows, offs) in enumerate(((host_src_rows, host_src_offsets), (host_dst_rows, host_dst_offsets))):
                    buf = self.c_rows[s][side]
                    if buf is None or buf.shape[0] < rows.shape[0]:
                        buf = self.c_rows[s][side] = torch.empty(max(rows.shape[0], 1), 3, device=self.dev)
                    buf[:rows.shape[0]].copy_(rows, non_blocking=True)
                    self.c_offs[s][side].copy_(offs, non_blocking=True)
                self.copied[s].record(self.copy_stream)
            with torch.cuda.stream(self.compute_stream):
                self.compute_stream.wait_event(self.copied[s])
                expand_rows(self.c_rows[s][0], self.c_offs[s][0], self.N, out=self.d_src[s])
                expand_rows(self.c_rows[s][1], self.c_offs[s][1], self.N, out=self.d_dst[s])
                self.outs[s] = icp_batch(self.d_src[s], self.d_dst[s], self.params, out=self.outs[s], workspace=self.ws, ext=ext)
                self.consumed[s].record(self.compute_stream)
                host_pose_out.copy_(self.outs[s].pose, non_blocking=True)
        self.k += 1
        return self.outs[s]

    def submit_packed(self, packed: "PackedCompactBatch", host_pose_out: torch.Tensor, ext=None):
        """The compact format as ONE pinned buffer (``pack_compact``): a single H2D copy per step instead of four (the two
        offset arrays are 4 KB each and would cost a copy latency of their own)."""
        s = self.k & 1
        with torch.cuda.device(self.dev):
            with torch.cuda.stream(self.copy_stream):
                if self.k >= 2:
                    self.copy_stream.wait_event(self.consumed[s])
                buf = self.c_rows[s][0]
                if buf is None or buf.dtype != torch.uint8 or buf.numel() < packed.buffer.numel():
                    buf = self.c_rows[s][0] = torch.empty(packed.buffer.numel(), device=self.dev, dtype=torch.uint8)
                buf[:packed.buffer.numel()].copy_(packed.buffer, non_blocking=True)
                self.copied[s].record(self.copy_stream)
            with torch.cuda.stream(self.compute_stream):
                self.compute_stream.wait_event(self.copied[s])
                v = packed.views(buf)
                expand_rows(v[0], v[1], self.N, out=self.d_src[s])
                expand_rows(v[2], v[3], self.N, out=self.d_dst[s])
                self.outs[s] = icp_batch(self.d_src[s], self.d_dst[s], self.params, out=self.outs[s], workspace=self.ws, ext=ext)
                self.consumed[s].record(self.compute_stream)
                host_pose_out.copy_(self.outs[s].pose, non_blocking=True)
        self.k += 1
        return self.outs[s]

    def synchronize(self):
        self.copy_stream.synchronize()
        self.compute_stream.synchronize()
