"""oracle/ref_loader.py -- TEST INFRASTRUCTURE.  Runs the UNMODIFIED reference Python on CPU.

The reference hot-path files (/root/reference/utils_match.py, utils_hist.py, utils_icp.py,
utils_icp_pytorch3d.py, utils_helper.py) import third-party packages that are absent from this image
(pytorch3d 0.7.4, open3d, seaborn, matplotlib, plotly, parmap, torchist, hdbscan) and a CUDA-only JIT
extension (hist_cuda).  ``load_reference()`` pre-seeds ``sys.modules`` with

* functional stubs for the leaves that carry arithmetic (``pytorch3d.ops.knn_points``,
  ``pytorch3d.ops.utils.{wmean,eyes,is_pointclouds,convert_pointclouds_to_tensor}``,
  ``pytorch3d.transforms.matrix_to_euler_angles``, ``hist_cuda.hist.hist``) -- each restating the
  published behaviour of the pinned third-party version, and
* inert modules for everything that is only imported for plotting,

then imports the reference modules from ``/root/reference`` through ``sys.path`` at run time.  Nothing
from the reference is copied into this repository.  This only works in the build container
(/root/reference is absent on the GPU box); it is used by ``oracle/gen_golden.py`` to produce the
committed fixtures under ``tests/golden/`` and by the CPU tests that pin ``icp_oracle.py`` to the
reference.
"""
from __future__ import annotations

import collections
import importlib
import os
import sys
import types

import torch

from . import leaves

REFERENCE_ROOT = os.environ.get("ICPF_REFERENCE_ROOT", "/root/reference")

_KNN = collections.namedtuple("_KNN", ["dists", "idx", "knn"])


# --------------------------------------------------------------------------- pytorch3d.ops leaves
def _knn_points(p1, p2, lengths1=None, lengths2=None, norm=2, K=1, version=-1,
                return_nn=False, return_sorted=True):
    """pytorch3d 0.7.4 ``knn_points`` for K=1, norm=2 (the only configuration the reference uses)."""
    if K != 1 or norm != 2:
        raise NotImplementedError("oracle stub covers K=1, norm=2 only")
    if p1.shape[0] != p2.shape[0]:
        raise ValueError("pts1 and pts2 must have the same batch dimension.")
    if p1.shape[2] != p2.shape[2]:
        raise ValueError("pts1 and pts2 must have the same point dimension.")
    d2, idx = leaves.knn1(p1, p2, lengths1, lengths2)
    d2 = d2.to(p1.dtype)[:, :, None]
    idx = idx[:, :, None]
    nn = None
    if return_nn:
        nn = torch.gather(p2, 1, idx.expand(-1, -1, p2.shape[2]))[:, :, None, :]
    return _KNN(dists=d2, idx=idx, knn=nn)


def _wmean(x, weight=None, dim=-2, keepdim=True, eps=1e-9):
    """pytorch3d 0.7.4 ``ops.utils.wmean``."""
    args = {"dim": dim, "keepdim": keepdim}
    if weight is None:
        return x.mean(**args)
    if any(xd != wd and xd != 1 and wd != 1 for xd, wd in zip(x.shape[-2::-1], weight.shape[::-1])):
        raise ValueError("wmean: weights are not compatible with the tensor")
    return (x * weight[..., None]).sum(**args) / weight[..., None].sum(**args).clamp(eps)


def _eyes(dim, N, device=None, dtype=torch.float32):
    return torch.eye(dim, device=device, dtype=dtype)[None].repeat(N, 1, 1)


def _is_pointclouds(pcl):
    return hasattr(pcl, "points_padded") and hasattr(pcl, "num_points_per_cloud")


def _convert_pointclouds_to_tensor(pcl):
    if _is_pointclouds(pcl):
        return pcl.points_padded(), pcl.num_points_per_cloud()
    if torch.is_tensor(pcl):
        num = pcl.shape[1] * torch.ones(pcl.shape[0], device=pcl.device, dtype=torch.int64)
        return pcl, num
    raise ValueError("The inputs X, Y should be either Pointclouds objects or tensors.")


def _matrix_to_euler_angles(matrix, convention):
    """pytorch3d 0.7.4 ``matrix_to_euler_angles`` for the one convention the reference uses ("ZYX",
    utils_match.py:184), restated from its published algorithm: central angle asin(-m20), the two outer angles by
    atan2 -- (atan2(m10, m00), asin(-m20), atan2(m21, m22))."""
    if convention != "ZYX":
        raise NotImplementedError("oracle stub covers convention='ZYX' only")
    z = torch.atan2(matrix[..., 1, 0], matrix[..., 0, 0])
    y = torch.asin(-matrix[..., 2, 0])
    x = torch.atan2(matrix[..., 2, 1], matrix[..., 2, 2])
    return torch.stack([z, y, x], dim=-1)


# --------------------------------------------------------------------------- hist_cuda leaf
def _hist(X, Y, min_x, min_y, min_z, max_x, max_y, max_z, len_x, len_y, len_z, mini_batch=8):
    """CPU stand-in for the CUDA-only ``HIST.hist`` (hist_cuda/hist.py:39-51)."""
    out = leaves.hist_votes(X, Y, (float(min_x), float(min_y), float(min_z)),
                            (float(max_x), float(max_y), float(max_z)), (len_x, len_y, len_z))
    return out.to(X.dtype)


class _Inert(types.ModuleType):
    """Module whose every attribute is another inert module / no-op callable."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        child = _Inert(self.__name__ + "." + name)
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):
        return None


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def install_stubs():
    """Seed ``sys.modules`` (idempotent)."""
    if "pytorch3d" in sys.modules and getattr(sys.modules["pytorch3d"], "_icpf_stub", False):
        return
    ops_utils = _module("pytorch3d.ops.utils", wmean=_wmean, eyes=_eyes, is_pointclouds=_is_pointclouds,
                        convert_pointclouds_to_tensor=_convert_pointclouds_to_tensor)
    ops = _module("pytorch3d.ops", knn_points=_knn_points, utils=ops_utils)
    s_utils = _module("pytorch3d.structures.utils", list_to_padded=None)
    structures = _module("pytorch3d.structures", utils=s_utils)
    transforms = _module("pytorch3d.transforms", matrix_to_euler_angles=_matrix_to_euler_angles)
    p3d = _module("pytorch3d", ops=ops, structures=structures, transforms=transforms, _icpf_stub=True)
    for m in (p3d, ops, ops_utils, structures, s_utils, transforms):
        sys.modules[m.__name__] = m

    hist_mod = _module("hist_cuda.hist", hist=_hist)
    hist_pkg = _module("hist_cuda", hist=hist_mod)
    hist_pkg.__path__ = []  # mark as package
    sys.modules["hist_cuda"] = hist_pkg
    sys.modules["hist_cuda.hist"] = hist_mod

    for name in ("open3d", "seaborn", "matplotlib", "matplotlib.pyplot", "plotly", "plotly.graph_objs",
                 "parmap", "torchist", "hdbscan"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _Inert(name)


def load_reference():
    """Return a namespace with the reference's hot-path modules, imported verbatim from REFERENCE_ROOT."""
    if not os.path.isdir(REFERENCE_ROOT):
        raise FileNotFoundError(f"{REFERENCE_ROOT} not present (expected on the GPU box)")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    names = ["utils_helper", "utils_icp_pytorch3d", "utils_icp", "utils_hist", "utils_match", "utils_flow",
             "utils_check"]
    return types.SimpleNamespace(**{n: importlib.import_module(n) for n in names})


def reference_available() -> bool:
    return os.path.isdir(REFERENCE_ROOT) and os.path.exists(os.path.join(REFERENCE_ROOT, "utils_match.py"))
