"""ctypes binding of libicpflow_b200.so (the C ABI declared in include/icpflow_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``python -m icp_flow_b200.build``; there is no CPU
fallback -- if the shared object is missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ICPF_LIB_PATH: an experimental build of the same ABI (tools/ab_kernel.py); the product path is the in-tree library
LIB_PATH = os.environ.get("ICPF_LIB_PATH") or os.path.join(_HERE, "libicpflow_b200.so")

# every symbol include/icpflow_b200.h declares (checked by tests/test_abi.py against the header text)
EXPORTS = (
    "icpf_version", "icpf_error_string", "icpf_default_params", "icpf_workspace_bytes",
    "icpf_icp_f32", "icpf_icp_ex_f32", "icpf_nn_f32", "icpf_transform_points_f32", "icpf_host_kabsch",
    "icpf_host_kabsch_sequence", "icpf_peer_push_f32", "icpf_expand_rows_f32", "icpf_hist_votes_f32", "icpf_hist_init_f32", "icpf_apply_icp_f32", "icpf_apply_icp_phase_f32", "icpf_hist_icp_f32",
    "icpf_match_eval_f32",
    "icpf_cluster_index_workspace_bytes", "icpf_cluster_index_f32", "icpf_sanity_check_f32", "icpf_sanity_check_cross_f32",
    "icpf_match_select_workspace_bytes", "icpf_match_select_f32", "icpf_gather_pairs_f32",
    "icpf_hdbscan_labels_host", "icpf_hdbscan_workspace_bytes", "icpf_hdbscan_mst_f32",
    "icpf_flow_f32", "icpf_dbscan_workspace_bytes", "icpf_dbscan_f32",
)


class IcpfParams(ctypes.Structure):
    """Mirror of ``struct icpf_params``."""

    _fields_ = [
        ("thres_dist", ctypes.c_double),
        ("max_iterations", ctypes.c_int32),
        ("relative_rmse_thr", ctypes.c_float),
        ("early_exit", ctypes.c_int32),
        ("batch_stop", ctypes.c_int32),
        ("nn_mode", ctypes.c_int32),
        ("reserved", ctypes.c_int32 * 2),
    ]


class IcpfHistBins(ctypes.Structure):
    """Mirror of ``struct icpf_hist_bins``."""

    _fields_ = [
        ("bins_x", ctypes.c_void_p),
        ("bins_y", ctypes.c_void_p),
        ("bins_z", ctypes.c_void_p),
        ("len", ctypes.c_int32 * 3),
        ("min", ctypes.c_float * 3),
        ("max", ctypes.c_float * 3),
        ("half_bin", ctypes.c_float),
    ]


class IcpfMatchGates(ctypes.Structure):
    """Mirror of ``struct icpf_match_gates``."""

    _fields_ = [("translation_frame", ctypes.c_double), ("thres_iou", ctypes.c_double), ("thres_rot", ctypes.c_double)]


class IcpfIcpExt(ctypes.Structure):
    """Mirror of ``struct icpf_icp_ext`` (per-call extensions of the ICP loop: fused peer gather, timing events)."""

    _fields_ = [("peer_pose_dev", ctypes.c_void_p), ("peer_world", ctypes.c_int32), ("peer_row0", ctypes.c_int32),
                ("start_event", ctypes.c_void_p), ("stop_event", ctypes.c_void_p)]


class IcpfError(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        msg = lib().icpf_error_string(code).decode()
        super().__init__(f"{where} failed: {msg} (code {code})")


_LIB = None


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build the sm_100a extension first (python -c 'import __graft_entry__ as g; "
            "g.build()' at the repo root). icp_flow_b200 has no CPU or PyTorch fallback."
        )
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64p = ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p
    L.icpf_version.restype = ctypes.c_int
    L.icpf_version.argtypes = []
    L.icpf_error_string.restype = ctypes.c_char_p
    L.icpf_error_string.argtypes = [ctypes.c_int]
    L.icpf_default_params.restype = None
    L.icpf_default_params.argtypes = [ctypes.POINTER(IcpfParams)]
    L.icpf_workspace_bytes.restype = ctypes.c_size_t
    L.icpf_workspace_bytes.argtypes = [i32] * 5
    L.icpf_icp_f32.restype = ctypes.c_int
    L.icpf_icp_f32.argtypes = [vp, vp, vp, vp, i32, i32, ctypes.POINTER(IcpfParams), vp, vp, vp, vp, vp, vp, vp, vp,
                               ctypes.c_size_t, vp]
    L.icpf_nn_f32.restype = ctypes.c_int
    L.icpf_nn_f32.argtypes = [vp, vp, i32, i32, i32, i32, i32, i64p, vp, vp]
    L.icpf_transform_points_f32.restype = ctypes.c_int
    L.icpf_transform_points_f32.argtypes = [vp, vp, i32, i32, vp, vp]
    L.icpf_match_eval_f32.restype = ctypes.c_int
    L.icpf_match_eval_f32.argtypes = [vp, vp, vp, i32, i32, ctypes.c_double, vp, vp, vp, vp, vp, vp,
                                      ctypes.POINTER(IcpfMatchGates), vp, vp]
    f3, i3 = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32)
    L.icpf_hist_votes_f32.restype = ctypes.c_int
    L.icpf_hist_votes_f32.argtypes = [vp, vp, i32, i32, i32, f3, f3, i3, vp, vp]
    L.icpf_hist_init_f32.restype = ctypes.c_int
    L.icpf_hist_init_f32.argtypes = [vp, vp, i32, i32, ctypes.POINTER(IcpfHistBins), i32, vp, vp, vp, vp, vp, vp,
                                     ctypes.c_size_t, vp]
    L.icpf_apply_icp_f32.restype = ctypes.c_int
    L.icpf_apply_icp_f32.argtypes = [vp, vp, vp, i32, i32, ctypes.POINTER(IcpfParams), i32, vp, vp, vp, vp, vp,
                                     ctypes.c_size_t, vp]
    L.icpf_apply_icp_phase_f32.restype = ctypes.c_int
    L.icpf_apply_icp_phase_f32.argtypes = [vp, vp, vp, i32, i32, ctypes.POINTER(IcpfParams), i32, i32, i32, i32, vp, vp, vp,
                                           vp, vp, vp, ctypes.c_size_t, vp]
    L.icpf_hist_icp_f32.restype = ctypes.c_int
    L.icpf_hist_icp_f32.argtypes = [vp, vp, i32, i32, ctypes.POINTER(IcpfHistBins), ctypes.POINTER(IcpfParams), vp, vp,
                                    vp, vp, ctypes.c_size_t, vp]
    L.icpf_icp_ex_f32.restype = ctypes.c_int
    L.icpf_icp_ex_f32.argtypes = L.icpf_icp_f32.argtypes + [ctypes.POINTER(IcpfIcpExt)]
    L.icpf_peer_push_f32.restype = ctypes.c_int
    L.icpf_peer_push_f32.argtypes = [vp, vp, i32, i32, i32, vp]
    L.icpf_expand_rows_f32.restype = ctypes.c_int
    L.icpf_expand_rows_f32.argtypes = [vp, vp, i32, i32, vp, vp]
    L.icpf_host_kabsch.restype = None
    L.icpf_host_kabsch.argtypes = [vp, i32, vp]
    L.icpf_host_kabsch_sequence.restype = None
    L.icpf_host_kabsch_sequence.argtypes = [vp, i32, vp]
    L.icpf_cluster_index_workspace_bytes.restype = ctypes.c_size_t
    L.icpf_cluster_index_workspace_bytes.argtypes = [i32, i32]
    L.icpf_cluster_index_f32.restype = ctypes.c_int
    L.icpf_cluster_index_f32.argtypes = [vp, i32, vp, i32, i32, vp, vp, vp, vp, ctypes.c_size_t, vp]
    L.icpf_sanity_check_f32.restype = ctypes.c_int
    L.icpf_sanity_check_f32.argtypes = [vp, vp, i32, vp, vp, i32, vp, i32, i32, ctypes.c_double, ctypes.c_double, vp,
                                        vp, vp, vp]
    L.icpf_sanity_check_cross_f32.restype = ctypes.c_int
    L.icpf_sanity_check_cross_f32.argtypes = [vp, vp, i32, vp, vp, i32, vp, i32, i32, i32, ctypes.c_double, ctypes.c_double,
                                              vp, vp, vp]
    L.icpf_match_select_workspace_bytes.restype = ctypes.c_size_t
    L.icpf_match_select_workspace_bytes.argtypes = [i32, i32]
    L.icpf_match_select_f32.restype = ctypes.c_int
    L.icpf_match_select_f32.argtypes = [vp, i32, vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, ctypes.c_double, vp, vp, vp, vp, vp,
                                        vp, ctypes.c_size_t, vp]
    L.icpf_hdbscan_labels_host.restype = ctypes.c_int
    L.icpf_hdbscan_labels_host.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    L.icpf_hdbscan_workspace_bytes.restype = ctypes.c_size_t
    L.icpf_hdbscan_workspace_bytes.argtypes = [i32]
    L.icpf_hdbscan_mst_f32.restype = ctypes.c_int
    L.icpf_hdbscan_mst_f32.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
    L.icpf_gather_pairs_f32.restype = ctypes.c_int
    L.icpf_gather_pairs_f32.argtypes = [vp, i32, vp, vp, i32, vp, i32, vp, vp, i32, vp, i32, i32, vp, vp, vp, vp, vp]
    L.icpf_flow_f32.restype = ctypes.c_int
    L.icpf_flow_f32.argtypes = [vp, i32, vp, i32, vp, i32, vp, i32, vp, vp, vp]
    L.icpf_dbscan_workspace_bytes.restype = ctypes.c_size_t
    L.icpf_dbscan_workspace_bytes.argtypes = [i32]
    L.icpf_dbscan_f32.restype = ctypes.c_int
    L.icpf_dbscan_f32.argtypes = [vp, i32, i32, ctypes.c_double, i32, vp, vp, vp, ctypes.c_size_t, vp]
    _LIB = L
    return L


def check(code: int, where: str) -> None:
    if code != 0:
        raise IcpfError(code, where)


def default_params() -> IcpfParams:
    p = IcpfParams()
    lib().icpf_default_params(ctypes.byref(p))
    return p
