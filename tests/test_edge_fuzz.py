"""Edge-case sweep over small, ragged and degenerate batches on both engines (tests/engines.py): empty clouds, single
points, duplicated points, collinear clouds, far-away origins, N that is no multiple of anything.  What must hold
whatever the input: no crash and no out-of-bounds access (the emulator engine also runs under AddressSanitizer,
tests/simt/memcheck.sh), finite rigid outputs, and the three correspondence-search modes agree bit for bit.
Index work (the NN seam, the cluster index) is compared exactly with numpy."""
import os
import types

import numpy as np
import pytest
import torch

from engines import is_simt, put
from icp_flow_b200 import ops

pytestmark = [pytest.mark.usefixtures("engine"), pytest.mark.order_last]


def _batch(rng, P, N, kind):
    src = np.full((P, N, 4), 1e8, np.float32); src[..., 3] = 0
    dst = src.copy()
    for p in range(P):
        n_s, n_d = (int(rng.integers(0, N + 1)) for _ in range(2))
        if kind == "full":
            n_s = n_d = N
        elif kind == "one_empty" and p == 0:
            n_s = 0
        elif kind == "one_empty" and p == 1:
            n_d = 0
        centre = rng.uniform(-40, 40, 3) * (250.0 if kind == "far" else 1.0)
        a = rng.uniform(-1, 1, (n_s, 3)) * [1.5, 0.8, 0.6]
        if kind == "collinear":
            a[:, 1:] = 0
        if kind == "duplicates" and n_s > 0:
            a[:] = a[0]
        k = min(n_s, n_d)
        b = np.concatenate([a[:k], rng.uniform(-1, 1, (n_d - k, 3)) * [1.5, 0.8, 0.6]]) if n_d else np.zeros((0, 3))
        b = b + rng.uniform(-0.04, 0.04, 3) + rng.normal(0, 0.004, b.shape)
        src[p, :n_s, :3] = a + centre; src[p, :n_s, 3] = 1
        dst[p, :n_d, :3] = b + centre; dst[p, :n_d, 3] = 1
    return src, dst


SHAPES = [(1, 1), (2, 2), (3, 5), (4, 17), (3, 33), (5, 64), (2, 127), (3, 129), (2, 300)]
KINDS = ["ragged", "full", "one_empty", "far", "collinear", "duplicates"]


@pytest.mark.parametrize("kind", KINDS)
def test_icp_modes_agree_and_stay_finite(kind):
    rng = np.random.default_rng(KINDS.index(kind))
    for P, N in SHAPES:
        src, dst = _batch(rng, P, N, kind)
        outs = []
        for mode in (1, 2, 3):
            r = ops.icp_batch(put(src), put(dst), ops.make_params(max_iterations=40, relative_rmse_thr=1e-6, nn_mode=mode))
            outs.append(r)
            R, T = r.R.cpu().numpy().astype(np.float64), r.T.cpu().numpy()
            assert np.isfinite(R).all() and np.isfinite(T).all(), (kind, P, N, mode)
            assert np.abs(R @ R.transpose(0, 2, 1) - np.eye(3)).max() < 1e-4, (kind, P, N, mode)
            assert (r.iterations.cpu().numpy() <= 40).all() and (r.iterations.cpu().numpy() >= 0).all()
        for b in outs[1:]:
            a = outs[0]
            assert torch.equal(a.R, b.R) and torch.equal(a.T, b.T) and torch.equal(a.iterations, b.iterations), (kind, P, N)
            assert torch.equal(a.conv_mask, b.conv_mask) and torch.equal(a.batch, b.batch), (kind, P, N)
            assert torch.equal(a.rmse.cpu().nan_to_num(-1.0), b.rmse.cpu().nan_to_num(-1.0)), (kind, P, N)


@pytest.mark.parametrize("kind", KINDS)
def test_hist_icp_and_match_eval_survive(kind):
    rng = np.random.default_rng(100 + KINDS.index(kind))
    args = types.SimpleNamespace(thres_dist=0.1, translation_frame=2.0, chunk_size=50, thres_iou=0.1, thres_rot=0.1)
    for P, N in SHAPES:
        src, dst = _batch(rng, P, N, kind)
        T = ops.hist_icp(args, put(src), put(dst))
        Tn = T.cpu().numpy()
        assert Tn.shape == (P, 4, 4) and np.isfinite(Tn).all(), (kind, P, N)
        assert np.array_equal(Tn[:, 3], np.tile(np.array([0, 0, 0, 1], np.float32), (P, 1))), (kind, P, N)
        ev = ops.match_eval(args, put(src), put(dst), put(T), return_accept=True)
        inl = ev[1].cpu().numpy()
        n_s, n_d = (src[:, :, 3] > 0).sum(1), (dst[:, :, 3] > 0).sum(1)
        assert (inl[:, 0] <= n_s).all() and (inl[:, 1] <= n_d).all() and (inl >= 0).all(), (kind, P, N)
        both = (n_s > 0) & (n_d > 0)
        assert np.isfinite(ev[0].cpu().numpy()[both]).all(), (kind, P, N)


def test_nn_seam_exact_on_ragged_shapes():
    rng = np.random.default_rng(7)
    for B, Ns, Nd in [(1, 1, 1), (2, 3, 7), (3, 130, 5), (2, 65, 257), (1, 1000, 31)]:
        a = rng.normal(0, 3, (B, Ns, 3)).astype(np.float32)
        b = rng.normal(0, 3, (B, Nd, 3)).astype(np.float32)
        b[:, Nd // 2] = b[:, 0]                                               # an exact tie: the lowest index wins
        idx, dist = ops.nearest_neighbor_batch(put(a), put(b))
        d = a[:, :, None, :] - b[:, None, :, :]
        d2 = d[..., 0] * d[..., 0]
        d2 = d2 + d[..., 1] * d[..., 1]
        d2 = d2 + d[..., 2] * d[..., 2]                                        # knn_points' fp32 op order
        want = d2.argmin(-1)
        assert np.array_equal(idx.cpu().numpy(), want), (B, Ns, Nd)
        np.testing.assert_allclose(dist.cpu().numpy(), np.sqrt(np.take_along_axis(d2, want[..., None], -1)[..., 0]), rtol=2e-7)


def test_scan_level_entry_points_on_empty_and_tiny_inputs():
    import icp_flow_b200 as E
    for n in (0, 1, 31, 257):
        pts = torch.randn(n, 3)
        lab = torch.randint(-1, 3, (n,)).float()
        idx = E.ScanIndex(put(pts), put(lab))
        order = idx.order.cpu().numpy()
        ln = lab.numpy()
        want = np.concatenate([np.nonzero(ln == k)[0] for k in range(idx.n_labels)]) if n else np.zeros(0, np.int64)
        assert np.array_equal(order[: len(want)], want), n
        flow = E.flow_estimation_torch(None, put(pts), None, put(lab), None, put(torch.zeros(0, 10)),
                                       put(torch.zeros(0, 4, 4)), put(torch.eye(4)))
        assert flow.shape == (n, 3) and float(flow.abs().sum()) == 0.0


def test_non_finite_points_do_not_spread_or_hang():
    """NaN / inf coordinates in a valid row (a LiDAR return gone wrong): no crash, no hang, no out-of-bounds access, and
    the OTHER pairs of the batch get exactly the result they get without the poisoned pair next to them (the batch
    stop aside: compared with a fixed iteration count)."""
    rng = np.random.default_rng(42)
    src, dst = _batch(rng, 6, 96, "full")
    clean = [ops.icp_batch(put(src[2:]), put(dst[2:]), ops.make_params(max_iterations=25, relative_rmse_thr=-1.0,
                                                                      early_exit=False, batch_stop=False, nn_mode=m))
             for m in (1, 3)]
    bad_s, bad_d = src.copy(), dst.copy()
    bad_s[0, 5, 0] = np.nan
    bad_s[0, 9, 2] = np.inf
    bad_d[1, 3, 1] = np.nan
    bad_d[1, 7, 0] = -np.inf
    for k, m in enumerate((1, 3)):
        r = ops.icp_batch(put(bad_s), put(bad_d), ops.make_params(max_iterations=25, relative_rmse_thr=-1.0,
                                                                  early_exit=False, batch_stop=False, nn_mode=m))
        assert torch.equal(r.R[2:], clean[k].R) and torch.equal(r.T[2:], clean[k].T), m
    args = types.SimpleNamespace(thres_dist=0.1, translation_frame=2.0, chunk_size=50, thres_iou=0.1, thres_rot=0.1)
    T = ops.hist_icp(args, put(bad_s), put(bad_d))
    assert T.shape == (6, 4, 4) and bool(torch.isfinite(T[2:]).all())
    ev = ops.match_eval(args, put(bad_s), put(bad_d), put(T))
    assert bool(torch.isfinite(ev[0][2:]).all())


def test_expand_rows_rebuilds_the_padded_batch_bit_for_bit():
    """The compact host format (valid xyz rows + CSR offsets, `ops.compact_rows`) expanded on the device is the padded
    batch pad_segment builds (utils_helper.py:185-196): same rows, flag 1, then (1e8, 1e8, 1e8, 0)."""
    rng = np.random.default_rng(11)
    for P, N in [(1, 1), (3, 5), (4, 64), (7, 129), (2, 512)]:
        src, _ = _batch(rng, P, N, "ragged")
        rows, offsets = ops.compact_rows(torch.from_numpy(src))
        assert rows.shape[0] == int((src[..., 3] > 0).sum()) and offsets[-1] == rows.shape[0]
        if rows.shape[0] == 0:
            rows = torch.zeros(1, 3)
        got = ops.expand_rows(put(rows), put(offsets), N).cpu().numpy()
        assert got.tobytes() == src.tobytes(), (P, N)


def test_packed_compact_batch_views_round_trip():
    """ops.pack_compact: both clouds' compact rows and offsets in one buffer; the views carved from a copy of the buffer
    expand to the padded batches bit for bit."""
    rng = np.random.default_rng(12)
    src, dst = _batch(rng, 5, 97, "ragged")
    if is_simt():
        import unittest.mock as mock
        with mock.patch.object(torch.Tensor, "pin_memory", lambda self: self):
            pk = ops.pack_compact(torch.from_numpy(src), torch.from_numpy(dst))
    else:
        pk = ops.pack_compact(torch.from_numpy(src), torch.from_numpy(dst))
    v = pk.views(put(pk.buffer.clone()))
    assert ops.expand_rows(v[0], v[1], 97).cpu().numpy().tobytes() == src.tobytes()
    assert ops.expand_rows(v[2], v[3], 97).cpu().numpy().tobytes() == dst.tobytes()
