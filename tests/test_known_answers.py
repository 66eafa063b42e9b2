"""Known-answer tests of the path's semantics (SURVEY.md section 7 lists them; the reference ships none).

Each case is checked on the CPU oracle, on the engine with `-m gpu`, and on the engine's kernels under the SIMT-on-CPU
emulator (tests/simt/) through the same helper, so all sides are held to the same analytic answers and not only to
each other.
"""
import types

import numpy as np
import pytest
import torch

import engines
from engines import put
from oracle import icp_oracle as O

TAU = 0.1


def _pad(points, N):
    out = np.full((N, 4), 1e8, np.float32)
    out[:, 3] = 0.0
    out[: len(points), :3] = points
    out[: len(points), 3] = 1.0
    return out


def _cloud(rng, n, centre=(20.0, -10.0, 1.0), extent=(2.0, 1.0, 1.2)):
    return (rng.uniform(-0.5, 0.5, size=(n, 3)) * np.array(extent) + np.array(centre)).astype(np.float32)


class OracleSide:
    name = "oracle"

    @staticmethod
    def hist_icp(src, dst, F):
        p = O.PathParams(thres_dist=TAU, translation_frame=F)
        T, dbg = O.hist_icp(torch.from_numpy(src), torch.from_numpy(dst), p, return_debug=True)
        return T.numpy(), dbg["init"].numpy(), dbg["rolled_back"].numpy()

    @staticmethod
    def icp(src, dst, iters=100):
        tr = O.icp_loop(torch.from_numpy(src), torch.from_numpy(dst), TAU, iters, 1e-6)
        return tr.R.numpy(), tr.T.numpy()


class EngineSide:
    name = "engine"

    @staticmethod
    def hist_icp(src, dst, F):
        from icp_flow_b200 import ops
        args = types.SimpleNamespace(thres_dist=TAU, translation_frame=F, chunk_size=50)
        s, d = put(torch.from_numpy(src)), put(torch.from_numpy(dst))
        T, dbg = ops.hist_icp(args, s, d, return_debug=True)
        # roll-back flags through the apply_icp seam on the swapped clouds
        n_s, n_d = (s[:, :, 3] > 0).sum(1), (d[:, :, 3] > 0).sum(1)
        sw = n_s > n_d
        a, c = s.clone(), d.clone()
        a[sw], c[sw] = d[sw], s[sw]
        _, adbg = ops.apply_icp(args, a, c, dbg["init"], return_debug=True)
        return T.cpu().numpy(), dbg["init"].cpu().numpy(), (adbg["flags"].cpu().numpy() & 1).astype(bool)

    @staticmethod
    def icp(src, dst, iters=100):
        from icp_flow_b200 import ops
        r = ops.icp_batch(put(torch.from_numpy(src)), put(torch.from_numpy(dst)),
                          ops.make_params(thres=TAU, max_iterations=iters))
        return r.R.cpu().numpy(), r.T.cpu().numpy()


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu), "simt"])
def side(request):
    """oracle: the CPU restatement; cuda: the engine on the GPU; simt: the engine's kernels under the SIMT-on-CPU
    emulator (tests/engines.py)."""
    if request.param == "oracle":
        yield OracleSide
    else:
        with engines.running(request.param):
            yield EngineSide


def test_identity_pair_gives_identity(side):
    """Identical clouds: R = I, T = 0 (to fp32), init translation exactly zero."""
    rng = np.random.default_rng(0)
    pts = _cloud(rng, 200)
    src = np.stack([_pad(pts, 256)])
    T, init, _ = side.hist_icp(src, src.copy(), 2.0)
    assert np.abs(init[0, :3, 3]).max() <= 1e-6
    assert np.abs(T[0] - np.eye(4)).max() < 2e-4           # |t| error is amplified by the 20 m lever arm
    moved = pts @ T[0, :3, :3].T + T[0, :3, 3]
    assert np.abs(moved - pts).max() < 2e-5


def test_pure_translation_on_the_bin_lattice_is_recovered(side):
    """dst = src + (0.7, -0.4, 0): the histogram peak decodes to that lattice point and ICP keeps it."""
    rng = np.random.default_rng(1)
    pts = _cloud(rng, 240)
    shift = np.array([0.7, -0.4, 0.0], np.float32)
    src = np.stack([_pad(pts, 256)])
    dst = np.stack([_pad(pts + shift, 256)])
    T, init, _ = side.hist_icp(src, dst, 2.0)
    assert np.abs(init[0, :3, 3] - shift).max() <= 0.1 + 1e-6           # within one bin of the lattice point
    moved = pts @ T[0, :3, :3].T + T[0, :3, 3]
    assert np.abs(moved - (pts + shift)).max() < 1e-3


def test_zero_inlier_pair_is_identity_icp_and_rolls_back(side):
    """Clouds 5 m apart in z (outside the histogram's z range and the ICP gate): no votes -> zero init translation,
    no inliers -> ICP returns the identity, error does not drop -> roll back to the init pose."""
    rng = np.random.default_rng(2)
    pts = _cloud(rng, 150)
    src = np.stack([_pad(pts, 256)])
    dst = np.stack([_pad(pts + np.array([0.0, 0.0, 5.0], np.float32), 256)])
    R, Tt = side.icp(src, dst)
    assert np.array_equal(R[0], np.eye(3, dtype=np.float32)) and np.array_equal(Tt[0], np.zeros(3, np.float32))
    T, init, rolled = side.hist_icp(src, dst, 2.0)
    assert np.abs(T[0] - init[0]).max() == 0.0 and bool(rolled[0])
    assert np.abs(T[0, :3, :3] - np.eye(3)).max() == 0.0


def test_padded_rows_do_not_matter(side):
    """The same clusters padded to 256 and to 640 rows give the same transforms."""
    rng = np.random.default_rng(3)
    a, b = _cloud(rng, 180), None
    ang = np.deg2rad(2.0)
    Rz = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], np.float32)
    c = a.mean(0)
    b = ((a - c) @ Rz.T + c + np.array([0.03, -0.02, 0.01], np.float32)).astype(np.float32)
    T1, _, _ = side.hist_icp(np.stack([_pad(a, 256)]), np.stack([_pad(b, 256)]), 2.0)
    T2, _, _ = side.hist_icp(np.stack([_pad(a, 640)]), np.stack([_pad(b, 640)]), 2.0)
    assert np.abs(T1 - T2).max() <= (0.0 if side.name == "engine" else 1e-6)
    moved = a @ T1[0, :3, :3].T + T1[0, :3, 3]
    assert np.abs(moved - b).max() < 5e-3


def test_swap_symmetry(side):
    """n_src > n_dst: the clouds are swapped internally and the result inverted (utils_match.py:139-154), so the
    transform still maps src onto dst."""
    rng = np.random.default_rng(4)
    big = _cloud(rng, 250)
    small = (big[:140] + np.array([0.2, 0.1, 0.0], np.float32)).astype(np.float32)
    T, _, _ = side.hist_icp(np.stack([_pad(big, 256)]), np.stack([_pad(small, 256)]), 2.0)
    moved = big[:140] @ T[0, :3, :3].T + T[0, :3, 3]
    assert np.abs(moved - small).max() < 2e-3
    assert np.array_equal(T[0, 3], np.array([0, 0, 0, 1], np.float32))


def test_planar_and_tiny_clusters_stay_finite_rigid(side):
    """Exactly planar cluster (rank-2 cross-covariance) and a 4-point cluster: finite proper rotations."""
    rng = np.random.default_rng(5)
    plane = _cloud(rng, 200, extent=(3.0, 2.0, 0.0))
    tiny = _cloud(rng, 4)
    src = np.stack([_pad(plane, 256), _pad(tiny, 256)])
    dst = np.stack([_pad(plane + np.array([0.02, 0.03, 0.0], np.float32), 256), _pad(tiny + 0.01, 256)])
    R, T = side.icp(src, dst)
    assert np.isfinite(R).all() and np.isfinite(T).all()
    for k in range(2):
        Rk = R[k].astype(np.float64)
        assert np.abs(Rk @ Rk.T - np.eye(3)).max() < 1e-5 and np.linalg.det(Rk) > 0.999
    moved = plane @ R[0] + T[0]
    assert np.abs(moved - (plane + np.array([0.02, 0.03, 0.0]))).max() < 1e-3


def test_histogram_bin_edges_oracle():
    """v == min is counted, v == max (the LAST BIN START the reference passes as max) is not; bin width is
    (max - min) / len, narrower than thres_dist.  A vote one ulp below max makes the
    reference compute bin index == len (an unguarded out-of-bounds write): oracle and engine clamp it to the last bin."""
    from oracle import leaves
    p = O.PathParams(thres_dist=0.1, translation_frame=2.0)
    bx, by, bz = O.bin_edges(p)
    assert len(bx) == 41 and len(bz) == 3
    mn, mx = float(bx.min()), float(bx.max())
    X = torch.tensor([[[mn, 0.0, 0.0, 1.0], [mx, 0.0, 0.0, 1.0], [np.nextafter(np.float32(mx), np.float32(0)), 0.0, 0.0, 1.0]]])
    Y = torch.tensor([[[0.0, 0.0, 0.0, 1.0]]])
    h = leaves.hist_votes(X, Y, (mn, float(by.min()), float(bz.min())), (mx, float(by.max()), float(bz.max())),
                          (len(bx), len(by), len(bz)))
    assert h.sum() == 2                                   # the v == max vote is dropped
    assert h[0, 0].sum() == 1 and h[0, len(bx) - 1].sum() == 1      # (v - min) / (max - min) * len == len -> clamped
    assert (mx - mn) / len(bx) < 0.1                      # kernel bin width < thres_dist
    assert 0.1 // 2 == 0.0                                # the "+ thres_dist // 2" of utils_hist.py:78 adds nothing


@pytest.mark.gpu
def test_vote_kernel_counts_equal_the_reference_cuda_kernel():
    """The reference's own `hist_cuda_kernel` (hist_cuda_core.cuh:23-64, built unmodified for sm_100a into oracle/_ref/ by
    oracle/Makefile) against icpf_hist_votes_f32 on the same inputs: every count of every bin identical."""
    import torch
    from icp_flow_b200 import ops, synth
    from oracle import ref_hist
    if not ref_hist.available():
        pytest.skip("oracle/_ref/libref_hist.so not built (needs /root/reference at build time)")
    dev = torch.device("cuda:0")
    for (P, N, F, seed) in [(9, 160, 2.0, 1), (20, 512, 6.666, 2), (5, 1024, 3.333, 3)]:
        src, dst, _ = synth.make_pairs(P, N, seed=seed, ragged=True, residual_only=False, wrong_frac=0.2)
        s, d = torch.from_numpy(src).to(dev), torch.from_numpy(dst).to(dev)
        hb = ops._hist_bins(0.1, F, dev)
        mine = ops.hist(d, s, *hb.c.min, *hb.c.max, *hb.lens)
        ref = ref_hist.hist(d, s, list(hb.c.min), list(hb.c.max), hb.lens)
        torch.cuda.synchronize()
        assert torch.equal(mine, ref), (P, N, F, int((mine != ref).sum()))
        assert float(ref.sum()) > 0
