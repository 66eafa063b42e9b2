"""CPU: install() rebinds exactly the reference's hot-path callables (names from SURVEY.md section 8b) and uninstall()
restores them.  Needs the reference sources, so it only runs in the build container."""
import pytest

from oracle import ref_loader


def test_install_rebinds_reference_seams():
    if not ref_loader.reference_available():
        pytest.skip("/root/reference not present (GPU box)")
    ref = ref_loader.load_reference()
    import icp_flow_b200
    from icp_flow_b200 import ops

    orig = (ref.utils_match.hist_icp, ref.utils_hist.estimate_init_pose, ref.utils_icp.apply_icp,
            ref.utils_icp_pytorch3d.iterative_closest_point, ref.utils_helper.nearest_neighbor_batch)
    try:
        mods = icp_flow_b200.install()
        assert {"utils_match", "utils_hist", "utils_icp", "utils_icp_pytorch3d"} <= set(mods)
        assert ref.utils_match.hist_icp is ops.hist_icp
        assert ref.utils_match.estimate_init_pose is ops.estimate_init_pose      # re-imported name inside utils_match
        assert ref.utils_hist.estimate_init_pose is ops.estimate_init_pose
        assert ref.utils_icp.apply_icp is ops.apply_icp
        assert ref.utils_icp.pytorch3d_icp is ops.pytorch3d_icp
        assert ref.utils_icp_pytorch3d.iterative_closest_point is ops.iterative_closest_point
        assert ref.utils_helper.nearest_neighbor_batch is orig[4]                 # helpers only on request
        # scan-level seams (rows f1-f3): pair enumeration, candidate filter, batch construction, flow recovery
        assert ref.utils_match.match_pcds is icp_flow_b200.match_pcds
        assert ref.utils_match.match_pairs is icp_flow_b200.match_pairs
        assert ref.utils_match.sanity_check is icp_flow_b200.sanity_check        # re-imported name inside utils_match
        assert ref.utils_check.sanity_check is icp_flow_b200.sanity_check
        assert ref.utils_flow.flow_estimation_torch is icp_flow_b200.flow_estimation_torch
        icp_flow_b200.install(patch_helpers=True)
        assert ref.utils_match.nearest_neighbor_batch is ops.nearest_neighbor_batch
    finally:
        icp_flow_b200.uninstall()
    assert (ref.utils_match.hist_icp, ref.utils_hist.estimate_init_pose, ref.utils_icp.apply_icp,
            ref.utils_icp_pytorch3d.iterative_closest_point, ref.utils_helper.nearest_neighbor_batch) == orig


def test_signatures_match_the_reference():
    """Same parameter names, order and defaults as the callables being replaced."""
    if not ref_loader.reference_available():
        pytest.skip("/root/reference not present (GPU box)")
    import inspect
    ref = ref_loader.load_reference()
    from icp_flow_b200 import ops

    def params(f, n=None):
        ps = list(inspect.signature(f).parameters.values())
        return [(p.name, p.default) for p in (ps if n is None else ps[:n])]

    assert params(ops.iterative_closest_point) == params(ref.utils_icp_pytorch3d.iterative_closest_point)
    assert params(ops.hist_icp, 3) == params(ref.utils_match.hist_icp)
    assert params(ops.estimate_init_pose, 3) == params(ref.utils_hist.estimate_init_pose)
    assert params(ops.apply_icp, 4) == params(ref.utils_icp.apply_icp)
    assert params(ops.pytorch3d_icp) == params(ref.utils_icp.pytorch3d_icp)
    assert params(ops.nearest_neighbor_batch) == params(ref.utils_helper.nearest_neighbor_batch)
    assert params(ops.transform_points_batch) == params(ref.utils_helper.transform_points_batch)
    import icp_flow_b200 as E
    assert params(E.sanity_check) == params(ref.utils_check.sanity_check)
    assert params(E.match_pairs) == params(ref.utils_match.match_pairs)
    assert params(E.match_pcds) == params(ref.utils_match.match_pcds)
    assert params(E.match_eval, 4) == params(ref.utils_match.match_eval)
    assert params(E.flow_estimation_torch) == params(ref.utils_flow.flow_estimation_torch)
    assert params(E.flow_estimation, 8) == params(ref.utils_flow.flow_estimation)
    import hist_cuda.hist as ref_hist          # the stub keeps the reference's signature (hist_cuda/hist.py:39)
    assert [p for p, _ in params(ops.hist)] == [p for p, _ in params(ref_hist.hist)]
