"""ICP pre-alignment (the reference's default icp_register_first=True).  VTK is absent:
the oracle restates vtkIterativeClosestPointTransform + vtkLandmarkTransform ("parity unpinned",
oracle/icp_port.py header)."""
import numpy as np
import pytest

import pyfocusr_b200.mesh as fmesh
from oracle import icp_port as ip


def _rot(rv):
    from scipy.spatial.transform import Rotation

    return Rotation.from_rotvec(rv).as_matrix()


# --------------------------------------------------------------------------------------------- CPU
def test_oracle_pieces():
    rng = np.random.RandomState(0)
    p = rng.randn(60, 3)
    r, t = _rot([0.4, -0.3, 0.9]), np.array([1.0, 2.0, -3.0])
    m = ip.landmark_transform(p, p @ r.T + t)
    assert np.allclose(m[:3, :3], r, atol=1e-13) and np.allclose(m[:3, 3], t, atol=1e-13)
    m = ip.landmark_transform(p, 1.7 * (p @ r.T) + t, similarity=True)
    assert np.allclose(m[:3, :3], 1.7 * r, atol=1e-13) and np.allclose(m[:3, 3], t, atol=1e-12)
    # closest point on a triangle: never farther than any sampled point of the triangle, and on the triangle
    a, b, c = rng.randn(3, 3)
    u = rng.dirichlet([0.3, 0.3, 0.3], 5000) @ np.stack([a, b, c])
    for _ in range(100):
        q = 2.0 * rng.randn(3)
        cp, d2 = ip.closest_points_on_triangles(q, a, b, c)
        assert d2 <= np.min(np.sum((u - q) ** 2, axis=1)) + 1e-12
        w = np.linalg.lstsq(np.stack([a, b, c]).T, cp, rcond=None)[0]
        assert abs(w.sum() - 1.0) < 1e-9 and w.min() > -1e-9
    # the surface distance goes down with the iteration count
    tm = fmesh.perturbed_ellipsoid(6, 0, semi_axes=(43.0, 25.0, 33.0))
    src = tm.points @ _rot([0.1, -0.15, 0.08]).T + np.array([3.0, -2.0, 1.0])

    def surf(pts):
        cp = ip.closest_points_on_mesh(pts[::3], tm.points, tm.tris)
        return float(np.mean(np.linalg.norm(cp - pts[::3], axis=1)))

    d = [surf(ip.apply_transform(src, ip.icp_transform(tm.points, tm.tris, src, max_iterations=k, max_landmarks=100)))
         for k in (1, 10, 40)]
    assert d[0] > d[1] > d[2]


# --------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def torch():
    import torch as t

    if not t.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return t


@pytest.mark.gpu
@pytest.mark.parametrize("similarity", [False, True])
def test_icp_matches_oracle(torch, similarity):
    from pyfocusr_b200 import _device

    tm = fmesh.perturbed_ellipsoid(8, 0, semi_axes=(43.0, 25.0, 33.0))
    sm = fmesh.perturbed_ellipsoid(8, 1, semi_axes=(43.0, 25.0, 33.0))
    src = (1.08 if similarity else 1.0) * (sm.points @ _rot([0.12, -0.1, 0.2]).T) + np.array([4.0, -3.0, 2.0])
    for iters, lm, tol in ((1, 1000, 1e-11), (12, 100, 1e-9), (40, 57, 1e-7)):
        ref = ip.icp_transform(tm.points, tm.tris, src, max_iterations=iters, max_landmarks=lm, similarity=similarity)
        mat, moved = _device.icp(tm.points, tm.tris, src, max_iterations=iters, max_landmarks=lm, similarity=similarity)
        # closest points are unique, the fit is a closed form: fp64 agreement, loosening slowly with the horizon
        assert np.max(np.abs(mat - ref)) <= tol * 50.0, (iters, np.max(np.abs(mat - ref)))
        assert np.max(np.abs(moved.cpu().numpy() - ip.apply_transform(src, ref))) <= tol * 50.0
        assert np.array_equal(mat[3], [0.0, 0.0, 0.0, 1.0])
    ref0 = ip.icp_transform(tm.points, tm.tris, src, max_iterations=3, max_landmarks=80, similarity=similarity,
                            start_by_matching_centroids=False)
    mat0, _ = _device.icp(tm.points, tm.tris, src, max_iterations=3, max_landmarks=80, similarity=similarity,
                          start_by_matching_centroids=False)
    assert np.max(np.abs(mat0 - ref0)) <= 1e-9


@pytest.mark.gpu
def test_focusr_literal_defaults(torch):
    """`Focusr(target, source)` with nothing else -- ICP, curvature features, CPD -- runs without VTK or cycpd."""
    import pyfocusr_b200 as pyfocusr

    mt = fmesh.perturbed_ellipsoid(8, 0, semi_axes=(43.0, 25.0, 33.0))
    ms = fmesh.perturbed_ellipsoid(8, 1, semi_axes=(43.0, 25.0, 33.0))
    moved = pyfocusr.PolyData(ms.points @ _rot([0.05, 0.02, -0.04]).T + np.array([2.0, -1.0, 1.5]), ms.tris)
    np.random.seed(0)
    f = pyfocusr.Focusr(mt, moved, rigid_reg_max_iterations=5, non_rigid_max_iterations=5)
    ref = ip.icp_transform(mt.points, mt.tris, moved.points, max_iterations=100, max_landmarks=1000)
    assert np.max(np.abs(f._icp_transform.matrix - ref)) <= 1e-6
    assert f._icp_transform.GetMatrix().GetElement(3, 3) == 1.0
    assert np.max(np.abs(f.graph_source.points - ip.apply_transform(moved.points, ref))) <= 1e-6

    def surf(pts):
        cp = ip.closest_points_on_mesh(pts, mt.points, mt.tris)
        return float(np.mean(np.linalg.norm(cp - pts, axis=1)))

    assert surf(f.graph_source.points) < 0.5 * surf(moved.points)          # the pre-alignment pulled the source onto the target
    f.align_maps()
    assert f.corresponding_target_idx_for_each_source_pt.shape == (ms.points.shape[0],)
    assert f.graph_target.n_extra_features == 2
    with pytest.raises(TypeError):
        pyfocusr.Focusr(mt, moved, icp_registration_mode="affine")
