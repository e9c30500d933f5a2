"""Curvature node features (the reference's default list_features_to_calc=["curvature"]).
VTK is absent: the oracle restates vtkCurvatures ("parity unpinned", oracle/curvature_port.py header)."""
import numpy as np
import pytest

import pyfocusr_b200.mesh as fmesh
from oracle import curvature_port as cv
from oracle import port


def _damaged(seed=0, nu=8):
    rng = np.random.RandomState(seed)
    e = fmesh.perturbed_ellipsoid(nu, 1)
    t = e.tris.copy()
    keep = np.ones(len(t), bool)
    keep[rng.choice(len(t), 15, replace=False)] = False
    t = t[keep]
    fl = rng.choice(len(t), 10, replace=False)
    t[fl] = t[fl][:, [0, 2, 1]]
    t = np.concatenate([t, t[:3]])                      # duplicated faces: edges with two neighbours
    pts = np.concatenate([e.points, [[0.0, 0.0, 0.0]]])  # and an unreferenced vertex
    return pts, t


# --------------------------------------------------------------------------------------------- CPU
def test_oracle_vectorised_equals_loops_and_sphere_values():
    m = fmesh.icosphere(6, radius=2.0)
    a, b = cv.curvatures_loops(m.points, m.tris), cv.curvatures(m.points, m.tris)
    for k in a:
        assert np.max(np.abs(a[k] - b[k])) <= 1e-13
    assert np.all(np.abs(b["gauss"] - 0.25) < 0.04) and np.all(np.abs(b["mean"] - 0.5) < 0.08)   # 1/R^2, 1/R
    pts, t = _damaged()
    a, b = cv.curvatures_loops(pts, t), cv.curvatures(pts, t)
    for k in a:
        assert np.max(np.abs(a[k] - b[k])) <= 1e-12 * max(1.0, np.max(np.abs(a[k])))
    assert b["gauss"][-1] == 0.0 and b["mean"][-1] == 0.0                                            # unreferenced vertex
    # flat patch: zero curvature inside
    g = np.stack(np.meshgrid(np.arange(6.0), np.arange(6.0), indexing="ij"), -1).reshape(-1, 2)
    flat = np.concatenate([g, np.zeros((36, 1))], axis=1)
    quads = [(i * 6 + j, i * 6 + j + 1, (i + 1) * 6 + j + 1, (i + 1) * 6 + j) for i in range(5) for j in range(5)]
    tri = np.array([(a_, b_, c_) for a_, b_, c_, d_ in quads] + [(a_, c_, d_) for a_, b_, c_, d_ in quads])
    c = cv.curvatures(flat, tri)
    inner = [i * 6 + j for i in range(1, 5) for j in range(1, 5)]
    assert np.max(np.abs(c["gauss"][inner])) < 1e-12 and np.max(np.abs(c["mean"][inner])) < 1e-12


# --------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def torch():
    import torch as t

    if not t.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return t


def _check(out, ref, tol=1e-11):
    for k in ("gauss", "mean", "minimum", "maximum"):
        got, want = out[k].cpu().numpy(), ref[k]
        # min / max switch to 0 where H^2 - K changes sign: compare those only where the discriminant is clear
        if k in ("minimum", "maximum"):
            disc = ref["mean"] ** 2 - ref["gauss"]
            ok = np.abs(disc) > 1e-9 * (1.0 + np.abs(ref["gauss"]))
            got, want = got[ok], want[ok]
        # sqrt(H^2 - K) near 0 amplifies a 1e-15 rounding difference by 1 / (2 sqrt(1e-9)): 1e-10 for min / max
        assert np.max(np.abs(got - want) / (1.0 + np.abs(want))) <= (tol if k in ("gauss", "mean") else 1e-10), k


@pytest.mark.gpu
def test_curvatures_match_oracle(torch, shipped_meshes):
    from pyfocusr_b200 import _device

    for name in ("target_mesh", "source_mesh_15k"):
        m = shipped_meshes[name]
        _check(_device.curvatures(m.points, m.tris), cv.curvatures(m.points, m.tris))
    pts, t = _damaged()
    _check(_device.curvatures(pts, t), cv.curvatures(pts, t))
    # a batch with global vertex ids gives each mesh's own values
    a, b = fmesh.icosphere(5, radius=3.0), fmesh.perturbed_ellipsoid(7, 2)
    pts = np.concatenate([a.points, b.points])
    tris = np.concatenate([a.tris, b.tris + a.points.shape[0]])
    out = _device.curvatures(pts, tris)
    ra, rb = cv.curvatures(a.points, a.tris), cv.curvatures(b.points, b.tris)
    _check({k: v[: a.points.shape[0]] for k, v in out.items()}, ra)
    _check({k: v[a.points.shape[0]:] for k, v in out.items()}, rb)
    # deterministic: two runs are bitwise equal
    again = _device.curvatures(pts, tris)
    assert all(torch.equal(out[k], again[k]) for k in out)
    from pyfocusr_b200 import _lib

    bad = tris.copy()
    bad[0, 0] = pts.shape[0] + 5
    with pytest.raises(_lib.FocusrB200Error):
        _device.curvatures(pts, bad)


@pytest.mark.gpu
def test_graph_curvature_features_and_default_focusr(torch, shipped_meshes):
    """Graph(list_features_to_calc=["curvature"]) and the reference's literal default Focusr(...) call."""
    import pyfocusr_b200 as pyfocusr

    m = shipped_meshes["target_mesh"]
    np.random.seed(0)
    g = pyfocusr.Graph(m, n_spectral_features=3, list_features_to_calc=["curvature"], include_features_in_adj_matrix=True)
    assert g.n_extra_features == 2
    ref = [port.normalized_node_feature(f) for f in cv.curvature_features(m.points, m.tris, "curvature")]
    for got, want in zip(g.node_features, ref):
        # near-umbilic vertices: sqrt of a discriminant ~1e-14 is ~1e-7 (see _check); everything else agrees to 1e-11
        assert np.max(np.abs(got - want)) <= 1e-6 and np.median(np.abs(got - want)) <= 1e-10
    g.get_graph_spectrum()
    # adjacency in the feature-augmented space (graph.py:166-175) from the oracle with OUR features
    aug = port.feature_augmented_points(m.points, g.node_features)
    a_ref = port.adjacency(aug, m.tris)
    a = g.adjacency_matrix.tocsr()
    a.sort_indices()
    assert np.array_equal(a.indices, a_ref.indices) and np.array_equal(a.data, a_ref.data)
    ms = shipped_meshes["source_mesh"]
    f = pyfocusr.Focusr(m, ms, icp_register_first=False, rigid_reg_max_iterations=10, non_rigid_max_iterations=10,
                        n_coords_spectral_registration=800)      # everything else: the reference's defaults
    assert f.graph_target.n_extra_features == 2 and f.graph_source.n_extra_features == 2
    f.align_maps()
    assert f.corresponding_target_idx_for_each_source_pt.shape == (ms.points.shape[0],)
    assert np.isfinite(f.weighted_avg_transformed_points).all()
    g1 = pyfocusr.Graph(m, list_features_to_calc=["min_curvature", "max_curvature"])
    assert g1.n_extra_features == 2 and np.array_equal(g1.node_features[0], g.node_features[0])
