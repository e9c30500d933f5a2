"""Generate ``tests/golden/*.npz`` from the UNMODIFIED reference and pin ``oracle/port.py``.

Runs only where ``/root/reference`` exists (the build container).  It
  1. imports the reference package with the stub modules in ``oracle/stubs`` (vtk, itkwidgets,
     cycpd, matplotlib are not installable here; none is touched on the hot path);
  2. runs the reference's own classes (``Graph``, ``eigsort``, ``Focusr`` methods) on the four
     shipped meshes with seeded ``np.random`` (so ``rand_idxs`` are reproducible);
  3. asserts the restatement in ``oracle/port.py`` reproduces every stage (bit-identical where
     the arithmetic is deterministic, <=1e-9 relative for ARPACK output whose start vector the
     reference draws from OS entropy);
  4. asserts the reference's only published known-answer vector (example notebook cell 2);
  5. writes the fixtures the GPU box needs (it has no ``/root/reference``).

Usage:  python oracle/make_golden.py
"""
from __future__ import annotations

import contextlib
import hashlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path[:0] = [os.path.join(HERE, "stubs"), REF, ROOT]

from oracle import port  # noqa: E402
from pyfocusr_b200.mesh import read_vtk_mesh  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# examples/Example_registering_two_bone_meshes.ipynb, cell 2 stored output
NOTEBOOK_TARGET = [8.39246263e-04, 1.63007145e-03, 2.12549101e-03, 3.13941439e-03,
                   3.77495258e-03, 4.01682329e-03]
NOTEBOOK_SOURCE = [8.31236570e-04, 1.64152416e-03, 2.11362458e-03, 3.09029787e-03,
                   3.88535401e-03, 3.92405051e-03]


SCALAR = "thickness_change_(mm)"  # point scalar carried by every shipped mesh (data/*.vtk:34998)


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def run_reference_pair(ref, mesh_t, mesh_s, n_spec, n_extra, seed_t, seed_s, n_samples=5000):
    """The reference's own code path: Focusr.__init__ (minus ICP) + align_maps (minus CPD)."""
    n = n_spec + n_extra
    graphs = []
    for mesh, seed in ((mesh_t, seed_t), (mesh_s, seed_s)):
        np.random.seed(seed)
        g = quiet(ref.Graph, mesh, n_spectral_features=n, n_rand_samples=n_samples,
                  list_features_to_calc=[], feature_weights=np.eye(2))
        quiet(g.get_graph_spectrum)
        graphs.append(g)
    gt, gs = graphs
    raw = dict(vecs_t=gt.eig_vecs.copy(), vecs_s=gs.eig_vecs.copy())
    f = object.__new__(ref.Focusr)  # ctor needs VTK ICP; set what align_maps reads
    f.graph_target, f.graph_source = gt, gs
    f.n_spectral_features, f.n_extra_spectral = n_spec, n_extra
    f.n_total_spectral_features = n
    f.target_eigenmap_as_reference = True
    f.get_weighted_spectral_coords = True
    f.initial_correspondence_type = f.final_correspondence_type = "kd"
    f.graph_smoothing_iterations, f.projection_smooth_iterations = 300, 40
    sorter = ref.eigsort(graph_target=gt, graph_source=gs, n_features=n, target_as_reference=True)
    f.Q = quiet(sorter.sort_eigenmaps)
    f.calc_spectral_coords()
    f.get_initial_correspondences()
    idx0 = np.asarray(f.corresponding_target_idx_for_each_source_pt).copy()
    f.get_smoothed_correspondences()
    f.get_weighted_final_node_locations()
    f.get_nearest_neighbour_final_node_locations()
    return gt, gs, f, sorter, raw, idx0


def check_graph(ref_g, mesh, tag):
    """Pin A / degree / L / mean_filter of the port against the reference Graph (bitwise)."""
    a = port.adjacency(mesh.points, mesh.tris)
    ar = ref_g.adjacency_matrix.tocsr()
    ar.sort_indices()
    assert np.array_equal(ar.indptr, a.indptr) and np.array_equal(ar.indices, a.indices), tag
    assert np.array_equal(ar.data, a.data), tag + " A values"
    deg = port.row_sums_sequential(a)
    assert np.array_equal(ref_g.degree_matrix.diagonal(), deg), tag + " degree"
    assert np.array_equal(ref_g.degree_matrix_inv.diagonal(), port.degree_inv(deg)), tag
    lap = port.laplacian(a, deg)
    lr = ref_g.laplacian_matrix.copy()
    lr.sort_indices()
    assert np.array_equal(lr.indptr, lap.indptr) and np.array_equal(lr.indices, lap.indices), tag
    assert np.array_equal(lr.data, lap.data), tag + " L values"
    assert np.array_equal(ref_g.normed_points, port.normed_points(mesh.points)), tag
    # mean_filter_graph on xyz, 7 iterations, bitwise
    assert np.array_equal(ref_g.mean_filter_graph(mesh.points, 7),
                          port.mean_filter(a, mesh.points, 7)), tag + " mean_filter"
    e = port.directed_edges(mesh.tris)
    key = e[:, 0] * mesh.points.shape[0] + e[:, 1]
    asym = abs(a - a.T)
    asym.eliminate_zeros()
    return dict(a=a, deg=deg, lap=lap, n_oneway=int(asym.nnz),
                n_dup=int(key.size - np.unique(key).size), n_isolated=int(np.sum(deg == 0)))


def feature_case(ref, mt, ms):
    """a mesh scalar as extra node feature, inside the adjacency weights
    (graph.py:166-175) and appended to the spectral coordinates (focusr.py:218-269)."""
    tag, n_spec, n = "5k_feat", 3, 6
    graphs = []
    for mesh, seed in ((mt, 0), (ms, 1)):
        np.random.seed(seed)
        g = quiet(ref.Graph, mesh, n_spectral_features=n, n_rand_samples=5000, list_features_to_calc=[],
                  list_features_to_get_from_mesh=[SCALAR], feature_weights=np.eye(1), include_features_in_adj_matrix=True)
        quiet(g.get_graph_spectrum)
        graphs.append(g)
    gt, gs = graphs
    out = {}
    infos = []
    for side, g, mesh in (("t", gt, mt), ("s", gs, ms)):
        feat = port.normalized_node_feature(mesh.point_scalars[SCALAR])
        assert np.array_equal(feat, g.node_features[0]), side
        aug = port.feature_augmented_points(mesh.points, [feat])
        a = port.adjacency(aug, mesh.tris)
        ar = g.adjacency_matrix.tocsr()
        ar.sort_indices()
        assert np.array_equal(ar.indices, a.indices) and np.array_equal(ar.data, a.data), side + " feature adjacency"
        lap = port.laplacian(a)
        lr = g.laplacian_matrix.copy()
        lr.sort_indices()
        assert np.array_equal(lr.data, lap.data), side
        out[f"{tag}_{side}_A_sha"] = np.array(sha(a.data) + sha(a.indices) + sha(a.indptr))
        out[f"{tag}_{side}_eig_vals"] = g.eig_vals.copy()
        out[f"{tag}_{side}_eig_vecs_raw_normed"] = g.eig_vecs.copy()
        out[f"{tag}_{side}_rand_idxs"] = np.asarray(g.rand_idxs)
        infos.append((a, feat))
    raw_t, raw_s = gt.eig_vecs.copy(), gs.eig_vecs.copy()
    f = object.__new__(ref.Focusr)
    f.graph_target, f.graph_source = gt, gs
    f.n_spectral_features, f.n_extra_spectral, f.n_total_spectral_features = n_spec, 3, n
    f.target_eigenmap_as_reference = f.get_weighted_spectral_coords = True
    f.initial_correspondence_type = f.final_correspondence_type = "kd"
    f.graph_smoothing_iterations, f.projection_smooth_iterations, f.feature_smoothing_iterations = 300, 40, 40
    sorter = ref.eigsort(graph_target=gt, graph_source=gs, n_features=n, target_as_reference=True)
    f.Q = quiet(sorter.sort_eigenmaps)
    f.calc_spectral_coords()
    quiet(f.append_features_to_spectral_coords)
    f.get_initial_correspondences()
    idx0 = np.asarray(f.corresponding_target_idx_for_each_source_pt).copy()
    f.get_smoothed_correspondences()
    # port on the same pre-sort vectors
    vt, vs = raw_t.copy(), raw_s.copy()
    srt = port.sort_eigenmaps(mt.points, ms.points, gt.rand_idxs, gs.rand_idxs, gt.eig_vals, gs.eig_vals, vt, vs, n, True)
    assert np.array_equal(f.Q, srt["Q"])
    w = port.spectral_weights(srt["Q"], gs.eig_vals, gt.eig_vals, n_spec)
    tc = port.features_as_coords(infos[0][0], [infos[0][1]], port.spectral_coords(vt, w, n_spec), 40)
    sc = port.features_as_coords(infos[1][0], [infos[1][1]], port.spectral_coords(vs, w, n_spec), 40)
    assert np.array_equal(tc, f.target_spectral_coords) and np.array_equal(sc, f.source_spectral_coords)
    cs = port.correspondence_stage(dict(A=infos[0][0]), dict(A=infos[1][0]), mt.points, ms.points, tc, sc)
    assert np.array_equal(cs["initial_idx"], idx0)
    assert np.array_equal(cs["final_idx"], f.corresponding_target_idx_for_each_source_pt)
    for name in ("Q", "target_matches", "source_matches", "flipped_pairs"):
        out[f"{tag}_{name}"] = srt[name]
    out[f"{tag}_spectral_weights"] = w
    out[f"{tag}_target_coords_sha"] = np.array(sha(tc))
    out[f"{tag}_source_coords_sha"] = np.array(sha(sc))
    out[f"{tag}_initial_idx"] = idx0.astype(np.int32)
    out[f"{tag}_final_idx"] = cs["final_idx"].astype(np.int32)
    print(tag, "feature adjacency / features-as-coords / correspondences: port == reference (bitwise); eig",
          np.sort(gt.eig_vals)[:3])
    return out


def main():
    if not os.path.isdir(REF):
        raise SystemExit("needs /root/reference (build container only)")
    import pyfocusr as ref  # the unmodified reference

    os.makedirs(GOLD, exist_ok=True)
    names = ["target_mesh", "source_mesh", "target_mesh_15k", "source_mesh_15k"]
    meshes = {n: read_vtk_mesh(os.path.join(REF, "data", n + ".vtk")) for n in names}
    np.savez_compressed(
        os.path.join(GOLD, "meshes.npz"),
        **{n + "_points": m.points for n, m in meshes.items()},
        **{n + "_tris": m.tris for n, m in meshes.items()},
        **{n + "_scalar": m.point_scalars[SCALAR] for n, m in meshes.items()},
    )

    out = {}
    cases = [("5k", "target_mesh", "source_mesh", 3, 3), ("15k", "target_mesh_15k", "source_mesh_15k", 3, 3),
             ("5k_n13", "target_mesh", "source_mesh", 10, 3)]
    for tag, tn, sn, n_spec, n_extra in cases:
        mt, ms = meshes[tn], meshes[sn]
        n = n_spec + n_extra
        gt, gs, f, sorter, raw, idx0 = run_reference_pair(ref, mt, ms, n_spec, n_extra, 0, 1)
        info_t = check_graph(gt, mt, tag + " target")
        info_s = check_graph(gs, ms, tag + " source")
        for side, g, info in (("t", gt, info_t), ("s", gs, info_s)):
            # --- eigensolve: port vs reference (random ARPACK v0 -> tolerance, not bitwise)
            vals, vecs = port.recursive_eig(info["lap"], n + 1, n, 1)
            o = np.argsort(vals)
            ro = np.argsort(g.eig_vals)
            assert vals.shape == g.eig_vals.shape, (tag, side, vals.shape, g.eig_vals.shape)
            rel = np.max(np.abs(vals[o] - g.eig_vals[ro]) / g.eig_vals[ro])
            assert rel < 1e-8, (tag, side, rel)
            out[f"{tag}_{side}_eig_vals"] = g.eig_vals.copy()
            out[f"{tag}_{side}_eig_vecs_raw_normed"] = raw["vecs_" + side]  # before eigsort
            out[f"{tag}_{side}_rand_idxs"] = np.asarray(g.rand_idxs)
            out[f"{tag}_{side}_A_sha"] = np.array(sha(info["a"].data) + sha(info["a"].indices) + sha(info["a"].indptr))
            out[f"{tag}_{side}_L_sha"] = np.array(sha(info["lap"].data) + sha(info["lap"].indices) + sha(info["lap"].indptr))
            out[f"{tag}_{side}_deg_sha"] = np.array(sha(info["deg"]))
            out[f"{tag}_{side}_counts"] = np.array([info["a"].nnz, info["lap"].nnz, info["n_oneway"],
                                                    info["n_dup"], info["n_isolated"]])
            print(tag, side, "N", g.n_points, "nnzA", info["a"].nnz, "nnzL", info["lap"].nnz,
                  "oneway", info["n_oneway"], "dup", info["n_dup"], "isolated", info["n_isolated"],
                  "m", g.eig_vals.size, "eig rel diff port/ref %.2e" % rel)
        if tag == "5k":
            assert np.allclose(np.sort(gt.eig_vals), NOTEBOOK_TARGET, rtol=5e-9, atol=0)
            assert np.allclose(np.sort(gs.eig_vals), NOTEBOOK_SOURCE, rtol=5e-9, atol=0)
            print("notebook known-answer eigenvalues reproduced")

        # --- eigsort: port on the reference's own (pre-sort) eigenvectors, bitwise
        vt, vs = raw["vecs_t"].copy(), raw["vecs_s"].copy()
        srt = port.sort_eigenmaps(mt.points, ms.points, gt.rand_idxs, gs.rand_idxs, gt.eig_vals,
                                  gs.eig_vals, vt, vs, n, True)
        for name in ("c_lambda", "c_hist", "c_hist_f", "c_spatial", "c_spatial_f"):
            assert np.array_equal(getattr(sorter, name), srt[name]), (tag, name)
        assert np.array_equal(f.Q, srt["Q"]), tag
        assert np.array_equal(gs.eig_vecs, vs) and np.array_equal(gt.eig_vecs, vt), tag
        w = port.spectral_weights(srt["Q"], gs.eig_vals, gt.eig_vals, n_spec)
        assert np.array_equal(w, f.spectral_weights), tag
        tc = port.spectral_coords(vt, w, n_spec)
        sc = port.spectral_coords(vs, w, n_spec)
        assert np.array_equal(tc, f.target_spectral_coords), tag
        assert np.array_equal(sc, f.source_spectral_coords), tag
        # --- correspondences (CPD = identity), bitwise
        cs = port.correspondence_stage(dict(A=info_t["a"]), dict(A=info_s["a"]), mt.points, ms.points, tc, sc)
        assert np.array_equal(cs["initial_idx"], idx0), tag
        assert np.array_equal(cs["smoothed_target_coords"], f.smoothed_target_coords), tag
        assert np.array_equal(cs["source_projected_on_target"], f.source_projected_on_target), tag
        assert np.array_equal(cs["final_idx"], f.corresponding_target_idx_for_each_source_pt), tag
        assert np.array_equal(cs["weighted_avg_transformed_points"], f.weighted_avg_transformed_points), tag
        assert np.array_equal(cs["nearest_neighbor_transformed_points"], f.nearest_neighbor_transformed_points), tag
        # brute-force KNN (product tie rule) == cKDTree on the reference's features
        for refs, qs, want in ((tc, sc, idx0),
                               (cs["smoothed_target_coords"], cs["source_projected_on_target"], cs["final_idx"])):
            _, bi = port.knn_bruteforce(refs, qs, 1)
            assert np.array_equal(bi[:, 0], want), tag + " brute-force vs cKDTree"
        d3, i3 = port.knn_bruteforce(cs["smoothed_target_coords"], cs["source_projected_on_target"], 3)
        assert np.array_equal(i3, cs["knn3_idx"]) and np.array_equal(d3, cs["knn3_dist"]), tag + " k=3"
        print(tag, "eigsort / coords / smoothing / KNN / positions: port == reference (bitwise);",
              "matches", srt["target_matches"], srt["source_matches"], "flipped", srt["flipped_pairs"].tolist())

        for name in ("c_lambda", "c_hist", "c_hist_f", "c_spatial", "c_spatial_f", "Q",
                     "target_matches", "source_matches", "flipped_pairs"):
            out[f"{tag}_{name}"] = srt[name]
        out[f"{tag}_spectral_weights"] = w
        out[f"{tag}_initial_idx"] = idx0.astype(np.int32)
        out[f"{tag}_final_idx"] = cs["final_idx"].astype(np.int32)
        out[f"{tag}_knn3_idx"] = cs["knn3_idx"].astype(np.int32)
        out[f"{tag}_smoothed_target_sha"] = np.array(sha(cs["smoothed_target_coords"]))
        out[f"{tag}_source_projected_sha"] = np.array(sha(cs["source_projected_on_target"]))
        out[f"{tag}_weighted_avg_sha"] = np.array(sha(cs["weighted_avg_transformed_points"]))
        out[f"{tag}_sorted_vecs_s_sha"] = np.array(sha(vs))

    out.update(feature_case(ref, meshes["target_mesh"], meshes["source_mesh"]))
    np.savez_compressed(os.path.join(GOLD, "reference_outputs.npz"), **out)
    for fn in sorted(os.listdir(GOLD)):
        print(fn, os.path.getsize(os.path.join(GOLD, fn)) // 1024, "KiB")


if __name__ == "__main__":
    main()
