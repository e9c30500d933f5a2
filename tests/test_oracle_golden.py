"""The CPU oracle (oracle/port.py) against the golden fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py, run where /root/reference exists) and the reference's only published
known-answer vector (example notebook cell 2).  Runs without a GPU and without /root/reference."""
import hashlib

import numpy as np
import pytest
from scipy.spatial import KDTree

from oracle import port

NOTEBOOK_TARGET = [8.39246263e-04, 1.63007145e-03, 2.12549101e-03, 3.13941439e-03, 3.77495258e-03, 4.01682329e-03]
NOTEBOOK_SOURCE = [8.31236570e-04, 1.64152416e-03, 2.11362458e-03, 3.09029787e-03, 3.88535401e-03, 3.92405051e-03]
CASES = {"5k": ("target_mesh", "source_mesh"), "15k": ("target_mesh_15k", "source_mesh_15k")}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("tag", ["5k", "15k"])
def test_laplacian_assembly_matches_reference_hashes(tag, shipped_meshes, golden):
    for side, name in zip("ts", CASES[tag]):
        m = shipped_meshes[name]
        a = port.adjacency(m.points, m.tris)
        deg = port.row_sums_sequential(a)
        lap = port.laplacian(a, deg)
        assert sha(a.data) + sha(a.indices) + sha(a.indptr) == str(golden["%s_%s_A_sha" % (tag, side)])
        assert sha(lap.data) + sha(lap.indices) + sha(lap.indptr) == str(golden["%s_%s_L_sha" % (tag, side)])
        assert sha(deg) == str(golden["%s_%s_deg_sha" % (tag, side)])
        cnt = golden["%s_%s_counts" % (tag, side)]
        assert a.nnz == cnt[0] and lap.nnz == cnt[1] and int(np.sum(deg == 0)) == cnt[4]
    # SURVEY.md section 8 table
    assert golden["15k_t_counts"].tolist() == [89964, 104962, 12, 3, 0]
    assert golden["15k_s_counts"].tolist() == [89944, 104938, 8, 2, 2]


def test_eigenvalues_known_answer_and_retry(shipped_meshes, golden):
    """Notebook cell 2 (5k meshes) and the 15k source's retry to k=14 -> 11 eigenpairs."""
    for name, want in (("target_mesh", NOTEBOOK_TARGET), ("source_mesh", NOTEBOOK_SOURCE)):
        m = shipped_meshes[name]
        vals, vecs = port.recursive_eig(port.laplacian(port.adjacency(m.points, m.tris)), 7, 6, 1)
        assert np.allclose(np.sort(vals), want, rtol=5e-9, atol=0)
        assert vecs.shape == (5000, 6)
    m = shipped_meshes["source_mesh_15k"]
    vals, _ = port.recursive_eig(port.laplacian(port.adjacency(m.points, m.tris)), 7, 6, 1)
    gold = np.sort(golden["15k_s_eig_vals"])
    assert vals.size == 11 and np.allclose(np.sort(vals), gold, rtol=1e-8, atol=0)
    assert port.retry_k_final(7, 6, 1, 3) == 14 and port.retry_k_final(7, 6, 1, 1) == 7
    assert port.retry_k_final(7, 6, 1, 9) == 21 and port.retry_k_final(14, 13, 1, 1) == 14


@pytest.mark.parametrize("tag,n,n_spec", [("5k", 6, 3), ("15k", 6, 3), ("5k_n13", 13, 10)])
def test_eigsort_and_correspondences_bitwise(tag, n, n_spec, shipped_meshes, golden):
    names = CASES.get(tag, CASES["5k"])
    mt, ms = shipped_meshes[names[0]], shipped_meshes[names[1]]
    vt, vs = golden[tag + "_t_eig_vecs_raw_normed"].copy(), golden[tag + "_s_eig_vecs_raw_normed"].copy()
    srt = port.sort_eigenmaps(mt.points, ms.points, golden[tag + "_t_rand_idxs"], golden[tag + "_s_rand_idxs"],
                              golden[tag + "_t_eig_vals"], golden[tag + "_s_eig_vals"], vt, vs, n, True)
    for name in ("c_lambda", "c_hist", "c_hist_f", "c_spatial", "c_spatial_f", "Q", "target_matches", "source_matches"):
        assert np.array_equal(srt[name], golden["%s_%s" % (tag, name)]), name
    assert np.array_equal(srt["flipped_pairs"], golden[tag + "_flipped_pairs"])
    assert sha(vs) == str(golden[tag + "_sorted_vecs_s_sha"])
    w = port.spectral_weights(srt["Q"], golden[tag + "_s_eig_vals"], golden[tag + "_t_eig_vals"], n_spec)
    assert np.array_equal(w, golden[tag + "_spectral_weights"])
    if tag == "5k_n13":
        return
    cs = port.correspondence_stage(dict(A=port.adjacency(mt.points, mt.tris)), dict(A=port.adjacency(ms.points, ms.tris)),
                                   mt.points, ms.points, port.spectral_coords(vt, w, n_spec), port.spectral_coords(vs, w, n_spec))
    assert np.array_equal(cs["initial_idx"], golden[tag + "_initial_idx"])
    assert np.array_equal(cs["final_idx"], golden[tag + "_final_idx"])
    assert np.array_equal(cs["knn3_idx"], golden[tag + "_knn3_idx"])
    assert sha(cs["smoothed_target_coords"]) == str(golden[tag + "_smoothed_target_sha"])
    assert sha(cs["source_projected_on_target"]) == str(golden[tag + "_source_projected_sha"])
    assert sha(cs["weighted_avg_transformed_points"]) == str(golden[tag + "_weighted_avg_sha"])


def test_bruteforce_knn_contract():
    rng = np.random.RandomState(3)
    for dim in (1, 3, 10):
        refs, qs = rng.standard_normal((800, dim)), rng.standard_normal((300, dim))
        d, i = port.knn_bruteforce(refs, qs, 3)
        dk, ik = KDTree(refs).query(qs, k=3)
        assert np.array_equal(i, ik) and np.allclose(d, dk, rtol=1e-15, atol=0)
    refs = np.concatenate([rng.standard_normal((50, 3))] * 2)
    _, i = port.knn_bruteforce(refs, refs[:50], 2)
    assert np.array_equal(i[:, 0], np.arange(50)) and np.array_equal(i[:, 1], np.arange(50, 100))  # lower index first


def test_large_golden_present():
    import os

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "large_eigs.npz"))
    v = z["nu316_k11"]
    assert v.size == 10
    # exact multiplets of the sphere: 3 + 5 (+2 of the 7-fold l=3 level)
    assert np.ptp(v[:3]) / v[0] < 1e-7 and np.ptp(v[3:8]) / v[3] < 1e-7 and np.ptp(v[8:]) / v[8] < 1e-7
    assert z["nu100_seed5_k65"].size == 64
