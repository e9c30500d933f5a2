"""Shared checker of the stages after the eigensolver against the CPU oracle (test infrastructure).

eigsort's costs depend on the arbitrary sign of the target eigenvectors, so the oracle is fed OUR
pre-sort eigenvectors.  The chain of custody is then:
  (a) eigsort: Q within 1e-9 of the oracle's, the permuted / flipped eigenvectors bitwise equal;
  (b) spectral coordinates within 1e-9 (relative to the largest coordinate) of the oracle's;
  (c) with the oracle's own coordinates, every initial correspondence that differs from ours is a distance tie
      within 1e-9 -- rounding in the cost sums can only flip (near-)ties, nothing else;
  (d) fed OUR coordinates, the oracle's correspondence stage (KNN, 300 + 40 smoothing passes, KNN, k = 3 weighted
      positions) reproduces ours EXACTLY: indices equal, coordinates bitwise equal.  No tolerance on indices anywhere.
"""
import numpy as np


def check_pair_against_oracle(port, out, p, n_pairs, pts_t, tris_t, pts_s, tris_s, n_features, ns, smooth_t, smooth_s):
    """``out``: SpectralBatch.run(..., keep_presort=True) result; pair ``p`` of ``n_pairs``.  Returns the number of
    initial correspondences that were (near-)tie flips in step (c)."""
    off = out["graph"].mesh_off_host
    nf = out["eigs_info"]["n_found"]
    vals = out["eig_vals"].cpu().numpy()
    o_t, o_s = int(off[p]), int(off[n_pairs + p])
    n_t, n_s = pts_t.shape[0], pts_s.shape[0]
    nt_total = int(off[n_pairs])
    pre = out["eig_vecs_presort"]
    vt = pre[o_t:o_t + n_t, :nf[p]].cpu().numpy().copy()
    vs = pre[o_s:o_s + n_s, :nf[n_pairs + p]].cpu().numpy().copy()
    srt = port.sort_eigenmaps(pts_t, pts_s, out["idx_t"][p], out["idx_s"][p], vals[p, :nf[p]],
                              vals[n_pairs + p, :nf[n_pairs + p]], vt, vs, n_features, True)
    # (a)
    assert np.max(np.abs(out["Q"][p].cpu().numpy() - srt["Q"]) / srt["Q"]) <= 1e-9
    assert int(out["eigsort_status"].item()) == 0
    post = out["eig_vecs"][o_s:o_s + n_s, :n_features].cpu().numpy()
    assert np.array_equal(post, vs[:, :n_features])
    # (b)
    w = port.spectral_weights(srt["Q"], vals[n_pairs + p], vals[p], ns)
    ct, cs = port.spectral_coords(vt, w, ns), port.spectral_coords(vs, w, ns)
    coords = out["coords"].cpu().numpy()
    our_t, our_s = coords[o_t:o_t + n_t], coords[o_s:o_s + n_s]
    scale = max(np.max(np.abs(ct)), np.max(np.abs(cs)))
    assert np.max(np.abs(our_t - ct)) <= 1e-9 * scale and np.max(np.abs(our_s - cs)) <= 1e-9 * scale
    # (c)
    q0 = o_s - nt_total
    ours0 = out["initial_idx"][q0:q0 + n_s].cpu().numpy()
    _, ref0 = port.kd_correspondence(ct, cs)
    diff = np.nonzero(ours0 != ref0)[0]
    if diff.size:
        d_ours = np.linalg.norm(ct[ours0[diff]] - cs[diff], axis=1)
        d_ref = np.linalg.norm(ct[ref0[diff]] - cs[diff], axis=1)
        assert np.max(np.abs(d_ours - d_ref)) <= 1e-9 * scale, "an initial correspondence differs by more than a tie"
    # (d)
    ref = port.correspondence_stage(dict(A=port.adjacency(pts_t, tris_t)), dict(A=port.adjacency(pts_s, tris_s)),
                                    pts_t, pts_s, our_t, our_s, smooth_t, smooth_s)
    assert np.array_equal(ours0, ref["initial_idx"])
    assert np.array_equal(out["smoothed_target_coords"][o_t:o_t + n_t].cpu().numpy(), ref["smoothed_target_coords"])
    assert np.array_equal(out["source_projected_on_target"][q0:q0 + n_s].cpu().numpy(), ref["source_projected_on_target"])
    assert np.array_equal(out["final_idx"][q0:q0 + n_s].cpu().numpy(), ref["final_idx"])
    assert np.array_equal(out["knn3_idx"][q0:q0 + n_s].cpu().numpy(), ref["knn3_idx"])
    assert np.array_equal(out["weighted_avg_transformed_points"][q0:q0 + n_s].cpu().numpy(),
                          ref["weighted_avg_transformed_points"])
    return int(diff.size)
