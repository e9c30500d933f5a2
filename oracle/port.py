"""CPU oracle for the PyFOCUSR spectral-correspondence hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a vectorised numpy/scipy *restatement* of the reference's algorithm (the
reference is pure Python over scipy/numpy, so the oracle is Python too).  It exists so that
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs can check and time the reference's CPU path on the GPU box, where
``/root/reference`` does not exist.  Nothing under ``pyfocusr_b200/`` imports it: the product
path is the CUDA library and fails loudly without it.

Pinning (see ``oracle/make_golden.py``): every function below is asserted against the
*unmodified* reference (imported from ``/root/reference`` with the stub modules in
``oracle/stubs``) on the four shipped meshes -- adjacency / degree / Laplacian / smoothing
/ KNN / final positions bit-identical, eigenvalues to <= 1e-9 relative (ARPACK start vector
is random in the reference) -- and against the reference's only published known-answer
vector, the 12 eigenvalues printed in ``examples/Example_registering_two_bone_meshes.ipynb``
cell 2.  The arithmetic that lives in third-party code (scipy 1.18.1:
ARPACK ``eigs``, ``cKDTree``, ``wasserstein_distance``, ``linear_sum_assignment``) is called
exactly as the reference calls it, at the cited call sites.

Every function cites the reference ``file:line`` it follows (paths relative to
``/root/reference/pyfocusr``).
"""
from __future__ import annotations

import numpy as np
from scipy import sparse
from scipy.optimize import linear_sum_assignment
from scipy.sparse.linalg import eigs
from scipy.spatial import KDTree
from scipy.stats import wasserstein_distance

MIN_EIG_VAL = 1e-10  # graph.py:369


# ---------------------------------------------------------------------------------------
# A0  Graph.__init__ geometry                                            graph.py:58-67
# ---------------------------------------------------------------------------------------
def normed_points(points):
    """graph.py:63-67  ``(p - min) / mean(ptp)``."""
    ptp = np.ptp(points, axis=0)
    return (points - np.min(points, axis=0)) / np.mean(ptp)


def rand_idxs(n_points, n_rand_samples):
    """graph.py:274-290  (consumes the global numpy RNG exactly like the reference)."""
    if n_rand_samples > n_points:
        return np.arange(n_points)
    return np.random.choice(n_points, size=n_rand_samples, replace=False)


# ---------------------------------------------------------------------------------------
# A1  weighted adjacency                                               graph.py:148-178
# ---------------------------------------------------------------------------------------
def directed_edges(tris):
    """Edge order of graph.py:156-161 (cell-major; VTK polygon edges 01,12,20)."""
    t = np.asarray(tris, dtype=np.int64)
    return np.stack([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]], axis=1).reshape(-1, 2)


def adjacency(points, tris):
    """graph.py:163-178: ``A[p1,p2] = 1/sqrt(sum((x1-x2)^2))``, assignment (duplicates collapse),
    one direction per cell edge.  ``np.sum`` over 3 squares is ``(d0^2+d1^2)+d2^2`` (no FMA).
    Returns canonical CSR (sorted columns), float64 data / int32 indices."""
    p = np.asarray(points, dtype=np.float64)
    e = directed_edges(tris)
    n = p.shape[0]
    d = p[e[:, 0]] - p[e[:, 1]]
    # points may carry range-scaled node features behind xyz (graph.py:166-175); np.sum over < 8
    # squares is the plain left-to-right sum
    s2 = d[:, 0] * d[:, 0]
    for c in range(1, p.shape[1]):
        s2 = s2 + d[:, c] * d[:, c]
    with np.errstate(divide="ignore"):
        w = 1.0 / np.sqrt(s2)
    key = e[:, 0] * n + e[:, 1]
    # last write wins; duplicates of one (p1,p2) carry the identical value, so any pick is exact
    uk, ui = np.unique(key, return_index=True)
    a = sparse.csr_matrix((w[ui], (uk // n, uk % n)), shape=(n, n))
    a.sort_indices()
    return a


# ---------------------------------------------------------------------------------------
# A2  degree                                                           graph.py:216-219
# ---------------------------------------------------------------------------------------
def row_sums_sequential(a):
    """``A.sum(axis=1)`` of the reference == left-to-right sum over ascending columns from 0.0
    (pinned bitwise in make_golden.py; ``np.add.reduceat`` is NOT bit-equal)."""
    a = a.tocsr()
    n = a.shape[0]
    cnt = np.diff(a.indptr)
    acc = np.zeros(n)
    for p in range(int(cnt.max()) if n else 0):
        rows = np.nonzero(cnt > p)[0]
        acc[rows] = acc[rows] + a.data[a.indptr[rows] + p]
    return acc


def degree_inv(deg):
    """graph.py:219  ``(d + 1e-8) ** -1`` (bitwise equal to ``1.0/(d+1e-8)``)."""
    return (deg + 1e-8) ** -1


# ---------------------------------------------------------------------------------------
# A3/A4  G = D^-1 (default branch), L = G @ (D - A)             graph.py:213-214,221-226
# ---------------------------------------------------------------------------------------
def laplacian(a, deg=None):
    """graph.py:225-226 with G = degree_matrix_inv.  ``L_ij = dinv_i * (-w_ij)``,
    ``L_ii = dinv_i * d_i``; explicit zeros dropped (zero-degree rows are empty).
    Returned with sorted indices (scipy emits each row descending; compare after sorting)."""
    a = a.tocsr()
    n = a.shape[0]
    if deg is None:
        deg = row_sums_sequential(a)
    dinv = degree_inv(deg)
    rows = np.repeat(np.arange(n), np.diff(a.indptr))
    data = np.concatenate([dinv[rows] * (-a.data), dinv * deg])
    r = np.concatenate([rows, np.arange(n)])
    c = np.concatenate([a.indices, np.arange(n)])
    keep = data != 0
    lap = sparse.csr_matrix((data[keep], (r[keep], c[keep])), shape=(n, n))
    lap.sort_indices()
    return lap


# ---------------------------------------------------------------------------------------
# B1  recursive_eig                                                    graph.py:357-389
# ---------------------------------------------------------------------------------------
def recursive_eig(matrix, k, n_k_needed, k_buffer=1, sigma=1e-10, which="LM", v0=None):
    """graph.py:357-389 verbatim in behaviour (prints dropped).  ``v0`` (not in the reference,
    default None = reference behaviour) lets golden generation be reproducible."""
    eig_vals, eig_vecs = eigs(matrix, k=k, sigma=sigma, which=which, ncv=4 * k, v0=v0)
    n_good = sum(eig_vals > MIN_EIG_VAL)
    if n_good < n_k_needed:
        k += k_buffer + n_k_needed
        eig_vals, eig_vecs = recursive_eig(matrix, k, n_k_needed, k_buffer, sigma, which, v0)
    keep = np.where(eig_vals > MIN_EIG_VAL)[0]
    return np.real(eig_vals[keep]), np.real(eig_vecs[:, keep])


def retry_k_final(k, n_k_needed, k_buffer, z):
    """The k the recursion of graph.py:374-379 ends at when the spectrum holds ``z``
    eigenvalues <= 1e-10."""
    while k - min(z, k) < n_k_needed:
        k += k_buffer + n_k_needed
    return k


# ---------------------------------------------------------------------------------------
# B2  eigenvector normalisation                                        graph.py:254-257
# ---------------------------------------------------------------------------------------
def normalize_eigvecs(v):
    return (v - np.min(v, axis=0)) / np.ptp(v, axis=0) - 0.5


# ---------------------------------------------------------------------------------------
# sub-sampling helpers                                                 graph.py:263-272
# ---------------------------------------------------------------------------------------
def eig_val_gap(eig_vals):
    return np.mean(np.diff(eig_vals))


def rand_normalized_points(points, idxs):
    s = points[idxs, :]
    return (s - np.min(s, axis=0)) / np.ptp(s, axis=0)


# ---------------------------------------------------------------------------------------
# F1  mean_filter_graph                                                graph.py:320-354
# ---------------------------------------------------------------------------------------
def mean_filter(a, values, iterations=300):
    """graph.py:349-354.  ``average_mat = diag(1/(1+rowsum)) @ (A + I)``; scipy stores each row
    of the product in *descending* column order and ``@`` accumulates ``y += a*x`` in stored
    order without FMA, which the CUDA smoothing kernel reproduces bit-for-bit."""
    a = a.tocsr()
    d_inv = sparse.diags(1.0 / (1 + row_sums_sequential(a)))
    out = values
    average_mat = d_inv @ (a + sparse.eye(a.shape[0]))
    for _ in range(iterations):
        out = average_mat @ out
    return out


# ---------------------------------------------------------------------------------------
# C1-C5  eigsort                                                       eigsort.py:9-249
# ---------------------------------------------------------------------------------------
def c_lambda(eig_vals_t, eig_vals_s, n):
    """eigsort.py:142-160 (gap averages over *all* returned eigenvalues of each graph)."""
    gap = (eig_val_gap(eig_vals_t) + eig_val_gap(eig_vals_s)) / 2
    lt = np.asarray(eig_vals_t)[:n, None]
    ls = np.asarray(eig_vals_s)[None, :n]
    return np.exp((lt - ls) ** 2 / (2 * gap**2))


def c_hist(rt, rs, n):
    """eigsort.py:162-189: W1 distance between log-shifted sampled eigenvector columns."""
    eps = np.finfo(float).eps
    c = np.zeros((n, n))
    cf = np.zeros((n, n))
    for i in range(n):
        u = np.log(rt[:, i] + 0.5 + eps)
        for j in range(n):
            c[i, j] = wasserstein_distance(u, np.log(rs[:, j] + 0.5 + eps))
            cf[i, j] = wasserstein_distance(u, np.log(-rs[:, j] + 0.5 + eps))
    return c, cf


def c_spatial(pts_t01, pts_s01, rt, rs, n):
    """eigsort.py:191-233: NN of sampled target xyz in sampled source xyz, then per column pair
    ``sqrt(sum((+-s_j[idx] - t_i)^2)) / n_samples``."""
    _, idx = KDTree(pts_s01).query(pts_t01)
    c = np.zeros((n, n))
    cf = np.zeros((n, n))
    m = rt.shape[0]
    for i in range(n):
        for j in range(n):
            c[i, j] = np.sqrt(np.sum((rs[idx, j] - rt[:, i]) ** 2)) / m
            cf[i, j] = np.sqrt(np.sum((-rs[idx, j] - rt[:, i]) ** 2)) / m
    return c, cf, idx


def eigen_sort_decide(cl, ch, chf, cs, csf, target_as_reference=True):
    """eigsort.py:66-105.  Returns ``(Q_pairs, target_matches, source_matches, flipped_pairs)``."""
    c = cs * cl * ch
    c_f = csf * cl * chf
    q = np.min((c, c_f), axis=0)
    s = c > c_f
    tf, sf = np.where(s)
    if target_as_reference:
        tm, sm = linear_sum_assignment(q)
    else:
        sm, tm = linear_sum_assignment(q.T)
    q_pairs = q[tm, sm]
    flipped = [p2 for p1 in zip(tf, sf) for p2 in zip(tm, sm) if p2 == p1]
    return q_pairs, tm, sm, flipped


def eigen_sort_apply(vecs_t, vecs_s, tm, sm, flipped, target_as_reference=True):
    """eigsort.py:108-122 (in place; fancy-index RHS is copied first, so a true permutation)."""
    for m0, m1 in flipped:
        if target_as_reference:
            vecs_s[:, m1] = vecs_s[:, m1] * -1
        else:
            vecs_t[:, m0] = vecs_t[:, m0] * -1
    if target_as_reference:
        vecs_s[:, tm] = vecs_s[:, sm]
    else:
        vecs_t[:, sm] = vecs_t[:, tm]


def sort_eigenmaps(pts_t, pts_s, idx_t, idx_s, vals_t, vals_s, vecs_t, vecs_s, n,
                   target_as_reference=True):
    """eigsort.py:34-41 + 235-249.  Mutates ``vecs_*`` like the reference; returns a dict."""
    rt, rs = vecs_t[idx_t, :], vecs_s[idx_s, :]
    pt, ps = rand_normalized_points(pts_t, idx_t), rand_normalized_points(pts_s, idx_s)
    cl = c_lambda(vals_t, vals_s, n)
    ch, chf = c_hist(rt, rs, n)
    cs, csf, nn = c_spatial(pt, ps, rt, rs, n)
    q, tm, sm, flipped = eigen_sort_decide(cl, ch, chf, cs, csf, target_as_reference)
    eigen_sort_apply(vecs_t, vecs_s, tm, sm, flipped, target_as_reference)
    return dict(Q=q, c_lambda=cl, c_hist=ch, c_hist_f=chf, c_spatial=cs, c_spatial_f=csf,
                target_matches=np.asarray(tm), source_matches=np.asarray(sm),
                flipped_pairs=np.asarray(flipped, dtype=np.int64).reshape(-1, 2), nn_idx=nn)


# ---------------------------------------------------------------------------------------
# D1  spectral weights / coords                                        focusr.py:459-508
# ---------------------------------------------------------------------------------------
def spectral_weights(q, vals_s, vals_t, ns):
    w = q[:ns] * np.max((vals_s[:ns], vals_t[:ns]), axis=0)
    sigma = np.mean(w)
    return np.exp(-(w**2) / (2 * sigma**2))


def spectral_coords(vecs, weights, ns, weighted=True):
    return vecs[:, :ns] * weights[None, :] if weighted else vecs[:, :ns]


# ---------------------------------------------------------------------------------------
# E1/E2  KD correspondence                                             focusr.py:351-353
# ---------------------------------------------------------------------------------------
def kd_correspondence(target_pts, source_pts, k=1):
    return KDTree(target_pts).query(source_pts, k=k)


def knn_bruteforce(refs, queries, k=1, chunk=512):
    """Exact fp64 k-NN with the product's tie rule (lower index wins): direct differences
    ``sum((a_i - b_i)^2)`` accumulated left to right, never ``|a|^2+|b|^2-2ab``
    (SURVEY.md section 7.3-7).  Used to state the tie-break contract cKDTree leaves open."""
    refs = np.asarray(refs, dtype=np.float64)
    queries = np.asarray(queries, dtype=np.float64)
    nq = queries.shape[0]
    idx = np.zeros((nq, k), dtype=np.int64)
    dist = np.zeros((nq, k))
    for s in range(0, nq, chunk):
        q = queries[s : s + chunk]
        d2 = np.zeros((q.shape[0], refs.shape[0]))
        for c in range(refs.shape[1]):
            diff = q[:, c, None] - refs[None, :, c]
            d2 = d2 + diff * diff
        order = np.argsort(d2, axis=1, kind="stable")[:, :k]
        idx[s : s + chunk] = order
        dist[s : s + chunk] = np.sqrt(np.take_along_axis(d2, order, axis=1))
    return dist, idx


# ---------------------------------------------------------------------------------------
# E3/E4  final positions                                               focusr.py:401-431
# ---------------------------------------------------------------------------------------
def weighted_final_positions(smoothed_target, source_projected, target_points, n_closest=3):
    """focusr.py:401-426, vectorised over source points (the reference loops in Python).
    Per point: any zero distance -> copy that target point; else inverse-distance weights,
    ``sum_k(p_k*w_k) / (w_0+w_1+w_2)`` with both sums left to right."""
    dist, idx = KDTree(smoothed_target).query(source_projected, k=n_closest)
    out = np.zeros((source_projected.shape[0], 3))
    zero = dist == 0
    has0 = zero.any(axis=1)
    first0 = np.argmax(zero, axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = 1 / dist
        num = np.zeros_like(out)
        den = np.zeros(dist.shape[0])
        for j in range(n_closest):
            num = num + target_points[idx[:, j], :] * w[:, j, None]
            den = den + w[:, j]
        out[:] = num / den[:, None]
    rows = np.nonzero(has0)[0]
    out[rows] = target_points[idx[rows, first0[rows]]]
    return out, dist, idx


# ---------------------------------------------------------------------------------------
# whole spectral stage (Focusr.__init__ + align_maps minus ICP / CPD)
# ---------------------------------------------------------------------------------------
def graph_spectrum(points, tris, n_spectral_features, norm_eig_vecs=True, v0=None):
    """graph.py:228-257."""
    a = adjacency(points, tris)
    deg = row_sums_sequential(a)
    lap = laplacian(a, deg)
    vals, vecs = recursive_eig(lap, k=n_spectral_features + 1, n_k_needed=n_spectral_features,
                               k_buffer=1, v0=v0)
    if norm_eig_vecs:
        vecs = normalize_eigvecs(vecs)
    return dict(A=a, deg=deg, L=lap, eig_vals=vals, eig_vecs=vecs)


def correspondence_stage(gt, gs, pts_t, pts_s, target_coords, source_coords,
                         graph_smoothing_iterations=300, projection_smooth_iterations=40):
    """focusr.py:545-562 with ``kd`` correspondences (defaults)."""
    _, idx0 = kd_correspondence(target_coords, source_coords)
    smoothed_t = mean_filter(gt["A"], pts_t, graph_smoothing_iterations)
    src_proj = mean_filter(gs["A"], smoothed_t[idx0, :], projection_smooth_iterations)
    _, idx1 = kd_correspondence(smoothed_t, src_proj)
    wavg, d3, i3 = weighted_final_positions(smoothed_t, src_proj, pts_t)
    nearest = pts_t[idx1, :]
    return dict(initial_idx=idx0, smoothed_target_coords=smoothed_t,
                source_projected_on_target=src_proj, final_idx=idx1,
                weighted_avg_transformed_points=wavg, nearest_neighbor_transformed_points=nearest,
                knn3_dist=d3, knn3_idx=i3)


def spectral_stage(pts_t, tris_t, pts_s, tris_s, n_spectral_features=3, n_extra_spectral=3,
                   n_coords_spectral_ordering=5000, idx_t=None, idx_s=None,
                   graph_smoothing_iterations=300, projection_smooth_iterations=40,
                   target_as_reference=True, weighted=True):
    """The north-star path for one pair, CPD = identity."""
    n = n_spectral_features + n_extra_spectral
    gt = graph_spectrum(pts_t, tris_t, n)
    gs = graph_spectrum(pts_s, tris_s, n)
    if idx_t is None:
        idx_t = rand_idxs(pts_t.shape[0], n_coords_spectral_ordering)
    if idx_s is None:
        idx_s = rand_idxs(pts_s.shape[0], n_coords_spectral_ordering)
    srt = sort_eigenmaps(pts_t, pts_s, idx_t, idx_s, gt["eig_vals"], gs["eig_vals"],
                         gt["eig_vecs"], gs["eig_vecs"], n, target_as_reference)
    w = spectral_weights(srt["Q"], gs["eig_vals"], gt["eig_vals"], n_spectral_features)
    tc = spectral_coords(gt["eig_vecs"], w, n_spectral_features, weighted)
    sc = spectral_coords(gs["eig_vecs"], w, n_spectral_features, weighted)
    out = correspondence_stage(gt, gs, pts_t, pts_s, tc, sc, graph_smoothing_iterations,
                               projection_smooth_iterations)
    out.update(graph_target=gt, graph_source=gs, eigsort=srt, spectral_weights=w,
               target_spectral_coords=tc, source_spectral_coords=sc)
    return out


# ---------------------------------------------------------------------------------------------------
# optional node features (mesh scalars)                  graph.py:88-142,166-175; focusr.py:218-269
# ---------------------------------------------------------------------------------------------------
def normalized_node_feature(values, norm_using_std=True, cap_std=3, norm_range_0_to_1=True):
    """graph.py:121-142."""
    f = np.array(values, dtype=np.float64)
    if norm_using_std:
        f = (f - np.mean(f)) / np.std(f)
        if cap_std is not False:
            f[f > cap_std] = cap_std
            f[f < -cap_std] = -cap_std
    if norm_range_0_to_1:
        f = (f - np.min(f)) / np.ptp(f)
    return f


def feature_augmented_points(points, node_features):
    """graph.py:113-119,166-175: xyz followed by the features scaled to the mean xyz range."""
    scale = np.mean(np.ptp(points, axis=0))
    return np.concatenate([points] + [(f * scale)[:, None] for f in node_features], axis=1)


def features_as_coords(a, node_features, coords, iterations=40):
    """focusr.py:228-262 for one graph: smooth, rescale to [0, 1], scale to the spectral range."""
    out = np.zeros((coords.shape[0], len(node_features)))
    for k, f in enumerate(node_features):
        g = mean_filter(a, f, iterations)
        g = g - np.min(g)
        g = g / np.max(g)
        out[:, k] = np.ptp(coords) * g
    return np.concatenate((coords, out), axis=1)
